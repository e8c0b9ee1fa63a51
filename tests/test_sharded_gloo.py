"""Agent sharding + the single per-step all-gather, world_size 2 over gloo on CPU.

The compute backend is injected: the test-only host build of the device algorithm stands in for
the CUDA kernels, so what is verified here is the HOST logic of
multiagent_planning_b200/sharded.py (partition, in-place gather layout, goal test on the gathered
buffer, status exchange): the 2-rank run must equal the 1-rank run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scenario(N):
    from multiagent_planning_b200 import scenarios
    pmin, pmax = scenarios.density_arena(N, density=1.5)
    po, pf = scenarios.random_test(N, pmin, pmax, 0.35, 2.0, seed=21)
    return pmin, pmax, po, pf


def _worker(rank, world, port, N, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multiagent_planning_b200 import _lib, sharded
        from tests.host_emul.backend import EmulBackend
        pmin, pmax, po, pf = _scenario(N)
        P = _lib.Params()
        _lib.lib().dmpcb200_default_params(P, 0)
        sh = sharded.ShardedDMPC(N, P, pmin, pmax, po, pf, backend_factory=EmulBackend)
        for _ in range(steps):
            sh.step()
        reached, md = sh.reached_goal()
        st = sh.gather_status()
        if rank == 0:
            q.put(dict(l=sh.horizons(), md=md, status=st, n_allgather=sh.n_allgather, steps=sh.steps,
                       part=(sh.n0, sh.n1)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [1, 31, 48])   # N = 1: the second rank owns an EMPTY block
def test_two_ranks_equal_one_rank(emul, N):
    steps = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["n_allgather"] == steps and res["steps"] == steps     # exactly one all-gather per step
    assert res["part"] == (0, -(-N // 2))

    # single-rank reference through the same code
    from multiagent_planning_b200 import _lib, sharded
    from tests.host_emul.backend import EmulBackend
    pmin, pmax, po, pf = _scenario(N)
    P = _lib.Params()
    _lib.lib().dmpcb200_default_params(P, 0)
    one = sharded.ShardedDMPC(N, P, pmin, pmax, po, pf, rank=0, world=1, backend_factory=EmulBackend)
    for _ in range(steps):
        one.step()
    assert np.array_equal(one.horizons(), res["l"])
    assert one.reached_goal()[1] == res["md"]
    assert np.array_equal(one.gather_status(), res["status"])


def test_partition():
    from multiagent_planning_b200.sharded import partition
    assert partition(500, 8) == (63, [(0, 63), (63, 126), (126, 189), (189, 252), (252, 315), (315, 378),
                                      (378, 441), (441, 500)])
    assert partition(100, 1) == (100, [(0, 100)])
    blk, parts = partition(9, 4)
    assert blk == 3 and parts[-1] == (9, 9)
    blk, parts = partition(17, 8)          # two empty trailing ranks: valid, they only take part in the exchange
    assert blk == 3 and parts[5] == (15, 17) and parts[6] == parts[7] == (17, 17)
    with pytest.raises(ValueError):
        partition(0, 2)


def test_bad_world_size_raises_identically_on_every_rank():
    """a partition that does not fit the padded buffer must raise before any backend is created (so that no
    rank is left waiting in the all-gather): same exception for every rank index"""
    from multiagent_planning_b200 import _lib, dmpc, sharded
    pmin, pmax, po, pf = _scenario(64)
    P = _lib.Params()
    _lib.lib().dmpcb200_default_params(P, 0)
    made = []
    for rank in range(3):   # 3 blocks of 22 = 66 rows > 64 padded rows
        with pytest.raises(dmpc.DmpcError, match="exceed the padded"):
            sharded.ShardedDMPC(64, P, pmin, pmax, po, pf, rank=rank, world=3,
                                backend_factory=lambda **kw: made.append(kw))
    assert not made
