"""One-process-per-GPU DMPC: agents sharded in contiguous blocks, ONE all-gather of predicted
horizons per MPC step (torch.distributed, NCCL over NVLink / NVSwitch).

Reference: the C++ clusters of dmpc/cpp/dmpc.cpp:1600-1625 (contiguous agent blocks per thread),
the per-step exchange `prev_obs = obs` (dmpc.cpp:1681) / `l = new_l` (test/failure_rate.m:124) and
the goal test on the exchanged data (ReachedGoal.m, reached_goalv2 dmpc.cpp:1868-1882).

Every rank holds the full replicated horizon buffer l (N x K x 3 fp64, = MATLAB's 3 x K x N) and
the state of its own block only.  A step is: scan + QP kernels on the local block writing the
block's new horizons straight into the send slice of the next buffer, then one in-place
all_gather_into_tensor.  The goal test runs redundantly on every rank from the gathered buffer
(first column of every horizon = new position), so it needs no second collective; per-agent
status words are exchanged only when the caller asks (every `check_every` steps).

torch is plumbing here (device tensors over the library's own buffers, streams, the process
group); the compute is libdmpc_b200.so.  The backend is injectable so that the host logic can be
exercised with world_size 2 over gloo on a CPU-only machine by the test-suite.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import dmpc


def partition(N: int, world: int):
    """Contiguous blocks of ceil(N/world) agents (dmpc.cpp:1600-1625).  Trailing blocks may be short or EMPTY
    (N = 17 on 8 ranks: blocks of 3, ranks 6 and 7 own nothing): an empty rank still holds a replica of the
    horizon buffer and takes part in the exchange, it just solves no agent (the library accepts n0 == n1)."""
    if N < 1 or world < 1:
        raise ValueError("partition: need N >= 1 and world >= 1")
    blk = -(-N // world)
    return blk, [(min(r * blk, N), min((r + 1) * blk, N)) for r in range(world)]


class _DevArray:
    """Zero-copy torch view of library-owned device memory (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False),
                                             version=2)


class CudaBackend:
    """Local block on one B200 through the C-ABI (dmpcb200_step_dev / dmpcb200_goal_dev)."""

    def __init__(self, N, params, pmin, pmax, pf, n0, n1, rows, device):
        self.N, self.K, self.n0, self.n1 = N, int(params.K), n0, n1
        self.device = torch.device("cuda", device)
        self.solver = dmpc.Solver(N, params, n0=n0, n1=n1, device=device, pmin=pmin, pmax=pmax, pf=pf)
        s = self.solver
        npad = -(-N // 32) * 32
        if rows > npad:
            raise dmpc.DmpcError("world size must divide 32 (1, 2, 4, 8)")
        view = lambda which, shape, ts="<f8": torch.as_tensor(_DevArray(s.device_ptr(which), shape, ts),
                                                              device=self.device)
        self.l = [view(0, (npad, self.K, 3)), view(1, (npad, self.K, 3))]
        self.st = [[view(w, (N, 3)) for w in (2, 3, 4)], [view(w, (N, 3)) for w in (8, 9, 10)]]
        self.status = view(6, (N,), "<i4")
        self.goal_out = view(7, (2,))

    def init(self, po):
        self.solver.init_horizons(po)  # fills l[0], st[0] for all agents

    def step_local(self, cur):
        nx = cur ^ 1
        s = self.solver
        st, sn = self.st[cur], self.st[nx]
        s.step_dev(st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), self.l[cur].data_ptr(),
                   self.l[nx].data_ptr(), sn[0].data_ptr(), sn[1].data_ptr(), sn[2].data_ptr(),
                   self.status.data_ptr(), stream=torch.cuda.current_stream(self.device).cuda_stream)

    def goal(self, nx):
        self.solver.goal_dev(self.l[nx].data_ptr(), 3 * self.K, self.goal_out.data_ptr(),
                             stream=torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        self.solver.close()


class ShardedDMPC:
    def __init__(self, N, params, pmin, pmax, po, pf, rank=None, world=None, group=None, backend_factory=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.N, self.K, self.P = int(N), int(params.K), params
        # every rank computes the same partition from (N, world): a bad configuration raises on ALL ranks
        # before any backend exists, never on one rank while the others wait in the all-gather
        self.blk, parts = partition(self.N, self.world)
        if self.blk * self.world > -(-self.N // 32) * 32:
            raise dmpc.DmpcError(f"{self.world} blocks of {self.blk} agents exceed the padded horizon buffer "
                                 f"({-(-self.N // 32) * 32} rows): use a world size of 1, 2, 4 or 8")
        self.n0, self.n1 = parts[self.rank]
        self.rows = self.blk * self.world  # agent rows in the gathered buffer (>= N, tail rows inert)
        self.pf = np.asarray(pf, float).reshape(3, N, order="F")
        if backend_factory is None:
            dev = torch.cuda.current_device()
            backend_factory = lambda **kw: CudaBackend(device=dev, **kw)
        self.be = backend_factory(N=self.N, params=params, pmin=pmin, pmax=pmax, pf=self.pf, n0=self.n0,
                                  n1=self.n1, rows=self.rows)
        self.be.init(np.asarray(po, float).reshape(3, N, order="F"))
        self.cur = 0
        self.steps = 0
        self.n_allgather = 0
        self._inplace = dist.get_backend(group) == "nccl" if dist.is_initialized() else False

    # one MPC step: local solves + the single all-gather
    def step(self):
        be, nx = self.be, self.cur ^ 1
        be.step_local(self.cur)
        if self.world > 1:
            out = be.l[nx][: self.rows]
            send = out[self.rank * self.blk:(self.rank + 1) * self.blk]
            if self._inplace:
                dist.all_gather_into_tensor(out.view(-1), send.reshape(-1), group=self.group)
            elif out.is_cuda:
                # test configuration (gloo with device buffers: two ranks on one GPU): stage through the host
                got = torch.empty(out.shape, dtype=out.dtype)
                dist.all_gather_into_tensor(got.view(-1), send.reshape(-1).cpu(), group=self.group)
                out.copy_(got)
            else:
                dist.all_gather_into_tensor(out.view(-1), send.clone().reshape(-1), group=self.group)
            self.n_allgather += 1
        self.cur = nx
        self.steps += 1

    def reached_goal(self):
        """ReachedGoal.m on the gathered buffer; one host read."""
        self.be.goal(self.cur)
        g = self.be.goal_out.cpu()
        return bool(g[1] != 0), float(g[0])

    def gather_status(self):
        """status words of all agents (one small all-gather; not on the per-step path)."""
        loc = torch.zeros(self.blk, dtype=torch.int32, device=self.be.status.device)
        loc[: self.n1 - self.n0] = self.be.status[self.n0:self.n1]
        if self.world == 1:
            return loc[: self.N].cpu().numpy()
        out = torch.zeros(self.rows, dtype=torch.int32, device=loc.device)
        dist.all_gather_into_tensor(out, loc, group=self.group)
        return out[: self.N].cpu().numpy()

    def run(self, max_steps, check_every=8, stop_on_fail=False):
        """`while ~reached_goal && k < max_K`: the goal / failure look happens every check_every steps
        (the state keeps stepping in between; a reached swarm sits at its goal)."""
        reached, fail = False, -1
        for k in range(max_steps):
            self.step()
            if (k + 1) % check_every == 0 or k + 1 == max_steps:
                reached, _ = self.reached_goal()
                if stop_on_fail:
                    st = self.gather_status()
                    bad = np.nonzero(((st & dmpc.ST_SOLVED) == 0) | ((st & dmpc.ST_OUTBOUND) != 0))[0]
                    if bad.size:
                        fail = int(bad[0])
                        break
                if reached:
                    break
        return dict(steps=self.steps, reached=reached, first_fail_agent=fail)

    def horizons(self):
        """current l as numpy (3, K, N)"""
        t = self.be.l[self.cur][: self.N].cpu().numpy()
        return np.asfortranarray(t.transpose(2, 1, 0))

    def local_state(self):
        st = self.be.st[self.cur]
        return tuple(np.asfortranarray(x[self.n0:self.n1].cpu().numpy().T) for x in st)

    def capture_graph(self):
        """CUDA graph of two steps (even -> odd -> even) including the NCCL all-gathers."""
        assert self.cur == 0, "capture at an even step"
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step()
            self.step()
        self.steps -= 2  # capture does not execute
        self.n_allgather -= 2 if self.world > 1 else 0
        self.cur = 0
        return g

    def close(self):
        self.be.close()
