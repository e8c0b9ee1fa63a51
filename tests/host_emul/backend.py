"""TEST-ONLY backend for multiagent_planning_b200.sharded.ShardedDMPC: runs the device algorithm's
host build (tests/host_emul/emul.cpp) on CPU torch tensors so that the sharding / all-gather host
logic can be exercised with world_size 2 over gloo where no GPU exists."""
import numpy as np
import torch

from . import emul


class EmulBackend:
    def __init__(self, N, params, pmin, pmax, pf, n0, n1, rows):
        self.N, self.K, self.n0, self.n1 = N, int(params.K), n0, n1
        self.P = emul.params_from(params)
        self.pmin, self.pmax, self.pf = np.asarray(pmin, float), np.asarray(pmax, float), np.asarray(pf, float)
        self.l = [torch.zeros(rows, self.K, 3, dtype=torch.float64) for _ in range(2)]
        self.st = [[torch.zeros(N, 3, dtype=torch.float64) for _ in range(3)] for _ in range(2)]
        self.status = torch.zeros(N, dtype=torch.int32)
        self.goal_out = torch.zeros(2, dtype=torch.float64)

    def init(self, po):
        # initDMPC.m: p(:,i) = po + t_i (pf - po)/10
        t = np.arange(self.K) * self.P.h
        p = po[:, None, :] + (t[None, :, None] * (self.pf - po)[:, None, :]) / self.P.init_div  # (3,K,N)
        self.l[0][: self.N] = torch.from_numpy(np.ascontiguousarray(p.transpose(2, 1, 0)))
        self.st[0][0][:] = torch.from_numpy(np.ascontiguousarray(po.T))
        self.st[0][1].zero_()
        self.st[0][2].zero_()

    def _np(self, t):  # (N,3) -> (3,N)
        return np.asfortranarray(t.numpy().T)

    def step_local(self, cur):
        nx = cur ^ 1
        l_prev = np.asfortranarray(self.l[cur][: self.N].numpy().transpose(2, 1, 0))
        o = emul.step(self.P, self._np(self.st[cur][0]), self._np(self.st[cur][1]), self._np(self.st[cur][2]),
                      self.pf, l_prev, self.pmin, self.pmax, n0=self.n0, n1=self.n1)
        sl = slice(self.n0, self.n1)
        self.l[nx][sl] = torch.from_numpy(np.ascontiguousarray(o["l_new"].transpose(2, 1, 0)))[sl]
        for i, k in enumerate(("p1", "v1", "a1")):
            self.st[nx][i][sl] = torch.from_numpy(np.ascontiguousarray(o[k].T))[sl]
        self.status[sl] = torch.from_numpy(o["status"])[sl]

    def goal(self, nx):
        p = self.l[nx][: self.N, 0, :].numpy().T
        md = float(np.sqrt(((p - self.pf) ** 2).sum(0)).max())
        self.goal_out[0] = md
        self.goal_out[1] = 1.0 if md < self.P.goal_tol else 0.0

    def close(self):
        pass
