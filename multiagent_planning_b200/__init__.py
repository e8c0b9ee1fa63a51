"""multiagent_planning_b200 -- B200-native (sm_100a CUDA) DMPC per-agent QP hot path of
carlosluis/multiagent_planning, behind the reference's own function surface.

    from multiagent_planning_b200 import dmpc
    s = dmpc.Solver(N, dmpc.default_params(), pmin=pmin, pmax=pmax, pf=pf)
    s.init_horizons(po); s.run(150)

`dmpc` holds the host mirror (reference-named functions + Solver), `_lib` the ctypes binding of
libdmpc_b200.so (include/dmpc_b200.h), `sharded` the one-process-per-GPU runner (torch.distributed).
"""
from . import _lib  # noqa: F401
from . import dmpc  # noqa: F401

__all__ = ["dmpc", "_lib"]
