"""poismf_b200 — B200-native (sm_100a) implementation of poismf's alternating-sweep hot path.

Product surface:
  * C ABI            include/poismf_b200.h  (libpoismf_b200.so, built in-tree)
  * host drop-in     poismf_b200/host/poismf_host.c  (run_poismf / predict_multiple / topN
                     with the reference's own prototypes)
  * Python mirror    poismf_b200.c_funs  (same names as the reference's Cython wrapper)
  * device handle    poismf_b200.device.DeviceFit
  * sharding layer   poismf_b200.sharding

Nothing here falls back to the CPU: a missing library or a missing GPU raises.
"""
from . import _lib  # noqa: F401
from ._lib import FLAG_NO_CACHED, FLAG_NO_LOCKSTEP, FLAG_STRICT, SIDE_CSC, SIDE_CSR, make_params  # noqa: F401

__all__ = ["c_funs", "device", "synth", "make_params", "FLAG_STRICT", "FLAG_NO_CACHED", "FLAG_NO_LOCKSTEP", "SIDE_CSR", "SIDE_CSC"]
