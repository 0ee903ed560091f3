"""openkite_b200 -- B200-native batched kite-dynamics engine (drop-in for openKITE's kite_model hot path).

Product layout:
  csrc/        hand-written sm_100a CUDA kernels + the extern "C" layer (include/kite_b200.h)
  engine.py    ctypes binding used by tests / bench (torch tensors as device memory)
  build.py     in-tree nvcc build of libkite_b200.so
  sharding.py  block partition of the global index range + gather of per-unit results (torch.distributed)
  collocation.py  host-side Chebyshev operators (constant matrices handed to kite_colloc_eval)
The C++ host mirror of the reference API (KiteDynamics, ODESolver, KiteEKF, Chebyshev) lives in include/openkite/.
"""
from .engine import (Engine, KiteError, KiteParams, NmpcCost, load_properties, load_library, KITE, KITE_ID, RIGID_BODY,  # noqa: F401
                     U_CONST, U_PER_STEP, U_SHARED, U_SYNTH, LIB_PATH)
