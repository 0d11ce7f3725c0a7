"""chipmunk2d_b200 -- B200-native cpSpaceStep hot path behind the Chipmunk2D C API.

Layout:
  csrc/    CUDA kernels + the C ABI (include/cpb200.h) -> lib/libcpb200.so
  host/    C99 host layer implementing the public Chipmunk2D API (include/chipmunk/chipmunk.h)
           on top of the C ABI -> lib/libchipmunk_b200.so
  scenes/  scene blob format + public-API scene loader (compiled against both libraries)
  engine.py / api.py / scenes.py  ctypes plumbing for tests and bench
"""
from .engine import World, Scene, EngineError, load_engine  # noqa: F401
