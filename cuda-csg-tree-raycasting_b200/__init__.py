"""cuda-csg-tree-raycasting_b200 — B200-native CSG ray caster (hot path only).

csrc/      CUDA kernel (sm_100a), host scene code and the C ABI  -> libcsg_b200.so
host/      C++ mirror of the reference's Raycaster / CSGTree / Camera interface over the C ABI
binding.py ctypes binding used by tests and bench.py
"""
from .binding import *  # noqa: F401,F403
from .binding import lib, LIB_PATH, EXPORTS  # noqa: F401
