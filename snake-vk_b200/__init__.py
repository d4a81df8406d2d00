"""snake-vk_b200 — B200-native path-tracing pass behind Snake-VK's `RT` interface.

The product is `libsnkpt.so` (CUDA kernels for sm_100a + the C ABI of include/snkpt.h);
this package is the thin Python host mirror used by the tests and bench.py. Import
fails loudly when the CUDA library has not been built: there is no CPU fallback.
The directory name has a hyphen, so import it with
`importlib.import_module("snake-vk_b200")`.
"""
from . import scenes, sharding, frontend  # noqa: F401
from .api import RT, PathTracer, SnkError, load_library, lib_path  # noqa: F401
