"""spirit_b200: B200-native (sm_100a CUDA, fp64) implementation of Spirit's Heisenberg gradient / LLG / GNEB hot path
behind Spirit's own C API. The product is spirit_b200/libSpirit.so (built by spirit_b200/build.py); this package
is the thin host-side driver (ctypes) used by the tests and bench.py. There is no CPU fallback."""
from . import capi, session  # noqa: F401
from .capi import load_product  # noqa: F401
from .session import Session  # noqa: F401
