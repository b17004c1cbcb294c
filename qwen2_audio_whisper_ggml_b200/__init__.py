"""B200-native (sm_100a) mel + Qwen2-Audio/Whisper-large-v3 encoder front-end.

Python is only the thin host mirror used by tests and bench.py; the product is the C-ABI library
``libq2w_b200.so`` (``include/q2w_b200.h``) and the reference-compatible C API (``include/qwen2-whisper.h``).
There is no CPU fallback: importing works anywhere (so the CPU test tier can check symbols), but every
compute call raises if the CUDA library or an sm_100 device is missing.
"""
from .lib import load_library, Q2WError  # noqa: F401
from .api import Context, default_context_params  # noqa: F401

__all__ = ["load_library", "Q2WError", "Context", "default_context_params"]
