"""B200-native hot path of jfour1e/SBI-for-Diffusion-Models.

Host layer: Python/PyTorch with the reference's function signatures and (z, x) layout.
Compute: hand-written sm_100a CUDA in ``_lib/libddm_b200.so`` behind the C ABI declared in
``include/ddm_b200.h``.  No CPU fallback.
"""
from . import constants, run_config  # noqa: F401

__all__ = ["constants", "run_config"]
