"""Host driver of the device pulse-side generator (``ddm_pulses_pcg64``).

The reference draws pulse trains trial by trial in Python from a NumPy ``Generator``
(models/rt_choice_model.py:62-91 -> models/choice_model.py:43-60).  Here the same NumPy
PCG64 stream is continued on the GPU: trial ``i`` of a call owns draws ``[i(P+1), (i+1)(P+1))``
after the generator's current state, so the matrix is identical to NumPy's and the host
``Generator`` is advanced by the draws consumed -- later NumPy draws from the same ``rng``
continue exactly where the reference's would.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native
from .simulator import compute_device

_M64 = (1 << 64) - 1


def pcg64_state(rng: np.random.Generator) -> Tuple[int, int]:
    bg = rng.bit_generator
    st = bg.state
    if st.get("bit_generator") != "PCG64":
        raise TypeError(
            f"pulse generation continues a NumPy PCG64 stream on the GPU; got bit generator "
            f"{st.get('bit_generator')!r}. Use np.random.default_rng(seed).")
    return int(st["state"]["state"]), int(st["state"]["inc"])


def success_threshold(p_success: float) -> int:
    """ceil(clip(p,0,1) * 2**53):  k * 2**-53 < p  <=>  k < threshold  (choice_model.py:56,58)."""
    p = min(max(float(p_success), 0.0), 1.0)
    return int(math.ceil(p * 9007199254740992.0))


def pulses_from_state(state: int, inc: int, first_trial: int, n_trials: int, n_pulses: int, p_success: float,
                      *, device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rows [first_trial, first_trial+n) of the stream that starts at (state, inc).

    ``out`` may be a column slice of a wider matrix (e.g. ``z[:, 5:]``): its row stride is
    passed to the kernel, so z is assembled in place without a concatenation."""
    if n_trials < 0:
        raise ValueError("n_trials must be >= 0")
    if n_pulses < 0:
        raise ValueError("n_pulses must be >= 0")
    L = _native.lib()
    dev = compute_device(device if out is None else out.device)
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((n_trials, n_pulses), dtype=torch.float32, device=dev)
        if out.shape != (n_trials, n_pulses) or out.dtype != torch.float32 or (n_pulses > 1 and out.stride(1) != 1):
            raise ValueError("out must be (n_trials, n_pulses) fp32 with unit column stride")
        ld = out.stride(0) if n_trials > 1 else max(n_pulses, 1)
        rc = L.ddm_pulses_pcg64(ctypes.c_uint64(state >> 64), ctypes.c_uint64(state & _M64),
                                ctypes.c_uint64(inc >> 64), ctypes.c_uint64(inc & _M64),
                                ctypes.c_uint64(first_trial), n_trials, n_pulses,
                                ctypes.c_uint64(success_threshold(p_success)),
                                out.data_ptr() if out.numel() else None, ld,
                                torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "ddm_pulses_pcg64")
    return out


def generate_pulse_matrix_device(rng: np.random.Generator, n_trials: int, n_pulses: int, *, p_success: float,
                                 device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device twin of ``generate_pulse_matrix_numpy``: same values, tensor stays on the GPU,
    and ``rng`` ends in the state NumPy would have left it in."""
    state, inc = pcg64_state(rng)
    s = pulses_from_state(state, inc, 0, n_trials, n_pulses, p_success, device=device, out=out)
    if n_trials > 0 and n_pulses > 0:
        rng.bit_generator.advance(n_trials * (n_pulses + 1))
    return s
