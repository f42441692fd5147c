"""CPU oracle for the pulse-DDM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Nothing under
``sbi_for_diffusion_models_b200/`` imports it; the product path fails loudly when the
CUDA library is missing instead of falling back to anything here.

Two independent restatements of the reference simulator live here:

* ``sim_scalar_c``      -- one trial at a time in plain C (``ddm_oracle.c``), fast enough
                          to check 1e5-trial GPU batches in seconds;
* ``sim_lockstep_torch`` -- batched torch ops in lock-step over the whole batch, i.e. the
                          reference's own execution strategy (19 tensor-op dispatches and
                          two host syncs per Euler step).  This is what ``bench.py --impl
                          reference`` times as the reference's CPU implementation.

Both follow ``/root/reference/src/sbi_for_diffusion_models/models/rt_choice_model.py:112-221``
(citations on each function).  Parity status: PINNED -- ``tests/golden/*.npz`` hold outputs
of the imported reference (``tests/golden/make_golden.py``) and
``tests/test_oracle_golden.py`` checks both restatements against them bit for bit.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Callable, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libddm_oracle.so")

# reference constants.py:3-5
T_MAX = 8.0
DT_CHOICE = 5e-4
PULSE_INTERVAL = 0.1


def build(force: bool = False) -> str:
    """Compile ddm_oracle.c with gcc (no FMA contraction)."""
    src = os.path.join(_HERE, "ddm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "_build/libddm_oracle.so"], check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, f32, u64, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p
        L.ddm_oracle_sim_f32.argtypes = [vp, i64, vp, i64, i64, i64, i64, i64, f32, f32, f32, f32,
                                         vp, ctypes.c_int, vp, vp]
        L.ddm_oracle_sim_f32.restype = ctypes.c_int
        L.ddm_oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.ddm_oracle_philox4x32_10.restype = None
        L.ddm_oracle_philox_words.argtypes = [u64, u64, i64, i64, vp]
        L.ddm_oracle_philox_words.restype = None
        L.ddm_oracle_pulses_pcg64.argtypes = [u64, u64, u64, u64, u64, i64, i64, u64, vp, i64]
        L.ddm_oracle_pulses_pcg64.restype = None
        L.ddm_oracle_pcg64_advance.argtypes = [vp, vp, u64, u64, u64]
        L.ddm_oracle_pcg64_advance.restype = None
        _lib = L
    return _lib


# --------------------------------------------------------------------------------------
# schedule and fp32 scalar rounding
# --------------------------------------------------------------------------------------

def schedule(dt: float = DT_CHOICE, t_max: float = T_MAX,
             pulse_interval: float = PULSE_INTERVAL) -> Tuple[int, int, int]:
    """(n_max, steps_per_pulse, P).  rt_choice_model.py:52-53, :59."""
    n_max = int(np.floor(float(t_max) / float(dt)))
    spp = max(int(np.round(float(pulse_interval) / float(dt))), 1)
    return n_max, spp, (n_max + spp - 1) // spp


def fp32_scalars(dt: float = DT_CHOICE, t_max: float = T_MAX, mu_sensory: float = 1.0):
    """Python floats as float32 tensors see them (rt_choice_model.py:135,137,141,146-147,186)."""
    f = lambda v: float(np.float32(v))
    return dict(dt=f(dt), t_max=f(t_max), t_nd_hi=f(float(t_max) - 1e-6),
                noise_scale=f(float(mu_sensory) * float(np.sqrt(dt))))


# --------------------------------------------------------------------------------------
# restatement 1: scalar C
# --------------------------------------------------------------------------------------

def sim_scalar_c(theta, pulses, noise, *, dt: float = DT_CHOICE, t_max: float = T_MAX,
                 pulse_interval: float = PULSE_INTERVAL, mu_sensory: float = 1.0,
                 log_rt: bool = False):
    """theta (N,5) f32, pulses (N,P) or (1,P) f32, noise (n_max,N) f32 -> x (N,2) f32, hit_step (N,) i64."""
    theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float32))
    pulses = np.ascontiguousarray(np.asarray(pulses, dtype=np.float32))
    noise = np.ascontiguousarray(np.asarray(noise, dtype=np.float32))
    if theta.ndim == 1:
        theta = theta[None, :]
    if pulses.ndim == 1:
        pulses = pulses[None, :]
    N = theta.shape[0]
    n_max, spp, _ = schedule(dt, t_max, pulse_interval)
    assert noise.shape == (n_max, N), (noise.shape, (n_max, N))
    ld_p = 0 if (pulses.shape[0] == 1 and N > 1) else pulses.shape[1]
    sc = fp32_scalars(dt, t_max, mu_sensory)
    x = np.empty((N, 2), dtype=np.float32)
    steps = np.empty((N,), dtype=np.int64)
    rc = lib().ddm_oracle_sim_f32(theta.ctypes.data, theta.shape[1], pulses.ctypes.data, ld_p,
                                  N, pulses.shape[1], n_max, spp, sc["dt"], sc["t_max"],
                                  sc["t_nd_hi"], sc["noise_scale"], noise.ctypes.data,
                                  int(bool(log_rt)), x.ctypes.data, steps.ctypes.data)
    if rc != 0:
        raise ValueError("ddm_oracle_sim_f32 rejected its arguments")
    return x, steps


# --------------------------------------------------------------------------------------
# restatement 2: batched lock-step torch (the reference's execution strategy)
# --------------------------------------------------------------------------------------

def sim_lockstep_torch(theta: torch.Tensor, pulses: torch.Tensor,
                       draw: Optional[Callable[[int, int], torch.Tensor]] = None, *,
                       dt: float = DT_CHOICE, t_max: float = T_MAX,
                       pulse_interval: float = PULSE_INTERVAL, mu_sensory: float = 1.0):
    """Whole batch advanced one Euler step at a time until no trial is active.

    ``draw(k, N)`` supplies the k-th (N,) noise vector; default ``torch.randn`` (the
    reference's global-generator behaviour, rt_choice_model.py:186).
    Returns (x (N,2) f32 raw rt / choice, hit_step (N,) i64, steps_executed).
    """
    th = theta.to(torch.float32).reshape(-1, 5)
    N = th.shape[0]
    n_max, spp, P = schedule(dt, t_max, pulse_interval)
    s = pulses.to(torch.float32)
    if s.ndim == 1:
        s = s[None, :]
    if s.shape[0] == 1 and N > 1:
        s = s.expand(N, -1)
    if s.shape[0] != N or s.shape[1] < P:
        raise ValueError("pulse matrix does not match the batch / schedule")
    if draw is None:
        draw = lambda k, n: torch.randn((n,), dtype=torch.float32)

    start_frac = th[:, 0].clamp(0.0, 1.0)             # :131
    leak = th[:, 1]                                   # :132
    gain = th[:, 2].abs()                             # :133
    bound = th[:, 3].abs().clamp_min(1e-6)            # :134
    t_nd = th[:, 4].clamp(0.0, float(t_max) - 1e-6)   # :135
    window = torch.floor((float(t_max) - t_nd) / dt).to(torch.int64).clamp(0, n_max)  # :141
    acc = start_frac * bound                          # :144
    scale = float(mu_sensory) * np.sqrt(dt)           # :146-147

    done = torch.zeros(N, dtype=torch.bool)
    which = torch.zeros(N, dtype=torch.int64)
    when = torch.zeros(N, dtype=torch.int64)
    executed = 0
    for k in range(n_max):
        live = (~done) & (k < window)                 # :182
        if not bool(live.any()):
            break
        executed += 1
        acc = acc + (-leak * acc) * dt + draw(k, N) * scale   # :186-187
        if k % spp == 0:                              # :190-192
            acc = acc + gain * s[:, k // spp] * live.to(torch.float32)
        top = live & (acc >= bound)                   # :195
        bottom = live & (acc <= 0.0)                  # :196
        crossed = top | bottom
        if bool(crossed.any()):                       # :199-204
            when = torch.where(crossed, torch.full_like(when, k + 1), when)
            which = torch.where(top, torch.ones_like(which), which)
            which = torch.where(bottom, torch.zeros_like(which), which)
            done = done | crossed
    when = torch.where(done, when, window)            # :206-212
    which = torch.where(done, which, torch.full_like(which, 2))  # :215
    rt = (t_nd + when.to(torch.float32) * dt).clamp(1e-6, float(t_max))  # :218
    return torch.stack([rt, which.to(torch.float32)], dim=-1), when, executed


def pack_x(rt_choice: torch.Tensor, log_rt: bool) -> torch.Tensor:
    """rt_choice_model.py:338-342."""
    rt = rt_choice[:, 0:1].to(torch.float32).clamp_min(1e-6)
    if log_rt:
        rt = torch.log(rt)
    return torch.cat([rt, rt_choice[:, 1:2].to(torch.int64).to(torch.float32)], dim=1)


# --------------------------------------------------------------------------------------
# deterministic, platform-independent test noise
# --------------------------------------------------------------------------------------

_SM_GAMMA = np.uint64(0x9E3779B97F4A7C15)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + _SM_GAMMA
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_noise(seed: int, n_steps: int, n_trials: int) -> np.ndarray:
    """(n_steps, n_trials) float32, mean 0 / sd 1, built from integer arithmetic only.

    Each value is (u0+u1+u2+u3 - 131070) * fp32(sqrt(3)/65536) with 16-bit u's taken
    from splitmix64(seed, index): exactly reproducible on any IEEE machine, so golden
    fixtures only need to store the seed.  (Irwin-Hall, not Gaussian: the bit-exact
    parity tests do not care; distribution tests use real normals.)
    """
    idx = np.arange(n_steps * n_trials, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(idx + np.uint64(seed) * np.uint64(0xD1342543DE82EF95))
    total = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
             + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48))).astype(np.int64)
    centred = (total - 131070).astype(np.float32)
    return (centred * np.float32(math.sqrt(3.0) / 65536.0)).reshape(n_steps, n_trials)


# --------------------------------------------------------------------------------------
# NumPy PCG64 pulse sides
# --------------------------------------------------------------------------------------

def pcg64_state(rng: np.random.Generator) -> Tuple[int, int]:
    st = rng.bit_generator.state
    assert st["bit_generator"] == "PCG64"
    return int(st["state"]["state"]), int(st["state"]["inc"])


def p_threshold(p_success: float) -> int:
    """ceil(clip(p,0,1) * 2**53): k*2**-53 < p  <=>  k < threshold (choice_model.py:56,58)."""
    p = min(max(float(p_success), 0.0), 1.0)
    return int(math.ceil(p * 9007199254740992.0))


def pulses_pcg64_c(state: int, inc: int, first_trial: int, n: int, P: int, p_success: float) -> np.ndarray:
    out = np.empty((n, P), dtype=np.float32)
    m = (1 << 64) - 1
    lib().ddm_oracle_pulses_pcg64(state >> 64, state & m, inc >> 64, inc & m, first_trial, n, P,
                                  p_threshold(p_success), out.ctypes.data, P)
    return out


def pulses_loop_numpy(rng: np.random.Generator, n: int, P: int, p_success: float) -> np.ndarray:
    """Per-trial NumPy draws in the reference's order (choice_model.py:56-59, rt_choice_model.py:88-91)."""
    out = np.empty((n, P), dtype=np.float32)
    p = float(np.clip(p_success, 0.0, 1.0))
    for i in range(n):
        if P <= 0:
            continue
        side = 1.0 if rng.random() < 0.5 else -1.0
        ok = rng.random(size=P) < p
        out[i] = np.where(ok, side, -side)
    return out


# --------------------------------------------------------------------------------------
# Philox
# --------------------------------------------------------------------------------------

def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    o = np.empty(4, dtype=np.uint32)
    lib().ddm_oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def philox_words(seed: int, trial_offset: int, n_trials: int, n_steps: int) -> np.ndarray:
    out = np.empty((n_steps, n_trials), dtype=np.uint32)
    lib().ddm_oracle_philox_words(seed, trial_offset, n_trials, n_steps, out.ctypes.data)
    return out


# --------------------------------------------------------------------------------------
# synthetic theta: torch-only stand-in for the pipeline prior
# --------------------------------------------------------------------------------------

def prior_sample(n: int, seed: int = 0) -> torch.Tensor:
    """a0~Beta(2,2), lam~LogNormal(-1,1), v~LogNormal(0,1), B~LogNormal(2.75,.5), tau~Beta(2,2)
    (rt_choice_model_pipeline.py:38-46), drawn with a private generator-free recipe so
    the stream does not depend on the torch version: Beta(2,2) as the median of three
    uniforms, LogNormal from Box-Muller in float64."""
    rs = np.random.RandomState(seed)  # legacy MT19937 stream is frozen by NumPy policy
    u = rs.random_sample((n, 3))
    a0 = np.median(u, axis=1)
    u = rs.random_sample((n, 3))
    tau = np.median(u, axis=1)
    g = rs.standard_normal((n, 3))
    lam = np.exp(-1.0 + g[:, 0])
    v = np.exp(g[:, 1])
    B = np.exp(2.75 + 0.5 * g[:, 2])
    return torch.from_numpy(np.stack([a0, lam, v, B, tau], axis=1).astype(np.float32))


def sim_rng_c(theta, pulses, seed: int, *, dt: float = DT_CHOICE, t_max: float = T_MAX,
              pulse_interval: float = PULSE_INTERVAL, mu_sensory: float = 1.0):
    """Scalar-C simulator with its own CPU normal generator (distribution tests, CPU baseline)."""
    theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float32))
    pulses = np.ascontiguousarray(np.asarray(pulses, dtype=np.float32))
    if theta.ndim == 1:
        theta = theta[None, :]
    if pulses.ndim == 1:
        pulses = pulses[None, :]
    N = theta.shape[0]
    n_max, spp, _ = schedule(dt, t_max, pulse_interval)
    ld_p = 0 if (pulses.shape[0] == 1 and N > 1) else pulses.shape[1]
    sc = fp32_scalars(dt, t_max, mu_sensory)
    x = np.empty((N, 2), dtype=np.float32)
    steps = np.empty((N,), dtype=np.int64)
    L = lib()
    L.ddm_oracle_sim_rng_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
    L.ddm_oracle_sim_rng_f32.restype = ctypes.c_int
    rc = L.ddm_oracle_sim_rng_f32(theta.ctypes.data, theta.shape[1], pulses.ctypes.data, ld_p, N, pulses.shape[1],
                                  n_max, spp, sc["dt"], sc["t_max"], sc["t_nd_hi"], sc["noise_scale"],
                                  ctypes.c_uint64(seed), x.ctypes.data, steps.ctypes.data)
    if rc != 0:
        raise ValueError("ddm_oracle_sim_rng_f32 rejected its arguments")
    return x, steps
