"""Drop-in for the reference's ``models/rt_choice_model.py`` public functions, backed by the
sm_100a kernels.  Same names, argument meaning, return layout and ``ValueError``s; the time
loop (reference :181-204) runs in ``ddm_sim_f32`` and pulse sampling (reference :88-91) in
``ddm_pulses_pcg64``.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from .. import constants
from ..pulses import generate_pulse_matrix_device
from ..run_config import RUN_CONFIG_PARAMS
from ..simulator import Schedule, simulate_trials

cfg = RUN_CONFIG_PARAMS


def pulse_schedule(*, dt: Optional[float] = None) -> Tuple[int, int]:
    """(n_max, steps_per_pulse) of the Euler grid (reference :45-54)."""
    s = Schedule.from_constants(dt=dt)
    return s.n_max, s.steps_per_pulse


def n_pulses_max_from_schedule(n_max: int, steps_per_pulse: int) -> int:
    """Pulse slots a trial of n_max steps can reach (reference :57-59)."""
    return -(-int(n_max) // int(steps_per_pulse))


def generate_pulse_matrix_numpy(rng: np.random.Generator, n_trials: int, n_pulses: int, *,
                                p_success: float = cfg.P_SUCCESS) -> np.ndarray:
    """(n_trials, n_pulses) float32 of +-1, identical to the reference's matrix for the same
    ``rng`` (reference :62-91); generated on the GPU, ``rng`` advanced accordingly."""
    if n_trials < 0:
        raise ValueError("n_trials must be >= 0")
    if n_pulses < 0:
        raise ValueError("n_pulses must be >= 0")
    return generate_pulse_matrix_device(rng, n_trials, n_pulses, p_success=p_success).cpu().numpy()


def as_pulse_tensor(pulse_sides: Union[np.ndarray, Tensor], *, device, dtype: torch.dtype = torch.float32) -> Tensor:
    """(N,P) tensor view of a pulse matrix or a single (P,) train (reference :94-109)."""
    s = pulse_sides if isinstance(pulse_sides, Tensor) else torch.from_numpy(np.asarray(pulse_sides))
    if s.ndim == 1:
        s = s.view(1, -1)
    if s.ndim != 2:
        raise ValueError(f"pulse_sides must have shape (N,P) or (P,), got {tuple(s.shape)}")
    return s.to(device=device, dtype=dtype)


def rt_choice_model_simulator_torch(theta: Tensor, rng: np.random.Generator | None = None, *,
                                    mu_sensory: float = 1.0,
                                    pulse_sides: Optional[Union[np.ndarray, Tensor]] = None,
                                    p_success: float = cfg.P_SUCCESS, seed: Optional[int] = None,
                                    noise: Optional[Tensor] = None) -> Tensor:
    """theta (N,5) or (5,) -> x (N,2) fp32 ``[rt, choice in {0,1,2}]`` on theta's device
    (reference :251-283).  ``pulse_sides=None`` samples the stimulus from ``rng``
    (marginalising it, reference :157-163).  ``seed`` / ``noise`` are extensions: the Philox
    key (default: drawn from torch's global generator) and shared noise for exact replay."""
    if theta.ndim == 1:
        theta = theta.view(1, -1)
    if theta.shape[-1] != 5:
        raise ValueError(f"Expected theta shape (N,5) or (5,), got {tuple(theta.shape)}")
    sched = Schedule.from_constants(mu_sensory)
    if pulse_sides is None:
        if rng is None:
            rng = np.random.default_rng()
        pulse_sides = generate_pulse_matrix_device(rng, theta.shape[0], sched.n_pulses, p_success=p_success)
    x = simulate_trials(theta, pulse_sides, mu_sensory=float(mu_sensory), seed=seed, noise=noise, schedule=sched)
    return x.to(theta.device)


def rt_choice_model_simulator(theta: np.ndarray, rng: np.random.Generator, *, mu_sensory: float = 1.0,
                              pulse_sides: Optional[Union[np.ndarray, Tensor]] = None,
                              p_success: float = cfg.P_SUCCESS) -> tuple[float, int]:
    """Single-trial NumPy API (reference :224-248)."""
    th = torch.tensor(np.asarray(theta), dtype=torch.float32).view(1, 5)
    x = rt_choice_model_simulator_torch(th, rng, mu_sensory=float(mu_sensory), pulse_sides=pulse_sides,
                                        p_success=float(p_success))
    return float(x[0, 0]), int(x[0, 1])


def simulate_session_data_rt_choice(theta_true: Tensor, num_trials: int, rng: np.random.Generator | None = None, *,
                                    mu_sensory: float = 1.0,
                                    pulse_sides: Optional[Union[np.ndarray, Tensor]] = None,
                                    p_success: float = cfg.P_SUCCESS, return_pulse_sides: bool = False,
                                    seed: Optional[int] = None, noise: Optional[Tensor] = None,
                                    ) -> Union[Tensor, Tuple[Tensor, Tensor]]:
    """One session of ``num_trials`` iid trials at ``theta_true``: raw ``[rt, choice]``
    (reference :286-329); with ``return_pulse_sides`` also the (num_trials, P) stimulus."""
    if rng is None:
        rng = np.random.default_rng()
    theta_rep = theta_true.view(1, -1).to(torch.float32).expand(num_trials, -1)
    if pulse_sides is None:
        sched = Schedule.from_constants(mu_sensory)
        pulse_sides = generate_pulse_matrix_device(rng, num_trials, sched.n_pulses, p_success=p_success)
    x = rt_choice_model_simulator_torch(theta_rep, rng=rng, mu_sensory=mu_sensory, pulse_sides=pulse_sides,
                                        p_success=p_success, seed=seed, noise=noise)
    if return_pulse_sides:
        return x, as_pulse_tensor(pulse_sides, device=x.device, dtype=torch.float32)
    return x


def pack_x_rt_choice(rt_choice: torch.Tensor, *, log_rt: bool) -> torch.Tensor:
    """``[rt, choice]`` -> ``[max(rt,1e-6) or its log, choice]`` (reference :332-342).  The
    simulator entry points fuse this into the kernel epilogue; this helper is for callers
    that already hold raw simulator output."""
    rt = rt_choice[:, 0:1].to(torch.float32).clamp_min(1e-6)
    choice = rt_choice[:, 1:2].to(torch.int64).to(torch.float32)
    return torch.cat([torch.log(rt) if log_rt else rt, choice], dim=1)
