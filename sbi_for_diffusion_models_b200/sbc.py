"""SBC building blocks on the hot path (reference mnle.py:98-104 and :183-206).

``run_sbc`` itself is third-party-sampler glue and stays out of scope; what it does per
dataset on this path -- draw (theta_true, ds_seed), simulate a session, rank posterior samples
-- is provided here in batched, shardable form: all sessions of a shard run in ONE simulator
launch.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .pulses import generate_pulse_matrix_device
from .simulator import Schedule, compute_device, simulate_trials


def compute_ranks(theta_true: torch.Tensor, posterior_samples: torch.Tensor) -> torch.Tensor:
    """rank_d = #{s : samples[s, d] < theta_true[d]}  (reference mnle.py:98-104)."""
    return (posterior_samples < theta_true.view(-1)[None, :]).sum(dim=0).to(torch.int64)


def draw_sbc_datasets(prior_theta, num_datasets: int, seed: int = 0) -> Tuple[torch.Tensor, np.ndarray]:
    """Pre-draw every dataset's (theta_true, ds_seed) in the reference's order
    (mnle.py:161-162 seeds, :185 prior draw, :188 ds_seed) so that datasets can be sharded:
    the i-th pair is what the reference's sequential loop would have used."""
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    thetas, seeds = [], []
    for _ in range(num_datasets):
        thetas.append(prior_theta.sample((1,)).view(5).to(torch.float32).cpu())
        seeds.append(int(rng.integers(0, 2**31 - 1)))
    return torch.stack(thetas), np.asarray(seeds, dtype=np.int64)


@torch.no_grad()
def simulate_sbc_sessions(thetas_true: torch.Tensor, ds_seeds, num_trials: int, *, mu_sensory: float,
                          p_success: float, noise_seed: int = 0, first_dataset: int = 0, device=None):
    """Sessions for datasets [first_dataset, first_dataset + D): x (D, T, 2) raw [rt, choice] and
    pulses (D, T, P) on the compute device.  Pulses of dataset i come from
    ``default_rng(ds_seeds[i])`` exactly as in ``simulate_session_data_rt_choice``
    (reference rt_choice_model.py:312-316); all D*T trials run in one launch, Philox-indexed by
    global trial number (first_dataset + i) * T + t so results do not depend on the sharding."""
    dev = compute_device(device)
    sched = Schedule.from_constants(mu_sensory)
    D, T, P = thetas_true.shape[0], int(num_trials), sched.n_pulses
    pulses = torch.empty((D, T, P), dtype=torch.float32, device=dev)
    for i in range(D):
        generate_pulse_matrix_device(np.random.default_rng(int(ds_seeds[i])), T, P, p_success=p_success, out=pulses[i])
    theta_rows = thetas_true.to(device=dev, dtype=torch.float32).repeat_interleave(T, dim=0)
    x = simulate_trials(theta_rows, pulses.view(D * T, P), mu_sensory=mu_sensory, seed=noise_seed,
                        trial_offset=first_dataset * T, schedule=sched, device=dev)
    return x.view(D, T, 2), pulses
