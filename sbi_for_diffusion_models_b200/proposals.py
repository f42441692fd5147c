"""Training proposals over z = [theta(5), pulse_sides(P)] (reference proposals.py:9-74).

Same classes and semantics; the pulse part is generated on the GPU from the proposal's
NumPy PCG64 stream (bit-identical to the reference's draws) and, when the proposal is
given a CUDA ``device``, z is assembled in place on that device with no host round trip.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.distributions import Distribution

from .pulses import generate_pulse_matrix_device


class PulseSequenceProposal(Distribution):
    """Distribution over +-1 pulse trains of length P: one 50/50 correct side per train, each
    pulse on that side with probability ``p_success``.  Sampling only; ``log_prob`` is 0."""

    arg_constraints = {}
    has_rsample = False

    def __init__(self, P: int, p_success: float, seed: int = 0, device=None):
        super().__init__(validate_args=False)
        self.P = int(P)
        self.p_success = float(p_success)
        self.rng = np.random.default_rng(seed)
        self._device = device

    @property
    def event_shape(self):
        return torch.Size([self.P])

    def sample_into(self, out: torch.Tensor) -> torch.Tensor:
        """Fill a CUDA (n, P) fp32 view (any row stride) with the next n trains of the stream."""
        return generate_pulse_matrix_device(self.rng, out.shape[0], self.P, p_success=self.p_success, out=out)

    def sample(self, sample_shape=torch.Size()):
        shape = tuple(sample_shape)
        n = int(np.prod(shape)) if len(shape) > 0 else 1
        s = generate_pulse_matrix_device(self.rng, n, self.P, p_success=self.p_success)
        if len(shape) > 0:
            s = s.view(*shape, self.P)
        # reference quirk kept: sample(()) has shape (1, P), not (P,)
        target = self._device if self._device is not None else "cpu"
        return s.to(target)

    def log_prob(self, value):
        return torch.zeros(value.shape[:-1], device=value.device, dtype=torch.float32)


class ExtendedProposal(Distribution):
    """Joint proposal over z = [theta, pulses]; theta comes from the user's prior."""

    arg_constraints = {}
    has_rsample = False

    def __init__(self, theta_prior: Distribution, pulse_proposal: PulseSequenceProposal, device=None):
        super().__init__(validate_args=False)
        self.theta_prior = theta_prior
        self.pulse_proposal = pulse_proposal
        self._device = device

    @property
    def event_shape(self):
        return torch.Size([5 + self.pulse_proposal.P])

    def sample(self, sample_shape=torch.Size()):
        shape = tuple(sample_shape)
        theta = self.theta_prior.sample(sample_shape)
        P = self.pulse_proposal.P
        on_gpu = self._device is not None and torch.device(self._device).type == "cuda"
        if on_gpu and len(shape) > 0:
            n = int(np.prod(shape))
            z = torch.empty((n, 5 + P), dtype=torch.float32, device=self._device)
            z[:, :5] = theta.reshape(n, 5).to(device=z.device, dtype=torch.float32, non_blocking=True)
            self.pulse_proposal.sample_into(z[:, 5:])
            return z.view(*shape, 5 + P)
        pulses = self.pulse_proposal.sample(sample_shape)
        z = torch.cat([theta.to(torch.float32).to(pulses.device), pulses.to(torch.float32)], dim=-1)
        return z if self._device is None else z.to(self._device)

    def log_prob(self, z):
        return self.theta_prior.log_prob(z[..., :5]) + self.pulse_proposal.log_prob(z[..., 5:])
