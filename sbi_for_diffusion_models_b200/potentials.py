"""Posterior potential over the five global parameters (reference potentials.py:7-117) with
the MNLE likelihood evaluated by the fused CUDA kernel.

``ConditionedMNLELogLikelihood(estimator, local_theta, device)`` and
``ThetaOnlyPosteriorPotential(conditioned_loglike=..., prior_theta=..., x_o=..., device=...,
temperature=...)`` keep the reference's constructor arguments, methods and return shapes, so
``run_inference_mcmc`` (reference mnle.py:61-87) can build them unchanged.  The reference
expands (C thetas) x (T trials) into T*C rows of 85 numbers and calls ``estimator.log_prob``
(potentials.py:100-113); here nothing is materialised: the kernel assembles row r = t*C + c
from ``theta[c]``, ``local_theta[t]`` and ``x_o[t]`` and returns the sum over t.

Gradients: a call that needs d loglik / d theta (NUTS: ``track_gradients=True`` with a theta that
requires grad, reference potentials.py:33, 112) runs value and the five partials in reverse mode on the
tensor cores (``mnle_loglik_sum_grad_tc_f32``; ``DeviceMNLE.loglik_sum_and_grad`` also offers the fp32
forward-mode kernel ``mnle_loglik_sum_grad_f32``) behind a ``torch.autograd.Function``; everything else
runs the tensor-core forward kernel.
"""
from __future__ import annotations

import torch
from torch.distributions import Distribution

from .mnle_net import DeviceMNLE, PackedMNLE


_converted = {}   # id(source object) -> (weak reference to it, its DeviceMNLE): convert a foreign estimator once


def as_device_estimator(estimator) -> DeviceMNLE:
    """Accept our estimator types, or anything exposing an MNLE-shaped ``state_dict()`` (converted once per
    source object: every conversion allocates device copies of the weights)."""
    import weakref
    if isinstance(estimator, DeviceMNLE):
        return estimator
    if isinstance(estimator, PackedMNLE) or hasattr(estimator, "state_dict"):
        hit = _converted.get(id(estimator))
        if hit is not None and hit[0]() is estimator:
            return hit[1]
        dev = DeviceMNLE(estimator if isinstance(estimator, PackedMNLE) else PackedMNLE.from_state_dict(estimator.state_dict()))
        try:
            key = id(estimator)
            _converted[key] = (weakref.ref(estimator, lambda _r, key=key: _converted.pop(key, None)), dev)
        except TypeError:     # not weak-referenceable: no cache
            pass
        return dev
    raise TypeError(f"cannot run {type(estimator).__name__} on the GPU: expected DeviceMNLE, PackedMNLE or a module "
                    "whose state_dict() holds the reference's MNLE architecture")


def prior_log_prob(prior, theta: torch.Tensor) -> torch.Tensor:
    """``prior.log_prob(theta)`` on theta's device; priors whose parameters live on the CPU (the
    reference builds its prior there, rt_choice_model_pipeline.py:38-46) are evaluated on the CPU
    and the result is moved back."""
    try:
        return prior.log_prob(theta)
    except RuntimeError:
        if theta.device.type == "cpu":
            raise
        return prior.log_prob(theta.cpu()).to(theta.device)


class _LoglikSumWithGrad(torch.autograd.Function):
    """sum_t log p(x_t | theta_c, pulses_t) with its Jacobian taken from the gradient kernel (value and
    d out[c] / d theta[c] in one call): each output depends on its own theta row only, so
    d L / d theta[c] = grad_output[c] * d out[c] / d theta[c]."""

    @staticmethod
    def forward(ctx, theta, estimator, x, pulses):
        out, grad = estimator.loglik_sum_and_grad(theta.to(dtype=torch.float32), x, pulses)
        ctx.save_for_backward(grad)
        ctx.theta_dtype = theta.dtype
        return out.to(theta.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        (grad,) = ctx.saved_tensors
        return (grad_output.to(grad.dtype).unsqueeze(1) * grad).to(ctx.theta_dtype), None, None, None


class ConditionedMNLELogLikelihood(torch.nn.Module):
    """sum_t log p(x_t | global_theta, local_theta_t) for every row of ``global_theta``.
    Pickles as CPU data only (packed weights + the pulse buffer)."""

    def __init__(self, estimator, local_theta: torch.Tensor, device: str = "cpu"):
        super().__init__()
        self.estimator = as_device_estimator(estimator)
        self.device = device
        self.register_buffer("local_theta", local_theta.to(device=device, dtype=torch.float32))

    def forward(self, global_theta: torch.Tensor, x_o: torch.Tensor, track_gradients: bool = True) -> torch.Tensor:
        x = x_o.to(dtype=torch.float32)
        if x.dim() == 3:
            assert x.shape[1] == 1, "This implementation supports a single observed x batch (num_xs=1)."
            x = x[:, 0, :]
        num_trials = x.shape[0]
        assert self.local_theta.shape[0] == num_trials, (
            f"local_theta must have shape (num_trials, P). Got {tuple(self.local_theta.shape)}")
        if track_gradients and torch.is_grad_enabled() and global_theta.requires_grad:
            return _LoglikSumWithGrad.apply(global_theta, self.estimator, x, self.local_theta).to(self.device)
        theta = global_theta.detach().to(dtype=torch.float32)
        ll = self.estimator.loglik_sum(theta, x, self.local_theta)
        return ll.to(self.device)


class ThetaOnlyPosteriorPotential:
    """log prior(theta) + loglik(theta) / temperature; rows outside the prior's support are
    returned as-is (-inf) without evaluating the likelihood (reference potentials.py:33-57)."""

    def __init__(self, *, conditioned_loglike, prior_theta: Distribution, x_o: torch.Tensor, device: str = "cpu",
                 temperature: float = 1.0):
        self.conditioned_loglike = conditioned_loglike
        self.prior_theta = prior_theta
        self._x_o = x_o.to(device=device, dtype=torch.float32)
        self.device = device
        self.temperature = float(temperature)

    def return_x_o(self):
        return self._x_o

    def set_x_o(self, x_o: torch.Tensor):
        self._x_o = x_o.to(self.device, dtype=torch.float32)
        return self

    def set_x(self, x: torch.Tensor):
        return self.set_x_o(x)

    def __call__(self, theta: torch.Tensor, x_o: torch.Tensor = None, track_gradients: bool = True) -> torch.Tensor:
        if x_o is not None:
            self.set_x_o(x_o)
        if theta.ndim == 1:
            theta = theta.view(1, -1)
        theta = theta.to(self.device, dtype=torch.float32)
        log_prior = prior_log_prob(self.prior_theta, theta)
        inside = torch.isfinite(log_prior)
        if not bool(inside.any()):
            return log_prior
        with torch.set_grad_enabled(bool(track_gradients)):
            ll = self.conditioned_loglike(theta[inside], self._x_o, track_gradients=bool(track_gradients)).reshape(-1)
        out = log_prior.clone()
        out[inside] = out[inside] + ll / self.temperature
        return out
