"""Prior over the five global parameters theta = [a0, lam, v, B, tau].

The reference builds it with ``sbi.utils.MultipleIndependent`` over five 1-D torch
distributions (rt_choice_model_pipeline.py:34-46).  sbi is not part of this package, so the
same object is provided here on torch alone: ``sample((n,)) -> (n, 5)`` concatenates the
components' draws in order (one ``sample`` call per component, like MultipleIndependent), and
``log_prob`` sums the components and is ``-inf`` outside the support, which is what the potential
relies on (reference potentials.py:43-46).
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch.distributions import Beta, Distribution, LogNormal


class MultipleIndependentPrior(Distribution):
    """Product of independent 1-D distributions, event shape (len(dists),)."""

    arg_constraints = {}
    has_rsample = False

    def __init__(self, dists: Sequence[Distribution]):
        self.dists = list(dists)
        for d in self.dists:
            if tuple(d.batch_shape) not in ((), (1,)) or tuple(d.event_shape) != ():
                raise ValueError("every component must be a scalar (batch shape () or (1,)) distribution")
        super().__init__(batch_shape=torch.Size(), event_shape=torch.Size([len(self.dists)]), validate_args=False)

    def sample(self, sample_shape=torch.Size()) -> torch.Tensor:
        shape = torch.Size(sample_shape)
        cols = [_sample_component(d, shape).reshape(shape + (1,)) for d in self.dists]
        return torch.cat(cols, dim=-1)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        if value.shape[-1] != len(self.dists):
            raise ValueError(f"last dimension must be {len(self.dists)}, got {tuple(value.shape)}")
        fused = self._fused(value.device)
        if fused is not None:
            # Beta / LogNormal columns in closed form over the whole (N, D) matrix: a dozen launches
            # instead of ~8 per component (this sits inside every sampler step)
            is_beta, a1, b1, lbeta, mu, inv2s2, lnc = fused
            x = value.to(a1.dtype)
            ok = (x > 0) & ((x < 1) | ~is_beta)
            xs = torch.where(ok, x, 0.5)
            logx = torch.log(xs)
            beta = a1 * logx + b1 * torch.log1p(-torch.where(is_beta, xs, 0.5)) - lbeta
            logn = -logx - lnc - (logx - mu) ** 2 * inv2s2
            total = torch.where(is_beta, beta, logn).sum(-1)
            return torch.where(ok.all(-1), total, -float("inf")).to(value.dtype)
        total = torch.zeros(value.shape[:-1], dtype=value.dtype, device=value.device)
        inside = torch.ones(value.shape[:-1], dtype=torch.bool, device=value.device)
        for i, d in enumerate(self._components(value.device)):
            v = value[..., i]
            ok = d.support.check(v)
            inside &= ok
            safe = torch.where(ok, v, self._interior[i])
            total = total + d.log_prob(safe).reshape(v.shape)
        return torch.where(inside, total, torch.full_like(total, -float("inf")))

    def _fused(self, device):
        """Per-column constants of the closed-form path, or None if a component is neither Beta nor
        LogNormal."""
        key = "fused:" + str(device)
        cache = self.__dict__.setdefault("_by_device", {})
        if key not in cache:
            if not all(isinstance(d, (Beta, LogNormal)) for d in self.dists):
                cache[key] = None
            else:
                f = lambda xs: torch.tensor(xs, dtype=torch.float32, device=device)
                isb = [isinstance(d, Beta) for d in self.dists]
                a = [float(d.concentration1.reshape(-1)[0]) if b else 1.0 for d, b in zip(self.dists, isb)]
                bb = [float(d.concentration0.reshape(-1)[0]) if b else 1.0 for d, b in zip(self.dists, isb)]
                mu = [0.0 if b else float(d.loc.reshape(-1)[0]) for d, b in zip(self.dists, isb)]
                sg = [1.0 if b else float(d.scale.reshape(-1)[0]) for d, b in zip(self.dists, isb)]
                lbeta = [math.lgamma(x) + math.lgamma(y) - math.lgamma(x + y) for x, y in zip(a, bb)]
                cache[key] = (torch.tensor(isb, device=device), f(a) - 1.0, f(bb) - 1.0, f(lbeta), f(mu),
                              f([0.5 / (s_ * s_) for s_ in sg]), f([math.log(s_) + 0.5 * math.log(2 * math.pi) for s_ in sg]))
        return cache[key]

    def _components(self, device):
        """Component distributions with their parameters on ``device``, built once per device (no
        host-to-device copies on the hot path, so the potential can be recorded in a CUDA graph)."""
        key = str(device)
        cache = self.__dict__.setdefault("_by_device", {})
        if key not in cache:
            cache[key] = [_on_device(d, device) for d in self.dists]
            self._interior = [_interior_point(d) for d in self.dists]
        return cache[key]

    @property
    def support(self):
        raise NotImplementedError("use log_prob(...) == -inf to test the support")


_cuda_params = {}   # id(distribution) -> (distribution, kind, python floats): read once, no device sync per draw


def _sample_component(d: Distribution, shape: torch.Size) -> torch.Tensor:
    """``d.sample(shape)``, except for Beta / LogNormal components whose parameters live on a CUDA device:

    * torch's samplers synchronise the host with the device on every call (``torch.normal`` checks ``std.min() >= 0``
      with ``.item()``), which serialises a training-set loop that otherwise only enqueues work;
    * torch's CUDA Beta sampler goes through per-element gamma rejection loops (~15 ns per draw) and made the prior,
      not the simulator or the PCIe link, the bottleneck of a 1e7-trial training set.

    There a LogNormal is drawn as exp(loc + scale * randn) and a Beta(a, b) with small integer parameters as the a-th
    smallest of a + b - 1 uniforms (its exact law), from parameters read to the host once.  CPU priors (the
    reference's) keep torch's samplers and therefore torch's random stream."""
    if isinstance(d, (Beta, LogNormal)):
        ref = d.concentration1 if isinstance(d, Beta) else d.loc
        if ref.is_cuda and ref.numel() == 1:
            hit = _cuda_params.get(id(d))
            if hit is None or hit[0] is not d:
                if isinstance(d, Beta):
                    hit = (d, "beta", float(d.concentration1.reshape(-1)[0]), float(d.concentration0.reshape(-1)[0]))
                else:
                    hit = (d, "lognormal", float(d.loc.reshape(-1)[0]), float(d.scale.reshape(-1)[0]))
                _cuda_params[id(d)] = hit
            _, kind, p0, p1 = hit
            if kind == "lognormal":
                return torch.randn(tuple(shape), device=ref.device, dtype=ref.dtype).mul_(p1).add_(p0).exp_()
            if p0 == int(p0) and p1 == int(p1) and 1 <= p0 and 1 <= p1 and p0 + p1 <= 6:
                n = int(p0 + p1 - 1)
                u = torch.rand(tuple(shape) + (n,), device=ref.device, dtype=ref.dtype)
                if n == 1:
                    return u[..., 0]
                if n == 3 and p0 == 2:      # the pipeline prior's Beta(2, 2): the median of three uniforms
                    a, b, c = u.unbind(-1)
                    return torch.maximum(torch.minimum(a, b), torch.minimum(torch.maximum(a, b), c))
                return torch.sort(u, dim=-1).values[..., int(p0) - 1]
    return d.sample(shape)


def _on_device(d: Distribution, device) -> Distribution:
    """The same distribution with its parameters on ``device`` (Beta / LogNormal only need this)."""
    if isinstance(d, Beta):
        return Beta(d.concentration1.to(device), d.concentration0.to(device), validate_args=False)
    if isinstance(d, LogNormal):
        return LogNormal(d.loc.to(device), d.scale.to(device), validate_args=False)
    return d


def _interior_point(d: Distribution) -> float:
    if isinstance(d, Beta):
        return 0.5
    if isinstance(d, LogNormal):
        return 1.0
    return float(d.mean.reshape(()).detach())


def build_prior_theta(device=None) -> MultipleIndependentPrior:
    """a0 ~ Beta(2,2), lam ~ LogNormal(-1,1), v ~ LogNormal(0,1), B ~ LogNormal(2.75,0.5),
    tau ~ Beta(2,2)  (reference rt_choice_model_pipeline.py:38-46).  ``device``: where the components'
    parameters live, i.e. where ``sample`` draws (the reference's prior is a CPU object; a CUDA one keeps
    1e7-1e9-trial proposal draws off the host)."""
    t = lambda v: torch.tensor([v], device=device)
    return MultipleIndependentPrior([
        Beta(t(2.0), t(2.0)),
        LogNormal(t(-1.0), t(1.0)),
        LogNormal(t(0.0), t(1.0)),
        LogNormal(t(2.75), t(0.5)),
        Beta(t(2.0), t(2.0)),
    ])
