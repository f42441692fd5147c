"""Prior over the five global parameters theta = [a0, lam, v, B, tau].

The reference builds it with ``sbi.utils.MultipleIndependent`` over five 1-D torch
distributions (rt_choice_model_pipeline.py:34-46).  sbi is not part of this package, so the
same object is provided here on torch alone: ``sample((n,)) -> (n, 5)`` concatenates the
components' draws in order (one ``sample`` call per component, like MultipleIndependent), and
``log_prob`` sums the components and is ``-inf`` outside the support, which is what the potential
relies on (reference potentials.py:43-46).
"""
from __future__ import annotations

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
        cols = [d.sample(shape).reshape(shape + (1,)) for d in self.dists]
        return torch.cat(cols, dim=-1)

    def log_prob(self, value: torch.Tensor) -> torch.Tensor:
        if value.shape[-1] != len(self.dists):
            raise ValueError(f"last dimension must be {len(self.dists)}, got {tuple(value.shape)}")
        total = torch.zeros(value.shape[:-1], dtype=value.dtype, device=value.device)
        inside = torch.ones(value.shape[:-1], dtype=torch.bool, device=value.device)
        for i, d in enumerate(self.dists):
            v = value[..., i]
            ok = _on_device(d, v.device).support.check(v)
            inside &= ok
            safe = torch.where(ok, v, _interior_point(d).to(v))
            total = total + _on_device(d, v.device).log_prob(safe).reshape(v.shape)
        return torch.where(inside, total, torch.full_like(total, -float("inf")))

    @property
    def support(self):
        raise NotImplementedError("use log_prob(...) == -inf to test the support")


def _on_device(d: Distribution, device) -> Distribution:
    """The same distribution with its parameters on ``device`` (Beta / LogNormal only need this)."""
    if isinstance(d, Beta):
        return Beta(d.concentration1.to(device), d.concentration0.to(device), validate_args=False)
    if isinstance(d, LogNormal):
        return LogNormal(d.loc.to(device), d.scale.to(device), validate_args=False)
    return d


def _interior_point(d: Distribution) -> torch.Tensor:
    if isinstance(d, Beta):
        return torch.tensor(0.5)
    if isinstance(d, LogNormal):
        return torch.tensor(1.0)
    return d.mean.reshape(()).detach()


def build_prior_theta() -> MultipleIndependentPrior:
    """a0 ~ Beta(2,2), lam ~ LogNormal(-1,1), v ~ LogNormal(0,1), B ~ LogNormal(2.75,0.5),
    tau ~ Beta(2,2)  (reference rt_choice_model_pipeline.py:38-46)."""
    return MultipleIndependentPrior([
        Beta(torch.tensor([2.0]), torch.tensor([2.0])),
        LogNormal(torch.tensor([-1.0]), torch.tensor([1.0])),
        LogNormal(torch.tensor([0.0]), torch.tensor([1.0])),
        LogNormal(torch.tensor([2.75]), torch.tensor([0.5])),
        Beta(torch.tensor([2.0]), torch.tensor([2.0])),
    ])
