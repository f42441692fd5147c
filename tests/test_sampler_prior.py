"""Host logic of the inference drivers: the torch-only prior and the many-chain slice sampler
(pure torch, so checked on the CPU against closed forms)."""
import math

import numpy as np
import pytest
import torch

from sbi_for_diffusion_models_b200.priors import MultipleIndependentPrior, build_prior_theta
from sbi_for_diffusion_models_b200.samplers import VectorizedSliceSampler


def test_prior_matches_its_components_and_is_minus_inf_outside():
    prior = build_prior_theta()
    torch.manual_seed(0)
    th = prior.sample((1000,))
    assert tuple(th.shape) == (1000, 5) and th.dtype == torch.float32
    assert tuple(prior.sample((1,)).view(5).shape) == (5,)          # the reference's usage (mnle.py:185)
    assert tuple(prior.sample().shape) == (5,)
    lp = prior.log_prob(th)
    from torch.distributions import Beta, LogNormal
    want = (Beta(2.0, 2.0).log_prob(th[:, 0]) + LogNormal(-1.0, 1.0).log_prob(th[:, 1]) + LogNormal(0.0, 1.0).log_prob(th[:, 2])
            + LogNormal(2.75, 0.5).log_prob(th[:, 3]) + Beta(2.0, 2.0).log_prob(th[:, 4]))
    assert torch.allclose(lp, want, atol=1e-5)
    bad = th[:6].clone()
    bad[0, 0], bad[1, 0], bad[2, 1], bad[3, 2], bad[4, 3], bad[5, 4] = -0.1, 1.2, -1.0, 0.0, -3.0, 1.5
    assert torch.equal(prior.log_prob(bad), torch.full((6,), -float("inf")))
    assert math.isfinite(prior.log_prob(th[0]).item())              # 1-D theta
    # the draw order is one sample() per component, in order -> reproducible under a torch seed
    torch.manual_seed(3)
    a = prior.sample((4,))
    torch.manual_seed(3)
    assert torch.equal(a, prior.sample((4,)))
    with pytest.raises(ValueError):
        MultipleIndependentPrior([torch.distributions.Normal(torch.zeros(2), torch.ones(2))])


def test_slice_sampler_recovers_a_correlated_gaussian_and_respects_bounds():
    g = torch.Generator().manual_seed(1)
    cov = torch.tensor([[1.0, 0.8], [0.8, 2.0]], dtype=torch.float64)
    prec = torch.linalg.inv(cov)
    mean = torch.tensor([0.5, -1.0], dtype=torch.float64)

    def logp(x):
        d = x - mean
        return -0.5 * torch.einsum("ni,ij,nj->n", d, prec, d)

    init = torch.randn((256, 2), generator=g, dtype=torch.float64)
    s = VectorizedSliceSampler(logp, init, generator=g)
    draws = s.run(40, warmup=30).reshape(-1, 2)
    assert torch.allclose(draws.mean(0), mean, atol=0.06)
    assert torch.allclose(torch.cov(draws.T), cov, atol=0.12)
    assert s.n_evals > 0

    # truncated target: exp(-x) on (0, 3) x uniform on (-1, 1); -inf outside
    def logq(x):
        ok = (x[:, 0] > 0) & (x[:, 0] < 3) & (x[:, 1] > -1) & (x[:, 1] < 1)
        return torch.where(ok, -x[:, 0], torch.full_like(x[:, 0], -float("inf")))

    init = torch.stack([torch.rand(256, generator=g, dtype=torch.float64) * 3, torch.rand(256, generator=g, dtype=torch.float64) * 2 - 1], 1)
    d2 = VectorizedSliceSampler(logq, init, generator=g).run(40, warmup=20).reshape(-1, 2)
    assert float(d2[:, 0].min()) > 0 and float(d2[:, 0].max()) < 3 and float(d2[:, 1].abs().max()) < 1
    want_mean = (1 - 4 * math.exp(-3)) / (1 - math.exp(-3))           # E[x] of exp(-x) truncated to (0, 3)
    assert abs(float(d2[:, 0].mean()) - want_mean) < 0.05
    assert abs(float(d2[:, 1].var()) - 1 / 3) < 0.03
    with pytest.raises(ValueError, match="finite"):
        VectorizedSliceSampler(logq, torch.full((4, 2), 5.0, dtype=torch.float64))


def test_slice_sampler_is_uniform_in_rank_for_a_known_posterior():
    """SBC in miniature on a conjugate model: theta ~ N(0,1), x | theta ~ N(theta, 1); the rank of
    theta_true among posterior draws must be uniform."""
    g = torch.Generator().manual_seed(5)
    n_data, n_draws = 400, 31
    theta_true = torch.randn(n_data, generator=g, dtype=torch.float64)
    x = theta_true + torch.randn(n_data, generator=g, dtype=torch.float64)

    def logp(th):     # one chain per dataset
        return -0.5 * th[:, 0] ** 2 - 0.5 * (x - th[:, 0]) ** 2

    s = VectorizedSliceSampler(logp, torch.zeros((n_data, 1), dtype=torch.float64), init_width=1.0, generator=g)
    s.width[:] = 1.0
    assert tuple(s.width.shape) == (n_data, 1)            # every chain tunes its own widths
    draws = s.run(n_draws, warmup=20, thin=3)[:, :, 0]                   # (n_draws, n_data)
    ranks = (draws < theta_true[None, :]).sum(0).numpy()
    hist = np.bincount(ranks // 4, minlength=8)[:8]                     # 32 possible ranks -> 8 bins
    chi2 = float(((hist - n_data / 8) ** 2 / (n_data / 8)).sum())
    assert chi2 < 24.3, (chi2, hist)                                     # chi2(7) 0.999 quantile
