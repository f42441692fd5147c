"""CPU specification of the MNLE density estimator's ``log_prob`` -- TEST INFRASTRUCTURE.

PARITY UNPINNED as a whole; the SPLINE (``rqs_forward``, where the numerics are) is pinned against an independent
third-party implementation of the same routine: ``transformers.models.vits.modeling_vits.
_unconstrained_rational_quadratic_spline`` (in this image; VITS took it from nflows) agrees with it to 4e-15 in
float64 and a few ulp in float32 (``tests/test_mnle_spec.py``).  What stays unpinned is the wiring around the spline
(layer shapes, z-scoring, the categorical head, the state_dict names).

The arithmetic of this path lives in third-party packages that are not
under /root/reference and are not installed here (no network): ``sbi==0.25.0``
(``MixedDensityEstimator`` built by ``likelihood_nn(model="mnle", ...)``), which uses
``nflows==0.14`` / ``pyknos==0.16.0`` for the neural spline flow (pins: reference
``uv.lock``; call sites ``mnle.py:31-39`` (builder), ``potentials.py:113`` (log_prob)).  The
reference ships no tests, weights or golden log-probs for it.  This module restates the
published algorithms those packages implement:

* Durkan et al. 2019, "Neural Spline Flows" -- monotone rational-quadratic spline with
  linear tails (nflows ``rational_quadratic_spline`` / ``unconstrained_...``);
* Boelts et al. 2022, "Flexible and efficient simulation-based inference for models of
  decision-making" -- MNLE = categorical net for the choice x conditional flow for log RT;

with the hyper-parameters of the reference's call (``hidden_features=128, num_transforms=10,
num_bins=24, log_transform_x=True, z_score_theta="independent", z_score_x="independent"``) and
sbi's defaults for the rest (2 hidden layers + sigmoid in the categorical net, one hidden
128x128 layer + ReLU in each spline conditioner, tail_bound 10, min bin width / height /
derivative 1e-3).  It is the oracle for the CUDA kernel; agreement of this spec with a real
sbi estimator can only be checked by someone who has sbi (``tools/compare_with_sbi.py``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn.functional as F

HIDDEN = 128
NUM_BINS = 24
NUM_TRANSFORMS = 10
COND_DIM = 85            # theta(5) + pulses(80)
CTX_DIM = COND_DIM + 1   # + choice
SPLINE_OUT = 3 * NUM_BINS - 1
TAIL_BOUND = 10.0
MIN_BIN = 1e-3
MIN_DERIV = 1e-3
PROB_EPS = float(torch.finfo(torch.float32).eps)   # Categorical(probs=...) clamps to [eps, 1-eps]


def init_params(seed: int = 0, n_choices: int = 3, scale: float = 1.0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random parameters in the layout the CUDA packer consumes.  ``scale`` > 1 sharpens
    the splines (larger conditioner outputs) to stress the numerics like a trained net does."""
    g = torch.Generator().manual_seed(seed)

    def lin(n_out, n_in, s=1.0):
        bound = s / math.sqrt(n_in)
        w = (torch.rand((n_out, n_in), generator=g, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand((n_out,), generator=g, dtype=torch.float64) * 2 - 1) * bound
        return w.to(dtype), b.to(dtype)

    p: Dict[str, torch.Tensor] = {}
    # z-scoring buffers: theta stats like the pipeline prior's, pulses ~ mean 0 / std 1
    p["cond_mean"] = torch.cat([torch.tensor([0.5, 0.6, 1.6, 17.7, 0.5], dtype=torch.float64),
                                torch.zeros(80, dtype=torch.float64)]).to(dtype)
    p["cond_std"] = torch.cat([torch.tensor([0.22, 0.8, 2.1, 9.4, 0.22], dtype=torch.float64),
                               torch.ones(80, dtype=torch.float64)]).to(dtype)
    p["cat.W0"], p["cat.b0"] = lin(HIDDEN, COND_DIM, scale)
    p["cat.W1"], p["cat.b1"] = lin(HIDDEN, HIDDEN, scale)
    p["cat.W2"], p["cat.b2"] = lin(HIDDEN, HIDDEN, scale)
    p["cat.Wo"], p["cat.bo"] = lin(n_choices, HIDDEN, scale)
    p["flow.mu_y"] = torch.tensor(0.35, dtype=dtype)      # mean / std of log rt in a training set
    p["flow.sigma_y"] = torch.tensor(1.1, dtype=dtype)
    for k in range(NUM_TRANSFORMS):
        p[f"flow.{k}.W1"], p[f"flow.{k}.b1"] = lin(HIDDEN, CTX_DIM, scale)
        p[f"flow.{k}.W2"], p[f"flow.{k}.b2"] = lin(HIDDEN, HIDDEN, scale)
        p[f"flow.{k}.W3"], p[f"flow.{k}.b3"] = lin(SPLINE_OUT, HIDDEN, 4.0 * scale)
    return p


def cast_params(p: Dict[str, torch.Tensor], dtype) -> Dict[str, torch.Tensor]:
    return {k: v.to(dtype) for k, v in p.items()}


def rqs_forward(u: torch.Tensor, q: torch.Tensor):
    """One unconstrained rational-quadratic spline with linear tails on scalar inputs.
    u (R,), q (R, 3K-1) raw conditioner outputs -> (u_out (R,), logabsdet (R,))."""
    K = NUM_BINS
    uw = q[:, :K] / math.sqrt(HIDDEN)
    uh = q[:, K:2 * K] / math.sqrt(HIDDEN)
    ud = q[:, 2 * K:]
    const = math.log(math.exp(1.0 - MIN_DERIV) - 1.0)
    ud = F.pad(ud, (1, 1), value=const)

    inside = (u >= -TAIL_BOUND) & (u <= TAIL_BOUND)
    ui = torch.where(inside, u, torch.zeros_like(u))

    def knots(unnorm):
        w = MIN_BIN + (1.0 - MIN_BIN * K) * torch.softmax(unnorm, dim=-1)
        cw = torch.cumsum(w, dim=-1)
        cw = F.pad(cw, (1, 0), value=0.0)
        cw = 2.0 * TAIL_BOUND * cw - TAIL_BOUND
        cw[:, 0] = -TAIL_BOUND
        cw[:, -1] = TAIL_BOUND
        return cw, cw[:, 1:] - cw[:, :-1]

    cw, w = knots(uw)
    ch, h = knots(uh)
    d = MIN_DERIV + F.softplus(ud)

    edges = cw.clone()
    edges[:, -1] += 1e-6
    b = (ui[:, None] >= edges).sum(dim=-1) - 1
    b = b.clamp(0, K - 1)[:, None]
    in_cw, in_w = cw.gather(1, b)[:, 0], w.gather(1, b)[:, 0]
    in_ch, in_h = ch.gather(1, b)[:, 0], h.gather(1, b)[:, 0]
    delta = in_h / in_w
    d0, d1 = d.gather(1, b)[:, 0], d.gather(1, b + 1)[:, 0]

    th = (ui - in_cw) / in_w
    t1 = th * (1.0 - th)
    den = delta + (d0 + d1 - 2.0 * delta) * t1
    out = in_ch + in_h * (delta * th * th + d0 * t1) / den
    dnum = delta * delta * (d1 * th * th + 2.0 * delta * t1 + d0 * (1.0 - th) * (1.0 - th))
    lad = torch.log(dnum) - 2.0 * torch.log(den)
    return torch.where(inside, out, u), torch.where(inside, lad, torch.zeros_like(lad))


def log_prob(p: Dict[str, torch.Tensor], x: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """x (R,2) = [rt seconds, choice], cond (R,85) -> log p(rt, choice | cond), shape (R,).
    Computed in the dtype of the parameters (float32 spec, float64 for error budgets)."""
    dt = p["cat.W0"].dtype
    x, cond = x.to(dt), cond.to(dt)
    zt = (cond - p["cond_mean"]) / p["cond_std"]
    choice = x[:, 1]
    # categorical head
    h = torch.sigmoid(zt @ p["cat.W0"].T + p["cat.b0"])
    h = torch.sigmoid(h @ p["cat.W1"].T + p["cat.b1"])
    h = torch.sigmoid(h @ p["cat.W2"].T + p["cat.b2"])
    probs = torch.softmax(h @ p["cat.Wo"].T + p["cat.bo"], dim=-1)
    pc = probs.gather(1, choice.to(torch.int64)[:, None])[:, 0]
    lp_choice = torch.log(pc.clamp(PROB_EPS, 1.0 - PROB_EPS))
    # flow on log rt, conditioned on [z-scored condition, choice]
    ctx = torch.cat([zt, choice[:, None]], dim=1)
    y = torch.log(x[:, 0])
    u = (y - p["flow.mu_y"]) / p["flow.sigma_y"]
    lad = -torch.log(p["flow.sigma_y"]) * torch.ones_like(u)
    for k in range(NUM_TRANSFORMS):
        a = torch.relu(ctx @ p[f"flow.{k}.W1"].T + p[f"flow.{k}.b1"])
        a = torch.relu(a @ p[f"flow.{k}.W2"].T + p[f"flow.{k}.b2"])
        q = a @ p[f"flow.{k}.W3"].T + p[f"flow.{k}.b3"]
        u, l = rqs_forward(u, q)
        lad = lad + l
    base = -0.5 * u * u - 0.5 * math.log(2.0 * math.pi)
    return lp_choice + base + lad - y


def potential_rows(theta: torch.Tensor, x_o: torch.Tensor, pulses: torch.Tensor):
    """Row expansion of the reference's ConditionedMNLELogLikelihood.forward
    (potentials.py:96-110): row r = t*C + c holds [theta_c, pulses_t] and x_t."""
    C, T = theta.shape[0], x_o.shape[0]
    cond = torch.cat([theta.repeat(T, 1), pulses.repeat_interleave(C, dim=0)], dim=-1)
    return x_o.repeat_interleave(C, dim=0), cond


def loglik_sum(p: Dict[str, torch.Tensor], theta: torch.Tensor, x_o: torch.Tensor, pulses: torch.Tensor):
    """sum_t log p(x_t | theta_c, pulses_t) -> (C,)  (potentials.py:113-115)."""
    C, T = theta.shape[0], x_o.shape[0]
    xr, cond = potential_rows(theta, x_o, pulses)
    return log_prob(p, xr, cond).reshape(T, C).sum(0)


class SpecEstimator(torch.nn.Module):
    """Object with the estimator call shape the reference uses (potentials.py:113):
    ``log_prob(x (1,R,2), condition=(R,85)) -> (1,R)``."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        super().__init__()
        self.params = params

    def log_prob(self, x, condition):
        return log_prob(self.params, x.reshape(-1, 2), condition).unsqueeze(0)
