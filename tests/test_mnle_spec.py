"""CPU checks of the MNLE specification (oracle/mnle_spec.py) and of the host-side packer.
Parity with real sbi is UNPINNED (sbi is not installable here): these tests pin the spec to the
published spline / MNLE definitions and to the reference's row layout, nothing more."""
import math

import numpy as np
import pytest
import torch

from oracle import ddm_oracle as orc
from oracle import mnle_spec as ms
from sbi_for_diffusion_models_b200.mnle_net import PackedMNLE


def test_spline_is_monotone_identity_in_tails_and_logdet_is_log_derivative():
    g = torch.Generator().manual_seed(0)
    u = torch.cat([torch.linspace(-12, 12, 193, dtype=torch.float64), torch.tensor([-10.0, 10.0, 0.0], dtype=torch.float64)])
    u.requires_grad_(True)
    q = torch.randn(u.shape[0], ms.SPLINE_OUT, dtype=torch.float64, generator=g) * 4
    out, lad = ms.rqs_forward(u, q)
    (grad,) = torch.autograd.grad(out.sum(), u)
    assert bool((grad > 0).all())
    assert float((grad.log() - lad.detach()).abs().max()) < 1e-10
    outside = (u.detach().abs() > 10)
    assert torch.equal(out[outside], u.detach()[outside]) and float(lad[outside].abs().max()) == 0.0
    # the spline maps [-10, 10] onto itself with unit slope at both ends (linear tails are C1)
    ends = torch.tensor([-10.0, 10.0], dtype=torch.float64, requires_grad=True)
    o, l = ms.rqs_forward(ends, q[:2])
    assert torch.allclose(o, ends.detach(), atol=1e-9) and float(l.abs().max()) < 1e-9


def test_density_normalises():
    """Integral over rt of sum_choice p(rt, choice | z) = 1 (trapezoid in log rt, float64)."""
    p = ms.cast_params(ms.init_params(3, scale=1.5), torch.float64)
    cond = torch.cat([orc.prior_sample(1, seed=1)[0].double(), torch.ones(80, dtype=torch.float64)])[None, :]
    y = torch.linspace(-14, 14, 40001, dtype=torch.float64)
    total = 0.0
    for c in range(3):
        x = torch.stack([torch.exp(y), torch.full_like(y, float(c))], dim=1)
        lp = ms.log_prob(p, x, cond.expand(y.shape[0], -1))
        total += torch.trapezoid(torch.exp(lp + y), y).item()     # p(rt) d rt = p(rt) rt d log rt
    assert abs(total - 1.0) < 5e-5   # trapezoid error on a 7e-4 grid


def test_fp32_spec_tracks_fp64_spec():
    p = ms.init_params(0)
    C, T = 64, 50
    theta = orc.prior_sample(C, seed=3)
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(123)), 0, T, 80, 0.75))
    x, _ = orc.sim_rng_c(np.repeat(np.array([[0.45, 0.6, 1.3, 14.0, 0.25]], np.float32), T, 0), pulses.numpy(), 7)
    x = torch.from_numpy(x)
    a = ms.loglik_sum(p, theta, x, pulses).double()
    b = ms.loglik_sum(ms.cast_params(p, torch.float64), theta, x, pulses)
    assert float(((a - b).abs() / b.abs()).max()) < 2e-5


def test_row_layout_matches_reference_fixture(golden):
    """potentials.py:96-115: row r = t*C + c, reshape (T, C), sum over t -- checked with the stub
    estimator the fixture was generated with."""
    g = golden("potential_layout")
    pul, x_obs, thetas = (torch.from_numpy(g[k].astype(np.float32)) for k in ("pulses", "x_obs", "thetas"))
    xr, cond = ms.potential_rows(thetas, x_obs, pul)
    w = torch.arange(1, 86, dtype=torch.float32) * 1e-2
    stub = (cond * w).sum(-1) + 3.0 * xr[:, 0] - 0.5 * xr[:, 1]
    ll = stub.reshape(x_obs.shape[0], thetas.shape[0]).sum(0)
    assert torch.allclose(ll, torch.from_numpy(g["ll"]), rtol=1e-6, atol=1e-5)


def test_packer_folds_zscoring_exactly():
    p = ms.cast_params(ms.init_params(5), torch.float64)
    pk = PackedMNLE.from_params(p)
    assert pk.packed.dtype == np.float32 and pk.packed.size == PackedMNLE.packed_floats(3) == 412491
    W0 = pk.packed[:128 * 85].reshape(128, 85).astype(np.float64)
    b0 = pk.packed[128 * 85:128 * 85 + 128].astype(np.float64)
    z = np.random.RandomState(0).randn(85)
    zt = (z - p["cond_mean"].numpy()) / p["cond_std"].numpy()
    want = p["cat.W0"].numpy() @ zt + p["cat.b0"].numpy()
    assert np.allclose(W0 @ z + b0, want, rtol=0, atol=2e-6)
    assert pk.packed[-2] == np.float32(p["flow.mu_y"]) and pk.packed[-1] == np.float32(p["flow.sigma_y"])


def _fake_sbi_state_dict(p, y_as="affine", ctx86=None):
    """A state_dict laid out the way sbi 0.25.0 / nflows are believed to (UNVERIFIED, see from_state_dict):
    ``Standardize`` modules carry ``_mean`` / ``_std``, nflows ``AffineTransform`` carries ``_shift`` / ``_scale``
    with x -> x * scale + shift."""
    sd = {}
    sd["net.discrete_net.embedding._mean"] = p["cond_mean"]
    sd["net.discrete_net.embedding._std"] = p["cond_std"]
    for i, n in enumerate(("0", "1", "2", "o")):
        sd[f"net.discrete_net.layers.{i}.weight"], sd[f"net.discrete_net.layers.{i}.bias"] = p[f"cat.W{n}"], p[f"cat.b{n}"]
    if y_as == "affine":
        sd["net.continuous_net._transform.0._shift"] = (-p["flow.mu_y"] / p["flow.sigma_y"]).reshape(1)
        sd["net.continuous_net._transform.0._scale"] = (1.0 / p["flow.sigma_y"]).reshape(1)
    else:
        sd["net.continuous_net._transform.0._mean"] = p["flow.mu_y"].reshape(1)
        sd["net.continuous_net._transform.0._std"] = p["flow.sigma_y"].reshape(1)
    if ctx86 is not None:
        sd["net.continuous_net._embedding_net.0._mean"], sd["net.continuous_net._embedding_net.0._std"] = ctx86
    for k in range(10):
        base = f"net.continuous_net._transform.{k + 1}.conditioner"
        for i in range(3):
            sd[f"{base}.{2 * i}.weight"], sd[f"{base}.{2 * i}.bias"] = p[f"flow.{k}.W{i + 1}"], p[f"flow.{k}.b{i + 1}"]
    return sd


def test_state_dict_import_is_shape_driven_and_strict():
    p = ms.cast_params(ms.init_params(6), torch.float64)
    a = PackedMNLE.from_params(p)
    for y_as in ("affine", "standardize"):          # (x * scale + shift) or ((x - mean) / std) for log rt
        b = PackedMNLE.from_state_dict(_fake_sbi_state_dict(p, y_as))
        assert np.array_equal(a.packed, b.packed) and b.n_choices == 3
    broken = _fake_sbi_state_dict(p)
    del broken["net.continuous_net._transform.4.conditioner.0.weight"]
    with pytest.raises(ValueError, match="does not look like"):
        PackedMNLE.from_state_dict(broken)
    extra = _fake_sbi_state_dict(p)
    extra["head.weight"], extra["head.bias"] = torch.zeros(7, 128), torch.zeros(7)      # a layer the layout has no place for
    with pytest.raises(ValueError, match="does not look like"):
        PackedMNLE.from_state_dict(extra)
    no_std = {k: v for k, v in _fake_sbi_state_dict(p, "standardize").items() if "_std" not in k}
    with pytest.raises(ValueError, match="standardisation"):
        PackedMNLE.from_state_dict(no_std)
    twice = _fake_sbi_state_dict(p)
    twice["other._mean"], twice["other._std"] = p["cond_mean"], p["cond_std"]           # two 85-wide candidates: refuse
    with pytest.raises(ValueError, match="ambiguous"):
        PackedMNLE.from_state_dict(twice)
    with pytest.raises(ValueError):
        PackedMNLE(np.zeros(10, np.float32), 3)


def test_flow_side_context_standardisation_is_folded_into_the_conditioners():
    """A flow that z-scores all 86 context columns itself (choice included): first layers W (c - m) / s + b."""
    p = ms.cast_params(ms.init_params(8), torch.float64)
    g = torch.Generator().manual_seed(1)
    m86 = torch.randn(86, generator=g, dtype=torch.float64)
    s86 = torch.rand(86, generator=g, dtype=torch.float64) + 0.5
    pk = PackedMNLE.from_state_dict(_fake_sbi_state_dict(p, ctx86=(m86, s86)))
    from sbi_for_diffusion_models_b200.mnle_net import unpack_params
    u = unpack_params(torch.from_numpy(pk.packed.copy()), 3)
    ctx = torch.randn(86, generator=g, dtype=torch.float64)
    for k in (0, 9):
        want = p[f"flow.{k}.W1"] @ ((ctx - m86) / s86) + p[f"flow.{k}.b1"]
        got = u[f"flow.{k}.W1"].double() @ ctx + u[f"flow.{k}.b1"].double()
        assert float((got - want).abs().max()) < 5e-6
    z = ctx[:85]                                     # the categorical net keeps the 85-wide z-scoring
    want = p["cat.W0"] @ ((z - p["cond_mean"]) / p["cond_std"]) + p["cat.b0"]
    assert float((u["cat.W0"].double() @ z + u["cat.b0"].double() - want).abs().max()) < 5e-6


def test_estimator_state_round_trips_like_the_reference_checkpoints(tmp_path, monkeypatch):
    """reference mnle.py:241-297: torch.save({"state_dict": est.state_dict(), "config": cfg}) / load_state_dict."""
    import pickle
    from sbi_for_diffusion_models_b200 import mnle
    from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE
    from sbi_for_diffusion_models_b200.run_config import RUN_CONFIG_PARAMS as cfg
    est = DeviceMNLE(PackedMNLE.from_params(ms.init_params(2)))
    sd = est.state_dict()
    assert set(sd) == {"packed_params", "n_choices"} and sd["packed_params"].numel() == 412491
    other = DeviceMNLE(PackedMNLE.from_params(ms.init_params(3)))
    other.load_state_dict(sd)
    assert np.array_equal(other.packed.packed, est.packed.packed)
    two = DeviceMNLE(PackedMNLE.from_params(ms.init_params(4, n_choices=2)))     # a different head size loads too
    two.load_state_dict(sd)
    assert two.packed.n_choices == 3 and np.array_equal(two.packed.packed, est.packed.packed)
    clone = pickle.loads(pickle.dumps(est))
    assert np.array_equal(clone.packed.packed, est.packed.packed) and set(clone.state_dict()) == set(sd)
    monkeypatch.setenv("HOME", str(tmp_path))
    assert mnle.load_model(cfg) is None                                          # nothing saved yet (mnle.py:264-266)
    path = mnle.save_model(est, cfg, "net.pt")
    assert path == str(tmp_path / "models" / "net.pt")
    back = mnle.load_model(cfg, "net.pt")
    assert np.array_equal(back.packed.packed, est.packed.packed) and back.packed.n_choices == 3


def test_sbc_helpers_cpu():
    from sbi_for_diffusion_models_b200.sbc import compute_ranks, draw_sbc_datasets
    samples = torch.tensor([[0.1, 5.0], [0.2, 1.0], [0.3, 2.0], [0.4, 9.0]])
    assert compute_ranks(torch.tensor([0.25, 2.0]), samples).tolist() == [2, 1]     # strict <, per dimension
    assert compute_ranks(torch.tensor([[0.25, 2.0]]), samples).dtype == torch.int64

    class Prior:
        def sample(self, shape):
            return torch.rand((shape[0], 5))

    th, seeds = draw_sbc_datasets(Prior(), 6, seed=0)
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    for i in range(6):                              # reference order: prior draw, then ds_seed (mnle.py:185-188)
        assert torch.equal(th[i], torch.rand((1, 5)).view(5))
        assert seeds[i] == int(rng.integers(0, 2**31 - 1))


def test_spline_matches_the_nflows_derived_implementation_shipped_in_transformers():
    """An independent pin for the part of A8 where the numerics are: the unconstrained rational-quadratic spline with
    linear tails.  sbi 0.25.0 evaluates it through nflows 0.14 (``nflows.transforms.splines.rational_quadratic``),
    which is not in this image; Hugging Face ``transformers`` IS, and its VITS model carries the same routine (VITS
    took its ``transforms.py`` from nflows: softmax widths / heights with the 1e-3 floors, cumulative knots pinned at
    +-tail_bound, softplus derivatives with the boundary constant log(exp(1 - 1e-3) - 1), the 1e-6 nudge of the last
    edge in the bin search, identity outside the tails).  The spec's ``rqs_forward`` must agree with it to rounding,
    in float64 and in float32, on random parameters at the scale of a trained net and on the edge cases."""
    vits = pytest.importorskip("transformers.models.vits.modeling_vits")
    ref = vits._unconstrained_rational_quadratic_spline
    K = ms.NUM_BINS
    g = torch.Generator().manual_seed(5)
    for dtype, tol in ((torch.float64, 1e-13), (torch.float32, 1e-6)):   # measured: 4e-15 and 4e-6 (a few ulp at |u| ~ 10)
        for scale in (1.0, 30.0, 120.0):      # raw conditioner outputs; 120 / sqrt(128) ~ 10: sharp bins
            R = 4000
            q = (torch.randn(R, 3 * K - 1, generator=g) * scale).to(dtype)
            u = (torch.rand(R, generator=g) * 24.0 - 12.0).to(dtype)           # a sixth of them outside the tails
            u[:8] = torch.tensor([-ms.TAIL_BOUND, ms.TAIL_BOUND, 0.0, -12.0, 12.0, 9.999999, -9.999999, 1e-30], dtype=dtype)
            got_u, got_lad = ms.rqs_forward(u, q)
            want_u, want_lad = ref(u.clone(), q[:, :K] / math.sqrt(ms.HIDDEN), q[:, K:2 * K] / math.sqrt(ms.HIDDEN),
                                   q[:, 2 * K:].clone(), reverse=False, tail_bound=ms.TAIL_BOUND, min_bin_width=ms.MIN_BIN,
                                   min_bin_height=ms.MIN_BIN, min_derivative=ms.MIN_DERIV)
            assert torch.allclose(got_u, want_u, rtol=tol, atol=tol * 10), (dtype, scale, float((got_u - want_u).abs().max()))
            assert torch.allclose(got_lad, want_lad, rtol=tol, atol=tol * 10), (dtype, scale, float((got_lad - want_lad).abs().max()))
            outside = (u < -ms.TAIL_BOUND) | (u > ms.TAIL_BOUND)
            assert torch.equal(got_u[outside], u[outside]) and float(got_lad[outside].abs().max()) == 0.0
