"""MNLE training step through the C ABI against float64 autograd of the CPU
specification (oracle/mnle_spec.py; parity unpinned against sbi itself, see that module).

Tolerances.  forward="fp32" (CUDA-core forward, fp32 reverse mode, bf16 hi/lo tensor-core weight
gradients): each gradient tensor within 2e-3 of its own largest entry of float64 autograd, loss within
1e-5 relative.  forward="tc" (default: tcgen05 forward, operands carry ~17 bits): activations sit
within ~1e-5 of fp32, so ReLU units that close to their kink get the other mask -- about a hundred
(row, unit) pairs in these batches, each moving single gradient entries by one row's contribution:
1e-2 of the largest entry, loss within 1e-4 relative (5e-2 / 1e-3 on the sharpened stress net)."""
import numpy as np
import pytest
import torch

from oracle import ddm_oracle as orc
from oracle import mnle_spec as ms
from sbi_for_diffusion_models_b200.mnle_train import MNLETrainer, init_raw_params, train_mnle

pytestmark = pytest.mark.gpu


def _data(R, seed=0):
    theta = orc.prior_sample(R, seed=seed + 2)
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(seed + 1)), 0, R, 80, 0.75))
    cond = torch.cat([theta, pulses], dim=1)
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(np.stack([np.exp(rs.uniform(-3, 2.1, R)), rs.randint(0, 3, R)], 1).astype(np.float32))
    x[:5, 0] = torch.tensor([1e-6, 8.0, 7.999999, 1e-3, 3e-5])      # linear tails of the spline
    return x, cond


def _trainer(p, **kw):
    return MNLETrainer(int(p["cat.Wo"].shape[0]), cond_mean=p["cond_mean"], cond_std=p["cond_std"],
                       mu_y=float(p["flow.mu_y"]), sigma_y=float(p["flow.sigma_y"]), init=p, **kw)


def _relu_margin(p, x, cond):
    """Smallest |pre-activation| over the ReLU layers (float64).  A unit whose pre-activation is
    within fp32 rounding (~1e-7) of zero is switched on in one precision and off in the other, which
    moves single gradient entries by a whole row's contribution: an artefact of comparing across the
    kink, not an error.  The seeds below keep clear of it; this asserts that they do."""
    pd = ms.cast_params(p, torch.float64)
    ctx = torch.cat([(cond.double() - pd["cond_mean"]) / pd["cond_std"], x[:, 1:2].double()], 1)
    m = float("inf")
    for k in range(ms.NUM_TRANSFORMS):
        pre1 = ctx @ pd[f"flow.{k}.W1"].T + pd[f"flow.{k}.b1"]
        pre2 = torch.relu(pre1) @ pd[f"flow.{k}.W2"].T + pd[f"flow.{k}.b2"]
        m = min(m, float(pre1.abs().min()), float(pre2.abs().min()))
    return m


def _spec_loss_and_grads(p, x, cond):
    p64 = {k: v.double().clone().requires_grad_(k not in ("cond_mean", "cond_std", "flow.mu_y", "flow.sigma_y"))
           for k, v in p.items()}
    loss = -ms.log_prob(p64, x, cond).mean()
    loss.backward()
    return float(loss), {k: v.grad for k, v in p64.items() if v.grad is not None}


@pytest.mark.parametrize("forward", ["fp32", "tc"])
@pytest.mark.parametrize("seed,scale,R", [(1, 1.0, 300), (1, 2.0, 1000), (2, 1.0, 64 * 130 + 5)],
                         ids=["ragged", "sharp", "many-chunks"])
def test_loss_and_gradient_match_float64_autograd(seed, scale, R, forward):
    p = ms.init_params(seed, scale=scale)
    x, cond = _data(R, seed)
    if R < 2000:
        assert _relu_margin(p, x, cond) > 5e-7
    want_loss, want = _spec_loss_and_grads(p, x, cond)
    tr = _trainer(p)
    xd, cd = x.cuda(), tr.standardise(cond)
    stats = tr.nll(xd, cd, forward=forward).cpu()
    loss_tol = {("fp32", 1.0): 1e-5, ("fp32", 2.0): 1e-5, ("tc", 1.0): 1e-4, ("tc", 2.0): 1e-3}[(forward, scale)]
    grad_tol = {("fp32", 1.0): 2e-3, ("fp32", 2.0): 2e-2, ("tc", 1.0): 1e-2, ("tc", 2.0): 5e-2}[(forward, scale)]
    assert abs(float(stats[0]) - want_loss) <= loss_tol * abs(want_loss) + 1e-6
    got = tr.named_grads()
    assert set(want) == set(got) - {"flow.mu_y", "flow.sigma_y"}
    for name, g64 in want.items():
        err = float((got[name].double() - g64).abs().max())
        ref = float(g64.abs().max())
        # the sharpened net (weights x 2) is the numerics stress of the forward tests too: rows next to a
        # bin edge or a ReLU kink land on the other side in fp32
        assert err <= grad_tol * ref + 1e-9, (name, err, ref)
    assert float(got["flow.mu_y"]) == 0.0 and float(got["flow.sigma_y"]) == 0.0   # buffers, not trained
    ss = sum(float((g.double() ** 2).sum()) for g in want.values())
    assert abs(float(stats[1]) - ss) <= (1e-3 if (scale == 1.0 and forward == "fp32") else 2e-2) * ss
    # bit-reproducible (fixed-order reductions, no atomics)
    g1 = tr.grad.clone()
    tr.nll(xd, cd, forward=forward)
    assert torch.equal(g1, tr.grad)
    # loss-only pass leaves the gradient alone and agrees on the loss
    tr.grad.zero_()
    s2 = tr.nll(xd, cd, grad=False, forward=forward).cpu()
    assert float(s2[0]) == float(stats[0]) and float(s2[1]) == 0.0 and float(tr.grad.abs().max()) == 0.0


def test_row_index_gathers_the_minibatch():
    p = ms.init_params(3)
    x, cond = _data(700, 3)
    tr = _trainer(p)
    xd, cd = x.cuda(), tr.standardise(cond)
    idx = torch.randperm(700, generator=torch.Generator().manual_seed(0))[:257].cuda()
    s_idx = tr.nll(xd, cd, idx).clone()
    g_idx = tr.grad.clone()
    s_gat = tr.nll(xd[idx].contiguous(), cd[idx].contiguous()).clone()
    assert torch.equal(s_idx, s_gat) and torch.equal(g_idx, tr.grad)
    wide = torch.cat([cd, torch.zeros(700, 3, device="cuda")], dim=1)       # strided condition rows
    tr.nll(xd, wide[:, :85], idx)
    assert torch.equal(g_idx, tr.grad)
    with pytest.raises(ValueError):
        tr.nll(xd, cd, idx.int())
    with pytest.raises(ValueError):
        tr.nll(xd, cd[:, :80])


def test_adam_update_matches_torch_optim():
    p = ms.init_params(4)
    x, cond = _data(512, 4)
    tr = _trainer(p)
    xd, cd = x.cuda(), tr.standardise(cond)
    n = tr.params.numel() - 2
    ref = torch.nn.Parameter(tr.params[:n].detach().cpu().clone())
    opt = torch.optim.Adam([ref], lr=5e-4)
    for step in range(4):
        tr.nll(xd, cd)
        ref.grad = tr.grad[:n].detach().cpu().clone()
        clip = 5.0 if step % 2 == 0 else 1e-3       # the second value actually clips
        torch.nn.utils.clip_grad_norm_([ref], clip)
        opt.step()
        tr.adam(lr=5e-4, clip_max_norm=clip)
        diff = float((tr.params[:n].cpu() - ref.detach()).abs().max())
        assert diff < 2e-7, (step, diff)
    assert float(tr.params[n]) == pytest.approx(float(p["flow.mu_y"])) and float(tr.params[n + 1]) == pytest.approx(
        float(p["flow.sigma_y"]))


def test_training_trajectory_follows_the_spec_under_torch_autograd():
    """Five full-batch Adam steps on the device vs the same steps through the float64 spec."""
    p = ms.init_params(1)
    x, cond = _data(640, 1)
    assert _relu_margin(p, x, cond) > 5e-7
    tr = _trainer(p)
    xd, cd = x.cuda(), tr.standardise(cond)
    frozen = ("cond_mean", "cond_std", "flow.mu_y", "flow.sigma_y")
    p64 = {k: v.double().clone().requires_grad_(k not in frozen) for k, v in p.items()}
    opt = torch.optim.Adam([v for k, v in p64.items() if k not in frozen], lr=5e-4)
    for step in range(5):
        loss = -ms.log_prob(p64, x, cond).mean()
        got = float(tr.nll(xd, cd, forward="fp32")[0])
        # (units crossing a ReLU kink between the two precisions make the trajectories drift apart slowly)
        assert abs(got - float(loss)) <= (2e-5 if step < 2 else 5e-4) * abs(float(loss)), (step, got, float(loss))
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([v for k, v in p64.items() if k not in frozen], 5.0)
        opt.step()
        tr.adam(lr=5e-4, clip_max_norm=5.0)
    # Adam's first steps move every parameter by ~lr * sign(g): entries whose gradient is at the fp32
    # noise floor may step the other way, so the parameters are compared in bulk, the loss strictly
    raw = tr.raw_params()
    diffs = torch.cat([(raw[k].double() - v.detach()).abs().reshape(-1) for k, v in p64.items() if k not in frozen])
    assert float(diffs.quantile(0.9)) < 2e-5 and float(diffs.max()) <= 2 * 5 * 5e-4


def test_train_mnle_end_to_end_on_simulated_trials():
    """Reference flow (rt_choice_model_pipeline.py:66-82): simulate a training set, train, use the estimator."""
    from sbi_for_diffusion_models_b200 import simulator as sim
    from sbi_for_diffusion_models_b200.pulses import generate_pulse_matrix_device
    from sbi_for_diffusion_models_b200.run_config import RunConfig

    N = 6000
    theta = orc.prior_sample(N, seed=11)
    pulses = generate_pulse_matrix_device(np.random.default_rng(0), N, 80, p_success=0.75)
    x = sim.simulate_trials(theta, pulses, seed=5).cpu()
    z = torch.cat([theta, pulses.cpu()], dim=1)
    cfg = RunConfig(TRAIN_BATCH_SIZE=1024)
    est, info = train_mnle(cfg, None, z, x, max_num_epochs=6, return_summary=True)
    hist = info["history"]
    assert info["n_choices"] == 3 and info["epochs"] == 6 and info["steps"] == 6 * (5400 // 1024)
    assert hist[-1][0] < hist[0][0] - 0.05 and hist[-1][1] < hist[0][1]          # learning
    assert all(np.isfinite(h).all() for h in hist)
    lp = est.log_prob(x[:500].unsqueeze(0), condition=z[:500])
    assert tuple(lp.shape) == (1, 500) and bool(torch.isfinite(lp).all())
    # same seed, same result: the whole loop is deterministic
    est2 = train_mnle(cfg, None, z, x, max_num_epochs=2, seed=0)
    est3 = train_mnle(cfg, None, z, x, max_num_epochs=2, seed=0)
    assert np.array_equal(est2.packed.packed, est3.packed.packed)
    with pytest.raises(ValueError):
        train_mnle(cfg, None, z[:, :80], x)


def test_init_follows_torch_linear_defaults():
    p = init_raw_params(3, seed=0)
    assert tuple(p["cat.W0"].shape) == (128, 85) and tuple(p["flow.9.W3"].shape) == (71, 128)
    assert float(p["flow.0.W1"].abs().max()) <= 1 / np.sqrt(86) and float(p["flow.0.b1"].abs().max()) <= 1 / np.sqrt(86)
    assert float(p["cat.Wo"].abs().max()) > 0.5 / np.sqrt(128)


@pytest.mark.parametrize("forward", ["fp32", "tc"])
@pytest.mark.parametrize("R", [1, 127, 300, 64 * 130 + 5])
def test_workspace_and_gradient_writes_stay_in_bounds(R, forward):
    """Guard bands around every buffer the step writes (workspace of exactly the advertised size,
    gradient, stats): ragged last tiles must not spill (compute-sanitizer is not available on the
    GPU pool, so the bounds are checked this way)."""
    from sbi_for_diffusion_models_b200 import _native
    p = ms.init_params(5)
    x, cond = _data(max(R, 8), 5)
    tr = _trainer(p)
    xd, cd = x.cuda()[:R].contiguous(), tr.standardise(cond)[:R].contiguous()
    G = 4096
    need = _native.lib().mnle_train_workspace_floats(tr.n_choices, R)
    big_ws = torch.full((need + 2 * G,), 12345.0, device="cuda")
    tr._ws = big_ws[G:G + need]
    n = tr.grad.numel()
    big_g = torch.full((n + 2 * G,), 12345.0, device="cuda")
    tr.grad = big_g[G:G + n]
    big_s = torch.full((2 + 2 * G,), 12345.0, device="cuda")
    tr.stats = big_s[G:G + 2]
    stats = tr.nll(xd, cd, forward=forward)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(stats).all()) and bool(torch.isfinite(tr.grad).all())
    for big, m in ((big_ws, need), (big_g, n), (big_s, 2)):
        assert bool((big[:G] == 12345.0).all()) and bool((big[G + m:] == 12345.0).all())
    tr.adam()
    torch.cuda.synchronize()
    assert bool((big_g[:G] == 12345.0).all()) and bool((big_g[G + n:] == 12345.0).all())


@pytest.mark.parametrize("K", [2, 8])
def test_training_step_with_other_numbers_of_choices(K):
    """K = 2: a training set without censored trials (sbi sizes the categorical head from the data)."""
    p = ms.init_params(20 + K, n_choices=K)
    R = 333
    x, cond = _data(R, {2: 8, 8: 9}[K])
    x[:, 1] = torch.from_numpy(np.random.RandomState(K).randint(0, K, R).astype(np.float32))
    assert _relu_margin(p, x, cond) > 5e-7       # (data seeds chosen to keep clear of the ReLU kinks)
    want_loss, want = _spec_loss_and_grads(p, x, cond)
    tr = _trainer(p)
    assert tr.n_choices == K
    xd, cd = x.cuda(), tr.standardise(cond)
    # (tc: with only 333 rows one ReLU unit masked differently moves an entry by 1/333 of a row's gradient,
    # see the module docstring; the categorical tensors, which are what K changes, get the tight bound)
    for forward, loss_tol, grad_tol in (("fp32", 1e-5, 2e-3), ("tc", 1e-4, 2e-2)):
        stats = tr.nll(xd, cd, forward=forward).cpu()
        assert abs(float(stats[0]) - want_loss) <= loss_tol * abs(want_loss) + 1e-6, forward
        got = tr.named_grads()
        for name, g64 in want.items():
            err, ref = float((got[name].double() - g64).abs().max()), float(g64.abs().max())
            tol = 2e-3 if name.startswith("cat.") else grad_tol      # sigmoid nets have no kinks
            assert err <= tol * ref + 1e-9, (forward, name, err, ref)
    tr.adam()
    assert bool(torch.isfinite(tr.params).all())
