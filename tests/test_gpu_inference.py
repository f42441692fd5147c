"""Inference drivers on the GPU: batched potential = per-dataset potentials, run_inference_mcmc /
run_sbc keep the reference's call shapes (mnle.py:52-95, 128-237) and produce sane posteriors."""
import numpy as np
import pytest
import torch

from oracle import ddm_oracle as orc
from oracle import mnle_spec as ms
from sbi_for_diffusion_models_b200.mnle import run_inference_mcmc, run_sbc, sbc_shard
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
from sbi_for_diffusion_models_b200.priors import build_prior_theta
from sbi_for_diffusion_models_b200.run_config import RunConfig
from sbi_for_diffusion_models_b200.sbc import draw_sbc_datasets, simulate_sbc_sessions

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def est():
    return DeviceMNLE(PackedMNLE.from_params(ms.init_params(0)))


@pytest.mark.parametrize("D,T,C", [(3, 50, 128), (5, 7, 130), (2, 1, 1), (4, 50, 37)])
def test_batched_potential_equals_per_dataset_calls(est, D, T, C):
    prior = build_prior_theta()
    thetas, seeds = draw_sbc_datasets(prior, D, seed=1)
    x, pulses = simulate_sbc_sessions(thetas, seeds, T, mu_sensory=1.0, p_success=0.75, noise_seed=2)
    torch.manual_seed(0)
    th = prior.sample((D * C,)).view(D, C, 5).cuda()
    got = est.loglik_sum_batched(th, x, pulses)
    assert tuple(got.shape) == (D, C) and got.is_cuda
    for d in range(D):
        assert torch.equal(got[d], est.loglik_sum(th[d], x[d], pulses[d], kernel="tc"))      # same kernel, same bits
    assert torch.equal(got, est.loglik_sum_batched(th, x, pulses))                              # reproducible
    with pytest.raises(ValueError, match="theta must be"):
        est.loglik_sum_batched(th[0], x, pulses)
    with pytest.raises(ValueError, match="pulses must be"):
        est.loglik_sum_batched(th, x, pulses[:, :-1] if T > 1 else pulses[:1])


def test_run_inference_mcmc_shapes_and_posterior_concentrates(est):
    cfg = RunConfig(POSTERIOR_SAMPLES=600, WARMUP_STEPS=15, NUM_CHAINS=2)
    prior = build_prior_theta()
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(123)), 0, 50, 80, 0.75))
    x, _ = orc.sim_rng_c(np.repeat(np.array([[0.45, 0.6, 1.3, 14.0, 0.25]], np.float32), 50, 0), pulses.numpy(), 7)
    x = torch.from_numpy(x)
    torch.manual_seed(0)
    g = torch.Generator(device="cuda").manual_seed(0)
    samples = run_inference_mcmc(cfg, prior, est, x, pulses, generator=g)
    assert tuple(samples.shape) == (600, 5) and samples.device.type == "cpu" and samples.dtype == torch.float32
    assert bool(torch.isfinite(prior.log_prob(samples)).all())         # every draw inside the support
    # the draws follow the posterior of THIS (random-init) estimator: higher potential than prior draws
    from sbi_for_diffusion_models_b200.potentials import ConditionedMNLELogLikelihood, ThetaOnlyPosteriorPotential
    pot = ThetaOnlyPosteriorPotential(conditioned_loglike=ConditionedMNLELogLikelihood(est, pulses, "cpu"),
                                      prior_theta=prior, x_o=x, device="cpu", temperature=1.0)
    torch.manual_seed(1)
    assert float(pot(samples, track_gradients=False).median()) > float(pot(prior.sample((600,)), track_gradients=False).median()) + 1.0


def test_run_sbc_reference_outputs(est, tmp_path):
    cfg = RunConfig(WARMUP_STEPS=5, NUM_TRIALS_OBS=20)
    prior = build_prior_theta()
    out = run_sbc(cfg, prior_theta=prior, density_estimator=est, device="cpu", num_datasets=4,
                  posterior_samples_per_dataset=150, seed=0, outdir=str(tmp_path / "sbc"))
    assert out["thetas_true"].shape == (4, 5) and out["thetas_true"].dtype == np.float32
    assert out["ranks"].shape == (4, 5) and out["ranks"].dtype == np.int64
    assert len(out["all_samples"]) == 4 and all(tuple(s.shape) == (150, 5) for s in out["all_samples"])
    assert (out["ranks"] >= 0).all() and (out["ranks"] <= 150).all()
    # datasets are the reference's: theta_true_i and ds_seed_i from seed 0 in its order (mnle.py:161-189)
    want_thetas, _ = draw_sbc_datasets(prior, 4, seed=0)
    assert np.array_equal(out["thetas_true"], want_thetas.numpy())
    for i in range(4):
        assert np.array_equal(out["ranks"][i], (out["all_samples"][i] < want_thetas[i][None, :]).sum(0).numpy())
    assert (tmp_path / "sbc" / "sbc_ranks.npy").exists() and (tmp_path / "sbc" / "sbc_thetas_true.npy").exists()


def test_sbc_result_does_not_depend_on_the_sharding(est):
    """Any split of the datasets into shards (GPUs) reproduces the unsplit run bit for bit: Philox trial
    offsets in the simulator, starting points drawn for all datasets, counter-based sampler uniforms,
    per-dataset slice widths, and a potential kernel whose rows do not depend on their neighbours."""
    cfg = RunConfig(WARMUP_STEPS=4, NUM_TRIALS_OBS=12)
    prior = build_prior_theta()
    N, C, S = 5, 128, 200
    thetas, seeds = draw_sbc_datasets(prior, N, seed=3)
    init_all = prior.sample((N * C,)).to(torch.float32)
    whole_r, whole_s = sbc_shard(cfg, prior, est, thetas, seeds, init_all, 0, N, S, 3)
    for lo, hi in ((0, 2), (2, 5), (4, 5), (3, 3)):
        r, s = sbc_shard(cfg, prior, est, thetas, seeds, init_all, lo, hi, S, 3)
        assert torch.equal(r, whole_r[lo:hi]) and torch.equal(s, whole_s[lo:hi]), (lo, hi)
    assert tuple(whole_s.shape) == (N, S, 5) and bool((whole_r >= 0).all()) and bool((whole_r <= S).all())


def test_fused_sampler_kernels_give_the_bits_of_the_torch_implementation():
    """csrc/mnle_sampler.cu (two kernels around the potential call of an iteration) against the elementwise torch
    implementation of the same per-chain state machine: same uniforms, same arithmetic in the same order -> same
    draws, bit for bit, with and without CUDA-graph replay; and with the MNLE potential both paths agree too."""
    from sbi_for_diffusion_models_b200.samplers import PhiloxUniforms, VectorizedSliceSampler
    dev = torch.device("cuda")
    torch.manual_seed(0)
    N, D = 1000, 3
    mean = torch.tensor([0.5, -1.0, 2.0], device=dev)
    prec = torch.linalg.inv(torch.tensor([[1.0, 0.6, 0.0], [0.6, 2.0, 0.3], [0.0, 0.3, 0.5]], device=dev))

    def logp(x):       # correlated Gaussian truncated to x_2 > 0: -inf outside, like a prior's support
        d = x - mean
        lp = -0.5 * torch.einsum("ni,ij,nj->n", d, prec, d)
        return torch.where(x[:, 2] > 0, lp, torch.full_like(lp, -float("inf")))

    init = torch.randn((N, D), device=dev).abs() + 0.1
    draws = {}
    for fused, graph in ((False, False), (True, False), (True, True), (False, True)):
        s = VectorizedSliceSampler(logp, init, uniforms=PhiloxUniforms(77, 0, N, dev), fused=fused, use_graph=graph)
        draws[(fused, graph)] = s.run(7, warmup=9, thin=2)
        assert tuple(draws[(fused, graph)].shape) == (7, N, D) and s.n_evals > 0
    base = draws[(False, False)]
    for key, got in draws.items():
        assert torch.equal(got, base), key
    assert float(base[..., 2].min()) > 0
    got_mean = base.reshape(-1, D).mean(0)
    assert bool(((got_mean - mean).abs() < torch.tensor([0.25, 0.25, 0.6], device=dev)).all())   # (short run: coarse)
