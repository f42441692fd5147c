"""Device time of one batched potential call at the SBC shape (D datasets x 128 chains x T=50)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sbi_for_diffusion_models_b200.mnle import _BatchedPotential
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
from sbi_for_diffusion_models_b200.priors import build_prior_theta
from sbi_for_diffusion_models_b200.samplers import GraphedLogProb
from sbi_for_diffusion_models_b200.sbc import draw_sbc_datasets, simulate_sbc_sessions

torch.cuda.set_device(0)
est = DeviceMNLE(PackedMNLE.from_params(bench.random_mnle_params(0)))
prior = build_prior_theta()
for D in (1, 8, 125):
    C, T = 128, 50
    thetas, seeds = draw_sbc_datasets(prior, D, seed=1)
    x, pulses = simulate_sbc_sessions(thetas, seeds, T, mu_sensory=1.0, p_success=0.75, noise_seed=2)
    th = prior.sample((D * C,)).cuda()
    pot = _BatchedPotential(est, prior, x, pulses, 1.0)
    gp = GraphedLogProb(pot, th)
    def timeit(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps
    k = timeit(lambda: est.loglik_sum_batched(th.view(D, C, 5), x, pulses))
    p = timeit(lambda: pot(th))
    g = timeit(lambda: gp(th))
    pr = timeit(lambda: prior.log_prob(th))
    print(f"D={D}: rows {D*C*T}: loglik_sum_batched {k[0]:.3f} ms | potential (prior + kernel) {p[0]:.3f} ms | graphed {g[0]:.3f} ms "
          f"(graph active: {gp.graph is not None}) | prior alone {pr[0]:.3f} ms")
