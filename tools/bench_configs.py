"""Measured numbers for the BASELINE configs that are not bench.py's headline line (one GPU):
configs[2] shape (long pulse schedule dt = 1e-4: n_max = 80000, steps_per_pulse = 1000) and
configs[4] shape (run_sbc: simulate + sample + rank per dataset, batched over datasets).
Prints one JSON object; `python tools/bench_configs.py [--datasets D] [--long-trials N]`."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from sbi_for_diffusion_models_b200.mnle import run_sbc  # noqa: E402
from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE  # noqa: E402
from sbi_for_diffusion_models_b200.priors import build_prior_theta  # noqa: E402
from sbi_for_diffusion_models_b200.run_config import RunConfig  # noqa: E402
from sbi_for_diffusion_models_b200.simulator import Schedule, simulate_trials  # noqa: E402


def long_schedule(n):
    dev = torch.device("cuda", 0)
    sched = Schedule.from_constants(dt=1e-4)
    z = bench.build_workload(n, 0, dev)
    x = torch.empty((n, 2), dtype=torch.float32, device=dev)
    for i in range(2):
        simulate_trials(z[:, :5], z[:, 5:], seed=i, out=x, schedule=sched)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for i in range(reps):
        simulate_trials(z[:, :5], z[:, 5:], seed=10 + i, out=x, schedule=sched)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    _, st = simulate_trials(z[:, :5], z[:, 5:], seed=10, out=x, schedule=sched, return_stats=True)
    return {"workload": f"configs[2] shape on one GPU: {n} trials, dt=1e-4 (n_max={sched.n_max}, "
                        f"steps_per_pulse={sched.steps_per_pulse}, P={sched.n_pulses})",
            "ms_per_launch": ms, "useful_steps_per_s": st.useful_steps / (ms * 1e-3), "trials_per_s": n / (ms * 1e-3),
            "mean_steps_per_trial": st.useful_steps / n, "lane_efficiency": st.lane_efficiency}


def sbc(datasets, samples, warmup):
    est = DeviceMNLE(PackedMNLE.from_params(bench.random_mnle_params(0)))
    cfg = RunConfig(WARMUP_STEPS=warmup)
    prior = build_prior_theta()
    run_sbc(cfg, prior_theta=prior, density_estimator=est, num_datasets=2, posterior_samples_per_dataset=128, save=False)
    torch.cuda.synchronize()
    import sbi_for_diffusion_models_b200.samplers as smp
    evals = {"n": 0}
    orig = smp.VectorizedSliceSampler.run

    def counting(self, *a, **kw):
        out = orig(self, *a, **kw)
        evals["n"] += self.n_evals
        return out

    smp.VectorizedSliceSampler.run = counting
    t0 = time.perf_counter()
    out = run_sbc(cfg, prior_theta=prior, density_estimator=est, num_datasets=datasets,
                  posterior_samples_per_dataset=samples, save=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    smp.VectorizedSliceSampler.run = orig
    rows = datasets * 128 * cfg.NUM_TRIALS_OBS
    return {"workload": f"configs[4] shape on one GPU: run_sbc over {datasets} datasets x 128 chains, T={cfg.NUM_TRIALS_OBS}, "
                        f"{warmup} warm-up sweeps + {-(-samples // 128)} draws per chain, random-init MNLE",
            "seconds": dt, "datasets_per_s": datasets / dt, "potential_calls": evals["n"],
            "ms_per_potential_call": dt / evals["n"] * 1e3, "rows_per_call": rows,
            "rows_per_s": rows * evals["n"] / dt, "rank_mean": out["ranks"].mean(0).tolist()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--datasets", type=int, default=125)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--sbc-warmup", type=int, default=20)
    ap.add_argument("--long-trials", type=int, default=2_000_000)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    print(json.dumps({"long_schedule": long_schedule(a.long_trials), "sbc": sbc(a.datasets, a.samples, a.sbc_warmup)}, indent=1))
