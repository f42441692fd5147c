#!/usr/bin/env python
"""Train the MNLE the tests, the smoke and the bench evaluate -- with this repository's own path.

    python tools/train_reference_net.py [--trials 1000000] [--max-epochs 400] [--out tests/golden/mnle_trained.npz]

The reference only ever evaluates a TRAINED estimator (mnle.py:41-48 -> potentials.py:113).  This
script makes one the way the reference's pipeline does (rt_choice_model_pipeline.py:58-90):
``ExtendedProposal`` over the pipeline prior and the pulse proposal ->
``simulate_training_set_with_conditions`` -> ``train_mnle`` (batch 4096, 10 % validation, early stop
after 20 stale epochs), all on the GPU, and stores the packed fp32 parameters (1.65 MB, z-scoring
folded into the first layers) plus the training summary.  Then it reports how far the CUDA kernels
sit from the float64 CPU spec ON THAT NET at configs[3] (T = 50 trials x C = 1024 chains).

Needs a B200 (no CPU fallback).  The committed fixture was produced by exactly this command line;
seeds are fixed, so re-running it regenerates the same training set (the trained weights can differ
in the last bits only if the kernels' summation order changes).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=1_000_000)
    ap.add_argument("--max-epochs", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="tests/golden/mnle_trained.npz")
    args = ap.parse_args()

    from sbi_for_diffusion_models_b200.data_simulator import simulate_observed_session, simulate_training_set_with_conditions
    from sbi_for_diffusion_models_b200.mnle_train import train_mnle
    from sbi_for_diffusion_models_b200.priors import build_prior_theta
    from sbi_for_diffusion_models_b200.proposals import ExtendedProposal, PulseSequenceProposal
    from sbi_for_diffusion_models_b200.run_config import RUN_CONFIG_PARAMS as cfg

    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    P = 80
    prior = build_prior_theta()
    proposal = ExtendedProposal(prior, PulseSequenceProposal(P, cfg.P_SUCCESS, seed=args.seed, device="cuda"), device="cuda")
    t0 = time.time()
    z, x = simulate_training_set_with_conditions(proposal, args.trials, 1 << 18, "cuda", mu_sensory=cfg.MU_SENSORY,
                                                 p_success=cfg.P_SUCCESS, P=P, log_rt=cfg.LOG_RT_MANUALLY,
                                                 seed=20261018)
    t_sim = time.time() - t0
    t0 = time.time()
    est, summary = train_mnle(cfg, proposal, z, x, "cuda", seed=args.seed, max_num_epochs=args.max_epochs,
                              return_summary=True, show_train_summary=False)
    torch.cuda.synchronize()
    t_train = time.time() - t0
    hist = summary.pop("history")
    summary.update(trials=args.trials, simulate_s=round(t_sim, 2), train_s=round(t_train, 2),
                   first_epoch=hist[0], last_epoch=hist[-1], seed=args.seed)
    print("training summary:", json.dumps(summary))
    packed = est.packed.packed
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    np.savez(args.out, packed=packed, n_choices=np.int64(est.packed.n_choices), summary=json.dumps(summary))
    print("wrote", args.out, packed.size, "floats")
    report_accuracy(est, cfg, prior)


def report_accuracy(est, cfg, prior):
    """tcgen05 / fp32 kernels vs the float64 spec on the trained net (checker code: oracle/)."""
    from oracle import mnle_spec as ms
    from sbi_for_diffusion_models_b200.data_simulator import simulate_observed_session
    from sbi_for_diffusion_models_b200.mnle_net import unpack_params

    # the packed buffer already has the z-scoring folded into the first layers: identity z-scoring in the spec
    p32 = dict(unpack_params(torch.from_numpy(est.packed.packed.copy()), est.packed.n_choices))
    p32["cond_mean"], p32["cond_std"] = torch.zeros(85), torch.ones(85)
    p64 = ms.cast_params(p32, torch.float64)
    theta_true = torch.tensor([0.45, 0.6, 1.3, 14.0, 0.25])
    x_o, pulses = simulate_observed_session(theta_true, 50, "cuda", mu_sensory=cfg.MU_SENSORY, p_success=cfg.P_SUCCESS,
                                            P=80, seed=123, log_rt=False, noise_seed=7)
    torch.manual_seed(3)
    sets = {"prior chains": prior.sample((1024,)),
            "chains near theta_true": theta_true * (1 + 0.05 * torch.randn(1024, 5))}
    for name, theta in sets.items():
        want = ms.loglik_sum(p64, theta, x_o, pulses)
        xr, cond = ms.potential_rows(theta, x_o, pulses)
        want_rows = ms.log_prob(p64, xr, cond)
        for kernel in ("tc", "simt"):
            got = est.loglik_sum(theta, x_o, pulses, kernel=kernel).double().cpu()
            rel = ((got - want).abs() / want.abs())
            rows = est.log_prob(xr, condition=cond, kernel=kernel)[0].double().cpu()
            err = (rows - want_rows).abs()
            print(f"{name:24s} {kernel:5s} sums: max rel {float(rel.max()):.2e} mean rel {float(rel.mean()):.2e} "
                  f"| rows: max abs {float(err.max()):.2e} mean abs {float(err.mean()):.2e} "
                  f"| sum range [{float(want.min()):.1f}, {float(want.max()):.1f}]")


if __name__ == "__main__":
    main()
