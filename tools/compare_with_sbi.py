#!/usr/bin/env python
"""Pin the MNLE path against a real sbi estimator -- for someone who HAS sbi installed.

sbi 0.25.0 / nflows 0.14 / pyknos 0.16.0 are not in this repository's build environment (no
network), so the MNLE arithmetic here is restated from the published algorithms
(oracle/mnle_spec.py) and its parity with sbi is UNPINNED.  This script closes that gap on a
machine with sbi (and a B200 for the CUDA columns):

    pip install sbi==0.25.0
    python tools/compare_with_sbi.py [--no-cuda]

It builds the estimator exactly as the reference does (mnle.py:31-39), on a small simulated
training set (for the z-scoring buffers and the number of choice categories), optionally trains it
for a few epochs (``--train-epochs``, so that the weights are not at their initial values), evaluates
``estimator.log_prob`` on held-out rows, imports the ``state_dict`` with
``PackedMNLE.from_state_dict`` and prints the largest differences of (a) the CPU spec, (b) the fp32
CUDA kernel, (c) the tcgen05 kernel against sbi's own numbers.  Expected: ~1e-5 (fp32 noise).
Anything larger means the restatement (or the shape-driven state_dict import) does not match that
sbi version and must be fixed before trusting MNLE numbers from this package.

It also WRITES ``tests/golden/mnle_sbi.npz`` (``--out``): the estimator's ``state_dict`` (keys in registration
order + tensors), the held-out rows and sbi's own log-probs.  Commit that file and
``tests/test_mnle_sbi_fixture.py`` turns from "skipped: parity unpinned" into the pin of the whole MNLE path
(spec on the CPU; fp32, precise and tcgen05 kernels on the GPU) against real sbi numbers.

NOT run in this repository's CI: it cannot be (sbi is absent).  It is shipped, not claimed.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cuda", action="store_true")
    ap.add_argument("--train-rows", type=int, default=4096)
    ap.add_argument("--test-rows", type=int, default=2048)
    ap.add_argument("--train-epochs", type=int, default=0, help="train with sbi for this many epochs first")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                  "tests", "golden", "mnle_sbi.npz"))
    args = ap.parse_args()
    try:
        from sbi.neural_nets import likelihood_nn
    except Exception as e:  # pragma: no cover
        raise SystemExit(f"sbi is not importable here ({e!r}); install sbi==0.25.0 to run this comparison")
    from oracle import ddm_oracle as orc
    from oracle import mnle_spec as ms

    torch.manual_seed(0)
    n = args.train_rows + args.test_rows
    theta = orc.prior_sample(n, seed=1)
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(0)), 0, n, 80, 0.75))
    x, _ = orc.sim_rng_c(theta.numpy(), pulses.numpy(), seed=3)
    x = torch.from_numpy(x)
    z = torch.cat([theta, pulses], dim=1)
    tr = slice(0, args.train_rows)
    te = slice(args.train_rows, n)

    # the reference's builder call (mnle.py:31-39)
    build = likelihood_nn(model="mnle", log_transform_x=True, z_score_theta="independent", z_score_x="independent",
                          hidden_features=128, num_transforms=10, num_bins=24)
    est = build(z[tr], x[tr])           # sbi: builder(batch_theta, batch_x) with theta := condition z
    if args.train_epochs > 0:           # a few optimiser steps the way sbi's trainer takes them (-mean log_prob, Adam)
        opt = torch.optim.Adam(est.parameters(), lr=5e-4)
        est.train()
        for _ in range(args.train_epochs):
            for a in range(0, args.train_rows, 512):
                opt.zero_grad()
                loss = -est.log_prob(x[tr][a:a + 512].unsqueeze(0), condition=z[tr][a:a + 512]).mean()
                loss.backward()
                opt.step()
    est.eval()
    with torch.no_grad():
        want = est.log_prob(x[te].unsqueeze(0), condition=z[te]).reshape(-1).double()

    sd = est.state_dict()
    keys = list(sd.keys())
    import sbi
    np.savez_compressed(args.out, keys=np.array(keys), x=x[te].numpy(), z=z[te].numpy(), log_prob=want.numpy(),
                        sbi_version=np.array(getattr(sbi, "__version__", "unknown")),
                        **{f"t{i}": sd[k].detach().cpu().numpy() for i, k in enumerate(keys)})
    print(f"wrote {args.out}: {len(keys)} state_dict tensors, {want.numel()} rows (sbi {getattr(sbi, '__version__', '?')})")

    from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
    packed = PackedMNLE.from_state_dict(sd)
    print(f"imported state_dict: {packed.packed.size} packed floats, {packed.n_choices} choice categories")

    # (a) CPU spec on the imported parameters: unpack the folded buffer back into spec names
    p = spec_params_from_packed(packed)
    got = ms.log_prob(ms.cast_params(p, torch.float64), x[te], z[te])
    report("CPU spec (float64)", got, want)
    if not args.no_cuda:
        dev = DeviceMNLE(packed)
        for kernel in ("precise", "simt", "tc"):
            got = dev.log_prob(x[te], condition=z[te], kernel=kernel)[0].double()
            report(f"CUDA {kernel}", got, want)


def report(name, got, want):
    err = (got - want).abs()
    print(f"{name:22s} max |diff| {float(err.max()):.3e}   mean {float(err.mean()):.3e}   "
          f"(log-prob range [{float(want.min()):.2f}, {float(want.max()):.2f}])")


def spec_params_from_packed(packed):
    """The packed buffer (z-scoring already folded into the first layers) as a mnle_spec parameter
    dict with identity z-scoring."""
    from sbi_for_diffusion_models_b200.mnle_net import COND_DIM, CTX_DIM, HIDDEN, NUM_TRANSFORMS, SPLINE_OUT
    buf = torch.from_numpy(packed.packed.copy())
    o = [0]

    def take(*shape):
        k = int(np.prod(shape))
        t = buf[o[0]:o[0] + k].reshape(*shape)
        o[0] += k
        return t

    K = packed.n_choices
    p = {"cond_mean": torch.zeros(COND_DIM), "cond_std": torch.ones(COND_DIM)}
    p["cat.W0"], p["cat.b0"] = take(HIDDEN, COND_DIM), take(HIDDEN)
    p["cat.W1"], p["cat.b1"] = take(HIDDEN, HIDDEN), take(HIDDEN)
    p["cat.W2"], p["cat.b2"] = take(HIDDEN, HIDDEN), take(HIDDEN)
    p["cat.Wo"], p["cat.bo"] = take(K, HIDDEN), take(K)
    for k in range(NUM_TRANSFORMS):
        p[f"flow.{k}.W1"], p[f"flow.{k}.b1"] = take(HIDDEN, CTX_DIM), take(HIDDEN)
        p[f"flow.{k}.W2"], p[f"flow.{k}.b2"] = take(HIDDEN, HIDDEN), take(HIDDEN)
        p[f"flow.{k}.W3"], p[f"flow.{k}.b3"] = take(SPLINE_OUT, HIDDEN), take(SPLINE_OUT)
    p["flow.mu_y"], p["flow.sigma_y"] = take(1)[0], take(1)[0]
    return p


if __name__ == "__main__":
    main()
