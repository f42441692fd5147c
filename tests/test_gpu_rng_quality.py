"""High-power checks of the simulator's native noise (Philox4x32-10 -> six 21-bit Box-Muller fields per block on
the MUFU unit) -- the part of the path that can only match the reference in distribution.

1. 2e6 trials per theta, GPU (native noise) vs the scalar C oracle with its own CPU generator: two-sample KS
   on RT | choice and chi-square on the choice counts.  The thetas sit where a defective normal generator
   would show first: a bound a few noise steps wide (exits in the first steps are tail events of single
   normals), a decision window of only 100 steps, a strong leak, and a typical session parameter.
2. 1e9 dumped normals against the EXACT law of the generator -- the radius uniform lives on 2^21 levels, so
   |z| <= 5.40 and the far tail is a short list of atoms: even moments to order 8 and the mass beyond 4 and 5
   sigma must match that discrete law within sampling error, and the discrete law's own distance from N(0,1)
   (the documented bias bound) is asserted as well.
"""
import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch
from scipy import stats

from oracle import ddm_oracle as orc
from sbi_for_diffusion_models_b200 import simulator as sim

pytestmark = pytest.mark.gpu

THETAS = {
    "narrow bound (B = 0.08, exits within a few steps)": [0.5, 0.3, 0.01, 0.08, 0.3],
    "late start (t_nd = 7.95: 100-step window)": [0.5, 0.5, 0.3, 1.0, 7.95],
    "strong leak (lam = 3)": [0.5, 3.0, 0.6, 1.2, 0.2],
    "typical session": [0.45, 0.6, 1.3, 14.0, 0.25],
}
N_TRIALS = 2_000_000


def _oracle_trials(theta, pulses, n, seed):
    """n trials of one theta through the C oracle on all host cores (ctypes releases the GIL)."""
    workers = max(1, len(os.sched_getaffinity(0)))
    chunk = -(-n // (4 * workers))
    th = np.repeat(np.asarray([theta], np.float32), chunk, 0)

    def run(i):
        m = min(chunk, n - i * chunk)
        return orc.sim_rng_c(th[:m], pulses, seed + 7919 * i)[0]
    with ThreadPoolExecutor(workers) as ex:
        return np.concatenate(list(ex.map(run, range(-(-n // chunk)))))


@pytest.mark.parametrize("name", list(THETAS))
def test_two_million_trials_match_the_cpu_oracle_in_distribution(name):
    theta = THETAS[name]
    pulses = orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(5)), 0, 1, 80, 0.75)
    want = _oracle_trials(theta, pulses, N_TRIALS, seed=11)
    th = torch.tensor([theta]).expand(N_TRIALS, 5)
    got = sim.simulate_trials(th, torch.from_numpy(pulses), seed=20261018).cpu().numpy()
    alpha = 1e-3 / 12          # 4 thetas x (choice counts + RT | choice for two bounds), Bonferroni
    cg, cw = np.bincount(got[:, 1].astype(int), minlength=3), np.bincount(want[:, 1].astype(int), minlength=3)
    keep = (cg + cw) > 0
    if keep.sum() > 1:
        p = stats.chi2_contingency(np.stack([cg[keep], cw[keep]]))[1]
        assert p > alpha, (name, "choice counts", cg.tolist(), cw.tolist(), p)
    for c in (0, 1):
        a, b = got[got[:, 1] == c, 0], want[want[:, 1] == c, 0]
        if min(len(a), len(b)) < 1000:
            continue
        res = stats.ks_2samp(a, b)
        assert res.pvalue > alpha, (name, f"RT | choice {c}", len(a), len(b), res.statistic, res.pvalue)


def _discrete_law():
    """Exact law of the generator's radius: u = k / 2^21, k = 1 .. 2^21, r = sqrt(-2 ln u); angle uniform (2^21
    levels: its even cosine moments equal the continuous ones to 1e-12).  -> E z^(2j), P(|z| > 4), P(|z| > 5)."""
    k = np.arange(1, 2 ** 21 + 1, dtype=np.float64)
    r2 = -2.0 * np.log(k / 2.0 ** 21)
    cos_m = {1: 0.5, 2: 3.0 / 8, 3: 5.0 / 16, 4: 35.0 / 128}          # E cos^(2j) phi
    moments = {2 * j: float(np.mean(r2 ** j)) * cos_m[j] for j in (1, 2, 3, 4)}
    tail = {}
    for c in (4.0, 5.0):
        r = np.sqrt(r2[r2 > c * c])
        tail[c] = float(np.sum(2.0 / math.pi * np.arccos(c / r))) / 2.0 ** 21     # P(|r cos phi| > c)
    return moments, tail


def test_one_billion_normals_follow_the_generators_exact_law():
    n_trials, steps_per_chunk, chunks = 1 << 20, 120, 8                  # 8 x 1.26e8 = 1.007e9 normals
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    beyond = torch.zeros(2, dtype=torch.float64, device="cuda")
    zmax = 0.0
    for c in range(chunks):
        z = sim.philox_normals(99, n_trials, steps_per_chunk, trial_offset=c * n_trials).double()
        z2 = z * z
        sums += torch.stack([z2.sum(), (z2 * z2).sum(), (z2 * z2 * z2).sum(), (z2 * z2 * z2 * z2).sum()])
        beyond += torch.stack([(z2 > 16.0).sum(), (z2 > 25.0).sum()]).double()
        zmax = max(zmax, float(z.abs().max()))
        assert abs(float(z.mean())) < 6.0 / math.sqrt(z.numel())
    n = float(chunks * n_trials * steps_per_chunk)
    got = (sums / n).tolist()
    law_m, law_t = _discrete_law()
    gauss_m = {2: 1.0, 4: 3.0, 6: 15.0, 8: 105.0}
    gauss_m2 = {2: 3.0, 4: 105.0, 6: 10395.0, 8: 2027025.0}              # E z^(2k) of N(0,1), for the sampling error
    for i, k in enumerate((2, 4, 6, 8)):
        sd = math.sqrt((gauss_m2[k] - gauss_m[k] ** 2) / n)
        # the MUFU lg2 / sqrt / sin / cos approximations move a moment by ~1e-6 relative: far below 5 sd here
        assert abs(got[i] - law_m[k]) < 5.0 * sd, (k, got[i], law_m[k], sd)
    # the law itself against N(0,1): what truncation at 5.40 sigma and the 2^-21 radius grid cost
    # (variance -3.9e-6, 4th moment -3.2e-5 relative, 6th -1.8e-4, 8th -7.3e-4)
    assert abs(law_m[2] - 1.0) < 5e-6 and abs(law_m[4] - 3.0) < 1.2e-4
    assert abs(law_m[6] - 15.0) < 3e-3 and abs(law_m[8] - 105.0) < 0.1
    for j, c in enumerate((4.0, 5.0)):
        expect = law_t[c] * n
        assert abs(float(beyond[j]) - expect) < 5.0 * math.sqrt(expect), (c, float(beyond[j]), expect)
        gauss = 2.0 * stats.norm.sf(c)
        # mass beyond 4 sigma: -0.19 % of the Gaussian's 6.33e-5; beyond 5 sigma (seven atoms of the radius): -14 % of 5.7e-7
        assert abs(law_t[c] / gauss - 1.0) < (2.5e-3 if c == 4.0 else 0.15), (c, law_t[c], gauss)
    assert zmax <= math.sqrt(2.0 * 21.0 * math.log(2.0)) * (1.0 + 1e-5)   # |z| <= sqrt(-2 ln 2^-21) = 5.396
