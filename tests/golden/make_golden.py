"""Generate tests/golden/*.npz by importing the UNMODIFIED reference from /root/reference/src.

Run in the build container only (``python tests/golden/make_golden.py``); the GPU box has
no /root/reference, so tests read the committed fixtures and never this script's imports.

Shared noise: the reference has no noise argument (rt_choice_model.py:186 calls
``torch.randn((N,))`` once per executed step), so ``torch.randn`` is swapped for a server
that hands out row k of a pre-built (n_max, N) tensor on its k-th call inside each
``_simulate_rt_choice_batch_torch`` call.  The tensor itself is
``oracle.ddm_oracle.synthetic_noise(seed, n_max, N)`` -- integer-built, so a fixture only
stores the seed.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

# data_simulator.py:4 / mnle.py:6 import matplotlib but the simulator never uses it
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))

from oracle import ddm_oracle as orc  # noqa: E402

import sbi_for_diffusion_models.models.rt_choice_model as ref_rt  # noqa: E402
import sbi_for_diffusion_models.data_simulator as ref_ds  # noqa: E402
import sbi_for_diffusion_models.proposals as ref_prop  # noqa: E402
import sbi_for_diffusion_models.potentials as ref_pot  # noqa: E402


class SharedNoise:
    """Serve pre-built noise rows to the reference, one simulator call at a time."""

    def __init__(self, base_seed: int):
        self.base_seed = int(base_seed)
        self.calls = 0
        self.seeds = []
        self._rows = None
        self._k = 0
        self._orig_sim = ref_rt._simulate_rt_choice_batch_torch
        self._orig_randn = torch.randn

    def _randn(self, shape, **kw):
        row = self._rows[self._k]
        self._k += 1
        assert tuple(shape) == (row.shape[0],)
        return row

    def _sim(self, theta, **kw):
        n = theta.shape[0]
        n_max, _ = ref_rt.pulse_schedule(dt=float(ref_rt.DT_CHOICE))
        seed = self.base_seed + self.calls
        self.calls += 1
        self.seeds.append(seed)
        self._rows = torch.from_numpy(orc.synthetic_noise(seed, n_max, n))
        self._k = 0
        return self._orig_sim(theta, **kw)

    def __enter__(self):
        ref_rt._simulate_rt_choice_batch_torch = self._sim
        torch.randn = self._randn
        return self

    def __exit__(self, *exc):
        ref_rt._simulate_rt_choice_batch_torch = self._orig_sim
        torch.randn = self._orig_randn


def run_ref(theta, pulses, seed):
    with SharedNoise(seed):
        x = ref_rt.rt_choice_model_simulator_torch(torch.as_tensor(theta), mu_sensory=1.0,
                                                   pulse_sides=torch.as_tensor(pulses))
    return x.numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path)} bytes")


def pulses_ref(seed, n, P, p=0.75):
    return ref_rt.generate_pulse_matrix_numpy(np.random.default_rng(seed), n, P, p_success=p)


def edge_thetas():
    rows = [
        # a0,   lam,   v,    B,     t_nd
        [0.5, 0.3, 1.0, 12.0, 7.9999995],   # t_nd clamps to 8-1e-6 -> n_steps = 0
        [0.5, 0.3, 1.0, 12.0, 9.0],         # above the clamp
        [0.5, 0.3, 1.0, 12.0, -1.0],        # below the clamp -> 0
        [0.0, 0.3, 1.0, 12.0, 0.2],         # starts on the lower bound
        [1.0, 0.3, 1.0, 12.0, 0.2],         # starts on the upper bound
        [-0.3, 0.3, 1.0, 12.0, 0.2],        # a0 clamps to 0
        [1.7, 0.3, 1.0, 12.0, 0.2],         # a0 clamps to 1
        [0.5, -0.4, 1.0, 12.0, 0.2],        # negative leak (unstable), no clamp on lam
        [0.5, 0.3, -2.0, 12.0, 0.2],        # v through abs
        [0.5, 0.3, 1.0, -12.0, 0.2],        # B through abs
        [0.5, 0.3, 1.0, 0.0, 0.2],          # B floors at 1e-6
        [0.5, 0.3, 1.0, 1e-9, 0.2],
        [0.5, 0.0, 0.0, 1e6, 0.0],          # unreachable bounds, full window
        [0.5, 0.0, 0.0, 1e6, 3.99975],      # unreachable bounds, mid window
        [0.5, 50.0, 1.0, 12.0, 0.1],        # strong leak
        [0.5, 1999.0, 1.0, 12.0, 0.1],      # lam*dt ~ 1
        [0.5, 4100.0, 1.0, 12.0, 0.1],      # lam*dt > 2: oscillating blow-up
        [0.5, 0.3, 40.0, 12.0, 0.1],        # first kick crosses a bound
        [0.5, 0.3, 1.0, 0.05, 0.1],         # tiny bound, early hits
        [0.999999, 0.0, 0.0, 5.0, 0.3],
        [1e-7, 0.0, 0.0, 5.0, 0.3],
        [0.5, 0.3, 3.0, 30.0, 7.9],         # short window (200 steps)
        [0.5, 0.3, 3.0, 30.0, 7.9995],      # one-step window
        [0.5, 0.3, 3.0, 30.0, 7.99975],     # window rounds to 0
    ]
    return np.asarray(rows, dtype=np.float32)


def main():
    torch.manual_seed(0)
    n_max, spp = ref_rt.pulse_schedule()
    P = ref_rt.n_pulses_max_from_schedule(n_max, spp)
    assert (n_max, spp, P) == (16000, 200, 80)

    # 1. prior-shaped thetas, default schedule
    N = 384
    theta = orc.prior_sample(N, seed=1).numpy()
    pulses = pulses_ref(7, N, P)
    x = run_ref(theta, pulses, 11)
    save("sim_prior", theta=theta, pulses=pulses.astype(np.int8), noise_seed=np.int64(11), x=x,
         x_packed_log=ref_rt.pack_x_rt_choice(torch.from_numpy(x), log_rt=True).numpy(),
         x_packed_raw=ref_rt.pack_x_rt_choice(torch.from_numpy(x), log_rt=False).numpy())

    # 2. edge cases
    th = edge_thetas()
    pl = pulses_ref(8, th.shape[0], P)
    save("sim_edges", theta=th, pulses=pl.astype(np.int8), noise_seed=np.int64(12),
         x=run_ref(th, pl, 12))

    # 3. decision-window sweep: unreachable bounds so hit_step == n_steps shows up in rt
    rs = np.random.RandomState(3)
    t_nd = np.concatenate([
        np.linspace(0.0, 8.0, 801),
        rs.uniform(0.0, 8.0, size=1200),
        8.0 - np.arange(1, 48) * 5e-4,            # exact multiples of dt from the end
        np.nextafter(np.float32(8.0) - np.arange(1, 48, dtype=np.float32) * np.float32(5e-4), np.float32(0)),
        np.nextafter(np.float32(8.0) - np.arange(1, 48, dtype=np.float32) * np.float32(5e-4), np.float32(9)),
    ]).astype(np.float32)
    th = np.zeros((t_nd.shape[0], 5), dtype=np.float32)
    th[:, 0] = 0.5
    th[:, 3] = 1e6
    th[:, 4] = t_nd
    pl = pulses_ref(9, 1, P)
    save("sim_window", theta=th, pulses=pl.astype(np.int8), noise_seed=np.int64(13),
         x=run_ref(th, pl, 13))

    # 4. one broadcast pulse row, and a longer-than-needed pulse matrix (tail ignored, :178)
    N = 96
    theta = orc.prior_sample(N, seed=2).numpy()
    pl1 = pulses_ref(10, 1, P)
    pl96 = pulses_ref(11, N, 96)
    save("sim_shapes", theta=theta, pulses_row=pl1.astype(np.int8), pulses_wide=pl96.astype(np.int8),
         noise_seed_row=np.int64(14), x_row=run_ref(theta, pl1, 14),
         noise_seed_wide=np.int64(15), x_wide=run_ref(theta, pl96, 15))

    # 5. non-binary pulse values (the simulator multiplies whatever it is given, :192)
    N = 64
    theta = orc.prior_sample(N, seed=3).numpy()
    plf = (pulses_ref(12, N, P) * rs.choice([0.0, 0.5, 1.0, 2.0], size=(N, P))).astype(np.float32)
    save("sim_realpulses", theta=theta, pulses=plf, noise_seed=np.int64(16), x=run_ref(theta, plf, 16))

    # 6. other schedules (module constants patched the way a user would edit constants.py)
    sched = {}
    for tag, dt, interval in (("dt1e-3", 1e-3, 0.1), ("dt2e-3", 2e-3, 0.1), ("dt1e-3_i50", 1e-3, 0.05),
                              ("dt2.5e-3_i30", 2.5e-3, 0.0325)):
        old = (ref_rt.DT_CHOICE, ref_rt.PULSE_INTERVAL)
        ref_rt.DT_CHOICE, ref_rt.PULSE_INTERVAL = dt, interval
        try:
            nm, sp = ref_rt.pulse_schedule(dt=dt)
            Pn = ref_rt.n_pulses_max_from_schedule(nm, sp)
            N = 128
            theta = orc.prior_sample(N, seed=4).numpy()
            pl = pulses_ref(13, N, Pn)
            sched[tag + "_x"] = run_ref(theta, pl, 17)
            sched[tag + "_pulses"] = pl.astype(np.int8)
            sched[tag + "_meta"] = np.asarray([dt, interval, nm, sp, Pn], dtype=np.float64)
        finally:
            ref_rt.DT_CHOICE, ref_rt.PULSE_INTERVAL = old
    save("sim_schedules", theta=orc.prior_sample(128, seed=4).numpy(), noise_seed=np.int64(17), **sched)

    # 7. PCG64 pulse streams
    pc = {}
    for seed in (0, 123, 2**31 - 2):
        pc[f"seed{seed}"] = pulses_ref(seed, 64, P).astype(np.int8)
    for p in (0.0, 0.5, 1.0, 0.3):
        pc[f"p{p}"] = pulses_ref(5, 32, P, p).astype(np.int8)
    rng = np.random.default_rng(21)
    pc["stream_a"] = ref_rt.generate_pulse_matrix_numpy(rng, 10, P, p_success=0.75).astype(np.int8)
    pc["stream_b"] = ref_rt.generate_pulse_matrix_numpy(rng, 7, 33, p_success=0.75).astype(np.int8)
    pc["stream_c"] = ref_rt.generate_pulse_matrix_numpy(rng, 5, P, p_success=0.75).astype(np.int8)
    prop = ref_prop.PulseSequenceProposal(P=P, p_success=0.75, seed=0)
    pc["proposal_first"] = prop.sample((6,)).numpy().astype(np.int8)
    pc["proposal_second"] = prop.sample((3, 2)).numpy().astype(np.int8)
    pc["proposal_scalar"] = prop.sample().numpy().astype(np.int8)
    save("pulses_pcg64", **pc)

    # 8. session helpers
    theta_true = torch.tensor([0.45, 0.6, 1.3, 14.0, 0.25])
    with SharedNoise(30):
        x_o, pulses_o = ref_ds.simulate_observed_session(theta_true, 50, "cpu", mu_sensory=1.0,
                                                         p_success=0.75, P=P, seed=123, log_rt=False)
    with SharedNoise(31):
        x_log, _ = ref_ds.simulate_observed_session(theta_true, 50, "cpu", mu_sensory=1.0,
                                                    p_success=0.75, P=P, seed=123, log_rt=True)
    with SharedNoise(32):
        xs, ss = ref_rt.simulate_session_data_rt_choice(theta_true, 40, rng=np.random.default_rng(77),
                                                        mu_sensory=1.0, p_success=0.75,
                                                        return_pulse_sides=True)
    save("sessions", theta_true=theta_true.numpy(), x_o=x_o.numpy(), pulses_o=pulses_o.numpy().astype(np.int8),
         x_o_log=x_log.numpy(), sess_x=xs.numpy(), sess_pulses=ss.numpy().astype(np.int8))

    # 9. training-set shell: 300 simulations in batches of 128 (three simulator calls)
    class PriorStub:
        def __init__(self):
            self.k = 0

        def sample(self, shape=torch.Size()):
            n = int(np.prod(shape)) if len(shape) else 1
            self.k += 1
            return orc.prior_sample(n, seed=100 + self.k)

        def log_prob(self, th):
            return torch.zeros(th.shape[:-1])

    proposal = ref_prop.ExtendedProposal(PriorStub(), ref_prop.PulseSequenceProposal(P=P, p_success=0.75, seed=0))
    with SharedNoise(40) as sn:
        z_all, x_all = ref_ds.simulate_training_set_with_conditions(
            proposal, 300, 128, "cpu", mu_sensory=1.0, p_success=0.75, P=P, log_rt=False)
        seeds = list(sn.seeds)
    save("training_set", z=z_all.numpy(), x=x_all.numpy(), noise_seeds=np.asarray(seeds, dtype=np.int64))

    # 10. potential row layout with a stub estimator that encodes (x, condition) into a number
    class StubEstimator(torch.nn.Module):
        def log_prob(self, x, condition):
            w = torch.arange(1, condition.shape[1] + 1, dtype=torch.float32) * 1e-2
            return ((condition * w).sum(-1) + 3.0 * x[0, :, 0] - 0.5 * x[0, :, 1]).unsqueeze(0)

    T, C = 7, 5
    pul = torch.from_numpy(pulses_ref(50, T, P))
    x_obs = torch.stack([torch.linspace(0.3, 2.0, T), torch.tensor([0., 1., 2., 1., 0., 0., 1.])], dim=1)
    thetas = orc.prior_sample(C, seed=9)
    cll = ref_pot.ConditionedMNLELogLikelihood(StubEstimator(), pul, "cpu")
    ll = cll(thetas, x_obs, track_gradients=False)
    save("potential_layout", pulses=pul.numpy().astype(np.int8), x_obs=x_obs.numpy(), thetas=thetas.numpy(),
         ll=ll.numpy())

    # 11. SBC ranks helper is in mnle.py (needs sbi); restate input/output by hand from mnle.py:98-104
    print("done")


if __name__ == "__main__":
    main()
