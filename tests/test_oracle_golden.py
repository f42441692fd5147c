"""The CPU oracle (both restatements) against outputs of the imported reference.

Fixtures: tests/golden/*.npz written by tests/golden/make_golden.py.  Bit-exact.
"""
import numpy as np
import pytest
import torch

from oracle import ddm_oracle as orc


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _assert_same(got, want, what):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    bad = np.nonzero((_bits(got) != _bits(want)).any(axis=-1))[0]
    assert bad.size == 0, f"{what}: {bad.size} rows differ, first {bad[:5]}: {got[bad[:3]]} vs {want[bad[:3]]}"


def _both(theta, pulses, seed, **sched):
    n_max, _, _ = orc.schedule(**sched)
    N = theta.shape[0]
    noise = orc.synthetic_noise(int(seed), n_max, N)
    xc, steps = orc.sim_scalar_c(theta, pulses, noise, **sched)
    rows = torch.from_numpy(noise)
    xt, when, _ = orc.sim_lockstep_torch(torch.from_numpy(theta), torch.from_numpy(np.asarray(pulses, np.float32)),
                                         lambda k, n: rows[k], **sched)
    assert np.array_equal(steps, when.numpy())
    return xc, xt.numpy()


@pytest.mark.parametrize("name", ["sim_prior", "sim_edges", "sim_window", "sim_realpulses"])
def test_simulator_cases(golden, name):
    g = golden(name)
    xc, xt = _both(g["theta"], g["pulses"].astype(np.float32), g["noise_seed"])
    _assert_same(xc, g["x"], name + " (C)")
    _assert_same(xt, g["x"], name + " (torch)")


def test_outcome_mix_is_nontrivial(golden):
    x = golden("sim_prior")["x"]
    counts = np.bincount(x[:, 1].astype(int), minlength=3)
    assert (counts > 20).all(), counts


def test_window_formula_covers_every_length(golden):
    g = golden("sim_window")
    # unreachable bounds: rt - t_nd encodes n_steps; the sweep must hit the boundary values
    assert (g["x"][:, 1] == 2).all()
    assert g["x"][:, 0].max() <= 8.0


def test_shapes(golden):
    g = golden("sim_shapes")
    xc, xt = _both(g["theta"], g["pulses_row"].astype(np.float32), g["noise_seed_row"])
    _assert_same(xc, g["x_row"], "broadcast row (C)")
    _assert_same(xt, g["x_row"], "broadcast row (torch)")
    xc, xt = _both(g["theta"], g["pulses_wide"].astype(np.float32), g["noise_seed_wide"])
    _assert_same(xc, g["x_wide"], "wide (C)")
    _assert_same(xt, g["x_wide"], "wide (torch)")


@pytest.mark.parametrize("tag", ["dt1e-3", "dt2e-3", "dt1e-3_i50", "dt2.5e-3_i30"])
def test_schedules(golden, tag):
    g = golden("sim_schedules")
    dt, interval, n_max, spp, P = g[tag + "_meta"]
    assert orc.schedule(dt=dt, pulse_interval=interval) == (int(n_max), int(spp), int(P))
    xc, xt = _both(g["theta"], g[tag + "_pulses"].astype(np.float32), g["noise_seed"], dt=float(dt),
                   pulse_interval=float(interval))
    _assert_same(xc, g[tag + "_x"], tag + " (C)")
    _assert_same(xt, g[tag + "_x"], tag + " (torch)")


def test_pack(golden):
    g = golden("sim_prior")
    x = torch.from_numpy(g["x"])
    _assert_same(orc.pack_x(x, False).numpy(), g["x_packed_raw"], "pack raw")
    _assert_same(orc.pack_x(x, True).numpy(), g["x_packed_log"], "pack log (torch.log on the same host)")
    # C path uses libm logf: allow 2 ulp against torch's vectorised log
    noise = orc.synthetic_noise(int(g["noise_seed"]), 16000, g["theta"].shape[0])
    xc, _ = orc.sim_scalar_c(g["theta"], g["pulses"].astype(np.float32), noise, log_rt=True)
    d = np.abs(_bits(xc[:, 0]).astype(np.int64) - _bits(g["x_packed_log"][:, 0]).astype(np.int64))
    tiny = np.abs(g["x_packed_log"][:, 0]) < 1e-3
    assert d[~tiny].max() <= 2
    assert np.allclose(xc[tiny, 0], g["x_packed_log"][tiny, 0], atol=3e-7)
    assert np.array_equal(xc[:, 1], g["x_packed_log"][:, 1])


def test_training_set_shell(golden):
    g = golden("training_set")
    z, x, seeds = g["z"], g["x"], g["noise_seeds"]
    assert z.shape == (300, 85) and x.shape == (300, 2) and len(seeds) == 3
    start = 0
    for seed, bs in zip(seeds, (128, 128, 44)):
        zz = z[start:start + bs]
        noise = orc.synthetic_noise(int(seed), 16000, bs)
        xc, _ = orc.sim_scalar_c(zz[:, :5], zz[:, 5:], noise)
        _assert_same(xc, x[start:start + bs], f"batch at {start}")
        start += bs


def test_sessions(golden):
    g = golden("sessions")
    th = np.repeat(g["theta_true"][None, :], 50, axis=0)
    x, _ = orc.sim_scalar_c(th, g["pulses_o"].astype(np.float32), orc.synthetic_noise(30, 16000, 50))
    _assert_same(x, g["x_o"], "observed session")
    rng = np.random.default_rng(123)
    assert np.array_equal(orc.pulses_loop_numpy(rng, 50, 80, 0.75), g["pulses_o"].astype(np.float32))
    th = np.repeat(g["theta_true"][None, :], 40, axis=0)
    x, _ = orc.sim_scalar_c(th, g["sess_pulses"].astype(np.float32), orc.synthetic_noise(32, 16000, 40))
    _assert_same(x, g["sess_x"], "session data")
