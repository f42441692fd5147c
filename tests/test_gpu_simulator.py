"""Parity of the CUDA simulator (through the C ABI) with the CPU oracle and the reference's
golden vectors.  Run on the B200 box: ``pytest -m gpu``."""
import os

import numpy as np
import pytest
import torch

from oracle import ddm_oracle as orc
from sbi_for_diffusion_models_b200 import simulator as sim
from sbi_for_diffusion_models_b200.simulator import Schedule, simulate_trials

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


def _assert_same(got, want, what):
    got = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    bad = np.nonzero((_bits(got) != _bits(want)).reshape(len(want), -1).any(axis=1))[0]
    assert bad.size == 0, f"{what}: {bad.size}/{len(want)} rows differ, first {bad[:5]}: {got[bad[:3]]} vs {want[bad[:3]]}"


def _gpu_shared(theta, pulses, seed, sched=None, **kw):
    sched = sched or Schedule.from_constants()
    noise = torch.from_numpy(orc.synthetic_noise(int(seed), sched.n_max, theta.shape[0]))
    return simulate_trials(torch.from_numpy(theta), torch.from_numpy(np.asarray(pulses, np.float32)), noise=noise,
                           schedule=sched, **kw)


# ---------------------------------------------------------------- shared noise: bit-exact ---

@pytest.mark.parametrize("name", ["sim_prior", "sim_edges", "sim_window", "sim_realpulses"])
def test_golden_shared_noise(golden, name):
    g = golden(name)
    x, steps, stats = _gpu_shared(g["theta"], g["pulses"], g["noise_seed"], return_steps=True, return_stats=True)
    _assert_same(x, g["x"], name)
    assert stats.useful_steps == int(steps.sum())
    if name == "sim_realpulses":
        assert stats.generic_rows > 0          # non-binary rows took the global-read kick path
    if name == "sim_prior":
        assert stats.generic_rows == 0


def test_golden_shapes(golden):
    g = golden("sim_shapes")
    _assert_same(_gpu_shared(g["theta"], g["pulses_row"], g["noise_seed_row"]), g["x_row"], "broadcast row")
    _assert_same(_gpu_shared(g["theta"], g["pulses_row"][0], g["noise_seed_row"]), g["x_row"], "1-D pulses")
    _assert_same(_gpu_shared(g["theta"], g["pulses_wide"], g["noise_seed_wide"]), g["x_wide"], "wide pulses")


@pytest.mark.parametrize("tag", ["dt1e-3", "dt2e-3", "dt1e-3_i50", "dt2.5e-3_i30"])
def test_golden_schedules(golden, tag):
    g = golden("sim_schedules")
    dt, interval, n_max, spp, P = g[tag + "_meta"]
    sched = Schedule.from_constants(dt=float(dt), pulse_interval=float(interval))
    assert (sched.n_max, sched.steps_per_pulse, sched.n_pulses) == (int(n_max), int(spp), int(P))
    _assert_same(_gpu_shared(g["theta"], g[tag + "_pulses"], g["noise_seed"], sched), g[tag + "_x"], tag)


def test_golden_log_rt_epilogue(golden):
    g = golden("sim_prior")
    x = _gpu_shared(g["theta"], g["pulses"], g["noise_seed"], log_rt=True).cpu().numpy()
    want = g["x_packed_log"]
    assert np.array_equal(x[:, 1], want[:, 1])
    ulp = np.abs(_bits(x[:, 0]).astype(np.int64) - _bits(want[:, 0]).astype(np.int64))
    near0 = np.abs(want[:, 0]) < 1e-3
    assert ulp[~near0].max() <= 2, "fp32 log: tolerance 2 ulp against torch.log on the CPU"
    assert np.allclose(x[near0, 0], want[near0, 0], atol=3e-7)


def test_oracle_shared_noise_large():
    """8192 prior-shaped trials, default schedule, against the scalar C oracle."""
    n = 8192
    theta = orc.prior_sample(n, seed=21).numpy()
    pulses = orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(5)), 0, n, 80, 0.75)
    noise = orc.synthetic_noise(99, 16000, n)
    want, want_steps = orc.sim_scalar_c(theta, pulses, noise)
    x, steps = simulate_trials(torch.from_numpy(theta), torch.from_numpy(pulses), noise=torch.from_numpy(noise),
                               return_steps=True)
    _assert_same(x, want, "8192 trials")
    assert np.array_equal(steps.cpu().numpy().astype(np.int64), want_steps)


def test_strided_z_views_and_theta_broadcast():
    n = 700
    z = torch.empty((n, 85))
    z[:, :5] = orc.prior_sample(n, seed=22)
    z[:, 5:] = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(6)), 0, n, 80, 0.75))
    noise = orc.synthetic_noise(5, 16000, n)
    want, _ = orc.sim_scalar_c(z[:, :5].numpy(), z[:, 5:].numpy(), noise)
    zc = z.cuda()
    got = simulate_trials(zc[:, :5], zc[:, 5:], noise=torch.from_numpy(noise))   # ld = 85 for both views
    _assert_same(got, want, "views of z")
    th1 = z[3:4, :5]
    want, _ = orc.sim_scalar_c(np.repeat(th1.numpy(), n, 0), z[:, 5:].numpy(), noise)
    got = simulate_trials(th1.cuda().expand(n, 5), zc[:, 5:], noise=torch.from_numpy(noise))   # ld_theta = 0
    _assert_same(got, want, "broadcast theta")


def test_empty_and_single():
    x = simulate_trials(torch.zeros((0, 5)), torch.ones((0, 80)))
    assert tuple(x.shape) == (0, 2)
    x = simulate_trials(torch.tensor([0.5, 0.3, 1.0, 12.0, 0.2]), torch.ones(80), seed=1)
    assert tuple(x.shape) == (1, 2) and x[0, 1].item() in (0.0, 1.0, 2.0)


def test_value_errors_match_reference():
    th = torch.zeros((4, 5))
    with pytest.raises(ValueError, match="Expected theta shape"):
        simulate_trials(torch.zeros((4, 4)), torch.ones((4, 80)))
    with pytest.raises(ValueError, match="first dim must match batch size"):
        simulate_trials(th, torch.ones((3, 80)))
    with pytest.raises(ValueError, match="needs at least 80"):
        simulate_trials(th, torch.ones((4, 79)))
    with pytest.raises(ValueError, match="must have shape"):
        simulate_trials(th, torch.ones((4, 2, 80)))


# ------------------------------------------------------------------- native Philox noise ---

def test_philox_words_match_oracle():
    seed, off = 0x0123456789ABCDEF, (1 << 32) - 37       # trial index crosses 2^32
    got = sim.philox_words(seed, 100, 41, trial_offset=off).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, orc.philox_words(seed, off, 100, 41))


def test_philox_normals_are_standard_normal():
    from scipy import stats
    z = sim.philox_normals(7, 4096, 512).cpu().numpy().astype(np.float64)
    assert np.isfinite(z).all() and np.abs(z).max() < 5.8
    flat = z.ravel()
    assert abs(flat.mean()) < 4 / np.sqrt(flat.size)
    assert abs(flat.var() - 1.0) < 5e-3
    assert abs(stats.skew(flat)) < 0.01 and abs(stats.kurtosis(flat)) < 0.02
    assert stats.kstest(flat[:: 7], "norm").pvalue > 0.01
    # no correlation between the two outputs of a Box-Muller pair, nor along time
    assert abs(np.corrcoef(z[0::4].ravel(), z[1::4].ravel())[0, 1]) < 0.005
    assert abs(np.corrcoef(z[:-1].ravel(), z[1:].ravel())[0, 1]) < 0.005
    # nor across neighbouring trials
    assert abs(np.corrcoef(z[:, :-1].ravel(), z[:, 1:].ravel())[0, 1]) < 0.005


def test_native_run_replays_through_the_oracle_bit_for_bit():
    """Philox path == shared-noise path == CPU oracle when the oracle is fed the normals the
    kernel consumed."""
    n, seed, off = 4096, 1234567, 10**10
    theta = orc.prior_sample(n, seed=23).numpy()
    pulses = orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(8)), 0, n, 80, 0.75)
    th, pl = torch.from_numpy(theta).cuda(), torch.from_numpy(pulses).cuda()
    x_native, steps = simulate_trials(th, pl, seed=seed, trial_offset=off, return_steps=True)
    normals = sim.philox_normals(seed, n, 16000, trial_offset=off)
    x_inject = simulate_trials(th, pl, noise=normals)
    assert torch.equal(x_native, x_inject)
    want, want_steps = orc.sim_scalar_c(theta, pulses, normals.cpu().numpy())
    _assert_same(x_native, want, "native vs oracle replay")
    assert np.array_equal(steps.cpu().numpy().astype(np.int64), want_steps)


def test_results_do_not_depend_on_batching():
    n, seed = 5000, 77
    theta = orc.prior_sample(n, seed=24).cuda()
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(9)), 0, n, 80, 0.75)).cuda()
    whole = simulate_trials(theta, pulses, seed=seed)
    parts = [simulate_trials(theta[a:b], pulses[a:b], seed=seed, trial_offset=a)
             for a, b in ((0, 1), (1, 1300), (1300, 1301), (1301, 5000))]
    assert torch.equal(whole, torch.cat(parts))
    again = simulate_trials(theta, pulses, seed=seed)
    assert torch.equal(whole, again)
    other = simulate_trials(theta, pulses, seed=seed + 1)
    assert not torch.equal(whole, other)


def test_torch_manual_seed_controls_default_key():
    theta = orc.prior_sample(256, seed=25)
    pulses = torch.ones((1, 80))
    torch.manual_seed(3)
    a = simulate_trials(theta, pulses)
    torch.manual_seed(3)
    b = simulate_trials(theta, pulses)
    c = simulate_trials(theta, pulses)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_ks_equivalence_per_theta():
    """RT | choice and choice frequencies under native Philox noise vs the CPU oracle with
    its own generator: two-sample KS p > 0.01 per theta (north_star), Bonferroni over the
    32 x 3 comparisons for the overall assertion (>= 32 thetas x >= 20 000 trials)."""
    from scipy import stats
    n_theta, n_trials = 32, 20000
    thetas = orc.prior_sample(n_theta, seed=31).numpy()
    thetas[:, 4] = np.minimum(thetas[:, 4], 0.6)
    fails, pvals = [], []
    for k in range(n_theta):
        pulses = orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(100 + k)), 0, 1, 80, 0.75)
        th = np.repeat(thetas[k:k + 1], n_trials, axis=0)
        cpu, _ = orc.sim_rng_c(th, pulses, seed=500 + k)
        gpu = simulate_trials(torch.from_numpy(thetas[k]).cuda().view(1, 5).expand(n_trials, 5),
                              torch.from_numpy(pulses), seed=900 + k).cpu().numpy()
        counts = np.stack([np.bincount(a[:, 1].astype(int), minlength=3) for a in (cpu, gpu)])
        keep = counts.sum(0) > 0
        p_choice = stats.chi2_contingency(counts[:, keep]).pvalue if keep.sum() > 1 else 1.0
        pvals.append(p_choice)
        for c in (0, 1):
            a, b = cpu[cpu[:, 1] == c, 0], gpu[gpu[:, 1] == c, 0]
            if min(len(a), len(b)) >= 200:
                pvals.append(stats.ks_2samp(a, b).pvalue)
        if min(pvals[-3:]) < 0.01:
            fails.append((k, pvals[-3:]))
    pvals = np.asarray(pvals)
    assert pvals.min() > 0.01 / len(pvals), f"distribution mismatch: {fails}"
    assert (pvals < 0.01).mean() < 0.08, f"too many small p-values: {fails}"


def test_pure_diffusion_matches_closed_forms():
    """lam = 0, v = 0: Brownian motion on [0, B] -- P(upper) = a0 and E[exit time] = a(B-a)/sigma^2."""
    n = 200000
    a0, B = 0.3, 1.5
    theta = torch.tensor([a0, 0.0, 0.0, B, 0.0]).cuda().view(1, 5).expand(n, 5)
    x, steps = simulate_trials(theta, torch.ones((1, 80)), seed=5, return_steps=True)
    x = x.cpu().numpy()
    assert (x[:, 1] != 2).mean() > 0.999
    p_up = (x[:, 1] == 1).mean()
    assert abs(p_up - a0) < 4 * np.sqrt(a0 * (1 - a0) / n) + 0.01      # + O(sqrt(dt)) overshoot bias
    mean_t = x[:, 0].mean()
    assert abs(mean_t - a0 * B * (B - a0 * B)) < 0.03


# ---------------------------------------------------------------- full-size properties ---

def test_million_trials_properties():
    n = 1 << 20
    theta = orc.prior_sample(n, seed=41).cuda()
    rng = np.random.default_rng(3)
    from sbi_for_diffusion_models_b200.pulses import generate_pulse_matrix_device
    pulses = generate_pulse_matrix_device(rng, n, 80, p_success=0.75)
    x, steps, stats = simulate_trials(theta, pulses, seed=2024, return_steps=True, return_stats=True)
    assert torch.isfinite(x).all()
    ch = x[:, 1]
    assert bool(((ch == 0) | (ch == 1) | (ch == 2)).all())
    assert float(x[:, 0].min()) >= 1e-6 and float(x[:, 0].max()) <= 8.0
    assert stats.useful_steps == int(steps.to(torch.int64).sum())
    # rt is exactly t_nd + hit_step * dt in fp32 (reference :218)
    t_nd = theta[:, 4].clamp(0.0, float(np.float32(8.0 - 1e-6)))
    rt = (t_nd + steps.to(torch.float32) * float(np.float32(5e-4))).clamp(1e-6, 8.0)
    assert torch.equal(rt, x[:, 0])
    frac = torch.bincount(ch.to(torch.int64), minlength=3).float() / n
    assert abs(frac[0] - 0.567) < 0.01 and abs(frac[1] - 0.259) < 0.01 and abs(frac[2] - 0.175) < 0.01
    assert abs(steps.float().mean().item() - 5080) < 60
    assert stats.lane_efficiency > 0.6, stats   # 3.5 trials per lane: the drain tail is ~25 %; 0.995 at 1e8 trials


def test_full_size_training_set_properties():
    """BASELINE configs[1] at its full size (1e8 trials, z = 34 GB resident; 2^25 trials when the device
    has less than 60 GB free): size-independent properties, and a checksum of checksums -- the whole set
    in one launch equals the same set simulated in eight slices with trial offsets."""
    import bench
    free, _ = torch.cuda.mem_get_info()
    n = 100_000_000 if free > 60e9 else 1 << 25
    dev = torch.device("cuda", torch.cuda.current_device())
    z = bench.build_workload(n, 0, dev)
    x = torch.empty((n, 2), device=dev)
    steps, stats = simulate_trials(z[:, :5], z[:, 5:], seed=77, out=x, return_steps=True, return_stats=True)[1:]
    assert bool(torch.isfinite(x).all())
    ch = x[:, 1]
    assert bool(((ch == 0) | (ch == 1) | (ch == 2)).all())
    assert float(x[:, 0].min()) >= 1e-6 and float(x[:, 0].max()) <= 8.0
    assert stats.useful_steps == int(steps.to(torch.int64).sum()) and stats.generic_rows == 0
    t_nd = z[:, 4].clamp(0.0, float(np.float32(8.0 - 1e-6)))
    rt = (t_nd + steps.to(torch.float32) * float(np.float32(5e-4))).clamp(1e-6, 8.0)
    assert torch.equal(rt, x[:, 0])                                        # reference :218, exactly
    del rt, t_nd
    frac = torch.bincount(ch.to(torch.int64), minlength=3).double() / n
    assert abs(frac[0] - 0.5651) < 2e-3 and abs(frac[1] - 0.2558) < 2e-3 and abs(frac[2] - 0.1791) < 2e-3
    assert abs(steps.double().mean().item() - 5116) < 10
    assert stats.lane_efficiency > (0.99 if n == 100_000_000 else 0.97)
    whole = int(x.view(torch.int32).to(torch.int64).sum())
    parts = 0
    xs = torch.empty((n // 8 + 8, 2), device=dev)
    for k in range(8):
        a, b = n * k // 8, n * (k + 1) // 8
        simulate_trials(z[a:b, :5], z[a:b, 5:], seed=77, trial_offset=a, out=xs[:b - a])
        parts += int(xs[:b - a].view(torch.int32).to(torch.int64).sum())
    assert parts == whole


def test_long_schedule_dt_1e_4_shared_noise_and_native():
    """configs[2]'s long pulse schedule: dt = 1e-4 -> n_max = 80000, steps_per_pulse = 1000, P = 80."""
    sched = Schedule.from_constants(dt=1e-4)
    assert (sched.n_max, sched.steps_per_pulse, sched.n_pulses) == (80000, 1000, 80)
    n = 192
    theta = orc.prior_sample(n, seed=51).numpy()
    pulses = orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(12)), 0, n, 80, 0.75)
    noise = orc.synthetic_noise(321, sched.n_max, n)
    want, want_steps = orc.sim_scalar_c(theta, pulses, noise, dt=1e-4)
    x, steps = simulate_trials(torch.from_numpy(theta), torch.from_numpy(pulses), noise=torch.from_numpy(noise),
                               schedule=sched, return_steps=True)
    _assert_same(x, want, "dt=1e-4 shared noise")
    assert np.array_equal(steps.cpu().numpy().astype(np.int64), want_steps)
    # native noise on the same schedule: replay through the oracle
    xn = simulate_trials(torch.from_numpy(theta), torch.from_numpy(pulses), seed=9, schedule=sched)
    normals = sim.philox_normals(9, n, sched.n_max)
    want_n, _ = orc.sim_scalar_c(theta, pulses, normals.cpu().numpy(), dt=1e-4)
    _assert_same(xn, want_n, "dt=1e-4 native replay")


def test_stress_schedule_dt_1e_6_native_replay():
    """SURVEY 8d stress variant: dt = 1e-6 (the unused constants.py:2 value) -> n_max = 8e6 steps,
    steps_per_pulse = 1e5.  Native noise, replayed through the C oracle with the dumped normals;
    censored trials run all 8e6 steps (step counters, Philox block index and kick schedule far from
    the default's ranges)."""
    sched = Schedule.from_constants(dt=1e-6)
    assert sched.steps_per_pulse == 100000 and sched.n_pulses == 80 and sched.n_max >= 7999999
    n = 12
    theta = orc.prior_sample(n, seed=71).numpy()
    theta[0] = [0.5, 0.0, 0.0, 1e4, 0.1]          # never reaches a bound: censored after the whole window
    theta[1] = [0.5, 0.3, 2.0, 30.0, 7.9999]      # window of a few steps
    pulses = orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(16)), 0, n, 80, 0.75)
    x, steps = simulate_trials(torch.from_numpy(theta), torch.from_numpy(pulses), seed=31, schedule=sched,
                               return_steps=True)
    normals = sim.philox_normals(31, n, sched.n_max)
    want, want_steps = orc.sim_scalar_c(theta, pulses, normals.cpu().numpy(), dt=1e-6)
    del normals
    _assert_same(x, want, "dt=1e-6 native replay")
    assert np.array_equal(steps.cpu().numpy().astype(np.int64), want_steps)
    assert want[0, 1] == 2.0 and want_steps[0] >= 7.8e6


def test_streaming_host_pipeline_equals_resident_launch(monkeypatch):
    """Host-resident z through one persistent streaming kernel per batch == ddm_sim_f32 on resident z:
    fp32 rows over the link (ddm_sim_stream_f32) and host-packed 32-byte records (ddm_pack_z_host +
    ddm_sim_packed_f32), bit for bit."""
    from sbi_for_diffusion_models_b200.simulator import HostPipeline
    n = 300000 + 17
    z = torch.empty((n, 85))
    z[:, :5] = orc.prior_sample(n, seed=61)
    z[:, 5:] = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(13)), 0, n, 80, 0.75))
    zh = z.pin_memory()
    sched = Schedule.from_constants()
    want = simulate_trials(z[:, :5], z[:, 5:], seed=4242, trial_offset=1000).cpu()
    from sbi_for_diffusion_models_b200 import simulator as simmod
    auto_packs = simmod.pack_threads() >= simmod.pack_min_threads()
    for packed in (False, True, None):
        xh = torch.full((n, 2), -7.0).pin_memory()
        pipe = HostPipeline(85, max_batch=1 << 17, chunk=1 << 14)      # 3 batches, 8 chunks each, ragged tail
        pipe.run(zh, xh, sched=sched, seed=4242, trial_offset=1000, packed=packed)
        pipe.synchronize()
        assert torch.equal(xh, want), packed
        assert pipe.launches == 3
        is_packed = packed is True or (packed is None and auto_packs)
        assert pipe.packed_batches == (3 if is_packed else 0)
        assert pipe.h2d_bytes == n * (32 if is_packed else 340)
    # the same pipeline object again (slots, staging blocks and events are reused); this time in the
    # order used under tools that serialise launches (Nsight Compute, CUDA_LAUNCH_BLOCKING=1)
    monkeypatch.setenv("DDM_INGEST_ORDER", "copies_first")
    assert simmod.launches_block()
    xh2 = torch.empty((n, 2)).pin_memory()
    pipe.run(zh, xh2, sched=sched, seed=4242, trial_offset=1000, packed=True)
    pipe.synchronize()
    assert torch.equal(xh2, want)


def test_launch_blocking_probe_and_starved_launch(monkeypatch):
    """The ingest order is decided by a probe, not by sniffing the environment: a one-thread kernel waits for a
    word the host raises once the launch call has returned.  Here launches do not block; in a child process with
    CUDA_LAUNCH_BLOCKING=1 they do.  And a streaming launch whose rows never arrive (copies enqueued too late, for
    whatever reason) gives up after the timeout, is reported when its slot is checked, and switches the process to
    copies-first."""
    import ctypes
    import subprocess
    import sys
    from sbi_for_diffusion_models_b200 import _native
    from sbi_for_diffusion_models_b200 import simulator as simmod
    L = _native.lib()
    out = ctypes.c_int(-1)
    assert L.ddm_probe_launch_blocking(50_000, ctypes.byref(out)) == 0 and out.value == 0
    monkeypatch.delenv("DDM_INGEST_ORDER", raising=False)
    monkeypatch.setattr(simmod, "_launches_block", None)
    assert simmod.launches_block() is False
    code = ("import ctypes; from sbi_for_diffusion_models_b200 import _native; o = ctypes.c_int(-1); "
            "assert _native.lib().ddm_probe_launch_blocking(50000, ctypes.byref(o)) == 0; print('blocking', o.value)")
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert "blocking 1" in res.stdout, res.stdout + res.stderr
    # a starved launch: `ready` is never raised; 50 ms timeout instead of the default 20 s
    n = 4096
    z = torch.cat([orc.prior_sample(n, seed=5), torch.ones(n, 80)], 1).pin_memory()
    xh = torch.empty((n, 2)).pin_memory()
    pipe = simmod.HostPipeline(85, max_batch=1 << 12, chunk=1 << 12)
    assert L.ddm_sim_set_stream_timeout_us(50_000) == 0
    try:
        real = L.ddm_ingest_packed
        monkeypatch.setattr(L, "ddm_ingest_packed", lambda *a: 0)      # the copies "never" get enqueued
        pipe.run(z, xh, sched=Schedule.from_constants(), seed=1, packed=True)
        with pytest.raises(RuntimeError, match="timed out waiting"):
            pipe.synchronize()
        assert simmod.launches_block() is True                         # copies first from now on
        monkeypatch.setattr(L, "ddm_ingest_packed", real)
        pipe.run(z, xh, sched=Schedule.from_constants(), seed=1, packed=True)
        pipe.synchronize()
        assert torch.equal(xh, simulate_trials(z[:, :5], z[:, 5:], seed=1).cpu())
    finally:
        assert L.ddm_sim_set_stream_timeout_us(20_000_000) == 0
        monkeypatch.setattr(simmod, "_launches_block", None)


def test_small_batch_kernel_is_bit_identical_to_the_throughput_kernel():
    """Launches of <= 8192 trials run producer / consumer CTAs (one warp integrates 32 trials in lock-step, seven
    warps generate the noise of the next 42 steps into shared memory); larger ones the persistent lane-per-trial
    kernel.  Same Philox indexing, same operations: identical bits, hit steps and step counts -- default, long and
    unaligned schedules, broadcast and non-binary pulses, log rt, ragged sizes."""
    from sbi_for_diffusion_models_b200 import _native
    L = _native.lib()
    rs = np.random.RandomState(3)
    cases = [(1, {}), (31, {}), (33, {}), (1000, {}), (10007, {}), (40000, {}),
             (3000, {"dt": 1e-4}), (2000, {"dt": 7e-4, "pulse_interval": 0.0497}), (500, {"t_max": 1.0}),
             # several kicks per 24-step chunk, sign masks (63 pulses) and memory pulses (267 pulses > 96 mask bits)
             (1500, {"dt": 1e-3, "pulse_interval": 0.008, "t_max": 0.5}), (700, {"dt": 2.5e-3, "pulse_interval": 0.03}),
             (900, {"dt": 1e-3, "pulse_interval": 0.001, "t_max": 0.09})]
    try:
        for n, kw in cases:
            sched = Schedule.from_constants(1.0, **kw)
            theta = orc.prior_sample(n, seed=70 + n)
            theta[: min(n, 4)] = torch.tensor([[0.5, 0.2, 0.3, 1e-7, 0.1], [0.0, -0.5, 2.0, 5.0, 7.9999995], [1.0, 3.0, 0.0, 30.0, 0.0],
                                               [0.3, 0.7, 1.0, 12.0, 9.0]])[: min(n, 4)]
            pulses = torch.from_numpy(np.where(rs.rand(n, sched.n_pulses) < 0.75, 1.0, -1.0).astype(np.float32))
            variants = [("rows", pulses, False), ("log rt", pulses, True), ("broadcast", pulses[:1], False)]
            odd = pulses.clone()
            odd[::7, ::5] = 0.25
            variants.append(("non-binary", odd, False))
            for name, pl, log_rt in variants:
                out = {}
                for mode, cap in (("small", 1 << 20), ("throughput", 0)):
                    assert L.ddm_sim_set_small_batch_max(cap) == 0
                    out[mode] = sim.simulate_trials(theta, pl, seed=5 + n, trial_offset=11, schedule=sched, log_rt=log_rt,
                                                    return_steps=True, return_stats=True)
                (xa, sa, ta), (xb, sb, tb) = out["small"], out["throughput"]
                assert torch.equal(xa.view(torch.int32), xb.view(torch.int32)), (n, kw, name)
                assert torch.equal(sa, sb) and ta.useful_steps == tb.useful_steps, (n, kw, name)
    finally:
        assert L.ddm_sim_set_small_batch_max(8192) == 0


def test_packed_ingest_falls_back_for_non_binary_pulses():
    """A batch with a pulse value other than +-1 cannot be packed to sign bits: it is re-run through
    the fp32 rows (reference semantics a += v * s, rt_choice_model.py:192)."""
    from sbi_for_diffusion_models_b200.simulator import HostPipeline
    n = 70000
    z = torch.empty((n, 85))
    z[:, :5] = orc.prior_sample(n, seed=62)
    z[:, 5:] = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(14)), 0, n, 80, 0.75))
    z[40000, 5] = 0.25          # first pulse of one trial in the second batch
    z[40001, 5 + 3] = -2.0
    zh, xh = z.pin_memory(), torch.empty((n, 2)).pin_memory()
    pipe = HostPipeline(85, max_batch=1 << 15, chunk=1 << 13)
    pipe.run(zh, xh, sched=Schedule.from_constants(), seed=99, packed=True)
    pipe.synchronize()
    want = simulate_trials(z[:, :5], z[:, 5:], seed=99).cpu()
    assert torch.equal(xh, want)
    assert pipe.launches == 3 + 1 and pipe.packed_batches == 3


def test_packed_kernel_on_resident_records():
    """ddm_sim_packed_f32 without the streaming flag (records already in HBM), incl. hit_step."""
    import ctypes
    from sbi_for_diffusion_models_b200 import _native
    n = 5000
    theta = orc.prior_sample(n, seed=63)
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(15)), 0, n, 80, 0.75))
    z = torch.cat([theta, pulses], 1).contiguous()
    rec = torch.empty((n, 8), dtype=torch.int32)
    L = _native.lib()
    assert L.ddm_pack_z_host(z.data_ptr(), 85, n, 80, rec.data_ptr(), 2) == 0
    sched = Schedule.from_constants()
    recd = rec.cuda()
    x = torch.empty((n, 2), device="cuda")
    steps = torch.empty((n,), dtype=torch.int32, device="cuda")
    ws = torch.zeros(8, dtype=torch.int64, device="cuda")
    rc = L.ddm_sim_packed_f32(recd.data_ptr(), n, sched.n_max, sched.steps_per_pulse, sched.dt, sched.t_max, sched.t_nd_hi,
                              sched.noise_scale, ctypes.c_uint64(5), ctypes.c_uint64(17), 0, x.data_ptr(), steps.data_ptr(),
                              ws.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    want, want_steps = simulate_trials(theta, pulses, seed=5, trial_offset=17, return_steps=True)
    assert torch.equal(x, want) and torch.equal(steps, want_steps)
    assert L.ddm_sim_packed_f32(recd.data_ptr(), n, 16000, 100, sched.dt, sched.t_max, sched.t_nd_hi, sched.noise_scale,
                                ctypes.c_uint64(5), ctypes.c_uint64(0), 0, x.data_ptr(), None, ws.data_ptr(), None,
                                None) == _native.DDM_ERR_INVALID          # 160 pulses do not fit a record


def test_outputs_stay_in_bounds(monkeypatch):
    """Guard bands around the device outputs of the simulator, the Philox dumps and the pulse generator
    at ragged sizes (compute-sanitizer is not available on the GPU pool)."""
    from sbi_for_diffusion_models_b200.pulses import generate_pulse_matrix_device
    real_empty, bands, G = torch.empty, [], 2048

    def guarded(*size, **kw):
        dev, dt = kw.get("device"), kw.get("dtype")
        if dev is None or torch.device(dev).type != "cuda" or dt not in (torch.float32, torch.int32):
            return real_empty(*size, **kw)
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        n = int(np.prod(shape))
        big = torch.full((n + 2 * G,), 12345, dtype=dt, device=dev)
        bands.append((big, n))
        return big[G:G + n].view(shape)

    monkeypatch.setattr(torch, "empty", guarded)
    for n in (1, 31, 33, 1000, 70001):
        theta = orc.prior_sample(n, seed=n)
        pulses = generate_pulse_matrix_device(np.random.default_rng(n), n, 80, p_success=0.75)
        assert tuple(pulses.shape) == (n, 80)
        x, steps = sim.simulate_trials(theta, pulses, seed=3, return_steps=True)
        assert tuple(x.shape) == (n, 2) and bool(torch.isfinite(x).all())
        x = sim.simulate_trials(theta, pulses[:, :80].cpu().repeat(1, 2)[:, :100], seed=3, log_rt=True)   # P > 80, generic path
        assert bool(torch.isfinite(x).all())
    sim.philox_normals(5, 77, 13)
    sim.philox_words(5, 77, 13)
    torch.cuda.synchronize()
    assert len(bands) >= 20
    for big, n in bands:
        assert bool((big[:G] == 12345).all()) and bool((big[G + n:] == 12345).all()), n
