"""Device pulse generator and the reference-shaped Python API, on the GPU."""
import numpy as np
import pytest
import torch

from oracle import ddm_oracle as orc
from sbi_for_diffusion_models_b200 import data_simulator as ds
from sbi_for_diffusion_models_b200 import proposals
from sbi_for_diffusion_models_b200.models import rt_choice_model as rt
from sbi_for_diffusion_models_b200.pulses import generate_pulse_matrix_device, pulses_from_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 123, 2**31 - 2])
def test_pulses_match_reference_fixture(golden, seed):
    g = golden("pulses_pcg64")
    got = rt.generate_pulse_matrix_numpy(np.random.default_rng(seed), 64, 80, p_success=0.75)
    assert got.dtype == np.float32 and np.array_equal(got, g[f"seed{seed}"].astype(np.float32))


@pytest.mark.parametrize("p", [0.0, 0.5, 1.0, 0.3])
def test_pulses_probabilities(golden, p):
    g = golden("pulses_pcg64")
    got = rt.generate_pulse_matrix_numpy(np.random.default_rng(5), 32, 80, p_success=p)
    assert np.array_equal(got, g[f"p{p}"].astype(np.float32))


def test_stream_continues_like_numpy(golden):
    g = golden("pulses_pcg64")
    rng = np.random.default_rng(21)
    for key, (n, P) in (("stream_a", (10, 80)), ("stream_b", (7, 33)), ("stream_c", (5, 80))):
        assert np.array_equal(rt.generate_pulse_matrix_numpy(rng, n, P, p_success=0.75), g[key].astype(np.float32))
    ref = np.random.default_rng(21)
    orc.pulses_loop_numpy(ref, 10, 80, 0.75), orc.pulses_loop_numpy(ref, 7, 33, 0.75), orc.pulses_loop_numpy(ref, 5, 80, 0.75)
    assert rng.random() == ref.random()         # the host generator was advanced exactly


def test_large_matrix_against_numpy_and_row_ranges():
    rng = np.random.default_rng(99)
    state, inc = orc.pcg64_state(rng)
    n = 50000
    want = rng.random((n, 81))                  # stream-equivalent to the per-trial loop
    side = np.where(want[:, :1] < 0.5, 1.0, -1.0)
    want = np.where(want[:, 1:] < 0.75, side, -side).astype(np.float32)
    got = pulses_from_state(state, inc, 0, n, 80, 0.75)
    assert np.array_equal(got.cpu().numpy(), want)
    part = pulses_from_state(state, inc, 31000, 999, 80, 0.75)
    assert np.array_equal(part.cpu().numpy(), want[31000:31999])
    z = torch.zeros((n, 85), device="cuda")
    pulses_from_state(state, inc, 0, n, 80, 0.75, out=z[:, 5:])
    assert np.array_equal(z[:, 5:].cpu().numpy(), want) and float(z[:, :5].abs().sum()) == 0.0
    for P in (1, 31, 32, 33, 96, 200):
        r = np.random.default_rng(P)
        st, ic = orc.pcg64_state(r)
        assert np.array_equal(pulses_from_state(st, ic, 0, 77, P, 0.75).cpu().numpy(),
                              orc.pulses_loop_numpy(r, 77, P, 0.75))


def test_edge_sizes_and_errors():
    assert rt.generate_pulse_matrix_numpy(np.random.default_rng(0), 0, 80).shape == (0, 80)
    rng = np.random.default_rng(0)
    assert rt.generate_pulse_matrix_numpy(rng, 5, 0).shape == (5, 0)
    assert rng.random() == np.random.default_rng(0).random()      # nothing was drawn
    with pytest.raises(ValueError, match="n_trials"):
        rt.generate_pulse_matrix_numpy(np.random.default_rng(0), -1, 80)
    with pytest.raises(ValueError, match="n_pulses"):
        rt.generate_pulse_matrix_numpy(np.random.default_rng(0), 1, -1)
    with pytest.raises(TypeError, match="PCG64"):
        generate_pulse_matrix_device(np.random.Generator(np.random.MT19937(0)), 4, 80, p_success=0.75)


def test_proposals_match_reference_fixture(golden):
    g = golden("pulses_pcg64")
    prop = proposals.PulseSequenceProposal(P=80, p_success=0.75, seed=0)
    a = prop.sample((6,))
    b = prop.sample((3, 2))
    c = prop.sample()
    assert a.device.type == "cpu" and a.dtype == torch.float32
    assert np.array_equal(a.numpy(), g["proposal_first"].astype(np.float32))
    assert tuple(b.shape) == (3, 2, 80) and np.array_equal(b.numpy(), g["proposal_second"].astype(np.float32))
    assert tuple(c.shape) == (1, 80) and np.array_equal(c.numpy(), g["proposal_scalar"].astype(np.float32))
    assert float(prop.log_prob(a).abs().sum()) == 0.0 and tuple(prop.log_prob(b).shape) == (3, 2)


class _PriorStub:
    def __init__(self):
        self.k = 0

    def sample(self, shape=torch.Size()):
        n = int(np.prod(shape)) if len(shape) else 1
        self.k += 1
        return orc.prior_sample(n, seed=100 + self.k)

    def log_prob(self, th):
        return torch.zeros(th.shape[:-1])


def test_training_set_shell_matches_reference(golden, capsys):
    g = golden("training_set")
    P = 80
    for device in (None, "cuda"):
        prop = proposals.ExtendedProposal(_PriorStub(), proposals.PulseSequenceProposal(P=P, p_success=0.75, seed=0),
                                          device=device)
        z_all, x_all = ds.simulate_training_set_with_conditions(prop, 300, 128, "cpu", mu_sensory=1.0, p_success=0.75,
                                                                P=P, log_rt=False, seed=11)
        assert z_all.device.type == "cpu" and x_all.device.type == "cpu"
        assert np.array_equal(z_all.numpy(), g["z"])        # same prior draws + same PCG64 pulse stream
        assert tuple(x_all.shape) == (300, 2)
    out = capsys.readouterr().out
    assert "Simulated 128/300" in out and "Unique outcomes in training (choice)" in out
    # x given z: replay each reference batch with its shared noise through sim_wrapper
    start = 0
    z = torch.from_numpy(g["z"])
    for seed, bs in zip(g["noise_seeds"], (128, 128, 44)):
        noise = torch.from_numpy(orc.synthetic_noise(int(seed), 16000, bs))
        x = ds.sim_wrapper(z[start:start + bs], mu_sensory=1.0, p_success=0.75, P=P, log_rt=False, noise=noise)
        assert x.device.type == "cpu"
        assert np.array_equal(x.numpy().view(np.uint32), g["x"][start:start + bs].view(np.uint32))
        start += bs
    # batch size does not change x for a fixed z / seed
    a = ds.sim_wrapper(z, mu_sensory=1.0, p_success=0.75, P=P, log_rt=False, seed=5)
    b = torch.cat([ds.sim_wrapper(z[:100], mu_sensory=1.0, p_success=0.75, P=P, log_rt=False, seed=5),
                   ds.sim_wrapper(z[100:], mu_sensory=1.0, p_success=0.75, P=P, log_rt=False, seed=5,
                                  trial_offset=100)])
    assert torch.equal(a, b)


def test_sessions_match_reference(golden):
    g = golden("sessions")
    th = torch.from_numpy(g["theta_true"])
    x_o, pulses_o = ds.simulate_observed_session(th, 50, "cpu", mu_sensory=1.0, p_success=0.75, P=80, seed=123,
                                                 log_rt=False, noise=torch.from_numpy(orc.synthetic_noise(30, 16000, 50)))
    assert x_o.device.type == "cpu" and pulses_o.device.type == "cpu"
    assert np.array_equal(pulses_o.numpy(), g["pulses_o"].astype(np.float32))
    assert np.array_equal(x_o.numpy().view(np.uint32), g["x_o"].view(np.uint32))
    xs, ss = rt.simulate_session_data_rt_choice(th, 40, rng=np.random.default_rng(77), mu_sensory=1.0, p_success=0.75,
                                                return_pulse_sides=True,
                                                noise=torch.from_numpy(orc.synthetic_noise(32, 16000, 40)))
    assert np.array_equal(ss.cpu().numpy(), g["sess_pulses"].astype(np.float32))
    assert np.array_equal(xs.cpu().numpy().view(np.uint32), g["sess_x"].view(np.uint32))


def test_single_trial_numpy_api_and_pack():
    rt_val, choice = rt.rt_choice_model_simulator(np.array([0.5, 0.3, 1.0, 12.0, 0.2]), np.random.default_rng(0))
    assert 1e-6 <= rt_val <= 8.0 and choice in (0, 1, 2)
    x = torch.tensor([[0.5, 1.0], [1e-9, 0.0], [8.0, 2.0]])
    assert torch.equal(rt.pack_x_rt_choice(x, log_rt=False), torch.tensor([[0.5, 1.0], [1e-6, 0.0], [8.0, 2.0]]))
    assert torch.allclose(rt.pack_x_rt_choice(x, log_rt=True)[:, 0], torch.log(torch.tensor([0.5, 1e-6, 8.0])))
    assert rt.pulse_schedule() == (16000, 200) and rt.n_pulses_max_from_schedule(16000, 200) == 80


def test_cuda_prior_draws_follow_the_prior():
    """A prior whose parameters live on the GPU draws Beta(2, 2) as an order statistic of uniforms (exact law,
    ~50x cheaper than torch's CUDA gamma rejection sampler); the other components and log_prob are torch's."""
    from scipy import stats
    from sbi_for_diffusion_models_b200.priors import build_prior_theta
    torch.manual_seed(3)
    prior = build_prior_theta("cuda")
    th = prior.sample((400_000,))
    assert th.is_cuda and tuple(th.shape) == (400_000, 5)
    t = th.cpu().double().numpy()
    laws = [stats.beta(2, 2), stats.lognorm(s=1.0, scale=np.exp(-1.0)), stats.lognorm(s=1.0, scale=1.0),
            stats.lognorm(s=0.5, scale=np.exp(2.75)), stats.beta(2, 2)]
    for i, law in enumerate(laws):
        assert stats.kstest(t[:, i], law.cdf).pvalue > 1e-4, i
    assert bool(torch.isfinite(prior.log_prob(th)).all())
    assert tuple(prior.sample(()).shape) == (5,) and tuple(prior.sample((3, 2)).shape) == (3, 2, 5)


def test_training_set_comes_home_as_records_or_rows_with_the_same_bits(monkeypatch):
    """Large sets: z crosses the link as 32-byte records packed by the GPU (ddm_pack_z_dev) and is rebuilt by the host
    cores (ddm_unpack_z_host); DDM_TRAINSET_D2H=rows copies the fp32 rows.  Same (z, x), also when a block holds a
    pulse value other than +-1 (that block is copied as it is)."""
    import contextlib
    import io
    from sbi_for_diffusion_models_b200 import data_simulator as ds
    from sbi_for_diffusion_models_b200.priors import build_prior_theta
    from sbi_for_diffusion_models_b200.proposals import ExtendedProposal, PulseSequenceProposal

    class Odd(ExtendedProposal):
        """every 7th call plants a pulse value that is not +-1"""
        calls = 0

        def sample(self, shape=torch.Size()):
            z = super().sample(shape)
            Odd.calls += 1
            if Odd.calls % 7 == 0:
                z[-1, 9] = 0.5
            return z

    def run(mode, cls, n, bs):
        monkeypatch.setenv("DDM_TRAINSET_D2H", mode)
        monkeypatch.setattr(ds, "_LAUNCH_ROWS", 40_000)                    # several blocks, ragged tail
        Odd.calls = 0
        torch.manual_seed(11)
        prop = cls(build_prior_theta("cuda"), PulseSequenceProposal(80, 0.75, seed=2, device="cuda"), device="cuda")
        with contextlib.redirect_stdout(io.StringIO()):
            return ds.simulate_training_set_with_conditions(prop, n, bs, "cuda", mu_sensory=1.0, p_success=0.75, P=80,
                                                            log_rt=False, seed=5)
    for cls in (ExtendedProposal, Odd):
        za, xa = run("records", cls, 130_001, 10_000)
        zb, xb = run("rows", cls, 130_001, 10_000)
        assert torch.equal(za, zb) and torch.equal(xa, xb), cls.__name__
        assert bool(((za[:, 5:].abs() == 1) | (za[:, 5:] == 0.5)).all())
    assert int((za[:, 5:] == 0.5).sum()) == 2
