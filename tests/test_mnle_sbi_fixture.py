"""The MNLE path against REAL sbi numbers -- when someone has produced them.

sbi 0.25.0 cannot be installed in this repository's build environment, so the MNLE arithmetic is restated from
the published algorithms (oracle/mnle_spec.py) and its parity with sbi is unpinned.  ``tools/compare_with_sbi.py``
(run where sbi is installed) writes ``tests/golden/mnle_sbi.npz``: an sbi estimator's ``state_dict``, held-out
rows and sbi's own ``log_prob``.  With that file present these tests pin the state_dict import, the CPU spec and
the CUDA kernels to sbi; without it they are skipped and say why."""
import os

import numpy as np
import pytest
import torch

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mnle_sbi.npz")
needs_fixture = pytest.mark.skipif(not os.path.exists(FIXTURE),
                                   reason="tests/golden/mnle_sbi.npz absent: MNLE parity with sbi 0.25.0 is UNPINNED "
                                          "(run tools/compare_with_sbi.py where sbi is installed and commit its output)")


def _load():
    d = np.load(FIXTURE, allow_pickle=False)
    sd = {str(k): torch.from_numpy(d[f"t{i}"]) for i, k in enumerate(d["keys"])}
    return sd, torch.from_numpy(d["x"]), torch.from_numpy(d["z"]), torch.from_numpy(d["log_prob"]).double()


def _spec_params(packed):
    from sbi_for_diffusion_models_b200.mnle_net import unpack_params
    p = {k: v.clone() for k, v in unpack_params(torch.from_numpy(packed.packed.copy()), packed.n_choices).items()}
    p["cond_mean"], p["cond_std"] = torch.zeros(85), torch.ones(85)
    return p


@needs_fixture
def test_cpu_spec_reproduces_sbi_log_probs():
    from oracle import mnle_spec as ms
    from sbi_for_diffusion_models_b200.mnle_net import PackedMNLE
    sd, x, z, want = _load()
    packed = PackedMNLE.from_state_dict(sd)
    got = ms.log_prob(ms.cast_params(_spec_params(packed), torch.float64), x, z)
    # sbi evaluates in fp32: its own numbers carry fp32 rounding (see tests/test_gpu_mnle.py on trained nets)
    assert float((got - want).abs().mean()) < 1e-3 and float((got - want).abs().max()) < 1e-1


@needs_fixture
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["precise", "simt", "tc"])
def test_cuda_kernels_reproduce_sbi_log_probs(kernel):
    from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
    sd, x, z, want = _load()
    est = DeviceMNLE(PackedMNLE.from_state_dict(sd))
    got = est.log_prob(x.unsqueeze(0), condition=z, kernel=kernel)[0].double()
    assert float((got - want).abs().mean()) < 1e-3 and float((got - want).abs().max()) < 1e-1
    # north_star: sums within 1e-4 relative (all held-out rows as one session)
    assert abs(float(got.sum() - want.sum())) < 1e-4 * float(want.abs().sum())
