"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares,
validates arguments before touching CUDA, and its host-only helper matches NumPy."""
import ctypes
import os
import re

import numpy as np
import pytest

from sbi_for_diffusion_models_b200 import _native
from sbi_for_diffusion_models_b200.build import build_native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build_native()
    return _native.lib()


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        text = open(os.path.join(inc, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b((?:ddm|mnle)_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_header_symbols_are_exported(L):
    declared = _declared()
    assert "ddm_sim_f32" in declared and "ddm_pulses_pcg64" in declared
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert set(_native.exported_symbols()) == set(declared)


def test_abi_version(L):
    assert L.ddm_abi_version() == 1
    assert L.ddm_sim_workspace_bytes() == 8 * _native.WS_WORDS


def test_invalid_arguments_are_rejected_before_cuda(L):
    ws = ctypes.create_string_buffer(64)
    args = dict(theta=1, ld_theta=5, pulses=1, ld_pulses=80, N=4, P=80, n_max=16000, spp=200)

    def call(**kw):
        a = dict(args, **kw)
        return L.ddm_sim_f32(a["theta"], a["ld_theta"], a["pulses"], a["ld_pulses"], a["N"], a["P"], a["n_max"],
                             a["spp"], 5e-4, 8.0, 7.999999, 0.0223, 0, 0, None, 0, 0, 1 << 12, None,
                             ctypes.addressof(ws) & ~7, None)

    assert call(P=79) == _native.DDM_ERR_INVALID          # reference rt_choice_model.py:173-176
    assert b"P=79" in L.ddm_last_error()
    assert call(N=-1) == _native.DDM_ERR_INVALID
    assert call(spp=0) == _native.DDM_ERR_INVALID
    assert call(ld_theta=3) == _native.DDM_ERR_INVALID
    assert call(ld_pulses=10) == _native.DDM_ERR_INVALID
    assert L.ddm_pulses_pcg64(0, 0, 0, 1, 0, -1, 80, 1, None, 80, None) == _native.DDM_ERR_INVALID
    assert b"n_trials" in L.ddm_last_error()             # reference rt_choice_model.py:83-84
    assert L.ddm_pulses_pcg64(0, 0, 0, 1, 0, 4, -2, 1, None, 80, None) == _native.DDM_ERR_INVALID
    with pytest.raises(ValueError):
        _native.check(_native.DDM_ERR_INVALID, "x")


def test_host_pcg64_advance_matches_numpy(L):
    rng = np.random.default_rng(2024)
    st = rng.bit_generator.state["state"]
    state, inc = int(st["state"]), int(st["inc"])
    m = (1 << 64) - 1
    for draws in (0, 1, 81, 10**6 + 7, 2**40 + 3):
        hi, lo = ctypes.c_uint64(state >> 64), ctypes.c_uint64(state & m)
        assert L.ddm_pcg64_advance(ctypes.byref(hi), ctypes.byref(lo), inc >> 64, inc & m, draws) == 0
        ref = np.random.default_rng(2024)
        ref.bit_generator.advance(draws)
        assert ((hi.value << 64) | lo.value) == int(ref.bit_generator.state["state"]["state"])


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sbi_for_diffusion_models_b200.simulator import simulate_trials
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        simulate_trials(torch.zeros(2, 5), torch.ones(2, 80))


def test_product_never_imports_oracle():
    """The checker must not be reachable from the product path (no import, include or dlopen)."""
    pkg = os.path.join(ROOT, "sbi_for_diffusion_models_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include[^\n]*oracle|libddm_oracle|ddm_oracle_", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f"{f} reaches into oracle/"


def test_host_packer_matches_numpy(L):
    """ddm_pack_z_host (host code, no GPU): 32-byte records [theta bits x 5, pulse sign masks x 3]."""
    rs = np.random.RandomState(0)
    N, ld = 40003, 90
    z = np.empty((N, ld), np.float32)
    z[:, :5] = rs.randn(N, 5)
    z[:, 5:] = np.where(rs.rand(N, ld - 5) < 0.5, 1.0, -1.0)
    out = np.zeros((N, 8), np.uint32)

    def expect(n_pulses):
        m = np.zeros((N, 3), np.uint64)
        for c in range(96):
            bit = (z[:, 5 + c] > 0) if c < n_pulses else np.ones(N, bool)
            m[:, c // 32] |= bit.astype(np.uint64) << np.uint64(c % 32)
        return m.astype(np.uint32)

    for n_pulses, threads in ((80, 1), (80, 5), (0, 2), (1, 2), (31, 3), (32, 3), (33, 1), (64, 4), (79, 2), (85, 16)):
        assert L.ddm_pack_z_host(z.ctypes.data, ld, N, n_pulses, out.ctypes.data, threads) == 0
        assert np.array_equal(out[:, :5].view(np.float32), z[:, :5])
        assert np.array_equal(out[:, 5:], expect(n_pulses)), (n_pulses, threads)
    z[7, 5 + 33], z[9, 5 + 79], z[11, 5 + 81], z[12, 5] = 0.5, -1.0000001, 3.0, np.nan
    assert L.ddm_pack_z_host(z.ctypes.data, ld, N, 80, out.ctypes.data, 4) == 3        # column 81 is past the schedule
    assert L.ddm_pack_z_host(z.ctypes.data, ld, 0, 80, out.ctypes.data, 4) == 0
    assert L.ddm_pack_z_host(z.ctypes.data, ld, N, 97, out.ctypes.data, 4) == _native.DDM_ERR_INVALID
    assert L.ddm_pack_z_host(z.ctypes.data, 60, N, 80, out.ctypes.data, 4) == _native.DDM_ERR_INVALID


def test_host_unpacker_inverts_the_packer(L):
    """ddm_unpack_z_host (host code, no GPU): records [theta bits x 5, sign masks x 3] -> fp32 rows, every alignment
    and thread count, whole 16-row blocks (non-temporal stores) and ragged tails."""
    rs = np.random.RandomState(1)
    for N, P in ((1, 80), (15, 80), (16, 80), (1000, 80), (70001, 80), (333, 7), (4097, 96), (100, 0)):
        z = np.empty((N, 5 + P), np.float32)
        z[:, :5] = rs.randn(N, 5)
        z[:, 5:] = np.where(rs.rand(N, P) < 0.5, 1.0, -1.0)
        rec = np.empty((N, 8), np.uint32)
        assert L.ddm_pack_z_host(z.ctypes.data, 5 + P, N, P, rec.ctypes.data, 4) == 0
        buf = np.zeros(N * (5 + P) + 64, np.float32)
        off = (-buf.ctypes.data // 4) % 16
        for shift in (off, off + 1):                     # 64-byte aligned (fast path) and not
            out = buf[shift:shift + N * (5 + P)].reshape(N, 5 + P)
            for threads in (1, 5):
                out[:] = 7.0
                assert L.ddm_unpack_z_host(rec.ctypes.data, N, P, out.ctypes.data, 5 + P, threads) == 0
                assert np.array_equal(out.view(np.uint32), z.view(np.uint32)), (N, P, shift - off, threads)
    assert L.ddm_unpack_z_host(rec.ctypes.data, 10, 97, buf.ctypes.data, 200, 1) == _native.DDM_ERR_INVALID
