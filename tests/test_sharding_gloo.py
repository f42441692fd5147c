"""Host-side sharding logic on CPU with the gloo backend, world_size 2 and 3.  The simulator is
replaced by the CPU oracle (injected), so these tests cover ranges, stream jumps, offsets and the
gather -- everything but the kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ddm_oracle as orc
from sbi_for_diffusion_models_b200.sharding import all_gather_rows, gather_sbc, loglik_sum_sharded, shard_bounds


def test_shard_bounds_partition():
    for total in (0, 1, 7, 10, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


class OracleSim:
    """CPU stand-in for the CUDA simulator: deterministic in (z row, global trial index)."""

    @staticmethod
    def pulses(state, inc, first, n, P, p):
        return torch.from_numpy(orc.pulses_pcg64_c(state, inc, first, n, P, p))

    def __call__(self, z, *, P, mu_sensory, log_rt, seed, trial_offset):
        n = z.shape[0]
        out = np.empty((n, 2), np.float32)
        for i in range(n):   # per-trial stream keyed by the GLOBAL index, like Philox on the device
            x, _ = orc.sim_rng_c(z[i:i + 1, :5].numpy(), z[i:i + 1, 5:].numpy(), seed=seed * 1000003 + trial_offset + i)
            out[i] = x[0]
        return torch.from_numpy(out)


class Prior:
    def sample(self, shape):
        return torch.rand((shape[0], 5)) * torch.tensor([1.0, 1.0, 2.0, 20.0, 1.0]) + torch.tensor([0.0, 0.0, 0.0, 5.0, 0.0])


def _run_sharded():
    from sbi_for_diffusion_models_b200 import proposals
    from sbi_for_diffusion_models_b200.sharding import simulate_training_set_sharded
    torch.manual_seed(123)
    prop = proposals.ExtendedProposal(Prior(), proposals.PulseSequenceProposal(P=80, p_success=0.75, seed=4))
    z, x = simulate_training_set_sharded(prop, 23, 8, "cpu", mu_sensory=1.0, p_success=0.75, P=80, log_rt=False,
                                         seed=9, simulate=OracleSim())
    return z, x, prop.pulse_proposal.rng.random()


def _chains():
    return orc.prior_sample(7, seed=5)


def _spec_potential():
    """CPU stand-in for the device potential: the MNLE spec on a fixed 6-trial session."""
    from oracle import mnle_spec as ms
    p = ms.init_params(0)
    pulses = torch.from_numpy(orc.pulses_pcg64_c(*orc.pcg64_state(np.random.default_rng(123)), 0, 6, 80, 0.75))
    x, _ = orc.sim_rng_c(np.repeat(np.array([[0.45, 0.6, 1.3, 14.0, 0.25]], np.float32), 6, 0), pulses.numpy(), 7)
    return lambda th: ms.loglik_sum(p, th, torch.from_numpy(x), pulses).float()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, x, nxt = _run_sharded()
        lo, hi = shard_bounds(10, rank, world)
        th = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 5)
        rk = torch.arange(lo, hi, dtype=torch.int64)[:, None].repeat(1, 5) * 3
        th_all, rk_all = gather_sbc(th, rk, 10)
        ragged = all_gather_rows(torch.full((hi - lo, 2), float(rank)), 10)
        pot = loglik_sum_sharded(_spec_potential(), _chains(), None)
        few = loglik_sum_sharded(_spec_potential(), _chains()[:1], None)      # fewer chains than ranks
        q.put((rank, z.numpy(), x.numpy(), nxt, th_all.numpy(), rk_all.numpy(), ragged.numpy(), pot.numpy(), few.numpy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_single_process(world):
    z1, x1, nxt1 = _run_sharded()                     # world 1 (no process group)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pot1, few1 = _spec_potential()(_chains()).numpy(), _spec_potential()(_chains()[:1]).numpy()
    for rank, z, x, nxt, th_all, rk_all, ragged, pot, few in results:
        assert np.array_equal(z, z1.numpy()), f"rank {rank}: z differs from the single-process set"
        assert np.array_equal(x, x1.numpy()), f"rank {rank}: x differs"
        assert nxt == nxt1                            # host pulse generator advanced identically
        assert np.array_equal(th_all[:, 0], np.arange(10, dtype=np.float32))
        assert np.array_equal(rk_all[:, 0], np.arange(10) * 3) and rk_all.dtype == np.int64
        want = np.concatenate([np.full(shard_bounds(10, r, world)[1] - shard_bounds(10, r, world)[0], float(r)) for r in range(world)])
        assert np.array_equal(ragged[:, 0], want)
        assert pot.shape == (7,) and np.allclose(pot, pot1, rtol=1e-6) and np.allclose(few, few1, rtol=1e-6)
