"""Sharding of the hot path over the GPUs of one box: one process per GPU
(``torch.distributed``, NCCL over NVLink), contiguous global trial / dataset ranges per rank,
no data-path collective during simulation and ONE all-gather of the outputs at the end.

The reference is single-process: trials are independent
(rt_choice_model.py:112-221 has no cross-trial term) and SBC datasets are independent
(mnle.py:183-218).  Because the Philox counter is the GLOBAL trial index and the pulse stream is
jumped to the shard's first trial, the union of the shards is bit-identical to a single-process
run, for any world size.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks laid out by ``shard_bounds(total, r, world)`` in rank
    order.  Blocks are padded to the largest shard so one ``all_gather_into_tensor`` (NCCL) /
    ``all_gather`` (gloo) moves everything."""
    rank, world = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    tail = tuple(local.shape[1:])
    padded = local.new_zeros((biggest,) + tail)
    padded[: local.shape[0]] = local
    if local.is_cuda:
        flat = local.new_empty((world * biggest,) + tail)
        dist.all_gather_into_tensor(flat, padded, group=group)
        parts = list(flat.view((world, biggest) + tail).unbind(0))
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    return torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


class PeerGather:
    """The gathered output of a sharded simulation, filled by the simulator kernels themselves.

    Every rank owns ``x_all`` (sum of all shards, 2) fp32 in symmetric memory (``torch.distributed._symmetric_memory``:
    CUDA virtual-memory handles exchanged once, every rank's block mapped into every process over NVLink).  A
    rank simulates its shard with ``simulate_trials(..., out=pg.local, peer_blocks=pg.peers)``: each finished trial's
    8 bytes are stored to the local slot and to the same slot of every peer's ``x_all`` while the kernel keeps
    running (``ddm_sim_gather_f32``), so the path's only exchange costs no collective, no extra pass over
    x and no SMs.  ``barrier()`` (enqueued on the current stream) after the launch makes all slots of ``x_all``
    complete on every rank."""

    def __init__(self, shard_rows: int, group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = _world(group)
        self.rows = int(shard_rows)                     # equal shards: rank r owns rows [r * rows, (r + 1) * rows)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        shape = (self.world * self.rows, 2)
        self.x_all = symm.empty(shape, dtype=torch.float32, device=dev)
        self.handle = symm.rendezvous(self.x_all, group if group is not None else dist.group.WORLD)
        lo, hi = self.rank * self.rows, (self.rank + 1) * self.rows
        self.local = self.x_all[lo:hi]
        self.peers = [self.handle.get_buffer(r, shape, torch.float32)[lo:hi] for r in range(self.world) if r != self.rank]

    def barrier(self) -> None:
        self.handle.barrier(channel=0)


def _cuda_simulate(z, *, P, mu_sensory, log_rt, seed, trial_offset):
    from .simulator import simulate_trials
    return simulate_trials(z[:, :5], z[:, 5:5 + P], mu_sensory=mu_sensory, log_rt=log_rt, seed=seed,
                           trial_offset=trial_offset)


@torch.no_grad()
def simulate_training_set_sharded(proposal, num_simulations: int, batch_size: int, device, *, mu_sensory: float,
                                  p_success: float, P: int, log_rt: bool, seed: int, gather: bool = True,
                                  group=None, simulate: Optional[Callable] = None):
    """Sharded ``simulate_training_set_with_conditions`` (reference data_simulator.py:33-71).

    Every rank walks the reference's batch loop so that the theta prior's torch stream stays in
    lock-step on all ranks (seed torch identically everywhere), but only materialises pulses and
    simulates the rows of its own range [lo, hi).  The proposal must be an ``ExtendedProposal``
    over a ``PulseSequenceProposal`` (its PCG64 stream is jumped to each batch's first row).
    Returns (z, x) for the WHOLE set on every rank when ``gather`` else this rank's slice.
    """
    from .pulses import pcg64_state, pulses_from_state
    rank, world = _world(group)
    lo, hi = shard_bounds(num_simulations, rank, world)
    simulate = simulate or _cuda_simulate
    pp = proposal.pulse_proposal
    if pp.P != P:
        raise ValueError(f"proposal draws {pp.P} pulses per trial but P={P}")
    state, inc = pcg64_state(pp.rng)
    on_gpu = simulate is _cuda_simulate
    zs, xs = [], []
    for start in range(0, num_simulations, batch_size):
        bs = min(batch_size, num_simulations - start)
        theta = proposal.theta_prior.sample((bs,)).to(torch.float32)        # all ranks: keeps streams aligned
        a, b = max(start, lo), min(start + bs, hi)
        if a >= b:
            continue
        if on_gpu:
            z = torch.empty((b - a, 5 + P), dtype=torch.float32, device=device)
            z[:, :5] = theta[a - start:b - start].to(z.device, non_blocking=True)
            pulses_from_state(state, inc, a, b - a, P, pp.p_success, out=z[:, 5:])
        else:   # host-logic tests: the injected simulator also supplies the pulse rows
            z = torch.cat([theta[a - start:b - start], simulate.pulses(state, inc, a, b - a, P, pp.p_success)], dim=1)
        x = simulate(z, P=P, mu_sensory=mu_sensory, log_rt=log_rt, seed=seed, trial_offset=a)
        zs.append(z)
        xs.append(x)
    pp.rng.bit_generator.advance(num_simulations * (P + 1))                 # as if one process had drawn it all
    width = 5 + P
    ref = zs[0] if zs else torch.empty((0, width), device=device if on_gpu else "cpu")
    z_loc = torch.cat(zs) if zs else ref.new_empty((0, width))
    x_loc = torch.cat(xs) if xs else ref.new_empty((0, 2))
    if not gather:
        return z_loc, x_loc
    return all_gather_rows(z_loc, num_simulations, group), all_gather_rows(x_loc, num_simulations, group)


def gather_sbc(thetas_local: torch.Tensor, ranks_local: torch.Tensor, num_datasets: int, group=None):
    """All-gather the per-dataset results of an SBC run sharded with ``shard_bounds``:
    (thetas_true (D,5) f32, ranks (D,5) i64) in dataset order on every rank."""
    return (all_gather_rows(thetas_local, num_datasets, group), all_gather_rows(ranks_local, num_datasets, group))


def loglik_sum_sharded(loglik_sum: Callable[[torch.Tensor], torch.Tensor], theta: torch.Tensor, group=None) -> torch.Tensor:
    """MNLE potential with the chains split over the ranks (weights replicated, no exchange
    while computing, one all-gather of C floats).  ``loglik_sum(theta_rows) -> (n,)`` is the per-rank
    evaluation, e.g. ``lambda th: estimator.loglik_sum(th, x_o, pulses_o)``; every rank passes the same
    ``theta`` (C,5) and gets all C values back in chain order.  A chain's sum never leaves its rank, so
    the result equals the single-process one bit for bit."""
    rank, world = _world(group)
    C = theta.shape[0]
    lo, hi = shard_bounds(C, rank, world)
    local = loglik_sum(theta[lo:hi]) if hi > lo else theta.new_empty((0,), dtype=torch.float32)
    return all_gather_rows(local.reshape(-1), C, group)
