"""Training-set and observed-session shells (reference data_simulator.py:14-111) on top of
the CUDA simulator.  Signatures, (z, x) layout, CPU-resident return values, progress prints
and sanity asserts are the reference's; compute happens on the current CUDA device whatever
``device`` string the caller passes (the reference hard-codes "cpu" at its call sites).
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch
from torch.distributions import Distribution

from .pulses import generate_pulse_matrix_device
from . import _native
from .simulator import HostPipeline, Schedule, compute_device, next_seed, pack_threads, simulate_trials


_HOST_STREAM_MIN = 1 << 18   # CPU-resident z with at least this many rows is streamed in chunks
_HOST_BATCH_ROWS = 1 << 24   # rows per persistent streaming launch (every launch ends with the drain of its longest trials)
_pipelines = {}


def host_pipeline_counters() -> dict:
    """Totals over the host-streaming pipelines of this process: launches, packed batches and the bytes
    actually sent over the link (32 per trial with packed ingest, 4 * (5 + P) with fp32 rows)."""
    out = {"launches": 0, "packed_batches": 0, "h2d_bytes": 0}
    for pipe in _pipelines.values():
        out["launches"] += pipe.launches
        out["packed_batches"] += pipe.packed_batches
        out["h2d_bytes"] += pipe.h2d_bytes
    return out


def sim_wrapper(theta_and_pulses: torch.Tensor, *, mu_sensory: float, p_success: float, P: int, log_rt: bool,
                seed: Optional[int] = None, trial_offset: int = 0, noise: Optional[torch.Tensor] = None,
                ) -> torch.Tensor:
    """z = [theta(5), pulse_sides(P)] -> packed x = [rt (or log rt), choice] (reference :14-30).
    theta and pulses are passed to the kernel as strided views of z: no split copies.  A large
    CPU-resident z is streamed through the GPU in chunks with copies and kernels overlapped;
    x comes back on z's device, like the reference's."""
    z = theta_and_pulses
    n = z.shape[0]
    if (not z.is_cuda and noise is None and n >= _HOST_STREAM_MIN and z.dtype == torch.float32
            and z.shape[1] == 5 + P and z.stride(1) == 1):
        sched = Schedule.from_constants(mu_sensory)
        dev = compute_device(None)
        key = (dev, z.shape[1])
        pipe = _pipelines.get(key)
        if pipe is None:
            pipe = _pipelines[key] = HostPipeline(z.shape[1], max_batch=_HOST_BATCH_ROWS, device=dev)
        x = torch.empty((n, 2), dtype=torch.float32, pin_memory=True)
        packed = pipe.choose_packed(sched, n)
        t0 = time.perf_counter()
        pipe.run(z, x, sched=sched, seed=next_seed() if seed is None else seed, log_rt=log_rt,
                 trial_offset=trial_offset, packed=packed)
        pipe.synchronize()
        pipe.report(packed, n, time.perf_counter() - t0)
        return x
    x = simulate_trials(z[:, :5], z[:, 5:5 + P], mu_sensory=mu_sensory, log_rt=log_rt, seed=seed,
                        trial_offset=trial_offset, noise=noise)
    return x.to(z.device)


_LAUNCH_ROWS = 1 << 22   # rows simulated per launch by simulate_training_set_with_conditions
_copy_streams = {}


def _copy_stream(dev) -> "torch.cuda.Stream":
    s = _copy_streams.get(dev)
    if s is None:
        s = _copy_streams[dev] = torch.cuda.Stream(dev)
    return s


@torch.no_grad()
def simulate_training_set_with_conditions(proposal: Distribution, num_simulations: int, batch_size: int, device, *,
                                          mu_sensory: float, p_success: float, P: int, log_rt: bool,
                                          seed: Optional[int] = None):
    """Draw z ~ proposal in batches, simulate x | z, return CPU (z_all (N,5+P), x_all (N,2))
    (reference :33-71).  ``proposal.sample((bs,))`` is called batch by batch exactly as in the reference (so the
    proposal's random streams advance the same way), but the draws of up to 2^22 rows are gathered into one device
    block and simulated by ONE launch: with the reference's default batch of 4096 a launch per batch would drain
    to its longest trial (16 000 serial Euler steps) every time.  One Philox key covers the whole set and trials
    are indexed globally, so the result does not depend on ``batch_size`` or on this grouping for a given z."""
    dev = compute_device(device)
    if seed is None:
        seed = next_seed()
    # results go to pinned memory (asynchronous device->host copies on their own stream, overlapped with the next
    # block's draws and simulation) unless the whole set is small: pinning a few megabytes costs more than it saves
    pinned = num_simulations * (7 + P) * 4 >= (8 << 20)
    z_all = torch.empty((num_simulations, 5 + P), dtype=torch.float32, pin_memory=pinned)
    x_all = torch.empty((num_simulations, 2), dtype=torch.float32, pin_memory=pinned)
    stream = torch.cuda.current_stream(dev)
    copier = _copy_stream(dev) if pinned else stream
    # Large sets: z goes home as 32-byte records (theta bits + pulse sign masks, packed by the GPU) and the host cores
    # rebuild the fp32 rows while the GPU simulates the next block -- 40 instead of 348 bytes per trial over PCIe.
    # A block with a pulse value other than +-1 is copied as it is.
    as_records = pinned and P <= 96 and os.environ.get("DDM_TRAINSET_D2H", "records") == "records"
    L = _native.lib()
    n_threads = pack_threads()
    rec_host = torch.empty((num_simulations, 8), dtype=torch.int32, pin_memory=True) if as_records else None
    unpacker = ThreadPoolExecutor(max_workers=1) if as_records else None
    jobs = []
    in_flight = []       # (event, device blocks) until their device->host copies have run
    checks, outcomes = [], torch.zeros(3, dtype=torch.int64, device=dev)
    group = max(int(batch_size), _LAUNCH_ROWS)
    n_groups = -(-num_simulations // group)
    generic_host = torch.zeros((max(n_groups, 1),), dtype=torch.int64, pin_memory=True) if as_records else None

    def rebuild(g, g0, g1, ev, z_dev):
        """Worker thread: block g has landed -- rebuild its rows of z_all from the records (or fetch them as they are)."""
        ev.synchronize()
        if int(generic_host[g]) == 0:
            _native.check(L.ddm_unpack_z_host(rec_host[g0:g1].data_ptr(), g1 - g0, P, z_all[g0:g1].data_ptr(), 5 + P, n_threads),
                          "ddm_unpack_z_host")
        else:
            with torch.cuda.stream(copier):
                z_all[g0:g1].copy_(z_dev, non_blocking=True)
            copier.synchronize()

    for g, g0 in enumerate(range(0, num_simulations, group)):
        g1 = min(g0 + group, num_simulations)
        with torch.cuda.device(dev):
            z = torch.empty((g1 - g0, 5 + P), dtype=torch.float32, device=dev)
        for start in range(g0, g1, batch_size):
            bs = min(batch_size, g1 - start)
            z[start - g0:start - g0 + bs].copy_(proposal.sample((bs,)), non_blocking=True)
            if (start // batch_size) % 50 == 0:
                print(f"Simulated {start + bs:,}/{num_simulations:,}")
        x = sim_wrapper(z, mu_sensory=mu_sensory, p_success=p_success, P=P, log_rt=log_rt, seed=seed, trial_offset=g0)
        # the reference's sanity checks (:62-66), evaluated where the data is: one flag word per block
        choice = x[:, -1]
        counts = torch.stack([(choice == 0).sum(), (choice == 1).sum(), (choice == 2).sum()])    # (no host sync)
        checks.append(torch.stack([torch.isfinite(z).all(), torch.isfinite(x).all(), counts.sum() == choice.numel()]))
        outcomes += counts
        if as_records:
            with torch.cuda.device(dev):
                rec = torch.empty((g1 - g0, 8), dtype=torch.int32, device=dev)
                n_generic = torch.empty((1,), dtype=torch.int64, device=dev)
                _native.check(L.ddm_pack_z_dev(z.data_ptr(), 5 + P, g1 - g0, P, rec.data_ptr(), n_generic.data_ptr(),
                                               stream.cuda_stream), "ddm_pack_z_dev")
        if copier is not stream:
            copier.wait_stream(stream)
        with torch.cuda.stream(copier):
            if as_records:
                rec_host[g0:g1].copy_(rec, non_blocking=True)
                generic_host[g:g + 1].copy_(n_generic, non_blocking=True)
            else:
                z_all[g0:g1].copy_(z, non_blocking=True)
            x_all[g0:g1].copy_(x, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copier)
        if copier is not stream:
            for t in (z, x) + ((rec, n_generic) if as_records else ()):
                t.record_stream(copier)
        if as_records:
            jobs.append(unpacker.submit(rebuild, g, g0, g1, ev, z))
        in_flight.append((ev, z, x))
        while len(in_flight) > 2:      # bound the device memory held by blocks whose copies are still queued
            in_flight.pop(0)[0].synchronize()
    for ev, _, _ in in_flight:
        ev.synchronize()
    in_flight.clear()
    for job in jobs:
        job.result()
    if unpacker is not None:
        unpacker.shutdown()

    assert z_all.shape[0] == num_simulations
    assert x_all.shape[0] == num_simulations
    ok = torch.stack(checks).all(0).tolist() if checks else [True, True, True]
    assert ok[0], "non-finite values in z"
    assert ok[1], "non-finite values in x"
    assert ok[2], "choice outside {0, 1, 2}"

    print("Training x shape:", tuple(x_all.shape), " (N,2) = [rt(or log rt), choice]")
    print("Training z shape:", tuple(z_all.shape), " (N, 5+P) = [theta, pulses]")
    print("Unique outcomes in training (choice):", [float(c) for c, k in enumerate(outcomes.tolist()) if k > 0])
    return z_all, x_all


@torch.no_grad()
def simulate_observed_session(theta_true: torch.Tensor, num_trials: int, device, *, mu_sensory: float,
                              p_success: float, P: int, seed: int = 123, log_rt: bool,
                              noise_seed: Optional[int] = None, noise: Optional[torch.Tensor] = None):
    """One observed session at ``theta_true``: (x_o (T,2), pulses_o (T,P)) on the CPU
    (reference :74-99).  ``seed`` seeds the stimulus exactly as in the reference."""
    dev = compute_device(device)
    rng = np.random.default_rng(seed)
    pulses_o = generate_pulse_matrix_device(rng, num_trials, P, p_success=p_success, device=dev)
    theta_rep = theta_true.view(1, 5).expand(num_trials, 5)
    x_o = simulate_trials(theta_rep, pulses_o, mu_sensory=mu_sensory, log_rt=log_rt, seed=noise_seed, noise=noise,
                          device=dev)
    return x_o.cpu(), pulses_o.cpu()


def summarize_trials(name: str, x: torch.Tensor) -> None:
    """Print the RT range and the outcome histogram of a batch of trials (reference :102-111)."""
    counts = torch.bincount(x[:, 1].to(torch.int64), minlength=3)
    frac = counts.float() / counts.sum().clamp_min(1)
    print(f"{name}: n={len(x)}  rt[min,max]=({x[:, 0].min().item():.4f},{x[:, 0].max().item():.4f})  "
          f"choice counts={counts.tolist()}  frac={frac.tolist()}")
