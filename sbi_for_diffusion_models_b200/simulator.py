"""Host driver of the CUDA trial simulator (``ddm_sim_f32``).

This is the one place that turns torch tensors into raw device pointers.  It mirrors what
``_simulate_rt_choice_batch_torch`` does around its time loop
(/root/reference/src/sbi_for_diffusion_models/models/rt_choice_model.py:125-178): cast theta
to fp32, normalise the pulse matrix (1-D / one-row broadcast / width check) and raise the
same ``ValueError``s -- before anything is launched.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native, constants


@dataclass(frozen=True)
class Schedule:
    """Time grid of the simulator and the fp32 scalars the kernel consumes."""
    n_max: int
    steps_per_pulse: int
    n_pulses: int
    dt: float
    t_max: float
    t_nd_hi: float
    noise_scale: float

    @staticmethod
    def from_constants(mu_sensory: float = 1.0, *, dt: Optional[float] = None, t_max: Optional[float] = None,
                       pulse_interval: Optional[float] = None) -> "Schedule":
        # constants are read at call time, like the reference (rt_choice_model.py:137-138)
        dt = float(constants.DT_CHOICE if dt is None else dt)
        t_max = float(constants.T_MAX if t_max is None else t_max)
        pulse_interval = float(constants.PULSE_INTERVAL if pulse_interval is None else pulse_interval)
        n_max = int(np.floor(t_max / dt))                            # :52
        spp = max(int(np.round(pulse_interval / dt)), 1)             # :53
        f32 = lambda v: float(np.float32(v))
        return Schedule(n_max=n_max, steps_per_pulse=spp, n_pulses=(n_max + spp - 1) // spp,  # :59
                        dt=f32(dt), t_max=f32(t_max), t_nd_hi=f32(t_max - 1e-6),               # :135
                        noise_scale=f32(float(mu_sensory) * float(np.sqrt(dt))))                # :146-147,186


@dataclass
class SimStats:
    useful_steps: int      # sum over trials of hit_step
    lane_steps: int        # Euler steps issued by all lanes, busy or idle
    generic_rows: int      # trials whose pulse row was not +-1

    @property
    def lane_efficiency(self) -> float:
        return self.useful_steps / self.lane_steps if self.lane_steps else 0.0


def next_seed() -> int:
    """A 63-bit Philox key drawn from torch's global CPU generator, so ``torch.manual_seed``
    makes simulations reproducible just as it does for the reference's ``torch.randn``."""
    return int(torch.randint(0, 2**63 - 1, (1,), dtype=torch.int64).item())


def compute_device(device=None) -> torch.device:
    torch_ = _native.require_cuda()
    if device is None:
        return torch_.device("cuda", torch_.cuda.current_device())
    device = torch_.device(device)
    if device.type != "cuda":
        return torch_.device("cuda", torch_.cuda.current_device())
    return device if device.index is not None else torch_.device("cuda", torch_.cuda.current_device())


def _as_f32_rows(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """fp32 on `dev` with unit stride along the last dim (row stride is free)."""
    t = t.to(device=dev, dtype=torch.float32, non_blocking=True)
    if t.stride(-1) != 1 and t.shape[-1] > 1:
        t = t.contiguous()
    return t


def normalise_pulses(pulse_sides, n_trials: int, sched: Schedule, dev: torch.device) -> Tuple[torch.Tensor, int]:
    """rt_choice_model.py:94-109 and :165-178 -> (tensor on dev, row stride in floats; 0 = broadcast)."""
    s = pulse_sides if isinstance(pulse_sides, torch.Tensor) else torch.from_numpy(np.asarray(pulse_sides))
    if s.ndim == 1:
        s = s.view(1, -1)
    if s.ndim != 2:
        raise ValueError(f"pulse_sides must have shape (N,P) or (P,), got {tuple(s.shape)}")
    broadcast = s.shape[0] == 1 and n_trials > 1
    if not broadcast and s.shape[0] != n_trials:
        raise ValueError(
            f"pulse_sides first dim must match batch size N={n_trials} (or be 1 for broadcast), got {s.shape[0]}")
    if s.shape[1] < sched.n_pulses:
        raise ValueError(
            f"pulse_sides has P={s.shape[1]} pulses but simulator needs at least {sched.n_pulses} "
            f"for T_MAX={constants.T_MAX}s")
    s = _as_f32_rows(s, dev)
    return s, (0 if broadcast else s.stride(0))


def simulate_trials(theta: torch.Tensor, pulse_sides, *, mu_sensory: float = 1.0, log_rt: bool = False,
                    seed: Optional[int] = None, trial_offset: int = 0, noise: Optional[torch.Tensor] = None,
                    return_steps: bool = False, return_stats: bool = False, device=None,
                    schedule: Optional[Schedule] = None, out: Optional[torch.Tensor] = None,
                    peer_blocks=None):
    """Simulate one trial per row of ``theta`` on the GPU.

    theta (N,5) any float dtype, CPU or CUDA; pulse_sides (N,P) / (1,P) / (P,).
    Returns x (N,2) fp32 ``[rt (or log rt), choice]`` on the compute device, optionally
    followed by ``hit_step`` (N,) int32 and a :class:`SimStats`.

    ``noise`` (n_max, N) fp32 replaces the native Philox stream with shared noise: the
    output then equals the reference's bit for bit.  ``seed`` defaults to a draw from torch's
    global generator.

    ``peer_blocks``: (N,2) fp32 blocks in OTHER GPUs' memory (peer-mapped views, see
    ``sharding.PeerGather``); the kernel stores every result there as well -- the all-gather of a sharded run
    fused into the launch (``ddm_sim_gather_f32``).
    """
    L = _native.lib()
    dev = compute_device(device if device is not None else (theta.device if theta.is_cuda else None))
    if theta.ndim == 1:
        theta = theta.view(1, -1)
    if theta.shape[-1] != 5:
        raise ValueError(f"Expected theta shape (N,5) or (5,), got {tuple(theta.shape)}")
    sched = schedule or Schedule.from_constants(mu_sensory)
    n = theta.shape[0]
    with torch.cuda.device(dev):
        th = _as_f32_rows(theta, dev)
        s, ld_pulses = normalise_pulses(pulse_sides, n, sched, dev)
        if out is None:
            out = torch.empty((n, 2), dtype=torch.float32, device=dev)
        else:
            assert out.shape == (n, 2) and out.dtype == torch.float32 and out.is_contiguous() and out.device == dev
        steps = torch.empty((n,), dtype=torch.int32, device=dev) if return_steps else None
        ws = torch.empty((_native.WS_WORDS,), dtype=torch.int64, device=dev)
        noise_ptr, ld_noise = None, 0
        if noise is not None:
            noise = _as_f32_rows(noise, dev)
            if noise.ndim != 2 or noise.shape[0] < sched.n_max or noise.shape[1] < n:
                raise ValueError(f"noise must be (n_max={sched.n_max}, >=N={n}), got {tuple(noise.shape)}")
            noise_ptr, ld_noise = noise.data_ptr(), noise.stride(0)
        if seed is None:
            seed = 0 if noise is not None else next_seed()
        if peer_blocks:
            if noise is not None or return_steps:
                raise ValueError("the fused gather launch takes native noise and does not return hit steps")
            for b in peer_blocks:
                if tuple(b.shape) != (n, 2) or b.dtype != torch.float32 or not b.is_contiguous():
                    raise ValueError(f"peer blocks must be contiguous (N,2) fp32 with N={n}, got {tuple(b.shape)}")
            ptrs = (ctypes.c_void_p * len(peer_blocks))(*[b.data_ptr() for b in peer_blocks])
            rc = L.ddm_sim_gather_f32(th.data_ptr() if n else None, th.stride(0) if n else 5,
                                      s.data_ptr(), ld_pulses, n, s.shape[1],
                                      sched.n_max, sched.steps_per_pulse, sched.dt, sched.t_max, sched.t_nd_hi,
                                      sched.noise_scale, ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(trial_offset),
                                      int(bool(log_rt)), out.data_ptr() if n else None, ptrs, len(peer_blocks),
                                      ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _native.check(rc, "ddm_sim_gather_f32")
        else:
            rc = L.ddm_sim_f32(th.data_ptr() if n else None, th.stride(0) if n else 5,
                               s.data_ptr(), ld_pulses, n, s.shape[1],
                               sched.n_max, sched.steps_per_pulse, sched.dt, sched.t_max, sched.t_nd_hi,
                               sched.noise_scale, ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(trial_offset),
                               noise_ptr, ld_noise, int(bool(log_rt)), out.data_ptr() if n else None,
                               steps.data_ptr() if steps is not None and n else None,
                               ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _native.check(rc, "ddm_sim_f32")
        # keep inputs alive until the stream has consumed them
        for t in (th, s, noise):
            if t is not None:
                t.record_stream(torch.cuda.current_stream(dev))
    result = [out]
    if return_steps:
        result.append(steps)
    if return_stats:
        w = ws.cpu().tolist()
        result.append(SimStats(useful_steps=w[_native.WS_USEFUL_STEPS], lane_steps=w[_native.WS_LANE_STEPS],
                               generic_rows=w[_native.WS_GENERIC_ROWS]))
    return result[0] if len(result) == 1 else tuple(result)


def philox_normals(seed: int, n_trials: int, n_steps: int, *, trial_offset: int = 0, device=None) -> torch.Tensor:
    """(n_steps, n_trials) fp32: the normals ``simulate_trials`` consumes for this seed."""
    L = _native.lib()
    dev = compute_device(device)
    with torch.cuda.device(dev):
        out = torch.empty((n_steps, n_trials), dtype=torch.float32, device=dev)
        _native.check(L.ddm_philox_normals_f32(ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(trial_offset),
                                               n_trials, n_steps, out.data_ptr(), max(n_trials, 1),
                                               torch.cuda.current_stream(dev).cuda_stream), "ddm_philox_normals_f32")
    return out


def philox_words(seed: int, n_trials: int, n_steps: int, *, trial_offset: int = 0, device=None) -> torch.Tensor:
    """(n_steps, n_trials) raw Philox4x32-10 words as int32 bit patterns."""
    L = _native.lib()
    dev = compute_device(device)
    with torch.cuda.device(dev):
        out = torch.empty((n_steps, n_trials), dtype=torch.int32, device=dev)
        _native.check(L.ddm_philox_words_u32(ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(trial_offset),
                                             n_trials, n_steps, out.data_ptr(), max(n_trials, 1),
                                             torch.cuda.current_stream(dev).cuda_stream), "ddm_philox_words_u32")
    return out


def pack_min_threads() -> int:
    """Fewer host threads than this per process: send the fp32 rows over the link instead of packing.
    Measured on the B200 host (Xeon, 16 vCPU): the AVX-512 packer reads 16.6 GB/s of z per thread, 62 GB/s on 4
    threads and 165 GB/s on 16; the AVX2 one 9.9 / 37 / 94 GB/s -- against 55.6 GB/s of fp32 rows over PCIe 5 x16
    (``tools/pack_bw.py``).  So packing beats the link from 4 threads (AVX-512) or 8 threads (AVX2) up."""
    import os
    env = os.environ.get("DDM_PACK_MIN_THREADS")
    if env:
        return max(1, int(env))
    return 4 if _native.lib().ddm_pack_simd_bits() >= 512 else 8


def pack_threads() -> int:
    """CPU threads ``ddm_pack_z_host`` may use: DDM_PACK_THREADS, else this process's share of the cores."""
    import os
    env = os.environ.get("DDM_PACK_THREADS")
    if env:
        return max(1, int(env))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(32, cores // ranks))


_launches_block = None   # probed once per process; set to True for good after a streaming launch timed out


def launches_block() -> bool:
    """True when a kernel launch does not return until the kernel has finished (CUDA_LAUNCH_BLOCKING=1, or a
    tool that serialises launches: Nsight Compute, compute-sanitizer, a debugger).  The streaming ingest
    normally launches its persistent kernel FIRST and feeds it while it runs; when launches block, the launch
    would wait for rows nobody can enqueue any more, so the copies go first.  Decided by asking the driver, not
    by guessing from the environment: ``ddm_probe_launch_blocking`` launches a one-thread kernel that waits up to
    50 ms for a word the host raises the moment the launch call is back.  DDM_INGEST_ORDER=copies_first /
    launch_first overrides the probe; a streaming launch that ever times out waiting for its rows (any other cause
    of starvation) also switches the process to copies-first for good (``HostPipeline._check_slot``)."""
    import os
    global _launches_block
    order = os.environ.get("DDM_INGEST_ORDER")
    if order in ("copies_first", "launch_first"):
        return order == "copies_first"
    if _launches_block is None:
        out = ctypes.c_int(0)
        _native.check(_native.lib().ddm_probe_launch_blocking(50_000, ctypes.byref(out)), "ddm_probe_launch_blocking")
        _launches_block = bool(out.value)
    return _launches_block


class HostPipeline:
    """Streams a host-resident z = [theta, pulses] matrix through the GPU.

    Per batch (up to ``max_batch`` rows) ONE persistent streaming launch on the kernel stream consumes
    trials as they arrive: the copy stream carries the rows to the device in chunks, each followed by an
    8-byte copy that raises a device word ``ready`` to the number of rows delivered so far (a warp that
    claims rows beyond ``ready`` sleeps until the copy engine catches up).  PCIe ingest and simulation
    overlap inside a single launch, so a batch pays one drain phase instead of one per chunk.  Two slots
    are double-buffered across batches.  Global trial offsets make the result identical to a single
    launch over all rows.

    Packed ingest (default when the schedule needs <= 96 pulses and this process has at least
    ``pack_min_threads()`` cores to itself): pulse sides are +-1, so the 340
    bytes of an fp32 row carry 32 bytes of information and the PCIe link, not the kernel, bounds the
    fp32 path.  Each chunk is packed on the host cores (``ddm_pack_z_host``) into 32-byte records while
    the previous chunk is on the link and the kernel is already running; ``ddm_sim_packed_f32``
    consumes the records.  Same bits out.  A batch in which some row holds a value other than +-1 is
    re-run through the fp32 path."""

    def __init__(self, n_cols: int, max_batch: int = 1 << 22, chunk: int = 1 << 18, device=None):
        self.dev = compute_device(device)
        self.max_batch, self.chunk, self.n_cols = int(max_batch), int(chunk), int(n_cols)
        n_marks = -(-self.max_batch // self.chunk)
        with torch.cuda.device(self.dev):
            self.copy_stream = torch.cuda.Stream(self.dev)
            self.kernel_stream = torch.cuda.Stream(self.dev)
            self.slots = []
            for _ in range(2):
                self.slots.append({
                    "z": None, "pk": None, "pk_host": None,      # allocated on first use
                    "x": torch.empty((self.max_batch, 2), dtype=torch.float32, device=self.dev),
                    "ws": torch.zeros((_native.WS_WORDS,), dtype=torch.int64, device=self.dev),
                    "ready": torch.zeros((1,), dtype=torch.int64, device=self.dev),
                    "done": None, "host_free": None,
                })
            # constant pinned tables the flag copies read from (never rewritten: copies run later)
            self.marks = (torch.arange(1, n_marks + 1, dtype=torch.int64) * self.chunk).pin_memory()
            self.all_rows = torch.full((1,), 1 << 62, dtype=torch.int64).pin_memory()
        self.launches = 0
        self.packed_batches = 0
        self.h2d_bytes = 0
        self._pending = []
        # ingest mode chosen by measurement (choose_packed / report): seconds per row of each mode, calls so far
        self._times = {True: [], False: []}
        self._calls = 0

    PROBE_MIN_ROWS = 1 << 20    # only batches this large say anything about the ingest rate
    REPROBE_EVERY = 64          # the slower mode gets another look this often (the host's load may have changed)

    def choose_packed(self, sched: Schedule, n_rows: int) -> bool:
        """Packed records or fp32 rows for the next call?  Both give the same bits; which one is faster depends on
        the HOST: packing reads z with the CPU cores (340 + 64 bytes of DRAM traffic per trial) and sends 32 bytes
        over PCIe, fp32 rows are read by the copy engine (340 bytes of DRAM traffic and 340 over PCIe).  One process
        with 16 cores to itself packs at 165 GB/s against a 55 GB/s link; eight ranks sharing 32 cores and ~170 GB/s
        of host memory bandwidth are better off with the plain copies.  So: start from the thread-count rule, try
        each mode twice on large batches (the first call of a mode also allocates its staging blocks), keep the faster
        one, look again every REPROBE_EVERY calls."""
        import os
        if sched.n_pulses > 96:
            return False
        forced = os.environ.get("DDM_INGEST_MODE")
        if forced in ("packed", "rows"):
            return forced == "packed"
        first = pack_threads() >= pack_min_threads()
        if n_rows < self.PROBE_MIN_ROWS:
            return first
        # the first call of a mode also allocates its staging blocks: two samples per mode before it is judged
        for mode in (first, not first):
            if len(self._times[mode]) < 2:
                return mode
        best = min(self._times[True][1:]) <= min(self._times[False][1:])
        if self._calls % self.REPROBE_EVERY == self.REPROBE_EVERY - 1:
            return not best
        return best

    def report(self, packed: bool, n_rows: int, seconds: float) -> None:
        """Wall time of one whole call (run + synchronize) in the given mode."""
        self._calls += 1
        if n_rows >= self.PROBE_MIN_ROWS:
            t = self._times[packed]
            t.append(seconds / n_rows)
            if len(t) > 5:          # first sample (allocation) + the last four
                del t[1]

    def _slot_buffers(self, slot, packed: bool):
        with torch.cuda.device(self.dev):
            if packed and slot["pk"] is None:
                slot["pk"] = torch.empty((self.max_batch, 8), dtype=torch.int32, device=self.dev)
                slot["pk_host"] = torch.empty((self.max_batch, 8), dtype=torch.int32).pin_memory()
                slot["marks"] = torch.zeros((-(-self.max_batch // self.chunk),), dtype=torch.int64).pin_memory()
            if not packed and slot["z"] is None:
                slot["z"] = torch.empty((self.max_batch, self.n_cols), dtype=torch.float32, device=self.dev)

    def run(self, z_host: torch.Tensor, x_host: torch.Tensor, *, sched: Schedule, seed: int, log_rt: bool = False,
            trial_offset: int = 0, packed: Optional[bool] = None) -> None:
        """z_host (N, 5+P) fp32 CPU (pinned for full speed) -> x_host (N,2) fp32 CPU (pinned).
        Returns after everything has been enqueued; call ``synchronize`` before reading."""
        L = _native.lib()
        n = z_host.shape[0]
        assert z_host.dtype == torch.float32 and z_host.shape[1] == self.n_cols and z_host.stride(1) == 1
        assert x_host.shape == (n, 2) and x_host.dtype == torch.float32 and x_host.is_contiguous()
        P = self.n_cols - 5
        if P < sched.n_pulses:
            raise ValueError(f"pulse_sides has P={P} pulses but simulator needs at least {sched.n_pulses}")
        if packed is None:
            packed = sched.n_pulses <= 96 and pack_threads() >= pack_min_threads()
        elif packed and sched.n_pulses > 96:
            raise ValueError("packed ingest holds at most 96 pulse signs per trial")
        cur = torch.cuda.current_stream(self.dev)
        cs, ks = self.copy_stream, self.kernel_stream
        cs.wait_stream(cur)
        ks.wait_stream(cur)
        n_threads = pack_threads()
        redo = []
        for b, start in enumerate(range(0, n, self.max_batch)):
            slot = self.slots[b % 2]
            bs = min(self.max_batch, n - start)
            self._slot_buffers(slot, packed)
            if slot["done"] is not None:
                self._check_slot(slot)               # its previous batch: finished, and without a starved launch
                cs.wait_event(slot["done"])          # the previous kernel on this slot has finished
            with torch.cuda.stream(cs):
                slot["ready"].zero_()
                reset = torch.cuda.Event()
                reset.record(cs)
            common = (sched.n_max, sched.steps_per_pulse, sched.dt, sched.t_max, sched.t_nd_hi, sched.noise_scale,
                      ctypes.c_uint64(seed & (2**64 - 1)), ctypes.c_uint64(trial_offset + start), int(bool(log_rt)))

            def launch():
                # the kernel may start as soon as `ready` was reset: it sleeps on rows that have not arrived
                ks.wait_event(reset)
                with torch.cuda.stream(ks):
                    if packed:
                        rc = L.ddm_sim_packed_f32(slot["pk"].data_ptr(), bs, *common, slot["x"].data_ptr(), None,
                                                  slot["ws"].data_ptr(), slot["ready"].data_ptr(), ks.cuda_stream)
                    else:
                        zd = slot["z"]
                        rc = L.ddm_sim_stream_f32(zd.data_ptr(), self.n_cols, zd.data_ptr() + 20, self.n_cols, bs, P,
                                                  *common, slot["x"].data_ptr(), slot["ws"].data_ptr(),
                                                  slot["ready"].data_ptr(), ks.cuda_stream)
                    _native.check(rc, "ddm_sim_packed_f32" if packed else "ddm_sim_stream_f32")
                    self.launches += 1
                    x_host[start:start + bs].copy_(slot["x"][:bs], non_blocking=True)
                    slot["done"] = torch.cuda.Event()
                    slot["done"].record(ks)
                    slot["checked"] = False
                    self._pending.append(slot)

            if packed:
                if slot["host_free"] is not None:
                    slot["host_free"].synchronize()  # the copy engine has read the staging block's last contents
                copies_first = launches_block()
                if not copies_first:
                    launch()                         # resident and waiting while the host packs
                # marks[k] = rows delivered once chunk k has landed (the last one: "all of them")
                n_chunks = -(-bs // self.chunk)
                marks = slot["marks"]
                marks[:n_chunks] = self.marks[:n_chunks]
                marks[n_chunks - 1] = 1 << 62
                got = ctypes.c_int64(0)
                src = z_host[start:start + bs]
                rc = L.ddm_ingest_packed(src.data_ptr(), src.stride(0), bs, sched.n_pulses, self.chunk,
                                         slot["pk_host"].data_ptr(), slot["pk"].data_ptr(), slot["ready"].data_ptr(),
                                         marks.data_ptr(), n_threads, cs.cuda_stream, ctypes.byref(got))
                _native.check(rc, "ddm_ingest_packed")
                generic = got.value
                if copies_first:
                    launch()
                slot["host_free"] = torch.cuda.Event()
                slot["host_free"].record(cs)
                self.h2d_bytes += bs * 32
                self.packed_batches += 1
                if generic:
                    redo.append((start, bs))         # rows with pulse values other than +-1: fp32 path
            else:
                with torch.cuda.stream(cs):
                    for k, a in enumerate(range(0, bs, self.chunk)):
                        e = min(a + self.chunk, bs)
                        slot["z"][a:e].copy_(z_host[start + a:start + e], non_blocking=True)
                        slot["ready"].copy_(self.all_rows if e == bs else self.marks[k:k + 1], non_blocking=True)
                launch()
                self.h2d_bytes += bs * self.n_cols * 4
        cur.wait_stream(ks)
        cur.wait_stream(cs)
        for start, bs in redo:
            self.run(z_host[start:start + bs], x_host[start:start + bs], sched=sched, seed=seed, log_rt=log_rt,
                     trial_offset=trial_offset + start, packed=False)

    def _check_slot(self, slot) -> None:
        """Raise if the batch that last ran on ``slot`` gave up waiting for its rows (DDM_WS_ERROR), and make
        every later batch of this process enqueue its copies before its launch."""
        global _launches_block
        self._pending = [q for q in self._pending if q is not slot]
        if slot.get("checked", True):
            return
        slot["done"].synchronize()
        slot["checked"] = True
        if int(slot["ws"][_native.WS_ERROR].item()) != 0:
            _launches_block = True
            raise RuntimeError("streaming simulator launch timed out waiting for host->device copies "
                               "(its outputs are incomplete); later batches enqueue their copies first")

    def synchronize(self) -> None:
        torch.cuda.current_stream(self.dev).synchronize()
        pending, self._pending = self._pending, []
        for slot in pending:
            self._check_slot(slot)
