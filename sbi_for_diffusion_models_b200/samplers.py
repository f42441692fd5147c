"""Many-chain slice sampler that keeps every chain on the device.

The reference samples the posterior with sbi's ``MCMCPosterior`` (mnle.py:77-93; NUTS through
pyro, one process per chain, CPU) and its notebook uses sbi's ``slice_np_vectorized``.  Both are
third-party and host-bound.  Here the same job -- draw from exp(potential) -- is done by a
coordinate-wise slice sampler (Neal 2003: stepping out + shrinkage) whose N chains advance in
lock-step as one (N, D) tensor, so each step is ONE call of the potential over all chains: exactly
the (trials x chains) shape the fused MNLE kernel is built for.  Gradient free.

Pure torch: runs on whatever device ``init`` lives on (tests use the CPU, the product the GPU).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class GraphedLogProb:
    """``fn``: (N, D) -> (N,) captured once into a CUDA graph and replayed on static buffers.

    A slice-sampler step is one potential evaluation followed by a handful of tiny elementwise
    ops; at the sizes of one observed session the Python / launch path of the potential (prior,
    workspace allocation, two kernel launches) costs several times the kernels themselves, so the
    whole evaluation is recorded once and replayed.  Falls back to calling ``fn`` for inputs that
    are not CUDA tensors of the captured shape."""

    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], example: torch.Tensor, warmup: int = 2):
        self.fn = fn
        self.graph = None
        if not example.is_cuda:
            return
        self.x = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    fn(self.x)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    self.y = fn(self.x)
            self.graph = graph
        except Exception:       # fn synchronises or leaves the device (e.g. a CPU-only prior): call it directly
            self.graph = None
            torch.cuda.synchronize(example.device)
        torch.cuda.current_stream(example.device).wait_stream(side)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.graph is None or x.shape != self.x.shape or x.device != self.x.device or x.dtype != self.x.dtype:
            return self.fn(x)
        self.x.copy_(x)
        self.graph.replay()
        return self.y.clone()


class PhiloxUniforms:
    """Counter-based uniforms for chains [first_row, first_row + n_rows): draw k of global chain g is
    Philox4x32-10(key = seed + k // block, counter = (g, k % block)), whichever process owns the chain
    and whatever other chains run beside it -- so a sampler run sharded over GPUs reproduces the
    single-GPU run bit for bit.  Uses the simulator's device generator (``ddm_philox_words_u32``)."""

    def __init__(self, seed: int, first_row: int, n_rows: int, device, block: int = 256):
        self.seed, self.first, self.n, self.device, self.block = int(seed), int(first_row), int(n_rows), device, int(block)
        self._buf, self._blk = None, -1

    def __call__(self, k: int) -> torch.Tensor:
        """Draw number ``k`` of every chain: (n_rows,) fp32 in (0, 1)."""
        from .simulator import philox_words
        blk, j = divmod(int(k), self.block)
        if blk != self._blk:
            words = philox_words(self.seed + blk, self.n, self.block, trial_offset=self.first, device=self.device)
            # top 24 bits -> (0, 1): ((w >> 8) + 0.5) * 2^-24, exact in fp32
            self._buf = ((words >> 8) & 0xFFFFFF).to(torch.float32).add_(0.5).mul_(2.0 ** -24)
            self._blk = blk
        return self._buf[j]


class VectorizedSliceSampler:
    """``log_prob_fn``: (N, D) -> (N,), ``-inf`` outside the support.  ``init``: (N, D) starting
    points with finite log-probability."""

    def __init__(self, log_prob_fn: Callable[[torch.Tensor], torch.Tensor], init: torch.Tensor, *,
                 init_width: float = 0.1, max_step_out: int = 8, max_shrink: int = 64,
                 generator: Optional[torch.Generator] = None, uniforms: Optional[Callable[[int], torch.Tensor]] = None,
                 chain_groups: int = 1):
        if init.ndim != 2:
            raise ValueError(f"init must be (num_chains, dim), got {tuple(init.shape)}")
        self.f = log_prob_fn
        self.x = init.clone()
        self.gen, self.uniforms = generator, uniforms
        # chains come in ``chain_groups`` equal consecutive groups (SBC: one group per dataset), each
        # with its own slice widths, so that a group's trajectory does not depend on its neighbours
        self.groups = int(chain_groups)
        if init.shape[0] % self.groups:
            raise ValueError("the number of chains must be a multiple of chain_groups")
        self.N, self.D = init.shape
        self.lp = self._eval(self.x)
        if not bool(torch.isfinite(self.lp).all()):
            raise ValueError("every chain must start at a point of finite log-probability")
        per = self.N // self.groups
        self.width = float(init_width) * init.abs().view(self.groups, per, self.D).mean(dim=1).clamp_min(1e-3)   # (G, D)
        self.max_step_out, self.max_shrink = int(max_step_out), int(max_shrink)
        self.n_evals, self._updates = 0, 0

    # ------------------------------------------------------------------------------------
    def _eval(self, x: torch.Tensor) -> torch.Tensor:
        lp = self.f(x)
        self.n_evals = getattr(self, "n_evals", 0) + 1
        return torch.nan_to_num(lp, nan=-float("inf"), posinf=float("inf"), neginf=-float("inf"))

    def _rand(self, k: int) -> torch.Tensor:
        """Uniform number ``k`` of every chain.  With a counter-based source the index is explicit
        (update number, draw within the update), so chains that need more shrinkage steps than others
        do not shift anybody's later draws."""
        if self.uniforms is not None:
            return self.uniforms(k).to(self.x.dtype)
        return torch.rand((self.N,), dtype=self.x.dtype, device=self.x.device, generator=self.gen)

    def _with(self, d: int, v: torch.Tensor) -> torch.Tensor:
        y = self.x.clone()
        y[:, d] = v
        return y

    def _update_dim(self, d: int) -> torch.Tensor:
        """One slice update of coordinate ``d`` for all chains; returns the final bracket sizes."""
        x0 = self.x[:, d]
        w = self.width[:, d].repeat_interleave(self.N // self.groups)        # per-chain width of its group
        base = self._updates * (3 + self.max_shrink)
        self._updates += 1
        log_y = self.lp + torch.log(self._rand(base).clamp_min(1e-37))
        lo = x0 - w * self._rand(base + 1)
        hi = lo + w
        # stepping out with a limit (Neal 2003, fig. 3): at most m - 1 expansions in total, split at random between
        # the two ends (J to the left, m - 1 - J to the right), which keeps the update reversible for any m.  All
        # chains advance in lock-step, so the number of potential calls of an update is set by its slowest chain:
        # an uncapped search costs ~10 calls per end, every time, for the sake of a handful of chains.
        m = self.max_step_out
        J = torch.floor(m * self._rand(base + 2 + self.max_shrink)).clamp_(0, m - 1)
        K = (m - 1) - J
        grow = J > 0
        for _ in range(m - 1):
            if not bool(grow.any()):
                break
            grow = grow & (self._eval(self._with(d, lo)) > log_y)
            lo = torch.where(grow, lo - w, lo)
            J = J - grow.to(J.dtype)
            grow = grow & (J > 0)
        grow = K > 0
        for _ in range(m - 1):
            if not bool(grow.any()):
                break
            grow = grow & (self._eval(self._with(d, hi)) > log_y)
            hi = torch.where(grow, hi + w, hi)
            K = K - grow.to(K.dtype)
            grow = grow & (K > 0)
        # shrinkage
        todo = torch.ones_like(x0, dtype=torch.bool)
        new_x, new_lp = x0.clone(), self.lp.clone()
        for it in range(self.max_shrink):
            prop = lo + (hi - lo) * self._rand(base + 2 + it)
            lp = self._eval(self._with(d, torch.where(todo, prop, new_x)))
            ok = todo & (lp > log_y)
            new_x = torch.where(ok, prop, new_x)
            new_lp = torch.where(ok, lp, new_lp)
            todo = todo & ~ok
            if not bool(todo.any()):
                break
            left = todo & (prop < x0)
            lo = torch.where(left, prop, lo)
            hi = torch.where(todo & ~left, prop, hi)
        # chains that never found a point keep their state (probability ~0 with max_shrink = 64)
        self.x = self._with(d, new_x)
        self.lp = new_lp
        # the bracket as it stood when the chain accepted (shrunk around the slice): this is what sbi's slice
        # samplers average into the width while tuning.  (The stepped-out bracket is never smaller than the
        # width, so averaging THAT only ratchets the width up: 35 evaluations per update instead of ~15.)
        return hi - lo

    def sweep(self, tune: bool = False) -> None:
        for d in range(self.D):
            size = self._update_dim(d)
            if tune:   # running estimate of the typical slice width, as sbi's slice samplers do
                self.width[:, d] = 0.5 * self.width[:, d] + 0.5 * size.view(self.groups, -1).mean(dim=1).clamp_min(1e-6)

    @torch.no_grad()
    def run(self, num_samples_per_chain: int, *, warmup: int = 100, thin: int = 1) -> torch.Tensor:
        """-> (num_samples_per_chain, N, D) after ``warmup`` tuning sweeps."""
        for _ in range(int(warmup)):
            self.sweep(tune=True)
        out = torch.empty((int(num_samples_per_chain), self.N, self.D), dtype=self.x.dtype, device=self.x.device)
        for s in range(int(num_samples_per_chain)):
            for _ in range(max(int(thin), 1)):
                self.sweep()
            out[s] = self.x
        return out
