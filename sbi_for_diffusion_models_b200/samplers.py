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


class VectorizedSliceSampler:
    """``log_prob_fn``: (N, D) -> (N,), ``-inf`` outside the support.  ``init``: (N, D) starting
    points with finite log-probability."""

    def __init__(self, log_prob_fn: Callable[[torch.Tensor], torch.Tensor], init: torch.Tensor, *,
                 init_width: float = 0.1, max_step_out: int = 16, max_shrink: int = 64,
                 generator: Optional[torch.Generator] = None):
        if init.ndim != 2:
            raise ValueError(f"init must be (num_chains, dim), got {tuple(init.shape)}")
        self.f = log_prob_fn
        self.x = init.clone()
        self.gen = generator
        self.N, self.D = init.shape
        self.lp = self._eval(self.x)
        if not bool(torch.isfinite(self.lp).all()):
            raise ValueError("every chain must start at a point of finite log-probability")
        self.width = torch.full((self.D,), float(init_width), dtype=init.dtype, device=init.device) * \
            init.abs().mean(dim=0).clamp_min(1e-3)
        self.max_step_out, self.max_shrink = int(max_step_out), int(max_shrink)
        self.n_evals = 0

    # ------------------------------------------------------------------------------------
    def _eval(self, x: torch.Tensor) -> torch.Tensor:
        lp = self.f(x)
        self.n_evals = getattr(self, "n_evals", 0) + 1
        return torch.nan_to_num(lp, nan=-float("inf"), posinf=float("inf"), neginf=-float("inf"))

    def _rand(self) -> torch.Tensor:
        return torch.rand((self.N,), dtype=self.x.dtype, device=self.x.device, generator=self.gen)

    def _with(self, d: int, v: torch.Tensor) -> torch.Tensor:
        y = self.x.clone()
        y[:, d] = v
        return y

    def _update_dim(self, d: int) -> torch.Tensor:
        """One slice update of coordinate ``d`` for all chains; returns the bracket sizes."""
        x0 = self.x[:, d]
        w = self.width[d]
        log_y = self.lp + torch.log(self._rand().clamp_min(1e-37))
        lo = x0 - w * self._rand()
        hi = lo + w
        # stepping out: only chains whose bracket end is still inside the slice move
        grow = torch.ones_like(x0, dtype=torch.bool)
        for _ in range(self.max_step_out):
            grow = grow & (self._eval(self._with(d, lo)) > log_y)
            if not bool(grow.any()):
                break
            lo = torch.where(grow, lo - w, lo)
        grow = torch.ones_like(x0, dtype=torch.bool)
        for _ in range(self.max_step_out):
            grow = grow & (self._eval(self._with(d, hi)) > log_y)
            if not bool(grow.any()):
                break
            hi = torch.where(grow, hi + w, hi)
        size = hi - lo
        # shrinkage
        todo = torch.ones_like(x0, dtype=torch.bool)
        new_x, new_lp = x0.clone(), self.lp.clone()
        for _ in range(self.max_shrink):
            prop = lo + (hi - lo) * self._rand()
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
        return size

    def sweep(self, tune: bool = False) -> None:
        for d in range(self.D):
            size = self._update_dim(d)
            if tune:   # running estimate of the typical slice width, as sbi's slice samplers do
                self.width[d] = 0.5 * self.width[d] + 0.5 * size.mean().clamp_min(1e-6)

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
