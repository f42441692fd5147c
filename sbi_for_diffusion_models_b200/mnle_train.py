"""MNLE training on the device (reference mnle.py:16-50).

The reference hands ``(z_train, x_train)`` to ``sbi.inference.MNLE`` and calls
``trainer.train(training_batch_size=cfg.TRAIN_BATCH_SIZE)``.  What that does (sbi 0.25.0 defaults,
restated from its documentation -- sbi is not installable here, see ``oracle/mnle_spec.py``):

* per-dimension z-scoring of the condition and of ``log rt`` from the training set
  (``z_score_theta="independent"``, ``z_score_x="independent"``, ``log_transform_x=True``);
* one output class per distinct choice value in the training set;
* 10 % validation split, shuffled minibatches (incomplete last batch dropped), Adam(lr=5e-4) on
  ``-mean log_prob``, ``clip_grad_norm_(5.0)``, stop after 20 epochs without a better validation
  loss and return the best parameters.

Here the whole loop stays on the GPU: the standardised training set is resident in HBM, a
minibatch is a vector of row indices, and each step is ``mnle_train_nll_grad_f32`` (forward,
hand-written reverse mode, fixed-order gradient reduction) followed by ``mnle_train_adam_f32``; the
host reads one float per epoch (the validation loss).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import _native
from .mnle_net import COND_DIM, DeviceMNLE, PackedMNLE, packed_sections, unpack_params
from .simulator import compute_device


def init_raw_params(n_choices: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """``torch.nn.Linear`` default initialisation (weights and biases uniform in
    +-1/sqrt(fan_in)) for every layer of the estimator the reference builds (mnle.py:31-39)."""
    g = torch.Generator().manual_seed(int(seed))
    p: Dict[str, torch.Tensor] = {}
    for name, shape in packed_sections(n_choices):
        if not shape:
            continue
        if len(shape) == 2:
            fan_in = shape[1]
            bound = 1.0 / math.sqrt(fan_in)
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound   # bias: bound of its weight
    return p


def _column_mean_std(t: torch.Tensor, chunk: int = 1 << 22):
    """Per-column mean and unbiased std in float64 without a float64 copy of the whole (N, d) set
    (1e8 x 85 doubles would be 68 GB): shifted sums over row chunks."""
    n = t.shape[0]
    shift = t[:1].double()
    s1 = torch.zeros(t.shape[1], dtype=torch.float64, device=t.device)
    s2 = torch.zeros_like(s1)
    for a in range(0, n, chunk):
        d = t[a:a + chunk].double() - shift
        s1 += d.sum(0)
        s2 += (d * d).sum(0)
    mean = s1 / n
    var = (s2 - n * mean * mean) / max(n - 1, 1)
    return mean + shift[0], var.clamp_min(0.0).sqrt()


class MNLETrainer:
    """Device-resident parameters, Adam state and workspace of one estimator in training."""

    def __init__(self, n_choices: int, *, cond_mean, cond_std, mu_y: float, sigma_y: float,
                 init: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0, device=None):
        self.dev = compute_device(device)
        self.n_choices = int(n_choices)
        self.cond_mean = torch.as_tensor(cond_mean, dtype=torch.float64).reshape(COND_DIM).clone()
        self.cond_std = torch.as_tensor(cond_std, dtype=torch.float64).reshape(COND_DIM).clamp_min(1e-7).clone()
        raw = dict(init) if init is not None else init_raw_params(self.n_choices, seed)
        raw = {k: v for k, v in raw.items() if k not in ("cond_mean", "cond_std")}
        raw["flow.mu_y"], raw["flow.sigma_y"] = torch.tensor(float(mu_y)), torch.tensor(float(sigma_y))
        raw["cond_mean"], raw["cond_std"] = torch.zeros(COND_DIM), torch.ones(COND_DIM)   # identity fold
        packed = PackedMNLE.from_params(raw).packed
        with torch.cuda.device(self.dev):
            self.params = torch.from_numpy(packed).to(self.dev)
            self.grad = torch.zeros_like(self.params)
            self.m = torch.zeros_like(self.params)
            self.v = torch.zeros_like(self.params)
            self.stats = torch.zeros(2, dtype=torch.float32, device=self.dev)
        self._ws: Optional[torch.Tensor] = None
        self.step_count = 0

    # ---- data ---------------------------------------------------------------------------
    def standardise(self, z: torch.Tensor) -> torch.Tensor:
        """(N,85) condition -> float32 standardised copy on the device (what the first layers see)."""
        z = z.to(device=self.dev, dtype=torch.float32)
        return ((z - self.cond_mean.to(self.dev, torch.float32)) / self.cond_std.to(self.dev, torch.float32)).contiguous()

    def _workspace(self, R: int) -> torch.Tensor:
        need = _native.lib().mnle_train_workspace_floats(self.n_choices, R)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.float32, device=self.dev)
        return self._ws

    # ---- kernels --------------------------------------------------------------------------
    def nll(self, x: torch.Tensor, cond_std: torch.Tensor, idx: Optional[torch.Tensor] = None, *,
            grad: bool = True, forward: str = "tc") -> torch.Tensor:
        """stats (2,) on the device: [-mean log p over the minibatch, |grad|^2]; fills ``self.grad``
        when ``grad``.  ``x`` (N,2), ``cond_std`` (N,85) standardised, ``idx`` int64 rows or None.
        ``forward``: "tc" tensor cores (tcgen05), "fp32" CUDA cores (accuracy anchor)."""
        if forward not in ("tc", "fp32"):
            raise ValueError(f"unknown forward {forward!r}")
        if x.device != self.dev or cond_std.device != self.dev or x.dtype != torch.float32 or cond_std.dtype != torch.float32:
            raise ValueError("training data must be float32 tensors on the trainer's device")
        if x.ndim != 2 or x.shape[1] != 2 or not x.is_contiguous():
            raise ValueError(f"x must be contiguous (N,2), got {tuple(x.shape)}")
        if cond_std.ndim != 2 or cond_std.shape[1] != COND_DIM or cond_std.stride(1) != 1 or cond_std.shape[0] != x.shape[0]:
            raise ValueError(f"condition must be (N,{COND_DIM}) with N={x.shape[0]}, got {tuple(cond_std.shape)}")
        if idx is not None:
            if idx.dtype != torch.int64 or idx.device != self.dev or idx.ndim != 1 or not idx.is_contiguous():
                raise ValueError("idx must be a contiguous int64 vector on the trainer's device")
            R = idx.shape[0]
        else:
            R = x.shape[0]
        if R < 1:
            raise ValueError("empty minibatch")
        with torch.cuda.device(self.dev):
            ws = self._workspace(R)
            rc = _native.lib().mnle_train_nll_grad_f32(
                self.params.data_ptr(), self.n_choices, x.data_ptr(), cond_std.data_ptr(),
                cond_std.stride(0) if cond_std.shape[0] > 1 else COND_DIM, idx.data_ptr() if idx is not None else None, R,
                self.stats.data_ptr(), self.grad.data_ptr() if grad else None, ws.data_ptr(),
                1 if forward == "fp32" else 0, torch.cuda.current_stream(self.dev).cuda_stream)
        _native.check(rc, "mnle_train_nll_grad_f32")
        return self.stats

    def adam(self, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, clip_max_norm: Optional[float] = 5.0) -> None:
        self.step_count += 1
        with torch.cuda.device(self.dev):
            rc = _native.lib().mnle_train_adam_f32(
                self.params.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.n_choices,
                self.stats.data_ptr(), lr, betas[0], betas[1], eps, self.step_count,
                float(clip_max_norm) if clip_max_norm else 0.0, torch.cuda.current_stream(self.dev).cuda_stream)
        _native.check(rc, "mnle_train_adam_f32")

    # ---- results ---------------------------------------------------------------------------
    def raw_params(self) -> Dict[str, torch.Tensor]:
        """CPU copies by name (``oracle/mnle_spec.py`` naming) including the z-scoring buffers."""
        p = {k: v.detach().cpu().clone() for k, v in unpack_params(self.params, self.n_choices).items()}
        p["cond_mean"], p["cond_std"] = self.cond_mean.float(), self.cond_std.float()
        return p

    def named_grads(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach().cpu().clone() for k, v in unpack_params(self.grad, self.n_choices).items()}

    def estimator(self) -> DeviceMNLE:
        """Inference object (z-scoring folded into the first layers; tcgen05 potential)."""
        return DeviceMNLE(PackedMNLE.from_params(self.raw_params()), device=self.dev)


def train_mnle(cfg, proposal_z, z_train: torch.Tensor, x_train: torch.Tensor, device: str = "cpu", *,
               learning_rate: float = 5e-4, validation_fraction: float = 0.1, stop_after_epochs: int = 20,
               max_num_epochs: int = 2 ** 31 - 1, clip_max_norm: Optional[float] = 5.0, seed: int = 0,
               init: Optional[Dict[str, torch.Tensor]] = None, show_train_summary: bool = False,
               return_summary: bool = False):
    """Drop-in for the reference's ``train_mnle`` (mnle.py:16-50): same positional arguments, returns
    an estimator whose ``log_prob(x, condition=...)`` the potentials consume.  ``device`` is accepted
    for signature compatibility; the work runs on the current CUDA device (no CPU fallback).
    ``proposal_z`` is unused, as in sbi's likelihood training."""
    if not bool(getattr(cfg, "SBI_LOG_TRANSFORM_X", True)):
        raise NotImplementedError("the device trainer implements the reference's default SBI_LOG_TRANSFORM_X=True "
                                  "(raw rt in x, log transform inside the estimator)")
    dev = compute_device(device if str(device).startswith("cuda") else None)
    if z_train.ndim != 2 or z_train.shape[1] != COND_DIM:
        raise ValueError(f"z_train must be (N,{COND_DIM}), got {tuple(z_train.shape)}")
    if x_train.ndim != 2 or x_train.shape[1] != 2 or x_train.shape[0] != z_train.shape[0]:
        raise ValueError(f"x_train must be (N,2) with N={z_train.shape[0]}, got {tuple(x_train.shape)}")
    N = z_train.shape[0]
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    with torch.cuda.device(dev):
        x = x_train.to(device=dev, dtype=torch.float32).contiguous()
        z = z_train.to(device=dev, dtype=torch.float32)
        if not bool(torch.isfinite(x).all()) or not bool(torch.isfinite(z).all()):
            raise ValueError("training data contains non-finite values")
        if not bool((x[:, 0] > 0).all()):
            raise ValueError("rt must be positive (log_transform_x)")
        choices = torch.unique(x[:, 1])
        n_choices = int(choices.numel())
        if not torch.equal(choices.cpu(), torch.arange(n_choices, dtype=torch.float32)):
            raise ValueError(f"choices must be the integers 0..K-1, got {choices.cpu().tolist()}")
        cond_mean, cond_std = _column_mean_std(z)
        y_mean, y_std = _column_mean_std(x[:, :1].log())
        z_score_x = getattr(cfg, "Z_SCORE_X", "independent")
        mu_y, sigma_y = (float(y_mean), float(y_std)) if z_score_x not in (None, "none") else (0.0, 1.0)
        tr = MNLETrainer(n_choices, cond_mean=cond_mean.cpu(), cond_std=cond_std.cpu(), mu_y=mu_y, sigma_y=max(sigma_y, 1e-7),
                         init=init, seed=seed, device=dev)
        cond = tr.standardise(z)
        del z

        perm = torch.randperm(N, device=dev, generator=gen)
        n_val = int(validation_fraction * N)
        n_train = N - n_val
        train_idx, val_idx = perm[:n_train].contiguous(), perm[n_train:].contiguous()
        batch = min(int(cfg.TRAIN_BATCH_SIZE), n_train)
        n_batches = n_train // batch

        def validation_loss() -> float:
            if n_val == 0:
                return float("nan")
            tot, chunk = 0.0, 65536
            for a in range(0, n_val, chunk):
                idx = val_idx[a:a + chunk]
                tot += float(tr.nll(x, cond, idx, grad=False)[0]) * idx.shape[0]
            return tot / n_val

        best, best_params, since, epoch = float("inf"), tr.params.clone(), 0, 0
        history = []
        while epoch < max_num_epochs:
            order = train_idx[torch.randperm(n_train, device=dev, generator=gen)]
            train_loss = torch.zeros((), device=dev)
            for b in range(n_batches):
                stats = tr.nll(x, cond, order[b * batch:(b + 1) * batch])
                train_loss += stats[0]
                tr.adam(lr=learning_rate, clip_max_norm=clip_max_norm)
            epoch += 1
            val = validation_loss()
            history.append((float(train_loss) / max(n_batches, 1), val))
            if show_train_summary:
                print(f"epoch {epoch}: train {history[-1][0]:.4f}  validation {val:.4f}", flush=True)
            if n_val == 0:
                continue
            if val < best:
                best, since = val, 0
                best_params.copy_(tr.params)
            else:
                since += 1
                if since >= stop_after_epochs:
                    break
        if n_val > 0:
            tr.params.copy_(best_params)
        est = tr.estimator()
    if return_summary:
        return est, {"epochs": epoch, "best_validation_loss": best, "history": history, "n_choices": n_choices,
                     "steps": tr.step_count}
    return est
