"""Host-side helpers with the reference's names (utils.py): the loss-scaler object the engine is handed, the
gradient norm, the cosine schedule and a small meter/logger.  Only what the pretraining hot loop touches.

``NativeScalerWithGradNormCount`` keeps the reference call contract (utils.py:347-373):
``norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=model.parameters())`` and
``state_dict()["scale"]``.  The B200 path computes in bf16 with fp32 accumulation, so no loss scaling is needed:
the scale is the constant 1.0.  When the loss comes from the fused step (``model.pretrain_step``) the gradients are
already in the flat arena and ``backward()`` is skipped; the norm is one pass of ``mofo_sq_norm_f32`` over the arena
instead of 218 ``torch.norm`` calls (utils.py:387).
"""
from __future__ import annotations

import datetime
import math
import time
from collections import defaultdict, deque

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


class SmoothedValue:
    """A meter with the attribute set the engine and the reference's scripts read (utils.py:27-86): ``update(v, n)``,
    ``median`` / ``avg`` over a sliding window, ``global_avg`` over everything seen, ``max``, ``value`` and a format string
    over those names.  Implemented on a running (count, total) pair plus a bounded window."""

    __slots__ = ("window", "total", "count", "fmt")

    def __init__(self, window_size=20, fmt=None):
        self.window = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = "{median:.4f} ({global_avg:.4f})" if fmt is None else fmt

    def update(self, value, n=1):
        self.window.append(value)
        self.total += value * n
        self.count += n

    def synchronize_between_processes(self):
        """Sums (count, total) over the ranks; the window stays local (as in the reference)."""
        if not is_dist_avail_and_initialized():
            return
        pair = torch.tensor([float(self.count), self.total], dtype=torch.float64,
                            device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.barrier()
        dist.all_reduce(pair)
        self.count, self.total = int(pair[0].item()), float(pair[1].item())

    def _stat(self, fn, empty=0.0):
        return float(fn(list(self.window))) if self.window else empty

    median = property(lambda self: self._stat(np.median))
    avg = property(lambda self: self._stat(np.mean))
    max = property(lambda self: self._stat(max))
    value = property(lambda self: self.window[-1] if self.window else 0.0)
    global_avg = property(lambda self: self.total / self.count if self.count else 0.0)

    @property
    def deque(self):                         # name used by callers of the reference class
        return self.window

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


class MetricLogger:
    """Named ``SmoothedValue`` meters plus the ``log_every`` generator the training loops wrap their loader in
    (utils.py:89-170).  ``quiet`` suppresses the periodic lines (benchmarks, tests)."""

    def __init__(self, delimiter="\t", quiet=False):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter, self.quiet = delimiter, quiet

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def update(self, **named_values):
        for name, v in named_values.items():
            if v is None:
                continue
            v = v.item() if isinstance(v, torch.Tensor) else v
            if not isinstance(v, (int, float)):
                raise TypeError(f"meter {name!r}: expected a number, got {type(v).__name__}")
            self.meters[name].update(v)

    def __getattr__(self, name):
        meters = self.__dict__.get("meters", {})
        if name in meters:
            return meters[name]
        raise AttributeError(name)

    def __str__(self):
        return self.delimiter.join(f"{name}: {meter}" for name, meter in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def log_every(self, iterable, print_freq, header=None):
        header = header or ""
        total = len(iterable) if hasattr(iterable, "__len__") else -1
        step_time, wait_time = SmoothedValue(fmt="{avg:.4f}"), SmoothedValue(fmt="{avg:.4f}")
        t_begin = t_prev = time.time()
        for i, item in enumerate(iterable):
            wait_time.update(time.time() - t_prev)
            yield item
            step_time.update(time.time() - t_prev)
            if not self.quiet and (i % print_freq == 0 or i == total - 1):
                remaining = datetime.timedelta(seconds=int(step_time.global_avg * max(total - i, 0)))
                peak_mb = torch.cuda.max_memory_allocated() / 2 ** 20 if torch.cuda.is_available() else 0
                print(self.delimiter.join([header, f"[{i}/{total}]", f"eta: {remaining}", str(self), f"time: {step_time}",
                                           f"data: {wait_time}", f"max mem: {peak_mb:.0f}"]))
            t_prev = time.time()
        if not self.quiet:
            elapsed = time.time() - t_begin
            print(f"{header} Total time: {datetime.timedelta(seconds=int(elapsed))} ({elapsed / max(total, 1):.4f} s / it)")


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    """utils.py:376-388 (L2 only on this path), one kernel per gradient tensor."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if len(grads) == 0:
        return torch.tensor(0.)
    assert float(norm_type) == 2.0
    acc = torch.zeros(1, dtype=torch.float32, device=grads[0].device)
    for g in grads:
        _lib.sq_norm_f32(g.contiguous(), acc)
    return acc.sqrt()[0]


class NativeScalerWithGradNormCount:
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scale = 1.0

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True,
                 arena=None, loss_guard=False, staged_sq_norm=None, return_sq=False):
        """``arena``: the model's flat gradient arena when the gradients were produced by the fused step (the engine
        passes it); otherwise ``loss.backward()`` runs autograd through the model's kernel backward.
        ``staged_sq_norm``: the squared-norm accumulator of ``FusedAdamW.begin_staged`` when the update already ran stage
        by stage behind the gradient exchange (mofo_b200/dp.py); only the norm is finished here."""
        if staged_sq_norm is not None:
            return staged_sq_norm if return_sq else staged_sq_norm.sqrt()[0]
        if arena is None:
            loss.backward(create_graph=create_graph)
        if not update_grad:
            return None
        clip = clip_grad is not None and clip_grad > 0
        if arena is not None and not clip and getattr(optimizer, "fused_mofo", False):
            # no clip coefficient to derive first: the optimizer kernel accumulates sum(g^2) while it reads the gradients
            acc = torch.zeros(1, dtype=torch.float32, device=arena.device)
            optimizer.step(loss_guard=loss.detach().reshape(-1)[:1] if loss_guard else None, sq_norm_out=acc)
            return acc if return_sq else acc.sqrt()[0]        # return_sq: the caller takes the root on the host
        if arena is not None:
            acc = torch.zeros(1, dtype=torch.float32, device=arena.device)
            _lib.sq_norm_f32(arena, acc)
            norm = acc.sqrt()[0]
        else:
            assert parameters is not None
            parameters = list(parameters)
            norm = get_grad_norm_(parameters)
        coef = None
        if clip_grad is not None and clip_grad > 0:          # engine passes max_norm (0 = off, as in the reference CLI)
            coef = torch.clamp(clip_grad / (norm + 1e-6), max=1.0).reshape(1)
        if getattr(optimizer, "fused_mofo", False):
            # one kernel: clip coefficient and the non-finite-loss guard are applied on the device
            optimizer.step(clip_coef=coef, loss_guard=loss.detach().reshape(-1)[:1] if loss_guard else None)
            return norm
        if coef is not None:
            if arena is not None:
                arena.mul_(coef)
            else:
                for p in parameters:
                    if p.grad is not None:
                        p.grad.mul_(coef)
        optimizer.step()
        return norm

    def state_dict(self):
        return {"scale": self._scale}

    def load_state_dict(self, state_dict):
        self._scale = float(state_dict.get("scale", 1.0)) if isinstance(state_dict, dict) else 1.0
        self._scale = 1.0


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep, warmup_epochs=0, start_warmup_value=0,
                     warmup_steps=-1):
    """Per-iteration schedule with the reference's contract (utils.py:391-408): a linear warm-up from ``start_warmup_value``
    to ``base_value`` over ``warmup_epochs * niter_per_ep`` iterations (``warmup_steps`` > 0 overrides the length; the ramp is
    only emitted when ``warmup_epochs`` > 0), then half a cosine from ``base_value`` to ``final_value``."""
    total = epochs * niter_per_ep
    n_warm = warmup_steps if warmup_steps > 0 else warmup_epochs * niter_per_ep
    ramp = np.linspace(start_warmup_value, base_value, n_warm) if warmup_epochs > 0 else np.array([])
    n_cos = total - n_warm
    phase = np.arange(n_cos) / max(n_cos, 1)
    cos = final_value + 0.5 * (base_value - final_value) * (1.0 + np.cos(np.pi * phase))
    schedule = np.concatenate((ramp, cos))
    assert len(schedule) == total
    return schedule
