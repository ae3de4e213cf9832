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
    """utils.py:27-86"""

    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total = 0.0
        self.count = 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        t = t.tolist()
        self.count = int(t[0])
        self.total = t[1]

    @property
    def median(self):
        return float(np.median(list(self.deque))) if self.deque else 0.0

    @property
    def avg(self):
        return float(np.mean(list(self.deque))) if self.deque else 0.0

    @property
    def global_avg(self):
        return self.total / max(self.count, 1)

    @property
    def max(self):
        return max(self.deque) if self.deque else 0.0

    @property
    def value(self):
        return self.deque[-1] if self.deque else 0.0

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max, value=self.value)


class MetricLogger:
    """utils.py:89-170 (same meters / log_every behaviour; printing only)."""

    def __init__(self, delimiter="\t", quiet=False):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter
        self.quiet = quiet

    def update(self, **kwargs):
        for k, v in kwargs.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[k].update(v)

    def __getattr__(self, attr):
        if attr in self.meters:
            return self.meters[attr]
        if attr in self.__dict__:
            return self.__dict__[attr]
        raise AttributeError(attr)

    def __str__(self):
        return self.delimiter.join("{}: {}".format(n, str(m)) for n, m in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None):
        i = 0
        header = header or ''
        start = end = time.time()
        iter_time = SmoothedValue(fmt='{avg:.4f}')
        data_time = SmoothedValue(fmt='{avg:.4f}')
        n = len(iterable) if hasattr(iterable, "__len__") else -1
        for obj in iterable:
            data_time.update(time.time() - end)
            yield obj
            iter_time.update(time.time() - end)
            if not self.quiet and (i % print_freq == 0 or i == n - 1):
                eta = str(datetime.timedelta(seconds=int(iter_time.global_avg * max(n - i, 0))))
                mem = torch.cuda.max_memory_allocated() / (1024.0 * 1024.0) if torch.cuda.is_available() else 0
                print(self.delimiter.join([header, f"[{i}/{n}]", f"eta: {eta}", str(self), f"time: {iter_time}",
                                           f"data: {data_time}", f"max mem: {mem:.0f}"]))
            i += 1
            end = time.time()
        total = time.time() - start
        if not self.quiet:
            print('{} Total time: {} ({:.4f} s / it)'.format(header, str(datetime.timedelta(seconds=int(total))), total / max(n, 1)))


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
    """utils.py:391-408"""
    warmup_schedule = np.array([])
    warmup_iters = warmup_epochs * niter_per_ep
    if warmup_steps > 0:
        warmup_iters = warmup_steps
    if warmup_epochs > 0:
        warmup_schedule = np.linspace(start_warmup_value, base_value, warmup_iters)
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    schedule = np.array([final_value + 0.5 * (base_value - final_value) * (1 + math.cos(math.pi * i / (len(iters))))
                         for i in iters])
    schedule = np.concatenate((warmup_schedule, schedule))
    assert len(schedule) == epochs * niter_per_ep
    return schedule
