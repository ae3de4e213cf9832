"""Runs the UNMODIFIED reference (baseline/_ref, staged by setup_ref.py) through its own public entry points:
``timm.create_model`` -> ``optim_factory.create_optimizer`` -> ``utils.NativeScalerWithGradNormCount`` ->
``engine_for_pretraining.train_one_epoch_BB`` (optionally under its own DistributedDataParallel wrapper,
run_mae_pretraining_BB.py:229-231), on synthetic batches shaped like what its DataLoader yields
(``videos f32 [B,3,16,224,224]``, ``bbox int64 [B,16,4]``, ``mask f64 [B,1568]`` from its own
``TubeMaskingGenerator_BB`` after ``np.random.seed(10)``, transforms.py:139 / datasets.py:56-58).

Used by bench.py (``gpu_reference`` record = the Blackwell library kernels the reference already reaches, SURVEY §8d;
``--impl reference`` = the same call on the host cores) and by the parity tests.  Nothing here is on the product path.

The only thing changed at run time is the autocast dtype: the reference hard-codes ``torch.cuda.amp.autocast()``
(fp16, engine_for_pretraining.py:299); ``amp="bf16"`` / ``"fp32"`` swap that attribute for the duration of the call
(no reference source is edited).
"""
from __future__ import annotations

import contextlib
import functools
import importlib
import io
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_MODULES = ("masking_generator", "modeling_finetune", "modeling_pretrain", "utils", "optim_factory", "engine_for_pretraining")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + ".py")) for m in _MODULES)


_ns = None


def load():
    """Imports the staged reference modules under their own names (they import each other that way)."""
    global _ns
    if _ns is not None:
        return _ns
    if not available():
        raise RuntimeError("baseline/_ref is not staged: run `python baseline/setup_ref.py` where /root/reference exists")
    sys.path.insert(0, HERE)
    try:
        import shims
    finally:
        sys.path.pop(0)
    shims.install()
    for m in _MODULES:
        if m in sys.modules and not getattr(sys.modules[m], "__file__", "").startswith(REF_DIR):
            raise RuntimeError(f"module name {m!r} is already bound to {sys.modules[m].__file__}; cannot load the reference")
    sys.path.insert(0, REF_DIR)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            mods = {m: importlib.import_module(m) for m in _MODULES}
    finally:
        sys.path.remove(REF_DIR)
    # entry points captured now: a later import of the B200 look-alike may register the same names with timm's registry
    entry = {n: getattr(mods["modeling_pretrain"], n) for n in ("pretrain_mae_small_patch16_224",
             "pretrain_videomae_base_patch16_224", "pretrain_videomae_large_patch16_224")}

    def create_model(model_name, pretrained=False, **kwargs):
        """timm.create_model for the reference's registry entries (timm 0.4.12 drops None-valued kwargs)."""
        return entry[model_name](pretrained=pretrained, **{k: v for k, v in kwargs.items() if v is not None})

    _ns = types.SimpleNamespace(create_model=create_model, **mods)
    return _ns


def load_finetune_engine():
    """The reference's engine_for_finetuning (and the mixup module it imports), loaded like the modules above."""
    load()
    if "engine_for_finetuning" in sys.modules and getattr(sys.modules["engine_for_finetuning"], "__file__", "").startswith(REF_DIR):
        return sys.modules["engine_for_finetuning"]
    sys.path.insert(0, REF_DIR)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return importlib.import_module("engine_for_finetuning")
    finally:
        sys.path.remove(REF_DIR)


def create_model(name="pretrain_videomae_base_patch16_224"):
    """run_mae_pretraining_BB.py:138-148 (get_model)."""
    return load().create_model(name, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=4)


def create_optimizer(model, lr, weight_decay=0.05):
    """run_mae_pretraining_BB.py:233-234 with the CLI defaults (:53-63)."""
    args = types.SimpleNamespace(opt="adamw", opt_eps=1e-8, opt_betas=(0.9, 0.95), weight_decay=weight_decay, lr=lr, momentum=0.9)
    with contextlib.redirect_stdout(io.StringIO()):
        return load().optim_factory.create_optimizer(args, model)


def synthetic_boxes(B, rng):
    """integer w,h ~ U{32..160}, x1 ~ U{0..224-w}, y1 ~ U{0..224-h}; one box replicated to 16 frames."""
    w = rng.integers(32, 161, B); h = rng.integers(32, 161, B)
    x1 = (rng.random(B) * (224 - w + 1)).astype(np.int64); y1 = (rng.random(B) * (224 - h + 1)).astype(np.int64)
    return np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float64)


def reference_masks(boxes, grid=(8, 14, 14), mask_ratio=0.9, mask_ratio_bb=0.75):
    """The reference's own generator, called the way its dataset does: np.random.seed(10) then __call__(bbox[16,4])."""
    ref = load()
    gen = ref.masking_generator.TubeMaskingGenerator_BB(grid, mask_ratio, mask_ratio_bb)
    out = []
    state = np.random.get_state()
    try:
        for b in range(len(boxes)):
            np.random.seed(10)
            out.append(gen(np.repeat(boxes[b][None], 16, 0)))
    finally:
        np.random.set_state(state)
    return np.stack(out)


def synthetic_batches(B, n, seed, device, pin=False, size=224):
    """n batches as the reference's DataLoader would collate them; videos on ``device`` (or pinned host memory)."""
    mean = torch.tensor((0.485, 0.456, 0.406))[None, :, None, None, None]
    std = torch.tensor((0.229, 0.224, 0.225))[None, :, None, None, None]
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    batches = []
    for _ in range(n):
        vid = ((torch.rand(B, 3, 16, size, size, generator=g) - mean) / std).contiguous()
        boxes = synthetic_boxes(B, rng)
        mask = torch.from_numpy(reference_masks(boxes))
        bbox = torch.from_numpy(boxes).long()[:, None, :].expand(B, 16, 4).contiguous()
        if pin:
            vid, mask = vid.pin_memory(), mask.pin_memory()
        else:
            vid, mask = vid.to(device), mask.to(device)
        batches.append((vid, bbox, mask))
    return batches


def make_scaler(device):
    """utils.NativeScalerWithGradNormCount() (run_mae_pretraining_BB.py:235).  Its torch.cuda.amp.GradScaler() is bound to
    the CPU backend when the run is on the host cores (a CUDA-bound scaler disables itself without a GPU and then has no
    "scale" for engine_for_pretraining.py:427 to read)."""
    ref = load()
    orig = torch.cuda.amp.GradScaler
    if torch.device(device).type == "cpu":
        torch.cuda.amp.GradScaler = functools.partial(torch.amp.GradScaler, "cpu")
    try:
        with _quiet():
            return ref.utils.NativeScalerWithGradNormCount()
    finally:
        torch.cuda.amp.GradScaler = orig


class _Loader:
    def __init__(self, batches, n):
        self.batches, self.n = batches, n

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            yield self.batches[i % len(self.batches)]


@contextlib.contextmanager
def autocast_mode(amp):
    """amp: 'fp16' (as authored), 'bf16' or 'fp32' (autocast disabled)."""
    orig = torch.cuda.amp.autocast
    if amp == "bf16":
        torch.cuda.amp.autocast = functools.partial(torch.autocast, "cuda", dtype=torch.bfloat16)
    elif amp == "fp32":
        torch.cuda.amp.autocast = functools.partial(torch.autocast, "cuda", enabled=False)
    elif amp != "fp16":
        raise ValueError(amp)
    try:
        yield
    finally:
        torch.cuda.amp.autocast = orig


@contextlib.contextmanager
def _quiet():
    import warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        yield


def train_epoch(model, optimizer, scaler, batches, steps, device, amp="fp16", lr_values=None, log_writer=None):
    """One call of the reference's train_one_epoch_BB over ``steps`` synthetic batches; returns its stats dict.
    max_norm=None is what run_mae_pretraining_BB.py:279 passes (args.clip_grad, default None :59)."""
    ref = load()
    sync = torch.cuda.synchronize
    if not torch.cuda.is_available():           # engine_for_pretraining.py:429 needs a CUDA runtime; host-only boxes have none
        torch.cuda.synchronize = lambda *a, **k: None
    try:
        return _train_epoch(ref, model, optimizer, scaler, batches, steps, device, amp, lr_values, log_writer)
    finally:
        torch.cuda.synchronize = sync


def _train_epoch(ref, model, optimizer, scaler, batches, steps, device, amp, lr_values, log_writer):
    with autocast_mode(amp), _quiet():
        return ref.engine_for_pretraining.train_one_epoch_BB(
            model, _Loader(batches, steps), optimizer, torch.device(device), 0, scaler, max_norm=None, patch_size=16,
            normlize_target=True, log_writer=log_writer, lr_scheduler=None, start_steps=0, lr_schedule_values=lr_values,
            wd_schedule_values=None)


def run(device, model_name="pretrain_videomae_base_patch16_224", batch=32, steps=50, warmup=10, amp="fp16", host_inputs=False,
        ddp=False, seed=0, lr=None, pool=2):
    """Timed reference run.  Returns {"clips_per_s", "ms_per_step", "loss", ...} for this rank (caller max-reduces)."""
    import torch.distributed as dist
    ref = load()
    device = torch.device(device)
    world = dist.get_world_size() if ddp else 1
    rank = dist.get_rank() if ddp else 0
    torch.manual_seed(seed)
    with _quiet():
        model = create_model(model_name).to(device)
    core = model
    if ddp:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], find_unused_parameters=True)
    lr = lr if lr is not None else 1.5e-4 * batch * world / 256
    opt = create_optimizer(core, lr)
    scaler = make_scaler(device)
    batches = synthetic_batches(batch, pool, 1234 + rank, device, pin=host_inputs)
    cuda = device.type == "cuda"
    if warmup:
        train_epoch(model, opt, scaler, batches, warmup, device, amp)
    if cuda:
        torch.cuda.synchronize(device)
        if ddp:
            dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    t0 = time.perf_counter()
    stats = train_epoch(model, opt, scaler, batches, steps, device, amp)
    if cuda:
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1)
    else:
        ms = 1e3 * (time.perf_counter() - t0)
    out = {"clips_per_s": batch * steps / (ms * 1e-3), "ms_per_step": ms / steps, "loss": float(stats["loss"]),
           "amp": amp, "inputs": "pinned host" if host_inputs else "device-resident", "steps": steps, "warmup": warmup,
           "batch": batch, "model": model_name}
    del model, opt, scaler, batches
    if cuda:
        torch.cuda.empty_cache()
    return out


def single_step(model, batch, device, amp="fp32"):
    """ONE iteration of the reference's train_one_epoch_BB on ``batch`` with its own optimizer at lr = 0 (AdamW then leaves
    every parameter bit-identical: p*(1-0*wd) - 0*update), so the call's side effects are exactly what parity needs:
    the loss it logs, the model output (forward hook) and the unscaled gradients left in ``p.grad``."""
    core = model.module if hasattr(model, "module") else model
    opt = create_optimizer(core, lr=0.0)
    scaler = make_scaler(device)
    grabbed = {}
    h = core.register_forward_hook(lambda m, i, o: grabbed.__setitem__("out", o.detach()))
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    if amp == "fp32":
        torch.backends.cudnn.allow_tf32 = False          # the Conv3d patch embedding would otherwise run in TF32
        torch.backends.cuda.matmul.allow_tf32 = False
    try:
        stats = train_epoch(model, opt, scaler, [batch], 1, device, amp)
    finally:
        h.remove()
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    grads = {n: p.grad.detach().clone() for n, p in core.named_parameters()}
    return {"loss": float(stats["loss"]), "out": grabbed["out"], "grads": grads, "grad_norm": float(stats["grad_norm"])}


class LossLog:
    """log_writer stand-in (engine_for_pretraining.py:452-459): keeps the per-step loss."""

    def __init__(self):
        self.losses = []

    def update(self, head="scalar", **kw):
        if "loss" in kw:
            self.losses.append(float(kw["loss"]))

    def set_step(self, step=None):
        pass
