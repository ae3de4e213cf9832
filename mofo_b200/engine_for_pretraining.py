"""B200-native look-alike of the reference's ``engine_for_pretraining.py``.

``train_one_epoch_BB`` keeps the reference signature and return value (engine_for_pretraining.py:215-218,468) and the
per-step protocol of its loop body (:228-464): table-driven lr/wd, H2D of the batch, target construction + model +
MSE, non-finite-loss exit, optimizer step through ``loss_scaler``, the same meters.  What changes is where the work
runs: masking index lists, target (un-normalise / patchify / per-tube normalise, :258-288), forward, loss (:299-304)
and backward are ONE fused sequence of sm_100a kernels (``model.pretrain_step``), gradients land in a flat arena, and
the data-parallel mean is an NCCL all-reduce of arena slices overlapped with backward (mofo_b200/dp.py) instead of
DDP's reducer.  The dead ``video_masks`` work of :243-249 (result unused) is not performed.

Two paths, chosen per call:
  * fused (default)  — ``loss_scaler`` is ``mofo_b200.utils.NativeScalerWithGradNormCount`` (bf16: scale == 1).
  * autograd-compat  — any other scaler object (e.g. the reference's GradScaler-based one): labels come from the
    target kernel, ``model(videos, mask)`` runs through autograd (kernel backward), and the scaler drives
    ``backward()`` / ``step()`` exactly as in the reference; DDP wrappers work unchanged on this path.
Aliases for the two names ``run_mae_pretraining_BB.py`` calls but the reference never defines are exported.

Host / device pipelining (fused path with ``FusedAdamW``): the reference reads ``loss.item()`` and calls
``torch.cuda.synchronize()`` every step (:306, :429), which leaves the GPU idle while Python prepares the next step.
Here the loss and the gradient norm of step i are copied to pinned host memory asynchronously and READ WHILE STEP i+1
RUNS: every step's values still reach the meters / log writer, and a non-finite loss still ends training with the
reference's message and ``sys.exit(1)`` - one step later, and the device-side guard of the optimizer kernel has
already skipped that step's update.  ``MOFO_SYNC_EVERY_STEP=1`` restores the strict per-step synchronisation.
"""
from __future__ import annotations

import math
import os
import sys
from typing import Iterable

import torch
import torch.nn as nn

from . import _lib
from . import utils as _utils
from .dp import GradSync

__all__ = ["train_one_epoch_BB", "train_one_epoch_BB_no_global_union_gradual", "build_labels", "fused_step", "masks_from_bbox"]


def _core(model):
    return model.module if hasattr(model, "module") else model


def build_labels(videos, msk_idx, normalize_target=True):
    """labels f32 [B, N_mask, 1536] from the target kernel (engine_for_pretraining.py:258-288)."""
    B = videos.shape[0]
    n = msk_idx.shape[1]
    out = torch.empty(B * n, 1536, dtype=torch.float32, device=videos.device)
    _lib.target_mse(videos, msk_idx, None, None, None, None, normalize_target, 1.0, out)
    return out.view(B, n, 1536)


_MASK_GEN = {}


def masks_from_bbox(bbox, device, grid=(8, 14, 14), mask_ratio=0.9, mask_ratio_bb=0.75, seed=10):
    """Tube masks for a batch, generated ON THE GPU from the batch's motion boxes: (mask uint8 [B,N], vis_idx, msk_idx).

    The reference builds the mask in DataLoader workers (datasets.py:56-58 -> masking_generator.py:43-85) right after
    ``np.random.seed(10)`` (transforms.py:139), i.e. the mask is a deterministic function of the clip's first-frame box.
    CUDA is not usable in forked workers, so instead of moving the generator there this reproduces exactly what the
    worker would have returned - same box predicate, same MT19937 word stream of seed ``seed`` for every clip - from the
    ``bbox`` tensor the batch already carries ([B,16,4] or [B,4]; float boxes reproduce the worker bit for bit, the
    truncated LongTensor of kinetics.py:1064 reproduces what the generator returns for the truncated box)."""
    from . import masking_generator as mg
    key = (tuple(grid), float(mask_ratio), float(mask_ratio_bb), int(seed), str(device))
    ent = _MASK_GEN.get(key)
    if ent is None:
        gen = mg.TubeMaskingGenerator_BB(grid, mask_ratio, mask_ratio_bb, device=device)
        words = torch.from_numpy(mg.mt19937_words(seed, mg.words_per_clip(grid[1], grid[2])).view("int32")).to(device)
        ent = _MASK_GEN[key] = (gen, words)
    gen, words = ent
    B = bbox.shape[0]
    mask, vis_idx, msk_idx, _ = gen.generate_batch(bbox, words[None, :].expand(B, -1).contiguous())
    return mask, vis_idx, msk_idx


class _CudaPrefetcher:
    """Wraps the data loader: the H2D copies of batch i+1 (pinned host memory -> device, `non_blocking`) are issued on
    a copy stream while step i computes, so `videos.to(device)` (engine_for_pretraining.py:239-240) costs no step time
    when the host keeps up.  Yields batches whose tensors already live on `device`.

    The device side is two fixed staging slots per tensor (no allocator traffic, no `record_stream` bookkeeping): the
    copy into a slot waits for an event recorded on the compute stream once the step that consumed the slot's previous
    content has been enqueued."""

    def __init__(self, loader, device, mask_fn=None):
        self.loader, self.device = loader, device
        self.mask_fn = mask_fn          # GPU-mask mode: (videos, bbox) -> (vis_idx, msk_idx), run on the copy stream one batch ahead
        self.quiet = getattr(loader, "quiet", False)
        self.raw = bool(getattr(loader, "mofo_raw_frames", False))      # batches of raw uint8 frames (transforms.RawClipLoader)
        self.pre = getattr(loader, "preprocessor", None)
        self.slots = [{}, {}]           # per slot: field index -> device tensor
        self.consumed = [None, None]    # per slot: event after the consumer's work (compute stream)

    def __len__(self):
        return len(self.loader)

    def _stage(self, slot, k, t, stream):
        if not (isinstance(t, torch.Tensor) and t.device.type == "cpu"):
            return t
        buf = self.slots[slot].get(k)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            self.slots[slot][k] = buf
        with torch.cuda.stream(stream):
            buf.copy_(t, non_blocking=True)
        return buf

    def _load(self, it, stream, slot):
        try:
            videos, bbox, mask = next(it)
        except StopIteration:
            return None
        if self.consumed[slot] is not None:
            stream.wait_event(self.consumed[slot])
        if self.raw:
            # raw uint8 frames [B,T,H,W,3] + boxes + crops: H2D of the frames, then crop / resize / normalise / box transform
            # (mofo_clip_preprocess) on the copy stream, into this slot's clip buffer
            frames = self._stage(slot, 0, videos, stream)
            B, T = frames.shape[0], frames.shape[1]
            clip = self.slots[slot].get("clip")
            if clip is None or clip.shape[0] != B or clip.shape[2] != T:
                clip = torch.empty(B, 3, T, self.pre.size, self.pre.size, dtype=torch.float32, device=self.device)
                self.slots[slot]["clip"] = clip
            with torch.cuda.stream(stream):
                vid, boxes = self.pre(frames, bbox, mask, out=clip)
                boxes.record_stream(torch.cuda.current_stream(self.device))
            return vid, boxes, self._masks_ahead(vid, boxes, stream)
        vid = self._stage(slot, 0, videos, stream)
        if self.mask_fn is not None:
            return vid, bbox, self._masks_ahead(vid, bbox, stream)
        return vid, bbox, self._stage(slot, 2, mask, stream)

    def _masks_ahead(self, vid, bbox, stream):
        """GPU-mask mode: the (sequential, ~45 us) mask kernel of the NEXT batch runs on the copy stream beside the current
        step instead of at the head of its own step."""
        if self.mask_fn is None or not (isinstance(vid, torch.Tensor) and vid.is_cuda):
            return None
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(stream):
            vis_idx, msk_idx = self.mask_fn(vid, bbox)
            for t in (vis_idx, msk_idx):
                t.record_stream(main)
        return ("mofo_idx", vis_idx, msk_idx)

    def __iter__(self):
        stream = torch.cuda.Stream(device=self.device)
        stream.wait_stream(torch.cuda.current_stream(self.device))
        it = iter(self.loader)
        i = 0
        nxt = self._load(it, stream, 0)
        while nxt is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_stream(stream)                          # batch i has landed
            batch = nxt
            nxt = self._load(it, stream, (i + 1) & 1)        # overlaps with the step the caller is about to run
            yield batch
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))   # the caller has enqueued everything that reads batch i
            self.consumed[i & 1] = ev
            i += 1


def fused_step(core, optimizer, loss_scaler, sync, videos, vis_idx, msk_idx, normalize_target=True, max_norm=0, return_sq=False):
    """One fused training step on device-resident inputs: forward + target/MSE + backward (CUDA graphs), the gradient
    exchange overlapped with backward, gradient norm and the optimizer update.  Returns (loss, grad_norm) as CUDA tensors
    without synchronising.  With ``FusedAdamW`` and no clipping the update runs stage by stage on the exchange stream
    right behind each slice's all-reduce (``FusedAdamW.begin_staged``); with any other optimizer (or clipping, which
    needs the whole norm first) the scaler drives ``optimizer.step()`` after the last slice, as utils.py:353-367 does.
    With a non-fused optimizer the caller must check the loss before the update (the engine does)."""
    runner = core._runner
    runner._ensure_device(videos.device)
    arena = runner.grad_arena()
    fused_opt = getattr(optimizer, "fused_mofo", False)
    if fused_opt:
        optimizer.attach(core)
    sync.sync_parameters(core, optimizer)                   # first step only: rank 0's weights everywhere (DDP ctor)
    clip = max_norm is not None and max_norm > 0
    # measured on one B200 (bench.py, B=32): 13.49 ms/step staged vs 13.26 ms with the single launch after backward - with no
    # exchange to hide behind, the side-stream update only competes with backward for HBM; so it is on for world > 1 only
    staged = fused_opt and not clip and os.environ.get("MOFO_STAGED_OPT", "1" if sync.world > 1 else "0") == "1"
    on_stage = acc = None
    if staged:
        on_stage, acc = optimizer.begin_staged(loss_guard=runner.buf("loss", (1,), torch.float32))
    sync.begin(arena, runner.stage_end, on_stage=on_stage)
    loss = core.pretrain_step(videos, vis_idx=vis_idx, msk_idx=msk_idx, normalize_target=normalize_target,
                              grad_scale=sync.grad_scale, zero_grad=True,
                              stage_done=sync.stage_done if (sync.world > 1 or on_stage is not None) else None)
    sync.finish()
    if not fused_opt:
        return loss, None
    # return_sq: the SQUARED norm comes back when the scaler can provide it (no clipping) - the engine's pipelined path
    # takes the root on the host instead of launching a one-element sqrt kernel every step
    grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=None, arena=arena, loss_guard=True,
                            staged_sq_norm=acc, return_sq=return_sq and not clip)
    return loss, grad_norm


def train_one_epoch_BB(model: torch.nn.Module, data_loader: Iterable, optimizer: torch.optim.Optimizer,
                       device: torch.device, epoch: int, loss_scaler, max_norm: float = 0, patch_size: int = 16,
                       normlize_target: bool = True, log_writer=None, lr_scheduler=None, start_steps=None,
                       lr_schedule_values=None, wd_schedule_values=None, loss_weight=None):
    model.train()
    metric_logger = _utils.MetricLogger(delimiter="  ", quiet=getattr(data_loader, "quiet", False))
    metric_logger.add_meter('lr', _utils.SmoothedValue(window_size=1, fmt='{value:.6f}'))
    metric_logger.add_meter('min_lr', _utils.SmoothedValue(window_size=1, fmt='{value:.6f}'))
    header = 'Epoch: [{}]'.format(epoch)
    print_freq = 10
    if patch_size != 16:
        raise NotImplementedError("the target kernel is specialised for 16x16 patches, tubelet 2")
    core = _core(model)
    fused = isinstance(loss_scaler, _utils.NativeScalerWithGradNormCount)
    train_one_epoch_BB.last_path = "fused" if fused else "autograd-compat"      # which of the two paths this call took
    sync = GradSync() if fused else None
    start_steps = start_steps or 0
    pipelined = (fused and getattr(optimizer, "fused_mofo", False) and torch.device(device).type == "cuda"
                 and os.environ.get("MOFO_SYNC_EVERY_STEP", "0") != "1")
    gpu_masks = fused and (os.environ.get("MOFO_GPU_MASKS", "0") == "1" or bool(getattr(data_loader, "mofo_gpu_masks", False)))
    gpu_mask_ratios = getattr(data_loader, "mofo_mask_ratios", (0.9, 0.75))     # (mask_ratio, mask_ratio_BB) of the run
    pending = None                      # (event, pinned [loss, grad_norm], host-side values) of the step in flight
    if pipelined:
        pinned = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(2)]   # loss, grad norm, bad mask rows
        events = [torch.cuda.Event() for _ in range(2)]

    def record(loss_value, grad_norm, loss_scale_value, max_lr, min_lr, weight_decay_value):
        if not math.isfinite(loss_value):
            print("Loss is {}, stopping training".format(loss_value))
            sys.exit(1)
        metric_logger.update(loss=loss_value)
        metric_logger.update(loss_scale=loss_scale_value)
        metric_logger.update(lr=max_lr)
        metric_logger.update(min_lr=min_lr)
        metric_logger.update(weight_decay=weight_decay_value)
        metric_logger.update(grad_norm=grad_norm)
        if log_writer is not None:
            log_writer.update(loss=loss_value, head="loss")
            log_writer.update(loss_scale=loss_scale_value, head="opt")
            log_writer.update(lr=max_lr, head="opt")
            log_writer.update(min_lr=min_lr, head="opt")
            log_writer.update(weight_decay=weight_decay_value, head="opt")
            log_writer.update(grad_norm=grad_norm, head="opt")
            log_writer.set_step()

    def flush(p):
        ev, pin, rest, squared = p
        ev.synchronize()
        core.check_mask_rows(int(pin[2]))        # the reference raises on that step (x[~mask].reshape, modeling_pretrain.py:90)
        record(float(pin[0]), math.sqrt(max(float(pin[1]), 0.0)) if squared else float(pin[1]), *rest)

    def group_stats():
        min_lr, max_lr = 10., 0.
        for group in optimizer.param_groups:
            min_lr = min(min_lr, group["lr"])
            max_lr = max(max_lr, group["lr"])
        weight_decay_value = None
        for group in optimizer.param_groups:
            if group["weight_decay"] > 0:
                weight_decay_value = group["weight_decay"]
        return max_lr, min_lr, weight_decay_value

    def gpu_mask_fn(videos, bbox):
        pe = core.encoder.patch_embed
        grid = (videos.shape[2] // pe.tubelet_size, videos.shape[3] // pe.patch_size[0], videos.shape[4] // pe.patch_size[1])
        return masks_from_bbox(bbox, videos.device, grid, *gpu_mask_ratios)[1:]

    if torch.device(device).type == "cuda":
        data_loader = _CudaPrefetcher(data_loader, torch.device(device), mask_fn=gpu_mask_fn if gpu_masks else None)
    for step, batch in enumerate(metric_logger.log_every(data_loader, print_freq, header)):
        it = start_steps + step                                                         # :230
        if lr_schedule_values is not None or wd_schedule_values is not None:
            for i, param_group in enumerate(optimizer.param_groups):
                if lr_schedule_values is not None:
                    param_group["lr"] = lr_schedule_values[it] * param_group.get("lr_scale", 1.0)
                if wd_schedule_values is not None and param_group["weight_decay"] > 0:
                    param_group["weight_decay"] = wd_schedule_values[it]

        videos, bbox, bool_masked_pos = batch                                           # :238
        videos = videos.to(device, non_blocking=True)
        if not gpu_masks:
            bool_masked_pos = bool_masked_pos.to(device, non_blocking=True).flatten(1).to(torch.bool)

        if fused:
            core._runner._ensure_device(videos.device)
            arena = core._runner.grad_arena()
            if gpu_masks:
                # opt-in (MOFO_GPU_MASKS=1 or data_loader.mofo_gpu_masks): the loader's mask is ignored and the masks are
                # generated on the GPU from the batch's boxes - lets the DataLoader keep num_workers > 0 with a no-op
                # mask transform (CUDA cannot run in forked workers), see masks_from_bbox; the prefetcher has normally
                # produced them one batch ahead on the copy stream
                if isinstance(bool_masked_pos, tuple) and bool_masked_pos[0] == "mofo_idx":
                    vis_idx, msk_idx = bool_masked_pos[1], bool_masked_pos[2]
                else:
                    vis_idx, msk_idx = gpu_mask_fn(videos, bbox)
            else:
                vis_idx, msk_idx = core.indices_from_mask(bool_masked_pos)
            sq = pipelined and not (max_norm is not None and max_norm > 0)
            loss, grad_norm = fused_step(core, optimizer, loss_scaler, sync, videos, vis_idx, msk_idx, normlize_target, max_norm,
                                         return_sq=sq)
            if getattr(optimizer, "fused_mofo", False):
                # the fused optimizer skips the update on the device when the loss is not finite, so the whole step
                # (incl. the parameter update) is enqueued before the single host read of the loss
                if pipelined:
                    slot = step & 1
                    pinned[slot][0:1].copy_(loss.detach().reshape(1), non_blocking=True)    # :306, read one step later
                    pinned[slot][1:2].copy_(grad_norm.detach().reshape(1), non_blocking=True)
                    if core._bad_rows is not None:
                        pinned[slot][2:3].copy_(core._bad_rows, non_blocking=True)
                    events[slot].record()
                    mine = (events[slot], pinned[slot], (loss_scaler.state_dict()["scale"],) + group_stats(), sq)
                    if pending is not None:
                        flush(pending)              # step i-1's loss / grad norm: the device is busy with step i
                    pending = mine
                    if lr_scheduler is not None:
                        lr_scheduler.step_update(start_steps + step)
                    continue
                loss_value = loss.item()                                                # :306 (the step's D2H read)
                if not math.isfinite(loss_value):
                    print("Loss is {}, stopping training".format(loss_value))
                    sys.exit(1)
            else:
                loss_value = loss.item()
                if not math.isfinite(loss_value):
                    print("Loss is {}, stopping training".format(loss_value))
                    sys.exit(1)
                grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=core.parameters(), arena=arena)
        else:
            vis_idx, msk_idx = core.indices_from_mask(bool_masked_pos)
            if videos.dtype == torch.uint8:
                videos = _lib.normalize_u8(videos.contiguous(), torch.empty(videos.shape, dtype=torch.float32, device=videos.device))
            with torch.no_grad():
                labels = build_labels(videos.float().contiguous(), msk_idx, normlize_target)
            outputs = model(videos, bool_masked_pos)
            loss = nn.functional.mse_loss(outputs.float(), labels)
            loss_value = loss.item()
            if not math.isfinite(loss_value):
                print("Loss is {}, stopping training".format(loss_value))
                sys.exit(1)
            optimizer.zero_grad()
            is_second_order = hasattr(optimizer, 'is_second_order') and optimizer.is_second_order
            grad_norm = loss_scaler(loss, optimizer, clip_grad=max_norm, parameters=model.parameters(),
                                    create_graph=is_second_order)
        loss_scale_value = loss_scaler.state_dict()["scale"]

        torch.cuda.synchronize()                                                        # :429
        core.check_mask_rows()

        record(loss_value, grad_norm, loss_scale_value, *group_stats())
        if lr_scheduler is not None:
            lr_scheduler.step_update(start_steps + step)

    if pending is not None:
        flush(pending)
    core.check_mask_rows()
    metric_logger.synchronize_between_processes()                                       # :466
    if not metric_logger.quiet:
        print("Averaged stats:", metric_logger)
    return {k: meter.global_avg for k, meter in metric_logger.meters.items()}


# run_mae_pretraining_BB.py:271 calls this undefined name; the intended binding is train_one_epoch_BB
train_one_epoch_BB_no_global_union_gradual = train_one_epoch_BB
