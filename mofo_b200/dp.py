"""Data-parallel gradient exchange for the fused pretraining step (replaces DDP's bucketed all-reduce,
run_mae_pretraining_BB.py:229-231).

One process per GPU.  The model's gradients live in ONE flat fp32 arena laid out in backward-completion order
(``_Runner.backward_order``): decoder first, encoder blocks last.  As backward finishes a stage, ``stage_done(k)``
records an event on the compute stream and enqueues ``all_reduce(sum)`` of that contiguous arena slice on a separate
communication stream (NCCL over NVLink/NVSwitch), so the exchange of the decoder's gradients overlaps the encoder's
backward and only the last slice is exposed.  The 1/world_size factor of DDP's mean is folded into the loss gradient
(``grad_scale``) so no extra pass over the arena is needed.

There is no data-path collective besides this all-reduce.

``begin(..., on_stage=f)``: ``f(k, lo, hi)`` runs on the communication stream right after stage k's slice ``arena[lo:hi]``
has been reduced - the fused optimizer uses it to update that slice's parameters (and add its share of the gradient
norm) while backward is still producing the next stage, so that only the LAST slice's all-reduce + update is exposed.
With one process the same hook runs on a side stream (no collective), overlapping the optimizer with backward.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.comm_stream = None
        self.hook_stream = None
        self.arena = None
        self.stage_end = None
        self.prev_end = 0
        self.launched = 0

    def sync_parameters(self, model, optimizer=None):
        """What the DistributedDataParallel constructor does in the reference (run_mae_pretraining_BB.py:229-231 after
        per-rank seeding :170-171): every rank adopts rank 0's parameters (and optimizer moments), once per model.
        Without it, replicas initialised from different seeds would silently train apart."""
        core = model.module if hasattr(model, "module") else model
        if self.world == 1 or getattr(core, "_mofo_params_synced", False):
            return
        with torch.no_grad():
            if optimizer is not None and getattr(optimizer, "fused_mofo", False) and getattr(optimizer, "_attached", None) is core:
                for t in (optimizer.p_arena, optimizer.m_arena, optimizer.v_arena):
                    dist.broadcast(t, src=0, group=self.pg)
            else:
                for p in core.parameters():
                    dist.broadcast(p.data, src=0, group=self.pg)
        runner = getattr(core, "_runner", None)
        if runner is not None:
            runner.wversion = None              # bf16 operand copies are re-cast from the adopted fp32 values
        core._mofo_params_synced = True

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def begin(self, arena, stage_end, on_stage=None):
        self.arena, self.stage_end, self.prev_end, self.launched = arena, stage_end, 0, 0
        self.on_stage = on_stage
        self.used_stream = False
        if (self.world > 1 or on_stage is not None) and arena.is_cuda and self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=arena.device)
        if on_stage is not None and arena.is_cuda and getattr(self, "hook_stream", None) is None:
            # the hook (optimizer update of a reduced slice) runs on its OWN stream behind an event, so that the next
            # slice's all-reduce does not queue behind it on the communication stream
            self.hook_stream = torch.cuda.Stream(device=arena.device)

    def stage_done(self, k):
        """Called by the backward pass right after the last gradient of arena stage ``k`` has been enqueued."""
        on_stage = getattr(self, "on_stage", None)
        if self.world == 1 and on_stage is None:
            return
        lo, end = self.prev_end, self.stage_end[k]
        sl = self.arena[lo:end]
        self.prev_end = end
        if sl.numel() == 0:
            return
        if sl.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream(sl.device))
            with torch.cuda.stream(self.comm_stream):
                if self.world > 1:
                    dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.pg)
            if on_stage is not None:
                self.hook_stream.wait_stream(self.comm_stream)
                with torch.cuda.stream(self.hook_stream):
                    on_stage(k, lo, end)
            self.used_stream = True
        else:                                   # gloo / CPU tensors (host-logic tests)
            if self.world > 1:
                dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.pg)
            if on_stage is not None:
                on_stage(k, lo, end)
        self.launched += 1

    def finish(self):
        """Make the compute stream wait for every outstanding slice (no host sync)."""
        if self.arena is None or (self.world == 1 and getattr(self, "on_stage", None) is None):
            return
        if self.prev_end < self.arena.numel():          # stages that never reported (defensive)
            self.stage_done(len(self.stage_end) - 1)
        if self.arena.is_cuda and self.comm_stream is not None:
            torch.cuda.current_stream(self.arena.device).wait_stream(self.comm_stream)
            if getattr(self, "on_stage", None) is not None and getattr(self, "hook_stream", None) is not None:
                torch.cuda.current_stream(self.arena.device).wait_stream(self.hook_stream)
