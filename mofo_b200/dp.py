"""Data-parallel gradient exchange for the fused pretraining step (replaces DDP's bucketed all-reduce,
run_mae_pretraining_BB.py:229-231).

One process per GPU.  The model's gradients live in ONE flat fp32 arena laid out in backward-completion order
(``_Runner.backward_order``): decoder first, encoder blocks last.  As backward finishes a stage, ``stage_done(k)``
records an event on the compute stream and enqueues ``all_reduce(sum)`` of that contiguous arena slice on a separate
communication stream (NCCL over NVLink/NVSwitch), so the exchange of the decoder's gradients overlaps the encoder's
backward and only the last slice is exposed.  The 1/world_size factor of DDP's mean is folded into the loss gradient
(``grad_scale``) so no extra pass over the arena is needed.

There is no data-path collective besides this all-reduce.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.comm_stream = None
        self.arena = None
        self.stage_end = None
        self.prev_end = 0
        self.launched = 0

    @property
    def grad_scale(self):
        return 1.0 / self.world

    def begin(self, arena, stage_end):
        self.arena, self.stage_end, self.prev_end, self.launched = arena, stage_end, 0, 0
        if self.world > 1 and arena.is_cuda and self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=arena.device)

    def stage_done(self, k):
        """Called by the backward pass right after the last gradient of arena stage ``k`` has been enqueued."""
        if self.world == 1:
            return
        end = self.stage_end[k]
        sl = self.arena[self.prev_end:end]
        self.prev_end = end
        if sl.numel() == 0:
            return
        if sl.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream(sl.device))
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.pg)
        else:                                   # gloo / CPU tensors (host-logic tests)
            dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.pg)
        self.launched += 1

    def finish(self):
        """Make the compute stream wait for every outstanding slice (no host sync)."""
        if self.world == 1 or self.arena is None:
            return
        if self.prev_end < self.arena.numel():          # stages that never reported (defensive)
            self.stage_done(len(self.stage_end) - 1)
        if self.arena.is_cuda:
            torch.cuda.current_stream(self.arena.device).wait_stream(self.comm_stream)
