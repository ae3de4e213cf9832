#!/usr/bin/env python
"""End-to-end MOFO pretraining on synthetic clips through the reference-shaped API, mirroring the flow of
run_mae_pretraining_BB.py::main (create_model -> create_optimizer -> cosine schedules -> train_one_epoch_BB per epoch ->
save_model-style checkpoint with the reference's state_dict keys).  Single process or torchrun (one rank per GPU).

    python examples/pretrain_synthetic.py --model pretrain_videomae_base_patch16_224 --batch_size 32 --epochs 2 --steps_per_epoch 20
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mofo_b200 import engine_for_pretraining as engine
from mofo_b200 import masking_generator as mg
from mofo_b200 import modeling_pretrain as mp
from mofo_b200 import optim_factory, utils


class SyntheticClips:
    """Yields (videos f32 [B,3,16,224,224] pinned, bbox long [B,16,4], mask f64 [B,1568]) like the reference loader
    (kinetics.py:1062-1064); masks come from the GPU generator with per-clip MT19937 streams."""

    quiet = False

    def __init__(self, batch, steps, size, rank, device, mask_ratio=0.9, mask_ratio_bb=0.75):
        self.batch, self.steps, self.size, self.rank = batch, steps, size, rank
        hw = size // 16
        self.gen = mg.TubeMaskingGenerator_BB((8, hw, hw), mask_ratio, mask_ratio_bb, device=device)
        self.mean = torch.tensor((0.485, 0.456, 0.406))[None, :, None, None, None]
        self.std = torch.tensor((0.229, 0.224, 0.225))[None, :, None, None, None]

    def __len__(self):
        return self.steps

    def __iter__(self):
        g = torch.Generator().manual_seed(1234 + self.rank)
        rng = np.random.default_rng(4321 + self.rank)
        for i in range(self.steps):
            vid = ((torch.rand(self.batch, 3, 16, self.size, self.size, generator=g) - self.mean) / self.std).pin_memory()
            w = rng.integers(self.size // 7, self.size * 5 // 7 + 1, self.batch); h = rng.integers(self.size // 7, self.size * 5 // 7 + 1, self.batch)
            x1 = (rng.random(self.batch) * (self.size - w + 1)).astype(np.int64); y1 = (rng.random(self.batch) * (self.size - h + 1)).astype(np.int64)
            bb = np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float64)
            words = np.stack([mg.mt19937_words((self.rank * 100003 + i) * self.batch + b, mg.words_per_clip(self.gen.height, self.gen.width))
                              for b in range(self.batch)])
            mask = self.gen.generate_batch(bb, words, check=True)[0]
            yield vid, torch.from_numpy(bb).long()[:, None, :].expand(self.batch, 16, 4).contiguous(), mask.double().cpu()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="pretrain_videomae_base_patch16_224")
    ap.add_argument("--decoder_depth", type=int, default=4)
    ap.add_argument("--batch_size", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--warmup_epochs", type=int, default=1)
    ap.add_argument("--steps_per_epoch", type=int, default=10)
    ap.add_argument("--lr", type=float, default=1.5e-4)
    ap.add_argument("--min_lr", type=float, default=1e-5)
    ap.add_argument("--weight_decay", type=float, default=0.05)
    ap.add_argument("--opt", default="adamw")
    ap.add_argument("--opt_betas", type=float, nargs="+", default=[0.9, 0.95])
    ap.add_argument("--opt_eps", type=float, default=1e-8)
    ap.add_argument("--clip_grad", type=float, default=0.0)
    ap.add_argument("--output_dir", default="")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)
    model = mp.create_model(args.model, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=args.decoder_depth).to(device)
    args.lr = args.lr * args.batch_size * world / 256                       # run_mae_pretraining_BB.py:219-223
    optimizer = optim_factory.create_optimizer(args, model)
    loss_scaler = utils.NativeScalerWithGradNormCount()
    n = args.steps_per_epoch
    lr_values = utils.cosine_scheduler(args.lr, args.min_lr, args.epochs, n, warmup_epochs=args.warmup_epochs)
    wd_values = utils.cosine_scheduler(args.weight_decay, args.weight_decay, args.epochs, n)
    size = model.encoder.patch_embed.img_size[0]
    for epoch in range(args.epochs):
        loader = SyntheticClips(args.batch_size, n, size, rank, device)
        loader.quiet = rank != 0
        stats = engine.train_one_epoch_BB(model, loader, optimizer, device, epoch, loss_scaler, max_norm=args.clip_grad,
                                          patch_size=16, normlize_target=True, start_steps=epoch * n,
                                          lr_schedule_values=lr_values, wd_schedule_values=wd_values)
        if rank == 0:
            print({k: round(v, 6) for k, v in stats.items()})
    if args.output_dir and rank == 0:
        os.makedirs(args.output_dir, exist_ok=True)                          # utils.save_model layout (utils.py:411-428)
        torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict(), "epoch": args.epochs - 1,
                    "scaler": loss_scaler.state_dict(), "args": vars(args)}, os.path.join(args.output_dir, f"checkpoint-{args.epochs - 1}.pth"))
    if world > 1:
        dist.destroy_process_group()
    return stats


if __name__ == "__main__":
    main()
