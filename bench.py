#!/usr/bin/env python
"""Benchmark of the MOFO pretraining step (BASELINE.json metric: ViT-B pretrain clips/s, 16x224^2, mask 0.9).

    python bench.py --gpus N --steps K --warmup W            # this repo (N>1: launched with torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the UNMODIFIED reference (baseline/_ref) on host cores

A step = one pass of the hot path over one batch of 32 synthetic clips per GPU: tube masking (GPU kernel, per-clip
MT19937 words) -> fused forward + target/MSE + backward -> gradient all-reduce (N>1) -> grad-norm -> AdamW.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MOFO ViT-B pretrain clips/s (16x224^2, tubelet 2x16x16, mask 0.9 / BB 0.75)"
TRAIN_GFLOP_PER_CLIP = {"pretrain_videomae_base_patch16_224": 202.3, "pretrain_mae_small_patch16_224": 64.05,
                        "pretrain_videomae_large_patch16_224": 484.4}     # BASELINE.md §3 (algorithmic, fwd+bwd)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU (BASELINE configs[1]: 32)")
    ap.add_argument("--model", default="pretrain_videomae_base_patch16_224")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "finetune", "motion"],
                    help="finetune = BASELINE configs[4]: vit_base_patch16_224 classifier fwd+bwd on all 1568 tokens, batch 8 (not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline-leg", action="store_true", help="profiling runs: skip the per-launch-event leg (roofline = null)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="wall budget of the oracle-port fallback of the reference arm")
    ap.add_argument("--ref-clips", type=int, default=8, help="clips per step of the CPU reference arm (fixed, independent of N)")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference's own eager step on the same GPU(s)")
    ap.add_argument("--gpu-ref-steps", type=int, default=30)
    ap.add_argument("--gpu-ref-warmup", type=int, default=8)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        window = "timed region"
        if len(inside) < 3:            # very short timed region: widen to the surrounding (equally loaded) warm-up / legs
            inside = [r for t, r in self.rows if self.t0 is not None and self.t0 - 1.0 <= t <= self.t1 + 1.0]
            window = "timed region +-1 s (region shorter than the sampling period)"
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path = the oracle port (pure-Python reference cannot travel)
# ------------------------------------------------------------------------------------------------------------------
def oracle_cpu_step_fn(model_name, threads):
    import numpy as np
    import torch
    from oracle import mask_oracle as mo
    from oracle import model_oracle as mdl
    from oracle import target_oracle as tgt
    torch.set_num_threads(threads)
    cfg = mdl.CONFIGS[model_name]
    sd = mdl.random_state_dict(cfg, seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1.5e-4, betas=(0.9, 0.95), weight_decay=0.05)
    state = {"i": 0}

    def step(n_clips):
        i = state["i"]; state["i"] += 1
        vid = tgt.synthetic_clip(n_clips, seed=1234 + i)
        boxes = tgt.synthetic_boxes(n_clips, seed=4321 + i)
        t0 = time.perf_counter()
        masks = np.stack([mo.tube_mask_bb(boxes[b], np.random.RandomState(i * 64 + b)._bit_generator.random_raw(800), cfg.grid)[0]
                          for b in range(n_clips)])
        mask = torch.from_numpy(masks).bool()
        with torch.no_grad():
            labels = tgt.build_labels(vid, mask)
        opt.zero_grad()
        loss = tgt.mse_loss(mdl.forward(cfg, params, vid, mask), labels)
        loss.backward()
        opt.step()
        return time.perf_counter() - t0, float(loss)
    return step


def reference_cpu(args, clips, steps, warmup):
    """The reference's own train_one_epoch_BB (baseline/_ref, unmodified: its model, create_optimizer, GradScaler-based
    scaler, target construction, autograd backward, AdamW) on the host cores, fp32 (CUDA autocast does not touch CPU ops)."""
    import torch
    from baseline import refrun
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    r = refrun.run("cpu", args.model, batch=clips, steps=steps, warmup=warmup, amp="fp16", pool=2)
    r["cores"] = cores
    return r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    from baseline import refrun
    if refrun.available():
        n = min(args.ref_clips, args.batch)
        r = reference_cpu(args, n, args.steps, args.warmup)
        val = r["clips_per_s"]
        sample = (f"{n} clips/step (fixed, independent of --gpus) x {args.steps} steps of the {args.batch}-clip batch; the unmodified "
                  "reference from baseline/_ref through timm.create_model -> optim_factory.create_optimizer -> "
                  "engine_for_pretraining.train_one_epoch_BB (fp32 on CPU, AdamW step included)")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args), "global_batch": args.batch * args.gpus, "parallelism": f"dp{args.gpus}"},
                "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "reference", "sample": sample},
                "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "final_loss": r["loss"]}
        print(json.dumps(line), flush=True)
        return
    step = oracle_cpu_step_fn(args.model, cores)
    t1, _ = step(1)                                   # probe (also first-touch warm-up)
    total_steps = args.steps + args.warmup
    n = int(max(1, min(args.batch, args.ref_budget_s / max(total_steps * t1, 1e-9))))
    for _ in range(args.warmup):
        step(n)
    ts = [step(n)[0] for _ in range(args.steps)]
    tot = sum(ts)
    val = n * args.steps / tot
    sample = f"{n} clip(s)/step x {args.steps} steps of the {args.batch}-clip batch; oracle port (torch fp32 CPU autograd + numpy mask), AdamW step included"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": args.batch * args.gpus, "parallelism": f"dp{args.gpus}"},
            "cpu_baseline": {"value": val, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    return (f"BASELINE configs[1]: ViT-B MOFO pretrain ({args.model}, decoder depth 4), batch {args.batch}/GPU, synthetic "
            f"16x224x224 clips, tubelet 2x16x16, mask 0.9 / BB 0.75, full step incl. AdamW")


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def param_groups(model, weight_decay=0.05):
    """optim_factory.get_parameter_groups (optim_factory.py:49-88): no decay for 1-D params, biases and the skip list."""
    skip = model.no_weight_decay()
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        (no_decay if (p.ndim == 1 or name.endswith(".bias") or name in skip) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0, "lr_scale": 1.0}, {"params": decay, "weight_decay": weight_decay, "lr_scale": 1.0}]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mofo_b200 import _lib
    from mofo_b200 import engine_for_pretraining as eng
    from mofo_b200 import masking_generator as mg
    from mofo_b200 import modeling_pretrain as mp
    from mofo_b200 import utils as U
    from mofo_b200.dp import GradSync
    _lib.load()

    B = args.batch
    torch.manual_seed(0)                                  # identical init on every rank (DDP broadcasts rank 0's)
    model = mp.create_model(args.model, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, decoder_depth=4).to(dev)
    model.train()
    lr = 1.5e-4 * B * world / 256                         # run_mae_pretraining_BB.py:219-223
    from mofo_b200.optim_factory import FusedAdamW
    opt = FusedAdamW(param_groups(model), lr=lr, betas=(0.9, 0.95), weight_decay=0.05).attach(model)
    scaler = U.NativeScalerWithGradNormCount()
    gen = mg.TubeMaskingGenerator_BB((8, 14, 14), 0.9, 0.75, device=dev)
    mean = torch.tensor((0.485, 0.456, 0.406), device=dev)[None, :, None, None, None]
    std = torch.tensor((0.229, 0.224, 0.225), device=dev)[None, :, None, None, None]

    # synthetic inputs: POOL rotating batches so every step reads inputs that are not L2-resident
    POOL = 4
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    rng = np.random.default_rng(4321 + rank)
    pool = []
    for i in range(POOL):
        vid = (torch.rand(B, 3, 16, 224, 224, generator=g, device=dev) - mean) / std
        w = rng.integers(32, 161, B); h = rng.integers(32, 161, B)
        x1 = (rng.random(B) * (224 - w + 1)).astype(np.int64); y1 = (rng.random(B) * (224 - h + 1)).astype(np.int64)
        bb = torch.from_numpy(np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float64)).to(dev)
        words = np.stack([mg.mt19937_words((rank * POOL + i) * B + b, 768) for b in range(B)])
        words = torch.from_numpy(words.view(np.int32)).to(dev)
        pool.append((vid.contiguous(), bb, words))
    sync = GradSync()
    sync.sync_parameters(model, opt)
    runner = model._runner
    runner._ensure_device(dev)
    arena = runner.grad_arena()

    # the mask of step i+1 is sampled on a side stream while step i runs (the reference's DataLoader workers likewise
    # prepare masks ahead of the step that consumes them); every step's mask kernel is inside the timed region
    mask_stream = torch.cuda.Stream(device=dev)
    ahead = {}

    def sample_masks(i):
        _, bb, words = pool[i % POOL]                      # static inputs: nothing on the main stream to wait for
        with torch.cuda.stream(mask_stream):
            out = gen.generate_batch(bb, words)
            ev = torch.cuda.Event()
            ev.record(mask_stream)
        return out, ev

    def device_step(i):
        vid = pool[i % POOL][0]
        (mask, vis_idx, msk_idx, used), ev = ahead.pop(i) if i in ahead else sample_masks(i)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        for t in (vis_idx, msk_idx):
            t.record_stream(cur)
        ahead.clear()
        ahead[i + 1] = sample_masks(i + 1)                 # enqueued before step i, runs beside it
        loss, _ = eng.fused_step(model, opt, scaler, sync, vid, vis_idx, msk_idx, True, 0, return_sq=True)
        return loss

    def dp_check():
        """N>1 only, during warm-up: the staged, overlapped all-reduce must leave on every rank the same gradients as ONE
        plain all-reduce of the rank-local gradients (no dropped or doubly reduced slice), and all ranks must agree bit
        for bit.  Uses the eager (non-staged-optimizer) path so the parameters are not touched."""
        vid, bb, words = pool[0]
        mask, vis_idx, msk_idx, used = gen.generate_batch(bb, words)
        sync.begin(arena, runner.stage_end)
        model.pretrain_step(vid, vis_idx=vis_idx, msk_idx=msk_idx, normalize_target=True, grad_scale=sync.grad_scale,
                            zero_grad=True, stage_done=sync.stage_done)
        sync.finish()
        staged = arena.clone()
        model.pretrain_step(vid, vis_idx=vis_idx, msk_idx=msk_idx, normalize_target=True, grad_scale=sync.grad_scale,
                            zero_grad=True, stage_done=None)
        local = arena.clone()
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        rel = ((staged.double() - local.double()).norm() / local.double().norm()).item()
        chk = torch.stack([staged.double().sum(), staged.double().pow(2).sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        assert rel <= 1e-5 and same, f"data-parallel gradient check failed: rel {rel:.3e}, identical across ranks: {same}"
        return {"staged_vs_single_allreduce_rel_l2": rel, "identical_across_ranks": same, "tolerance": 1e-5}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                   # started before warm-up: nvidia-smi needs ~0.3 s to produce its first row
    dp = dp_check() if world > 1 else None
    for i in range(args.warmup):
        loss = device_step(i)
    barrier()
    assert torch.isfinite(loss).item(), "non-finite loss in warm-up"

    launches0 = _lib.launch_count
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.t0 = time.time()
    e0.record()
    for i in range(args.steps):
        loss = device_step(args.warmup + i)
    e1.record()
    barrier()
    sampler.t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = B * world * args.steps / (ms / 1e3)
    final_loss = loss.item()
    # roofline leg: the same K steps again, now with a CUDA-event pair around every launch of the dominant kernel
    # class (kept out of the region above because ~270 extra event records per step perturb the step time)
    e2 = torch.cuda.Event(enable_timing=True); e3 = torch.cuda.Event(enable_timing=True)
    barrier()
    n_gemm, gemm_ms, gemm_flops, ms_instrumented = 0, 0.0, 0.0, 1.0
    if not args.no_roofline_leg:
        model.use_cuda_graph = False          # per-launch events need individually launched kernels
        device_step(0); device_step(1)
        barrier()
        with _lib.timing("mofo_gemm_tn") as ktimer:
            e2.record()
            for i in range(args.steps):
                device_step(args.warmup + args.steps + i)
            e3.record()
            barrier()
        n_gemm, gemm_ms, gemm_flops = ktimer.summary()
        ms_instrumented = e2.elapsed_time(e3)
        model.use_cuda_graph = True
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the public engine API with HOST (pinned) batches; H2D + loss D2H inside the timed region ----------
    e2e = None
    if not args.no_e2e:
        class HostLoader:
            quiet = True

            def __init__(self, n):
                self.n = n
                self.batches = []
                for i in range(2):
                    vid, bb, words = pool[i]
                    mask = gen.generate_batch(bb, words)[0]
                    self.batches.append((vid.cpu().pin_memory(), bb.long().cpu()[:, None, :].expand(B, 16, 4).contiguous(),
                                         mask.double().cpu().pin_memory()))

            def __len__(self):
                return self.n

            def __iter__(self):
                for i in range(self.n):
                    yield self.batches[i % 2]
        # pinned-host -> device copy bandwidth of this box (the e2e path moves one fp32 batch per step)
        hb = pool[0][0].cpu().pin_memory(); db = torch.empty_like(pool[0][0])
        db.copy_(hb, non_blocking=True); torch.cuda.synchronize()
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        c0.record(); db.copy_(hb, non_blocking=True); c1.record(); torch.cuda.synchronize()
        h2d_gbps = hb.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del hb, db
        def run_epoch(loader):
            eng.train_one_epoch_BB(model, loader, opt, dev, 0, scaler, max_norm=0, patch_size=16, normlize_target=True, start_steps=0)

        def timed_epoch(make_loader):
            run_epoch(make_loader(max(3, args.warmup)))
            loader = make_loader(args.steps)
            barrier()
            t0 = time.perf_counter()
            run_epoch(loader)
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return dt.item()

        # extension: the host ships raw uint8 clips, ToTorchFormatTensor+GroupNormalize run on the GPU
        class HostLoaderU8(HostLoader):
            def __init__(self, n):
                self.n = n
                gcpu = torch.Generator().manual_seed(77 + rank)
                self.batches = []
                for i in range(2):
                    _, bb, words = pool[i]
                    mask = gen.generate_batch(bb, words)[0]
                    u8 = torch.randint(0, 256, (B, 3, 16, 224, 224), dtype=torch.uint8, generator=gcpu).pin_memory()
                    self.batches.append((u8, bb.long().cpu()[:, None, :].expand(B, 16, 4).contiguous(), mask.double().cpu().pin_memory()))
        dt_u8 = timed_epoch(HostLoaderU8)
        loader = HostLoader(max(3, args.warmup))
        run_epoch(loader)
        loader = HostLoader(args.steps)
        barrier()
        t0 = time.perf_counter()
        eng.train_one_epoch_BB(model, loader, opt, dev, 0, scaler, max_norm=0, patch_size=16, normlize_target=True, start_steps=0)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        vid_bytes = B * 3 * 16 * 224 * 224 * 4
        # headline e2e = the engine fed what a decoder hands over: uint8 clips (1 B / sample; ToTorchFormatTensor +
        # GroupNormalize run on the GPU, mofo_normalize_u8); the same call fed pre-normalised fp32 clips - what the
        # reference's CPU DataLoader workers produce - is reported beside it
        fp32_rate = B * world * args.steps / dt.item()
        e2e = {"value": B * world * args.steps / dt_u8, "unit": "clips/s",
               "h2d_bytes_per_step": vid_bytes // 4 + B * 1568 * 8, "d2h_bytes_per_step": 12,
               "ms_per_step": 1e3 * dt_u8 / args.steps, "h2d_pinned_gbps": h2d_gbps,
               "input": "pinned host uint8 clips [B,3,16,224,224] + f64 masks [B,1568] per step",
               "api": "mofo_b200.engine_for_pretraining.train_one_epoch_BB (H2D of batch i+1 overlaps step i on a copy stream; loss, "
                      "grad norm and the mask-row check of every step are copied to pinned host memory and read by the host while "
                      "the next step runs)",
               "fp32_input": {"value": fp32_rate, "unit": "clips/s", "h2d_bytes_per_step": vid_bytes + B * 1568 * 8,
                              "ms_per_step": 1e3 * dt.item() / args.steps,
                              "h2d_ms_per_step_at_the_measured_rate": vid_bytes / h2d_gbps / 1e6,
                              "note": "same engine call fed pre-normalised fp32 clips (4 B / sample), as the reference's DataLoader yields them"}}

    # ---- the reference's own eager step on the same GPU(s) (SURVEY §8d "GPU reference baseline"): unmodified modules from
    # baseline/_ref through train_one_epoch_BB, fp16 autocast + GradScaler as authored and under bf16 autocast; DDP
    # (find_unused_parameters=True, run_mae_pretraining_BB.py:229-231) at N>1.  Reported beside our numbers, never mixed in.
    gpu_ref = None
    if not args.no_gpu_reference:
        from baseline import refrun
        if refrun.available():
            torch.cuda.empty_cache()
            gpu_ref = {}
            for tag, amp, host in (("fp16_as_authored", "fp16", False), ("bf16_autocast", "bf16", False), ("bf16_autocast_host_inputs", "bf16", True)):
                try:
                    r = refrun.run(dev, args.model, batch=B, steps=args.gpu_ref_steps, warmup=args.gpu_ref_warmup, amp=amp,
                                   host_inputs=host, ddp=world > 1)
                    tm = torch.tensor([r["ms_per_step"]], device=dev)
                    if world > 1:
                        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                    r["ms_per_step"] = tm.item()
                    r["value"] = B * world / (tm.item() * 1e-3)
                    r["unit"] = "clips/s"
                    del r["clips_per_s"]
                    gpu_ref[tag] = r
                except Exception as e:                      # noqa: BLE001 - reported, never fatal for our own line
                    gpu_ref[tag] = {"error": repr(e)[:300]}
            gpu_ref["how"] = ("unmodified reference modules (baseline/_ref) via timm.create_model + optim_factory.create_optimizer + "
                              "utils.NativeScalerWithGradNormCount + engine_for_pretraining.train_one_epoch_BB, eager PyTorch "
                              f"{torch.__version__} (cuBLAS/cuDNN/ATen kernels), same B, synthetic clips, CUDA events around the epoch call"
                              + (", DistributedDataParallel(find_unused_parameters=True)" if world > 1 else ""))
        else:
            gpu_ref = {"unavailable": "baseline/_ref not staged (python baseline/setup_ref.py needs /root/reference)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    gflop = TRAIN_GFLOP_PER_CLIP.get(args.model)
    step_tflops = (value / world) * gflop / 1e3 if gflop else None
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    line = {"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": B * world, "parallelism": f"dp{world}"},
            "notes": {"l2": f"inputs larger than L2: {POOL} rotating {B * 3 * 16 * 224 * 224 * 4 >> 20} MiB batches per GPU",
                      "optimizer": "mofo_b200.optim_factory.FusedAdamW: one mofo_adamw_step kernel over the flat fp32 arenas, also "
                                   "emitting the bf16 W / W^T operand copies (SURVEY §8f-1)",
                      "launch": "fused step replayed as CUDA graphs (one per gradient-sync stage); roofline leg launches kernels individually"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "final_loss": final_loss, "gpu_reference": gpu_ref, "dp_check": dp,
            "roofline": {"bound": "tensor", "kernel": "gemm_tn2_kernel / gemm_tn_kernel (tcgen05 cta_group::2 and ::1, every fused-epilogue instance behind mofo_gemm_tn)",
                         "achieved": gemm_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": gemm_tflops / peaks["bf16_sustained"] if gemm_tflops else None,
                         "frac_of_burst_peak": gemm_tflops / peaks["bf16_burst"] if gemm_tflops else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the shipped kernel, `ncu --set full`
                         # (profiles/r02_ncu_gemm_epilogues_v3.txt): decoder fc1 + GELU instance gemm_tn2_kernel<256,1>
                         # [50176 x 1536, K = 384]: 39.8 MB read + 254.1 MB written vs 347.9 MB algorithmic (A 38.5 + W 1.2 +
                         # two bf16 outputs 2 x 154.1; the tail of the outputs was still dirty in L2) -> no re-reads
                         "traffic": 293.9e6, "traffic_instance": "gemm_tn2_kernel<256,BIAS_GELU_BF16> M=50176 N=1536 K=384 (algorithmic 347.9e6 B)",
                         "launches_timed": n_gemm, "kernel_ms_per_step": gemm_ms / args.steps if n_gemm else None,
                         "share_of_step": gemm_ms / ms_instrumented if n_gemm else None,
                         "instrumented_ms_per_step": ms_instrumented / args.steps if n_gemm else None,
                         "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step); frac_of_burst_peak uses bf16_tflops"},
            "step_mfu": {"algorithmic_tflops_per_gpu": step_tflops, "frac_of_measured_sustained": step_tflops / peaks["bf16_sustained"] if step_tflops else None,
                         "frac_of_nominal_2250": step_tflops / 2250.0 if step_tflops else None, "gflop_per_clip": gflop}}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        from baseline import refrun
        if refrun.available():
            n = min(args.ref_clips, B)
            r = reference_cpu(args, n, 3, 1)
            line["cpu_baseline"] = {"value": r["clips_per_s"], "unit": "clips/s", "cores": cores, "kind": "reference",
                                    "sample": f"3 steps (after 1 warm-up) of {n} clips of the same workload: the unmodified reference "
                                              "(baseline/_ref) train_one_epoch_BB on the host cores, fp32, AdamW step included"}
        else:
            step = oracle_cpu_step_fn(args.model, cores)
            t1, _ = step(1)
            n = 2 if t1 < 6 else 1
            ts = [step(n)[0] for _ in range(max(1, min(3, int(20 / max(t1 * n, 1e-9)))))]
            line["cpu_baseline"] = {"value": n * len(ts) / sum(ts), "unit": "clips/s", "cores": cores, "kind": "port",
                                    "sample": f"{len(ts)} step(s) of {n} clip(s) of the same workload (oracle port: numpy mask + torch fp32 CPU fwd/target/MSE/bwd + AdamW)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_finetune(args):
    """BASELINE configs[4] (dense-attention stress): classifier forward + cross-entropy + backward on all 1568
    tokens, B = 8 per GPU (FINETUNE.md:25), through the public nn.Module API; CUDA events, inputs resident in HBM (two
    rotating batches + a 256 MiB L2 flush write between steps).  The reference's own module on the same GPU is timed beside it."""
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from mofo_b200 import _lib
    from mofo_b200 import modeling_finetune as mf
    _lib.load()
    B = 8 if args.batch == 32 else args.batch
    name = "vit_base_patch16_224" if args.model.startswith("pretrain") else args.model
    kw = dict(num_classes=174, all_frames=16, tubelet_size=2, drop_rate=0.0, drop_path_rate=0.0, attn_drop_rate=0.0,
              use_mean_pooling=True, init_scale=0.001)
    torch.manual_seed(0)
    model = mf.create_model(name, pretrained=False, drop_block_rate=None, **kw).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(1234)
    xs = [torch.randn(B, 3, 16, 224, 224, generator=g, device=dev) for _ in range(2)]
    ys = [torch.randint(0, 174, (B,), device=dev, generator=g) for _ in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(m, i, amp=None):
        flush.zero_()
        for p in m.parameters():
            p.grad = None
        if amp is None:
            loss = torch.nn.functional.cross_entropy(m(xs[i % 2]), ys[i % 2])
        else:
            with torch.autocast("cuda", dtype=amp):
                loss = torch.nn.functional.cross_entropy(m(xs[i % 2]).float(), ys[i % 2])
        loss.backward()
        return loss

    def timed(m, amp=None):
        for i in range(args.warmup):
            step(m, i, amp)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count
        e0.record()
        for i in range(args.steps):
            loss = step(m, i, amp)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, loss.item(), _lib.launch_count - n0
    ms, loss, launches = timed(model)
    gflop = 1078.0                                   # ViT-B finetune fwd+bwd per clip
    line = {"metric": "MOFO ViT-B finetune clips/s (16x224^2, all 1568 tokens, fwd+bwd, no optimizer)", "value": B / (ms * 1e-3), "unit": "clips/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: {name} classifier fwd + CE + bwd on all 1568 tokens, batch {B}, dense-attention stress", "global_batch": B, "parallelism": "dp1"},
            "gpu_launches": launches, "final_loss": loss,
            "step_mfu": {"algorithmic_tflops_per_gpu": B / (ms * 1e-3) * gflop / 1e3, "gflop_per_clip": gflop,
                         "frac_of_nominal_2250": B / (ms * 1e-3) * gflop / 1e3 / 2250.0},
            "notes": {"l2": "256 MiB flush write between steps", "includes": "L2 flush fill (0.04 ms) inside the timed region for both arms"}}
    del model
    torch.cuda.empty_cache()
    from baseline import refrun
    if refrun.available() and not args.no_gpu_reference:
        ref = refrun.load()
        torch.manual_seed(0)
        rm = getattr(ref.modeling_finetune, name)(pretrained=False, **kw).to(dev).train()
        line["gpu_reference"] = {}
        for tag, amp in (("fp16_autocast_as_authored", torch.float16), ("bf16_autocast", torch.bfloat16)):
            rms, rloss, _ = timed(rm, amp)
            line["gpu_reference"][tag] = {"ms_per_step": rms, "value": B / (rms * 1e-3), "unit": "clips/s", "loss": rloss}
        line["gpu_reference"]["how"] = "unmodified reference modeling_finetune." + name + " (baseline/_ref), eager PyTorch, same inputs, same timing loop (no GradScaler: forward + CE + backward only)"
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# motion-box preprocessing, pixel stages (flow video -> motion map -> filtered gray map)
# ------------------------------------------------------------------------------------------------------------------
MOTION_METRIC = "MOFO motion-box preprocessing frames/s (240x320 flow video -> motion map -> gaussian/threshold/gaussian -> gray)"
MOTION_T, MOTION_H, MOTION_W, MOTION_WS = 48, 240, 320, 8


def motion_workload():
    return (f"SURVEY 8f-4: one SSv2-shaped optical-flow video per step ({MOTION_T} frames of {MOTION_H}x{MOTION_W}x3 uint8, ws = {MOTION_WS}): "
            "motion_map_creator.py:160-228 then bounding_box_creator_SSV.py:125-166 on every frame")


def motion_cpu(frames, steps, warmup):
    """The reference's own CPU path (baseline/motion_ref.py: its motion_sts functions + scipy + cv2) on the first `frames`
    frames of a synthetic video, one host thread (scipy.ndimage and this cv2 path are single-threaded)."""
    from baseline import motion_ref
    flows = motion_ref.synthetic_flow_video(3, MOTION_T, MOTION_H, MOTION_W)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        mm = motion_ref.reference_motion_map(flows[:max(frames, MOTION_WS)], MOTION_WS)[:frames]
        for f in mm:
            motion_ref.reference_filter_frame(f.copy())
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return frames * len(ts) / sum(ts), 1e3 * sum(ts) / len(ts)


def run_motion_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from baseline import motion_ref
    if not motion_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/motion_sts.py, scipy or cv2 missing"}), flush=True)
        return
    frames = 12
    val, ms = motion_cpu(frames, args.steps, min(args.warmup, 1))
    sample = (f"{frames} of the video's {MOTION_T} frames per step x {args.steps} steps: the reference's motion_sts.py (baseline/_ref) + "
              "scipy.ndimage + cv2 replaying motion_map_creator.py:160-228 and bounding_box_creator_SSV.py:125-166")
    print(json.dumps({"impl": "reference", "metric": MOTION_METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "u8 / f64", "data": "synthetic", "config": {"workload": motion_workload()},
                      "cpu_baseline": {"value": val, "unit": "frames/s", "cores": 1, "kind": "reference", "sample": sample},
                      "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


def run_motion(args):
    """One step = one flow video through mofo_motion_map + mofo_motion_box_filter (9 kernel launches).  Inputs resident in HBM:
    16 rotating videos (177 MB > L2).  e2e: pinned host flows in, gray maps back to pinned host memory, inside the timed region.
    Independent videos: N GPUs would run N replicas (no exchange); measured on one."""
    if int(os.environ.get("RANK", "0")) != 0:
        return                      # replicas only: videos are independent, no exchange step; one replica is measured (rank 0)
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    from mofo_b200 import _lib
    from mofo_b200 import motion_boxes as mb
    _lib.load()
    peaks = measured_peaks()
    T, H, W, ws = MOTION_T, MOTION_H, MOTION_W, MOTION_WS
    g = torch.Generator(device="cpu").manual_seed(7)
    host = [torch.randint(96, 160, (T, H, W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(4)]
    for hbuf in host:                                        # a moving patch with a different flow value
        hbuf[:, 60:150, 80:200, :2] += 70
    vids = [host[i % 4].to(dev) + (i // 4) for i in range(16)]
    flt = mb.MotionMapFilter()
    gray_host = torch.empty(T, H, W, dtype=torch.uint8).pin_memory()

    def step(i):
        return flt.filter(mb.motion_map(vids[i % 16], ws=ws))

    sampler = ClockSampler(0)
    sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count
    sampler.t0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    sampler.t1 = time.time()
    launches = _lib.launch_count - n0
    ms = e0.elapsed_time(e1) / args.steps
    # per-stage device times (CUDA events on the launching stream, separate leg)
    ea = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ta = tb = 0.0
    for i in range(args.steps):
        ea[0].record(); mm = mb.motion_map(vids[i % 16], ws=ws); ea[1].record(); flt.filter(mm); ea[2].record()
        torch.cuda.synchronize()
        ta += ea[0].elapsed_time(ea[1]); tb += ea[1].elapsed_time(ea[2])
    ta /= args.steps; tb /= args.steps
    # e2e through the module API with host buffers
    for i in range(2):
        d = host[i % 4].to(dev, non_blocking=True)
        gray_host.copy_(flt.filter(mb.motion_map(d, ws=ws))[1], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        d = host[i % 4].to(dev, non_blocking=True)
        gray_host.copy_(flt.filter(mb.motion_map(d, ws=ws))[1], non_blocking=True)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    px = T * H * W
    bytes_a = px * 6                                          # stage A: 3 B read + 3 B written per pixel and frame
    slots_b = px * 3 * 2 * ((4 + 1) + (120 + 1)) * 3          # stage B: DMUL + DADD per tap pair, 3 passes each for sigma 1 and 30
    fp64_peak = 148 * 64 * (clocks["sm_mhz"] or 1900.0) * 1e6   # DP lanes x SMs x clock: one DMUL or DADD per lane and cycle
    line = {"metric": MOTION_METRIC, "value": T / (ms * 1e-3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 / f64", "data": "synthetic",
            "config": {"workload": motion_workload(), "l2": "16 rotating videos (177 MB) > L2",
                       "parallelism": "replicas only (independent videos, no collective); one replica measured"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": T / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": T * H * W * 3, "d2h_bytes_per_step": T * H * W,
                    "api": "mofo_b200.motion_boxes.motion_map + MotionMapFilter.filter on pinned host uint8 flows; gray maps copied back to pinned host memory"},
            "roofline": {"bound": "hbm", "kernel": "motion_map_kernel (stage A; stage B is bound by the float64 pipe, see stage_b)",
                         "achieved": bytes_a / (ta * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": bytes_a / (ta * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r02_ncu_motion_full.txt):
                         # the 11.06 MB video is read once; the map (11.06 MB) was still in L2 when the next kernel consumed it
                         "traffic": 11.08e6,
                         "algorithmic_bytes_per_launch": bytes_a, "kernel_ms": ta, "peak_source": peaks["source"]},
            "stage_b": {"kernels": "3 x gauss_pass (sigma 1) + box_stats + 3 x gauss_pass (sigma 30, 241 taps in scipy's float64 order) + gray",
                        "ms": tb, "share_of_step": tb / (ta + tb), "fp64_issue_slots": slots_b,
                        "achieved_fp64_slots_per_s": slots_b / (tb * 1e-3), "frac_of_fp64_issue_peak": slots_b / (tb * 1e-3) / fp64_peak,
                        "note": "bit-exactness with scipy forbids FMA contraction and reordering: one DMUL + one DADD per tap pair"}}
    if not args.no_cpu_baseline:
        from baseline import motion_ref
        if motion_ref.available():
            frames = 12
            val, cms = motion_cpu(frames, 2, 1)
            line["cpu_baseline"] = {"value": val, "unit": "frames/s", "cores": 1, "kind": "reference",
                                    "sample": f"2 steps (after 1 warm-up) of {frames} of the video's {T} frames: the reference's motion_sts.py + scipy.ndimage + cv2 "
                                              "(baseline/motion_ref.py), single thread as in the reference's per-video worker"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.workload == "motion":
        run_motion_reference(a) if a.impl == "reference" else run_motion(a)
    elif a.workload == "finetune" and a.impl != "reference":
        run_finetune(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
