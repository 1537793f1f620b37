"""Benchmark of the FSNet training-step hot path (BASELINE.json metric: images/sec on 640x192 triplets).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm: cfg2a (configs/kitti_wpose_synthetic.py: ResNet-18, 192x640, 4 scales, 16 bins, B=12 per
GPU), one step = BaseTrainingHook.__call__ (zero_grad, forward, backward, clip, Adam).  `value` is
measured with the batch resident in HBM; `e2e` includes the pinned-host -> device copy of every step's
batch and a device -> host read of the loss.  `roofline` is the fused warp-SSIM forward kernel
(algorithmic bytes of SURVEY.md 8(d) / CUDA-event time of its launches inside the timed steps).
`--impl reference` times the CPU restatement of the reference (oracle/, pinned to the reference by
tests/golden) on the host cores.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

B_PER_GPU, H, W = 12, 192, 640
SCALES = (0, 1, 2, 3)
CONFIG = os.path.join(REPO, "configs", "kitti_wpose_synthetic.py")
# --workload: the default is BASELINE.json's configs[1] (what the driver runs); the others are extra, informational lines
WORKLOADS = {
    "cfg2a": dict(config="kitti_wpose_synthetic.py", B=12, H=192, W=640, fisheye=False, gflop_per_image=17.02, enc_convs=20, enc_gflop_per_image=8.883,
                  text="cfg2a kitti_wpose_synthetic: ResNet-18 depth net, 192x640, 4 scales, 16 bins, dataset poses, fwd+bwd+clip(35)+Adam"),
    "cfg2b": dict(config="kitti_posenet_synthetic.py", B=12, H=192, W=640, fisheye=False, gflop_per_image=17.02 + 2 * 9.776,
                  text="cfg2b kitti_posenet_synthetic: MonoDepthMeta, ResNet-18 depth net + ResNet-18 PoseNet (2 pairs), 192x640, 4 scales, "
                       "fwd+bwd+clip(35)+Adam"),
    # BASELINE.json configs[2] and [3] at their stated sizes (per-GPU batch 8); forward-convolution FLOPs are counted from the launches
    "cfg3": dict(config="kitti360_r50_synthetic.py", B=8, H=192, W=768, fisheye=False, enc_convs=53,
                 text="cfg3 kitti360_r50_synthetic: ResNet-50 depth net, 192x768, 4 scales, 16 bins, dataset poses, fwd+bwd+clip(35)+Adam"),
    "cfg4": dict(config="nusc_wpose_synthetic.py", B=8, H=320, W=640, fisheye=False, enc_convs=20,
                 text="cfg4 nusc_wpose_synthetic: ResNet-18 depth net, 320x640, 4 scales, 16 bins, dataset poses, fwd+bwd+clip(35)+Adam"),
    "cfg5": dict(config="kitti360_fisheye_synthetic.py", B=4, H=512, W=512, fisheye=True, gflop_per_image=18.95 + 24.16,
                 text="cfg5 kitti360_fisheye_synthetic: FishEyeDecoder (MEI camera), ResNet-18, 512x512, 4 scales, 64 bins, dataset poses, "
                      "is_log_image (default True), fwd+bwd+clip(1)+Adam"),
}


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def oracle_step_rate(batch, steps, warmup, split=None):
    """The oracle's full training step on the host cores -> triplets/s.  ``split`` (a dict) receives the mean seconds per phase:
    encoder forward, decoder forward, loss chain forward, backward, clip + Adam."""
    from oracle import fsnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    topo = O.Topology(height=H, width=W, scales=SCALES)
    trainer = O.OracleTrainer(topo, seed=123, lr=1e-4, clip=35.0)
    times, phases = [], {"encoder_fwd": 0.0, "decoder_fwd": 0.0, "loss_fwd": 0.0, "backward": 0.0, "clip_adam": 0.0}
    for i in range(warmup + steps):
        data = O.synthetic_batch(batch, H, W, seed=1234 + i)
        noise = O.tie_break_noise(batch, H, W, SCALES, seed=i)
        t = [time.perf_counter()]
        # OracleTrainer.step, phase by phase (same calls, same order)
        trainer.opt.zero_grad()
        feats = O.resnet_forward(trainer.sd, "depth_backbone.", data[("image", 0)], topo.depth)
        t.append(time.perf_counter())
        outputs = O.decoder_forward(trainer.sd, "head.depth_decoder.", feats, topo, data["P2"])
        t.append(time.perf_counter())
        out = O.loss_chain(outputs, data, {f: data[("relative_pose", f)] for f in topo.frame_ids[1:]}, topo, noise, False)
        t.append(time.perf_counter())
        out["loss"].mean().backward()
        t.append(time.perf_counter())
        torch.nn.utils.clip_grad_norm_(trainer.params, trainer.clip)
        trainer.opt.step()
        t.append(time.perf_counter())
        if i >= warmup:
            times.append(t[-1] - t[0])
            for k, a, b in zip(phases, t[:-1], t[1:]):
                phases[k] += b - a
    if split is not None:
        split.update({k: v / len(times) for k, v in phases.items()})
    return batch * len(times) / sum(times), sum(times) / len(times)


def run_reference(args, rank):
    if rank != 0:
        return
    batch = 4 if (args.steps + args.warmup) <= 12 else 2
    split = {}
    rate, sec = oracle_step_rate(batch, args.steps, args.warmup, split)
    cores = os.cpu_count() or 1
    sample = (f"B={batch} of {B_PER_GPU} triplets per step (192x640, ResNet-18, 4 scales), full step fwd+bwd+clip+Adam; "
              "deviation from BASELINE.md: the CPU restatement of the reference (oracle/, pinned to the reference by 28 golden "
              "fixtures), not the reference package itself (/root/reference does not exist on the GPU box), at a bounded batch")
    line = {
        "impl": "reference", "metric": f"images/sec ({W}x{H} triplets), full training step", "value": rate, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2a kitti_wpose 192x640 R18 4-scale n=16 (CPU restatement of the reference, oracle/)", "batch": batch},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample,
                         "seconds_per_step_split": split},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class LossReader:
    """Device -> host read of every step's loss inside the timed region.  sync=1: ``loss.item()`` per step (the host blocks on
    the GPU and its launch work for the next step is exposed).  sync=0: every step's loss is copied to a pinned host slot on the
    step's stream and read once the NEXT step has been queued -- what a logging loop does; every value still reaches the host
    inside the timed region (``last()`` drains the final one)."""

    def __init__(self, sync):
        self.sync = bool(sync)
        self.slots = [torch.zeros(1, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self.value = None

    def push(self, loss):
        if self.sync:
            self.value = loss.item()
            return
        if self.i > 0:
            j = (self.i - 1) % 2
            self.events[j].synchronize()
            self.value = float(self.slots[j][0])
        j = self.i % 2
        self.slots[j].copy_(loss.detach().reshape(1).double(), non_blocking=True)
        self.events[j].record()
        self.i += 1

    def last(self):
        if not self.sync and self.i > 0:
            j = (self.i - 1) % 2
            self.events[j].synchronize()
            self.value = float(self.slots[j][0])
        return self.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="fsnet_b200", choices=["fsnet_b200", "reference"])
    ap.add_argument("--backend", default=os.environ.get("FSNET_CONV_BACKEND", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run every step eagerly (no CUDA graph replay)")
    ap.add_argument("--prefetch", type=int, default=int(os.environ.get("FSNET_BENCH_PREFETCH", 1)),
                    help="e2e leg: upload batch k+1 on a side stream while step k runs (fsnet_b200.data.loading.DevicePrefetcher, "
                         "the default loader stage of scripts/train.py); 0 = the reference's serial upload inside the hook")
    ap.add_argument("--e2e-input", default=os.environ.get("FSNET_BENCH_E2E_INPUT", "uint8"), choices=["uint8", "float32"],
                    help="what the e2e leg uploads per step: uint8 frames + augmentation plan, normalised on the device by fsnet_augment_frames "
                         "(SURVEY 8(f) N3, default) or the six float32 image tensors per sample the reference's loader delivers")
    ap.add_argument("--e2e-sync", type=int, default=0,
                    help="e2e leg: 1 = block on every step's loss (loss.item()); 0 (default) = copy every step's loss to pinned host "
                         "memory asynchronously and read it one step later, as a logging loop would")
    ap.add_argument("--workload", default="cfg2a", choices=sorted(WORKLOADS), help="cfg2a = BASELINE.json configs[1] (default)")
    args = ap.parse_args()
    global B_PER_GPU, H, W, CONFIG
    wl = WORKLOADS[args.workload]
    B_PER_GPU, H, W, CONFIG = wl["B"], wl["H"], wl["W"], os.path.join(REPO, "configs", wl["config"])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", init_method="env://")

    from fsnet_b200 import _lib
    from fsnet_b200.networks import ops
    from fsnet_b200.data.synthetic import make_batch, make_fisheye_batch
    from vision_base.utils.builder import build
    from vision_base.utils.utils import cfg_from_file, set_random_seed

    if args.backend in ("auto", "tc"):
        if not ops.tc_available():
            raise SystemExit("libfsnet_b200.so lacks the tcgen05 convolution kernels: rebuild with `python -m fsnet_b200.build`")
        ops.set_backend("tc")
    else:                                             # "torch": the same module tree through stock PyTorch / cuDNN (comparator)
        sys.path.insert(0, os.path.join(REPO, "tools"))
        import torch_reference_backend
        torch_reference_backend.enable()
    tf32 = os.environ.get("FSNET_BENCH_TF32", "0") == "1"     # comparator only: what stock PyTorch does by default on this GPU class
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    cfg = cfg_from_file(CONFIG)
    set_random_seed(123)
    model = build(**cfg.meta_arch)
    if world > 1:
        # SyncBN as in the reference (scripts/train.py:101).  Gradients are averaged by the hook's flat NCCL
        # all-reduce instead of DDP's reducer so that the whole step stays one CUDA graph (identical initial
        # weights on every rank come from the shared seed).
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.to(dev)
    model.train()
    from vision_base.networks.optimizers.optimizers import build_optimizer
    optimizer = build_optimizer(model, **cfg.optimizer)
    use_graph = ops.COMPARATOR is None and not args.no_graph
    hook = build(**dict(cfg.trainer.training_hook, cuda_graph=use_graph))
    probe_hook = build(**dict(cfg.trainer.training_hook, cuda_graph=False))     # eager steps for per-kernel event timing

    host = (make_fisheye_batch if wl["fisheye"] else make_batch)(B_PER_GPU, H, W, seed=1234 + rank)
    stage = None
    if args.e2e_input == "uint8" and args.prefetch:
        # the e2e leg's host batch: decoded uint8 frames, the 0/1 mask as uint8 and a 16-double plan per sample (identity geometry, no
        # colour jitter) instead of ('image', f) / ('original_image', f) in float32 and an fp64 mask; the device stage rebuilds those
        import types
        import numpy as np
        from fsnet_b200.data.device_augment import DeviceAugmentStage, GEOM_RESIZE, PLAN_SIZE
        from fsnet_b200.data.synthetic import IMAGENET_MEAN, IMAGENET_STD
        frames = [0, 1, -1]
        u8 = torch.stack([(host[("original_image", f)] * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1) for f in frames], 1)
        plan = torch.zeros(B_PER_GPU, PLAN_SIZE, dtype=torch.float64)
        plan[:, 0], plan[:, 1], plan[:, 2], plan[:, 3], plan[:, 13], plan[:, 14], plan[:, 15] = 1.0, 1.0, W, H, H, W, GEOM_RESIZE
        host_e2e = {k: v for k, v in host.items() if not (isinstance(k, tuple) and k[0] in ("image", "original_image")) and k != "patched_mask"}
        host_e2e.update(frames_u8=u8.contiguous(), aug_plan=plan, mask_u8=host["patched_mask"].to(torch.uint8))
        stage = DeviceAugmentStage(types.SimpleNamespace(frames=frames, output_h=H, output_w=W, mean=np.array(IMAGENET_MEAN, dtype=np.float32),
                                                         std=np.array(IMAGENET_STD, dtype=np.float32)))
    else:
        host_e2e = host
    pinned = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host_e2e.items()}
    resident = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_e2e.values() if torch.is_tensor(v))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pinned_f32 = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()} if stage is not None else None

    def timed(n, use_host, hook=hook, head_start=False, pinned=pinned, stage=stage):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        last = None
        if use_host and args.prefetch:
            # the same n host->device copies and n loss read-backs, inside the timed region; the copy of step i+1 is issued
            # before step i is launched, so it runs on the copy engine under step i
            from fsnet_b200.data.loading import DevicePrefetcher
            reader = LossReader(args.e2e_sync)
            for i, data in enumerate(DevicePrefetcher((dict(pinned) for _ in range(n)), dev, device_transform=stage)):
                out = hook(data, model, optimizer, None, None, i, 0)
                reader.push(out["loss"])
            last = reader.last()
            n = 0
        reader = LossReader(args.e2e_sync) if use_host else None
        for i in range(n):
            if head_start:
                # eager steps are host bound (~30 ms of Python per step): a 60 ms spin kernel lets the host queue the whole step,
                # so that the per-kernel CUDA events below bracket back-to-back GPU execution, not launch latency
                torch.cuda._sleep(120_000_000)
            data = dict(pinned) if use_host else dict(resident)
            out = hook(data, model, optimizer, None, None, i, 0)
            if use_host:
                reader.push(out["loss"])             # device -> host read of the step's result
        if reader is not None and n:
            last = reader.last()
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, last

    timed(max(args.warmup, 5 if use_graph else 3), False)     # includes the eager warm-up calls and the capture
    _lib.reset_counters()
    timed(1, False, probe_hook)                                # one eager step: counts this arm's kernel launches per step
    launches_per_step = _lib.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, _ = timed(args.steps, False)
    launches = launches_per_step * args.steps
    timed(2, True)
    ms_e2e, last_loss = timed(args.steps, True)
    # informational: the same e2e leg fed with what the reference's loader delivers (six float32 images + fp64 mask per sample)
    ms_e2e_f32 = None
    if pinned_f32 is not None:
        timed(2, True, pinned=pinned_f32, stage=None)
        ms_e2e_f32, _ = timed(args.steps, True, pinned=pinned_f32, stage=None)
    clocks = sampler.stop() if rank == 0 else None          # sampled over both timed regions (value and e2e)
    # roofline leg: the same training steps run eagerly so that the kernels can be bracketed with CUDA events on their
    # stream (events cannot be read back from inside a replayed graph)
    # (is_log_image heads need the forward-only kernel for their hm outputs: forward and backward are separate launches)
    LOSS_ENTRY = "fsnet_warp_ssim_fwdbwd" if not wl["fisheye"] else "fsnet_warp_ssim_mei_bwd"
    _lib.profile_entry(LOSS_ENTRY, True)
    def conv_tag(a):
        # (tensor-core products of the launch: 3 = forward, FLOPs 2*N*Ho*Wo*Cout*Cin*KH*KW from the call's own arguments; the network
        # stems read 3 / 6 image channels padded to 8)
        cin = a[0].c if a[0].c != 8 else 3
        return a[9], 2.0 * a[12].n * a[12].h * a[12].w * a[4] * cin * a[5] * a[6]
    _lib.profile_entry("fsnet_conv", True, tag=conv_tag)
    timed(min(args.steps, 5), False, probe_hook, head_start=True)
    kern_us = _lib.profile_results(LOSS_ENTRY)                       # per-launch CUDA-event times (us), scale order
    conv_rows = _lib.profile_results("fsnet_conv", with_tags=True)
    _lib.profile_entry(LOSS_ENTRY, False)
    _lib.profile_entry("fsnet_conv", False)
    probe_steps = min(args.steps, 5)

    if rank != 0:
        return _finish(world)
    imgs = B_PER_GPU * world * args.steps
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    peak, peak_src = peaks()
    # backward-by-recomputation bytes B*HW*(40 + 16/4^s); the fused launch also produces the forward sums
    bytes_per_launch = sum(B_PER_GPU * H * W * (40 + 16 / 4 ** s) for s in SCALES) / len(SCALES)
    avg_us = sum(kern_us) / max(len(kern_us), 1) if kern_us else float("nan")
    achieved = bytes_per_launch / (avg_us * 1e-6) / 1e9 if kern_us else None
    traffic = None
    prof = os.path.join(REPO, "profiles", "r2_loss_pair_ncu.json")
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    # tensor-bound leg: all forward convolutions of one step (34 launches), SURVEY.md 8(d): 17.02 GFLOP per image at cfg2
    fwd_us = sum(t for t, tag in conv_rows if tag[0] == 3) / max(probe_steps, 1)
    counted_flops = sum(tag[1] for t, tag in conv_rows if tag[0] == 3) / max(probe_steps, 1)
    conv_flops = wl["gflop_per_image"] * 1e9 * B_PER_GPU if "gflop_per_image" in wl else counted_flops
    tf_peak = 1364.9
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            tf_peak = float(json.load(f).get("bf16_tflops_sustained", tf_peak))
    conv_tf = conv_flops / (fwd_us * 1e-6) / 1e12 if fwd_us > 0 else None
    # encoder only (what north_star's 50 % tensor-pipe target is quoted on): the first `enc_convs` forward launches of every step
    fwd_rows = [t for t, tag in conv_rows if tag[0] == 3]
    fwd_flops = [tag[1] for t, tag in conv_rows if tag[0] == 3]
    per_step = len(fwd_rows) // max(probe_steps, 1)
    enc_n = wl.get("enc_convs", 0)
    enc_us = (sum(sum(fwd_rows[i * per_step:i * per_step + enc_n]) for i in range(probe_steps)) / max(probe_steps, 1)) if (enc_n and per_step) else 0.0
    enc_flops = wl["enc_gflop_per_image"] * 1e9 * B_PER_GPU if "enc_gflop_per_image" in wl else (sum(fwd_flops[:enc_n]) if per_step else 0.0)
    enc_tf = enc_flops / (enc_us * 1e-6) / 1e12 if enc_us > 0 else None
    line = {
        "metric": f"images/sec ({W}x{H} triplets), full training step", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (convs: " + ops.precision_note() + ")", "data": "synthetic",
        "config": {"workload": wl["text"], "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * world,
                   "parallelism": f"dp{world}" + (" (SyncBN statistics + flat gradient all-reduce over NCCL, inside the step graph)" if world > 1 else ""),
                   "conv_backend": "tc" if ops.COMPARATOR is None else "torch (comparator)",
                   "cuda_graph": use_graph, "e2e_prefetch": bool(args.prefetch),
                   "e2e_input": ("uint8 frames + mask + augmentation plan, warped / normalised on the device (fsnet_augment_frames) on the upload stream"
                                 if stage is not None else "float32 images (6 per sample) + fp64 mask, as the reference's loader delivers them"),
                   "e2e_loss_read": "blocking .item() per step" if args.e2e_sync else "async copy to pinned memory per step, read one step later",
                   "l2": "no explicit flush: one step touches >2 GB of activations, far beyond the 126 MB L2"},
        "roofline": {"kernel": "loss_pair_kernel<0> via fsnet_warp_ssim_fwdbwd (fused warp-SSIM forward+backward, one launch per scale)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "peak_source": peak_src, "launches_timed": len(kern_us), "avg_launch_us": avg_us,
                     "algorithmic_bytes_per_launch": bytes_per_launch},
        "roofline_conv": {"kernel": "conv_halo_kernel<3,*> + conv_tc_kernel<3> (all forward convolutions of one step, tcgen05 implicit GEMM, 3 bf16 products per K-step)",
                          "bound": "tensor", "achieved": conv_tf, "peak": tf_peak, "unit": "TFLOP/s",
                          "frac": (conv_tf / tf_peak) if conv_tf else None, "us_per_step": fwd_us,
                          "algorithmic_flops_per_step": conv_flops, "flops_counted_from_launches": counted_flops,
                          "note": "useful fp32-equivalent FLOPs; the tensor pipe executes 3x as many (bf16x3 split)",
                          "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
        "roofline_conv_encoder": {"kernel": f"conv_halo_kernel<3,4> + conv_tc_kernel<3>, the {enc_n} convolutions of the ResNet encoder (forward)", "bound": "tensor",
                                  "achieved": enc_tf, "achieved_on_pipe": (3 * enc_tf) if enc_tf else None, "peak": tf_peak, "unit": "TFLOP/s",
                                  "frac": (enc_tf / tf_peak) if enc_tf else None, "frac_on_pipe": (3 * enc_tf / tf_peak) if enc_tf else None,
                                  "us_per_step": enc_us, "algorithmic_flops_per_step": enc_flops,
                                  "note": "achieved = useful fp32-equivalent FLOPs; on_pipe = x3 (three bf16 products per K-step)"},
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss,
                "float32_input": (None if ms_e2e_f32 is None else
                                  {"value": imgs / (ms_e2e_f32 / 1e3), "ms_per_step": ms_e2e_f32 / args.steps,
                                   "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v)),
                                   "note": "informational: the same leg uploading the reference loader's float32 tensors"})},
        "gpu_launches": launches, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline and args.workload == "cfg2a":
        split = {}
        # the same workload (B=12 per step), 1 warm-up + 8 timed full steps: ~10-15 s of CPU work on the box's host cores
        rate, sec = oracle_step_rate(B_PER_GPU, 8, 1, split)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"B={B_PER_GPU} triplets per step (the bench workload), 1 warm-up + 8 timed full steps of {sec:.2f} s "
                                          "(oracle/ = CPU restatement of the reference in torch fp32, pinned by reference-generated "
                                          "goldens; not the reference package, which does not travel to the GPU box)",
                                "seconds_per_step_split": split}
        # informational: the SAME module tree and loss kernels with the convolutions / BatchNorm / activations through stock PyTorch
        # (cuDNN), eager, on this GPU -- fp32 as this arm computes, and TF32 as PyTorch runs convolutions by default
        line["gpu_reference"] = gpu_reference(args)
    print(json.dumps(line), flush=True)
    _finish(world)


def gpu_reference(args):
    import subprocess
    out = {"what": "same modules, parameters and loss kernels; network arithmetic through torch.nn.functional / cuDNN (tools/torch_reference_backend.py), "
                   "eager, batch resident in HBM; a comparator, not a product path"}
    for name, flag in (("cudnn_fp32", "0"), ("cudnn_tf32", "1")):
        try:
            env = dict(os.environ, FSNET_BENCH_TF32=flag)
            res = subprocess.run([sys.executable, os.path.abspath(__file__), "--backend", "torch", "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                                  "--workload", args.workload], env=env, capture_output=True, text=True, timeout=300)
            d = json.loads(res.stdout.strip().splitlines()[-1])
            out[name] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"]}
        except Exception as e:       # noqa: BLE001 -- informational leg: never fails the bench line
            out[name] = {"unavailable": repr(e)[:200]}
    return out


def _finish(world):
    """destroy_process_group() blocks for ever once NCCL collectives live inside a captured CUDA graph (seen on
    this stack); the step results are already on the host, so synchronise, flush and leave."""
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
