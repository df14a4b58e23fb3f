#!/usr/bin/env python
"""bench.py — cUNet G+D training throughput (images/s) at 256x256 on N B200s, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels)
  python bench.py --impl reference [--gpus N] ...                 reference arm: the oracle
        restatement of the reference's own PyTorch CPU path on the box's host cores

A "step" is one training iteration of the reference trainer (t_cls_train.py:288-312 D update +
:226-286 G update, supervised branch, estimator term omitted: SURVEY §8d) on one synthetic batch:
batch 64 per GPU at 256x256 (BASELINE.json configs[1]; `--size 512 --batch 32` = configs[4]), replayed
from its CUDA graph (`--no-graph`: kernel by kernel).  Weak scaling: the per-GPU batch is fixed.
One JSON line on stdout (rank 0).  Besides the contract's keys the line carries: `roofline` (dominant
kernel, all-layer and per-layer fractions of the measured bf16 peak, the HBM-bound kernels),
`cpu_baseline` (oracle train step + the config-1 single-image forward on the host cores),
`gpu_comparator` (the oracle trainer through PyTorch/cuDNN on the same GPU: fp32-TF32 and bf16
autocast), `estimator_plugged_step`, `transfer` / `transfer_dedup` (one image x 1024 signals, device
resident and end to end with every output copied back to pinned host memory), `per_rank` (step time,
SM clock, power of every rank).  `--no-allreduce` is an ablation for the scaling analysis only.
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

UNIT = "images/s"


def metric_name(size):
    return f"cunet_gd_train_images_per_sec_{size}x{size}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--nc", type=int, default=5)
    ap.add_argument("--ref-batch", type=int, default=2, help="images per CPU step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline / transfer legs")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch the iteration kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-allreduce", action="store_true",
                    help="ablation for the scaling analysis: skip the gradient all-reduces (replicas drift)")
    ap.add_argument("--bucket-mb", type=int, default=8,
                    help="size of the gradient all-reduce buckets (scaling analysis)")
    ap.add_argument("--no-comparator", action="store_true",
                    help="skip the PyTorch/cuDNN same-box comparator and the estimator-plugged step")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


def g_conv_flops(H, W):
    """2*MAC of the generator's 15 convolutions, one image, forward (84.81 GF @256)."""
    plan = [(3, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 2), (128, 256, 4), (256, 256, 4),
            (256, 512, 8), (512, 512, 8), (768, 256, 4), (256, 256, 4), (384, 128, 2), (128, 128, 2),
            (192, 64, 1), (64, 64, 1)]
    total = sum(2 * 9 * ci * co * (H // d) * (W // d) for ci, co, d in plan)
    return total + 2 * 64 * 3 * H * W


def d_conv_flops(H, W):
    """2*MAC of the discriminator's 8 convolutions, one image, forward (5.50 GF @256)."""
    total, h, w = 0, H, W
    for cin, cout in ((3, 64), (64, 128), (128, 256), (256, 512)):
        total += 2 * 9 * cin * cin * h * w
        h, w = h // 2, w // 2
        total += 2 * 9 * cin * cout * h * w
    return total


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), \
            float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's CPU train step on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_train_step_rate(size, nc, batch, steps, warmup):
    import torch
    from oracle import train_oracle as T
    from weather_unet_b200 import Conditional_UNet
    from weather_unet_b200.disc import SNDisc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    g_sd = Conditional_UNet(nc).state_dict()
    torch.manual_seed(100)
    d_sd = SNDisc(nc).state_dict()
    tr = T.Trainer(g_sd, d_sd, lr=1e-4)
    g = torch.Generator().manual_seed(1234)
    images = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    c_real = torch.eye(nc)[torch.randint(0, nc, (batch,), generator=g)]
    c_tgt = torch.eye(nc)[torch.randint(0, nc, (batch,), generator=g)]
    for _ in range(warmup):
        tr.step(images, c_real, c_tgt)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(images, c_real, c_tgt)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores


def cpu_forward_rate(size, nc, reps=5, warmup=2):
    """BASELINE configs[0]: generator forward on ONE image with a one-hot condition on the CPU, eval
    mode (demo.py:52-54,79), random-init weights; oracle restatement, fp32, all host cores."""
    import torch
    from oracle import cunet_oracle as O
    from weather_unet_b200 import Conditional_UNet
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    sd = Conditional_UNet(nc).state_dict()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(1, 3, size, size, generator=g) * 2 - 1
    c = torch.eye(nc)[:1]
    with torch.no_grad():
        for _ in range(warmup):
            O.forward(sd, x, c, train=False)
        t0 = time.perf_counter()
        for _ in range(reps):
            O.forward(sd, x, c, train=False)
        dt = time.perf_counter() - t0
    return reps / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warm = max(1, min(args.warmup, 2))
    rate, sec, cores = cpu_train_step_rate(args.size, args.nc, args.ref_batch, steps, warm)
    sample = (f"{steps} timed CPU iterations of the oracle restatement of the reference train step "
              f"(fp32, torch CPU) at batch {args.ref_batch}, {args.size}x{args.size}, after {warm} warm-up")
    line = {
        "impl": "reference", "metric": metric_name(args.size), "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"G+D train step, batch {args.ref_batch} (bounded CPU sample), "
                               f"{args.size}x{args.size}, nc={args.nc}, CPU host cores"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1])), pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# same-box comparators ("the real bar"): the reference's arithmetic through PyTorch + cuDNN
# ------------------------------------------------------------------------------------------------
def comparator_legs(args, dev, resident, trainer, G, D):
    """(i) The oracle trainer — the restatement pinned bit-for-bit to the reference's modules — on
    THIS GPU through PyTorch 2.x / cuDNN: fp32 with TF32 convolutions (what the unmodified reference
    gets on an Ampere-or-later GPU) and under torch.autocast(bf16); same batch, CUDA-event timed.
    (ii) Our iteration with the estimator term plugged in (t_cls_train.py:247-256): a frozen
    random-init torchvision ResNet-101 (num_classes = nc) through PyTorch, the second number SURVEY
    §8d asks for."""
    import torch
    from oracle import train_oracle as T
    from weather_unet_b200.train_step import GDTrainStep
    B, S, nc = args.batch, args.size, args.nc
    out = {}

    def timed(fn, warm, it):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / it

    x, cr, ct = resident[0]
    comp = {}
    for name, autocast in (("fp32_tf32", False), ("autocast_bf16", True)):
        try:
            torch.cuda.empty_cache()
            tr = T.Trainer({k: v.detach().clone() for k, v in G.state_dict().items()},
                           {k: v.detach().clone() for k, v in D.state_dict().items()}, lr=1e-4)
            old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
            xin = x.contiguous(memory_format=torch.channels_last) if autocast else x

            def one():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    tr.step(xin, cr, ct)
            ms = timed(one, 2, 3)
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
            comp[name] = {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
            del tr
        except Exception as e:  # report, never fail the headline
            comp[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
    comp["what"] = ("oracle trainer (pinned bit-exact to the reference modules) on this GPU via PyTorch "
                    f"{torch.__version__} + cuDNN {torch.backends.cudnn.version()}, batch {B}, {S}x{S}, "
                    "torch.optim.Adam, cudnn.benchmark on, losses read with .item() like the reference; "
                    "3 timed iterations after 2 warm-up")
    out["gpu_comparator"] = comp
    torch.cuda.empty_cache()
    try:
        import torchvision
        torch.manual_seed(5)
        est = torchvision.models.resnet101(num_classes=nc).to(dev).eval()
        for p in est.parameters():
            p.requires_grad_(False)
        est = est.to(memory_format=torch.channels_last)

        def estimator(img):  # frozen ResNet-101 through PyTorch / cuDNN under bf16 autocast
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return est(img.contiguous(memory_format=torch.channels_last)).float()
        tr2 = GDTrainStep(G, D, lr=1e-4, estimator=estimator)
        ms = timed(lambda: tr2.step(x, cr, ct), 3, args.steps)
        out["estimator_plugged_step"] = {
            "value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
            "note": "G+D iteration with g_loss_w = MSE(estimator(fake), target) (t_cls_train.py:247-256): "
                    "frozen random-init torchvision resnet101(num_classes=nc) forward + input-gradient "
                    "through PyTorch/cuDNN (bf16 autocast, channels_last); launched kernel by kernel"}
        del tr2, est
    except Exception as e:
        out["estimator_plugged_step"] = {"error": f"{type(e).__name__}: {e}"[:200]}
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from weather_unet_b200 import Conditional_UNet, _ops as K
    from weather_unet_b200.disc import SNDisc
    from weather_unet_b200.train_step import GDTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    B, S, nc = args.batch, args.size, args.nc

    torch.manual_seed(0)
    G = Conditional_UNet(nc).to(dev).train()
    torch.manual_seed(100)
    D = SNDisc(nc).to(dev).train()
    use_graph = not args.no_graph
    trainer = GDTrainStep(G, D, lr=1e-4, static_grads=use_graph, bucket_bytes=args.bucket_mb << 20)
    if args.no_allreduce and trainer.g_buckets is not None:
        trainer.g_buckets.collective = trainer.d_buckets.collective = False

    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4  # a small ring of distinct pinned host batches
    host = []
    for _ in range(n_host):
        img = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).pin_memory()
        cr = torch.eye(nc)[torch.randint(0, nc, (B,), generator=gen)].pin_memory()
        ct = torch.eye(nc)[torch.randint(0, nc, (B,), generator=gen)].pin_memory()
        host.append((img, cr, ct))
    resident = [tuple(t.to(dev) for t in h) for h in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(v):
        if world == 1:
            return [v]
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = v
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    # ---- warm-up, then the device-resident timed region
    n_warm = max(5, args.warmup)  # >= 3 required; 5 lets clocks / allocator / cuDNN autotune settle
    graphed = None
    if use_graph:
        # the whole iteration as ONE CUDA graph (train_step.GraphedGDStep): n_warm eager iterations,
        # then the capture; every timed step = copy the batch into the static buffers + one replay
        from weather_unet_b200.train_step import GraphedGDStep
        graphed = GraphedGDStep(trainer, *resident[0], warmup=n_warm)
        run_step = graphed.step
        for i in range(3):
            run_step(*resident[i % n_host])
    else:
        run_step = trainer.step
        for i in range(n_warm):
            trainer.step(*resident[i % n_host])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()  # every rank samples its own GPU; rank 0's goes into `clocks`, all into `per_rank`
    n0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        losses = run_step(*resident[i % n_host])
    e1.record()
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    ms_own = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_own)
    clocks = sampler.stop()
    per_rank = {"ms_per_step": [round(v / args.steps, 3) for v in gather_ranks(ms_own)],
                "sm_mhz": gather_ranks(float(clocks.get("sm_mhz") or 0.0)),
                "power_w_max": gather_ranks(float(clocks.get("power_w_max") or 0.0))}
    # kernels of THIS library per timed region (a graph replay re-launches what the capture recorded)
    launches = (graphed.library_launches * args.steps) if graphed is not None else K.launch_count() - n0
    value = world * B * args.steps / (ms_total / 1e3)
    last = {k: float(v) for k, v in losses.items()}

    # ---- end to end: pinned host batch -> device every step, losses read back every step.
    # The loop is what a training script with a pinned-memory loader does: the NEXT batch is copied
    # host -> device on a copy stream while the current iteration computes (two device staging slots),
    # and the losses of iteration i are read on the host (pinned buffer, event) while iteration i+1
    # runs; every copy and every read happens inside the timed region, once per step.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [[torch.empty_like(t, device=dev) for t in host[0]] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    loss_host = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    d2h_bytes = loss_host[0].numel() * loss_host[0].element_size()
    main_stream = torch.cuda.current_stream()

    def upload(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])  # the iteration that used this slot has finished
            for dst, src in zip(slots[k], host[i % n_host]):
                dst.copy_(src, non_blocking=True)
            ready[k].record(copy_stream)

    for ev in consumed:
        ev.record(main_stream)
    barrier()
    read_back = []
    e0.record()
    upload(0)
    for i in range(args.steps):
        k = i % 2
        if i + 1 < args.steps:
            upload(i + 1)
        main_stream.wait_event(ready[k])
        losses = run_step(*slots[k])
        consumed[k].record(main_stream)
        loss_host[k].copy_(torch.stack([losses["d_loss"], losses["g_loss"]]), non_blocking=True)
        loss_done[k].record(main_stream)
        if i >= 1:  # host reads the previous iteration's losses while this one runs
            loss_done[1 - k].synchronize()
            read_back.append(loss_host[1 - k].tolist())
    loss_done[(args.steps - 1) % 2].synchronize()
    read_back.append(loss_host[(args.steps - 1) % 2].tolist())
    e1.record()
    barrier()
    assert len(read_back) == args.steps
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    extras = {}
    if not args.no_extras and world == 1:
        # NOT the headline: the same iteration with ONE generator forward shared by the D and the G
        # update (train_step.GDTrainStep(share_fake=True)) — statistically equivalent to the
        # reference's two forwards with independent dropout draws, not bitwise; reported for context
        shared = GDTrainStep(G, D, lr=1e-4, share_fake=True)
        for i in range(3):
            shared.step(*resident[i % n_host])
        barrier()
        e0.record()
        for i in range(args.steps):
            shared.step(*resident[i % n_host])
        e1.record()
        barrier()
        ms_sh = e0.elapsed_time(e1)
        extras["shared_forward_variant"] = {
            "value": B * args.steps / (ms_sh / 1e3), "unit": UNIT, "ms_per_step": ms_sh / args.steps,
            "note": "1 G forward per iteration instead of the reference's 2 (different dropout "
                    "sample for the D update): not the reference schedule, not the headline"}
    roof = None
    peak_burst, peak_sust, hbm, peak_src = peaks()
    if not args.no_extras:
        # ---- roofline of the dominant kernel family, timed alone with CUDA events on this stream:
        # the implicit-GEMM convolution at the generator's largest layer (dconv_up1.0: 192 -> 64 at
        # full resolution, 14.5 GFLOP per image forward), inputs far larger than L2
        # the kernels below are "timed alone" against the BURST peak: let the board leave the power-capped
        # state the training loop put it in (MEASURED_PEAKS' burst figure was taken from idle as well)
        torch.cuda.synchronize()
        time.sleep(2.0)

        def time_ms(fn, it=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(it):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / it

        # (the dominant layer, dconv_up1.0, goes first: right after the pause, before the others warm
        # the board up again)
        layers = [("dconv_up1.0", 128, 64, 64, 1),
                  ("dconv_down1.2", 64, 0, 64, 1), ("dconv_down2.0", 64, 0, 128, 2),
                  ("dconv_down2.2", 128, 0, 128, 2), ("dconv_down3.0", 128, 0, 256, 4),
                  ("dconv_down3.2", 256, 0, 256, 4), ("dconv_down4.0", 256, 0, 512, 8),
                  ("dconv_down4.2", 512, 0, 512, 8), ("dconv_up3.0", 512, 256, 256, 4),
                  ("dconv_up3.2", 256, 0, 256, 4), ("dconv_up2.0", 256, 128, 128, 2),
                  ("dconv_up2.2", 128, 0, 128, 2), ("dconv_up1.2", 64, 0, 64, 1)]
        per_layer, tot = {}, {"fprop": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
        for name, c0, c1, cout, d in layers:
            # every layer is "a kernel timed alone" against the burst peak: a short idle gap before
            # each one, so that the layers at the end of the list are not measured on a board the
            # earlier ones have driven into its power cap (dconv_up1.2 read 10 % low in round 1)
            torch.cuda.synchronize()
            time.sleep(0.4)
            h = S // d
            s0 = torch.randn(B, h, h, c0, device=dev).to(torch.bfloat16)
            s1 = torch.randn(B, h, h, c1, device=dev).to(torch.bfloat16) if c1 else None
            dy = torch.randn(B, h, h, cout, device=dev).to(torch.bfloat16)
            wf, wd = K.pack_conv3x3_weights(torch.randn(cout, c0 + c1, 3, 3, device=dev) * 0.05)
            bias = torch.zeros(cout, device=dev)
            fl = 2.0 * 9 * (c0 + c1) * cout * h * h * B
            tf = time_ms(lambda: K.conv3x3(s0, s1, wf, bias, True, None, cout))
            if c1:
                td = time_ms(lambda: (K.conv3x3(dy, None, wd[:c0], None, False, None, c0),
                                      K.conv3x3(dy, None, wd[c0:], None, False, None, c1)))
            else:
                td = time_ms(lambda: K.conv3x3(dy, None, wd, None, False, s0, c0))
            tw = time_ms(lambda: K.conv3x3_wgrad(s0, s1, dy))
            per_layer[name] = {"gflop": fl / 1e9, "fprop_tflops": fl / tf / 1e9,
                               "dgrad_tflops": fl / td / 1e9, "wgrad_tflops": fl / tw / 1e9}
            for k, t in (("fprop", tf), ("dgrad", td), ("wgrad", tw)):
                tot[k][0] += fl
                tot[k][1] += t
            del s0, s1, dy
        dom = per_layer["dconv_up1.0"]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "conv3x3_igemm_v2_kernel<64, 4> @ dconv_up1.0 fprop (192->64, "
                f"{S}x{S}, batch {B})", "achieved": dom["fprop_tflops"], "peak": peak_burst,
                "unit": "TFLOP/s", "frac": dom["fprop_tflops"] / peak_burst, "traffic": traffic,
                "traffic_source": "profiles/dominant_kernel_traffic.json (one ncu --set full capture of this "
                                  "launch; a citation, not measured in this run)",
                "peak_source": peak_src + ", burst (kernel timed alone)",
                "all_layers": {k: {"tflops": v[0] / v[1] / 1e9, "frac": v[0] / v[1] / 1e9 / peak_burst}
                               for k, v in tot.items()},
                "per_layer": per_layer}
        # ---- the HBM-bound fused ops of the generator, each timed alone at its largest site:
        # algorithmic bytes (tensors read + written once) / time vs the measured HBM peak
        def hbm_line(name, nbytes, fn):
            t = time_ms(fn)
            return {"kernel": name, "algorithmic_bytes": nbytes, "ms": t, "achieved_gbs": nbytes / t / 1e6,
                    "frac_of_hbm_peak": nbytes / t / 1e6 / hbm}

        px = B * S * S
        img_r = resident[0][0]
        w1 = torch.randn(64, 3, 3, 3, device=dev) * 0.2
        b1 = torch.zeros(64, device=dev)
        dy64 = torch.randn(B, S, S, 64, device=dev).to(torch.bfloat16)
        y64 = torch.randn(B, S, S, 64, device=dev).abs().to(torch.bfloat16)
        gp64 = torch.randn(B, S // 2, S // 2, 64, device=dev).to(torch.bfloat16)
        x128 = torch.randn(B, S // 2, S // 2, 128, device=dev).to(torch.bfloat16)
        gu128 = torch.randn(B, S, S, 128, device=dev).to(torch.bfloat16)
        lw, lb = torch.randn(512, nc, device=dev) * 0.3, torch.zeros(512, device=dev)
        cond = resident[0][1]
        wl, bl = torch.randn(3, 64, 1, 1, device=dev) * 0.1, torch.zeros(3, device=dev)
        _, st_ad = K.adain_up_drop(x128, cond, lw, lb, 1e-5, 0.3, 7, None)
        hbm_ops = [
            hbm_line("conv_k27_fprop (3->64 + ReLU, tcgen05 tf32)", px * (12 + 128),
                     lambda: K.conv_first(img_r, w1, b1)),
            hbm_line("conv_k27_wgrad (3->64, tcgen05)", px * (12 + 128),
                     lambda: K.conv_first_wgrad(img_r, dy64)),
            hbm_line("adain_stats + style + adain_up_drop_fwd (128 ch, 128^2 -> 256^2, Philox dropout)",
                     px // 4 * 256 * 2 + px * (256 + 16),
                     lambda: K.adain_up_drop(x128, cond, lw, lb, 1e-5, 0.3, 7, None)),
            hbm_line("adain backward: adjoint + style + apply (128 ch)",
                     px * (256 + 16) + px // 4 * 256 * 5,
                     lambda: K.adain_up_drop_bwd(gu128, x128, cond, lw, lb, st_ad)),
            hbm_line("maxpool2_fwd (64 ch, 256^2)", px * 128 + px // 4 * 128, lambda: K.maxpool2(y64)),
            hbm_line("maxpool2_bwd (64 ch, merges skip + pooled gradients, ReLU mask)",
                     px * 128 * 3 + px // 4 * 128, lambda: K.maxpool2_bwd(y64, gp64, dy64)),
            hbm_line("conv_last_tanh_fprop (64->3)", px * (128 + 12),
                     lambda: K.conv_last_tanh(y64, wl, bl)),
        ]
        roof["hbm_bound_ops"] = {"peak_gbs": hbm, "peak_source": peak_src, "ops": hbm_ops}
        del dy64, y64, gp64, x128, gu128
        # ---- batched transfer inference (inference/inf_1year_signals.py:98-107): one image x 1024
        # signals, sharded by batch across ranks, no collective; train-mode dropout as the reference
        # runs it (the script never calls transfer.eval()).  Two numbers per variant: device-resident
        # (`value`) and end to end (`e2e`): the signals come from pinned host memory and EVERY output
        # image goes back to pinned host memory (the reference hands each one to save_image), the
        # device -> host copies running on a copy stream under the next chunk's compute.
        n_sig = 1024
        per_rank = n_sig // world
        img1 = (torch.rand(1, 3, S, S, generator=gen) * 2 - 1).to(dev)
        sig_host = torch.randn(per_rank, nc, generator=gen).pin_memory()
        sig = sig_host.to(dev)
        chunk = min(128, per_rank)
        rep = img1.repeat(chunk, 1, 1, 1)  # materialised copies, as the reference's DataLoader collates
        out_host = torch.empty((per_rank, 3, S, S), dtype=torch.float32).pin_memory()
        d2h_stream = torch.cuda.Stream(device=dev)
        n_chunks = (per_rank + chunk - 1) // chunk
        out_slots = [torch.empty((chunk, 3, S, S), dtype=torch.float32, device=dev) for _ in range(2)]
        slot_free = [torch.cuda.Event() for _ in range(2)]
        reps_tr = 10

        def transfer(dedup, e2e):
            with torch.no_grad():
                for ci, j in enumerate(range(0, per_rank, chunk)):
                    n = min(chunk, per_rank - j)
                    if e2e:
                        sg = sig_host[j:j + n].to(dev, non_blocking=True)
                        main_stream.wait_event(slot_free[ci % 2])
                    else:
                        sg = sig[j:j + n]
                    y = G(img1 if dedup else rep[:n], sg)
                    if e2e:
                        out_slots[ci % 2][:n].copy_(y)
                        done = torch.cuda.Event()
                        done.record(main_stream)
                        with torch.cuda.stream(d2h_stream):
                            d2h_stream.wait_event(done)
                            out_host[j:j + n].copy_(out_slots[ci % 2][:n], non_blocking=True)
                            slot_free[ci % 2].record(d2h_stream)
            if e2e:
                main_stream.wait_stream(d2h_stream)

        # tensor-core work of one transfer image (forward only): full generator, or decoder + the
        # per-condition half of the concat convolutions when the encoder is shared
        gf_img = g_conv_flops(S, S)
        enc = sum(2 * 9 * ci * co * (S // d) * (S // d) for ci, co, d in
                  [(3, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 2), (128, 256, 4), (256, 256, 4),
                   (256, 512, 8), (512, 512, 8)])
        for key, dedup in (("transfer", False), ("transfer_dedup", True)):
            res = {}
            for e2e in (False, True):
                for ev in slot_free:
                    ev.record(d2h_stream)
                transfer(dedup, e2e)
                barrier()
                e0.record()
                for _ in range(reps_tr):
                    transfer(dedup, e2e)
                e1.record()
                barrier()
                res[e2e] = max_over_ranks(e0.elapsed_time(e1)) / reps_tr
            rate = n_sig / (res[False] / 1e3)
            extras[key] = {
                "metric": f"batched_transfer_images_per_sec_{S}x{S}", "value": rate,
                "unit": UNIT, "signals": n_sig, "reps": reps_tr, "ms_per_sweep": res[False],
                "e2e": {"value": n_sig / (res[True] / 1e3), "unit": UNIT,
                        "h2d_bytes_per_sweep": per_rank * nc * 4 * world,
                        "d2h_bytes_per_sweep": per_rank * 3 * S * S * 4 * world,
                        "ms_per_sweep": res[True]},
                "tensor_roofline": {
                    "executed_conv_tflops": (gf_img * rate if not dedup else
                                             ((gf_img - enc) * rate + enc * rate / chunk)) / 1e12,
                    "faithful_conv_tflops": gf_img * rate / 1e12,
                    "frac_of_sustained_peak": (gf_img * rate if not dedup else
                                               ((gf_img - enc) * rate + enc * rate / chunk)) / 1e12 / world
                    / peak_sust},
                "mode": ("train-mode dropout (faithful), sharded by batch, no collective; " +
                         ("encoder and skip tensors computed once per image (SURVEY §8 f3), "
                          "bit-identical output" if dedup else "replicated image batch, full compute"))}
        del out_host, out_slots

    if not args.no_extras and not args.no_comparator and world == 1:
        extras.update(comparator_legs(args, dev, resident, trainer, G, D))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = 8  # about 10-15 s of host work at batch 2 on the GPU boxes' 16 cores
        rate, sec, cores = cpu_train_step_rate(S, nc, args.ref_batch, n_cpu, 1)
        cpu_base = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{n_cpu} timed CPU iterations (1 warm-up) of the oracle restatement of the "
                              f"reference train step, fp32 torch CPU, batch {args.ref_batch}, {S}x{S}",
                    "config1_forward": {
                        "value": cpu_forward_rate(S, nc), "unit": UNIT,
                        "what": "BASELINE configs[0]: generator forward, ONE image, one-hot condition, "
                                "eval mode (demo.py:52-54,79), fp32 torch CPU, mean of 5 after 2 warm-up"}}

    if rank == 0:
        gf, df = g_conv_flops(S, S), d_conv_flops(S, S)
        g_bwd = 2 * gf - 2 * 9 * 3 * 64 * S * S                 # no data gradient for the image
        flops_step = B * (2 * gf + g_bwd + 3 * df + 6 * df)     # reference-faithful count (388.8 GF/img)
        executed = B * (2 * gf + g_bwd                          # 2 G forward + 1 G backward
                        + 3 * df                                # 3 D forward
                        + 5 * df)                               # 2 D full backward + 1 dgrad-only
        line = {
            "metric": metric_name(S), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"cUNet G + SNDisc D training iteration (t_cls_train.py supervised "
                                   f"branch, estimator term omitted), batch {B}/GPU, {S}x{S}, nc={nc}",
                       "per_gpu_batch": B, "global_batch": B * world, "image": f"{S}x{S}",
                       "parallelism": f"dp{world}" + (" (ABLATION: gradient all-reduce disabled)"
                                                      if args.no_allreduce else ""),
                       "l2": "inputs and activations (GBs per step) far exceed the 126 MB L2; no flush",
                       "launch": ("one CUDA graph replay per iteration (train_step.GraphedGDStep)"
                                  if graphed is not None else "kernel by kernel"),
                       "generator": "sm_100a kernels (this repo)",
                       "discriminator": "sm_100a kernels (this repo; bf16 activations, fp32 spectral norm), "
                                        "512-wide projection head on PyTorch"},
            "clocks": clocks, "per_rank": per_rank, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
            "step_tflops": {"executed_conv_flops_per_step": executed,
                            # `executed` is per GPU and every GPU runs its own step: per-GPU rate
                            "achieved": executed * args.steps / (ms_total / 1e3) / 1e12,
                            "frac_of_sustained_peak": executed * args.steps / (ms_total / 1e3) / 1e12
                            / peak_sust,
                            "reference_faithful_flops_per_step": flops_step},
            "losses_last_step": last,
        }
        if roof is not None:
            line["roofline"] = roof
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        line.update(extras)
        emit(line)
    if world > 1:
        # A captured graph that contains NCCL kernels keeps the communicator busy on teardown (seen:
        # destroy_process_group() never returning at N = 2).  The JSON line is out; release the graph
        # first, and never let teardown outlive a few seconds.
        import threading
        threading.Timer(15.0, lambda: os._exit(0)).start()
        graphed = None
        run_step = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


_JSON_FD = None


def emit(line):
    """The one JSON line, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner at
    # init) is sent to stderr by pointing fd 1 at fd 2 and keeping the original for emit()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
