#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 U-Net hot path (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one reference training iteration (scripts/train.py:108-133): zero_grad -> UNet forward ->
centre-crop + squeeze of target / weight map -> WeightedCrossEntropyLoss -> backward ->
SGD(lr 1e-4, momentum 0.99) step, on BASELINE.json configs[1]: batch 16 x 1 x 512 x 512 synthetic
DIC-C2DH-HeLa-shaped images with weight maps per GPU (weak scaling), bf16 tensor-core operands
with fp32 accumulation and fp32 master weights. Rank 0 prints ONE JSON line.

  value     img/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e       img/s through the public API with HOST (pinned) buffers: H2D of image / mask / weight
            map and D2H of the loss inside the timed region, every step
  roofline  the dominant kernel class timed in-step with CUDA events on the launching stream
            (ub_plan_profile_*), algorithmic FLOPs / duration against MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference (fp32 torch on the host cores), bounded sample
  --impl reference  times that CPU path alone (rank 0 only) and prints the same line shape
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unet_512x512_train_images_per_sec"
UNIT = "img/s"
BATCH_PER_GPU = 16
SIZE = 512
FLOP_PER_IMG_STEP = 671.28e9      # 3 x fwd - first-layer dgrad
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        d["_source"] = "measured"
        return d
    except Exception:
        d = dict(FALLBACK_PEAKS)
        d["_source"] = "fallback"
        return d


def forward_flops(size: int, base: int = 64, levels: int = 5, n_classes: int = 2) -> float:
    """Executed FLOPs (2 per MAC) of one eval forward of a size x size single-channel tile through
    the reference architecture (models/unet_model.py:73-85,105-146), layer by layer."""
    fl, h, cin = 0.0, size, 1
    skips = []
    for i in range(levels):
        co = base << i
        fl += 2.0 * (h - 2) ** 2 * co * 9 * cin
        fl += 2.0 * (h - 4) ** 2 * co * 9 * co
        h -= 4
        cin = co
        if i < levels - 1:
            skips.append(co)
            h //= 2
    for j in range(levels - 1):
        co = cin // 2
        fl += 2.0 * h * h * cin * 4 * co           # ConvTranspose2d(k=2, s=2)
        h *= 2
        fl += 2.0 * (h - 2) ** 2 * co * 9 * (co + skips[-1 - j])
        fl += 2.0 * (h - 4) ** 2 * co * 9 * co
        h -= 4
        cin = co
    return fl + 2.0 * h * h * n_classes * cin


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0])); smax.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_step_time(steps: int, warmup: int, threads: int | None = None):
    """The reference's CPU path on BASELINE.json configs[0]: one 1x1x512x512 sample per step,
    zero_grad -> forward -> centre-crop + squeeze -> loss -> backward -> SGD(1e-4, 0.99).step()
    (scripts/train.py:114-131).

    Runs the UNMODIFIED reference (`models/unet_model.py` UNet + `utils/losses.py`
    WeightedCrossEntropyLoss + `scripts/train.py` init_weights / center_crop_tensor) from
    baseline/_ref/ (oracle/install_ref.py) or /root/reference when either is present -> kind
    "reference"; otherwise the oracle port (oracle/unet_ref.py, pinned to the reference) -> "port".
    Returns (median s/step, total s, threads, kind)."""
    import torch

    from oracle import ref_loader, unet_ref

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    img, t, w = unet_ref.synthetic_batch(1, size=SIZE, seed=1234, cropped=False)
    ref = ref_loader.load_reference()
    times = []
    if ref is not None and ref.init_weights is not None:
        kind = "reference"
        torch.manual_seed(0)
        model = ref.UNet(1, 2)
        model.apply(ref.init_weights)
        model.train()
        crit = ref.WeightedCrossEntropyLoss()
        opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.99)     # scripts/train.py:97
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            out = model(img)
            size = out.shape[2:]
            tt = ref.center_crop_tensor(t, size).squeeze(1)
            ww = ref.center_crop_tensor(w, size).squeeze(1)
            loss = crit(out, tt, ww)
            loss.backward()
            opt.step()
            float(loss.item())
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    else:
        kind = "port"
        sd = unet_ref.make_state_dict(1, 2, seed=0)
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
                  if v.is_floating_point() and "running" not in k}
        full = dict(sd)
        full.update(params)
        opt = torch.optim.SGD(list(params.values()), lr=1e-4, momentum=0.99)
        size = (unet_ref.out_size(SIZE),) * 2
        tt = unet_ref.center_crop(t, size).squeeze(1)
        ww = unet_ref.center_crop(w, size).squeeze(1)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            bufs = {}
            logits = unet_ref.unet_forward(full, img, training=True, buffers_out=bufs)
            loss = unet_ref.weighted_cross_entropy(logits, tt, ww)
            loss.backward()
            opt.step()
            float(loss.item())
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    med = statistics.median(times)
    return med, sum(times), torch.get_num_threads(), kind


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    med, total, threads, kind = cpu_reference_step_time(args.steps, max(args.warmup, 1))
    value = 1.0 / med
    sample = (f"{args.steps} timed steps of one 1x1x{SIZE}x{SIZE} sample (batch-16 workload sampled "
              f"at 1 image/step; zero_grad + fwd + weighted CE + bwd + SGD step), fp32 torch CPU, "
              f"median step {med:.3f} s")
    what = ("the unmodified reference (models/unet_model.py, utils/losses.py, scripts/train.py "
            "helpers from baseline/_ref)" if kind == "reference" else "CPU oracle port of the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "UNet(1,2) 64->1024 train step (fwd + weighted CE + bwd + SGD), "
                               f"512x512, {what}", "batch_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--grad-wire", default=os.environ.get("UB_GRAD_WIRE", "fp32"), choices=["fp32", "bf16"],
                    help="payload dtype of the data-parallel gradient all-reduce")
    ap.add_argument("--stream-priority", type=int, default=None,
                    help="run on a non-default torch stream of this priority (-1 = above the library's "
                         "weight-gradient side stream) instead of the default stream")
    ap.add_argument("--no-wide", action="store_true", help="skip the BASELINE configs[4] wide-U-Net step")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused = unet_segmentation_b200.optim.FusedSGD, torch = torch.optim.SGD")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.stream_priority is not None:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=args.stream_priority))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    warmup = max(args.warmup, 3)

    from oracle import unet_ref  # synthetic data generator + seeded init only (not timed)
    from unet_segmentation_b200 import _lib, parallel
    from unet_segmentation_b200.loss import WeightedCrossEntropyLoss
    from unet_segmentation_b200.unet import UNet

    lib = _lib.load()
    N = args.batch
    model = UNet(1, 2)
    model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0))   # reference ctor + init_weights
    model = model.to(dev).train()
    parallel.broadcast_parameters(model)
    wire = {"fp32": None, "bf16": torch.bfloat16}[args.grad_wire]
    reducer = parallel.StageGradAllReducer(model, wire_dtype=wire) if world > 1 else None
    criterion = WeightedCrossEntropyLoss()
    if args.optimizer == "fused":     # same update rule as scripts/train.py:97, one fused kernel
        from unet_segmentation_b200.optim import FusedSGD

        opt = FusedSGD(model, lr=1e-4, momentum=0.99)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.99)   # scripts/train.py:97

    # synthetic data, seed = 1234 + rank; full-size mask / weight map as the dataset
    # hands them over (scripts/train.py:108-112), cropped on the device like train.py:118-126
    g = torch.Generator().manual_seed(1234 + rank)
    img_h = (0.4 + 0.2 * torch.rand(N, 1, SIZE, SIZE, generator=g)).pin_memory()
    yy, xx = torch.meshgrid(torch.arange(SIZE), torch.arange(SIZE), indexing="ij")
    m = torch.zeros(N, 1, SIZE, SIZE, dtype=torch.bool)
    for b in range(N):
        for _ in range(10):
            cy, cx = (torch.rand(2, generator=g) * SIZE).tolist()
            ry, rx = (SIZE * (0.08 + 0.12 * torch.rand(2, generator=g))).tolist()
            m[b, 0] |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1.0
    f_fg = m.float().mean().clamp(0.05, 0.95)
    mask_h = m.long().pin_memory()
    wmap_h = torch.where(m, 10 + 1 / f_fg, 10 + 1 / (1 - f_fg)).float().pin_memory()
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()
    h2d_bytes = img_h.numel() * 4 + mask_h.numel() * 8 + wmap_h.numel() * 4
    d2h_bytes = 4

    out_hw = unet_ref.out_size(SIZE)
    s0 = (SIZE - out_hw) // 2

    def crop(t):
        return t[:, :, s0:s0 + out_hw, s0:s0 + out_hw].squeeze(1)

    steps_run = [0]           # every training step executed (for the per-step all-reduce statistics)

    def step(img, mask, wmap):
        steps_run[0] += 1
        opt.zero_grad(set_to_none=True)
        logits = model(img)
        loss = criterion(logits, crop(mask), crop(wmap))
        loss.backward()
        opt.step()
        return loss

    img_d, mask_d, wmap_d = img_h.to(dev), mask_h.to(dev), wmap_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (value) ----------------
    for _ in range(warmup):
        loss = step(img_d, mask_d, wmap_d)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = int(lib.ub_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(img_d, mask_d, wmap_d)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.ub_launch_count()) - launches0
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(loss)
    ms_per_step = ms_total / args.steps
    value = world * N * args.steps / (ms_total * 1e-3)

    # ---------------- end-to-end with host buffers (e2e) ----------------
    # Every step's inputs come from pinned host memory and the loss goes back to the host. The
    # copies of step i+1 are issued on a side stream while step i computes (double buffering, what a
    # DataLoader with pin_memory + a prefetcher does); nothing is reused across steps.
    copy_stream = torch.cuda.Stream(device=dev)
    dbuf = [(torch.empty_like(img_d), torch.empty_like(mask_d), torch.empty_like(wmap_d))
            for _ in range(2)]
    free_ev = [None, None]      # compute finished with buffer set k

    def h2d(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            if free_ev[k] is not None:
                copy_stream.wait_event(free_ev[k])
            for dst, src in zip(dbuf[k], (img_h, mask_h, wmap_h)):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return k, ev

    def e2e_loop(n_steps):
        nxt = h2d(0)
        for i in range(n_steps):
            k, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if i + 1 < n_steps:
                nxt = h2d(i + 1)
            loss = step(*dbuf[k])
            free_ev[k] = torch.cuda.Event()
            free_ev[k].record(torch.cuda.current_stream())
            loss_h.copy_(loss.detach(), non_blocking=True)

    e2e_loop(2)
    barrier()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * N * args.steps / (e2e_ms * 1e-3)

    # ---------------- end-to-end through the device-side sample pipeline (SURVEY §8f N3 + N4) -------
    # What the reference's DataLoader does per sample on the host (utils/dataset.py:69-115 with the
    # augmentation scripts/train.py:34-36 switches on) happens on the device: the host ships the
    # on-disk dtypes (uint8 frame + uint16 instance labels, 3 B/px instead of 16), the device derives
    # the weight maps, applies the elastic deformation (alpha 2000, sigma 20) and builds the
    # float / int64 / cropped tensors; then the same training step. Reported beside `e2e`.
    pipeline = None
    try:
        from unet_segmentation_b200 import input_pipeline as ip

        img8_h = (img_h[:, 0] * 255).round().to(torch.uint8).pin_memory()
        lbl_h = (m[:, 0].to(torch.int32) * 7).to(torch.uint16).pin_memory()
        gen_d = torch.Generator(device=dev).manual_seed(99 + rank)
        prep = ip.DeviceBatchPreparer(dev, (out_hw, out_hw), augment=(2000, 20), generator=gen_d)

        def pipe_step(image, target, weight):
            steps_run[0] += 1
            opt.zero_grad(set_to_none=True)
            loss = criterion(model(image), target, weight)
            loss.backward()
            opt.step()
            return loss

        def pipe_loop(n_steps):
            nxt = prep.submit(img8_h, lbl_h, None)
            for i in range(n_steps):
                k = nxt
                if i + 1 < n_steps:
                    nxt = prep.submit(img8_h, lbl_h, None)
                loss = pipe_step(*prep.get(k))
                loss_h.copy_(loss.detach(), non_blocking=True)

        pipe_loop(2)
        torch.cuda.synchronize()
        l0 = int(lib.ub_launch_count())
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        pipe_loop(args.steps)
        pe1.record()
        torch.cuda.synchronize()
        pipe_local_ms = pe0.elapsed_time(pe1)
        pipeline = {"unit": "img/s",
                    "h2d_bytes_per_step": img8_h.numel() + lbl_h.numel() * 2, "d2h_bytes_per_step": 4,
                    "gpu_launches_per_step": (int(lib.ub_launch_count()) - l0) / args.steps,
                    "stages": "uint8 frames + uint16 labels from pinned host memory -> ub_weight_map "
                              "-> ub_elastic_deform(2000, 20) -> ub_prepare_batch (all on the preparer's "
                              "stream, overlapping the previous step) -> training step"}
    except Exception as exc:   # the extra measurement must never take the contract line down
        pipe_local_ms = float("nan")
        pipeline = {"error": f"{type(exc).__name__}: {exc}"}
    # collectives stay outside the try block so that every rank reaches them
    barrier()
    pipe_ms = max_over_ranks(pipe_local_ms)
    if "error" not in pipeline:
        pipeline["value"] = world * N * args.steps / (pipe_ms * 1e-3)
        pipeline["ms_per_step"] = pipe_ms / args.steps

    # ---------------- in-step kernel timing (roofline) ----------------
    roofline, breakdown = None, None
    peaks = load_peaks()
    if rank == 0:
        plan = next(p for k, p in model._plans.items() if k[4])   # the training plan
        ncls = lib.ub_plan_profile_classes()
        import ctypes as C

        lib.ub_plan_profile_enable(plan.handle, 1)
    prof_steps = min(args.steps, 5)
    for _ in range(prof_steps):
        step(img_d, mask_d, wmap_d)
    barrier()
    if rank == 0:
        ms = (C.c_double * ncls)(); fl = (C.c_double * ncls)(); by = (C.c_double * ncls)()
        ln = (C.c_int * ncls)()
        _lib.check(lib.ub_plan_profile_collect(plan.handle, ms, fl, by, ln), "profile_collect")
        lib.ub_plan_profile_enable(plan.handle, 0)
        breakdown = {}
        for c in range(ncls):
            name = lib.ub_plan_profile_class_name(c).decode()
            if ln[c]:
                breakdown[name] = {"ms_per_step": ms[c] / prof_steps,
                                   "groups_per_step": ln[c] / prof_steps,
                                   "tflops": (fl[c] / (ms[c] * 1e-3) / 1e12) if fl[c] else None,
                                   "gbs": by[c] / (ms[c] * 1e-3) / 1e9}
        dom = max(breakdown, key=lambda k: breakdown[k]["ms_per_step"])
        d = breakdown[dom]
        tensor_bound = d["tflops"] is not None
        # DRAM traffic of that kernel class per launch (= per layer), from the committed ncu captures:
        # the per-class capture of one step (scripts_dev/traffic_by_class.py), else the --set full one
        traffic, traffic_src = None, None
        for fname in ("r02_traffic_by_class.json", "r01_traffic_by_class.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fname)) as fh:
                    tj = json.load(fh)
                if dom in tj:
                    traffic = tj[dom]["dram_bytes_per_step"] / d["groups_per_step"]
                    traffic_src = f"profiles/{fname}: {tj['source']}"
                    break
            except Exception:
                pass
        if tensor_bound:
            peak = float(peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"])
            roofline = {"kernel": dom, "bound": "tensor", "achieved": d["tflops"], "peak": peak,
                        "unit": "TFLOP/s", "frac": d["tflops"] / peak,
                        "frac_of_measured_burst": d["tflops"] / float(peaks["bf16_tflops"]),
                        "frac_of_datasheet_2250": d["tflops"] / 2250.0, "traffic": traffic,
                        "traffic_source": traffic_src,
                        "launches_per_step": d["groups_per_step"],
                        "algorithmic_bytes_per_launch": d["gbs"] * 1e9 * d["ms_per_step"] * 1e-3
                        / d["groups_per_step"],
                        "peak_source": f"{peaks['_source']} sustained bf16 (kernel timed in-step)",
                        "share_of_step": d["ms_per_step"] / ms_per_step}
        else:
            peak = float(peaks["hbm_gbs"])
            roofline = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": peak,
                        "unit": "GB/s", "frac": d["gbs"] / peak, "traffic": traffic,
                        "traffic_source": traffic_src,
                        "peak_source": f"{peaks['_source']} HBM copy",
                        "share_of_step": d["ms_per_step"] / ms_per_step}

    # ---------------- BASELINE configs[4]: wide U-Net (base 128) on 1024^2 crops, batch 8 per GPU ----------
    wide = None
    if not args.no_wide:
        try:
            wide_model = UNet(1, 2, base_channels=128)
            wide_model.load_state_dict(unet_ref.make_state_dict(1, 2, seed=0, base=128))
            wide_model = wide_model.to(dev).train()
            parallel.broadcast_parameters(wide_model)
            wide_red = parallel.StageGradAllReducer(wide_model, wire_dtype=wire) if world > 1 else None
            from unet_segmentation_b200.optim import FusedSGD as _FSGD

            wopt = _FSGD(wide_model, lr=1e-4, momentum=0.99)
            WN, WS = 8, 1024
            gw = torch.Generator().manual_seed(4321 + rank)
            wimg = (0.4 + 0.2 * torch.rand(WN, 1, WS, WS, generator=gw)).to(dev)
            wo = unet_ref.out_size(WS)
            wt = (torch.rand(WN, wo, wo, generator=gw) > 0.55).long().to(dev)
            ww = torch.where(wt > 0, 12.5, 11.66).float()

            def wide_step():
                wopt.zero_grad(set_to_none=True)
                loss = criterion(wide_model(wimg), wt, ww)
                loss.backward()
                wopt.step()
                return loss

            for _ in range(3):
                wide_step()
            wsteps = 5
            barrier()
            we0, we1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            we0.record()
            for _ in range(wsteps):
                wl = wide_step()
            we1.record()
            barrier()
            wide_local = we0.elapsed_time(we1)
            wide = {"final_loss": float(wl), "arena_gb": wide_model.arena_bytes() / 1e9}
            del wide_model, wopt, wide_red, wimg, wt, ww
            torch.cuda.empty_cache()
        except Exception as exc:
            wide_local = float("nan")
            wide = {"error": f"{type(exc).__name__}: {exc}"}
        barrier()
        wide_ms = max_over_ranks(wide_local)
        if "error" not in wide:
            WIDE_FLOP_PER_IMG = 14234.6e9          # base 128 at 1024^2, 3 x fwd - first dgrad
            wide.update({"workload": "UNet(1,2, base 128, depth 5) training step, batch 8/GPU x 1x1024x1024 "
                                     "(BASELINE configs[4]), FusedSGD inside the step",
                         "ms_per_step": wide_ms / 5, "img_per_s": world * 8 * 5 / (wide_ms * 1e-3),
                         "step_tflops_per_gpu": 8 * WIDE_FLOP_PER_IMG / (wide_ms / 5 * 1e-3) / 1e12,
                         "frac_of_bf16_sustained": 8 * WIDE_FLOP_PER_IMG / (wide_ms / 5 * 1e-3) / 1e12
                         / float(peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]),
                         "n_gpus": world, "steps": 5, "warmup": 3})

    # ---------------- overlap-tile inference (BASELINE metric ii), tiles sharded over ranks ----------
    infer = None
    if not args.no_infer:
        from unet_segmentation_b200 import tiling

        model.eval()
        gen = torch.Generator().manual_seed(7)
        for mod in model.modules():      # non-trivial running statistics, identical on every rank
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_((torch.randn(mod.num_features, generator=gen) * 0.1).to(dev))
                mod.running_var.copy_((0.5 + torch.rand(mod.num_features, generator=gen)).to(dev))
        infer = {}
        burst = float(peaks["bf16_tflops"])
        USEFUL_FLOP_PER_PX = 1.468e6          # halo-free lower bound
        for size in (1024, 8192):
            gi = torch.Generator().manual_seed(99)
            frame = 0.4 + 0.2 * torch.rand(512, 512, generator=gi)
            img = frame.repeat(size // 512, size // 512).to(dev)   # mosaic of 512^2 frames
            tile_in, ranks_used, bt = tiling.choose_plan(size, size, world)   # least modelled time
            tile_out, _, origins = tiling.plan_tiles(size, size, tile_in)
            n_tiles = len(origins)
            # plain launches, graph capture, first replay; a small image is timed over enough calls
            # (~0.1 s) that the clock state left behind by the preceding step does not decide the figure
            reps = 5 if size > 2048 else 50
            for _ in range(3 if size > 2048 else 10):
                tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt, rank=rank,
                                            world=world, ranks_used=ranks_used)
            barrier()
            l0 = int(lib.ub_launch_count())
            e0.record()
            for _ in range(reps):
                mask = tiling.overlap_tile_predict(model, img, tile_in=tile_in, batch_tiles=bt,
                                                   rank=rank, world=world, ranks_used=ranks_used)
            e1.record()
            barrier()
            ms_i = max_over_ranks(e0.elapsed_time(e1)) / reps
            exec_flop = n_tiles * forward_flops(tile_in)
            mpix = size * size / (ms_i * 1e-3) / 1e6
            bound = world * burst * 1e12 / USEFUL_FLOP_PER_PX / 1e6
            infer[f"{size}x{size}"] = {
                "mpix_per_s": mpix, "ms": ms_i, "tiles": n_tiles, "tile_in": tile_in,
                "tile_out": tile_out, "batch_tiles": bt, "ranks_used": ranks_used,
                "input_px_per_output_px": n_tiles * tile_in * tile_in / (size * size),
                "useful_tflops": size * size * USEFUL_FLOP_PER_PX / (ms_i * 1e-3) / 1e12,
                "executed_tflops": exec_flop / (ms_i * 1e-3) / 1e12,
                "halo_free_bound_mpix_per_s": bound, "frac_of_halo_free_bound": mpix / bound,
                "executed_frac_of_bf16_burst": exec_flop / (ms_i * 1e-3) / 1e12 / (ranks_used * burst),
                "gpu_launches_per_image": (int(lib.ub_launch_count()) - l0) / reps,
                "fg_fraction": float((mask > 0).float().mean()),
                "path": "ub_extract_tiles (mirror gather) -> eval forward replayed from a CUDA graph "
                        "(head + mask fused into the last conv) -> all-gather of uint8 tiles (N > 1) "
                        "-> ub_stitch_tiles"}
        # whole-image post-processing on a stitched 8192^2 mask: the reference's shipped
        # binary masks (~1 700 speckly components per 324^2 frame) tiled to the full image
        if rank == 0:
            try:
                import numpy as np

                from unet_segmentation_b200.postprocess import get_instance_masks

                blob = np.load(os.path.join(ROOT, "tests", "golden", "ccl_golden.npz"))
                keys = sorted(k for k in blob.files if k.startswith("mask"))
                reps_t = -(-8192 // 324)
                big = np.concatenate([np.concatenate([blob[keys[(r + c) % len(keys)]]
                                                      for c in range(reps_t)], axis=1)
                                      for r in range(reps_t)], axis=0)[:8192, :8192]
                big_d = torch.from_numpy(np.ascontiguousarray(big)).to(dev)
                for _ in range(2):
                    inst = get_instance_masks(big_d, min_size=15)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    inst = get_instance_masks(big_d, min_size=15)
                e1.record()
                torch.cuda.synchronize()
                ms_c = e0.elapsed_time(e1) / 5
                blob_ms = None
                e0.record()
                inst2 = get_instance_masks(mask, min_size=15)      # the network's own (one-blob) mask
                e1.record()
                torch.cuda.synchronize()
                blob_ms = e0.elapsed_time(e1)
                infer["ccl_stitched_8192"] = {
                    "ms": ms_c, "mpix_per_s": 8192 * 8192 / (ms_c * 1e-3) / 1e6,
                    "foreground_fraction": float((big_d > 0).float().mean()),
                    "labels_wrap_uint16": True,
                    "mask": "reference's shipped 01_RES masks (tests/golden/ccl_golden.npz) tiled to 8192^2",
                    "one_blob_mask_ms": blob_ms, "one_blob_max_label": int(inst2.to(torch.int32).max()),
                    "max_label": int(inst.to(torch.int32).max())}
                del big_d, inst, inst2
            except Exception as exc:
                infer["ccl_stitched_8192"] = {"error": f"{type(exc).__name__}: {exc}"}
        # the reference's per-frame predict loop (scripts/predict.py:73-112): one 512^2 frame from host
        # memory -> eval forward -> (softmax[1] > 0.5) mask -> get_instance_masks -> both back on the host
        if rank == 0:
            try:
                from unet_segmentation_b200.postprocess import get_instance_masks

                frame_h = ((frame - 0.5) / 0.5).reshape(1, 1, 512, 512).pin_memory()   # Normalize(.5,.5)
                mask_hb = torch.empty(324, 324, dtype=torch.uint8).pin_memory()
                inst_hb = torch.empty(324, 324, dtype=torch.uint16).pin_memory()
                x_dev = torch.empty(1, 1, 512, 512, device=dev)

                def predict_frame():
                    x_dev.copy_(frame_h, non_blocking=True)
                    _, mk = model.predict_mask(x_dev)
                    inst = get_instance_masks(mk[0], min_size=15)
                    mask_hb.copy_(mk[0], non_blocking=True)
                    inst_hb.copy_(inst, non_blocking=True)

                for _ in range(3):
                    predict_frame()
                torch.cuda.synchronize()
                reps_f = 50
                e0.record()
                for _ in range(reps_f):
                    predict_frame()
                e1.record()
                torch.cuda.synchronize()
                ms_eager = e0.elapsed_time(e1) / reps_f
                # the same loop through the public per-frame API: the whole device side (H2D, forward,
                # connected components, both D2H copies) recorded once and replayed per frame
                import numpy as np

                from unet_segmentation_b200.predict import FramePredictor

                fp = FramePredictor(model, (512, 512), min_size=15)
                frame_np = frame_h.numpy()
                for _ in range(3):
                    mk_np, inst_np = fp(frame_np)
                t0 = time.perf_counter()
                for _ in range(reps_f):
                    mk_np, inst_np = fp(frame_np)         # includes the host-side sync per frame
                ms_f = (time.perf_counter() - t0) * 1e3 / reps_f
                same = bool((mk_np == mask_hb.numpy()).all() and (inst_np == inst_hb.numpy()).all())
                fp8 = FramePredictor(model, (512, 512), min_size=15, batch=8)
                frames8 = np.repeat(frame_np, 8, axis=0)
                for _ in range(3):
                    fp8(frames8)
                t0 = time.perf_counter()
                for _ in range(20):
                    fp8(frames8)
                ms_f8 = (time.perf_counter() - t0) * 1e3 / (20 * 8)
                infer["predict_frame_512"] = {"ms_per_frame": ms_f, "frames_per_s": 1e3 / ms_f,
                                              "eager_ms_per_frame": ms_eager,
                                              "ms_per_frame_batch8": ms_f8, "frames_per_s_batch8": 1e3 / ms_f8,
                                              "timing": "host wall clock around 50 synchronous FramePredictor calls "
                                                        "(numpy in, numpy out); eager = CUDA events around the same "
                                                        "loop issued op by op",
                                              "stages": "H2D 1x1x512x512 f32 -> eval forward (mask fused) -> "
                                                        "8-connected labelling + <15 px filter -> D2H 324x324 "
                                                        "uint8 + uint16, one CUDA-graph launch per frame",
                                              "graph_equals_eager": same,
                                              "max_label": int(inst_np.max())}
            except Exception as exc:   # an extra measurement must not take the contract line down
                infer["predict_frame_512"] = {"error": f"{type(exc).__name__}: {exc}"}
        model.train()

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        med, total, threads, kind = cpu_reference_step_time(steps=5, warmup=2)
        cpu_baseline = {"value": 1.0 / med, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"5 timed + 2 warm-up steps of one 1x1x{SIZE}x{SIZE} sample "
                                  f"(configs[0]: zero_grad + fwd + weighted CE + bwd + SGD step); "
                                  f"{'the unmodified reference from baseline/_ref' if kind == 'reference' else 'oracle port of the reference'}"
                                  f" on {os.cpu_count()} host CPUs, median {med:.3f} s/step"}

    if rank == 0:
        sus = float(peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"UNet(1,2) 64->1024 bf16 training step, batch {N}/GPU x 1x"
                                   f"{SIZE}x{SIZE} + weight maps (BASELINE configs[1]"
                                   f"{'/[2]' if world > 1 else ''})",
                       "global_batch": world * N, "parallelism": f"dp{world}",
                       "optimizer": ("SGD(lr=1e-4, momentum=0.99), FusedSGD (update + bf16 operand "
                                     "refresh in one kernel), inside the step" if args.optimizer == "fused"
                                     else "SGD(lr=1e-4, momentum=0.99) torch foreach, inside the step"),
                       "bn": "per-rank batch statistics",
                       "grad_allreduce": (f"9 per-stage NCCL all-reduces, {args.grad_wire} payload" if world > 1 else None), "l2": "per-step working set 11 GB >> 126 MB L2",
                       "streams": "weight gradients on an internal side stream, overlapping the "
                                  "BN-backward / data-gradient chain (kernel_breakdown is measured "
                                  "with that overlap off)",
                       "first_conv": "fp32 CUDA-core (SURVEY F4)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "step_tflops": world * N * FLOP_PER_IMG_STEP / (ms_per_step * 1e-3) / 1e12,
            "step_frac_of_bf16_sustained": N * FLOP_PER_IMG_STEP / (ms_per_step * 1e-3) / 1e12 / sus,
            "kernel_breakdown": breakdown,
            "infer_overlap_tile": infer,
            "wide_unet": wide,
            "final_loss": final_loss,
            "e2e_device_pipeline": pipeline,
            "allreduce": ({"collectives_per_step": reducer.n_collectives / steps_run[0],
                           "bytes_per_step": reducer.bytes_reduced / steps_run[0]} if reducer else None),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
