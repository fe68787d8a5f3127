#!/usr/bin/env python
"""Headline benchmark: voxels/s of sliding-window inference (ROI 96^3, overlap 0.5, gaussian), BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

One "step" = one full pass of the hot path over one synthetic volume: every window extracted, the backbone run on
every patch batch, all logits stitched, labels produced.  Workloads (BASELINE.json configs):

    btcv       cfg2  1x1x512x512x200, Swin-UNETR-style backbone, K=14, N=400 windows     (default)
    wholebody  cfg3  1x1x512x512x1024, z-slab partitioned across the ranks, N=2100       (--workload wholebody)
    brats      cfg4  1x4x240x240x155, K=3, 5-model ensemble sharded over the ranks -> majority vote
    cfg1             1x1x128^3, UNet, overlap .25, N=8

With N > 1 ranks the headline shards VOLUMES (cfg5: one btcv volume per rank and step, Dice counts all-reduced;
no data-path collective, "scaling": "weak").  The same line carries the other measured legs (benchmarks/legs.py):

    strong_scaling  ONE wholebody volume cut over the N ranks (flat partition, peer-memory finalise): the north-star curve
    ensemble        the 5-model brats ensemble sharded over the ranks + majority vote
    parity          (N > 1) tests/multigpu_parity.py's checks against the CPU oracle, run once before timing
    kernels         (N = 1) every non-backbone kernel at cfg2 and cfg4 sizes against the HBM roofline
    stitch_only     (N = 1) the path with a cheap predictor: this repo vs the reference's ATen op sequence on the same GPU
    cpu_baseline    (N = 1) the reference's CPU path (oracle port) on the host cores: bounded sample + measured legs

--workload wholebody / brats print that leg as the headline instead; --no-legs prints the headline alone.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "btcv": dict(shape=(1, 1, 512, 512, 200), k=14, backbone="swin_unetr", overlap=0.5, cfg="configs[1]"),
    "wholebody": dict(shape=(1, 1, 512, 512, 1024), k=14, backbone="swin_unetr", overlap=0.5, cfg="configs[2]"),
    "brats": dict(shape=(1, 4, 240, 240, 155), k=3, backbone="swin_unetr", overlap=0.5, cfg="configs[3]"),
    "cfg1": dict(shape=(1, 1, 128, 128, 128), k=14, backbone="unet", overlap=0.25, cfg="configs[0]"),
}
ROI = 96
METRIC = "voxels/sec sliding-window inference (ROI 96^3, ov 0.5)"


def acc_kernel_name() -> str:
    """Which accumulation kernel the last mss_accumulate call of this process launched (mss_accumulate_last_path)."""
    from medicalsemseg_b200 import _lib
    return {0: "accumulate_kernel (general)", 1: "accumulate_cells_kernel (cell-uniform, cp.async rings)",
            2: "accumulate_rows_kernel (row-staged, cp.async.bulk ring)"}.get(int(_lib.load().mss_accumulate_last_path()), "?")


def ncu_traffic(key: str):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by scripts/ncu_digest.py runs); None when no capture exists for this workload."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(key)
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms through NVML while the timed region runs."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, threading.Event(), [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self) -> None:
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def result(self) -> dict:
        self.stop_flag.set()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:  # noqa: BLE001
            return local
    return local


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port of engine/utils.py + engine/test.py:140-141), bounded sample
# ---------------------------------------------------------------------------------------------------------------

def config_dict(args, wl) -> dict:
    """The workload description both arms print (identical dicts: same metric on the same configuration)."""
    from medicalsemseg_b200.grid import make_grid

    nb, cin, d, h, w = wl["shape"]
    n_win = make_grid((d, h, w), ROI, wl["overlap"]).n_windows * nb
    return {"workload": args.workload, "baseline_config": wl["cfg"], "shape": list(wl["shape"]), "roi": ROI,
            "overlap": wl["overlap"], "classes": wl["k"], "blend": "gaussian", "windows": n_win, "sw_batch": args.sw_batch,
            "backbone": wl["backbone"] + " (benchmarks/backbones.py, random init, seed 13, fp32 eager torch)",
            "l2_policy": f"inputs larger than L2: {4 * n_win * wl['k'] * ROI ** 3 / 1e9:.1f} GB of logits per step stream "
                         "through the 126 MB L2"}


def cpu_reference_sample(wl: dict, repeats: int, warmup: int, sw_batch: int = 4):
    """Times the oracle on a strip of the workload holding `n_s` windows with the same backbone on the host cores and
    extrapolates to the full volume: T_full = T_stitch+backbone * (N / n_s) + T_labels * (V / V_s).  cfg1 (8 windows)
    is timed whole."""
    from benchmarks.backbones import build_backbone
    from oracle import sliding_window as osw

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nb, cin, d, h, w = wl["shape"]
    whole = d * h * w <= 128 ** 3
    sd, sh = (d, h) if whole else (min(d, ROI), min(h, ROI))
    sample_shape = (1, cin, sd, sh, w)  # one row of windows along W
    model = build_backbone(wl["backbone"], cin, wl["k"])
    rs = np.random.RandomState(0)
    vol = torch.from_numpy(rs.standard_normal(sample_shape).astype(np.float32))
    affine = torch.ones(1, 3)
    _, starts = osw.window_grid((d, h, w), (ROI,) * 3, wl["overlap"])
    n_full = len(starts[0]) * len(starts[1]) * len(starts[2])
    _, s_starts = osw.window_grid(sample_shape[2:], (ROI,) * 3, wl["overlap"])
    n_s = len(s_starts[0]) * len(s_starts[1]) * len(s_starts[2])
    v_full, v_s = d * h * w, sd * sh * w
    times = []
    with torch.no_grad():
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            out = osw.sliding_window_inference(vol, affine, ROI, sw_batch, model, overlap=wl["overlap"], mode="gaussian")
            t1 = time.perf_counter()
            osw.labels_from_logits(out)
            t2 = time.perf_counter()
            if it >= warmup:
                times.append((t1 - t0) * (n_full / n_s) + (t2 - t1) * (v_full / v_s))
    t_full = float(np.mean(times))
    how = "the whole volume, measured" if whole else (
        f"a {sd}x{sh}x{w} strip = {n_s} of {n_full} windows; per-window time scaled by {n_full}/{n_s}, label time by voxels "
        "(EXTRAPOLATED: the full volume needs ~6 min per step on the host)")
    sample = (f"oracle = port of engine/utils.py + engine/test.py:140-141 (the reference itself needs MONAI and is not on the GPU "
              f"box) with the same {wl['backbone']} backbone on {how}; sw_batch {sw_batch}, torch {cores} threads, "
              f"{repeats} timed repeats after {warmup} warm-up")
    return v_full / t_full, t_full, cores, sample


def cpu_measured_legs(sw_batch: int) -> dict:
    """The CPU numbers that are measured, not extrapolated: cfg1 end to end with its UNet, and the
    non-backbone path of cfg2 over ALL 400 windows with the cheap predictor."""
    from benchmarks.backbones import build_backbone
    from benchmarks.legs import CheapPredictor
    from oracle import sliding_window as osw

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}
    wl = WORKLOADS["cfg1"]
    model = build_backbone(wl["backbone"], 1, wl["k"])
    vol = torch.from_numpy(np.random.RandomState(0).standard_normal(wl["shape"]).astype(np.float32))
    with torch.no_grad():
        ts = []
        for it in range(3):
            t0 = time.perf_counter()
            o = osw.sliding_window_inference(vol, torch.ones(1, 3), ROI, 4, model, overlap=wl["overlap"], mode="gaussian",
                                             tuple_input=False)
            osw.labels_from_logits(o)
            ts.append(time.perf_counter() - t0)
    t = float(np.mean(ts[1:]))
    out["cfg1_end_to_end"] = {"seconds": t, "voxels_per_s": 128 ** 3 / t, "windows": 8, "backbone": "unet", "cores": cores,
                              "what": "BASELINE.json configs[0] whole: oracle sliding window + UNet + softmax/argmax, 2 repeats after 1 warm-up"}
    wl = WORKLOADS["btcv"]
    vol = torch.from_numpy(np.random.RandomState(0).standard_normal(wl["shape"]).astype(np.float32))
    with torch.no_grad():
        t0 = time.perf_counter()
        o = osw.sliding_window_inference(vol, None, ROI, sw_batch, CheapPredictor(wl["k"]), overlap=wl["overlap"], mode="gaussian",
                                         tuple_input=False)
        t1 = time.perf_counter()
        osw.labels_from_logits(o)
        t2 = time.perf_counter()
    out["cfg2_stitch_only_all_windows"] = {
        "seconds": t2 - t0, "stitch_seconds": t1 - t0, "labels_seconds": t2 - t1, "voxels_per_s": 512 * 512 * 200 / (t2 - t0),
        "windows": 400, "cores": cores,
        "what": "non-backbone path of configs[1] over ALL 400 windows with the cheap predictor of benchmarks/legs.py, 1 run"}
    return out


def run_reference(args, wl) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, t_full, cores, sample = cpu_reference_sample(wl, max(args.steps, 1), min(args.warmup, 1), args.sw_batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_dict(args, wl),
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------

def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="btcv", choices=sorted(WORKLOADS))
    ap.add_argument("--sw-batch", type=int, default=8,
                    help="windows per backbone call (engine/utils.py sw_batch_size); 8 is 19%% faster than 4 on B200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="headline only (no strong-scaling / ensemble / kernels / stitch-only legs)")
    ap.add_argument("--group-gib", type=float, default=None, help="logits held per accumulate launch (default: auto)")
    ap.add_argument("--partition", default="flat", choices=["flat", "block"], help="wholebody, N > 1: how the volume is cut")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload], name=args.workload)
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import medicalsemseg_b200 as mss
    from benchmarks import legs
    from benchmarks.backbones import build_backbone

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = peaks()

    def finish(line=None):
        if rank == 0 and line is not None:
            print(json.dumps(line))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()

    # ---- parity of the multi-GPU paths against the CPU oracle, once, before anything is timed --------------------------
    parity = None
    if world > 1 and not args.no_legs:
        from tests import multigpu_parity
        res = multigpu_parity.run_checks(rank, world, dev)
        if rank == 0:
            parity = multigpu_parity.summary(res)

    if args.workload in ("wholebody", "brats"):  # that leg as the headline
        sampler = ClockSampler(physical_gpu_index(local))
        sampler.start()
        if args.workload == "wholebody":
            leg = legs.wholebody_leg(wl, rank, world, dev, dist, steps=args.steps, warmup=args.warmup, sw_batch=args.sw_batch,
                                     partition=args.partition)
        else:
            leg = legs.ensemble_leg(wl, rank, world, dev, dist, steps=args.steps, warmup=args.warmup, sw_batch=args.sw_batch)
        clocks = sampler.result()
        line = None
        if rank == 0:
            line = {"metric": METRIC, "value": leg["voxels_per_s"], "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, wl), "clocks": clocks,
                    "leg": leg, "parity": parity,
                    "e2e": None, "gpu_launches": None, "note": "leg-only run: the contract line is the default workload's"}
        finish(line)
        return

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    model = build_backbone(wl["backbone"], cin, k).to(dev)
    gen = torch.Generator().manual_seed(rank)  # rank 0 = the seed-0 volume of SURVEY.md section 8d
    host_vol = torch.randn(wl["shape"], generator=gen).pin_memory()
    dev_vol = host_vol.to(dev)
    host_labels = torch.empty((nb, d, h, w), dtype=torch.uint8).pin_memory()
    label_gt = (torch.arange(d * h * w, device=dev) // 4096 % k).to(torch.uint8).view(d, h, w)  # for the Dice leg (N>1)
    group_bytes = None if args.group_gib is None else int(args.group_gib * (1 << 30))
    v = nb * d * h * w

    def step(volume, stats=None, time_kernels=False):
        with torch.no_grad():
            return mss.sliding_window_infer(volume, model, ROI, wl["overlap"], "gaussian", sw_batch_size=args.sw_batch,
                                            stats=stats, time_kernels=time_kernels, group_bytes=group_bytes)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def dice_leg(labels):
        if dist is None:
            return
        counts = mss.dice_counts(labels[0], label_gt, k)
        dist.all_reduce(counts)  # cfg5's count all-reduce (per-volume Dice across ranks: metrics.gather_volume_counts)

    for _ in range(args.warmup):
        dice_leg(step(dev_vol))
    barrier()

    # --- timed region 1: inputs resident in HBM ------------------------------------------------------------------
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    stats = [mss.InferStats() for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.steps):
        dice_leg(step(dev_vol, stats[s], time_kernels=True))
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)

    # --- timed region 2: end to end through the public API with host buffers ---------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(args.steps):
        labels = step(host_vol.to(dev, non_blocking=True))
        dice_leg(labels)
        host_labels.copy_(labels, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.result()

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()

    line = None
    if rank == 0:
        kms = [s.kernel_ms() for s in stats]
        acc_launch_ms = [x for s in stats for x in s.kernel_launch_ms("accumulate")]
        n_win = stats[0].n_windows
        r = ROI**3
        n_acc = max(stats[0].n_accumulate_calls, 1)
        # algorithmic bytes of the accumulate kernel as built: every logit read once (4 N K R), plus the
        # uint8 label of every voxel written once; when the windows need several launches each launch boundary leaves a
        # slab of unfinished voxels (windows are enumerated D-slowest; groups end on layer boundaries, so the slab is the
        # half-roi overlap of two layers) whose fp32 sums are written by one launch and read back by the next
        acc_bytes = 4 * n_win * k * r + v + (n_acc - 1) * 8 * k * (ROI // 2) * h * w * nb
        acc_ms_step = float(np.mean([m.get("accumulate", 0.0) for m in kms]))
        achieved = acc_bytes / (acc_ms_step * 1e-3) / 1e9 if acc_ms_step > 0 else None
        cfg = config_dict(args, wl)
        line = {
            "metric": METRIC, "value": v * world * args.steps / (ms_total * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "sharding": {"volumes_per_step": world,
                         "how": "one volume per rank, Dice counts all-reduced (cfg5 style)" if world > 1 else "single GPU"},
            "e2e": {"value": v * world * args.steps / (ms_e2e * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": host_vol.numel() * 4, "d2h_bytes_per_step": host_labels.numel()},
            "gpu_launches": int(sum(s.gpu_launches for s in stats)),
            "clocks": clocks,
            "roofline": {
                "kernel": acc_kernel_name() + " (fused normalise+argmax)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_note": "the measured peak is a COPY (half reads, half writes); this kernel is > 99 % reads, which HBM serves a "
                             "little faster than a copy, so frac can read slightly above 1 (B200_PROFILING.md: 'a good kernel "
                             "can read a little above 1.0'); against the 7.7 TB/s HGX figure it is achieved / 7700",
                "traffic": ncu_traffic(f"accumulate_fused_labels_{args.workload}") if n_acc == 1 else None,
                "algorithmic_bytes_per_launch": acc_bytes / n_acc, "launches_per_step": n_acc,
                "avg_launch_ms": float(np.mean(acc_launch_ms)) if acc_launch_ms else None,
                "survey_formula_gbs": (12 * n_win * k * r / (acc_ms_step * 1e-3) / 1e9) if acc_ms_step > 0 else None,
            },
            "breakdown_ms_per_step": {name: float(np.mean([m.get(name, 0.0) for m in kms])) for name in
                                      ("predictor", "extract", "accumulate", "finalize")},
            "volumes_per_s": world * args.steps / (ms_total * 1e-3),
        }
        ext_ms = line["breakdown_ms_per_step"]["extract"]
        if ext_ms > 0:
            comp = 4 * v * cin + 4 * n_win * cin * r
            line["roofline_extract"] = {"achieved": comp / (ext_ms * 1e-3) / 1e9, "frac": comp / (ext_ms * 1e-3) / 1e9 / peak,
                                        "unit": "GB/s", "bytes": comp,
                                        "note": "compulsory bytes: volume read once (4 V Cin) + patches written once (4 N Cin R); "
                                                "re-reads of overlapping windows hit L2"}
    del model, dev_vol, label_gt
    torch.cuda.empty_cache()

    # ---- the other measured legs ----------------------------------------------------------------------------------------
    if not args.no_legs:
        wb = legs.wholebody_leg(dict(WORKLOADS["wholebody"], name="wholebody"), rank, world, dev, dist, steps=2, warmup=1,
                                sw_batch=args.sw_batch, partition=args.partition)
        torch.cuda.empty_cache()
        ens = legs.ensemble_leg(dict(WORKLOADS["brats"], name="brats"), rank, world, dev, dist, steps=2, warmup=1,
                                sw_batch=args.sw_batch)
        torch.cuda.empty_cache()
        if rank == 0:
            line["strong_scaling"] = wb
            line["ensemble"] = ens
            line["parity"] = parity
            if world == 1:
                line["kernels"] = legs.kernels_table(dev, peak)
                line["stitch_only"] = legs.stitch_only(wl, dev, sw_batch=args.sw_batch)
    if rank == 0 and not args.no_cpu_baseline and world == 1:  # the CPU baseline is a rank-0, N=1 number
        val, t_full, cores, sample = cpu_reference_sample(wl, repeats=3, warmup=1, sw_batch=args.sw_batch)
        line["cpu_baseline"] = {"value": val, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample}
        if not args.no_legs:
            line["cpu_baseline"]["measured"] = cpu_measured_legs(args.sw_batch)
            if "stitch_only" in line:
                cpu = line["cpu_baseline"]["measured"]["cfg2_stitch_only_all_windows"]
                line["stitch_only"]["cpu"] = cpu
                line["stitch_only"]["speedup_vs_cpu"] = cpu["seconds"] * 1e3 / line["stitch_only"]["ours"]["ms_per_volume"]
    finish(line)


if __name__ == "__main__":
    main()
