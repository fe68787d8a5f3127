"""bench.py --workload brats: BASELINE.json configs[3] - a 5-model ensemble on one BraTS-shaped 4-channel volume
(240x240x155, 3 classes), models sharded over the ranks (model m on rank m % world, no communication while they run),
label maps gathered on rank 0 (8.9 MB each) and merged by the majority-vote kernel (majority_vote.py:23-37)."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

N_MODELS = 5


def run_ensemble(args, wl, rank, world, dev, dist) -> None:
    import medicalsemseg_b200 as mss
    from bench import METRIC, ROI, ClockSampler, peaks, physical_gpu_index
    from benchmarks.backbones import build_backbone

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    mine = [m for m in range(N_MODELS) if m % world == rank]
    models = []
    for m in mine:
        models.append(build_backbone(wl["backbone"], cin, k, seed=13 + m).to(dev))  # five differently initialised members
    gen = torch.Generator().manual_seed(0)
    host_vol = torch.randn(wl["shape"], generator=gen).pin_memory()  # the same volume on every rank
    dev_vol = host_vol.to(dev)
    host_out = torch.empty((d, h, w), dtype=torch.uint8).pin_memory()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    v = nb * d * h * w

    def step(volume, stats=None, time_kernels=False):
        maps = []
        with torch.no_grad():
            for mdl in models:
                maps.append(mss.sliding_window_infer(volume, mdl, ROI, wl["overlap"], "gaussian", sw_batch_size=args.sw_batch,
                                                     stats=stats, time_kernels=time_kernels)[0].contiguous())
        if dist is not None:
            # every rank contributes a fixed number of slots so one all_gather moves everything (empty slots are skipped)
            slots = -(-N_MODELS // world)
            buf = torch.zeros((slots, d, h, w), dtype=torch.uint8, device=dev)
            for i, mp in enumerate(maps):
                buf[i] = mp
            out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
            dist.gather(buf, out, dst=0)
            if rank != 0:
                return None
            allmaps = [out[m % world][m // world] for m in range(N_MODELS)]
        else:
            allmaps = maps
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        voted = mss.majority_vote(allmaps, k)
        b.record()
        if stats is not None:
            stats.events.setdefault("vote", []).append((a, b))
            stats.gpu_launches += 1
        return voted

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev_vol)
    barrier()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    stats = [mss.InferStats() for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.steps):
        step(dev_vol, stats[s], True)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(args.steps):
        voted = step(host_vol.to(dev, non_blocking=True))
        if voted is not None:
            host_out.copy_(voted, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.result()
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    if rank == 0:
        peak, peak_src = peaks()
        kms = [s.kernel_ms() for s in stats]
        vote_ms = float(np.mean([m.get("vote", 0.0) for m in kms]))
        acc_ms = float(np.mean([m.get("accumulate", 0.0) for m in kms])) / max(len(models), 1)
        n_win = stats[0].n_windows
        r = ROI ** 3
        acc_bytes = 4 * n_win * k * r + v
        line = {
            "metric": METRIC, "value": v * args.steps / (ms_total * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "brats", "baseline_config": wl["cfg"], "shape": list(wl["shape"]), "roi": ROI,
                       "overlap": wl["overlap"], "classes": k, "blend": "gaussian", "windows": n_win, "sw_batch": args.sw_batch,
                       "ensemble": N_MODELS, "models_on_rank0": len(models),
                       "backbone": wl["backbone"] + " (random init, seeds 13..17, fp32 eager torch)",
                       "sharding": "model m on rank m % world; label maps gathered on rank 0, majority vote there",
                       "l2_policy": "one ensemble inference per step; per-kernel numbers with L2 eviction in profiles/"},
            "e2e": {"value": v * args.steps / (ms_e2e * 1e-3), "unit": "voxels/s", "h2d_bytes_per_step": host_vol.numel() * 4,
                    "d2h_bytes_per_step": host_out.numel()},
            "gpu_launches": int(sum(s.gpu_launches for s in stats)), "clocks": clocks,
            "roofline": {"kernel": "accumulate_kernel<float> (fused normalise+argmax, K=3)", "bound": "hbm",
                         "achieved": acc_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else None, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s",
                         "frac": (acc_bytes / (acc_ms * 1e-3) / 1e9 / peak) if acc_ms > 0 else None, "traffic": None},
            "vote": {"ms": vote_ms, "achieved_gbs": v * (N_MODELS + 1) / (vote_ms * 1e-3) / 1e9 if vote_ms > 0 else None},
            "breakdown_ms_per_step": {name: float(np.mean([m.get(name, 0.0) for m in kms])) for name in
                                      ("predictor", "extract", "accumulate", "vote")},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
