"""bench.py --workload wholebody with N > 1 ranks: ONE 512x512x1024 volume (BASELINE.json configs[2]) partitioned into
slabs along its long axis, halo accumulators exchanged over NCCL, every rank finalising the planes it owns.
Strong scaling: total work is fixed, value = voxels of the whole volume / max-over-ranks step time."""
from __future__ import annotations

import json
import os

import numpy as np
import torch


def run_wholebody(args, wl, rank, world, dev, dist) -> None:
    import medicalsemseg_b200 as mss
    from bench import METRIC, ROI, ClockSampler, peaks, physical_gpu_index
    from benchmarks.backbones import build_backbone
    from medicalsemseg_b200 import slab
    from medicalsemseg_b200.grid import make_grid

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    model = build_backbone(wl["backbone"], cin, k).to(dev)
    grid = make_grid((d, h, w), ROI, wl["overlap"])
    part = slab.partition(grid, world)
    ax = part.axis
    # every rank holds only its slab of the synthetic volume (same global random field: seeded per plane block)
    ext = list((d, h, w))
    ext[ax] = part.buf_hi[rank] - part.buf_lo[rank]
    gen = torch.Generator().manual_seed(1000 + rank)
    host_slab = torch.randn([nb, cin] + ext, generator=gen).pin_memory()
    own_planes = part.own_hi[rank] - part.own_lo[rank]
    own_shape = [nb, d, h, w]
    own_shape[1 + ax] = own_planes
    host_labels = torch.empty(own_shape, dtype=torch.uint8).pin_memory()
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def step(slab_any, stats=None, time_kernels=False):
        with torch.no_grad():
            st = slab.local_pass(slab_any, model, grid, part, rank, "gaussian", sw_batch_size=args.sw_batch, stats=stats,
                                 time_kernels=time_kernels, volume_is_slab=True)
            with st.timer("halo"):
                slab.exchange_halos(st.acc, part, rank, None)
            return slab.finalize_owned(st, part, rank)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    dev_slab = host_slab.to(dev)
    for _ in range(args.warmup):
        step(dev_slab)
    barrier()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    stats = [mss.InferStats() for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.steps):
        step(dev_slab, stats[s], True)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for s in range(args.steps):
        labels = step(host_slab)
        host_labels.copy_(labels, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.result()
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    launches = torch.tensor([sum(s.gpu_launches for s in stats)], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)
    kms = [s.kernel_ms() for s in stats]
    mine = torch.tensor([float(np.mean([m.get(n, 0.0) for m in kms])) for n in ("predictor", "extract", "accumulate", "halo", "finalize")],
                        dtype=torch.float64, device=dev)
    allb = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    if rank == 0:
        peak, peak_src = peaks()
        v = nb * d * h * w
        r = ROI**3
        n_win = stats[0].n_windows
        acc_ms = float(np.mean([m.get("accumulate", 0.0) for m in kms]))
        ext_vox = nb * int(np.prod(ext))
        acc_bytes = 4 * n_win * k * r + 4 * ext_vox * k  # logits read once + raw fp32 sums of the slab written once
        line = {
            "metric": METRIC, "value": v * args.steps / (ms_total * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "wholebody", "baseline_config": wl["cfg"], "shape": list(wl["shape"]), "roi": ROI,
                       "overlap": wl["overlap"], "classes": k, "blend": "gaussian", "windows": grid.n_windows,
                       "sw_batch": args.sw_batch, "backbone": wl["backbone"] + " (random init, seed 13, fp32 eager torch)",
                       "partition": {"axis": ax, "window_starts_per_rank": [hi - lo for lo, hi in zip(part.win_lo, part.win_hi)],
                                     "halo_planes": [part.halo(i)[1] - part.halo(i)[0] for i in range(world)]},
                       "l2_policy": "inputs larger than L2"},
            "e2e": {"value": v * args.steps / (ms_e2e * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": host_slab.numel() * 4, "d2h_bytes_per_step": host_labels.numel(),
                    "note": "bytes of rank 0; every rank copies its own slab in and its owned labels out"},
            "gpu_launches": int(launches.item()), "clocks": clocks,
            "roofline": {"kernel": "accumulate_kernel<float> (raw sums, rank 0)", "bound": "hbm",
                         "achieved": acc_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else None, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s",
                         "frac": (acc_bytes / (acc_ms * 1e-3) / 1e9 / peak) if acc_ms > 0 else None, "traffic": None},
            "breakdown_ms_per_step_by_rank": {n: [float(b[i]) for b in allb] for i, n in
                                              enumerate(("predictor", "extract", "accumulate", "halo", "finalize"))},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
