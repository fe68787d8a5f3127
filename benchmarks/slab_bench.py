"""bench.py --workload wholebody with N > 1 ranks: ONE 512x512x1024 volume (BASELINE.json configs[2]) partitioned into
blocks of windows (medicalsemseg_b200/block.py: 1-D slabs are the special case; 2x2x2 on 8 ranks), halo accumulators
exchanged over NCCL axis by axis, every rank finalising the box it owns.
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
    from medicalsemseg_b200 import block
    from medicalsemseg_b200.grid import make_grid

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    model = build_backbone(wl["backbone"], cin, k).to(dev)
    grid = make_grid((d, h, w), ROI, wl["overlap"])
    dims = None if not getattr(args, "block_dims", None) else tuple(int(x) for x in args.block_dims.split("x"))
    part = block.block_partition(grid, world, dims)
    # every rank holds only its block of the synthetic volume
    blo, bhi = part.box(rank, "buf")
    ext = [hh - ll for ll, hh in zip(blo, bhi)]
    gen = torch.Generator().manual_seed(1000 + rank)
    host_slab = torch.randn([nb, cin] + ext, generator=gen).pin_memory()
    olo, ohi = part.box(rank, "own")
    host_labels = torch.empty([nb] + [hh - ll for ll, hh in zip(olo, ohi)], dtype=torch.uint8).pin_memory()
    local = int(os.environ.get("LOCAL_RANK", "0"))

    def step(slab_any, stats=None, time_kernels=False):
        with torch.no_grad():
            st = block.local_pass(slab_any, model, grid, part, rank, "gaussian", sw_batch_size=args.sw_batch, stats=stats,
                                  time_kernels=time_kernels, volume_is_block=True, peer_group=None if p2p else False)
            with st.timer("halo"):
                halo_bytes[0] = (block.exchange_halos_p2p(st.acc, part, rank, None) if p2p
                                 else block.exchange_halos(st.acc, part, rank, None))
            return block.finalize_owned(st, part, rank)

    halo_bytes = [0]
    p2p = getattr(args, "halo", "nccl") == "p2p" and block.can_exchange_p2p(part)

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
                       "partition": {"ranks_per_axis_dhw": list(part.dims),
                                     "windows_per_rank": [part.n_windows(i) for i in range(world)],
                                     "halo": "p2p (symmetric memory, peer reads)" if p2p else "nccl send/recv + add",
                                     "halo_bytes_received_rank0": int(halo_bytes[0])},
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
