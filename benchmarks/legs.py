"""The measured legs bench.py puts on its JSON line beside the headline (NOT part of the product path):

* ``wholebody_leg``  - BASELINE.json configs[2]: ONE 512x512x1024 volume, window list cut over the ranks (flat partition,
                       peer-memory finalise) - the strong-scaling number of the north star.
* ``ensemble_leg``   - configs[3]: 5-model BraTS ensemble sharded over the ranks, majority vote on rank 0.
* ``kernels_table``  - every non-backbone kernel at cfg2 and cfg4 sizes against the HBM roofline (CUDA events, L2 flushed).
* ``stitch_only``    - the path with a cheap elementwise predictor: this repo's kernels vs the reference's ATen op sequence
                       on the same GPU (benchmarks/aten_baseline.py) - the speed-up attributable to the kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Any, Dict, List, Optional

import numpy as np
import torch

ROI = 96
N_MODELS = 5


class CheapPredictor:
    """logits[:, k] = x * a_k + b_k on the first input channel: as little backbone as a predictor can be (one fused-free
    torch op chain per class), so a step is the stitching path and almost nothing else."""

    def __init__(self, k: int) -> None:
        self.k = k
        self.a = torch.tensor([0.5 + 0.125 * i * (-1) ** i for i in range(k)], dtype=torch.float32)
        self.b = torch.tensor([0.03125 * i - 0.25 for i in range(k)], dtype=torch.float32)

    def __call__(self, model_in: Any, *a: Any, **kw: Any) -> torch.Tensor:
        x = model_in[0] if isinstance(model_in, (tuple, list)) else model_in
        dev = x.device
        return x[:, :1] * self.a.to(dev).view(1, -1, 1, 1, 1) + self.b.to(dev).view(1, -1, 1, 1, 1)


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def wholebody_leg(wl: dict, rank: int, world: int, dev: torch.device, dist: Any, *, steps: int, warmup: int, sw_batch: int,
                  model: Optional[torch.nn.Module] = None, partition: str = "flat") -> Optional[Dict[str, Any]]:
    """Strong scaling of ONE whole-body volume.  Every rank holds only its planes of the synthetic volume (seeded per
    plane range, so the volume is the same whatever the world size... per rank); timing is device time, max over ranks."""
    import medicalsemseg_b200 as mss
    from benchmarks.backbones import build_backbone
    from medicalsemseg_b200 import block, flat
    from medicalsemseg_b200.grid import make_grid

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    if model is None:
        model = build_backbone(wl["backbone"], cin, k).to(dev)
    grid = make_grid((d, h, w), ROI, wl["overlap"])
    v = nb * d * h * w
    info: Dict[str, Any] = {}
    if world == 1:
        gen = torch.Generator().manual_seed(1000)
        vol = torch.randn(wl["shape"], generator=gen).to(dev)

        def step(stats=None):
            with torch.no_grad():
                return mss.sliding_window_infer(vol, model, ROI, wl["overlap"], "gaussian", sw_batch_size=sw_batch, stats=stats,
                                                time_kernels=stats is not None)
        info["partition"] = "single GPU"
        info["windows_per_rank"] = [grid.n_windows]
    else:
        use_flat = partition == "flat"
        if use_flat:
            part = flat.flat_partition(grid, world)
            lo, hi = part.buf_lo[rank], part.buf_hi[rank]
            gen = torch.Generator().manual_seed(1000 + rank)
            slab = torch.randn((nb, cin, hi - lo, h, w), generator=gen).to(dev)
            try:
                with torch.no_grad():  # the first call allocates + rendezvouses the symmetric accumulators
                    flat.sliding_window_infer_flat(slab, model, ROI, wl["overlap"], "gaussian", sw_batch_size=sw_batch,
                                                   volume_is_slab=True, spatial=(d, h, w))
                ok = torch.ones(1, device=dev)
            except Exception as e:  # noqa: BLE001 - no symmetric memory on this box: the NCCL block path stands in
                info["flat_error"] = f"{type(e).__name__}: {e}"[:200]
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            use_flat = bool(ok.item() > 0)
        if use_flat:
            def step(stats=None):
                with torch.no_grad():
                    return flat.sliding_window_infer_flat(slab, model, ROI, wl["overlap"], "gaussian", sw_batch_size=sw_batch,
                                                          volume_is_slab=True, spatial=(d, h, w), stats=stats,
                                                          time_kernels=stats is not None)[0]
            info["partition"] = "flat: contiguous window ranges; finalise reads peer accumulators over NVLink (symmetric memory)"
            info["windows_per_rank"] = [part.n_windows(r) for r in range(world)]
            info["exchange"] = "p2p peer reads inside mss_finalize_gather"
        else:
            bpart = block.block_partition(grid, world)
            blo, bhi = bpart.box(rank, "buf")
            gen = torch.Generator().manual_seed(1000 + rank)
            blk = torch.randn([nb, cin] + [b - a for a, b in zip(blo, bhi)], generator=gen).to(dev)

            def step(stats=None):
                with torch.no_grad():
                    st = block.local_pass(blk, model, grid, bpart, rank, "gaussian", sw_batch_size=sw_batch, stats=stats,
                                          time_kernels=stats is not None, volume_is_block=True)
                    with st.timer("exchange+finalize"):
                        block.exchange_halos(st.acc, bpart, rank, None)
                        return block.finalize_owned(st, bpart, rank)
            info["partition"] = f"blocks {list(bpart.dims)} (D, H, W), NCCL halo exchange"
            info["windows_per_rank"] = [bpart.n_windows(r) for r in range(world)]
            info["exchange"] = "nccl send/recv + mss_halo_add"

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    stats = [mss.InferStats() for _ in range(steps)]
    e0, e1 = _events()
    barrier()
    e0.record()
    for s in range(steps):
        step(stats[s])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kms = [s.kernel_ms() for s in stats]
    names = ("predictor", "extract", "accumulate", "exchange+finalize", "finalize")
    mine = torch.tensor([ms] + [float(np.mean([m.get(n, 0.0) for m in kms])) for n in names], dtype=torch.float64, device=dev)
    if dist is not None:
        allb = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
    else:
        allb = [mine]
    ms_max = max(float(b[0]) for b in allb)
    if rank != 0:
        return None
    info.update({
        "workload": "wholebody", "baseline_config": wl["cfg"], "shape": list(wl["shape"]), "windows": grid.n_windows,
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
        "voxels_per_s": v * steps / (ms_max * 1e-3), "scaling": "strong",
        "accumulate_launches_per_step_rank0": stats[0].n_accumulate_calls,
        "ms_by_rank": {n: [float(b[i + 1]) for b in allb] for i, n in enumerate(names)},
        "speedup_vs_n1": None,
        "note": "speed-up = this voxels_per_s / the N=1 line's (the driver runs N=1,2,4,8 back to back)",
    })
    return info


def ensemble_leg(wl: dict, rank: int, world: int, dev: torch.device, dist: Any, *, steps: int, warmup: int,
                 sw_batch: int) -> Optional[Dict[str, Any]]:
    """configs[3]: model m on rank m % world, label maps gathered on rank 0 (8.9 MB each), majority vote there."""
    import medicalsemseg_b200 as mss
    from benchmarks.backbones import build_backbone

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    mine = [m for m in range(N_MODELS) if m % world == rank]
    models = [build_backbone(wl["backbone"], cin, k, seed=13 + m).to(dev) for m in mine]
    gen = torch.Generator().manual_seed(0)
    vol = torch.randn(wl["shape"], generator=gen).to(dev)
    v = nb * d * h * w
    slots = -(-N_MODELS // world)

    def step(stats=None):
        maps = []
        with torch.no_grad():
            for mdl in models:
                maps.append(mss.sliding_window_infer(vol, mdl, ROI, wl["overlap"], "gaussian", sw_batch_size=sw_batch, stats=stats,
                                                     time_kernels=stats is not None)[0].contiguous())
        if dist is not None:
            buf = torch.zeros((slots, d, h, w), dtype=torch.uint8, device=dev)
            for i, mp in enumerate(maps):
                buf[i] = mp
            out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
            dist.gather(buf, out, dst=0)
            if rank != 0:
                return None
            maps = [out[m % world][m // world] for m in range(N_MODELS)]
        a, b = _events()
        a.record()
        voted = mss.majority_vote(maps, k)
        b.record()
        if stats is not None:
            stats.events.setdefault("vote", []).append((a, b))
        return voted

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    stats = [mss.InferStats() for _ in range(steps)]
    e0, e1 = _events()
    barrier()
    e0.record()
    for s in range(steps):
        step(stats[s])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms = float(t.item())
    kms = [s.kernel_ms() for s in stats]
    return {
        "workload": "brats", "baseline_config": wl["cfg"], "shape": list(wl["shape"]), "ensemble": N_MODELS, "n_gpus": world,
        "models_on_rank0": len(models), "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
        "voxels_per_s": v * steps / (ms * 1e-3), "volumes_per_s": steps / (ms * 1e-3), "scaling": "strong",
        "sharding": "model m on rank m % world; label maps gathered on rank 0; mss_majority_vote there",
        "ms_rank0": {n: float(np.mean([m.get(n, 0.0) for m in kms])) for n in ("predictor", "extract", "accumulate", "vote")},
    }


class _Flusher:
    """Evicts the 126 MB L2 between timed launches (write 256 MB, then read 256 MB of other data)."""

    def __init__(self, dev):
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.src = torch.zeros(64 << 20, dtype=torch.int32, device=dev)

    def __call__(self):
        self.buf.fill_(1)
        self.src.sum()


def _timed(fn, reps: int, flush) -> float:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush()
        a, b = _events()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def kernels_table(dev: torch.device, peak: float, reps: int = 5) -> Dict[str, Any]:
    """Every north-star kernel at cfg2 (512x512x200, K=14) and cfg4 (240x240x155, K=3, 4 channels) sizes: median CUDA-event
    time of one launch with the L2 flushed before it, algorithmic bytes, fraction of the HBM peak."""
    import medicalsemseg_b200 as mss
    from medicalsemseg_b200 import _lib, inferer
    from medicalsemseg_b200.importance import importance_map

    lib = _lib.load()
    flush = _Flusher(dev)
    stream = torch.cuda.current_stream().cuda_stream
    r = ROI ** 3
    out: Dict[str, Any] = {}

    def rec(size: str, name: str, nbytes: int, ms: float, note: str = "") -> None:
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.setdefault(size, {})[name] = {"ms": round(ms, 4), "bytes": int(nbytes), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3),
                                          **({"note": note} if note else {})}

    for size, shape, k, m in (("cfg2", (1, 1, 512, 512, 200), 14, 5), ("cfg4", (1, 4, 240, 240, 155), 3, 5)):
        nb, cin, d, h, w = shape
        v = d * h * w
        plan = inferer.get_plan((d, h, w), ROI, 0.5, dev, nb)
        imp = importance_map((ROI,) * 3, "gaussian", 0.125, dev)
        n_win = plan.grid.n_windows
        B = 8
        vol = inferer._tma_ready(torch.randn(shape, device=dev), plan.grid, 0.0)
        st = inferer.Stitcher(plan, imp, fuse=_lib.FUSE_LABELS, sw_batch=B)
        n_ext = min(n_win, max(B, (1 << 30) // (4 * cin * r) // B * B))
        for mode, nm in ((1, "extract (auto: volume-stationary bulk copies, else TMA / shifted-vector)"), (0, "extract (shifted-vector)")):
            st.use_tma = mode
            ms = _timed(lambda: st.extract(vol, 0, n_ext, 0.0), reps, flush)
            # compulsory bytes: every input voxel the windows touch read once + every patch element written once
            rec(size, nm, 4 * v * cin * min(1.0, n_ext / n_win) + 4 * n_ext * cin * r, ms,
                f"{n_ext} windows; compulsory bytes 4 V Cin (share) + 4 N Cin R")
        n_batches = -(-n_win // B)
        logits = [torch.randn((min(B, n_win - i * B), k, ROI, ROI, ROI), device=dev) for i in range(n_batches)]
        lay = plan.layout(k)
        labels = torch.empty((nb, d, h, w), dtype=torch.uint8, device=dev)
        near = torch.zeros(1, dtype=torch.int64, device=dev)
        ptrs = (C.c_void_p * n_batches)(*[t.data_ptr() for t in logits])

        def acc(fuse, accbuf):
            _lib.check(lib.mss_accumulate(C.byref(lay), ptrs, n_batches, B, _lib.MSS_F32, 0, n_win, imp.data_ptr(),
                                          None if accbuf is None else accbuf.data_ptr(), fuse, labels.data_ptr(), w, 1e-5,
                                          near.data_ptr(), stream), "mss_accumulate")
        ms = _timed(lambda: acc(_lib.FUSE_LABELS, None), reps, flush)
        path = {0: "general", 1: "cell-uniform", 2: "row-staged"}.get(int(lib.mss_accumulate_last_path()), "?")
        rec(size, "accumulate fused->labels", 4 * n_win * k * r + v, ms, f"{n_win} windows, K={k}, one launch, {path} kernel")
        accbuf = torch.empty((nb, k, d, h, plan.pitch_w), device=dev)
        ms = _timed(lambda: acc(_lib.FUSE_LOGITS, accbuf), reps, flush)
        rec(size, "accumulate fused->logits", 4 * n_win * k * r + 4 * v * k, ms)
        del logits
        accbuf.normal_()
        lo, hi = _lib.I3(0, 0, 0), _lib.I3(d, h, w)
        ms = _timed(lambda: _lib.check(lib.mss_finalize_labels(C.byref(lay), accbuf.data_ptr(), imp.data_ptr(), 0, lo, hi,
                                                               labels.data_ptr(), w, None, None, 1e-5, near.data_ptr(), stream),
                                       "mss_finalize_labels"), reps, flush)
        rec(size, "finalize argmax->labels", v * (4 * k + 1), ms)
        del accbuf
        maps = [torch.randint(0, k, (v,), dtype=torch.uint8, device=dev) for _ in range(m)]
        ms = _timed(lambda: mss.majority_vote(maps, k), reps, flush)
        rec(size, f"majority_vote M={m}", v * (m + 1), ms, "one volume per launch")
        mb = [torch.randint(0, k, (8 * v,), dtype=torch.uint8, device=dev) for _ in range(m)]
        ms = _timed(lambda: mss.majority_vote(mb, k), reps, flush)
        rec(size, f"majority_vote M={m}, 8 volumes per launch", 8 * v * (m + 1), ms)
        del mb
        pred = torch.randint(0, k, (v,), dtype=torch.uint8, device=dev)
        lab8 = torch.randint(0, k, (v,), dtype=torch.uint8, device=dev)
        cnt = torch.zeros((3, k), dtype=torch.int64, device=dev)
        ms = _timed(lambda: mss.dice_counts(pred, lab8, k, out=cnt), reps, flush)
        rec(size, "dice_counts u8", 2 * v, ms, "one volume per launch")
        pb = torch.randint(0, k, (8, v), dtype=torch.uint8, device=dev)
        lb = torch.randint(0, k, (8, v), dtype=torch.uint8, device=dev)
        ms = _timed(lambda: mss.dice_counts_batched(pb, lb, k), reps, flush)
        rec(size, "dice_counts_batched 8 volumes", 16 * v, ms, "per-volume counts, one launch (cfg5: a rank's share)")
        del pb, lb
        rows, length = k * h, w * 48
        a = torch.randn(rows, length, device=dev)
        b = torch.randn(rows, length, device=dev)
        ms = _timed(lambda: _lib.check(lib.mss_halo_add(a.data_ptr(), length, b.data_ptr(), length, rows, length, stream), "halo"),
                    reps, flush)
        rec(size, "halo_add (48 planes x K)", 12 * rows * length, ms)
        torch.cuda.empty_cache()
    return out


def stitch_only(wl: dict, dev: torch.device, *, sw_batch: int, steps: int = 3, with_aten: bool = True) -> Dict[str, Any]:
    """The hot path with a cheap elementwise predictor on one GPU, host buffers on both ends: (a) this repo's kernels,
    (b) the reference's ATen op sequence + D2H + host argmax on the same GPU (benchmarks/aten_baseline.py)."""
    import medicalsemseg_b200 as mss
    from benchmarks.aten_baseline import aten_sliding_window_labels

    nb, cin, d, h, w = wl["shape"]
    k = wl["k"]
    v = nb * d * h * w
    pred = CheapPredictor(k)
    gen = torch.Generator().manual_seed(0)
    host_vol = torch.randn(wl["shape"], generator=gen).pin_memory()
    host_labels = torch.empty((nb, d, h, w), dtype=torch.uint8).pin_memory()
    res: Dict[str, Any] = {"workload": wl.get("name"), "predictor": "CheapPredictor: logits[:, k] = x * a_k + b_k (one elementwise op)",
                           "shape": list(wl["shape"]), "classes": k, "sw_batch": sw_batch}

    def ours():
        with torch.no_grad():
            lab = mss.sliding_window_infer(host_vol.to(dev, non_blocking=True), pred, ROI, wl["overlap"], "gaussian",
                                           sw_batch_size=sw_batch)
        host_labels.copy_(lab, non_blocking=True)
    for _ in range(2):
        ours()
    torch.cuda.synchronize()
    a, b = _events()
    a.record()
    for _ in range(steps):
        ours()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    res["ours"] = {"ms_per_volume": ms, "voxels_per_s": v / (ms * 1e-3), "what": "H2D volume -> extract -> predictor -> accumulate "
                   "(fused labels) -> D2H uint8 labels; CUDA events"}
    ours_labels = host_labels.clone()
    if with_aten:
        aten_sliding_window_labels(host_vol, pred, ROI, sw_batch, wl["overlap"], dev)  # warm-up
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(1, steps - 1)):
            t0 = time.perf_counter()
            ref_labels = aten_sliding_window_labels(host_vol, pred, ROI, sw_batch, wl["overlap"], dev)
            ts.append(time.perf_counter() - t0)
        t = float(np.mean(ts))
        mism = int((torch.from_numpy(ref_labels) != ours_labels[0]).sum())
        res["aten_gpu"] = {"ms_per_volume": t * 1e3, "voxels_per_s": v / t,
                           "what": "engine/utils.py:120-151 op sequence on the GPU (slice+cat, imp*logits scatter loop, K-replicated "
                                   "count map, divide) + engine/test.py:140-141 (softmax, D2H of fp32 probabilities, host np.argmax); "
                                   "wall clock with synchronize", "label_mismatch_vs_ours": mism}
        res["speedup_vs_aten_gpu"] = res["aten_gpu"]["ms_per_volume"] / ms
    return res
