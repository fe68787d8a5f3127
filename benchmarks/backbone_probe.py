"""Times the benchmark backbone's forward under a few stock PyTorch settings (no kernels of ours involved)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks.backbones import build_backbone

def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def main():
    dev = torch.device("cuda")
    for bs in (4, 8):
        x = torch.randn(bs, 1, 96, 96, 96, device=dev)
        for name, setup in [
            ("default", dict()),
            ("cudnn.benchmark", dict(bench=True)),
            ("channels_last_3d", dict(cl=True)),
            ("benchmark+channels_last", dict(bench=True, cl=True)),
            ("benchmark+tf32 matmul", dict(bench=True, tf32=True)),
            ("benchmark+bf16 autocast", dict(bench=True, amp=torch.bfloat16)),
        ]:
            torch.backends.cudnn.benchmark = bool(setup.get("bench"))
            torch.backends.cuda.matmul.allow_tf32 = bool(setup.get("tf32"))
            m = build_backbone("swin_unetr", 1, 14).to(dev)
            xin = x
            if setup.get("cl"):
                m = m.to(memory_format=torch.channels_last_3d)
                xin = x.contiguous(memory_format=torch.channels_last_3d)
            def f():
                with torch.no_grad():
                    if setup.get("amp"):
                        with torch.autocast("cuda", dtype=setup["amp"]):
                            return m(xin)
                    return m(xin)
            ms = timeit(f)
            print(f"bs={bs} {name:28s} {ms:8.2f} ms/forward  {ms/bs:7.2f} ms/window", flush=True)
    # where does the time go (default settings, bs=4)
    torch.backends.cudnn.benchmark = True
    m = build_backbone("swin_unetr", 1, 14).to(dev)
    x = torch.randn(4, 1, 96, 96, 96, device=dev)
    with torch.no_grad():
        m(x); m(x)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m(x)
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))

if __name__ == "__main__":
    main()
