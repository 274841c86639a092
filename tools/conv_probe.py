"""Probe (GPU): time library formulations of the canonical conv body at batch 32 inside CUDA graphs.
Not part of the product; informs which torch/cuDNN configuration rainbow_b200.model uses."""
import sys
import torch
import torch.nn.functional as F

dev = "cuda"
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
specs = [(4, 32, 8, 4), (32, 64, 4, 2), (64, 64, 3, 1)]
ws = [torch.randn(co, ci, k, k, device=dev) * 0.05 for ci, co, k, s in specs]
bs = [torch.randn(co, device=dev) * 0.05 for ci, co, k, s in specs]
x0 = torch.rand(B, 4, 84, 84, device=dev)


def timeit(fn, name, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for _ in range(reps):
                out = fn()
    except Exception as e:
        print(f"{name:45s} capture failed: {str(e)[:80]}")
        torch.cuda.synchronize()
        return None
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps)
    print(f"{name:45s} {us:8.1f} us per 3-conv pass")
    return out


def fwd_plain(x=x0):
    for (ci, co, k, s), w, b in zip(specs, ws, bs):
        x = F.relu(F.conv2d(x, w, b, stride=s))
    return x


def fwd_fused(x=x0):
    for (ci, co, k, s), w, b in zip(specs, ws, bs):
        x = torch.cudnn_convolution_relu(x, w, b, (s, s), (0, 0), (1, 1), 1)
    return x


def fwd_unfold(x=x0):
    for (ci, co, k, s), w, b in zip(specs, ws, bs):
        n, c, h, wd = x.shape
        ho, wo = (h - k) // s + 1, (wd - k) // s + 1
        cols = F.unfold(x, k, stride=s)                       # [n, c*k*k, ho*wo]
        y = torch.baddbmm(b.view(1, -1, 1), w.view(1, co, -1).expand(n, -1, -1), cols)
        x = F.relu(y).view(n, co, ho, wo)
    return x


for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        tag = f"tf32={int(tf32)} bench={int(bench)}"
        ref = timeit(fwd_plain, f"conv2d+relu NCHW           {tag}")
        o = timeit(fwd_fused, f"cudnn_convolution_relu NCHW {tag}")
        if o is not None and ref is not None:
            print("     max abs diff vs plain:", float((o - ref).abs().max()))
        xcl = x0.contiguous(memory_format=torch.channels_last)
        wcl = [w.contiguous(memory_format=torch.channels_last) for w in ws]

        def fwd_cl():
            x = xcl
            for (ci, co, k, s), w, b in zip(specs, wcl, bs):
                x = F.relu(F.conv2d(x, w, b, stride=s))
            return x
        timeit(fwd_cl, f"conv2d+relu channels_last   {tag}")
    timeit(fwd_unfold, f"unfold+baddbmm              tf32={int(tf32)}")

# backward (plain autograd) for reference
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    wr = [w.clone().requires_grad_(True) for w in ws]
    br = [b.clone().requires_grad_(True) for b in bs]
    gout = torch.randn(B, 64, 7, 7, device=dev)
    for p in wr + br:
        p.grad = torch.zeros_like(p)

    def fb():
        x = x0
        for (ci, co, k, s), w, b in zip(specs, wr, br):
            x = F.relu(F.conv2d(x, w, b, stride=s))
        x.backward(gout)
        return x
    timeit(fb, f"fwd+bwd autograd            tf32={int(tf32)}", reps=5)
