"""Time the soft-resampling kernels alone: K back-to-back launches inside ONE CUDA graph (no per-launch host or graph-launch
overhead), with a rotating set of inputs larger than L2 (--cold) or the same L2-resident input (default).
NFDPF_SOFT_GENERIC=1 selects the generic (run-time strided) kernels for an A/B comparison."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizing_flows_dpfs_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--N", type=int, default=1024)
ap.add_argument("--K", type=int, default=40)
ap.add_argument("--cold", action="store_true")
a = ap.parse_args()
B, N, K = a.B, a.N, a.K
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
nset = 8 if a.cold else 1          # 8 x (36 + 36) MB of distinct inputs / outputs > 126 MB of L2
sets = []
for s in range(nset):
    w = torch.softmax(torch.randn(B, N, device=dev, generator=g) * 2, -1)
    x = torch.randn(B, N, 2, device=dev, generator=g) * 20
    off = torch.rand(B, device=dev, generator=g) / N
    sets.append(dict(w=w, x=x, off=off, xo=torch.empty_like(x), wo=torch.empty_like(w), lw=torch.empty_like(w),
                     idx=torch.empty(B, N, dtype=torch.int64, device=dev), saved=torch.empty(B, 2, device=dev),
                     gx=torch.randn(B, N, 2, device=dev, generator=g), gw=torch.randn(B, N, device=dev, generator=g),
                     glw=torch.randn(B, N, device=dev, generator=g), dx=torch.empty_like(x), dw=torch.empty_like(w)))
mk = torch.linspace(0.0, (N - 1.0) / N, N, device=dev)


def fwd(s):
    L.call("nfdpf_soft_resample_fwd", L.ptr(s["x"]), L.ptr(s["w"]), L.ptr(s["off"]), L.ptr(mk), 0.5, B, N, 2, L.ptr(s["xo"]), L.ptr(s["wo"]),
           L.ptr(s["idx"]), L.ptr(s["saved"]), L.ptr(s["lw"]), None, L.stream())


def bwd(s):
    L.call("nfdpf_soft_resample_bwd", L.ptr(s["gx"]), L.ptr(s["gw"]), L.ptr(s["w"]), L.ptr(s["idx"]), L.ptr(s["saved"]), 0.5, B, N, 2,
           L.ptr(s["dx"]), L.ptr(s["dw"]), L.ptr(s["glw"]), None, L.stream())


for name, fn in (("fwd", fwd), ("bwd", bwd)):
    for s in sets:
        fwd(s)
        fn(s)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(K):
                fn(sets[k % nset])
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K * 1e3)
    byts = B * N * (36 if name == "fwd" else 36)
    print("soft_resample_%s B=%d N=%d %s: %.2f us per launch, %.0f GB/s algorithmic (36 B per particle)%s" %
          (name, B, N, "cold (inputs rotate through > L2)" if a.cold else "L2-resident", best, byts / best * 1e-3,
           " [generic kernels]" if os.environ.get("NFDPF_SOFT_GENERIC") else ""))
