"""Cost of the GEMM epilogues for ONE utterance (M = 1027, N = 4096, one-tile CTAs): plain / bias / bias + GELU at one
k-block and at K = 1024, per launch inside a graph:  python tools/epi_probe.py"""
import sys, torch
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo/tts-with-diffusion-model_b200")
from vall_e.b200 import lib as L
L.load()
dev, M, reps = "cuda", 1027, 48
def in_graph(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2] * 1e3 / reps
for K in (64, 1024):
    A = (torch.randn(M, K, device=dev) * 0.1).bfloat16(); W = (torch.randn(4096, K, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(4096, device=dev); out = torch.zeros(M, 4096, dtype=torch.bfloat16, device=dev)
    for name, epi in (("none", L.EPI_NONE), ("bias", L.EPI_BIAS), ("gelu", L.EPI_BIAS_GELU)):
        t = in_graph(lambda: L.gemm_bf16(out, A, W, None if epi == L.EPI_NONE else bias, None, epi))
        print(f"N=4096 K={K} epi={name}: {t:.2f} us", flush=True)
