"""In-graph cost of the denoiser's GEMMs for ONE utterance (M = 1027) per tiling: a CUDA graph of 48
back-to-back launches of one shape, replayed; microseconds per launch including launch gaps.
    python tools/gemm_small_probe.py            (spawns one process per forced tiling)"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) == 1:
    for tile in ("auto", "pair", "wide", "narrow", "n64"):
        env = dict(os.environ)
        if tile != "auto":
            env["VB200_GEMM_TILE"] = tile
        subprocess.run([sys.executable, __file__, tile], env=env, check=False)
    sys.exit(0)

import torch  # noqa: E402
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

L.load()
dev, M, reps = "cuda", int(os.environ.get("PROBE_M", 1027)), 48
torch.manual_seed(0)
shapes = [("qkv", 3072, 1024, L.EPI_NONE, torch.bfloat16), ("to_out", 1024, 1024, L.EPI_BIAS_RESIDUAL, torch.float32),
          ("ffn1", 4096, 1024, L.EPI_BIAS_GELU, torch.bfloat16), ("ffn2", 1024, 4096, L.EPI_BIAS_RESIDUAL, torch.float32)]
res = []
for name, N, K, epi, dt in shapes:
    # distinct weights per launch, as in the model (12 layers): nothing stays hot in a 126 MB L2 except activations
    A = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(1 if os.environ.get('PROBE_SAMEW') else 12)]
    bias = torch.zeros(N, device=dev)
    out = torch.zeros(M, N, dtype=dt, device=dev)

    def run():
        for i in range(reps):
            L.gemm_bf16(out, A, Ws[i % len(Ws)], None if epi == L.EPI_NONE else bias, out if epi == L.EPI_BIAS_RESIDUAL else None, epi)
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res.append(f"{name} {ts[len(ts) // 2] * 1e3 / reps:6.2f}")
print(("L2-hot W " if os.environ.get('PROBE_SAMEW') else "") + f"{sys.argv[1]:7s} M={M}: " + "  ".join(res) + "  us per launch (in graph)")
