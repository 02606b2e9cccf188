"""One utterance's FFN2 (M = 1027, N = 1024, K = 4096) a few times, for an `ncu --set full` capture of the
single-wave regime:  ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 6 -c 1 python tools/gemm_small_ncu.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

L.load()
dev, M, N, K = "cuda", 1027, 1024, 4096
torch.manual_seed(0)
A = (torch.randn(M, K, device=dev) * 0.1).bfloat16()
Ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(8)]
bias = torch.zeros(N, device=dev)
out = torch.zeros(M, N, device=dev)
for i in range(8):
    L.gemm_bf16(out, A, Ws[i], bias, out, L.EPI_BIAS_RESIDUAL)
torch.cuda.synchronize()
