"""Minimal launches for ncu captures: python tools/kernel_probe.py attn|gemm|posterior|adaln"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

L.load()
dev = "cuda"
what = sys.argv[1] if len(sys.argv) > 1 else "attn"
torch.manual_seed(0)
if what in ("attn", "attn_c3"):
    lens, heads = ([2527] * 8, 16) if what == "attn" else ([1027] * 32, 16)
    M, d = sum(lens), heads * 64
    qkv = torch.randn(M, 3 * d, device=dev).bfloat16()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    out = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        L.flash_attn_varlen(out, qkv, cu, max(lens), heads, 0.125)
elif what == "gemm":
    M = 32864
    A = torch.randn(M, 1024, device=dev).bfloat16()
    W = torch.randn(4096, 1024, device=dev).bfloat16()
    bias = torch.randn(4096, device=dev)
    out = torch.empty(M, 4096, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        L.gemm_bf16(out, A, W, bias, None, L.EPI_BIAS_GELU)
    W2 = torch.randn(1024, 4096, device=dev).bfloat16()
    x = torch.randn(M, 1024, device=dev)
    b2 = torch.randn(1024, device=dev)
    for _ in range(3):
        L.gemm_bf16(x, out, W2, b2, x, L.EPI_BIAS_RESIDUAL)
elif what == "gemm_big":     # FFN1 at the bench shape (256 utterances x 1027 rows)
    M = 262912
    A = torch.randn(M, 1024, device=dev).bfloat16()
    W = torch.randn(4096, 1024, device=dev).bfloat16()
    bias = torch.randn(4096, device=dev)
    out = torch.empty(M, 4096, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        L.gemm_bf16(out, A, W, bias, None, L.EPI_BIAS_GELU)
elif what == "gemm_qkv":     # to_qkv at the bench shape: plain bf16 output
    M = 262912
    A = torch.randn(M, 1024, device=dev).bfloat16()
    W = torch.randn(3072, 1024, device=dev).bfloat16()
    out = torch.empty(M, 3072, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        L.gemm_bf16(out, A, W, None, None, L.EPI_NONE)
elif what == "gemm_out":     # to_out at the bench shape: K = 1024, fp32 residual reduce-add epilogue
    M = 262912
    A = torch.randn(M, 1024, device=dev).bfloat16()
    W = torch.randn(1024, 1024, device=dev).bfloat16()
    bias = torch.randn(1024, device=dev)
    x = torch.randn(M, 1024, device=dev)
    for _ in range(3):
        L.gemm_bf16(x, A, W, bias, x, L.EPI_BIAS_RESIDUAL)
elif what == "head_sample":  # fused classifier + reverse step at the bench shape
    sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
    from vall_e.vall_e import d3pm
    S, K, d, B = 51, 1024, 1024, 256
    rows = 750 * B
    tab = d3pm.scalar_table(S, K, "absorbing").to(dev)
    W = (torch.randn(8 * K, d, device=dev) * 0.03).bfloat16()
    bias = torch.randn(8 * K, device=dev)
    head_in = torch.randn(rows, d, device=dev).bfloat16()
    x_t = torch.full((rows, 8), K // 2, dtype=torch.int32, device=dev)
    row_utt = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(750)
    utt = torch.zeros(B, L.U_STRIDE, dtype=torch.int32, device=dev)
    utt[:, L.U_RESP0] = torch.arange(B, device=dev, dtype=torch.int32) * 750
    t_utt = torch.full((B,), 30, dtype=torch.int32, device=dev)
    o = torch.empty(rows, 8, dtype=torch.int32, device=dev)
    for _ in range(3):
        L.head_posterior_sample(o, None, head_in, W, bias, x_t, row_utt, t_utt, utt, tab, 8, K, L.ABSORBING,
                                L.NOISE_PHILOX, seed=1)
elif what == "posterior":
    sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
    from vall_e.vall_e import d3pm
    S, K, rows = 51, 1024, 24000
    table = d3pm.scalar_table(S, K, "absorbing").to(dev)
    logits = torch.randn(rows, 8 * K, device=dev).half()
    x_t = torch.full((rows, 8), K // 2, dtype=torch.int32, device=dev)
    row_utt = torch.zeros(rows, dtype=torch.int32, device=dev)
    utt = torch.zeros(1, L.U_STRIDE, dtype=torch.int32, device=dev)
    t_utt = torch.tensor([30], dtype=torch.int32, device=dev)
    out = torch.empty(rows, 8, dtype=torch.int32, device=dev)
    for _ in range(3):
        L.posterior_sample_from_logits(out, None, logits, 8 * K, x_t, row_utt, t_utt, utt, table, rows, 8, K,
                                       L.ABSORBING, L.NOISE_PHILOX, seed=1)
elif what == "adaln":
    M, d = 32864, 1024
    x = torch.randn(M, d, device=dev)
    table = torch.randn(52, 2 * d, device=dev)
    out = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
    lv = torch.zeros(1, dtype=torch.int32, device=dev)
    ru = torch.zeros(M, dtype=torch.int32, device=dev)
    for _ in range(3):
        L.adaln(out, x, table, lv, ru)
torch.cuda.synchronize()
print("ok", what)
