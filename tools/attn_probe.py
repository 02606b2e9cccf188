"""One attention launch shape, a few iterations: the command ncu wraps (tools/attn_ab.py is the timing A/B).
python tools/attn_probe.py [B] [T] [iters]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
heads, dev = 16, "cuda"
torch.manual_seed(1)
lens = [T] * B
M, d = sum(lens), heads * 64
qkv = torch.randn(M, 3 * d, device=dev).bfloat16()
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
out = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        a.record()
    L.flash_attn_varlen(out, qkv, cu, T, heads, 0.125)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(f"B={B} T={T}: {ms:.3f} ms {B * 4 * T * T * d / ms / 1e9:.0f} TFLOP/s")
