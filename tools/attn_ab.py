"""A/B timing of attention-kernel variants built by tools/attn_variants.sh.
python tools/attn_ab.py [libname ...]   (no args: every build/variants/lib*.so, plus the shipped library)"""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
VAR = ROOT / "tts-with-diffusion-model_b200" / "build" / "variants"

CHILD = r'''
import sys, torch
sys.path.insert(0, "%s")
from vall_e.b200 import lib as L
L.load()
dev = "cuda"
def run(lens, heads=16, iters=20, amp=1.0):
    torch.manual_seed(1)
    M, d = sum(lens), heads * 64
    qkv = (torch.randn(M, 3 * d, device=dev) * amp).bfloat16()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    out = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
    f = lambda: L.flash_attn_varlen(out, qkv, cu, max(lens), heads, 0.125)
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    return ms, sum(4 * T * T * d for T in lens) / ms / 1e9, out, qkv, cu
# correctness against the CUDA-core kernel
for lens, amp in (([256, 255, 385, 16, 17], 1.0), ([1027, 1027], 1.0), ([2527], 1.0), ([1027, 300], 3.0), ([640], 6.0)):
    ms, tf, out, qkv, cu = run(lens, iters=1, amp=amp)
    ref = torch.empty_like(out)
    L.flash_attn_varlen(ref, qkv, cu, max(lens), 16, 0.125, variant="simt")
    torch.cuda.synchronize()
    err = (out.float() - ref.float()).abs().max().item()
    print(f"  check lens={lens} amp={amp}: max abs err {err:.2e}", "OK" if err < 2.5e-2 * amp else "FAIL")
for lens in ([1027] * 256, [2527] * 8, [1024] * 64, [1027]):
    best = min(run(lens, iters=20 if len(lens) > 1 else 200)[:2] for _ in range(3))     # best of 3: boxes differ in clocks
    print(f"  B={len(lens)} T={lens[0]}: {best[0]:.4f} ms {best[1]:.0f} TFLOP/s", flush=True)
''' % str(ROOT / "tts-with-diffusion-model_b200")

libs = [VAR / f"lib{n}.so" for n in sys.argv[1:]] or sorted(VAR.glob("lib*.so"))
libs = [None] + libs
for lib in libs:
    env = dict(os.environ)
    if lib is not None:
        env["VB200_LIB"] = str(lib)
    print("==", "shipped" if lib is None else lib.name, flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr[-2000:] if r.returncode else "", flush=True)
