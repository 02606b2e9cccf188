"""A/B of the interleaved-halves schedule against one session over the whole batch (C3 shape):
    python tools/interleave_ab.py [B] [reps]"""
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
sys.path.insert(0, str(ROOT))
from vall_e.b200 import lib as L  # noqa: E402
from vall_e.b200.engine import BatchLayout, InterleavedSession  # noqa: E402
from vall_e.vall_e.diffusion import Diffusion  # noqa: E402
from bench import MODEL, synth_batch  # noqa: E402

L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
torch.manual_seed(0)
S = 51
model = Diffusion(**MODEL, n_steps=S)
for blk in model.blocks:
    for sub in (blk.attn, blk.ffn):
        torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
model = model.to(dev)
eng = model.engine()
table = model._table(dev)
text, proms = synth_batch(B, 50, 225, seed=11)
text, proms = [t.to(dev) for t in text], [p.to(dev) for p in proms]
lens = [750] * B


def timed(ses, label):
    out = None
    for r in range(reps + 1):
        ses.x_t.fill_(model.mask_id)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ses.run(table, S, L.ABSORBING, noise=L.NOISE_PHILOX, seed=5)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if r:
            print(f"{label}: {ms:8.1f} ms per reverse loop  {B * 6000 / ms:8.1f}k tokens/s", flush=True)
        out = ses.x_t.clone()
    return out


single = eng.session(BatchLayout(text, proms, lens, dev))
ref = timed(single, "one session      ")
if os.environ.get("LITE_ALONE"):
    single.cosched, single.graph = True, None
    timed(single, "one session, small-footprint GEMM tiling")
del single
torch.cuda.empty_cache()
inter = InterleavedSession(eng, text, proms, lens)
if os.environ.get("NO_COSCHED"):
    for h in inter.halves:
        h.cosched = False
got = timed(inter, "interleaved halves" + (" (default GEMM tiling)" if os.environ.get("NO_COSCHED") else ""))
print("codes identical:", bool(torch.equal(ref, got)))
