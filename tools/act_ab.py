"""Throughput of the reverse loop for B utterances of the C3 shape, for A/B runs of VB200_ACT=bf16|f16
(or any other environment knob):  VB200_ACT=bf16 python tools/act_ab.py [B]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
sys.path.insert(0, str(ROOT))
from vall_e.b200 import lib as L  # noqa: E402
from vall_e.vall_e.diffusion import Diffusion  # noqa: E402
from bench import MODEL, synth_batch  # noqa: E402

L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
torch.manual_seed(0)
S = 51
model = Diffusion(**MODEL, n_steps=S)
for blk in model.blocks:
    for sub in (blk.attn, blk.ffn):
        torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
model = model.to(dev)
text, proms = synth_batch(B, 50, 225, seed=11)
text, proms = [t.to(dev) for t in text], [p.to(dev) for p in proms]
ses = model._session(text, proms, [750] * B, None)
table = model._table(dev)
for r in range(4):
    ses.x_t.fill_(model.mask_id)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ses.run(table, S, L.ABSORBING, noise=L.NOISE_PHILOX, seed=5)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if r:
        print(f"VB200_ACT={os.environ.get('VB200_ACT', 'default')} B={B}: {ms:8.1f} ms per reverse loop  {B * 6000 / ms:8.1f}k tokens/s", flush=True)
