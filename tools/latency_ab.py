"""Batch-1 denoise-step latency (one graph replay, the C2 shape) and a checksum of the generated
codes, for A/B runs of kernel variants:  VB200_LIB=.../libX.so python tools/latency_ab.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
sys.path.insert(0, str(ROOT))
from vall_e.b200 import lib as L  # noqa: E402
from vall_e.vall_e.diffusion import Diffusion  # noqa: E402
from bench import MODEL, synth_utterance  # noqa: E402

L.load()
dev = torch.device("cuda")
torch.manual_seed(0)
S = 51
model = Diffusion(**MODEL, n_steps=S)
for blk in model.blocks:
    for sub in (blk.attn, blk.ffn):
        torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
model = model.to(dev)
u = synth_utterance(7, 50, 225)
text, proms = [u[0].to(dev)], [u[1].to(dev)]
codes = model.generate_audio(text, proms, resp_lens=[750], seed=3)
chk = int((codes[0].double() * torch.arange(1, 6001, device=dev).view(750, 8)).sum().item())
ses = model._session(text, proms, [750], [0])
table = model._table(dev)
lat = []
for i in range(240):
    ses.t_utt.fill_(S // 2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ses.graph.replay()
    b.record()
    torch.cuda.synchronize()
    lat.append(a.elapsed_time(b))
lat = sorted(lat[40:])
print(f"{L.LIB_PATH.name}: p50 {lat[len(lat) // 2] * 1e3:.1f} us  p10 {lat[len(lat) // 10] * 1e3:.1f} us  p90 {lat[int(len(lat) * .9)] * 1e3:.1f} us  "
      f"codes checksum {chk}")
