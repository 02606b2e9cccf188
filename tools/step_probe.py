"""Denoise-step time of the full model for B utterances of the C3 shape (one graph replay per step), for A/B
runs of environment switches:   VB200_PDL=1 python tools/step_probe.py 32 [reps]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
import bench  # noqa: E402
from vall_e.vall_e.diffusion import Diffusion  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.manual_seed(0)
m = Diffusion(**bench.MODEL, n_steps=51, transition="absorbing")
for blk in m.blocks:
    for sub in (blk.attn, blk.ffn):
        torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
m = m.to(dev)
utts = [bench.synth_utterance(i, 50, 225) for i in range(B)]
text, proms = [u[0].to(dev) for u in utts], [u[1].to(dev) for u in utts]
best = None
for r in range(reps + 1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = m.generate_audio(text, proms, resp_lens=[750] * B, seed=1)
    b.record()
    torch.cuda.synchronize()
    if r:
        ms = a.elapsed_time(b) / 50
        best = ms if best is None else min(best, ms)
cs = int(torch.stack(out).sum().item())
print(f"B={B}: best of {reps}: {best:.4f} ms per denoise step, codes checksum {cs}", flush=True)
