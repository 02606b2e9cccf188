"""Soak test for launch-order races (programmatic dependent launch, loads hoisted in front of the wait): the same
batch is generated N times; every run's codes must equal the first run's.
    python tools/soak_determinism.py [runs]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
import bench  # noqa: E402
from vall_e.vall_e.diffusion import Diffusion  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.manual_seed(0)
m = Diffusion(**bench.MODEL, n_steps=51, transition="absorbing")
for blk in m.blocks:
    for sub in (blk.attn, blk.ffn):
        torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
m = m.to(dev)
for name, shapes in (("one utterance (C2)", [(50, 225, 750)]), ("ragged batch of 3", [(50, 225, 750), (7, 40, 130), (30, 100, 333)]),
                     ("two utterances", [(50, 225, 750), (50, 225, 750)])):
    utts = [bench.synth_utterance(i, a, b) for i, (a, b, _) in enumerate(shapes)]
    text, proms = [u[0].to(dev) for u in utts], [u[1].to(dev) for u in utts]
    lens = [c for _, _, c in shapes]
    ref = torch.cat(m.generate_audio(text, proms, resp_lens=lens, seed=5))
    bad = 0
    for r in range(runs):
        out = torch.cat(m.generate_audio(text, proms, resp_lens=lens, seed=5))
        bad += int(not torch.equal(out, ref))
    torch.cuda.synchronize()
    print(f"{name}: {runs} runs x 50 denoise steps, {bad} differ from the first", flush=True)
