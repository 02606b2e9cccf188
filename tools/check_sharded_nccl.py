"""Multi-GPU functional check (NCCL): `torchrun --nproc-per-node N tools/check_sharded_nccl.py`.
Every rank generates its shard of a ragged batch, one all-gather assembles the codes, and rank 0
compares them bit for bit with a single-process run of the whole batch (codes must not depend on
the number of ranks: Philox is keyed by the global utterance id)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200.shard import generate_sharded  # noqa: E402
from vall_e.vall_e.diffusion import Diffusion  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
K, d, S = 256, 256, 9
m = Diffusion(K, d_model=d, n_heads=4, n_layers=3, n_steps=S, transition="absorbing")
for blk in m.blocks:
    for sub in (blk.attn, blk.ffn):
        torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
m = m.to(dev)
g = torch.Generator().manual_seed(3)
lens = [(5, 9, 140), (7, 30, 300), (2, 1, 17), (11, 50, 260), (3, 8, 64), (9, 12, 500), (4, 4, 33)]
text = [torch.randint(1, K, (a,), generator=g).to(dev) for a, _, _ in lens]
proms = [torch.randint(0, K, (b, 8), generator=g).to(dev) for _, b, _ in lens]
resp = [c for _, _, c in lens]


def gen(t, p, r, gids):
    return m.generate_audio(t, p, resp_lens=r, seed=7, gids=gids)


codes = generate_sharded(gen, text, proms, resp, device=dev, d_model=d)
torch.cuda.synchronize()
if rank == 0:
    whole = m.generate_audio(text, proms, resp_lens=resp, seed=7, gids=list(range(len(lens))))
    ok = all(torch.equal(a, b) for a, b in zip(codes, whole))
    print(f"sharded over {world} ranks == single process: {ok}")
    assert ok
dist.barrier()
dist.destroy_process_group()
