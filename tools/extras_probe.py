"""In-graph cost of the once-per-denoise-step kernels of ONE utterance (embedding gather, response-row gather, fused
classifier + reverse step), each replayed back to back inside one CUDA graph:  python tools/extras_probe.py"""
import sys, torch
from pathlib import Path
ROOT = Path("/root/repo")
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
import bench
from vall_e.b200 import lib as L
from vall_e.vall_e.diffusion import Diffusion
dev = torch.device("cuda", 0); torch.cuda.set_device(0); torch.manual_seed(0)
m = Diffusion(**bench.MODEL, n_steps=51, transition="absorbing").to(dev)
u = bench.synth_utterance(7, 50, 225)
text, proms = [u[0].to(dev)], [u[1].to(dev)]
m.generate_audio(text, proms, resp_lens=[750], seed=3)
ses = m._session(text, proms, [750], [0]); eng = m.engine(); lay, ws, w = ses.lay, ses.ws, eng.w
table = m._table(dev)
def in_graph(fn, reps=48):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    ts = []
    for _ in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2] * 1e3 / reps
ses.t_utt.fill_(25)
import inspect
print("embed:", inspect.signature(L.embed_gather))
def emb():
    L.embed_gather(ws.x, w.text_w, w.prom_w, w.resp_w, w.sep, w.time_w, w.ensure_pe(lay.max_T), lay.text_ids, lay.prom_ids, ses.x_t, lay.utt, lay.row_utt, ses.t_utt, w.K, 8)
print(f"embed_gather: {in_graph(emb):.2f} us")
def gat():
    L.gather_rows_bf16(ws.head_in, ws.x, lay.resp_row_index)
print(f"gather_rows: {in_graph(gat):.2f} us")
K = w.n_out // 8
def head():
    L.head_posterior_sample(ses.x_t, ws.logits, ws.head_in, w.w_cls, w.b_cls, ses.x_t, lay.resp_row_utt, ses.t_utt, lay.utt, table, 8, K, L.ABSORBING, L.NOISE_PHILOX, None, 3)
print(f"head_sample: {in_graph(head, 24):.2f} us")
