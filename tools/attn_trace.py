"""clock64() timeline of the attention kernel's softmax warp 0 and MMA warp (debug build -DVB200_ATTN_TRACE,
built by this script's caller into build/variants/libtrace.so).  Prints per-phase cycle statistics and the
phase relation of the two CTAs that share an SM.   VB200_LIB=.../libtrace.so python tools/attn_trace.py [B] [T]"""
import ctypes as C
import os
import sys
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tts-with-diffusion-model_b200"))
from vall_e.b200 import lib as L  # noqa: E402

lib = L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
heads, dev = 16, "cuda"
torch.manual_seed(1)
M, d = B * T, heads * 64
qkv = torch.randn(M, 3 * d, device=dev).bfloat16()
cu = torch.arange(B + 1, dtype=torch.int32, device=dev) * T
out = torch.empty(M, d, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    L.flash_attn_varlen(out, qkv, cu, T, heads, 0.125)
torch.cuda.synchronize()
CTAS, EV = 1024, 8
try:
    BLK = int(lib.vb200_debug_attn_trace_blocks())      # persistent kernel: blocks numbered over all of a CTA's items
    PERSISTENT = True
except AttributeError:
    BLK, PERSISTENT = 16, False
buf = np.zeros(CTAS * 2 * BLK * EV, dtype=np.uint64)
fn = lib.vb200_debug_attn_trace
fn.argtypes = [C.c_void_p, C.c_size_t]
assert fn(buf.ctypes.data, buf.size) == 0
tr = buf.reshape(CTAS, 2, BLK, EV).astype(np.int64)
nblk = (T + 127) // 128
if PERSISTENT:
    # one stream of blocks per CTA; report steady-state blocks (not the first / last of an item) and item boundaries
    names = ["wait s_full", "LDTM", "max+shfl", "exp pass", "wait pv_done", "STTM+arrive", "to next block"]
    inner, edge = {k: [] for k in names}, {k: [] for k in names}
    per_inner, per_edge = [], []
    for c in range(CTAS):
        sm = tr[c, 0]
        if sm[0, 1] == 0:
            continue
        for gidx in range(1, BLK - 2):
            e, nx = sm[gidx], sm[gidx + 1]
            if e[6] == 0 or nx[0] == 0:
                break
            first_of_item = gidx % nblk == 0
            dst, per = (edge, per_edge) if first_of_item else (inner, per_inner)
            for k in range(6):
                dst[names[k]].append(e[k + 1] - e[k])
            dst[names[6]].append(nx[0] - e[6])
            per.append(nx[0] - e[0])
    for label, dd, per in (("inner blocks", inner, per_inner), ("first block of an item", edge, per_edge)):
        if not per:
            continue
        print(f"B={B} T={T} persistent, {label}: cycles per key block median {np.median(per):.0f} mean {np.mean(per):.0f}")
        for k in names:
            v = np.array(dd[k])
            print(f"  {k:14s} median {np.median(v):7.0f}  mean {v.mean():7.0f}  p90 {np.percentile(v, 90):7.0f}")
    # MMA warp: P V(g) issue relative to the softmax warp's p_full arrival, S(g+1) issue relative to its LDTM
    dpv, ds = [], []
    for c in range(CTAS):
        sm, mm = tr[c, 0], tr[c, 1]
        for gidx in range(1, BLK - 2):
            if sm[gidx, 6] and mm[gidx, 1]:
                dpv.append(mm[gidx, 1] - sm[gidx, 6])
            if sm[gidx, 2] and mm[gidx + 1, 0]:
                ds.append(mm[gidx + 1, 0] - sm[gidx, 2])
    print(f"  MMA warp: P V(g) issued {np.median(dpv):.0f} cycles after warp 0 arrived on p_full (p90 {np.percentile(dpv, 90):.0f}); "
          f"S(g+1) issued {np.median(ds):.0f} after warp 0 had its scores (p90 {np.percentile(ds, 90):.0f})")
    sys.exit(0)
names = ["wait s_full", "LDTM", "max+shfl", "exp pass", "wait pv_done", "STTM+arrive", "to next block"]
ph = defaultdict(list)
per_block = []
for c in range(CTAS):
    sm = tr[c, 0]
    if sm[0, 1] == 0:
        continue
    for j in range(1, nblk - 1):                      # steady state blocks
        e = sm[j]
        for k in range(6):
            ph[names[k]].append(e[k + 1] - e[k])
        ph[names[6]].append(sm[j + 1][0] - e[6])
        per_block.append(sm[j + 1][0] - e[0])
print(f"B={B} T={T}: steady-state cycles per key block of one CTA: median {np.median(per_block):.0f}, mean {np.mean(per_block):.0f}")
for k in names:
    v = np.array(ph[k])
    print(f"  {k:14s} median {np.median(v):7.0f}  mean {v.mean():7.0f}  p90 {np.percentile(v, 90):7.0f}")
# MMA warp: gap between S(j+1) issue and P V(j) issue, and CTA lifetime
life, first = [], []
for c in range(CTAS):
    mm, sm = tr[c, 1], tr[c, 0]
    if sm[0, 1] == 0:
        continue
    t_start = mm[BLK - 1, 6]
    life.append(sm[nblk - 1, 6] - t_start)
    first.append(sm[0, 1] - t_start)
print(f"  CTA start -> first scores ready: median {np.median(first):.0f}; start -> last P written: median {np.median(life):.0f}")
# phase relation of co-resident CTAs: for CTAs on the same SM whose lifetimes overlap, offset of their exp-pass starts
by_sm = defaultdict(list)
for c in range(CTAS):
    if tr[c, 0, 0, 1]:
        by_sm[int(tr[c, 1, BLK - 1, 7])].append(c)
offs = []
for sm_id, cs in by_sm.items():
    cs.sort(key=lambda c: tr[c, 1, BLK - 1, 6])
    for a, b in zip(cs, cs[1:]):
        ea = tr[a, 0, 1:nblk - 1, 3]                 # exp-pass start stamps
        eb = tr[b, 0, 1:nblk - 1, 3]
        period = np.median(np.diff(ea)) if len(ea) > 1 else 0
        for t in eb:
            if ea[0] <= t <= ea[-1] and period > 0:
                k = np.searchsorted(ea, t) - 1
                offs.append((t - ea[k]) / period)
offs = np.array(offs)
if len(offs):
    hist, _ = np.histogram(offs, bins=10, range=(0, 1))
    print("  exp-pass start of the co-resident CTA, as a fraction of this CTA's block period (0 = in phase, 0.5 = anti-phase):")
    print("   ", " ".join(f"{h:5d}" for h in hist))
