"""Benchmark of the D3PM denoising-sampler hot path (BASELINE.json metric: codec tokens/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5]
                    [--denoise-steps S] [--transition absorbing|uniform]

Workload (default c3 = BASELINE.json configs[2], the configuration the throughput metric is
quoted on; it fits one GPU): full denoiser (d=1024, 12 layers, 16 heads, K=1024, 8 levels), 256
utterances of 10 s (750 frames) + 3 s prompt (225 frames) + 50 phones, 50 denoise steps
(timesteps S=51, reverse loop t=50..1), absorbing transition, synthetic tokens, random-init
weights.  The 256 utterances are sharded by utterance over the N ranks (strong scaling: total
work fixed), no collective inside the loop, one all-gather of the codes at the end.

One "step" = one full pass of the hot path over one batch: all 50 denoise steps for every
utterance of the batch -> B * 750 * 8 codec tokens.  `value` = tokens / s with inputs resident
in HBM; `e2e` = the same through Diffusion.generate_audio() from pinned host tensors with the
codes read back to the host.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "tts-with-diffusion-model_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    #        B    T_txt T_prom T_resp timesteps transition
    "c2": (1, 50, 225, 750, 51, "absorbing"),
    "c3": (256, 50, 225, 750, 51, "absorbing"),
    "c4": (16, 50, 225, 2250, 51, "absorbing"),
    "c5": (64, 50, 225, 300, 26, "uniform"),      # BASELINE configs[4], one point of its sweep: S = 25, uniform
}
MODEL = dict(n_tokens=1024, d_model=1024, n_heads=16, n_layers=12)
METRIC, UNIT = "codec_tokens_per_sec", "tokens/s"


def synth_batch(n, t_txt, t_prom, seed):
    g = torch.Generator().manual_seed(seed)
    text = [torch.randint(1, MODEL["n_tokens"], (t_txt,), generator=g) for _ in range(n)]
    proms = [torch.randint(0, MODEL["n_tokens"], (t_prom, 8), generator=g) for _ in range(n)]
    return text, proms


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("bf16_tflops_sustained", 1378.2), d.get("hbm_gbs", 6450.3), "measured"
    return 1400.0, 6650.0, "fallback"   # B200_PROFILING.md fallback (sustained GEMM, copy)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_step(state):
    """One denoise step of ONE C2-shaped utterance with the oracle port of the reference
    (fp32 denoiser forward of base.py + dense fp16 p_sample of ar_discrete.py), all host threads."""
    from oracle import denoiser as on
    sd, orc, text, proms, x_t, t = state["sd"], state["orc"], state["text"], state["proms"], state["x_t"], state["t"]
    tt = torch.tensor([t])
    logits = on.diffusion_logits(sd, text, proms, [x_t], tt, MODEL["n_heads"], MODEL["n_layers"])[0]
    lg = logits.to(torch.float16)                                   # (T_r, 8, K): reference p_sample needs fp16
    noise = torch.rand(lg.shape)
    tok_t = torch.full((lg.shape[0],), t)
    samp, _ = orc.p_sample(lg, tok_t, x_t.to(torch.int32), noise)
    state["x_t"] = samp
    state["t"] = max(t - 1, 1)


def cpu_reference_setup(t_txt, t_prom, t_resp, timesteps, transition):
    from oracle import denoiser as on
    from oracle.d3pm import D3PM
    torch.set_num_threads(os.cpu_count() or 1)
    K = MODEL["n_tokens"]
    sd = on.random_state_dict(K, MODEL["d_model"], MODEL["n_layers"], timesteps + 1, n_resp_levels=8,
                              n_out=8 * K, seed=0, time_rows=timesteps + 1, bf16_round=False)
    text, proms = synth_batch(1, t_txt, t_prom, seed=1)
    x_t = torch.full((t_resp, 8), K // 2, dtype=torch.long)
    return dict(sd=sd, orc=D3PM(timesteps, K, transition), text=text, proms=proms, x_t=x_t, t=timesteps - 1)


def run_reference(args, wl):
    B, t_txt, t_prom, t_resp, timesteps, transition = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    state = cpu_reference_setup(t_txt, t_prom, t_resp, timesteps, transition)
    for _ in range(args.warmup):
        cpu_reference_step(state)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(state)
    dt = (time.perf_counter() - t0) / args.steps
    n_den = timesteps - 1
    value = t_resp * 8 / (dt * n_den)          # tokens of one utterance / time of its full reverse loop
    cores = torch.get_num_threads()
    sample = (f"1 utterance of the {wl} shape (T={t_txt + t_prom + t_resp + 2}), each step = 1 of its {n_den} denoise "
              f"steps (fp32 denoiser forward + dense fp16 p_sample); tokens/s = {t_resp * 8} / ({n_den} x step time)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(wl, n_gpus):
    B, t_txt, t_prom, t_resp, timesteps, transition = WORKLOADS[wl]
    return {"workload": f"{wl}: full D3PM denoiser d=1024 L=12 h=16 K=1024 x 8 levels, {B} utterances x "
                        f"({t_txt} phones + {t_prom} prompt frames + {t_resp} frames), {timesteps - 1} denoise steps, "
                        f"{transition}; sharded by utterance over {n_gpus} GPU(s)",
            "global_batch": B, "frames": t_resp, "denoise_steps": timesteps - 1, "transition": transition,
            "l2_policy": "per-step working set (activations > 1 GB per GPU at every N) exceeds the 126 MB L2; no flush needed",
            "operands": "16-bit tensor-core operands, fp32 accumulate: bf16 everywhere except the classifier GEMM "
                        "(fp16 rows and weights); fp32 residual stream",
            "parallelism": f"utterance-sharded x{n_gpus}, one final all-gather"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args, wl):
    import torch.distributed as dist
    from vall_e.b200 import lib as L
    from vall_e.b200.shard import partition, utterance_cost
    from vall_e.vall_e.diffusion import Diffusion

    B, t_txt, t_prom, t_resp, timesteps, transition = WORKLOADS[wl]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()

    torch.manual_seed(0)
    model = Diffusion(**MODEL, n_steps=timesteps, transition=transition)
    for blk in model.blocks:                      # AdaLN tables are zero-initialised; exercise the path
        for sub in (blk.attn, blk.ffn):
            torch.nn.init.normal_(sub.norm.emb.weight, std=0.02)
    model = model.to(dev)
    eng = model.engine()

    costs = [utterance_cost(t_txt, t_prom, t_resp)] * B
    mine = partition(costs, world)[rank]
    n_local = len(mine)
    resp_lens = [t_resp] * n_local
    n_total = args.warmup + args.steps
    # host batches in pinned memory (e2e arm) — a different synthetic batch every step
    batches = []
    for s in range(n_total):
        text, proms = synth_batch(n_local, t_txt, t_prom, seed=1000 * (1 + rank) + s)
        batches.append(([t.pin_memory() for t in text], [p.pin_memory() for p in proms],
                        [s * B + i for i in mine]))

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_codes(x_t):
        """the one collective of the path: all-gather of the generated codes (int16 bytes)"""
        if world == 1:
            return x_t
        send = x_t.to(torch.int16).view(torch.uint8).view(-1)
        n_max = (B + world - 1) // world * t_resp * 8 * 2
        buf = torch.zeros(n_max, dtype=torch.uint8, device=dev)
        buf[: send.numel()] = send
        out = torch.empty(world * n_max, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, buf)
        return out

    # ---------------- device-resident arm: inputs already in HBM when the clock starts
    ses = model._session(batches[0][0], batches[0][1], resp_lens, batches[0][2])
    table = model._table(dev)
    tr = L.ABSORBING if transition == "absorbing" else L.UNIFORM

    def device_step(i):
        ses.load([t.to(dev) for t in batches[i][0]], [p.to(dev) for p in batches[i][1]], None)   # untimed staging happens before
        ses.x_t.fill_(model.mask_id)

    def timed_device_step():
        ses.run(table, timesteps, tr, noise=L.NOISE_PHILOX, seed=args.seed, use_graph=True)
        gather_codes(ses.x_t)

    for i in range(args.warmup):
        device_step(i)
        timed_device_step()
    sync_all()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        sync_all()
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            device_step(args.warmup + k)      # restage inputs on device (outside the event bracket)
            ev[k][0].record()
            timed_device_step()
            ev[k][1].record()
        sync_all()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms = max_over_ranks(sum(step_ms) / len(step_ms))
    launches = (eng.launches - launches0)
    tokens_per_step = B * t_resp * 8
    value = tokens_per_step / (ms / 1e3)

    # ---------------- end-to-end arm: public API, pinned host inputs -> host codes, every step
    for i in range(min(1, args.warmup)):
        model.generate_audio(batches[i][0], batches[i][1], resp_lens=resp_lens, seed=args.seed, gids=batches[i][2], to_host=True)
    sync_all()
    e2e_ms = []
    for k in range(args.steps):
        bt, bp, gid = batches[args.warmup + k]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        a.record()
        codes = model.generate_audio(bt, bp, resp_lens=resp_lens, seed=args.seed, gids=gid, to_host=True)
        if world > 1:
            gather_codes(ses.x_t)
        b.record()
        torch.cuda.synchronize(dev)
        e2e_ms.append(a.elapsed_time(b))
    e2e = max_over_ranks(sum(e2e_ms) / len(e2e_ms))
    h2d = int(model.last_h2d_bytes)
    d2h = int(n_local * t_resp * 8 * 4)

    # ---------------- roofline of the dominant kernel (tcgen05 GEMM), measured live: one eager
    # pass of the same batch with CUDA events around every GEMM / attention launch
    eng.profile = []
    ses.x_t.fill_(model.mask_id)
    prof_steps = 2
    ses.t_utt.fill_(timesteps - 1)
    K_cls = eng.w.n_out // 8
    for _ in range(prof_steps):
        head_in = eng.forward(ses.lay, ses.ws, ses.x_t, ses.t_utt, use_time=True, head=False)
        ev = eng._prof_begin()
        L.head_posterior_sample(ses.x_t, ses.ws.logits, head_in, eng.w.w_cls, eng.w.b_cls, ses.x_t,
                                ses.lay.resp_row_utt, ses.t_utt, ses.lay.utt, table, 8, K_cls, tr, L.NOISE_PHILOX,
                                None, args.seed)
        eng._prof_end(ev, "head_sample", 2 * ses.lay.M_resp * eng.w.n_out * eng.w.d)
    torch.cuda.synchronize(dev)
    prof, eng.profile = eng.profile, None
    agg = {}
    for kind, flops, s, e in prof:
        a = agg.setdefault(kind, [0.0, 0.0, 0])
        a[0] += flops
        a[1] += s.elapsed_time(e)
        a[2] += 1
    peak_tf, peak_hbm, peak_src = measured_peaks()
    g = list(agg.get("gemm", [0.0, 1.0, 1]))
    hsamp = agg.get("head_sample")
    if hsamp:                 # the classifier GEMM (with the reverse step as its epilogue) is a tcgen05 GEMM too
        g = [g[0] + hsamp[0], g[1] + hsamp[1], g[2] + hsamp[2]]
    at = agg.get("attn", [0.0, 1.0, 1])
    gemm_tf = g[0] / (g[1] * 1e-3) / 1e12
    attn_tf = at[0] / (at[1] * 1e-3) / 1e12
    step_total_ms = ms / (timesteps - 1)
    traffic = None      # dram bytes per launch of the dominant GEMM from one committed `ncu --set full` capture
    tf = ROOT / "profiles" / "r1_gemm_traffic.json"
    if tf.exists():
        tj = json.loads(tf.read_text())
        traffic = tj.get("bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (QKV/out/FFN1/FFN2) + head_sample_kernel (classifier)",
                "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tf / peak_tf,
                "peak_source": f"{peak_src} bf16_tflops_sustained",
                "peak_note": "cuBLAS bf16 back to back for 4 s at the pool's power cap; the launches here are timed one by "
                             "one inside an eager pass of the step, so a lightly loaded box can exceed it (burst figure 1632)",
                "traffic": traffic,
                "traffic_note": "dram bytes of one FFN1 launch (largest GEMM) from profiles/r1_gemm_traffic.json; algorithmic 2.70 GB",
                "avg_launch_ms": g[1] / g[2], "launches_timed": g[2],
                "share_of_denoise_step": (g[1] / prof_steps) / step_total_ms,
                "attention": {"kernel": "flash_attn_kernel", "achieved": attn_tf, "unit": "TFLOP/s",
                              "frac": attn_tf / peak_tf, "avg_launch_ms": at[1] / at[2],
                              "share_of_denoise_step": (at[1] / prof_steps) / step_total_ms}}
    # the HBM-bound kernels of the step beside it (algorithmic bytes / launch time / measured copy bandwidth)
    if hsamp:
        roofline["head_sample"] = {"kernel": "head_sample_kernel (classifier GEMM + D3PM reverse step, one launch)",
                                   "bound": "tensor", "achieved": hsamp[0] / (hsamp[1] * 1e-3) / 1e12,
                                   "unit": "TFLOP/s", "frac": hsamp[0] / (hsamp[1] * 1e-3) / 1e12 / peak_tf,
                                   "avg_launch_ms": hsamp[1] / hsamp[2],
                                   "share_of_denoise_step": (hsamp[1] / prof_steps) / step_total_ms}
    for kind, name in (("norm", "adaln_rows_kernel"),):
        if kind in agg:
            byt, t_ms, n = agg[kind]
            gbs = byt / (t_ms * 1e-3) / 1e9
            roofline[kind] = {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s",
                              "frac": gbs / peak_hbm, "avg_launch_ms": t_ms / n,
                              "share_of_denoise_step": (t_ms / prof_steps) / step_total_ms}

    # ---------------- p50 denoise-step latency (BASELINE.json's second metric): C2 shape, batch 1,
    # one CUDA-graph replay = denoiser forward + posterior + sample, >= 20 warm iterations
    latency = None
    if rank == 0:
        b1, tt1, tp1, tr1, _, _ = WORKLOADS["c2"]
        text1, proms1 = synth_batch(1, tt1, tp1, seed=7)
        ses1 = model._session([t.to(dev) for t in text1], [p.to(dev) for p in proms1], [tr1], [0])
        ses1.x_t.fill_(model.mask_id)
        ses1.run(table, timesteps, tr, noise=L.NOISE_PHILOX, seed=args.seed, use_graph=True)   # captures the step graph
        lat = []
        for _ in range(40):
            ses1.t_utt.fill_(timesteps // 2)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ses1.graph.replay()
            b.record()
            torch.cuda.synchronize(dev)
            lat.append(a.elapsed_time(b))
        lat = sorted(lat[10:])
        latency = {"p50_denoise_step_ms": lat[len(lat) // 2], "p90_denoise_step_ms": lat[int(len(lat) * 0.9)],
                   "workload": f"c2: batch 1, T={tt1 + tp1 + tr1 + 2} rows ({tr1} frames x 8 levels), one graph replay per step",
                   "tokens_per_sec_batch1": tr1 * 8 / (lat[len(lat) // 2] * 1e-3 * (timesteps - 1))}

    # ---------------- CPU baseline beside it (rank 0, N=1 only): bounded sample of the oracle port
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        state = cpu_reference_setup(t_txt, t_prom, t_resp, timesteps, transition)
        cpu_reference_step(state)
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            cpu_reference_step(state)
        dt = (time.perf_counter() - t0) / n
        cpu_baseline = {"value": t_resp * 8 / (dt * (timesteps - 1)), "unit": UNIT, "cores": torch.get_num_threads(),
                        "kind": "port",
                        "sample": f"oracle port of the reference, 1 utterance of this shape, {n} of {timesteps - 1} denoise "
                                  f"steps timed ({dt:.2f} s each), tokens/s = {t_resp * 8} / ({timesteps - 1} x step)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(wl, world),
                "denoise_step_ms": step_total_ms, "wall_s_timed_region": t_wall,
                "clocks": clocks.summary(),
                "e2e": {"value": tokens_per_step / (e2e / 1e3), "unit": UNIT, "ms_per_step": e2e,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "roofline": roofline, "latency": latency, "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--denoise-steps", type=int, default=None,
                    help="override the workload's number of denoise steps (BASELINE configs[4] sweeps 10/25/50/100)")
    ap.add_argument("--transition", default=None, choices=["absorbing", "uniform"],
                    help="override the workload's transition (BASELINE configs[4] sweeps both)")
    args = ap.parse_args()
    if args.denoise_steps is not None or args.transition is not None:
        B, t_txt, t_prom, t_resp, timesteps, transition = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (B, t_txt, t_prom, t_resp,
                                    timesteps if args.denoise_steps is None else args.denoise_steps + 1,
                                    transition if args.transition is None else args.transition)
    if args.impl == "reference":
        run_reference(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
