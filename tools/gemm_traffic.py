"""Regenerates profiles/gemm_traffic.json — the `roofline.traffic` of bench.py — from a fresh `ncu --set full`
capture of ONE FFN1 launch at the bench shape (M = 262 912, N = 4 096, K = 1 024, GELU epilogue), and stamps it with
the digest of the GEMM sources so that bench.py only reports it for the build it was taken from.  On the GPU box:
    python tools/gemm_traffic.py            (writes gpurun_out/r2_gemm_ffn1.ncu-rep, profiles/gemm_traffic.json and
                                             profiles/r2_gemm_ffn1_ncu_raw.txt)"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rep = ROOT / "gpurun_out" / "r2_gemm_ffn1.ncu-rep"
rep.parent.mkdir(exist_ok=True)
subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:gemm_tcgen05",
                "-s", "2", "-c", "1", "-f", "-o", str(rep.with_suffix("")), sys.executable, str(ROOT / "tools" / "kernel_probe.py"),
                "gemm_big"], check=True, capture_output=True, text=True)
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
u = dict(zip(hdr, units))


def num(k):
    v = float(d[k].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u[k], 1)
    return v * scale


rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
M, N, K = 262912, 4096, 1024
algo = M * K * 2 + N * K * 2 + M * N * 2 + N * 4
import bench  # noqa: E402
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "launch__grid_size", "launch__registers_per_thread")
txt = [f"# ncu --set full --clock-control none, one FFN1 launch (gemm_tcgen05_kernel<BIAS_GELU, bf16, CTAS=2, 256>), M={M} N={N} K={K}",
       f"# Kernel: {d.get('Kernel Name')}"]
txt += [f"{k} = {d[k]} {u[k]}" for k in keep if k in d]
for dst in ("profiles", "gpurun_out"):          # gpurun only brings gpurun_out/ back: copy from there into profiles/
    (ROOT / dst / "r2_gemm_ffn1_ncu_raw.txt").write_text("\n".join(txt) + "\n")
out = {"bytes_per_launch": int(rd + wr), "algorithmic_bytes": algo, "dram_read": int(rd), "dram_write": int(wr),
       "kernel": f"gemm_tcgen05_kernel (FFN1), M={M} N={N} K={K}", "capture": "profiles/r2_gemm_ffn1_ncu_raw.txt (tools/gemm_traffic.py)",
       "kernel_source_digest": bench.kernel_source_digest("gemm_tcgen05.cu", "common.cuh"),
       "tensor_pipe_active_pct": float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan").replace(",", ""))}
for dst in ("profiles", "gpurun_out"):
    (ROOT / dst / "gemm_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))
