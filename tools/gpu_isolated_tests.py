"""Bring-up helper: runs every GPU test in its own process (a device-side trap poisons the CUDA
context, so one pytest process would turn the first kernel bug into N failures).
Usage on the GPU box: python tools/gpu_isolated_tests.py [pytest -k expression] > gpurun_out/tests.log"""
import subprocess
import sys

sel = sys.argv[1] if len(sys.argv) > 1 else ""
cmd = [sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--collect-only", "-q"]
if sel:
    cmd += ["-k", sel]
ids = [l.strip() for l in subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines() if "::" in l]
ids = list(dict.fromkeys(i.split("[")[0] for i in ids))      # one process per test function
print(f"{len(ids)} tests", flush=True)
fails = 0
for tid in ids:
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", tid, "-q", "--no-header", "-p", "no:cacheprovider", "--tb=short"],
                           capture_output=True, text=True, timeout=400)
        ok = r.returncode == 0
        tail = "" if ok else "\n".join((r.stdout + r.stderr).splitlines()[-60:])
    except subprocess.TimeoutExpired:
        ok, tail = False, "TIMEOUT"
    fails += not ok
    print(("PASS " if ok else "FAIL ") + tid, flush=True)
    if tail:
        print(tail, flush=True)
print(f"failed {fails} of {len(ids)}")
