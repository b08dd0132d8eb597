"""A/B of the DCCRN step and of its tensor-core GEMM launches alone between library builds (APS_B200_LIB)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or [""]
for rep in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["APS_B200_LIB"] = os.path.join(ROOT, lib)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "dccrn", "--steps", "10", "--warmup", "3",
                              "--cpu-seconds", "0.2"], capture_output=True, text=True, env=env, cwd=ROOT)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(f"{lib or 'default':40s} step {d['ms_per_step']:7.2f} ms   GEMM launches alone {d['roofline']['kernel_ms_per_step']:7.2f} ms", flush=True)
        except Exception as e:
            print(lib, "FAILED", e, out.stderr[-400:])
