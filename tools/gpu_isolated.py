#!/usr/bin/env python
"""Run each GPU test node in its own process with a timeout (a trapped kernel kills the CUDA context,
so isolation keeps one bad kernel from masking the rest).  Writes gpurun_out/isolated_<tag>.txt."""
import os
import subprocess
import sys
import time

def main():
    args = sys.argv[1:]
    tag = "run"
    if args and args[0].startswith("--tag="):
        tag = args.pop(0).split("=", 1)[1]
    targets = args or ["tests"]
    os.makedirs("gpurun_out", exist_ok=True)
    col = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *targets],
                         capture_output=True, text=True)
    nodes = [l.strip() for l in col.stdout.splitlines() if "::" in l]
    lines = []
    t00 = time.time()
    for n in nodes:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", n, "-p", "no:cacheprovider"],
                               capture_output=True, text=True, timeout=int(os.environ.get("FT_TEST_TIMEOUT", "300")))
            ok = r.returncode == 0
            tail = "" if ok else "\n".join((r.stdout + r.stderr).splitlines()[-25:])
        except subprocess.TimeoutExpired:
            ok, tail = False, "TIMEOUT"
        lines.append(f"{'PASS' if ok else 'FAIL'} {time.time() - t0:6.1f}s {n}")
        if tail:
            lines.append(tail)
        print(lines[-1 if not tail else -2], flush=True)
        if tail:
            print(tail, flush=True)
    npass = sum(l.startswith("PASS") for l in lines)
    nfail = sum(l.startswith("FAIL") for l in lines)
    lines.append(f"TOTAL pass={npass} fail={nfail} wall={time.time() - t00:.0f}s")
    print(lines[-1])
    with open(f"gpurun_out/isolated_{tag}.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    return 0

if __name__ == "__main__":
    sys.exit(main())
