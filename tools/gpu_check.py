#!/usr/bin/env python
"""Run every GPU test in its own process with a timeout, so one trapping kernel cannot hide the rest.
Usage (on the GPU box):  python tools/gpu_check.py [-k expr] [--timeout 120]"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-k", default=None)
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("paths", nargs="*", default=["tests"])
    args = ap.parse_args()
    cmd = [sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu"] + args.paths
    if args.k:
        cmd += ["-k", args.k]
    ids = [l.strip() for l in subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True).stdout.splitlines()
           if "::" in l]
    print(f"{len(ids)} gpu tests", flush=True)
    failed = []
    for tid in ids:
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "--no-header", "-m", "gpu", tid], cwd=ROOT,
                               capture_output=True, text=True, timeout=args.timeout)
            ok = r.returncode == 0
            tail = "" if ok else "\n".join((r.stdout + r.stderr).splitlines()[-25:])
        except subprocess.TimeoutExpired:
            ok, tail = False, "TIMEOUT"
        print(("PASS " if ok else "FAIL ") + tid, flush=True)
        if not ok:
            failed.append(tid)
            print(tail, flush=True)
    print(f"\n{len(ids) - len(failed)}/{len(ids)} passed")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
