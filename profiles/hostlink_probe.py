#!/usr/bin/env python
"""Host-link probe: what the box's host<->device path carries, without the library.

    python profiles/hostlink_probe.py [--out gpurun_out/r02_hostlink_probe]

Runs structured_light_calculation_b200/bin/hostlink_probe (profiles/hostlink_probe.cu) twice:
  1. one process, a thread per GPU: H2D / D2H / both, every GPU alone, pairs, quads, all; pinned,
     write-combined and 2 MB-page host memory; 64 MB and 4 MB copies; CPU memcpy bandwidth;
  2. one process per GPU, started at a common wall-clock time (what `torchrun bench.py` does).
Writes <out>.jsonl (every measurement) and <out>.txt (the table DESIGN.md quotes), and
<out>_ceiling.json: the per-N ceilings bench.py reads for `e2e.roofline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "structured_light_calculation_b200", "bin", "hostlink_probe")


def run_lines(cmd):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if res.returncode != 0:
        raise SystemExit(f"{' '.join(cmd)} failed ({res.returncode}): {res.stderr[-2000:]}")
    return [json.loads(ln) for ln in res.stdout.splitlines() if ln.startswith("{")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_hostlink_probe"))
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--reps", type=int, default=24)
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)

    rows = run_lines([EXE, "--plan", "full", "--mb", str(args.mb), "--reps", str(args.reps)])
    n_dev = rows[0]["devices"]

    # one process per GPU, N = 1, 2, 4, 8 together
    for n in [k for k in (1, 2, 4, 8) if k <= n_dev]:
        start = time.time() + 6.0 + 1.0 * n          # context creation of n processes
        procs = [subprocess.Popen([EXE, "--plan", "process", "--devices", str(d), "--start-at", f"{start:.3f}",
                                   "--spacing", "1.0", "--mb", str(args.mb), "--reps", str(args.reps)],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for d in range(n)]
        per_dir = {}
        for p in procs:
            out, err = p.communicate()
            if p.returncode != 0:
                raise SystemExit(f"process-mode probe failed: {err[-1000:]}")
            for ln in out.splitlines():
                if ln.startswith("{"):
                    r = json.loads(ln)
                    if r.get("kind") == "dma":
                        per_dir.setdefault(r["dir"], []).append(r)
        for d, lst in per_dir.items():
            rows.append({"kind": "dma", "mode": "processes", "gpus": [r["gpus"][0] for r in lst], "n": len(lst), "dir": d,
                         "mem": "pinned", "mb_per_copy": args.mb, "copies_per_dir": args.reps,
                         "per_gpu_gbs": [r["per_gpu_gbs"][0] for r in lst],
                         # the processes start together (wall clock) and run the same bytes: sum of their rates
                         "aggregate_gbs": round(sum(r["per_gpu_gbs"][0] for r in lst), 2),
                         "seconds": max(r["seconds"] for r in lst)})

    with open(args.out + ".jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")

    dma = [r for r in rows if r.get("kind") == "dma"]
    lines = [json.dumps(rows[0]), "",
             f"{'mode':10s} {'mem':20s} {'MB':>3s} {'gpus':22s} {'dir':6s} {'aggregate GB/s':>15s}   per GPU"]
    for r in dma:
        lines.append(f"{r['mode']:10s} {r['mem']:20s} {r['mb_per_copy']:3d} {str(r['gpus']):22s} {r['dir']:6s} "
                     f"{r['aggregate_gbs']:15.1f}   {r['per_gpu_gbs']}")
    lines.append("")
    for r in rows:
        if r.get("kind") == "cpu_memcpy":
            lines.append(f"cpu memcpy {r['threads']:3d} threads: {r['read_plus_write_gbs']:.1f} GB/s (read + write)")
    with open(args.out + ".txt", "w") as f:
        f.write("\n".join(lines) + "\n")

    # ceilings per N (GPUs 0..N-1 together, pinned memory): best of thread / process mode
    ceil = {}
    for n in (1, 2, 4, 8):
        if n > n_dev:
            continue
        want = list(range(n))
        entry = {}
        for d in ("h2d", "d2h", "bidir"):
            c = [r["aggregate_gbs"] for r in dma if r["gpus"] == want and r["dir"] == d and r["mem"] == "pinned"
                 and r["mb_per_copy"] == args.mb]
            if c:
                entry[d + "_gbs"] = max(c)
        ceil[str(n)] = entry
    with open(args.out + "_ceiling.json", "w") as f:
        json.dump({"how": "profiles/hostlink_probe.py: cudaMemcpyAsync of pinned 64 MB buffers, GPUs 0..N-1 together, "
                          "best of one-process-with-threads and one-process-per-GPU", "per_n": ceil}, f, indent=1)
    print("\n".join(lines))
    return 0


if __name__ == "__main__":
    sys.exit(main())
