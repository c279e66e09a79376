#!/usr/bin/env python
"""JOB suite benchmark (BASELINE.json config 5, 'JOB-plan total ms vs CPU'):
all 113 plans of plans.json on synthetic IMDB-shaped data, Contest::execute timed per query exactly
like the contest harness does (host pages in, host pages out; tests/read_sql.cpp:1234-1236), on the
B200 engine and -- on a subset sized to finish in minutes -- on the unmodified reference's CPU
execute() (oracle/_ref) for the same inputs, with a parity check of every compared result.

    python bench_job.py [--scale 0.05] [--cpu-queries 20] [--names 1a,13a,...]
Prints one JSON line.  Not the driver's headline (that is bench.py on config 2).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--cpu-queries", type=int, default=20, help="how many plans the CPU reference also runs")
    ap.add_argument("--names", default="")
    ap.add_argument("--repeat", type=int, default=2, help="GPU executes per plan (best is kept)")
    args = ap.parse_args()
    import radix_join_b200 as rj
    from radix_join_b200 import job
    from oracle import pyoracle as orc

    names = [n for n in args.names.split(",") if n] or list(job.workload()["queries"])
    kind = "reference" if orc.available("ref") else "port"
    ctx = rj.build_context(0)
    per_query, gpu_total, scan_total, out_total = {}, 0.0, 0, 0
    cpu_total = cpu_gpu_total = cpu_rows = 0.0
    cpu_left = args.cpu_queries
    t_gen = 0.0
    for name in names:
        t0 = time.perf_counter()
        plan, root_cols, scan_rows = job.make_job(name, scale=args.scale, seed=1)
        t_gen += time.perf_counter() - t0
        best = None
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            got = rj.execute(plan, ctx)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        entry = {"scan_rows": scan_rows, "out_rows": got.num_rows, "gpu_ms": round(best * 1e3, 3)}
        gpu_total += best
        scan_total += scan_rows
        out_total += got.num_rows
        # CPU arm on the first --cpu-queries plans that are not tiny, with a parity check
        if cpu_left > 0:
            t0 = time.perf_counter()
            want = orc.execute(plan, impl="ref" if kind == "reference" else "port")
            cpu = orc.last_execute_seconds() if kind == "reference" else time.perf_counter() - t0
            assert got.num_rows == want.num_rows and orc.result_equal(got, want), f"parity failure on {name}"
            entry["cpu_ms"] = round(cpu * 1e3, 1)
            cpu_total += cpu
            cpu_gpu_total += best
            cpu_rows += scan_rows
            cpu_left -= 1
        per_query[name] = entry
    rj.destroy_context(ctx)
    line = {
        "metric": "job_suite_total_ms", "unit": "ms", "higher_is_better": False,
        "value": round(gpu_total * 1e3, 2), "n_gpus": 1, "plans": len(names), "scale": args.scale,
        "scan_rows": scan_total, "output_rows": out_total, "mtuples_per_s": round(scan_total / 1e6 / gpu_total, 2),
        "timing": "host clock around execute(): host pages in, host pages out (H2D + D2H inside), best of %d" % args.repeat,
        "cpu_baseline": {"kind": kind, "cores": os.cpu_count(), "plans": args.cpu_queries - cpu_left,
                         "total_ms": round(cpu_total * 1e3, 1), "gpu_ms_same_plans": round(cpu_gpu_total * 1e3, 2),
                         "speedup_same_plans": round(cpu_total / cpu_gpu_total, 1) if cpu_gpu_total else None,
                         "parity": "every compared result bit-exact (multiset) vs the CPU implementation"},
        "data": "synthetic IMDB-shaped (radix_join_b200.job.make_inputs), generation %.1f s not timed" % t_gen,
        "per_query": per_query,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
