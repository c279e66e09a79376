#!/usr/bin/env python
"""JOB suite benchmark (BASELINE.json config 5) -- thin CLI over radix_join_b200.job_bench, the same
code `bench.py --workload job` runs.

    python bench_job.py [--scale 0.05] [--cpu-queries 20] [--names 1a,13a,...]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--cpu-queries", type=int, default=20)
    ap.add_argument("--names", default="")
    ap.add_argument("--repeat", type=int, default=2)
    a = ap.parse_args()
    from radix_join_b200 import job_bench
    args = argparse.Namespace(job_scale=a.scale, job_cpu_plans=a.cpu_queries, job_names=a.names, steps=a.repeat, warmup=0)
    job_bench.run(args, impl="ours")


if __name__ == "__main__":
    main()
