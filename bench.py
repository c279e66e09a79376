#!/usr/bin/env python
"""bench.py -- the reference's headline metric on a B200: join throughput (input tuples / s) of
Contest::execute on BASELINE.json config 2 (INT32 join, 64 Mi build x 512 Mi probe, Zipf(0.75),
INT64 + FP64 payloads with 1 % NULLs; synthetic, generated on device).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

One step = one full pass of the hot path: page decode -> radix partition -> shared-memory build+probe
-> gather + page encode, input pages resident in HBM when the timed region starts, result pages left in
HBM.  `e2e` is the same call with HOST pages in and HOST pages out (H2D / D2H inside the timed region).
`--impl reference` times the UNMODIFIED reference's CPU execute() (oracle/_ref) on a bounded sample of
the same workload on the host cores.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_BUILD, FULL_PROBE = 1 << 26, 1 << 29  # config 2
CPU_SAMPLE_DIV = 64                        # the CPU arms run config 2 at 1/64 scale (1 Mi x 8 Mi)
METRIC, UNIT = "join_throughput", "Mtuples/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (oracle/_ref) on a bounded sample of config 2
# --------------------------------------------------------------------------------------------------
def cpu_sample_plan(n_build, n_probe):
    """config 2 at reduced scale, host pages, same distributions (numpy twin of synthetic.make_c2_device)"""
    import numpy as np
    from oracle import pyoracle as orc
    from radix_join_b200 import synthetic as syn
    rng = np.random.default_rng(43)
    perm = rng.permutation(n_build).astype(np.int32)
    ra = syn.splitmix64_numpy(perm.astype(np.int64).view(np.uint64)).view(np.int64)
    zipf = syn.Zipf(n_build, 0.75)
    sk = perm[zipf.ranks(np.random.default_rng(44).random(n_probe), np)]
    bits = syn.splitmix64_numpy(np.arange(n_probe, dtype=np.uint64))
    inf = ((bits >> np.uint64(52)) & np.uint64(0x7FF)) == np.uint64(0x7FF)
    bits[inf] &= ~np.uint64(1 << 62)
    va = (rng.random(n_build) >= 0.01).astype(np.uint8)
    vb = (rng.random(n_probe) >= 0.01).astype(np.uint8)
    tl = orc.encode([orc.Cells.from_values(orc.INT32, perm), orc.Cells(orc.INT64, va, values=ra)], impl="port")
    tr = orc.encode([orc.Cells.from_values(orc.INT32, sk), orc.Cells(orc.FP64, vb, values=bits.view(np.float64))], impl="port")
    plan = syn.single_join_plan(payload=True)
    plan.new_input(tl)
    plan.new_input(tr)
    return plan


def time_reference(plan, n_tuples, steps, warmup):
    """seconds per execute() of the reference, timed like the contest harness does
    (steady_clock around Contest::execute only, tests/read_sql.cpp:1234-1236)"""
    from oracle import pyoracle as orc
    kind = "reference" if orc.available("ref") else "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        # all host threads: torchrun exports OMP_NUM_THREADS=1, so the count is set explicitly
        res = orc.execute(plan, impl="ref" if kind == "reference" else "port", n_threads=os.cpu_count() or 1)
        dt = orc.last_execute_seconds() if kind == "reference" else time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        rows = res.num_rows
        del res
    return kind, times, rows


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is a single-process CPU program: rank 0 alone runs it
    nb, np_ = FULL_BUILD // CPU_SAMPLE_DIV, FULL_PROBE // CPU_SAMPLE_DIV
    plan = cpu_sample_plan(nb, np_)
    kind, times, rows = time_reference(plan, nb + np_, args.steps, args.warmup)
    sec = sum(times) / len(times)
    value = (nb + np_) / 1e6 / sec
    cores = os.cpu_count()
    sample = f"config 2 at 1/{CPU_SAMPLE_DIV} scale: {nb} x {np_} rows, Zipf(0.75), INT64+FP64 payloads, {rows} output rows"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "c2_int32_join_64Mi_x_512Mi_zipf0.75_int64_fp64_payloads", "sample": sample},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_single_gpu(args):
    import torch
    import radix_join_b200 as rj
    from radix_join_b200 import synthetic as syn

    torch.cuda.set_device(0)
    nb, np_ = FULL_BUILD // args.scale, FULL_PROBE // args.scale
    ctx = rj.build_context(0)
    dt = syn.make_c2_device(ctx, nb, np_, checksum=True)
    inputs = rj.adopt_device(dt.plan, dt.device_pages, ctx, keep=dt.keep)
    in_bytes = sum(n * 8192 for cols in dt.device_pages for _, n in cols)
    stream = torch.cuda.ExternalStream(ctx.stream)

    def step():
        res = rj.execute_resident(dt.plan, inputs, ctx)
        rows, pages = res.num_rows, res.total_pages()
        res.free()
        return rows, pages

    rows, out_pages = dt.expected_rows, 0
    for _ in range(args.warmup):
        rows, out_pages = step()
    assert rows == dt.expected_rows, (rows, dt.expected_rows)

    ctx.profile_enable(True)
    ctx.profile_reset()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ctx.kernel_launches()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        rows, out_pages = step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - launches0
    clocks = sampler.stop()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ms_per_step = ms_total / args.steps
    value = (nb + np_) / 1e6 / (ms_per_step / 1e3)

    # ---- parity at full size (outside the timed region): multiset checksum of the result pages against
    #      the value derived from the generator without a join (synthetic.row_hash_torch) -------------
    res = rj.execute_resident(dt.plan, inputs, ctx)
    parity = {"rows": res.num_rows == dt.expected_rows,
              "multiset_checksum": syn.result_checksum(ctx, res) == dt.expected_checksum,
              "how": "two wrapping 64-bit sums over a per-row hash of the result pages vs the generator's direct-addressing expectation"}
    res.free()
    assert parity["rows"] and parity["multiset_checksum"], parity

    # ---- roofline of the dominant kernel class (CUDA events recorded around the launches, on the
    #      launching stream, inside the timed region above) -----------------------------------------
    peak, peak_src = measured_peak()
    stages = {}
    for name in ("row_offsets", "decode", "histogram", "scatter", "join", "gather", "encode"):
        st = prof[name]
        if st["launches"] == 0:
            continue
        ms = st["ms"] / args.steps
        gbs = st["bytes"] / args.steps / 1e9 / (ms / 1e3) if ms > 0 else 0.0
        stages[name] = {"ms_per_step": round(ms, 4), "algorithmic_gb_per_step": round(st["bytes"] / args.steps / 1e9, 4),
                        "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4), "launches_per_step": st["launches"] // args.steps}
    dominant = max(stages, key=lambda k: stages[k]["ms_per_step"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(f"scale{args.scale}", {}).get(dominant)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": stages[dominant]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": stages[dominant]["frac"], "traffic": traffic, "peak_source": peak_src,
                "stages": stages}

    # ---- end to end: host pages in (pinned, contiguous), host pages out -----------------------------
    # rj_execute_streamed: the probe table goes through the GPU in row windows, so the upload of window
    # k+1, the kernels of window k and the download of window k-1 overlap on the PCIe link.
    e2e = None
    if not args.no_e2e:
        import numpy as np
        host_plan, keep = syn.to_host_plan(dt)
        root = host_plan.nodes[host_plan.root]
        # pinned result buffers, sized once: dense page count of every output column + room for the
        # partly filled last page of each window
        caps = [-(-dt.expected_rows // int(ctx.lib.rj_fixed_rows_per_page(int(t)))) + 1024 for _, t in root.output_attrs]
        out_bufs = [torch.empty(cap * 8192, dtype=torch.uint8, pin_memory=True).numpy().reshape(-1, 8192) for cap in caps]
        used = [0] * len(caps)

        def alloc(column, _dtype, n_pages):
            lo = used[column]
            used[column] = lo + n_pages
            return out_bufs[column][lo:lo + n_pages]

        e2e_steps = max(1, min(args.steps, 3))
        times = []
        for i in range(1 + e2e_steps):
            used[:] = [0] * len(caps)
            t0 = time.perf_counter()
            e_rows, chunks = rj.execute_streamed(host_plan, ctx, alloc=alloc)
            if i > 0:
                times.append(time.perf_counter() - t0)
        assert e_rows == dt.expected_rows
        parity["e2e_multiset_checksum"] = syn.host_chunks_checksum(ctx, [t for _, t in root.output_attrs], chunks) == dt.expected_checksum
        assert parity["e2e_multiset_checksum"], parity
        sec = sum(times) / len(times)
        e2e = {"value": round((nb + np_) / 1e6 / sec, 2), "unit": UNIT, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": int(sum(used) * 8192), "ms_per_step": round(sec * 1e3, 2),
               "steps": len(times), "api": "rj_execute_streamed (256 MiB windows of the probe table)",
               "host_buffers": "pinned, contiguous per column; timed with the host clock around the call"}
        del keep

    # ---- CPU baseline: the reference's execute() on a bounded sample, host cores ---------------------
    cpu = None
    if not args.no_cpu_baseline:
        cb, cp = FULL_BUILD // CPU_SAMPLE_DIV, FULL_PROBE // CPU_SAMPLE_DIV
        kind, times, crow = time_reference(cpu_sample_plan(cb, cp), cb + cp, 1, 0)
        cpu = {"value": round((cb + cp) / 1e6 / times[0], 4), "unit": UNIT, "cores": os.cpu_count(), "kind": kind,
               "sample": f"config 2 at 1/{CPU_SAMPLE_DIV} scale: {cb} x {cp} rows, {crow} output rows, one execute()"}

    inputs.free()
    rj.destroy_context(ctx)
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": "c2_int32_join_64Mi_x_512Mi_zipf0.75_int64_fp64_payloads" + ("" if args.scale == 1 else f"_div{args.scale}"),
                   "build_rows": nb, "probe_rows": np_, "output_rows": rows, "input_pages_bytes": in_bytes,
                   "output_pages_bytes": out_pages * 8192,
                   "cache": "inputs (%.1f GB) and every intermediate are far larger than the 126 MB L2; no flush needed" % (in_bytes / 1e9),
                   "tuples": "build rows + probe rows (SURVEY 8d)"},
        "clocks": clocks, "gpu_launches": launches, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "parity": parity,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=1, help="divide config 2's row counts (debugging only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from radix_join_b200 import dist_bench
        return dist_bench.run(args, FULL_BUILD // args.scale, FULL_PROBE // args.scale, METRIC, UNIT, ClockSampler, measured_peak)
    return run_single_gpu(args)


if __name__ == "__main__":
    main()
