#!/usr/bin/env python
"""bench.py -- the reference's headline metric on a B200: join throughput (input tuples / s) of
Contest::execute on BASELINE.json config 2 (INT32 join, 64 Mi build x 512 Mi probe, Zipf(0.75),
INT64 + FP64 payloads with 1 % NULLs; synthetic, generated on device).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|job] [--scale S]

One step = one full pass of the hot path: page decode -> radix partition -> shared-memory build+probe
-> gather + page encode, input pages resident in HBM when the timed region starts, result pages left in
HBM (`value`).  `e2e` is `Contest::execute` itself (radix-join_b200/libcontest_b200.so) called by a C++
harness on a Plan of individually new-ed host pages, returning new-ed host pages: host gather, H2D,
kernels, D2H and host scatter all inside the timed region (tests/contest_harness.cpp).
`--impl reference` times the UNMODIFIED reference's CPU `Contest::execute` (oracle/_ref) through the same
harness on a bounded sample of the same workload on the host cores.  `--workload c1` is BASELINE.json
configs[0], which both arms run at full size (a same-config ratio); `--workload job` the 113-plan JOB
suite.  One JSON line is printed by rank 0.
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

METRIC, UNIT = "join_throughput", "Mtuples/s"
# BASELINE.json configs[1] (the headline) and configs[0] (fits both arms at full size: a same-config ratio)
WORKLOADS = {
    "c2": {"build": 1 << 26, "probe": 1 << 29, "name": "c2_int32_join_64Mi_x_512Mi_zipf0.75_int64_fp64_payloads"},
    "c1": {"build": 1_000_000, "probe": 10_000_000, "name": "c1_int32_join_1M_x_10M_unique_fk"},
}
HARNESS = os.path.join(ROOT, "oracle", "_ref", "contest_harness")  # tests/contest_harness.cpp, built by csrc/Makefile `contest`
PARITY_SAMPLE_DIV = 64   # GPU vs reference on the SAME Plan object, outside the timed region


def reference_div(workload, steps, warmup):
    """The reference arm's bounded sample: config 1 runs at full size (about 6-10 s per execute());
    config 2 needs ~300 GB of host RAM and ~4 min per execute() at full size in the reference's
    row-of-variants form, so it runs at 1/16 scale unless the requested step
    count would push the whole run past a few minutes (~15 s per execute() at 1/16)."""
    if workload == "c1":
        return 1
    div = 16
    while (steps + warmup) * 240.0 / div > 200.0 and div < 256:
        div *= 2
    return div


def run_harness(*args, timeout=3000):
    """tests/contest_harness.cpp: Contest::execute on individually new-ed pages (ours or the reference)"""
    if not os.path.exists(HARNESS):
        return None, "oracle/_ref/contest_harness is not built (needs the reference headers at build time)"
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the reference uses OpenMP
    out = subprocess.run([HARNESS, *args], capture_output=True, text=True, timeout=timeout, env=env)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    if not lines:
        return None, f"harness exit {out.returncode}: {out.stderr[-400:]}"
    res = json.loads(lines[-1])
    res["exit_code"] = out.returncode
    return res, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  nvidia-smi needs a few hundred
    ms to produce its first line, longer than a short timed region lasts, so the sampler starts before the
    warm-up steps and mark() / stop() bracket the timed region: samples whose timestamp falls inside it are
    used; if the region was too short to catch two samples, the samples of the warm-up steps (the same
    kernels, same load) are used as well and the line says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []
        self.t_start = self.t_mark = None

    def start(self):
        self.t_start = time.time()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """the timed region starts now"""
        self.t_mark = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        t_end = time.time()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(lo, hi):
            sm, mx, reasons = [], [], set()
            for t, ln in self.lines:
                if not (lo <= t <= hi):
                    continue
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        t0 = self.t_mark if self.t_mark is not None else self.t_start
        sm, mx, reasons = parse(t0, t_end + 0.06)  # a sample is stamped when it is read: allow one period
        window = "timed region"
        if len(sm) < 2:
            sm, mx, reasons = parse(self.t_start, t_end + 0.06)
            window = "warm-up + timed region (the timed region alone is shorter than two sampling periods)"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window}


# --------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (oracle/_ref) on a bounded sample of the workload
# --------------------------------------------------------------------------------------------------
def reference_run(workload, div, steps, warmup):
    """-> (Mtuples/s, ms per execute, kind, cores, sample text) of the reference's Contest::execute, timed
    like the contest harness does (steady_clock around the call only, tests/read_sql.cpp:1234-1236)"""
    w = WORKLOADS[workload]
    nb, np_ = w["build"] // div, w["probe"] // div
    cores = os.cpu_count() or 1
    res, why = run_harness("bench", workload, "--div", str(div), "--steps", str(steps), "--warmup", str(warmup),
                           "--impl", "reference")
    if res is not None and res.get("ok"):
        sec = res["ms_mean"] / 1e3
        scale = "full size" if div == 1 else f"1/{div} scale"
        sample = (f"{w['name']} at {scale}: {nb} x {np_} rows, {res['rows']} output rows, inputs = individually new-ed pages "
                  f"(ColumnInserter), result checked against the generator's multiset checksum")
        return (nb + np_) / 1e6 / sec, res["ms_mean"], "reference", cores, sample
    # no harness binary: the plain-C port of the oracle through ctypes (still the CPU path, kind "port")
    from oracle import pyoracle as orc
    from radix_join_b200 import synthetic as syn
    import numpy as np
    if workload != "c2":
        raise RuntimeError(f"reference arm unavailable: {why}")
    rng = np.random.default_rng(43)
    perm = rng.permutation(nb).astype(np.int32)
    ra = syn.splitmix64_numpy(perm.astype(np.int64).view(np.uint64)).view(np.int64)
    sk = perm[syn.Zipf(nb, 0.75).ranks(np.random.default_rng(44).random(np_), np)]
    bits = syn.splitmix64_numpy(np.arange(np_, dtype=np.uint64))
    bits[((bits >> np.uint64(52)) & np.uint64(0x7FF)) == np.uint64(0x7FF)] &= ~np.uint64(1 << 62)
    va, vb = (rng.random(nb) >= 0.01).astype(np.uint8), (rng.random(np_) >= 0.01).astype(np.uint8)
    plan = syn.single_join_plan(payload=True)
    plan.new_input(orc.encode([orc.Cells.from_values(orc.INT32, perm), orc.Cells(orc.INT64, va, values=ra)], impl="port"))
    plan.new_input(orc.encode([orc.Cells.from_values(orc.INT32, sk), orc.Cells(orc.FP64, vb, values=bits.view(np.float64))], impl="port"))
    kind = "reference" if orc.available("ref") else "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        r = orc.execute(plan, impl="ref" if kind == "reference" else "port", n_threads=cores)
        dt = orc.last_execute_seconds() if kind == "reference" else time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        rows = r.num_rows
        del r
    sec = sum(times) / len(times)
    return (nb + np_) / 1e6 / sec, sec * 1e3, kind, cores, f"{w['name']} at 1/{div} scale: {nb} x {np_} rows, {rows} output rows (harness unavailable: {why})"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference is a single-process CPU program: rank 0 alone runs it
    if args.workload == "job":
        import bench_job
        return bench_job.run(args, impl="reference")
    div = reference_div(args.workload, args.steps, args.warmup)
    value, ms, kind, cores, sample = reference_run(args.workload, div, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["name"], "sample": sample, "sample_div": div,
                   "api": "Contest::execute of the unmodified reference (oracle/_ref/libref_oracle.so) via tests/contest_harness.cpp"},
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
    w = WORKLOADS[args.workload]
    nb, np_ = w["build"] // args.scale, w["probe"] // args.scale
    ctx = rj.build_context(0)
    if args.workload == "c2":
        dt = syn.make_c2_device(ctx, nb, np_, checksum=True)
    else:
        dt = syn.make_c1_device(ctx, nb, np_, checksum=True)
    inputs = rj.adopt_device(dt.plan, dt.device_pages, ctx, keep=dt.keep)
    in_bytes = sum(n * 8192 for cols in dt.device_pages for _, n in cols)
    stream = torch.cuda.ExternalStream(ctx.stream)
    # config 1 (133 MB in, 81 MB out) would sit in the 126 MB L2 from one step to the next: write a
    # buffer larger than L2 between timed steps; config 2's 7.4 GB of inputs need no flush
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if in_bytes < (1 << 30) else None

    def step():
        res = rj.execute_resident(dt.plan, inputs, ctx)
        rows, pages = res.num_rows, res.total_pages()
        res.free()
        return rows, pages

    rows, out_pages = dt.expected_rows, 0
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(args.warmup):
        rows, out_pages = step()
    assert rows == dt.expected_rows, (rows, dt.expected_rows)

    ctx.profile_enable(True)
    ctx.profile_reset()
    launches0 = ctx.kernel_launches()
    torch.cuda.synchronize()
    sampler.mark()
    ms_total = 0.0
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            rows, out_pages = step()
        ev1.record(stream)
        torch.cuda.synchronize()
        ms_total = ev0.elapsed_time(ev1)
    else:
        for _ in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(1)  # 256 MB written: evicts the previous step's lines from L2 (untimed)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            rows, out_pages = step()
            ev1.record(stream)
            torch.cuda.synchronize()
            ms_total += ev0.elapsed_time(ev1)
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
    for name in ("row_offsets", "decode", "histogram", "scatter", "join", "join_emit", "gather", "encode"):
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
            traffic = json.load(f).get(f"{args.workload}_scale{args.scale}", {}).get(dominant)
    except Exception:
        pass
    alg_total = sum(v["algorithmic_gb_per_step"] for v in stages.values())
    # SURVEY 8d's whole-join figure (one-pass radix join with row ids and late materialisation): the yardstick the
    # previous round was judged on, kept so that rounds compare; the fused path's own compulsory traffic is smaller
    survey_gb = 102.2 * (nb + np_) / 1e9 if args.workload == "c2" else 61.6 * (nb + np_) / 1e9
    kernel_names = {"scatter": "scatter_carry_kernel (stage 'scatter': 2 passes x 2 sides = 4 launches)", "join_emit": "join_emit_kernel",
                    "decode": "decode_fixed_kernel", "histogram": "radix_hist_kernel", "join": "join_kernel", "encode": "encode_fixed_kernel"}
    roofline = {"bound": "hbm", "kernel": kernel_names.get(dominant, dominant), "stage": dominant, "achieved": stages[dominant]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": stages[dominant]["frac"], "traffic": traffic, "peak_source": peak_src,
                "numerator": "SURVEY 8d algorithmic bytes of the stage (the scatter is charged its two real passes against the ONE-pass numerator N*12 B plus the carried columns once)",
                "whole_step": {"algorithmic_gb": round(alg_total, 3), "achieved_gbs": round(alg_total / (ms_per_step / 1e3), 1),
                               "frac": round(alg_total / (ms_per_step / 1e3) / peak, 4),
                               "survey_8d_gb": round(survey_gb, 2), "survey_8d_frac": round(survey_gb / (ms_per_step / 1e3) / peak, 4)},
                "stages": stages}

    inputs.free()
    rj.destroy_context(ctx)
    del dt, inputs, flush
    torch.cuda.empty_cache()

    # ---- end to end: Contest::execute on individually new-ed host pages, new-ed host pages out ----------
    # (a separate process, like a contest harness: tests/contest_harness.cpp links nothing of this repo but
    # dlopens radix-join_b200/libcontest_b200.so).  Timed with steady_clock around the call, as
    # tests/read_sql.cpp:1234-1236 does; its result is checked against the generator's multiset checksum.
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 5))
        hres, why = run_harness("bench", args.workload, "--div", str(args.scale), "--steps", str(e2e_steps), "--warmup", "2")
        if hres is None or not hres.get("ok"):
            raise RuntimeError(f"end-to-end run through Contest::execute failed: {why or hres}")
        sec = hres["ms_mean"] / 1e3
        e2e = {"value": round((nb + np_) / 1e6 / sec, 2), "unit": UNIT, "h2d_bytes_per_step": hres["input_pages"] * 8192,
               "d2h_bytes_per_step": hres["output_pages"] * 8192, "ms_per_step": round(hres["ms_mean"], 2),
               "ms_best": round(hres["ms_best"], 2), "steps": e2e_steps, "warmup": 2,
               "api": "Contest::execute (libcontest_b200.so -> rj_execute_pages): individually new-ed pages in (ColumnInserter), new Page out",
               "host_threads": hres["threads"], "build_context_ms": hres["build_context_ms"],
               "timed": "steady_clock around Contest::execute, mean of the timed calls; the previous result is destroyed before the next call"}
        parity["e2e_multiset_checksum"] = bool(hres["checked"] and hres["ok"])

    # ---- the SAME Plan object through both implementations (outside the timed region) -------------------
    if not args.no_cpu_baseline and args.workload == "c2" and args.scale == 1:
        pres, why = run_harness("parity", "c2", "--div", str(PARITY_SAMPLE_DIV))
        parity["sample_vs_reference"] = {"ok": bool(pres and pres.get("ok")), "div": PARITY_SAMPLE_DIV,
                                         "rows": pres.get("rows") if pres else None,
                                         "how": "Contest::execute of this repo and of the unmodified reference on one Plan; results decoded by the reference's Table::from_columnar, sorted, compared exactly"}
        assert parity["sample_vs_reference"]["ok"], (pres, why)

    # ---- CPU baseline: the reference's Contest::execute on a bounded sample, host cores ------------------
    cpu = None
    if not args.no_cpu_baseline:
        div = 1 if args.workload == "c1" else 16 * args.scale
        cvalue, cms, kind, cores, sample = reference_run(args.workload, div, 1, 0)
        cpu = {"value": round(cvalue, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + ", one execute()",
               "ms": round(cms, 1)}

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": w["name"] + ("" if args.scale == 1 else f"_div{args.scale}"),
                   "build_rows": nb, "probe_rows": np_, "output_rows": rows, "input_pages_bytes": in_bytes,
                   "output_pages_bytes": out_pages * 8192,
                   "cache": ("a 256 MB buffer is written between timed steps (inputs + outputs would fit the 126 MB L2)" if in_bytes < (1 << 30) else
                             "inputs (%.1f GB) and every intermediate are far larger than the 126 MB L2; no flush needed" % (in_bytes / 1e9)),
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c1", "job"])
    ap.add_argument("--scale", type=int, default=1, help="divide the workload's row counts (debugging only)")
    ap.add_argument("--job-scale", type=float, default=1.0, help="--workload job: fraction of the IMDB row counts")
    ap.add_argument("--job-cpu-plans", type=int, default=12, help="--workload job: plans the CPU reference also runs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "job":
        import bench_job
        return bench_job.run(args, impl="ours")
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from radix_join_b200 import dist_bench
        w = WORKLOADS["c2"]
        return dist_bench.run(args, w["build"] // args.scale, w["probe"] // args.scale, METRIC, UNIT, ClockSampler, measured_peak)
    return run_single_gpu(args)


if __name__ == "__main__":
    main()
