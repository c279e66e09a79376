"""Development aid: where does execute() spend its time on one JOB plan?  python tools/job_probe.py 24b,17f [scale]"""
import sys, time
sys.path.insert(0, ".")
import radix_join_b200 as rj
from radix_join_b200 import job

names = sys.argv[1].split(",")
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
ctx = rj.build_context(0)
for name in names:
    plan, root_cols, scan_rows = job.make_job(name, scale=scale, seed=1)
    rj.execute(plan, ctx)
    ctx.profile_enable(True)
    ctx.profile_reset()
    t0 = time.perf_counter()
    got = rj.execute(plan, ctx)
    dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    print(name, "scan rows", scan_rows, "out", got.num_rows, "ms %.2f" % (dt * 1e3), "joins", sum(1 for n in plan.nodes if hasattr(n.data, "left")))
    for k, v in prof.items():
        if v["launches"]:
            print("   %-12s %8.3f ms  %4d launches" % (k, v["ms"], v["launches"]))
rj.destroy_context(ctx)
