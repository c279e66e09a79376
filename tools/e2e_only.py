import sys, time, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import radix_join_b200 as rj
from radix_join_b200 import synthetic as syn
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nb, np_ = (64 << 20) // scale, (512 << 20) // scale
ctx = rj.build_context(0)
dt = syn.make_c2_device(ctx, nb, np_)
host_plan, keep = syn.to_host_plan(dt)
root = host_plan.nodes[host_plan.root]
caps = [-(-dt.expected_rows // int(ctx.lib.rj_fixed_rows_per_page(int(t)))) + 1024 for _, t in root.output_attrs]
out_bufs = [torch.empty(cap * 8192, dtype=torch.uint8, pin_memory=True).numpy().reshape(-1, 8192) for cap in caps]
used = [0] * len(caps)
def alloc(column, _dtype, n_pages):
    lo = used[column]; used[column] = lo + n_pages
    return out_bufs[column][lo:lo + n_pages]
for mb in (512, 256, 128):
    for i in range(3):
        used[:] = [0] * len(caps)
        t0 = time.perf_counter()
        rows, _ = rj.execute_streamed(host_plan, ctx, chunk_bytes=mb << 20, alloc=alloc)
        print("streamed", mb, "MiB windows", i, "%.1f ms" % ((time.perf_counter() - t0) * 1e3), rows, file=sys.stderr)
