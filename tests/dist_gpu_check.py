"""Multi-GPU parity check, run under torchrun on a GPU box (not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Every rank joins its shard of a scaled config 2 through radix_join_b200.dist_join (CUDA ops + NCCL);
rank 0 appends the ranks' result pages and compares the multiset of rows with the single-GPU engine
AND the CPU oracle on the unsharded tables."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import radix_join_b200 as rj  # noqa: E402
from radix_join_b200 import dist_bench, dist_join as dj, synthetic as syn  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rj.build_context(local)
    ops = dj.CudaOps(ctx)
    outdir = os.environ.get("RJ_CHECK_DIR", "/tmp/rj_dist_check")
    os.makedirs(outdir, exist_ok=True)
    ok = True
    p2p = os.environ.get("RJ_DIST_EXCHANGE", "p2p") == "p2p"
    # 2^19 build rows: 8 radix bits, one scatter pass (the exchange delivers final partitions); 2^21: 10 bits, two passes
    for n_b, n_p in ((1 << 19, 1 << 22), (1 << 21, 1 << 23)):
        dt = syn.make_c2_device(ctx, n_b, n_p, rank=rank, world=world)
        build, probe = dist_bench._relations(dt)
        xchg = None
        if p2p:
            xchg = (dj.PeerExchange(ops.device, n_b, [torch.int64], [True]), dj.PeerExchange(ops.device, n_p, [torch.int64], [True]))
        single = want = inputs = None
        modes = [("exchange", lambda: dj.distributed_join(ops, build, probe, dist_bench.OUT_COLS, xchg=xchg))]
        if p2p:
            # the join's first scatter pass as the exchange (twice: the receive arrays are reused)
            modes.append(("fused", lambda: dj.distributed_join_fused(ops, build, probe, dist_bench.OUT_COLS, xchg, total_build_rows=n_b)))
            modes.append(("fused again", lambda: dj.distributed_join_fused(ops, build, probe, dist_bench.OUT_COLS, xchg)))

            def pushed():
                os.environ["RJ_DIST_MODE"] = "push"  # peer stores instead of the owners pulling
                try:
                    return dj.distributed_join_fused(ops, build, probe, dist_bench.OUT_COLS, xchg, total_build_rows=n_b)
                finally:
                    del os.environ["RJ_DIST_MODE"]
            modes.append(("fused push", pushed))
        if n_b == 1 << 19:
            # the broadcast of the (here: forced) small build side
            modes.append(("broadcast", lambda: dj.distributed_join(ops, build, probe, dist_bench.OUT_COLS, broadcast_max_rows=n_b)))
        for mode, run in modes:
            out = run()
            assert out is not None, f"{mode}: not eligible?"
            rows, cols, stats = out
            if rank == 0:
                print(f"{mode}:", stats["exchange"], flush=True)
            np.savez(os.path.join(outdir, f"rank{rank}.npz"), rows=rows, **{f"c{i}": c.to_numpy().reshape(-1) for i, c in enumerate(cols)})
            dist.barrier()
            if rank == 0:
                from oracle import pyoracle as orc
                parts = [np.load(os.path.join(outdir, f"rank{r}.npz")) for r in range(world)]
                total = int(sum(p["rows"] for p in parts))
                types = [0, 1, 2]
                got = rj.ColumnarTable(num_rows=total, columns=[
                    rj.Column(t, np.concatenate([p[f"c{i}"].reshape(-1, 8192) for p in parts])) for i, t in enumerate(types)])
                if single is None:
                    full = syn.make_c2_device(ctx, n_b, n_p)
                    inputs = rj.adopt_device(full.plan, full.device_pages, ctx, keep=full.keep)
                    res = rj.execute_resident(full.plan, inputs, ctx)
                    single = res.to_columnar()
                    res.free()
                    host_plan, _keep = syn.to_host_plan(full, pinned=False)
                    want = orc.execute(host_plan, impl="port")
                good = total == n_p == single.num_rows == want.num_rows
                good = good and orc.result_equal(got, single) and orc.result_equal(single, want)
                print(f"dist parity ({world} GPUs, {n_b} x {n_p}, {mode}): rows {total}, sent/rank {stats['sent_bytes']} B ->", "OK" if good else "MISMATCH", flush=True)
                ok = ok and good
            dist.barrier()
        if inputs is not None:
            inputs.free()
        del xchg, dt
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    ok = bool(int(flag))
    ops.close()
    dist.destroy_process_group()
    rj.destroy_context(ctx)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
