"""bench.py --gpus N (N > 1): config 2 sharded over N GPUs of one node, strong scaling.
Launched by torchrun, one rank per GPU; rank 0 prints the JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist

from . import dist_join as dj
from . import synthetic as syn
from .engine import build_context, destroy_context


def _relations(dt):
    (bk, ba), (pk, pb) = dt.device_pages
    build = dj.Relation(dt.n_build, (bk[0], bk[1], dj.INT32, False), [(ba[0], ba[1], dj.INT64, True)])
    probe = dj.Relation(dt.n_probe, (pk[0], pk[1], dj.INT32, False), [(pb[0], pb[1], dj.FP64, True)])
    return build, probe


OUT_COLS = [("b", "key", dj.INT32), ("b", 0, dj.INT64), ("p", 0, dj.FP64)]  # (R.k, R.a, S.b)
NVLINK_PEAK_GBS = 770.0  # peer copy bandwidth measured on this pool in round 1 (900 GB/s nominal per direction)


def _nvlink_roofline(stats, xchg_gb, scatter_ms, join_ms, hbm_peak, hbm_src):
    """the exchange kernel against the NVLink roofline: in the pull variant the owners' scatter pass 2 reads the
    regions out of the senders' memory, so its CUDA-event time (engine stage `scatter`, max over ranks) moves the
    bytes one rank exchanges per step"""
    pulled = "pass 2 reads" in stats["exchange"] or "pass 2 reads its regions" in stats["exchange"]
    achieved = xchg_gb / (scatter_ms / 1e3) if (pulled and scatter_ms > 0) else None
    return {"bound": "nvlink", "kernel": "scatter_carry_kernel (regions): scatter pass 2 pulling its regions from the peers" if pulled else "partition + exchange",
            "achieved": round(achieved, 1) if achieved else None, "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
            "frac": round(achieved / NVLINK_PEAK_GBS, 4) if achieved else None, "traffic": None,
            "stage_ms_max_over_ranks": {"scatter_pass2": round(scatter_ms, 3), "join_emit": round(join_ms, 3)},
            "note": f"bytes one rank exchanges per step: {xchg_gb:.3f} GB (>= {xchg_gb / NVLINK_PEAK_GBS * 1e3:.2f} ms at {NVLINK_PEAK_GBS:.0f} GB/s); the same kernel "
                    f"also partitions what it reads, so it is bound by the slower of NVLink and its own shared-memory work; "
                    f"single-GPU kernel rooflines are in the --gpus 1 line; HBM peak {hbm_peak} GB/s ({hbm_src})"}


def run(args, n_build, n_probe, metric, unit, ClockSampler, measured_peak):
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    # rank 0 prints exactly one line on stdout (NCCL_DEBUG=VERSION in this image would put a banner there)
    # NCCL's log (version banner, rank / channel set-up) goes to a file per rank next to stderr, not to stdout
    os.environ["NCCL_DEBUG"] = os.environ.get("RJ_NCCL_DEBUG", "INFO")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = build_context(local)
    ops = dj.CudaOps(ctx)
    dt = syn.make_c2_device(ctx, n_build, n_probe, rank=rank, world=world, checksum=True)
    build, probe = _relations(dt)
    in_bytes = sum(n * 8192 for cols in dt.device_pages for _, n in cols)

    # receive buffers in symmetric memory for the fused partition + exchange kernel; the collective
    # (NCCL all-to-all-v) exchange remains as the fallback if symmetric memory cannot be set up
    xchg, xchg_note = None, ""
    if os.environ.get("RJ_DIST_EXCHANGE", "p2p") == "p2p":
        try:
            cap_b = int(n_build / world * 1.25) + (1 << 20)
            cap_p = int(n_probe / world * 1.25) + (1 << 20)
            xchg = (dj.PeerExchange(ops.device, cap_b, [torch.int64], [True]),
                    dj.PeerExchange(ops.device, cap_p, [torch.int64], [True]))
        except Exception as e:  # noqa: BLE001
            xchg, xchg_note = None, f" (symmetric memory unavailable: {type(e).__name__}: {e})"
            if rank == 0:
                print("[rj dist] falling back to the NCCL exchange:", xchg_note, flush=True)

    def step():
        if xchg is not None and os.environ.get("RJ_DIST_FUSED", "1") != "0":
            out = dj.distributed_join_fused(ops, build, probe, OUT_COLS, xchg, total_build_rows=n_build)
            if out is not None:
                return out
        return dj.distributed_join(ops, build, probe, OUT_COLS, xchg=xchg)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up steps: nvidia-smi takes longer to start than a short timed region lasts
    for _ in range(args.warmup):
        rows, cols, stats = step()
    total_rows = torch.tensor([rows], dtype=torch.int64, device="cuda")
    dist.all_reduce(total_rows)
    assert int(total_rows) == n_probe, (int(total_rows), n_probe)
    # parity at full size: the ranks' result pages add up to the multiset checksum the generator predicts
    # (a rank's RESULT rows are not its INPUT shard's rows, only the sums over all ranks agree)
    def pages_of(c):
        return (c.type, c.tensor.data_ptr() if c.tensor is not None else c.result.column_device_ptr(c.col), c.n_pages)
    got = syn.pages_checksum(ctx, [pages_of(c) for c in cols])
    both = [None] * world
    dist.all_gather_object(both, (got, dt.expected_checksum))
    m64 = (1 << 64) - 1
    sum_got = tuple(sum(g[i] for g, _ in both) & m64 for i in range(3))
    sum_exp = tuple(sum(e[i] for _, e in both) & m64 for i in range(3))
    parity = {"rows": True, "multiset_checksum": sum_got == sum_exp,
              "how": "per-rank sums over a per-row hash of the result pages, added over ranks, vs the generator's join-free expectation"}
    assert parity["multiset_checksum"], (sum_got, sum_exp)

    launches0 = ctx.kernel_launches()
    ctx.profile_enable(True)   # CUDA events around the engine's stages (here: scatter pass 2 = the NVLink pull, join + pages)
    ctx.profile_reset()
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.mark()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        rows, cols, stats = step()
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)   # the job is as slow as its slowest rank
    launches = ctx.kernel_launches() - launches0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    stage_ms = torch.tensor([prof["scatter"]["ms"] / args.steps, prof["join_emit"]["ms"] / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(stage_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = float(ms) / args.steps
    value = (n_build + n_probe) / 1e6 / (ms_per_step / 1e3)
    sent = torch.tensor([stats["sent_bytes"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(sent, op=dist.ReduceOp.MAX)
    out_bytes = sum(c.nbytes for c in cols)

    # end to end: this rank's input pages start in pinned host memory, its result pages end there
    e2e = None
    if not args.no_e2e:
        it = iter(dt.keep)
        host_in, dev_in = [], []
        for cols_d in dt.device_pages:
            for _, n_pages in cols_d:
                d = next(it)
                h = torch.empty(n_pages * 8192, dtype=torch.uint8, pin_memory=True)
                h.copy_(d[: n_pages * 8192])
                host_in.append(h)
                dev_in.append(d)
        host_out = [torch.empty(max(c.nbytes, 8192), dtype=torch.uint8, pin_memory=True) for c in cols]
        times = []
        for i in range(1 + max(1, min(args.steps, 3))):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for h, d in zip(host_in, dev_in):
                d[: h.numel()].copy_(h, non_blocking=True)
            r, c, _ = step()
            for col, h in zip(c, host_out):
                col.copy_to_host(h)  # page counts are fixed by the row count (fixed rows per page)
            torch.cuda.synchronize()
            dt_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt_s, op=dist.ReduceOp.MAX)
            if i > 0:
                times.append(float(dt_s))
        sec = sum(times) / len(times)
        tot_in = torch.tensor([in_bytes, out_bytes], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot_in)
        # the floor of this number: what the box moves between pinned host memory and ALL the GPUs at once, both
        # directions busy (the ranks share PCIe switches and the host's memory system)
        probe_bytes = 256 << 20
        hp_in = torch.empty(probe_bytes, dtype=torch.uint8, pin_memory=True)
        hp_out = torch.empty(probe_bytes, dtype=torch.uint8, pin_memory=True)
        dp_in = torch.empty(probe_bytes, dtype=torch.uint8, device="cuda")
        dp_out = torch.empty(probe_bytes, dtype=torch.uint8, device="cuda")
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        best = None
        for _ in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _rep in range(4):
                with torch.cuda.stream(s_up):
                    dp_in.copy_(hp_in, non_blocking=True)
                with torch.cuda.stream(s_down):
                    hp_out.copy_(dp_out, non_blocking=True)
            torch.cuda.synchronize()
            dt_p = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt_p, op=dist.ReduceOp.MAX)
            best = float(dt_p) if best is None else min(best, float(dt_p))
        agg_gbs_each_way = 4 * probe_bytes * world / 1e9 / best
        e2e = {"value": round((n_build + n_probe) / 1e6 / sec, 2), "unit": unit, "h2d_bytes_per_step": int(tot_in[0]),
               "d2h_bytes_per_step": int(tot_in[1]), "ms_per_step": round(sec * 1e3, 2), "steps": len(times),
               "host_buffers": "pinned, contiguous per column and rank; host clock, max over ranks",
               "pcie_probe": {"aggregate_gbs_each_way": round(agg_gbs_each_way, 1),
                              "what": f"all {world} ranks copy 1 GiB host->device and 1 GiB device->host at the same time (pinned memory)",
                              "floor_ms": round(max(int(tot_in[0]), int(tot_in[1])) / 1e9 / agg_gbs_each_way * 1e3, 1)}}

    if rank == 0:
        peak, peak_src = measured_peak()
        xchg_gbs = float(sent) / 1e9
        line = {
            "metric": metric, "value": round(value, 2), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "c2_int32_join_64Mi_x_512Mi_zipf0.75_int64_fp64_payloads" + ("" if args.scale == 1 else f"_div{args.scale}"),
                       "build_rows": n_build, "probe_rows": n_probe, "output_rows": n_probe,
                       "parallelism": f"{world} ranks, rows sharded 1/{world}, ownership = top {dj.log2_exact(world)} hash bits, exchange = {stats['exchange']}{xchg_note}",
                       "cache": "per-rank inputs and intermediates are far larger than the 126 MB L2; no flush needed",
                       "tuples": "build rows + probe rows (SURVEY 8d)"},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": _nvlink_roofline(stats, xchg_gbs, float(stage_ms[0]), float(stage_ms[1]), peak, peak_src),
            "e2e": e2e, "cpu_baseline": None, "parity": parity,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    ops.close()
    dist.destroy_process_group()
    destroy_context(ctx)
    # torch still holds blocks / pinned buffers / events that were created while the engine's stream
    # was current; their destructors at interpreter exit race with CUDA's own teardown ("context is
    # destroyed" on some ranks).  The job is done and its line is printed: leave without teardown.
    import sys
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)
