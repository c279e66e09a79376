"""CPU test of the N > 1 path: world_size 2 (and 4) over gloo.  The sharding / ownership / exchange /
concatenation logic of radix_join_b200.dist_join runs unchanged; the device kernels are replaced by a
numpy stand-in defined HERE (test infrastructure -- the product binds the CUDA C-ABI and nothing else)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from helpers import FP64, INT32, INT64, orc, rj


class NumpyOps:
    """same interface as dist_join.CudaOps, numpy + oracle codec underneath (CPU tensors for gloo)"""

    def decode_fixed(self, pages, n_pages, type_, n_rows, want_valid):
        cells = orc.decode(rj.Column(type_, pages), n_rows, impl="port")
        vals = cells.values.view(np.int32 if type_ == INT32 else np.int64)
        valid = torch.from_numpy(cells.valid.copy()) if want_valid else None
        return torch.from_numpy(vals.copy()), valid  # validity kept as bytes in this stand-in

    def owner_partition(self, keys, valid, g):
        k = keys.numpy()
        ok = np.ones(len(k), bool) if valid is None else valid.numpy().astype(bool)
        owner = H.hash_keys(k.view(np.uint32)) >> np.uint32(32 - g)
        rows = np.nonzero(ok)[0]
        order = np.argsort(owner[rows], kind="stable")
        rows = rows[order].astype(np.int32)
        counts = np.bincount(owner[ok], minlength=1 << g)
        return torch.from_numpy(k[rows].copy()), torch.from_numpy(rows), torch.from_numpy(counts.astype(np.int64))

    def gather(self, values, valid, rows):
        r = rows.numpy().astype(np.int64)
        out = torch.from_numpy(values.numpy()[r].copy())
        return out, (torch.from_numpy(valid.numpy()[r].copy()) if valid is not None else None)

    def join_keys(self, build_keys, probe_keys):
        b, p = build_keys.numpy(), probe_keys.numpy()
        order = np.argsort(b, kind="stable")
        bs = b[order]
        lo, hi = np.searchsorted(bs, p, "left"), np.searchsorted(bs, p, "right")
        cnt = hi - lo
        pi = np.repeat(np.arange(len(p)), cnt)
        offs = np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        bi = order[np.repeat(lo, cnt) + offs]
        return torch.from_numpy(bi.astype(np.int32)), torch.from_numpy(pi.astype(np.int32))

    def encode_fixed(self, values, valid_bytes, rows, type_):
        r = rows.numpy().astype(np.int64)
        v = values.numpy()[r]
        valid = valid_bytes.numpy()[r] if valid_bytes is not None else np.ones(len(r), np.uint8)
        dt = {INT32: np.int32, INT64: np.int64, FP64: np.float64}[type_]
        col = orc.encode([orc.Cells(type_, valid, values=v.view(dt))], impl="port").columns[0] if len(r) else rj.Column(type_)
        return torch.from_numpy(col.pages.reshape(-1).copy()), col.n_pages


def make_tables(seed, n_b, n_p):
    rng = np.random.default_rng(seed)
    bk = orc.Cells(INT32, (rng.random(n_b) > 0.05).astype(np.uint8), values=rng.integers(0, n_b // 2, n_b).astype(np.int32))
    ba = H.random_cells(rng, INT64, n_b, null_frac=0.1)
    pk = orc.Cells(INT32, (rng.random(n_p) > 0.05).astype(np.uint8), values=rng.integers(0, n_b, n_p).astype(np.int32))
    pb = H.random_cells(rng, FP64, n_p, null_frac=0.1)
    return (bk, ba), (pk, pb)


def slice_cells(c, lo, hi):
    return orc.Cells(c.type, c.valid[lo:hi], values=c.values[lo:hi])


def worker(rank, world, port, outdir, n_b, n_p, broadcast_max_rows=0):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radix_join_b200 import dist_join as dj
    (bk, ba), (pk, pb) = make_tables(5, n_b, n_p)
    rels = []
    for (k, v), n, vt in (((bk, ba), n_b, INT64), ((pk, pb), n_p, FP64)):
        lo, hi = n * rank // world, n * (rank + 1) // world
        t = orc.encode([slice_cells(k, lo, hi), slice_cells(v, lo, hi)], impl="port")
        rels.append(dj.Relation(hi - lo, (t.columns[0].pages, t.columns[0].n_pages, INT32, True),
                                [(t.columns[1].pages, t.columns[1].n_pages, vt, True)]))
    out_cols = [("b", "key", INT32), ("b", 0, INT64), ("p", 0, FP64), ("p", "key", INT32)]
    rows, cols, stats = dj.distributed_join(NumpyOps(), rels[0], rels[1], out_cols, broadcast_max_rows=broadcast_max_rows)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), rows=rows, sent=stats["sent_bytes"], exchange=stats["exchange"],
             **{f"c{i}": c.to_numpy().reshape(-1) for i, c in enumerate(cols)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,broadcast_max_rows", [(2, 0), (4, 0), (2, 10_000), (4, 10_000), (2, 100)])
def test_distributed_join_matches_oracle(world, broadcast_max_rows):
    """hash-distributed exchange, and the broadcast of a small build side; a limit below the
    build side's size must fall back to the exchange"""
    n_b, n_p = 3000, 9000
    port = 29500 + os.getpid() % 2000 + world + (7 if broadcast_max_rows else 0) + (broadcast_max_rows == 100) * 11
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(worker, args=(world, port, d, n_b, n_p, broadcast_max_rows), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(world)]
    total = int(sum(p["rows"] for p in parts))
    types = [INT32, INT64, FP64, INT32]
    # the job's result = the ranks' page lists appended in rank order, column by column
    got = rj.ColumnarTable(num_rows=total, columns=[
        rj.Column(t, np.concatenate([p[f"c{i}"].reshape(-1, 8192) for p in parts])) for i, t in enumerate(types)])
    (bk, ba), (pk, pb) = make_tables(5, n_b, n_p)
    plan = H.single_join_plan(orc.encode([bk, ba], impl="port"), orc.encode([pk, pb], impl="port"),
                              [INT32, INT64], [INT32, FP64], 0, 0, True, out_cols=[0, 1, 3, 2])
    want = orc.execute(plan, impl="port")
    assert want.num_rows == total and total > 0
    assert orc.result_equal(got, want)
    assert all(int(p["sent"]) > 0 for p in parts)
    used_broadcast = "broadcast" in str(parts[0]["exchange"])
    assert used_broadcast == (broadcast_max_rows >= n_b)


def test_world_size_must_be_power_of_two():
    from radix_join_b200 import dist_join as dj
    assert dj.log2_exact(8) == 3
    with pytest.raises(ValueError):
        dj.log2_exact(6)


def test_bit_packing_roundtrip():
    from radix_join_b200 import dist_join as dj
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 1000):
        v = torch.from_numpy((rng.random(n) > 0.3).astype(np.uint8))
        words = dj.pack_bits(v)
        assert torch.equal(dj.unpack_bits(words, n), v)
        ref = np.packbits(v.numpy(), bitorder="little")
        assert np.array_equal(words.numpy().view(np.uint8)[: len(ref)], ref)
