"""Synthetic inputs of the BASELINE.json configs, generated ON DEVICE and written as 8 KB pages.

This is harness-side code -- the role `ColumnInserter` / `Table::from_csv` play for the reference
(include/plan.h:151-228, src/build_table.cpp:247-303): it produces the ColumnarTables a Plan scans.
torch is used for the random numbers (plumbing); the pages themselves are written by the engine's
`rj_gen_fixed_pages` kernel.  Shapes and seeds follow SURVEY.md section 8d:

  C1  R(k INT32) = permutation of [0, B), S(k INT32) = uniform draws from [0, B); no NULLs
  C2  R(k INT32, a INT64): k = permutation of [0, B) (seed 43), a = splitmix64(k), 1 % NULL in a
      S(k INT32, b FP64):  k = perm[rank], rank ~ Zipf(theta = 0.75) over B ranks (seed 44 + chunk),
                           b = splitmix64(row) as a finite double, 1 % NULL in b
Rows can be sharded: `rank`/`world` select a contiguous 1/world slice of both tables.
"""
import ctypes as C
import math

import numpy as np

from .plan import Column, ColumnarTable, DataType, Plan

_M64 = (1 << 64) - 1


def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def splitmix64_torch(x):
    import torch
    z = x + _s64(0x9E3779B97F4A7C15)
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(0x94D049BB133111EB)
    return z ^ ((z >> 31) & ((1 << 33) - 1))


def splitmix64_numpy(x):
    z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


class Zipf:
    """YCSB / Gray et al. 'Quickly generating billion-record synthetic databases' Zipfian sampler:
    rank in [0, n), P(rank = i) ~ 1 / (i + 1)^theta."""

    def __init__(self, n, theta):
        self.n, self.theta = n, theta
        # zeta(n, theta) in float64, chunked
        acc, step = 0.0, 1 << 24
        for lo in range(1, n + 1, step):
            hi = min(n + 1, lo + step)
            acc += float(np.sum(np.arange(lo, hi, dtype=np.float64) ** -theta))
        self.zetan = acc
        self.zeta2 = 1.0 + 0.5 ** theta
        self.alpha = 1.0 / (1.0 - theta)
        self.eta = (1.0 - (2.0 / n) ** (1.0 - theta)) / (1.0 - self.zeta2 / self.zetan)

    def ranks(self, u, xp):
        """u: uniform [0,1) float64 array (numpy or torch); xp: the array module"""
        uz = u * self.zetan
        body = (self.n * (self.eta * u - self.eta + 1.0) ** self.alpha)
        if xp is np:
            r = np.floor(body).astype(np.int64)
            r = np.where(uz < 1.0, 0, np.where(uz < self.zeta2, 1, r))
            return np.clip(r, 0, self.n - 1)
        r = xp.floor(body).to(xp.int64)
        r = xp.where(uz < 1.0, xp.zeros_like(r), xp.where(uz < self.zeta2, xp.ones_like(r), r))
        return r.clamp_(0, self.n - 1)


def _pack_valid_torch(valid_bool):
    """bool[n] -> uint32 words (as int32 tensor), LSB first, padded with one spare word"""
    import torch
    n = valid_bool.shape[0]
    pad = (-n) % 32
    v = torch.cat([valid_bool, torch.zeros(pad, dtype=torch.bool, device=valid_bool.device)]) if pad else valid_bool
    w = v.view(-1, 32).to(torch.int64)
    weights = (1 << torch.arange(32, device=v.device, dtype=torch.int64))
    words = (w * weights).sum(dim=1)
    words = torch.where(words >= (1 << 31), words - (1 << 32), words).to(torch.int32)
    return torch.cat([words, torch.zeros(1, dtype=torch.int32, device=v.device)])


def _not_null_torch(rows, salt, null_frac):
    """bool[n]: row is not NULL -- a function of the global row number only"""
    h = splitmix64_torch(rows * 2 + salt)
    return ((h >> 11) & 0xFFFFF) >= int(null_frac * (1 << 20))


def _finite_double_bits_torch(bits):
    import torch
    exp_all_ones = ((bits >> 52) & 0x7FF) == 0x7FF
    return torch.where(exp_all_ones, bits & ~(1 << 62), bits)


class DeviceTables:
    """Pages resident in HBM + the Plan that scans them."""

    def __init__(self):
        self.plan = None
        self.device_pages = []   # [table][column] = (device address, n_pages)
        self.keep = []           # torch tensors that own the memory
        self.n_build = 0
        self.n_probe = 0
        self.expected_rows = None
        self.expected_checksum = None  # multiset checksum of the join result (see row_hash_torch)


def _pages_from_dense(ctx, values, valid_words, n, dtype):
    """dense device array -> device pages via the engine's page writer"""
    import torch
    rpp = ctx.lib.rj_fixed_rows_per_page(int(dtype))
    n_pages = (n + rpp - 1) // rpp
    pages = torch.empty(max(n_pages, 1) * 8192, dtype=torch.uint8, device=values.device)
    torch.cuda.synchronize()
    ctx.check(ctx.lib.rj_gen_fixed_pages(ctx.handle, values.data_ptr(), valid_words.data_ptr() if valid_words is not None else None,
                                         n, int(dtype), pages.data_ptr(), None, None))
    torch.cuda.synchronize()
    return pages, n_pages


def single_join_plan(payload):
    plan = Plan()
    if payload:
        plan.new_scan_node(0, [(0, DataType.INT32), (1, DataType.INT64)])
        plan.new_scan_node(1, [(0, DataType.INT32), (1, DataType.FP64)])
        plan.new_join_node(True, 0, 1, 0, 0, [(0, DataType.INT32), (1, DataType.INT64), (3, DataType.FP64)])
    else:
        plan.new_scan_node(0, [(0, DataType.INT32)])
        plan.new_scan_node(1, [(0, DataType.INT32)])
        plan.new_join_node(True, 0, 1, 0, 0, [(0, DataType.INT32), (1, DataType.INT32)])
    plan.root = 2
    return plan


# ---- multiset checksum of a result: size-independent parity at full scale ---------------------------
# A result is a multiset of rows, so it is summarised by two wrapping 64-bit sums over a per-row hash
# (column order matters, row order does not; NULL hashes apart from every value).  For config 2 the
# expected checksum is derived from the generator WITHOUT joining: build keys are a permutation, so the
# build row of a probe key is a direct table lookup.
_NA = 0x6A09E667F3BCC909


def row_hash_torch(cols):
    """cols: [(int64 tensor, bool tensor or None)] -> int64 hash per row"""
    import torch
    h = None
    for c, (v, ok) in enumerate(cols):
        x = splitmix64_torch(v ^ _s64(0x9E3779B97F4A7C15 * (c + 1)))
        if ok is not None:
            x = torch.where(ok, x, torch.full_like(x, _s64(_NA + c)))
        h = x if h is None else splitmix64_torch(h + x)
    return h


def checksum_add(acc, h):
    s1 = (acc[0] + int(h.sum().item())) & _M64
    s2 = (acc[1] + int(splitmix64_torch(h).sum().item())) & _M64
    return (s1, s2, acc[2] + h.shape[0])


def _unpack_valid_torch(words, lo, hi):
    import torch
    idx = torch.arange(lo, hi, device=words.device, dtype=torch.int64)
    return ((words[idx >> 5].to(torch.int64) >> (idx & 31)) & 1).bool()


def decode_fixed_pages(ctx, pages_ptr, n_pages, dtype, num_rows, device="cuda"):
    """device pages of a fixed-width column -> (values tensor, validity words) via the engine's decode"""
    import torch
    row_start = torch.zeros(n_pages + 1, dtype=torch.int64, device=device)
    values = torch.zeros(max(num_rows, 1), dtype=torch.int32 if int(dtype) == int(DataType.INT32) else torch.int64, device=device)
    valid = torch.zeros((num_rows + 31) // 32 + 1, dtype=torch.int32, device=device)
    torch.cuda.synchronize()
    if n_pages:
        ctx.check(ctx.lib.rj_page_row_offsets(ctx.handle, pages_ptr, n_pages, int(dtype), row_start.data_ptr(), None, None))
        ctx.check(ctx.lib.rj_decode_fixed(ctx.handle, pages_ptr, n_pages, int(dtype), row_start.data_ptr(), values.data_ptr(),
                                          valid.data_ptr(), None))
    torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize()
    rows = int(row_start[-1].item()) if n_pages else 0
    return values, valid, rows


def pages_checksum(ctx, columns, acc=(0, 0, 0), chunk=1 << 25):
    """columns: [(DataType, device pages pointer, n_pages)] of ONE page list per column, row-aligned.
    Returns the running (sum, sum of mixed, rows) checksum."""
    import torch
    dec, n = [], None
    for dtype, ptr, n_pages in columns:
        # rows are not known up front for a window's page list: count them, then decode
        row_start = torch.zeros(n_pages + 1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        if n_pages:
            ctx.check(ctx.lib.rj_page_row_offsets(ctx.handle, ptr, n_pages, int(dtype), row_start.data_ptr(), None, None))
        torch.cuda.synchronize()
        rows = int(row_start[-1].item()) if n_pages else 0
        if n is None:
            n = rows
        if rows != n:
            raise AssertionError(f"columns decode to different row counts: {rows} vs {n}")
        dec.append(decode_fixed_pages(ctx, ptr, n_pages, dtype, rows)[:2])
    for lo in range(0, n or 0, chunk):
        hi = min(n, lo + chunk)
        cols = [(v[lo:hi].to(torch.int64), _unpack_valid_torch(w, lo, hi)) for v, w in dec]
        acc = checksum_add(acc, row_hash_torch(cols))
    return acc


def result_checksum(ctx, res):
    """checksum of a device-resident Result (all columns fixed-width)"""
    cols = [(res.column_type(c), res.column_device_ptr(c), res.column_pages(c)) for c in range(res.num_columns)]
    return pages_checksum(ctx, cols)


def host_chunks_checksum(ctx, types, chunks):
    """checksum of what rj.execute_streamed delivered: chunks[column] = list of host page arrays; the k-th
    arrays of all columns belong to the same window"""
    import torch
    acc = (0, 0, 0)
    n_windows = max((len(v) for v in chunks.values()), default=0)
    for k in range(n_windows):
        dev = [torch.from_numpy(chunks[c][k]).cuda() for c in range(len(types))]
        acc = pages_checksum(ctx, [(types[c], dev[c].data_ptr(), dev[c].numel() // 8192) for c in range(len(types))], acc)
    return acc


def make_c2_device(ctx, n_build, n_probe, rank=0, world=1, theta=0.75, null_frac=0.01, device="cuda", checksum=False):
    """Config 2 (or a scaled copy of it) in HBM.  With world > 1 each rank holds rows
    [rank*n/world, (rank+1)*n/world) of both tables.  checksum=True (world 1) also derives the expected
    multiset checksum of the join result from the generator."""
    import torch
    out = DeviceTables()
    g = torch.Generator(device=device)
    g.manual_seed(43)
    perm = torch.randperm(n_build, generator=g, device=device).to(torch.int32)
    b_lo, b_hi = n_build * rank // world, n_build * (rank + 1) // world
    p_lo, p_hi = n_probe * rank // world, n_probe * (rank + 1) // world
    # ---- R(k, a) ----
    rk = perm[b_lo:b_hi].contiguous()
    ra = splitmix64_torch(rk.to(torch.int64))
    # NULLs are a hash of the GLOBAL row number, so every sharding sees the same table
    ra_valid = _pack_valid_torch(_not_null_torch(torch.arange(b_lo, b_hi, device=device, dtype=torch.int64), 1043, null_frac))
    # ---- S(k, b) ---- generated in chunks keyed by the GLOBAL chunk index so every sharding sees the same table
    zipf = Zipf(n_build, theta)
    chunk = 1 << 24
    sk = torch.empty(p_hi - p_lo, dtype=torch.int32, device=device)
    for c0 in range((p_lo // chunk) * chunk, p_hi, chunk):
        gc = torch.Generator(device=device)
        gc.manual_seed(44 + c0 // chunk)
        u = torch.rand(chunk, generator=gc, device=device, dtype=torch.float64)
        keys = perm[zipf.ranks(u, torch)]
        lo, hi = max(c0, p_lo), min(c0 + chunk, p_hi)
        sk[lo - p_lo:hi - p_lo] = keys[lo - c0:hi - c0]
        del u, keys
    rows = torch.arange(p_lo, p_hi, device=device, dtype=torch.int64)
    sb = _finite_double_bits_torch(splitmix64_torch(rows))
    sb_ok = _not_null_torch(rows, 2044, null_frac)
    sb_valid = _pack_valid_torch(sb_ok)
    del rows
    if checksum:
        # (R.k, R.a, S.b) for every probe row of this shard: R.a = splitmix64(k), NULL iff the build row
        # holding k is (the whole permutation is known to every rank); shards add up to the job's checksum
        a_ok_by_key = torch.empty(n_build, dtype=torch.bool, device=device)
        a_ok_by_key[perm.to(torch.int64)] = _not_null_torch(torch.arange(0, n_build, device=device, dtype=torch.int64), 1043, null_frac)
        acc = (0, 0, 0)
        for lo in range(0, p_hi - p_lo, 1 << 25):
            hi = min(p_hi - p_lo, lo + (1 << 25))
            k = sk[lo:hi].to(torch.int64)
            acc = checksum_add(acc, row_hash_torch([(k, torch.ones_like(k, dtype=torch.bool)), (splitmix64_torch(k), a_ok_by_key[k]),
                                                    (sb[lo:hi], sb_ok[lo:hi])]))
        out.expected_checksum = acc
        del a_ok_by_key
    del sb_ok
    del perm
    tables = []
    for cols in (((rk, None, DataType.INT32), (ra, ra_valid, DataType.INT64)),
                 ((sk, None, DataType.INT32), (sb, sb_valid, DataType.FP64))):
        n = cols[0][0].shape[0]
        dev_cols, host_cols = [], []
        for values, valid, dt in cols:
            pages, n_pages = _pages_from_dense(ctx, values, valid, n, dt)
            out.keep.append(pages)
            dev_cols.append((pages.data_ptr(), n_pages))
            host_cols.append(Column(dt))  # placeholder: pages live on the device
        out.device_pages.append(dev_cols)
        tables.append(ColumnarTable(num_rows=n, columns=host_cols))
    out.plan = single_join_plan(payload=True)
    for t in tables:
        out.plan.new_input(t)
    out.n_build, out.n_probe = b_hi - b_lo, p_hi - p_lo
    out.expected_rows = n_probe if world == 1 else None  # every probe key exists on the build side
    return out


def make_c1_device(ctx, n_build, n_probe, device="cuda", checksum=False):
    """Config 1: single INT32 equi-join, unique foreign keys, no payload.  checksum=True also derives the
    expected multiset checksum of the result (k, k) for every probe key, without joining."""
    import torch
    out = DeviceTables()
    g = torch.Generator(device=device)
    g.manual_seed(42)
    rk = torch.randperm(n_build, generator=g, device=device).to(torch.int32)
    sk = torch.randint(0, n_build, (n_probe,), generator=g, device=device, dtype=torch.int64).to(torch.int32)
    tables = []
    for values in (rk, sk):
        n = values.shape[0]
        pages, n_pages = _pages_from_dense(ctx, values, None, n, DataType.INT32)
        out.keep.append(pages)
        out.device_pages.append([(pages.data_ptr(), n_pages)])
        tables.append(ColumnarTable(num_rows=n, columns=[Column(DataType.INT32)]))
    out.plan = single_join_plan(payload=False)
    for t in tables:
        out.plan.new_input(t)
    out.n_build, out.n_probe, out.expected_rows = n_build, n_probe, n_probe
    if checksum:
        acc = (0, 0, 0)
        for lo in range(0, n_probe, 1 << 25):
            k = sk[lo:lo + (1 << 25)].to(torch.int64)
            ok = torch.ones_like(k, dtype=torch.bool)
            acc = checksum_add(acc, row_hash_torch([(k, ok), (k, ok)]))
        out.expected_checksum = acc
    return out


def to_host_plan(dt: DeviceTables, pinned=True):
    """Copy the device pages into (pinned) host memory and return a Plan over HOST pages -- the input
    of the end-to-end measurement."""
    import torch
    plan = Plan()
    plan.nodes, plan.root = dt.plan.nodes, dt.plan.root
    keep = []
    it = iter(dt.keep)
    for t, dev_cols in zip(dt.plan.inputs, dt.device_pages):
        cols = []
        for c, (_, n_pages) in zip(t.columns, dev_cols):
            dev = next(it)
            host = torch.empty(n_pages * 8192, dtype=torch.uint8, pin_memory=pinned)
            host.copy_(dev[: n_pages * 8192])
            keep.append(host)
            cols.append(Column(c.type, host.numpy().reshape(-1, 8192)))
        plan.new_input(ColumnarTable(num_rows=t.num_rows, columns=cols))
    torch.cuda.synchronize()
    return plan, keep
