"""Multi-GPU single join: shard -> radix exchange -> local join.

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  Every rank holds a
contiguous 1/G slice of the rows of both relations.  Ownership of a tuple is decided by the TOP
log2(G) bits of the same 32-bit radix hash the single-GPU engine partitions on (the local partitioner
uses the LOW bits, so the two never interfere):

    1. decode the local pages                         (rj_page_row_offsets, rj_decode_fixed)
    2. histogram + scatter of (key, row) by owner     (rj_radix_histogram / rj_radix_scatter, shift = 32 - g)
       and gather of the payload columns in that order (rj_gather)      -- NULL keys are dropped here
    3. exchange: G x G count matrix, then all-to-all-v of keys / payloads / validity over NVLink
    4. local join of what was received                (rj_join_keys: partition -> build/probe in smem)
    5. gather + page encode of the output columns     (rj_encode_fixed)

The payloads travel with the tuples (early materialisation across the exchange), so step 5 only
touches local memory.  The result of the job is the concatenation of the ranks' page lists: a
ColumnarTable is just per-column page lists (reference include/plan.h:60-62,102-105).

`ops` abstracts the device work: `CudaOps` binds the engine's C-ABI stage entry points; the CPU tests
inject a numpy stand-in (tests/test_dist_gloo.py) to exercise the sharding / exchange logic with gloo.
"""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .plan import DataType, FlatPlan, Plan

INT32, INT64, FP64 = 0, 1, 2


class ResultPages:
    """One output column of this rank's share of the result: pages in device memory (a torch tensor, or a
    column of an engine result that is kept alive here)."""

    def __init__(self, n_pages, type_, tensor=None, result=None, col=None):
        self.n_pages, self.type, self.tensor, self.result, self.col = int(n_pages), int(type_), tensor, result, col

    @property
    def nbytes(self):
        return self.n_pages * 8192

    def copy_to_host(self, host):
        """host: pinned uint8 torch tensor (or numpy array) of at least nbytes"""
        if self.n_pages == 0:
            return
        if self.tensor is not None:
            dst = host if isinstance(host, torch.Tensor) else torch.from_numpy(host)
            dst.view(-1)[: self.nbytes].copy_(self.tensor.view(-1)[: self.nbytes], non_blocking=True)
            torch.cuda.current_stream().synchronize() if self.tensor.is_cuda else None
        else:
            arr = host.numpy() if isinstance(host, torch.Tensor) else host
            self.result.fetch_column(self.col, out=arr.reshape(-1)[: self.nbytes].reshape(-1, 8192))

    def to_numpy(self):
        out = np.empty((self.n_pages, 8192), dtype=np.uint8)
        if self.n_pages:
            if self.tensor is not None:
                out[:] = self.tensor.view(-1)[: self.nbytes].cpu().numpy().reshape(-1, 8192)
            else:
                self.result.fetch_column(self.col, out=out)
        return out


def log2_exact(n):
    g = n.bit_length() - 1
    if n < 1 or (1 << g) != n:
        raise ValueError("world size must be a power of two (ownership = top hash bits)")
    return g


class CudaOps:
    """Stage entry points of librj_b200.so on torch CUDA tensors (device memory plumbing only).
    Engine kernels, torch ops and NCCL (which synchronises with the current stream) all run on ONE
    stream, the engine context's; the host only waits where it needs a count."""

    def __init__(self, ctx):
        self.ctx, self.lib, self.h = ctx, ctx.lib, ctx.handle
        self.device = torch.device("cuda", torch.cuda.current_device())
        # one stream for everything: the engine context's stream becomes torch's current stream
        # (the legacy default stream has handle 0, which the C-ABI reads as "the context stream")
        torch.cuda.synchronize()
        self.torch_stream = torch.cuda.ExternalStream(ctx.stream)
        torch.cuda.set_stream(self.torch_stream)

    stream = None  # C-ABI: NULL = the context stream = torch's current stream (set above)

    def close(self):
        """give torch its default stream back (call before the engine context is destroyed)"""
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.default_stream())

    def sync(self):
        torch.cuda.current_stream().synchronize()

    def empty(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]

    def zeros(self, n, dtype):
        return torch.zeros(max(int(n), 1), dtype=dtype, device=self.device)[:int(n)]

    @staticmethod
    def _p(t):
        return None if t is None else t.data_ptr()

    def decode_fixed(self, pages_ptr, n_pages, type_, n_rows, want_valid):
        """-> (values tensor, validity bitmap words or None)"""
        start = self.empty(n_pages + 1, torch.int64)
        self.ctx.check(self.lib.rj_page_row_offsets(self.h, pages_ptr, n_pages, type_, self._p(start), None, self.stream))
        values = self.empty(n_rows, torch.int32 if type_ == INT32 else torch.int64)
        valid = self.zeros((n_rows + 31) // 32 + 1, torch.int32) if want_valid else None
        self.ctx.check(self.lib.rj_decode_fixed(self.h, pages_ptr, n_pages, type_, self._p(start), self._p(values),
                                                self._p(valid), self.stream))
        return values, valid

    def owner_partition(self, keys, valid, g):
        """group (key, row) by owner rank = top g hash bits -> (keys_out, rows_out, counts[G] on host)"""
        n, G = keys.numel(), 1 << g
        if g == 0:
            raise ValueError("owner_partition needs world > 1")
        hist = self.zeros(G, torch.int32)
        self.ctx.check(self.lib.rj_radix_histogram(self.h, self._p(keys), self._p(valid), n, 4, 32 - g, g, self._p(hist), self.stream))
        counts = hist.cpu().to(torch.int64)  # the one host round trip: the exchange needs the split sizes
        cursor = (torch.cumsum(counts, 0) - counts).to(torch.int32).to(self.device)
        total = int(counts.sum())
        keys_out, rows_out = self.empty(n, torch.int32), self.empty(n, torch.int32)
        self.ctx.check(self.lib.rj_radix_scatter(self.h, self._p(keys), self._p(valid), None, n, 4, 32 - g, g,
                                                 self._p(cursor), self._p(keys_out), self._p(rows_out), self.stream))
        return keys_out[:total], rows_out[:total], counts

    def owner_counts(self, keys, valid, g):
        """tuples per owner rank (host tensor int64[G])"""
        hist = self.zeros(1 << g, torch.int32)
        self.ctx.check(self.lib.rj_radix_histogram(self.h, self._p(keys), self._p(valid), keys.numel(), 4, 32 - g, g,
                                                   self._p(hist), self.stream))
        return hist.cpu().to(torch.int64)

    def scatter_to_peers(self, keys, valid, payloads, g, offsets, xchg):
        """Fused partition + exchange: ONE kernel groups the tuples by owner rank and writes the runs
        (keys, payload values, validity bytes) straight into the owners' receive buffers -- peer-mapped
        symmetric memory, i.e. coalesced stores over NVLink; no collective moves the data."""
        desc = _cabi.rj_scatter_multi_t()
        G = 1 << g
        for d in range(G):
            desc.keys_out[d] = xchg.keys.ptrs[d]
            desc.rows_out[d] = None
        n_pay = 0
        for i, (values, vbits) in enumerate(payloads):
            desc.pay_src[n_pay] = self._p(values)
            desc.pay_width[n_pay] = values.element_size()
            for d in range(G):
                desc.pay_dst[n_pay][d] = xchg.vals[i].ptrs[d]
            n_pay += 1
            if vbits is not None:
                desc.pay_src[n_pay] = self._p(vbits)
                desc.pay_width[n_pay] = 1
                for d in range(G):
                    desc.pay_dst[n_pay][d] = xchg.valids[i].ptrs[d]
                n_pay += 1
        desc.n_payload = n_pay
        cursor = offsets.to(torch.int32).to(self.device)
        self.ctx.check(self.lib.rj_radix_scatter_multi(self.h, self._p(keys), self._p(valid), keys.numel(), 4, 32 - g, g,
                                                       self._p(cursor), C.byref(desc), self.stream))

    def histogram(self, keys, valid, bits, out):
        """tuples per radix digit (the low `bits` hash bits) of this rank's slice, NULL keys excluded -> out[2**bits] (int32, zeroed by the caller)"""
        self.ctx.check(self.lib.rj_radix_histogram(self.h, self._p(keys), self._p(valid), keys.numel(), 4, 0, bits, self._p(out), self.stream))

    def exchange_scatter(self, keys, valid, payloads, shift, bits, cursor, g, xchg):
        """The FIRST scatter pass of the join, run as the exchange (k_scatter_carry.cu, peer destinations): digit d of
        the pass belongs to owner d >> (bits - g); keys, values and one validity byte per tuple are stored straight
        into the owner's receive arrays at cursor[d] -- coalesced runs over NVLink, no collective moves the data."""
        d = _cabi.rj_carry_scatter_t()
        d.d_keys, d.d_valid, d.n = self._p(keys), self._p(valid), keys.numel()
        d.shift, d.bits, d.d_cursor = shift, bits, self._p(cursor)
        G = 1 << g
        d.n_owners, d.owner_shift = G, bits - g
        for o in range(G):
            d.keys_dst_multi[o] = xchg.keys.ptrs[o]
        nv = nf = 0
        for i, (values, vbits) in enumerate(payloads):
            d.val_src[nv], d.val_width[nv] = self._p(values), values.element_size()
            for o in range(G):
                d.val_dst_multi[nv][o] = xchg.vals[i].ptrs[o]
            nv += 1
            if vbits is not None:
                d.flag_src[nf] = self._p(vbits)
                for o in range(G):
                    d.flag_dst_multi[nf][o] = xchg.valids[i].ptrs[o]
                nf += 1
        d.n_val, d.n_flag = nv, nf
        self.ctx.check(self.lib.rj_scatter_carry(self.h, C.byref(d), self.stream))

    def dist_layout(self, H, me, g, bits, p1, ptrs, widths):
        """pull_layout for both sides in ONE kernel (rj_dist_layout).  H: int32 [G * 2 * 2**bits]; ptrs: int64 [2, 5, 8];
        widths: int32 [2, 5].  -> dict of device tensors (see include/rj_b200.h)"""
        ndig, nloc = 1 << p1, (1 << bits) >> g
        out = {"cursor": self.empty(2 * ndig, torch.int32), "table": self.empty(2 * ndig * 5, torch.int64),
               "start": self.empty(2 * (ndig + 1), torch.int32), "tile": self.empty(2 * (ndig + 1), torch.int32),
               "group": self.empty(2 * ndig, torch.int32), "local_hist": self.empty(2 * nloc, torch.int32),
               "scalars": self.empty(4, torch.int64)}
        self.ctx.check(self.lib.rj_dist_layout(self.h, self._p(H), me, g, bits, p1, self._p(ptrs), self._p(widths), self._p(out["cursor"]),
                                               self._p(out["table"]), self._p(out["start"]), self._p(out["tile"]), self._p(out["group"]),
                                               self._p(out["local_hist"]), self._p(out["scalars"]), self.stream))
        return out

    def local_scatter(self, keys, valid, payloads, shift, bits, cursor, xchg, me):
        """scatter pass 1 into THIS rank's arrays of the exchange (symmetric memory: the owners of the regions read
        them from here in their second pass)"""
        d = _cabi.rj_carry_scatter_t()
        d.d_keys, d.d_valid, d.n = self._p(keys), self._p(valid), keys.numel()
        d.shift, d.bits, d.d_cursor = shift, bits, self._p(cursor)
        d.d_keys_out = xchg.keys.ptrs[me]
        nv = nf = 0
        for i, (values, vbits) in enumerate(payloads):
            d.val_src[nv], d.val_dst[nv], d.val_width[nv] = self._p(values), xchg.vals[i].ptrs[me], values.element_size()
            nv += 1
            if vbits is not None:
                d.flag_src[nf], d.flag_dst[nf] = self._p(vbits), xchg.valids[i].ptrs[me]
                nf += 1
        d.n_val, d.n_flag = nv, nf
        self.ctx.check(self.lib.rj_scatter_carry(self.h, C.byref(d), self.stream))

    def join_partitioned(self, sides, hists, local_bits, local_pass1_bits, hash_bits, out_cols):
        """sides = ((keys, [values], [validity bytes or None], [types]) for build, probe), grouped by their pass-1
        digit (or fully partitioned); hists = per-local-partition tuple counts (int32 tensors).
        -> (n_rows, [ResultPages]) or None when a table met duplicate build keys"""
        from .engine import Result
        cs = []
        for side in sides:
            sd = _cabi.rj_part_side_t()
            if isinstance(side, dict):
                # pull: the regions are read from the senders' arrays (see pull_layout)
                sd.d_keys, sd.n, sd.n_cols = None, side["n"], len(side["types"])
                for i, (t, nullable) in enumerate(zip(side["types"], side["nullable"])):
                    sd.d_vals[i], sd.types[i], sd.d_valid_bytes[i] = None, int(t), (1 if nullable else None)
                sd.n_sub = side["table"].shape[0]
                sd.d_src_table, sd.d_sub_start = self._p(side["table"]), self._p(side["start"])
                sd.d_sub_tile, sd.d_sub_group = self._p(side["tile"]), self._p(side["group"])
            else:
                keys, vals, valids, types = side
                sd.d_keys, sd.n, sd.n_cols = self._p(keys), keys.numel(), len(vals)
                for i, (v, vb, t) in enumerate(zip(vals, valids, types)):
                    sd.d_vals[i], sd.types[i], sd.d_valid_bytes[i] = self._p(v), int(t), self._p(vb)
            cs.append(sd)
        outs = (_cabi.rj_part_out_t * len(out_cols))()
        for i, (side, which, _type) in enumerate(out_cols):
            outs[i].side, outs[i].col = (0 if side == "b" else 1), (-1 if which == "key" else int(which))
        h = C.c_void_p()
        self.ctx.check(self.lib.rj_join_partitioned(self.h, C.byref(cs[0]), C.byref(cs[1]), self._p(hists[0]), self._p(hists[1]),
                                                    local_bits, local_pass1_bits, hash_bits, outs, len(out_cols), C.byref(h)))
        if not h.value:
            return None
        res = Result(self.ctx, h)
        cols = [ResultPages(res.column_pages(c), int(res.column_type(c)), result=res, col=c) for c in range(res.num_columns)]
        return res.num_rows, cols

    def gather(self, values, valid, rows):
        """values[rows], valid bits -> (gathered values, uint8 validity per row or None)"""
        n = rows.numel()
        out = self.empty(n, values.dtype)
        out_valid = self.empty((n + 31) // 32 + 1, torch.int32) if valid is not None else None
        self.ctx.check(self.lib.rj_gather(self.h, self._p(values), self._p(valid), self._p(rows), n,
                                          values.element_size(), self._p(out), self._p(out_valid), self.stream))
        if out_valid is None:
            return out, None
        as_bytes = self.empty(n, torch.uint8)  # all-to-all-v segments are not word aligned: ship bytes
        self.ctx.check(self.lib.rj_bitmap_to_bytes(self.h, self._p(out_valid), n, self._p(as_bytes), self.stream))
        return out, as_bytes

    def join_keys(self, build_keys, probe_keys):
        nb, np_ = build_keys.numel(), probe_keys.numel()
        cap = max(nb, np_, 1)
        m = C.c_uint64()
        for _ in range(2):
            ob, op = self.empty(cap, torch.int32), self.empty(cap, torch.int32)
            # rj_join_keys orders itself after `stream` and returns once the pairs are written
            self.ctx.check(self.lib.rj_join_keys(self.h, self._p(build_keys), None, nb, self._p(probe_keys), None, np_, 4,
                                                 cap, self._p(ob), self._p(op), C.byref(m), self.stream))
            if m.value <= cap:
                break
            cap = m.value
        return ob[:m.value], op[:m.value]

    def local_join_encode(self, bk, bvals, bvalids, pk, pvals, pvalids, out_cols):
        """The local join on what this rank owns, through the engine's whole path: the dense columns are
        adopted as two already-decoded base tables and joined by a 2 x Scan -> Join plan, so the root
        join carries the payloads in position order and encodes the result pages itself."""
        from .engine import ResidentInputs, Result
        keep, tables = [], (_cabi.rj_dense_table_t * 2)()
        for t, (keys, vals, valids) in enumerate(((bk, bvals, bvalids), (pk, pvals, pvalids))):
            cols = (_cabi.rj_dense_column_t * (1 + len(vals)))()
            cols[0].type, cols[0].d_values, cols[0].d_valid = INT32, self._p(keys), None
            for i, (v, vb) in enumerate(zip(vals, valids)):
                bits = None
                if vb is not None:
                    bits = self.empty((vb.numel() + 31) // 32 + 1, torch.int32)
                    self.ctx.check(self.lib.rj_bytes_to_bitmap(self.h, self._p(vb), vb.numel(), self._p(bits), self.stream))
                    keep.append(bits)
                cols[1 + i].type = self._types[t][i]
                cols[1 + i].d_values, cols[1 + i].d_valid = self._p(v), self._p(bits)
            keep.append(cols)
            tables[t].num_rows, tables[t].n_columns, tables[t].columns = keys.numel(), 1 + len(vals), cols
        plan = Plan()
        widths = [1 + len(bvals), 1 + len(pvals)]
        for t in range(2):
            plan.new_scan_node(t, [(0, DataType.INT32)] + [(1 + i, DataType(self._types[t][i])) for i in range(widths[t] - 1)])
        outs = []
        for side, which, type_ in out_cols:
            base = 0 if side == "b" else widths[0]
            outs.append((base + (0 if which == "key" else 1 + which), DataType(type_)))
        plan.root = plan.new_join_node(True, 0, 1, 0, 0, outs)
        hin = C.c_void_p()
        self.ctx.check(self.lib.rj_inputs_adopt_dense(self.h, tables, 2, C.byref(hin)))
        inputs = ResidentInputs(self.ctx, hin, keep=keep)
        flat = FlatPlan(plan)  # nodes only: the inputs live in the rj_inputs object
        hres = C.c_void_p()
        self.ctx.check(self.lib.rj_execute_resident(self.h, flat.pointer(), inputs.handle, C.byref(hres)))
        res = Result(self.ctx, hres)
        inputs.free()
        cols = [ResultPages(res.column_pages(c), int(res.column_type(c)), result=res, col=c) for c in range(res.num_columns)]
        return res.num_rows, cols

    _types = ((INT64,), (FP64,))  # payload types per side; set by distributed_join

    def encode_fixed(self, values, valid_bytes, rows, type_):
        """-> device tensor of pages (uint8, n_pages * 8192)"""
        n = rows.numel()
        rpp = self.lib.rj_fixed_rows_per_page(type_)
        n_pages = (n + rpp - 1) // rpp
        pages = self.empty(n_pages * 8192, torch.uint8)
        valid = None
        if valid_bytes is not None:
            valid = self.empty((valid_bytes.numel() + 31) // 32 + 1, torch.int32)
            self.ctx.check(self.lib.rj_bytes_to_bitmap(self.h, self._p(valid_bytes), valid_bytes.numel(), self._p(valid), self.stream))
        self.ctx.check(self.lib.rj_encode_fixed(self.h, self._p(values), self._p(valid), self._p(rows), n, type_,
                                                self._p(pages), self.stream))
        return pages, n_pages


def unpack_bits(words, n):
    """uint32 bitmap words (int32 tensor) -> uint8[n]"""
    shifts = torch.arange(32, device=words.device, dtype=torch.int32)
    bits = (words.view(-1, 1) >> shifts) & 1
    return bits.to(torch.uint8).view(-1)[:n].contiguous()


def pack_bits(valid_bytes):
    """uint8[n] -> uint32 bitmap words (int32 tensor, one spare word)"""
    n = valid_bytes.numel()
    pad = (-n) % 32
    v = valid_bytes
    if pad:
        v = torch.cat([v, torch.zeros(pad, dtype=v.dtype, device=v.device)])
    w = (v.view(-1, 32).to(torch.int64) << torch.arange(32, device=v.device, dtype=torch.int64)).sum(dim=1)
    w = torch.where(w >= (1 << 31), w - (1 << 32), w).to(torch.int32)
    return torch.cat([w, torch.zeros(1, dtype=torch.int32, device=v.device)])


def all_to_all_v(tensor, send_counts, recv_counts, group=None):
    """variable-size all-to-all of a 1-D tensor grouped by destination rank"""
    out = torch.empty(int(sum(recv_counts)), dtype=tensor.dtype, device=tensor.device)
    dist.all_to_all_single(out, tensor.contiguous(), output_split_sizes=[int(c) for c in recv_counts],
                           input_split_sizes=[int(c) for c in send_counts], group=group)
    return out


def exchange_counts(counts, group=None):
    """counts[d] = tuples this rank sends to rank d  ->  recv[s] = tuples rank s sends to this rank"""
    world = dist.get_world_size(group)
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    send = counts.to(torch.int64).to(device)
    recv = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(recv, send, group=group)
    return recv.cpu()


class _SymBuf:
    def __init__(self, n, dtype, device, group):
        import torch.distributed._symmetric_memory as symm
        self.tensor = symm.empty(n, dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.tensor, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]


class PeerExchange:
    """Receive buffers of one relation in symmetric memory: every rank can store into every other
    rank's buffers through NVLink.  Created once, reused by every join."""

    def __init__(self, device, cap_rows, payload_dtypes, payload_nullable, group=None):
        self.cap = int(cap_rows)
        self.keys = _SymBuf(self.cap, torch.int32, device, group)
        self.vals = [_SymBuf(self.cap, dt, device, group) for dt in payload_dtypes]
        self.valids = [_SymBuf(self.cap, torch.uint8, device, group) if nn else None for nn in payload_nullable]

    def barrier(self):
        self.keys.handle.barrier()  # device-side barrier on the current stream, all ranks


def shuffle_relation_p2p(ops, rel, g, xchg, group=None):
    """steps 1-3 with the exchange fused into the partition kernel (peer stores instead of NCCL)"""
    world, me = dist.get_world_size(group), dist.get_rank(group)
    kp, kn, kt, knull = rel.key
    keys, kvalid = ops.decode_fixed(kp, kn, kt, rel.n_rows, knull)
    payloads = [ops.decode_fixed(pp, pn, pt, rel.n_rows, pnull) for (pp, pn, pt, pnull) in rel.payloads]
    counts = ops.owner_counts(keys, kvalid, g)
    # G x G count matrix: row s = what rank s sends to each owner
    mat = torch.empty(world * world, dtype=torch.int64, device=ops.device)
    dist.all_gather_into_tensor(mat, counts.to(ops.device), group=group)
    mat = mat.view(world, world).cpu()
    offsets = mat[:me].sum(dim=0)          # where my run starts in each owner's buffer
    total = int(mat[:, me].sum())          # what I receive
    if int(mat.sum(dim=0).max()) > xchg.cap:
        raise RuntimeError("peer exchange buffers too small for this key distribution")
    xchg.barrier()                         # every rank is done reading the previous contents
    ops.scatter_to_peers(keys, kvalid, payloads, g, offsets, xchg)
    xchg.barrier()                         # every rank's stores have landed
    sent = int(counts.sum() - counts[me])
    sent_bytes = sent * 4
    vals, valids = [], []
    for i, (v, vbits) in enumerate(payloads):
        vals.append(xchg.vals[i].tensor[:total])
        sent_bytes += sent * v.element_size()
        if vbits is not None:
            valids.append(xchg.valids[i].tensor[:total])
            sent_bytes += sent
        else:
            valids.append(None)
    return xchg.keys.tensor[:total], vals, valids, sent_bytes


class Relation:
    """one rank's slice of a relation: a key column and payload columns, as device pages"""

    def __init__(self, n_rows, key, payloads):
        # key / payloads: (pages_ptr, n_pages, type, has_nulls)
        if int(key[2]) != 0:
            # the exchange, the owner histogram and the local join all move 4-byte keys
            raise ValueError("distributed joins take INT32 keys only (got key type %d)" % int(key[2]))
        self.n_rows, self.key, self.payloads = n_rows, key, payloads


def shuffle_relation(ops, rel, g, group=None):
    """steps 1-3 for one relation: returns (keys, [payload values], [payload validity bytes or None])
    of the tuples this rank OWNS, plus the bytes it sent"""
    trace = os.environ.get("RJ_DIST_TRACE") and hasattr(ops, "sync")
    marks = []

    def mark(name):
        if trace:
            ops.sync()
            marks.append((name, time.perf_counter()))

    mark("start")
    kp, kn, kt, knull = rel.key
    keys, kvalid = ops.decode_fixed(kp, kn, kt, rel.n_rows, knull)
    mark("decode_key")
    keys_o, rows_o, counts = ops.owner_partition(keys, kvalid, g)
    mark("owner_partition")
    recv = exchange_counts(counts, group)
    mark("counts")
    sent_bytes = 0
    out_keys = all_to_all_v(keys_o, counts, recv, group)
    mark("a2a_keys")
    sent_bytes += keys_o.numel() * 4
    vals, valids = [], []
    for (pp, pn, pt, pnull) in rel.payloads:
        v, vv = ops.decode_fixed(pp, pn, pt, rel.n_rows, pnull)
        mark("decode_payload")
        gv, gvalid = ops.gather(v, vv, rows_o)
        mark("gather_payload")
        vals.append(all_to_all_v(gv, counts, recv, group))
        sent_bytes += gv.numel() * gv.element_size()
        if gvalid is not None:
            valids.append(all_to_all_v(gvalid, counts, recv, group))
            sent_bytes += gvalid.numel()
        else:
            valids.append(None)
        mark("a2a_payload")
    if trace and dist.get_rank(group) == 0:
        print("[rj dist shuffle]", {n: round((t1 - t0) * 1e3, 2) for (n, t1), (_, t0) in zip(marks[1:], marks[:-1])}, flush=True)
    return out_keys, vals, valids, sent_bytes


def local_tuples(ops, rel, g):
    """this rank's tuples of a relation with NULL keys dropped: (keys, [payload values], [validity bytes or None])"""
    kp, kn, kt, knull = rel.key
    keys, kvalid = ops.decode_fixed(kp, kn, kt, rel.n_rows, knull)
    keys_o, rows_o, _counts = ops.owner_partition(keys, kvalid, g)  # grouped by owner: the order is irrelevant here
    vals, valids = [], []
    for (pp, pn, pt, pnull) in rel.payloads:
        v, vv = ops.decode_fixed(pp, pn, pt, rel.n_rows, pnull)
        gv, gvalid = ops.gather(v, vv, rows_o)
        vals.append(gv)
        valids.append(gvalid)
    return keys_o, vals, valids


def all_gather_v(tensor, group=None):
    """concatenation of every rank's 1-D tensor (different lengths), in rank order"""
    world = dist.get_world_size(group)
    n = int(tensor.numel())
    recv = exchange_counts(torch.full((world,), n, dtype=torch.int64), group)
    return all_to_all_v(tensor.repeat(world), [n] * world, recv, group)


def broadcast_relation(ops, rel, g, group=None):
    """Small relation: every rank receives ALL of its tuples (the build side is broadcast when
    it is small, the probe side then stays where it is and nothing else is exchanged)."""
    world = dist.get_world_size(group)
    keys, vals, valids = local_tuples(ops, rel, g)
    sent = 0
    out_keys = all_gather_v(keys, group)
    sent += keys.numel() * keys.element_size() * (world - 1)
    out_vals, out_valids = [], []
    for v, vb in zip(vals, valids):
        out_vals.append(all_gather_v(v, group))
        sent += v.numel() * v.element_size() * (world - 1)
        if vb is not None:
            out_valids.append(all_gather_v(vb, group))
            sent += vb.numel() * (world - 1)
        else:
            out_valids.append(None)
    return out_keys, out_vals, out_valids, sent


# --------------------------------------------------------------------------------------------------
# The fused path: the join's first scatter pass IS the exchange
# --------------------------------------------------------------------------------------------------
JOIN_TARGET_FILL = 2048   # kJoinTargetFill (csrc/rj_internal.h): build tuples per final partition
MAX_TOTAL_BITS, MAX_PASS_BITS = 15, 8


def choose_bits(n_build_total):
    """radix bits of the whole job, exactly as the single-GPU engine picks them (engine.cu: root_fused)"""
    bits = 0
    while (n_build_total >> bits) > JOIN_TARGET_FILL and bits < MAX_TOTAL_BITS:
        bits += 1
    return bits


def exchange_layout(H, me, g, bits, p1):
    """Where everything goes, from the all-gathered histograms alone (device tensor math, no host round trip).

    H[s, side, f] = tuples of relation `side` (0 build, 1 probe) on rank s whose final partition (the low `bits`
    hash bits) is f.  The exchange moves tuples by their top `p1` bits of f (the pass-1 digit d, or all of f when
    the join needs one pass); digit d is owned by rank d >> (p1 - g), so a rank ends up with a contiguous range of
    final partitions.  Inside an owner's receive arrays the digits lie in order, and inside a digit the senders'
    runs lie in rank order.  Returns
        cursor[side, d]      first index, in the owner's arrays, of THIS rank's run of digit d
        local_hist[side, :]  tuples per final partition of the range this rank owns (input of the local plan)
        owned[side]          tuples this rank receives
        per_owner[side, o]   tuples rank o receives (capacity check)
        sent[side]           tuples this rank stores into OTHER ranks' arrays"""
    G, nfin, ndig = H.shape[0], 1 << bits, 1 << p1
    per = ndig >> g
    Cnt = H.view(G, 2, ndig, nfin // ndig).sum(-1)          # [G, 2, ndig]
    tot = Cnt.sum(0).view(2, G, per)
    base = (torch.cumsum(tot, -1) - tot).view(2, ndig)      # start of digit d inside its owner's arrays
    cursor = base + Cnt[:me].sum(0)
    lo = me * (nfin >> g)
    local_hist = H.sum(0)[:, lo: lo + (nfin >> g)].contiguous()
    mine = Cnt[me].view(2, G, per).sum(-1)                  # what I send to each owner
    sent = mine.sum(-1) - mine[:, me]
    return cursor, local_hist, local_hist.sum(-1), tot.sum(-1), sent


SCATTER_TILE = 4096  # tuples per scatter tile (csrc/k_scatter_carry.cu)


def pull_layout(H, me, g, bits, p1, side, array_ptrs, array_widths):
    """The pull variant: every rank runs scatter pass 1 into its OWN arrays (grouped by pass-1 digit, cursor = the
    exclusive prefix of its own digit counts), and the owner of a digit reads that digit's G runs where they lie
    -- its second pass takes them as sub-regions (region-major, sender-minor).  From H alone:
        cursor[d]           where this rank's run of digit d starts in its own arrays
        table[x, a]         byte address of array a (keys, value 0, value 1, flag 0, flag 1) of sub-region x, biased
                            so that the element index is the sub-region's virtual position
        start / tile        exclusive prefixes of the sub-regions' tuple and tile counts (n_sub + 1 entries)
        group[x]            the pass-1 region (local numbering) sub-region x belongs to
    array_ptrs[a] = int64 tensor [G] of every rank's base address of array a (0 where the array does not exist)."""
    G, nfin, ndig = H.shape[0], 1 << bits, 1 << p1
    per = ndig >> g
    Cnt = H[:, side].reshape(G, ndig, nfin // ndig).sum(-1)              # [G, ndig]
    run_start = torch.cumsum(Cnt, -1) - Cnt                              # sender-local start of every digit run
    mine = torch.arange(per, device=H.device) + me * per
    cnt = Cnt[:, mine].T.reshape(-1)                                     # [per * G] region-major, sender-minor
    zero = torch.zeros(1, dtype=torch.int64, device=H.device)
    start = torch.cat([zero, torch.cumsum(cnt, 0)])
    tile = torch.cat([zero, torch.cumsum((cnt + SCATTER_TILE - 1) // SCATTER_TILE, 0)])
    group = torch.arange(per, device=H.device).repeat_interleave(G)
    delta = run_start[:, mine].T.reshape(-1) - start[:-1]                # element bias of every sub-region
    sender = torch.arange(G, device=H.device).repeat(per)
    table = torch.stack([array_ptrs[a][sender] + delta * array_widths[a] for a in range(5)], dim=1).contiguous()
    return run_start[me], table, start, tile, group


def distributed_join_fused(ops, build, probe, out_cols, xchg, group=None, total_build_rows=None):
    """Key / foreign-key join of two sharded relations where the first scatter pass of the single-GPU algorithm is
    the exchange: every rank histograms its slice over the job's radix digits, the histograms are all-gathered,
    each rank derives from them where its runs start in every owner's receive arrays, ONE kernel per relation
    scatters keys + carried columns into the owners' memory over NVLink, and each rank then runs the second pass
    and the fused build / probe / page-output kernel on the range of partitions it owns.  Host round trips: one
    (the owned tuple counts).  Returns None when the shape is not eligible (INT32 keys, at most two fixed-width
    carried columns per side, at least as many radix bits as ranks' bits) or a table met duplicate build keys --
    the caller then runs `distributed_join`."""
    world, me = dist.get_world_size(group), dist.get_rank(group)
    g = log2_exact(world)
    if len(build.payloads) > 2 or len(probe.payloads) > 2:
        return None
    if total_build_rows is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        tb = torch.tensor([build.n_rows], dtype=torch.int64, device=device)
        dist.all_reduce(tb, group=group)
        total_build_rows = int(tb)
    bits = choose_bits(total_build_rows)
    two = bits > MAX_PASS_BITS
    p1 = (bits + 1) // 2 if two else bits
    if p1 < g or bits - g < 0 or g == 0 or (two and p1 == g):
        return None
    trace = os.environ.get("RJ_DIST_TRACE") and hasattr(ops, "sync")
    t = [time.perf_counter()]

    def mark():
        if trace:
            ops.sync()
            t.append(time.perf_counter())

    # 1. decode, histogram over the job's digits
    rels, hist = [], ops.zeros(2 << bits, torch.int32)
    for side, rel in enumerate((build, probe)):
        kp, kn, kt, knull = rel.key
        keys, kvalid = ops.decode_fixed(kp, kn, kt, rel.n_rows, knull)
        pays = [ops.decode_fixed(pp, pn, pt, rel.n_rows, pnull) for (pp, pn, pt, pnull) in rel.payloads]
        ops.histogram(keys, kvalid, bits, hist[side << bits: (side + 1) << bits])
        rels.append((keys, kvalid, pays))
    mark()
    # 2. all ranks' histograms -> layout (device math), one small read-back
    H = torch.empty(world * (2 << bits), dtype=torch.int32, device=hist.device)
    dist.all_gather_into_tensor(H, hist, group=group)
    pull = two and os.environ.get("RJ_DIST_MODE", "pull") == "pull"
    payload_bytes = [sum((4 if p[2] == INT32 else 8) + (1 if p[3] else 0) for p in rel.payloads) for rel in (build, probe)]
    if pull:
        # 3. scatter pass 1 stays local (into this rank's symmetric arrays); 4. the owners PULL: their second pass
        #    reads every region's runs from the senders' memory over NVLink (TMA bulk loads), so the transfer
        #    overlaps the partitioning and nothing is copied twice.
        all_ptrs, all_widths = [], []
        for rel, x in zip((build, probe), xchg):
            if rel.n_rows > x.cap:
                raise RuntimeError("peer exchange buffers too small for this rank's slice")
            widths = [4] + [(4 if p[2] == INT32 else 8) for p in rel.payloads] + [0] * (2 - len(rel.payloads)) + [1, 1]
            ptrs = [x.keys.ptrs]
            ptrs += [x.vals[i].ptrs for i in range(len(rel.payloads))] + [[0] * world] * (2 - len(rel.payloads))
            flag_bufs = [x.valids[i] for i in range(len(rel.payloads)) if rel.payloads[i][3]]
            ptrs += [fb.ptrs for fb in flag_bufs] + [[0] * world] * (2 - len(flag_bufs))
            all_ptrs.append([list(p) + [0] * (8 - world) for p in ptrs])
            all_widths.append(widths)
        if hasattr(ops, "dist_layout"):
            key = (id(xchg[0]), id(xchg[1]), tuple(map(tuple, all_widths)))
            cache = getattr(ops, "_layout_consts", None)
            if cache is None or cache[0] != key:
                cache = (key, torch.tensor(all_ptrs, dtype=torch.int64, device=hist.device), torch.tensor(all_widths, dtype=torch.int32, device=hist.device))
                ops._layout_consts = cache
            lay = ops.dist_layout(H, me, g, bits, p1, cache[1], cache[2])
            ndig, nloc = 1 << p1, (1 << bits) >> g
            info = lay["scalars"].cpu()            # the one host round trip: what this rank owns
            n_own, n_sent = [int(info[0]), int(info[2])], [int(info[1]), int(info[3])]
            cursors = [lay["cursor"][s * ndig: (s + 1) * ndig] for s in range(2)]
            local_hist32 = [lay["local_hist"][s * nloc: (s + 1) * nloc] for s in range(2)]
            subs = [{"table": lay["table"][s * ndig * 5: (s + 1) * ndig * 5].view(ndig, 5), "start": lay["start"][s * (ndig + 1): (s + 1) * (ndig + 1)],
                     "tile": lay["tile"][s * (ndig + 1): (s + 1) * (ndig + 1)], "group": lay["group"][s * ndig: (s + 1) * ndig]} for s in range(2)]
        else:
            H64 = H.view(world, 2, 1 << bits).to(torch.int64)
            _cur, local_hist, owned, _per_owner, sent = exchange_layout(H64, me, g, bits, p1)
            local_hist32 = local_hist.to(torch.int32).contiguous()
            info = torch.cat([owned, sent]).cpu()
            n_own, n_sent = [int(x) for x in info[:2]], [int(x) for x in info[2:4]]
            cursors, subs = [], []
            for side in range(2):
                tptrs = [torch.tensor(p[:world], dtype=torch.int64, device=hist.device) for p in all_ptrs[side]]
                cur, table, start, tile, grp = pull_layout(H64, me, g, bits, p1, side, tptrs, all_widths[side])
                cursors.append(cur.to(torch.int32).contiguous())
                subs.append({"table": table, "start": start.to(torch.int32).contiguous(), "tile": tile.to(torch.int32).contiguous(),
                             "group": grp.to(torch.int32).contiguous()})
        mark()
        xchg[0].barrier()                  # every owner is done reading the previous contents of these arrays
        for side, (keys, kvalid, pays) in enumerate(rels):
            ops.local_scatter(keys, kvalid, pays, bits - p1, p1, cursors[side], xchg[side], me)
        xchg[0].barrier()                  # every rank's first pass is complete
        mark()
        sides, sent_bytes = [], 0
        for side, rel in enumerate((build, probe)):
            sides.append(dict(subs[side], n=n_own[side], types=[p[2] for p in rel.payloads], nullable=[bool(p[3]) for p in rel.payloads]))
            sent_bytes += n_sent[side] * (4 + payload_bytes[side])
        exchange_how = f"scatter pass 1 local, the owners' pass 2 reads its regions from the senders' memory (TMA over NVLink, {p1} of {bits} radix bits)"
    else:
        H64 = H.view(world, 2, 1 << bits).to(torch.int64)
        cursor, local_hist, owned, per_owner, sent = exchange_layout(H64, me, g, bits, p1)
        cursor32, local_hist32 = cursor.to(torch.int32).contiguous(), local_hist.to(torch.int32).contiguous()
        info = torch.cat([owned, per_owner.max(-1).values, sent]).cpu()
        n_own, worst, n_sent = [int(x) for x in info[:2]], [int(x) for x in info[2:4]], [int(x) for x in info[4:6]]
        if worst[0] > xchg[0].cap or worst[1] > xchg[1].cap:
            raise RuntimeError("peer exchange buffers too small for this key distribution")
        mark()
        # 3. the exchange = scatter pass 1 into the owners' arrays
        xchg[0].barrier()                      # every rank is done with the previous contents of its receive arrays
        for side, (keys, kvalid, pays) in enumerate(rels):
            ops.exchange_scatter(keys, kvalid, pays, bits - p1, p1, cursor32[side], g, xchg[side])
        xchg[0].barrier()                      # every rank's stores have landed
        mark()
        # 4. local: pass 2 inside every received region, then build + probe + page output
        sides, sent_bytes = [], 0
        for side, (rel, x) in enumerate(zip((build, probe), xchg)):
            n = n_own[side]
            vals = [x.vals[i].tensor[:n] for i in range(len(rel.payloads))]
            valids = [x.valids[i].tensor[:n] if rel.payloads[i][3] else None for i in range(len(rel.payloads))]
            types = [p[2] for p in rel.payloads]
            sides.append((x.keys.tensor[:n], vals, valids, types))
            sent_bytes += n_sent[side] * (4 + sum((4 if p[2] == INT32 else 8) + (1 if p[3] else 0) for p in rel.payloads))
        exchange_how = f"scatter pass 1 of the join written into the owners' memory (peer stores, {p1} of {bits} radix bits)"
    got = ops.join_partitioned(sides, (local_hist32[0], local_hist32[1]), bits - g, (p1 - g) if two else 0, bits, out_cols)
    mark()
    if got is None:
        if pull:
            return None  # duplicate build keys: the caller runs the general distributed join
        # duplicate build keys: the general local join on what was received (any order will do)
        ops._types = (tuple(p[2] for p in build.payloads), tuple(p[2] for p in probe.payloads))
        (bk, bvals, bvalids, _), (pk, pvals, pvalids, _) = sides
        got = ops.local_join_encode(bk, bvals, bvalids, pk, pvals, pvalids, out_cols)
    n_rows, cols = got
    stats = {"sent_bytes": sent_bytes, "owned_build": n_own[0], "owned_probe": n_own[1],
             "exchange": exchange_how}
    if trace:
        names = ["decode+histogram", "layout", "exchange", "pass 2 + join + pages"]
        stats["phase_ms"] = {n: round((b - a) * 1e3, 3) for n, a, b in zip(names, t[:-1], t[1:])}
        if me == 0:
            print("[rj dist fused]", stats["phase_ms"], flush=True)
    return n_rows, cols, stats


def distributed_join(ops, build, probe, out_cols, group=None, xchg=None, broadcast_max_rows=0):
    """Inner equi-join of two sharded relations.  out_cols: list of ("b"|"p", "key"|payload index, type).
    xchg = (PeerExchange for build, PeerExchange for probe) switches the exchange from NCCL all-to-all-v to
    the fused partition + peer-store kernel.  A build side of at most `broadcast_max_rows` rows in total is
    broadcast instead: every rank joins its own probe rows against the whole build side.
    Returns (n_rows, [ResultPages], stats) for THIS rank's share of the result."""
    world = dist.get_world_size(group)
    g = log2_exact(world)
    trace = os.environ.get("RJ_DIST_TRACE") and hasattr(ops, "sync")
    t = [time.perf_counter()]

    def mark():
        if trace:
            ops.sync()
            t.append(time.perf_counter())

    broadcast = False
    if broadcast_max_rows:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        total_b = torch.tensor([build.n_rows], dtype=torch.int64, device=device)
        dist.all_reduce(total_b, group=group)
        broadcast = int(total_b) <= broadcast_max_rows
    if broadcast:
        bk, bvals, bvalids, sent_b = broadcast_relation(ops, build, g, group)
        mark()
        pk, pvals, pvalids = local_tuples(ops, probe, g)
        sent_p = 0
        mark()
    elif xchg is not None:
        bk, bvals, bvalids, sent_b = shuffle_relation_p2p(ops, build, g, xchg[0], group)
        mark()
        pk, pvals, pvalids, sent_p = shuffle_relation_p2p(ops, probe, g, xchg[1], group)
        mark()
    else:
        bk, bvals, bvalids, sent_b = shuffle_relation(ops, build, g, group)
        mark()
        pk, pvals, pvalids, sent_p = shuffle_relation(ops, probe, g, group)
        mark()
    if hasattr(ops, "local_join_encode"):
        # whole-path local join (CUDA engine): partition -> build/probe -> carried payloads -> page encode
        ops._types = (tuple(p[2] for p in build.payloads), tuple(p[2] for p in probe.payloads))
        n_rows, cols = ops.local_join_encode(bk, bvals, bvalids, pk, pvals, pvalids, out_cols)
        mark()
    else:
        ob, op = ops.join_keys(bk, pk)
        mark()
        cols = []
        for side, which, type_ in out_cols:
            rows = ob if side == "b" else op
            if which == "key":
                values, valid = (bk if side == "b" else pk), None
            else:
                values = (bvals if side == "b" else pvals)[which]
                valid = (bvalids if side == "b" else pvalids)[which]
            pages, n_pages = ops.encode_fixed(values, valid, rows, type_)
            cols.append(ResultPages(n_pages, type_, tensor=pages))
        n_rows = int(ob.numel())
    mark()
    stats = {"sent_bytes": sent_b + sent_p, "owned_build": int(bk.numel()), "owned_probe": int(pk.numel()),
             "exchange": "build side broadcast (all-gather), probe side in place" if broadcast
                         else "peer stores (fused into the partition kernel)" if xchg is not None else "collective all-to-all-v"}
    if trace:
        names = ["shuffle_build", "shuffle_probe", "join(+encode)", "encode"]
        stats["phase_ms"] = {n: round((b - a) * 1e3, 3) for n, a, b in zip(names, t[:-1], t[1:])}
        if dist.get_rank(group) == 0:
            print("[rj dist]", stats["phase_ms"], flush=True)
    return n_rows, cols, stats
