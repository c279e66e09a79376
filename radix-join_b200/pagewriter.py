"""Harness-side page writer on the device -- the role of the reference's ColumnInserter<T>
(include/plan.h:151-335): dense values + validity in, 8 KB pages out, written by the engine's page
kernels (rj_gen_fixed_pages; rj_varchar_descriptors + rj_encode_varchar_plan / _write) instead of a host
loop.  Any packing that Table::from_columnar decodes is legal (src/build_table.cpp:312-436); fixed-width
pages hold 1984 / 1007 rows, VARCHAR pages follow the engine's parallel layout, strings longer than 8185
bytes become 0xffff / 0xfffe chains.

    pages = fixed_pages(ctx, values, valid, DataType.INT64)           # numpy in, (n_pages, 8192) uint8 out
    pages = varchar_pages(ctx, lengths, chars, valid)

Used by the synthetic JOB generator (radix_join_b200.job) when a context is passed: the host loop over
pages was the slowest part of building IMDB-sized inputs.  There is no CPU path in here.
"""
import ctypes as C

import numpy as np

from ._cabi import PAGE_SIZE
from .plan import DataType


def _pack_valid(valid, n, torch, device):
    """bool[n] -> uint32 bitmap words on the device (one spare word), or None"""
    if valid is None:
        return None
    v = np.asarray(valid, dtype=bool)
    if v.all():
        return None
    bits = np.packbits(v, bitorder="little")
    words = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
    words.view(np.uint8)[: len(bits)] = bits
    return torch.from_numpy(words.view(np.int32)).to(device)


def fixed_pages(ctx, values, valid, dtype):
    """INT32 / INT64 / FP64 column -> pages (host numpy array)"""
    import torch
    dtype = DataType(dtype)
    n = len(values)
    if n == 0:
        return np.zeros((0, PAGE_SIZE), dtype=np.uint8)
    device = torch.device("cuda", ctx.lib.rj_ctx_device(ctx.handle))
    np_dt = {DataType.INT32: np.int32, DataType.INT64: np.int64, DataType.FP64: np.float64}[dtype]
    d_vals = torch.from_numpy(np.ascontiguousarray(values, dtype=np_dt)).to(device)
    d_valid = _pack_valid(valid, n, torch, device)
    rpp = ctx.lib.rj_fixed_rows_per_page(int(dtype))
    n_pages = (n + rpp - 1) // rpp
    d_pages = torch.empty(n_pages * PAGE_SIZE, dtype=torch.uint8, device=device)
    torch.cuda.synchronize(device)
    ctx.check(ctx.lib.rj_gen_fixed_pages(ctx.handle, d_vals.data_ptr(), d_valid.data_ptr() if d_valid is not None else None,
                                         n, int(dtype), d_pages.data_ptr(), None, None))
    torch.cuda.synchronize(device)
    return d_pages.cpu().numpy().reshape(n_pages, PAGE_SIZE)


def varchar_pages(ctx, lengths, chars, valid):
    """lengths[n] (bytes per row; 0 for NULL rows), chars (flat uint8, rows back to back), valid[n] bool or None"""
    import torch
    n = len(lengths)
    if n == 0:
        return np.zeros((0, PAGE_SIZE), dtype=np.uint8)
    device = torch.device("cuda", ctx.lib.rj_ctx_device(ctx.handle))
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=off[1:])
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    if int(off[-1]) != len(chars):
        raise ValueError("lengths do not add up to the character buffer")
    d_off = torch.from_numpy(off).to(device)
    d_chars = torch.from_numpy(chars).to(device) if len(chars) else torch.zeros(16, dtype=torch.uint8, device=device)
    d_valid = _pack_valid(valid, n, torch, device)
    d_desc = torch.empty(n, dtype=torch.int64, device=device)
    torch.cuda.synchronize(device)
    ctx.check(ctx.lib.rj_varchar_descriptors(ctx.handle, d_off.data_ptr(), n, d_desc.data_ptr(), None))
    layout, n_pages = C.c_void_p(), C.c_uint64()
    ctx.check(ctx.lib.rj_encode_varchar_plan(ctx.handle, d_chars.data_ptr(), d_desc.data_ptr(),
                                             d_valid.data_ptr() if d_valid is not None else None, None, n,
                                             C.byref(layout), C.byref(n_pages), None))
    try:
        d_pages = torch.empty(max(int(n_pages.value), 1) * PAGE_SIZE, dtype=torch.uint8, device=device)
        ctx.check(ctx.lib.rj_encode_varchar_write(ctx.handle, layout, d_pages.data_ptr(), None))
        torch.cuda.synchronize(device)
    finally:
        ctx.lib.rj_encode_varchar_free(ctx.handle, layout)
    return d_pages[: int(n_pages.value) * PAGE_SIZE].cpu().numpy().reshape(int(n_pages.value), PAGE_SIZE)
