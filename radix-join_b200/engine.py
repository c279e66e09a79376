"""`build_context / destroy_context / execute` -- the reference's operator interface
(include/plan.h:337-344, src/execute.cpp:316-330) on top of the CUDA engine's C-ABI.

`execute(plan, ctx)` takes host pages and returns a ColumnarTable of host pages, exactly like the
reference; the split API (`upload` + `execute_resident`) keeps the base tables resident in HBM for
the device-only measurement of bench.py.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import PAGE_SIZE
from .plan import Column, ColumnarTable, DataType, FlatPlan, Plan


class EngineError(RuntimeError):
    """The engine's error contract: the reference throws std::runtime_error (src/execute.cpp:280)."""


class Context:
    """The opaque `void* context` of Contest::build_context()."""

    def __init__(self, device=0):
        """device: one index, or a list of indices = a device group driven by this process (rj_ctx_create_multi)"""
        self.lib = _cabi.load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*[int(d) for d in device])
            rc = self.lib.rj_ctx_create_multi(arr, len(device), C.byref(h))
        else:
            rc = self.lib.rj_ctx_create(int(device), C.byref(h))
        if rc != 0:
            msg = self.lib.rj_last_error(None)
            raise EngineError(msg.decode() if msg else f"rj_ctx_create failed ({rc})")
        self.handle = h

    def check(self, rc):
        if rc != 0:
            msg = self.lib.rj_last_error(self.handle)
            raise EngineError(msg.decode() if msg else f"engine error {rc}")

    def close(self):
        if self.handle:
            self.lib.rj_ctx_destroy(self.handle)
            self.handle = None

    @property
    def group_size(self):
        return int(self.lib.rj_ctx_group_size(self.handle))

    @property
    def sm_count(self):
        return self.lib.rj_ctx_sm_count(self.handle)

    @property
    def stream(self):
        """the cudaStream_t (as an int) all whole-path calls run on"""
        return int(self.lib.rj_ctx_stream(self.handle) or 0)

    def kernel_launches(self):
        return int(self.lib.rj_kernel_launch_count())

    # ---- profiling -------------------------------------------------------------------------------
    def profile_enable(self, on=True):
        self.check(self.lib.rj_profile_enable(self.handle, 1 if on else 0))

    def profile_reset(self):
        self.check(self.lib.rj_profile_reset(self.handle))

    def profile_read(self):
        stats = (_cabi.rj_stage_stat_t * _cabi.RJ_ST_COUNT)()
        self.check(self.lib.rj_profile_read(self.handle, stats))
        return {
            _cabi.STAGE_NAMES[i]: {"ms": stats[i].ms, "launches": int(stats[i].launches),
                                   "bytes": int(stats[i].bytes)}
            for i in range(_cabi.RJ_ST_COUNT)
        }


class Result:
    """Result pages resident in HBM until fetched."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.handle = handle

    @property
    def num_rows(self):
        return int(self.ctx.lib.rj_result_num_rows(self.handle))

    @property
    def num_columns(self):
        return int(self.ctx.lib.rj_result_num_columns(self.handle))

    def column_type(self, c):
        return DataType(self.ctx.lib.rj_result_column_type(self.handle, c))

    def column_pages(self, c):
        return int(self.ctx.lib.rj_result_column_pages(self.handle, c))

    def column_device_ptr(self, c):
        return int(self.ctx.lib.rj_result_column_device_ptr(self.handle, c))

    def total_pages(self):
        return sum(self.column_pages(c) for c in range(self.num_columns))

    def fetch_column(self, c, out=None):
        n = self.column_pages(c)
        if out is None:
            out = np.empty((n, PAGE_SIZE), dtype=np.uint8)
        if n:
            self.ctx.check(self.ctx.lib.rj_result_fetch(self.ctx.handle, self.handle, c, None,
                                                        out.ctypes.data))
        return out

    def to_columnar(self) -> ColumnarTable:
        t = ColumnarTable(num_rows=self.num_rows)
        for c in range(self.num_columns):
            t.columns.append(Column(self.column_type(c), self.fetch_column(c)))
        return t

    def free(self):
        if self.handle:
            self.ctx.lib.rj_result_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ResidentInputs:
    def __init__(self, ctx: Context, handle, keep=None):
        self.ctx = ctx
        self.handle = handle
        self._keep = keep

    def free(self):
        if self.handle:
            self.ctx.lib.rj_inputs_free(self.ctx.handle, self.handle)
            self.handle = None


def build_context(device=0) -> Context:
    """Contest::build_context (src/execute.cpp:326-328)"""
    return Context(device)


def destroy_context(ctx: Context):
    """Contest::destroy_context (src/execute.cpp:330)"""
    if ctx is not None:
        ctx.close()


def execute_to_device(plan: Plan, ctx: Context) -> Result:
    """Host pages in, result pages left in HBM."""
    flat = FlatPlan(plan)
    h = C.c_void_p()
    ctx.check(ctx.lib.rj_execute(ctx.handle, flat.pointer(), C.byref(h)))
    return Result(ctx, h)


def execute(plan: Plan, ctx: Context) -> ColumnarTable:
    """Contest::execute (src/execute.cpp:316-324): host pages in, an owning ColumnarTable out."""
    res = execute_to_device(plan, ctx)
    try:
        return res.to_columnar()
    finally:
        res.free()


def execute_streamed(plan: Plan, ctx: Context, chunk_bytes=0, alloc=None):
    """Host pages in, host pages out, with upload / kernels / download overlapped (rj_execute_streamed).

    `alloc(column, data_type, n_pages)` returns a C-contiguous uint8 array of n_pages * 8192 bytes for
    one window's result pages of a column (pinned memory makes the download a true DMA); default:
    numpy.  Returns (num_rows, pages) with pages[column] = list of the arrays that were filled.
    """
    flat = FlatPlan(plan)
    chunks = {}
    failure = []

    def sink(_user, column, dtype, n_pages):
        try:
            if alloc is None:
                buf = np.empty((n_pages, PAGE_SIZE), dtype=np.uint8)
            else:
                buf = alloc(column, DataType(dtype), n_pages)
            if buf.nbytes < n_pages * PAGE_SIZE or not buf.flags["C_CONTIGUOUS"]:
                raise ValueError("sink buffer too small or not contiguous")
            chunks.setdefault(column, []).append(buf)
            return buf.ctypes.data
        except Exception as e:  # an exception must not cross the C frame
            failure.append(e)
            return None

    cb = _cabi.rj_page_sink_t(sink)
    n = C.c_uint64(0)
    rc = ctx.lib.rj_execute_streamed(ctx.handle, flat.pointer(), int(chunk_bytes), cb, None, C.byref(n))
    if failure:
        raise failure[0]
    ctx.check(rc)
    return int(n.value), chunks


def execute_streamed_columnar(plan: Plan, ctx: Context, chunk_bytes=0) -> ColumnarTable:
    """execute_streamed, collected into an owning ColumnarTable (column types from the root's output_attrs)."""
    n, chunks = execute_streamed(plan, ctx, chunk_bytes)
    t = ColumnarTable(num_rows=n)
    root = plan.nodes[plan.root]
    for c, (_, dtype) in enumerate(root.output_attrs):
        parts = [b.reshape(-1, PAGE_SIZE) for b in chunks.get(c, [])]
        pages = np.concatenate(parts) if parts else None
        t.columns.append(Column(dtype, pages))
    return t


def upload(plan: Plan, ctx: Context) -> ResidentInputs:
    """Copy every input column's pages into HBM once (bench: inputs resident before the timer)."""
    flat = FlatPlan(plan)
    h = C.c_void_p()
    ctx.check(ctx.lib.rj_inputs_upload(ctx.handle, flat.tables, flat.n_tables, C.byref(h)))
    return ResidentInputs(ctx, h)


def adopt_device(plan: Plan, device_pages, ctx: Context, keep=None) -> ResidentInputs:
    """Use pages that already live in device memory: device_pages[table][column] = (address, n_pages)."""
    flat = FlatPlan(plan, device_pages=device_pages)
    h = C.c_void_p()
    ctx.check(ctx.lib.rj_inputs_adopt_device(ctx.handle, flat.tables, flat.n_tables, C.byref(h)))
    return ResidentInputs(ctx, h, keep=keep)


def execute_resident(plan: Plan, inputs: ResidentInputs, ctx: Context) -> Result:
    flat = FlatPlan(plan, device_pages=[[(0, 0) for _ in t.columns] for t in plan.inputs])
    h = C.c_void_p()
    ctx.check(ctx.lib.rj_execute_resident(ctx.handle, flat.pointer(), inputs.handle, C.byref(h)))
    return Result(ctx, h)


def tables_equal(a: ColumnarTable, b: ColumnarTable, ctx: Context):
    """Sorted-multiset comparison of two ColumnarTables on the device (the harness's `compare`,
    tests/read_sql.cpp:1159-1222): decode, hash every row, radix-sort, compare the rows at equal rank cell by
    cell.  -> (equal, number of differing cell / hash pairs)"""
    pa, pb = Plan(), Plan()
    pa.new_input(a)
    pb.new_input(b)
    fa, fb = FlatPlan(pa), FlatPlan(pb)
    eq, bad = C.c_int32(0), C.c_uint64(0)
    ctx.check(ctx.lib.rj_tables_equal(ctx.handle, fa.tables, fb.tables, C.byref(eq), C.byref(bad)))
    return bool(eq.value), int(bad.value)


def execute_pages(plan: Plan, ctx: Context) -> ColumnarTable:
    """Contest::execute's own data path (rj_execute_pages): every result page is allocated individually by a
    callback, the transfers are pipelined, and a device-group context runs eligible plans on all its GPUs."""
    flat = FlatPlan(plan)
    root = plan.nodes[plan.root]
    cols = [[] for _ in root.output_attrs]
    keep, failure = [], []
    import threading
    lock = threading.Lock()

    def new_pages(_user, n, out):
        try:
            bufs = [np.empty(PAGE_SIZE, dtype=np.uint8) for _ in range(n)]
            with lock:
                keep.extend(bufs)
            for i, b in enumerate(bufs):
                out[i] = b.ctypes.data
            return 0
        except Exception as e:  # noqa: BLE001
            failure.append(e)
            return 1

    def append(_user, column, _type, pages, n):
        cols[column].extend(int(pages[i]) for i in range(n))
        return 0

    def free_pages(_user, _n, _pages):
        return None

    alloc = _cabi.rj_page_alloc_t()
    f_new = _cabi.rj_page_alloc_t._fields_[1][1](new_pages)
    f_app = _cabi.rj_page_alloc_t._fields_[2][1](append)
    f_free = _cabi.rj_page_alloc_t._fields_[3][1](free_pages)
    alloc.user, alloc.new_pages, alloc.append, alloc.free_pages = None, f_new, f_app, f_free
    n = C.c_uint64(0)
    rc = ctx.lib.rj_execute_pages(ctx.handle, flat.pointer(), 0, C.byref(alloc), C.byref(n))
    if failure:
        raise failure[0]
    ctx.check(rc)
    by_addr = {b.ctypes.data: b for b in keep}
    t = ColumnarTable(num_rows=int(n.value))
    for c, (_, dtype) in enumerate(root.output_attrs):
        pages = np.stack([by_addr[a] for a in cols[c]]) if cols[c] else None
        t.columns.append(Column(dtype, pages))
    return t
