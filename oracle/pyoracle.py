"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the two CPU checkers (see oracle/oracle.h).

    impl="ref"  : the UNMODIFIED reference, oracle/_ref/libref_oracle.so   (oracle/ref_shim.cpp)
    impl="port" : the plain-C restatement, oracle/librj_oracle.so          (oracle/rj_oracle.c)

Plus the helpers every parity test needs: numpy cells <-> pages through the checker's own codec, and
a canonical (sorted-multiset) form of a result so that outputs whose row order is free
(reference tests/unit_tests.cpp:6-8, tests/read_sql.cpp:1155-1157) can be compared exactly.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import radix_join_b200 as rj  # noqa: E402  (struct layouts of include/rj_b200.h)
from radix_join_b200 import _cabi  # noqa: E402

PAGE_SIZE = 8192
INT32, INT64, FP64, VARCHAR = 0, 1, 2, 3
PORT_LIB = os.path.join(_HERE, "librj_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libref_oracle.so")


class orc_cells_t(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("reserved", C.c_uint32),
        ("rows", C.c_uint64),
        ("valid", C.c_void_p),
        ("values", C.c_void_p),
        ("str_off", C.c_void_p),
        ("chars", C.c_void_p),
    ]


_libs = {}


def build(ref=True, quiet=True):
    """Compile the checkers (oracle/Makefile).  The reference half needs /root/reference."""
    targets = ["port"] + (["ref"] if ref and os.path.isdir("/root/reference") else [])
    out = subprocess.run(["make", "-C", _HERE] + targets, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout[-2000:] + out.stderr[-4000:])
    if not quiet:
        print(out.stdout[-2000:])


def available(impl):
    return os.path.exists(REF_LIB if impl == "ref" else PORT_LIB)


def _load(impl):
    if impl in _libs:
        return _libs[impl]
    path = REF_LIB if impl == "ref" else PORT_LIB
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle {'ref' if impl == 'ref' else 'port'}`")
    lib = C.CDLL(path)
    p = "ref_" if impl == "ref" else "orc_"
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    sig = {
        "execute": (C.c_int, [C.POINTER(_cabi.rj_plan_t), C.c_int, C.POINTER(vp), C.c_char_p, C.c_size_t]),
        "encode": (C.c_int, [C.POINTER(orc_cells_t), u32, u64, C.POINTER(vp), C.c_char_p, C.c_size_t]),
        "result_num_rows": (u64, [vp]),
        "result_num_columns": (u32, [vp]),
        "result_column_type": (C.c_int32, [vp, u32]),
        "result_column_pages": (u64, [vp, u32]),
        "result_column_data": (vp, [vp, u32]),
        "result_free": (None, [vp]),
        "decode_sizes": (C.c_int, [C.POINTER(_cabi.rj_column_t), u64, C.POINTER(u64), C.c_char_p, C.c_size_t]),
        "decode_fill": (C.c_int, [C.POINTER(_cabi.rj_column_t), u64, vp, vp, vp, vp, C.c_char_p, C.c_size_t]),
    }
    ns = {}
    for name, (res, args) in sig.items():
        fn = getattr(lib, p + name)
        fn.restype, fn.argtypes = res, args
        ns[name] = fn
    if impl == "ref":
        lib.ref_last_execute_seconds.restype = C.c_double
        ns["last_execute_seconds"] = lib.ref_last_execute_seconds
    else:
        lib.orc_hash_int.restype, lib.orc_hash_int.argtypes = u64, [C.c_int64]
        lib.orc_num_buckets.restype, lib.orc_num_buckets.argtypes = u64, [u64, u64]
        ns["hash_int"], ns["num_buckets"] = lib.orc_hash_int, lib.orc_num_buckets
    _libs[impl] = ns
    return ns


class OracleError(RuntimeError):
    pass


# --------------------------------------------------------------------------------------------------
# cells: one decoded column in numpy form
# --------------------------------------------------------------------------------------------------
class Cells:
    """valid: uint8[rows]; fixed types: values int32/int64/float64[rows];
    VARCHAR: str_off uint64[rows+1] + chars uint8[total]."""

    def __init__(self, type, valid, values=None, str_off=None, chars=None):
        self.type = int(type)
        self.valid = np.ascontiguousarray(valid, dtype=np.uint8)
        self.rows = self.valid.shape[0]
        if self.type == VARCHAR:
            self.str_off = np.ascontiguousarray(str_off, dtype=np.uint64)
            self.chars = np.ascontiguousarray(chars, dtype=np.uint8)
            self.values = None
        else:
            dt = {INT32: np.int32, INT64: np.int64, FP64: np.float64}[self.type]
            self.values = np.ascontiguousarray(values, dtype=dt)
            self.str_off = self.chars = None

    @staticmethod
    def from_strings(strings):
        """strings: list of bytes/str/None"""
        valid = np.array([s is not None for s in strings], dtype=np.uint8)
        bs = [(s.encode() if isinstance(s, str) else s) if s is not None else b"" for s in strings]
        off = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        chars = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
        return Cells(VARCHAR, valid, str_off=off, chars=chars)

    @staticmethod
    def from_values(type, values, valid=None):
        values = np.asarray(values)
        if valid is None:
            valid = np.ones(values.shape[0], dtype=np.uint8)
        return Cells(type, valid, values=values)

    def bits(self):
        """fixed types: the value bit pattern as uint64 (NULL rows forced to 0)"""
        if self.type == INT32:
            b = self.values.view(np.uint32).astype(np.uint64)
        else:
            b = self.values.view(np.uint64).copy()
        b[self.valid == 0] = 0
        return b

    def struct(self):
        c = orc_cells_t()
        c.type, c.rows = self.type, self.rows
        c.valid = self.valid.ctypes.data
        if self.type == VARCHAR:
            c.str_off, c.chars = self.str_off.ctypes.data, self.chars.ctypes.data
        else:
            c.values = self.values.ctypes.data
        return c

    def to_python(self):
        out = []
        for i in range(self.rows):
            if not self.valid[i]:
                out.append(None)
            elif self.type == VARCHAR:
                out.append(bytes(self.chars[int(self.str_off[i]):int(self.str_off[i + 1])]))
            elif self.type == FP64:
                out.append(float(self.values[i]))
            else:
                out.append(int(self.values[i]))
        return out


def _result_to_columns(ns, h):
    n_rows = int(ns["result_num_rows"](h))
    cols = []
    for c in range(int(ns["result_num_columns"](h))):
        n_pages = int(ns["result_column_pages"](h, c))
        pages = np.zeros((n_pages, PAGE_SIZE), dtype=np.uint8)
        if n_pages:
            C.memmove(pages.ctypes.data, ns["result_column_data"](h, c), n_pages * PAGE_SIZE)
        cols.append(rj.Column(int(ns["result_column_type"](h, c)), pages))
    return rj.ColumnarTable(num_rows=n_rows, columns=cols)


def encode(cells, impl="port"):
    """Cells -> ColumnarTable through the checker's Table::to_columnar (build_table.cpp:456-681)."""
    ns = _load(impl)
    n_rows = cells[0].rows if cells else 0
    arr = (orc_cells_t * max(len(cells), 1))(*[c.struct() for c in cells])
    h, err = C.c_void_p(), C.create_string_buffer(256)
    if ns["encode"](arr, len(cells), n_rows, C.byref(h), err, 256) != 0:
        raise OracleError(err.value.decode())
    try:
        return _result_to_columns(ns, h)
    finally:
        ns["result_free"](h)


def _column_struct(col):
    s = _cabi.rj_column_t()
    s.type, s.n_pages = int(col.type), col.n_pages
    s.contiguous = col.pages.ctypes.data if col.n_pages else None
    return s


def decode(col, num_rows, impl="port"):
    """Column -> Cells through the checker's Table::from_columnar (build_table.cpp:312-436)."""
    ns = _load(impl)
    s, err, n_chars = _column_struct(col), C.create_string_buffer(256), C.c_uint64()
    if ns["decode_sizes"](C.byref(s), num_rows, C.byref(n_chars), err, 256) != 0:
        raise OracleError(err.value.decode())
    valid = np.zeros(num_rows, dtype=np.uint8)
    t = int(col.type)
    if t == VARCHAR:
        off = np.zeros(num_rows + 1, dtype=np.uint64)
        chars = np.zeros(max(int(n_chars.value), 1), dtype=np.uint8)
        rc = ns["decode_fill"](C.byref(s), num_rows, valid.ctypes.data, None, off.ctypes.data,
                               chars.ctypes.data, err, 256)
        cells = Cells(t, valid, str_off=off, chars=chars[:int(n_chars.value)])
    else:
        values = np.zeros(num_rows, dtype={INT32: np.int32, INT64: np.int64, FP64: np.float64}[t])
        rc = ns["decode_fill"](C.byref(s), num_rows, valid.ctypes.data, values.ctypes.data, None, None,
                               err, 256)
        cells = Cells(t, valid, values=values)
    if rc != 0:
        raise OracleError(err.value.decode())
    return cells


def decode_table(table, impl="port"):
    return [decode(c, table.num_rows, impl) for c in table.columns]


def execute(plan, impl="port", n_threads=0):
    """Run Contest::execute of the checker on a radix_join_b200.Plan; returns a ColumnarTable."""
    ns = _load(impl)
    flat = rj.FlatPlan(plan)
    h, err = C.c_void_p(), C.create_string_buffer(256)
    if ns["execute"](flat.pointer(), n_threads, C.byref(h), err, 256) != 0:
        raise OracleError(err.value.decode())
    try:
        return _result_to_columns(ns, h)
    finally:
        ns["result_free"](h)


def last_execute_seconds():
    return float(_load("ref")["last_execute_seconds"]())


# --------------------------------------------------------------------------------------------------
# canonical form: exact multiset comparison of results whose row order is free
# --------------------------------------------------------------------------------------------------
_P = np.uint64(0x100000001B3)


def _string_keys(cells):
    """64-bit polynomial hash per string (sort key only; equality is checked on the bytes)."""
    lens = (cells.str_off[1:] - cells.str_off[:-1]).astype(np.int64)
    total = int(lens.sum())
    keys = lens.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    if total == 0:
        return keys
    starts = cells.str_off[:-1].astype(np.int64)
    dense = np.zeros(len(lens), dtype=np.int64)  # start of each string in the densely packed order
    np.cumsum(lens[:-1], out=dense[1:])
    pos = np.arange(total, dtype=np.int64) - np.repeat(dense, lens)  # position inside its string
    src = np.repeat(starts, lens) + pos
    maxlen = int(lens.max())
    with np.errstate(over="ignore"):
        powers = np.cumprod(np.full(maxlen + 1, _P, dtype=np.uint64))
        contrib = (cells.chars[src].astype(np.uint64) + np.uint64(1)) * powers[pos]
        nz = lens > 0
        sums = np.add.reduceat(contrib, dense[nz])
        keys[nz] = keys[nz] + sums
    return keys


def _gather_strings(cells, perm):
    lens = (cells.str_off[1:] - cells.str_off[:-1]).astype(np.int64)[perm]
    new_off = np.zeros(len(perm) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    src0 = cells.str_off[:-1].astype(np.int64)[perm]
    idx = np.repeat(src0 - new_off[:-1], lens) + np.arange(int(new_off[-1]), dtype=np.int64)
    return new_off.astype(np.uint64), cells.chars[idx]


def canonical(cells_list):
    """Sort rows by content.  Returns a list of per-column tuples of numpy arrays that are equal
    (np.array_equal) for two tables iff the tables are equal as multisets of rows.
    FP64 is compared by bit pattern; NULL differs from every value, including "" and 0."""
    if not cells_list:
        return []
    n = cells_list[0].rows
    keys = []
    for c in cells_list:
        assert c.rows == n
        k = _string_keys(c) if c.type == VARCHAR else c.bits()
        k = k.copy()
        k[c.valid == 0] = 0
        keys.append(c.valid)
        keys.append(k)
    perm = np.lexsort(keys[::-1]) if n else np.zeros(0, dtype=np.int64)
    out = []
    for c in cells_list:
        if c.type == VARCHAR:
            off, chars = _gather_strings(c, perm)
            out.append((c.valid[perm], off, chars))
        else:
            out.append((c.valid[perm], c.bits()[perm]))
    return out


def tables_equal(a_cells, b_cells):
    """exact multiset equality of two decoded tables"""
    if len(a_cells) != len(b_cells):
        return False
    if a_cells and a_cells[0].rows != b_cells[0].rows:
        return False
    if any(x.type != y.type for x, y in zip(a_cells, b_cells)):
        return False
    ca, cb = canonical(a_cells), canonical(b_cells)
    return all(all(np.array_equal(u, v) for u, v in zip(x, y)) for x, y in zip(ca, cb))


def result_equal(a_table, b_table, impl="port"):
    """ColumnarTable vs ColumnarTable (num_rows, column count, column types, multiset of rows) --
    the checks of the reference harness, tests/read_sql.cpp:1159-1222."""
    if a_table.num_rows != b_table.num_rows or len(a_table.columns) != len(b_table.columns):
        return False
    if any(int(x.type) != int(y.type) for x, y in zip(a_table.columns, b_table.columns)):
        return False
    return tables_equal(decode_table(a_table, impl), decode_table(b_table, impl))
