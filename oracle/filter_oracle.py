"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's pre-filter evaluation.

  * `selected(cells, stmt)`  numpy / Python restatement of the COLUMN-WISE evaluation the reference runs in
    Table::from_csv (src/build_table.cpp:255): Comparison::eval over InnerColumns
    (src/statement.cpp:46-133, include/inner_column.h:170-325 and :386-562), LogicalOperation::eval
    (src/statement.cpp:186-200), bitmap_not / and / or (:8-44) and Comparison::like_match
    (include/statement.h:118-161).  One bool per row.
  * `selected_ref(table, stmt)`  the unmodified reference itself (oracle/_ref/libref_oracle.so: ref_filter
    in oracle/ref_shim.cpp builds the InnerColumns and calls Statement::eval) -- used to pin the restatement.
  * `emit(cells, mask)`  the rows that pass, in row order (from_inner_to_column, build_table.cpp:94-119).

Only tests/ may import this module; the product evaluates filters in csrc/k_filter.cu.
"""
import ctypes as C
import re

import numpy as np

from radix_join_b200 import _cabi
from radix_join_b200.plan import DataType, FlatPlan, Plan
from radix_join_b200.statement import Comparison, LogicalOperation

from . import pyoracle as orc

_SPECIAL = set("\\.^$|?*+()[]{}")


def like_regex(pattern: str):
    """include/statement.h:131-146: '%' -> '.*', '_' -> '.', the listed specials escaped; RE2::FullMatch with
    default options ('.' does not match a newline)"""
    out = []
    for ch in pattern:
        if ch == "%":
            out.append(".*")
        elif ch == "_":
            out.append(".")
        else:
            out.append("\\" + ch if ch in _SPECIAL else ch)
    return re.compile("".join(out))


def _compare(op, lhs, rhs):
    C_ = Comparison
    return {C_.EQ: lambda: lhs == rhs, C_.NEQ: lambda: lhs != rhs, C_.LT: lambda: lhs < rhs, C_.GT: lambda: lhs > rhs,
            C_.LEQ: lambda: lhs <= rhs, C_.GEQ: lambda: lhs >= rhs}[op]()


def selected(cells, stmt):
    """cells: list of oracle Cells (one per column); -> bool array, one entry per row"""
    n = len(cells[0].valid)
    if isinstance(stmt, LogicalOperation):
        a = selected(cells, stmt.children[0])
        if stmt.op_type == LogicalOperation.NOT:
            return ~a  # bitmap_not: a NULL row's false becomes true (statement.cpp:8-16)
        b = selected(cells, stmt.children[1])
        return (a & b) if stmt.op_type == LogicalOperation.AND else (a | b)
    col = cells[stmt.column]
    valid = col.valid.astype(bool)
    if stmt.op == Comparison.IS_NULL:
        return ~valid
    if stmt.op == Comparison.IS_NOT_NULL:
        return valid
    t = DataType(col.type)
    if t == DataType.VARCHAR:
        if not isinstance(stmt.value, (str, bytes)):
            raise TypeError("bad_variant_access")  # std::get<std::string> (statement.cpp:112)
        rhs = stmt.value.encode() if isinstance(stmt.value, str) else bytes(stmt.value)
        strings = col.to_python()  # bytes or None per row
        out = np.zeros(n, dtype=bool)
        if stmt.op in (Comparison.LIKE, Comparison.NOT_LIKE):
            rx = like_regex(rhs.decode("utf-8"))
            for i, s in enumerate(strings):
                if s is not None:
                    m = rx.fullmatch(s.decode("utf-8")) is not None
                    out[i] = m if stmt.op == Comparison.LIKE else not m
        else:
            for i, s in enumerate(strings):
                if s is not None:
                    out[i] = bool(_compare(stmt.op, s, rhs))  # bytes compare like std::string_view (unsigned chars)
        return out
    if t == DataType.FP64:
        if not isinstance(stmt.value, float):
            raise TypeError("bad_variant_access")  # std::get<double> (statement.cpp:93)
        rhs = np.float64(stmt.value)
        vals = col.values.view(np.float64)
    else:
        if not isinstance(stmt.value, int) or isinstance(stmt.value, bool):
            raise TypeError("bad_variant_access")  # std::get<int64_t> (statement.cpp:55,74)
        if t == DataType.INT32:
            rhs = np.int64(stmt.value).astype(np.int32)  # static_cast<int32_t> (statement.cpp:55)
            vals = col.values.view(np.int32)
        else:
            rhs = np.int64(stmt.value)
            vals = col.values.view(np.int64)
    with np.errstate(invalid="ignore"):
        return valid & _compare(stmt.op, vals, rhs)


def selected_ref(table, stmt):
    """the unmodified reference's column-wise eval on the decoded table"""
    lib = C.CDLL(orc.REF_LIB)
    plan = Plan()
    plan.new_input(table)
    flat = FlatPlan(plan)
    entries = []
    stmt.program(entries)
    prog = (_cabi.rj_pred_t * max(len(entries), 1))()
    for i, (e, _k) in enumerate(entries):
        prog[i] = e
    out = np.zeros(max(table.num_rows, 1), dtype=np.uint8)
    err = C.create_string_buffer(512)
    lib.ref_filter.restype = C.c_int
    lib.ref_filter.argtypes = [C.POINTER(_cabi.rj_table_t), C.POINTER(_cabi.rj_pred_t), C.c_uint32, C.c_void_p, C.c_char_p, C.c_size_t]
    rc = lib.ref_filter(flat.tables, prog, len(entries), out.ctypes.data, err, 512)
    if rc != 0:
        raise orc.OracleError(err.value.decode())
    return out[:table.num_rows].astype(bool)


def emit(cells, mask):
    """rows with mask set, in row order, as python lists per column (None = NULL)"""
    idx = np.nonzero(mask)[0]
    return [[c.to_python()[i] for i in idx] for c in cells]
