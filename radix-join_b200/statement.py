"""Host-side mirror of the reference's filter AST (include/statement.h:37-245) and of the filter + emit
tail of Table::from_csv (src/build_table.cpp:247-303), evaluated by the CUDA engine.

    stmt = LogicalOperation.makeAnd(Comparison(2, Comparison.GT, 2000), Comparison(1, Comparison.LIKE, "%(co-production)%"))
    filtered = filter_table(table, stmt, ctx)      # ColumnarTable of the rows that pass, in row order

Same names, argument meaning and error behaviour as the C++ classes: `Comparison(column, op, value)`
with the operators of `Comparison::Op`, `LogicalOperation` with `makeAnd / makeOr / makeNot`; a literal of
the wrong kind for the column raises like `std::get` throws (src/statement.cpp:55,74,93,112).  NOT flips
NULL rows to true exactly like `bitmap_not` (src/statement.cpp:8-16).  There is no CPU path here: the
predicates run in csrc/k_filter.cu.
"""
import ctypes as C

from . import _cabi
from .engine import Context, Result
from .plan import ColumnarTable, DataType, Plan


class Statement:
    """include/statement.h:37-43"""

    def program(self, out):
        raise NotImplementedError


class Comparison(Statement):
    """include/statement.h:45-116: `column` indexes the table's columns"""
    EQ, NEQ, LT, GT, LEQ, GEQ, LIKE, NOT_LIKE, IS_NULL, IS_NOT_NULL = range(10)

    def __init__(self, column, op, value=None):
        self.column, self.op, self.value = int(column), int(op), value

    def program(self, out):
        e = _cabi.rj_pred_t()
        e.kind, e.op, e.column, e.lit_type = 0, self.op, self.column, -1
        keep = None
        v = self.value
        if isinstance(v, bool):
            raise TypeError("a Literal is int64, double, string or monostate (include/statement.h:14)")
        if isinstance(v, int):
            e.lit_type, e.rhs_i = int(DataType.INT64), v
        elif isinstance(v, float):
            e.lit_type, e.rhs_d = int(DataType.FP64), v
        elif isinstance(v, (str, bytes)):
            keep = v.encode() if isinstance(v, str) else bytes(v)
            e.lit_type, e.rhs_s, e.rhs_s_len = int(DataType.VARCHAR), keep, len(keep)
        elif v is not None:
            raise TypeError("a Literal is int64, double, string or monostate (include/statement.h:14)")
        out.append((e, keep))


class LogicalOperation(Statement):
    """include/statement.h:185-245"""
    AND, OR, NOT = range(3)

    def __init__(self, op_type, children):
        self.op_type, self.children = int(op_type), list(children)

    @staticmethod
    def makeAnd(l, r):
        return LogicalOperation(LogicalOperation.AND, [l, r])

    @staticmethod
    def makeOr(l, r):
        return LogicalOperation(LogicalOperation.OR, [l, r])

    @staticmethod
    def makeNot(child):
        return LogicalOperation(LogicalOperation.NOT, [child])

    def program(self, out):
        # the table-wise eval reads children[0] and children[1] only (src/statement.cpp:186-200)
        need = 1 if self.op_type == LogicalOperation.NOT else 2
        if len(self.children) < need:
            raise ValueError("LogicalOperation: missing operand")
        for ch in self.children[:need]:
            ch.program(out)
        e = _cabi.rj_pred_t()
        e.kind, e.op, e.lit_type = 1, self.op_type, -1
        out.append((e, None))


def filter_table_to_device(table: ColumnarTable, stmt, ctx: Context) -> Result:
    """rows of `table` (host pages, all columns) that satisfy `stmt`, as result pages left in HBM"""
    plan = Plan()
    plan.new_input(table)
    from .plan import FlatPlan
    flat = FlatPlan(plan)
    entries = []
    if stmt is not None:
        stmt.program(entries)
    prog = (_cabi.rj_pred_t * max(len(entries), 1))()
    for i, (e, _keep) in enumerate(entries):
        prog[i] = e
    h = C.c_void_p()
    ctx.check(ctx.lib.rj_filter_table(ctx.handle, flat.tables, prog, len(entries), C.byref(h)))
    del entries  # (the literals stayed alive until here)
    return Result(ctx, h)


def filter_table(table: ColumnarTable, stmt, ctx: Context) -> ColumnarTable:
    """The filter + emit step of Table::from_csv (src/build_table.cpp:247-303): host pages in, host pages out."""
    res = filter_table_to_device(table, stmt, ctx)
    try:
        return res.to_columnar()
    finally:
        res.free()
