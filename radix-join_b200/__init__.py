"""radix-join_b200 -- B200-native radix hash-join engine behind the SIGMOD'25 contest API.

The product is the CUDA engine in csrc/ (librj_b200.so, C-ABI in include/rj_b200.h).  This package
is the Python host-side mirror of the reference's operator interface for the execute() path:

    ctx = build_context()               # Contest::build_context   (include/plan.h:339)
    out = execute(plan, ctx)            # Contest::execute         (include/plan.h:342)
    destroy_context(ctx)                # Contest::destroy_context (include/plan.h:340)

There is no CPU fallback: importing works without a GPU, but every call needs librj_b200.so and an
sm_100 device and fails loudly otherwise.
"""
from ._cabi import PAGE_SIZE, EngineMissing, load_library
from .engine import (Context, EngineError, ResidentInputs, Result, adopt_device, build_context,
                     destroy_context, execute, execute_pages, execute_resident, execute_streamed,
                     execute_streamed_columnar, execute_to_device, tables_equal, upload)
from .statement import Comparison, LogicalOperation, Statement, filter_table, filter_table_to_device
from .plan import (Column, ColumnarTable, DataType, FlatPlan, JoinNode, Plan, PlanNode, ScanNode)

__all__ = [
    "PAGE_SIZE", "EngineMissing", "load_library", "Context", "EngineError", "ResidentInputs", "Result",
    "adopt_device", "build_context", "destroy_context", "execute", "execute_pages", "execute_resident",
    "execute_streamed", "execute_streamed_columnar", "execute_to_device", "upload", "Column", "ColumnarTable", "DataType", "FlatPlan", "JoinNode",
    "Plan", "PlanNode", "ScanNode", "Comparison", "LogicalOperation", "Statement", "filter_table", "filter_table_to_device", "tables_equal",
]
