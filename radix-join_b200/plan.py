"""Host-side mirror of the reference's plan API (include/plan.h:27-149, include/attribute.h:8-13).

Same names, same argument meaning, same error behaviour as the C++ types a contest harness builds:
`Plan.new_scan_node / new_join_node / new_input`, `ColumnarTable{num_rows, columns}`,
`Column{type, pages}`, `DataType`.  Pages are rows of a `(n_pages, 8192)` uint8 numpy array (host
memory); nothing here touches the data, the engine below the C-ABI does all the work.
"""
import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple, Union

import numpy as np

from . import _cabi
from ._cabi import PAGE_SIZE


class DataType(enum.IntEnum):
    """include/attribute.h:8-13"""
    INT32 = 0
    INT64 = 1
    FP64 = 2
    VARCHAR = 3


@dataclass
class ScanNode:
    """include/plan.h:32-34"""
    base_table_id: int


@dataclass
class JoinNode:
    """include/plan.h:36-42"""
    build_left: bool
    left: int
    right: int
    left_attr: int
    right_attr: int


@dataclass
class PlanNode:
    """include/plan.h:44-52: output_attrs = [(source index, DataType)]"""
    data: Union[ScanNode, JoinNode]
    output_attrs: List[Tuple[int, DataType]]


class Column:
    """include/plan.h:60-100: a typed page list.  `pages` is a (n_pages, 8192) uint8 array."""

    def __init__(self, data_type, pages=None):
        self.type = DataType(data_type)
        if pages is None:
            pages = np.zeros((0, PAGE_SIZE), dtype=np.uint8)
        pages = np.ascontiguousarray(pages, dtype=np.uint8).reshape(-1, PAGE_SIZE)
        self.pages = pages

    @property
    def n_pages(self):
        return self.pages.shape[0]


@dataclass
class ColumnarTable:
    """include/plan.h:102-105"""
    num_rows: int = 0
    columns: List[Column] = field(default_factory=list)


class Plan:
    """include/plan.h:112-149"""

    def __init__(self):
        self.nodes: List[PlanNode] = []
        self.inputs: List[ColumnarTable] = []
        self.root = 0

    def new_join_node(self, build_left, left, right, left_attr, right_attr, output_attrs):
        self.nodes.append(PlanNode(JoinNode(bool(build_left), left, right, left_attr, right_attr),
                                   [(int(i), DataType(t)) for i, t in output_attrs]))
        return len(self.nodes) - 1

    def new_scan_node(self, base_table_id, output_attrs):
        self.nodes.append(PlanNode(ScanNode(base_table_id),
                                   [(int(i), DataType(t)) for i, t in output_attrs]))
        return len(self.nodes) - 1

    def new_input(self, table: ColumnarTable):
        self.inputs.append(table)
        return len(self.inputs) - 1


class FlatPlan:
    """A Plan flattened into the rj_plan_t of include/rj_b200.h.  Keeps every ctypes array and
    numpy buffer alive for as long as the object lives."""

    def __init__(self, plan: Plan, device_pages: Sequence[Sequence[int]] = None):
        self._keep = []
        n_nodes = len(plan.nodes)
        nodes = (_cabi.rj_node_t * max(n_nodes, 1))()
        for i, n in enumerate(plan.nodes):
            attrs = (_cabi.rj_attr_t * max(len(n.output_attrs), 1))()
            for a, (idx, t) in enumerate(n.output_attrs):
                attrs[a].index = idx
                attrs[a].type = int(t)
            self._keep.append(attrs)
            nodes[i].n_output_attrs = len(n.output_attrs)
            nodes[i].output_attrs = attrs
            if isinstance(n.data, JoinNode):
                nodes[i].is_join = 1
                nodes[i].build_left = 1 if n.data.build_left else 0
                nodes[i].left = n.data.left
                nodes[i].right = n.data.right
                nodes[i].left_attr = n.data.left_attr
                nodes[i].right_attr = n.data.right_attr
            else:
                nodes[i].is_join = 0
                nodes[i].base_table_id = n.data.base_table_id
        tables = (_cabi.rj_table_t * max(len(plan.inputs), 1))()
        for ti, t in enumerate(plan.inputs):
            cols = (_cabi.rj_column_t * max(len(t.columns), 1))()
            for ci, c in enumerate(t.columns):
                cols[ci].type = int(c.type)
                if device_pages is not None:
                    # columns already resident in HBM: (device address, n_pages)
                    addr, n_pages = device_pages[ti][ci]
                    cols[ci].n_pages = n_pages
                    cols[ci].contiguous = addr
                else:
                    cols[ci].n_pages = c.n_pages
                    cols[ci].contiguous = c.pages.ctypes.data if c.n_pages else None
                    self._keep.append(c.pages)
                cols[ci].pages = None
            self._keep.append(cols)
            tables[ti].num_rows = t.num_rows
            tables[ti].n_columns = len(t.columns)
            tables[ti].columns = cols
        self.nodes = nodes
        self.tables = tables
        self.n_tables = len(plan.inputs)
        self.c = _cabi.rj_plan_t()
        self.c.n_nodes = n_nodes
        self.c.n_inputs = len(plan.inputs)
        self.c.nodes = nodes
        self.c.inputs = tables
        self.c.root = plan.root

    def pointer(self):
        return C.byref(self.c)
