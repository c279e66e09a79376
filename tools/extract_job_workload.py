"""Derive the JOB workload definition used by configs 3-5 from the reference's workload files:
plans.json (PostgreSQL EXPLAIN trees), job/*.sql (select lists, aliases) and the IMDB schema in
tests/read_sql.cpp:21-139 (attributes_map).  Run in the CPU container (needs /root/reference):

    python tools/extract_job_workload.py      ->  radix-join_b200/job/job_workload.json

The output holds shapes only (tree structure, join columns, estimated cardinalities, column types);
it is what radix_join_b200.job turns into Plans the way the contest harness does
(tests/read_sql.cpp:861-1141, load_join_pipeline)."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "radix-join_b200", "job", "job_workload.json")


def schema():
    src = open(os.path.join(REF, "tests", "read_sql.cpp")).read()
    body = src[src.index("attributes_map = {"):src.index("};", src.index("attributes_map = {"))]
    out = {}
    for m in re.finditer(r'\{"(\w+)",\s*\{((?:\s*\{DataType::\w+,\s*"\w+"\},?)+)\}', body):
        out[m.group(1)] = [[c, t] for t, c in re.findall(r'DataType::(\w+),\s*"(\w+)"', m.group(2))]
    return out


def not_null_columns():
    """job/schema.sql: columns declared NOT NULL (everything else gets NULLs in the synthetic data)"""
    src = open(os.path.join(REF, "job", "schema.sql")).read()
    out = {}
    for m in re.finditer(r"CREATE TABLE (\w+) \((.*?)\);", src, flags=re.S):
        out[m.group(1)] = [l.split()[0] for l in m.group(2).split(",\n") if "NOT NULL" in l]
    return out


def main():
    plans = json.load(open(os.path.join(REF, "plans.json")))
    sch = schema()
    assert len(sch) == 21, len(sch)
    table_rows = {}
    queries = {}

    def convert(n, aliases):
        t = n["Node Type"]
        if t in ("Aggregate", "Gather"):
            return convert(n["Plans"][0], aliases)
        rows = n["Plan Rows"] * (2.4 if n.get("Parallel Aware") else 1.0)
        if t in ("Seq Scan", "Index Only Scan"):
            alias = n.get("Alias") or n["Relation Name"]
            table = n["Relation Name"]
            aliases[alias] = table
            if "Filter" not in n and "Index Cond" not in n:
                table_rows[table] = max(table_rows.get(table, 0), int(round(rows)))
            return {"scan": alias, "table": table, "rows": int(round(rows)), "filtered": "Filter" in n or "Index Cond" in n}
        if t != "Hash Join":
            raise SystemExit(f"unsupported node {t}")
        l, r = n["Plans"]
        # the Hash child is the build side (tests/read_sql.cpp:943-953)
        if l["Node Type"] == "Hash" and r["Node Type"] != "Hash":
            build_left, pl, pr = True, l["Plans"][0], r
        elif r["Node Type"] == "Hash" and l["Node Type"] != "Hash":
            build_left, pl, pr = False, l, r["Plans"][0]
        else:
            raise SystemExit("Hash Join without exactly one Hash child")
        m = re.fullmatch(r"\((\w+)\.(\w+) = (\w+)\.(\w+)\)", n["Hash Cond"])
        return {"join": [convert(pl, aliases), convert(pr, aliases)], "build_left": build_left,
                "cond": [[m.group(1), m.group(2)], [m.group(3), m.group(4)]], "rows": int(round(rows))}

    for name, plan in zip(plans["names"], plans["plans"]):
        sql = open(os.path.join(REF, plans["sql_directory"], name + ".sql")).read()
        aliases = {}
        tree = convert(plan["Plan"], aliases)
        outputs = re.findall(r"MIN\((\w+)\.(\w+)\)", sql, flags=re.I)
        assert outputs, name
        queries[name] = {"outputs": [list(o) for o in outputs], "aliases": aliases, "tree": tree}
    table_rows.setdefault("kind_type", 7)        # never scanned unfiltered
    table_rows.setdefault("comp_cast_type", 4)
    with open(OUT, "w") as f:
        json.dump({"schema": sch, "not_null": not_null_columns(), "table_rows": table_rows, "queries": queries}, f, separators=(",", ":"))
    n_join = sum(json.dumps(q["tree"]).count('"join"') for q in queries.values())
    n_scan = sum(json.dumps(q["tree"]).count('"scan"') for q in queries.values())
    print(f"{len(queries)} queries, {n_join} joins, {n_scan} scans, {len(sch)} tables ->", os.path.getsize(OUT), "bytes")
    print(sorted(table_rows.items(), key=lambda kv: -kv[1]))


if __name__ == "__main__":
    main()
