"""SQL statement -> operator plan.  Same entry point and plan keys as the reference's parse.py.

Mirrors /root/reference/parse.py: ``getIndex`` (:9-13) and ``sql_parse(tables, sql_statement)``
(:16-91).  For the two statement forms the reference implements, the returned dict has exactly its
keys — ``{"table", "select"}`` (:58) and ``{"select", "groupbys", "table", "g_col"}`` (:90) — with the
same column-index lists and aggregate codes (``prod 1, sum 2, max 3, min 4``, :81; the group column
itself contributes ``(g_col, 0)``, :73-75), and the same exception messages (:33,54,69,78,87).

Extensions (clauses the reference's README lists but parse.py ignores): ``where``
(AND / OR / NOT over column-vs-constant comparisons, BETWEEN, IN — handed to libhark in conjunctive
normal form), ``count``/``avg`` (codes 5, 6), ``having``,
``orderby``, ``join``, ``limit``, GROUP BY over several columns (``g_cols``), aggregates without GROUP BY (``global``), ``select *`` and the single-column select that crashes the
reference (:48-51 iterates a dict).  They appear as extra plan keys; a plan without them is
byte-for-byte what the reference would build.

``moz_sql_parser`` is not installed here; ``sqlmini.parse`` emits the same dict shapes.
"""

from .sqlmini import parse

FUNC_TO_FUT = {"prod": 1, "sum": 2, "max": 3, "min": 4}          # parse.py:81
FUNC_TO_FUT_EXT = {**FUNC_TO_FUT, "count": 5, "avg": 6}
CMP_TO_CODE = {"gt": 0, "gte": 1, "lt": 2, "lte": 3, "eq": 4, "neq": 5}
_FLIP = {"gt": "lt", "gte": "lte", "lt": "gt", "lte": "gte", "eq": "eq", "neq": "neq"}


def getIndex(elements, value):
    for i, v in enumerate(elements):
        if v == value:
            return i
    return -1


def _as_list(x):
    return x if isinstance(x, list) else [x]


def _strip_qualifier(name, aliases):
    """`f.col` -> (table_key, col); bare `col` -> (None, col)."""
    if isinstance(name, str) and "." in name:
        q, c = name.split(".", 1)
        return aliases.get(q, q), c
    return None, name


PRED_OP_MASK, PRED_OR, PRED_NOT = 0xFF, 0x100, 0x200       # include/hark.h HARK_PRED_*
MAX_PREDS = 16                                             # libhark's limit on one predicate list


def _leaf(expr, resolve):
    """{"gt": ["a", 4]} -> (col_index, op_code, ival, fval).  `resolve(operand) -> column index`."""
    (op, args), = expr.items()
    if op not in CMP_TO_CODE or not isinstance(args, list) or len(args) != 2:
        raise Exception(f"unsupported predicate operator {op} (column-vs-constant comparisons, AND / OR / NOT, "
                        f"BETWEEN and IN are supported)")
    lhs, rhs = args
    if isinstance(lhs, (int, float)) and not isinstance(rhs, (int, float)):
        lhs, rhs, op = rhs, lhs, _FLIP[op]
    if not isinstance(rhs, (int, float)):
        raise Exception(f"unsupported predicate {expr}: right-hand side must be a numeric constant")
    idx = resolve(lhs)
    ival = int(rhs) if float(rhs) == int(rhs) else None
    return (idx, CMP_TO_CODE[op], ival, float(rhs))


def _cnf(expr, resolve, neg=False):
    """Boolean expression -> conjunctive normal form: a list of clauses, each a list of leaf predicates
    (col, op | PRED_NOT?, ival, fval).  NOT is pushed down to the leaves (De Morgan) and stays a flag on the
    comparison, so NOT (x > c) keeps rows where x is NaN; OR over ANDs is distributed."""
    if not isinstance(expr, dict) or len(expr) != 1:
        raise Exception(f"unsupported predicate {expr}")
    (op, args), = expr.items()
    if op == "not":
        return _cnf(args, resolve, not neg)
    if op in ("and", "or"):
        parts = [_cnf(e, resolve, neg) for e in args]
        if (op == "and") != neg:                      # conjunction: concatenate the clause lists
            return [cl for p in parts for cl in p]
        out = [[]]                                    # disjunction: one clause per choice of a clause from each part
        for p in parts:
            out = [a + b for a in out for b in p]
            if sum(len(cl) for cl in out) > 4 * MAX_PREDS:
                raise Exception("WHERE / HAVING expression is too large once in conjunctive normal form")
        return out
    if op in ("between", "not_between"):
        x, lo, hi = args
        e = {"and": [{"gte": [x, lo]}, {"lte": [x, hi]}]}
        return _cnf(e, resolve, neg != (op == "not_between"))
    if op in ("in", "nin"):
        x, vals = args
        e = {"or": [{"eq": [x, v]} for v in _as_list(vals)]}
        return _cnf(e, resolve, neg != (op == "nin"))
    c, code, ival, fval = _leaf(expr, resolve)
    return [[(c, code | (PRED_NOT if neg else 0), ival, fval)]]


def _preds(expr, resolve):
    """WHERE / HAVING expression -> libhark predicate list (conjunctive normal form, hark.h hark_pred): clauses are
    AND-ed, a predicate carrying PRED_OR is OR-ed with the next one.  A plain conjunction comes out exactly as the
    flag-free list of its comparisons."""
    out = []
    for clause in _cnf(expr, resolve):
        uniq = list(dict.fromkeys(clause))
        for j, (c, code, ival, fval) in enumerate(uniq):
            out.append((c, code | (PRED_OR if j + 1 < len(uniq) else 0), ival, fval))
    if len(out) > MAX_PREDS:
        raise Exception(f"WHERE / HAVING needs {len(out)} comparisons in conjunctive normal form; at most {MAX_PREDS}")
    return out


def finalize_pred(pred, is_int_column):
    """Integer columns compare in int64: fold a fractional constant into an equivalent integer one."""
    import math
    col, code, ival, fval = pred
    op, flags = code & PRED_OP_MASK, code & ~PRED_OP_MASK
    if not is_int_column:
        return (col, code, 0 if ival is None else ival, fval)
    if ival is not None:
        return (col, code, ival, fval)
    # x > 2.5 <=> x > 2 ; x >= 2.5 <=> x > 2 ; x < 2.5 <=> x < 3 ; x <= 2.5 <=> x < 3 ; = never ; != always
    fl = math.floor(fval)
    if op in (0, 1):
        return (col, 0 | flags, fl, fval)
    if op in (2, 3):
        return (col, 2 | flags, fl + 1, fval)
    if op == 4:
        return (col, 2 | flags, -(2 ** 63), fval)      # x < INT64_MIN: never true
    return (col, 1 | flags, -(2 ** 63), fval)          # x >= INT64_MIN: always true


def sql_parse(tables, sql_statement):
    """Parses an SQL statement into the plan dict FutharkContext.sql dispatches on."""
    js_obj = parse(sql_statement)

    # ---- FROM (parse.py:28-33), extended with JOIN ----
    from_items = _as_list(js_obj["from"])
    aliases = {}

    def table_of(item):
        name = item["value"] if isinstance(item, dict) and "value" in item else item
        if isinstance(item, dict) and "name" in item:
            aliases[item["name"]] = name
        if name in tables:
            return name, tables[name]
        raise Exception(f"{name} is not in tables")

    table_name, table = table_of(from_items[0])
    columns = table.get_schema()
    plan_join = None
    if len(from_items) > 1:
        if len(from_items) > 2:
            raise Exception("only a single JOIN is supported")
        j = from_items[1]
        jkey = "inner join" if "inner join" in j else "join"
        name2, table2 = table_of(j[jkey])
        on = j["on"]
        if not (isinstance(on, dict) and "eq" in on):
            raise Exception("JOIN needs an equality ON condition")
        plan_join = {"name2": name2, "table2": table2, "on": on["eq"]}

    def col_index(col_name, tname=table_name, cols=columns):
        idx = getIndex(cols, col_name)
        if idx < 0:
            raise Exception(f"{col_name} is not in the schema of table {tname}")
        return idx

    if plan_join is not None:
        return _plan_join(js_obj, table_name, table, plan_join, aliases)

    def resolve_plain(operand):
        _, c = _strip_qualifier(operand, aliases)
        if not isinstance(c, str):
            raise Exception(f"unsupported predicate operand {operand}")
        return col_index(c)

    where = _preds(js_obj["where"], resolve_plain) if "where" in js_obj else []
    orderby = _as_list(js_obj["orderby"]) if "orderby" in js_obj else []
    extras = {}
    if where:
        extras["where"] = where
    if "limit" in js_obj:
        extras["limit"] = int(js_obj["limit"])

    # ---- SELECT DISTINCT a, b: GROUP BY a, b whose (COUNT) aggregate is dropped; rows come out in key order ----
    if "select_distinct" in js_obj:
        if "groupby" in js_obj or "having" in js_obj:
            raise Exception("SELECT DISTINCT together with GROUP BY / HAVING is not supported")
        sd = js_obj["select_distinct"]
        names = list(columns) if sd == "*" else [p["value"] for p in _as_list(sd)]
        if not all(isinstance(nm, str) for nm in names):
            raise Exception("SELECT DISTINCT takes plain columns")
        g_cols = [col_index(_strip_qualifier(nm, aliases)[1]) for nm in names]
        if len(set(g_cols)) != len(g_cols):
            raise Exception("SELECT DISTINCT lists a column twice")
        if orderby:
            keys = []
            for k in orderby:
                idx = col_index(_strip_qualifier(k["value"], aliases)[1])
                if idx not in g_cols:
                    raise Exception(f"{k['value']} must appear in the select list to be used in ORDER BY")
                keys.append((g_cols.index(idx), 1 if k.get("sort") == "desc" else 0))
            extras["orderby"] = keys
        return {"table": table.get_handle(), "select": [g_cols[0]], "groupbys": [FUNC_TO_FUT_EXT["count"]],
                "g_cols": g_cols, "distinct": True, **extras}

    # ---- plain SELECT (parse.py:42-58) ----
    if "groupby" not in js_obj.keys():
        pairs = _as_list(js_obj["select"]) if js_obj["select"] != "*" else []
        if pairs and all(isinstance(p.get("value"), dict) for p in pairs):
            # SELECT sum(a), count(*) ... with no GROUP BY: one group holding every (filtered) row
            sel, codes = [], []
            for p in pairs:
                for agg_func, agg_val in FUNC_TO_FUT_EXT.items():
                    if agg_func in p["value"]:
                        arg = p["value"][agg_func]
                        sel.append(0 if arg == "*" else col_index(_strip_qualifier(arg, aliases)[1]))
                        codes.append(agg_val)
            if len(sel) != len(pairs):
                raise Exception(f"unsupported aggregate in {js_obj['select']}")
            if "having" in js_obj or orderby:
                raise Exception("HAVING / ORDER BY need a GROUP BY clause")
            return {"table": table.get_handle(), "select": sel, "groupbys": codes, "global": True, **extras}
        fut_cols_selects = []
        if js_obj["select"] == "*":
            fut_cols_selects = list(range(len(columns)))
        else:
            for pair in _as_list(js_obj["select"]):
                if "value" in pair:
                    col_name = pair["value"]
                    if not isinstance(col_name, str):
                        raise Exception(f"{col_name} needs a GROUP BY clause")
                    fut_cols_selects += [col_index(_strip_qualifier(col_name, aliases)[1])]
        if orderby:
            extras["orderby"] = [(col_index(_strip_qualifier(k["value"], aliases)[1]),
                                  1 if k.get("sort") == "desc" else 0) for k in orderby]
        return {"table": table.get_handle(), "select": fut_cols_selects, **extras}

    # ---- GROUP BY (parse.py:60-90) ----
    fut_cols_selects = []
    typ_cols_selects = []
    gb = js_obj["groupby"]
    if isinstance(gb, list):
        return _plan_groupby_multi(js_obj, gb, table_name, table, columns, aliases, extras, orderby)
    g_col_name = _strip_qualifier(gb["value"], aliases)[1]
    g_col = getIndex(columns, g_col_name)
    if g_col < 0:
        raise Exception(f"{g_col_name} is not in the schema of table {table_name}")

    out_names = []      # output column descriptors, for HAVING / ORDER BY resolution
    select_pairs = _as_list(js_obj["select"]) if js_obj["select"] != "*" else [{"value": g_col_name}]
    for dic in select_pairs:
        val = dic["value"]
        if isinstance(val, str) and _strip_qualifier(val, aliases)[1] == g_col_name:
            fut_cols_selects += [g_col]
            typ_cols_selects += [0]
            out_names.append(g_col_name)
        elif isinstance(val, str):
            raise Exception(f"{val} is not an aggregation function or the columns thats grouped on")
        else:
            for agg_func, agg_val in FUNC_TO_FUT_EXT.items():
                if agg_func in val:
                    agg_col_name = val[agg_func]
                    if agg_col_name == "*":
                        agg_col = g_col
                    else:
                        agg_col_name = _strip_qualifier(agg_col_name, aliases)[1]
                        agg_col = getIndex(columns, agg_col_name)
                        if agg_col < 0:
                            raise Exception(f"{agg_col_name} is not in the schema of table {table_name}")
                    fut_cols_selects += [agg_col]
                    typ_cols_selects += [agg_val]
                    out_names.append((agg_func, agg_col))
        if "name" in dic and out_names:
            extras.setdefault("aliases", {})[dic["name"]] = len(out_names)      # output index (0 = key)

    def resolve_output(operand):
        """HAVING / ORDER BY operand -> output column index (0 = key, i = i-th select item)."""
        if isinstance(operand, dict):
            (f, c), = operand.items()
            c_idx = g_col if c == "*" else col_index(_strip_qualifier(c, aliases)[1])
            for i, nm in enumerate(out_names):
                if nm == (f, c_idx):
                    return i + 1
            raise Exception(f"{f}({c}) must appear in the select list to be used in HAVING / ORDER BY")
        _, c = _strip_qualifier(operand, aliases)
        if c in extras.get("aliases", {}):
            return extras["aliases"][c]
        if c == g_col_name:
            return 0
        raise Exception(f"{c} is not an output column of the GROUP BY")

    if "having" in js_obj:
        extras["having"] = _preds(js_obj["having"], resolve_output)
    if orderby:
        extras["orderby"] = [(resolve_output(k["value"]), 1 if k.get("sort") == "desc" else 0) for k in orderby]
    extras.pop("aliases", None)
    return {"select": fut_cols_selects, "groupbys": typ_cols_selects, "table": table.get_handle(), "g_col": g_col,
            **extras}


def _plan_groupby_multi(js_obj, gb, table_name, table, columns, aliases, extras, orderby):
    """GROUP BY a, b, ... — the several-columns case the reference leaves as a TODO (parse.py:64).  Plan keys:
    ``g_cols`` (instead of ``g_col``), ``select`` / ``groupbys`` as in the single-key plan (a selected group column
    carries code 0).  Output columns of the operator: the group columns, then one column per select item — the
    single-key layout (key, then the select items) with more keys."""
    def col_index(name):
        c = _strip_qualifier(name, aliases)[1]
        idx = getIndex(columns, c)
        if idx < 0:
            raise Exception(f"{c} is not in the schema of table {table_name}")
        return idx

    g_cols = [col_index(g["value"]) for g in gb]
    if len(set(g_cols)) != len(g_cols):
        raise Exception("GROUP BY lists a column twice")
    ng = len(g_cols)
    sel, codes, out_names = [], [], []
    aliases_out = {}
    select_pairs = _as_list(js_obj["select"]) if js_obj["select"] != "*" else [{"value": columns[g]} for g in g_cols]
    for dic in select_pairs:
        val = dic["value"]
        if isinstance(val, str):
            idx = col_index(val)
            if idx not in g_cols:
                raise Exception(f"{val} is not an aggregation function or the columns thats grouped on")
            sel.append(idx)
            codes.append(0)
            out_names.append(("key", idx))
        else:
            for agg_func, agg_val in FUNC_TO_FUT_EXT.items():
                if agg_func in val:
                    agg_col = g_cols[0] if val[agg_func] == "*" else col_index(val[agg_func])
                    sel.append(agg_col)
                    codes.append(agg_val)
                    out_names.append((agg_func, agg_col))
        if "name" in dic and out_names:
            aliases_out[dic["name"]] = ng + len(out_names) - 1

    def resolve_output(operand):
        if isinstance(operand, dict):
            (f, c), = operand.items()
            c_idx = g_cols[0] if c == "*" else col_index(c)
            for i, nm in enumerate(out_names):
                if nm == (f, c_idx):
                    return ng + i
            raise Exception(f"{f}({c}) must appear in the select list to be used in HAVING / ORDER BY")
        c = _strip_qualifier(operand, aliases)[1]
        if c in aliases_out:
            return aliases_out[c]
        idx = getIndex(columns, c)
        if idx in g_cols:
            return g_cols.index(idx)
        raise Exception(f"{c} is not an output column of the GROUP BY")

    if "having" in js_obj:
        extras["having"] = _preds(js_obj["having"], resolve_output)
    if orderby:
        extras["orderby"] = [(resolve_output(k["value"]), 1 if k.get("sort") == "desc" else 0) for k in orderby]
    return {"select": sel, "groupbys": codes, "table": table.get_handle(), "g_cols": g_cols, **extras}


def _plan_join(js_obj, name1, table1, pj, aliases):
    """FROM t1 JOIN t2 ON t1.a = t2.b, plain select list (join.fut:52 argument order) or
    GROUP BY over a t2 column with aggregates over t1 columns (hark_entry_join_groupby)."""
    name2, table2 = pj["name2"], pj["table2"]
    if name1 == name2:
        # both aliases would resolve to one schema entry: the ON columns and every qualified column of the second alias
        # would silently be read from the first.  Refuse instead of planning a wrong query.
        raise Exception(f"self-join of {name1} is not supported: register the table under a second name with create_table")
    schema = {name1: table1.get_schema(), name2: table2.get_schema()}

    def locate(name):
        q, c = _strip_qualifier(name, aliases)
        if q is not None:
            if q not in schema:
                raise Exception(f"{q} is not in tables")
            idx = getIndex(schema[q], c)
            if idx < 0:
                raise Exception(f"{c} is not in the schema of table {q}")
            return q, idx
        hits = [(t, getIndex(schema[t], c)) for t in (name1, name2) if getIndex(schema[t], c) >= 0]
        if not hits:
            raise Exception(f"{c} is not in the schema of table {name1}")
        if len(hits) > 1 and name1 != name2:
            raise Exception(f"{c} is ambiguous between {name1} and {name2}")
        return hits[0]

    (ta, ca), (tb, cb) = locate(pj["on"][0]), locate(pj["on"][1])
    if ta == name2 and tb == name1:
        (ta, ca), (tb, cb) = (tb, cb), (ta, ca)
    if not (ta == name1 and tb == name2):
        raise Exception("JOIN condition must compare one column of each table")
    plan = {"table": table1.get_handle(), "table2": table2.get_handle(), "join": (ca, cb)}
    if "limit" in js_obj:
        plan["limit"] = int(js_obj["limit"])
    orderby = _as_list(js_obj["orderby"]) if "orderby" in js_obj else []
    if "where" in js_obj:
        # predicate push-down: every clause of the conjunctive normal form goes to the table its columns belong to
        # and is applied BEFORE the join ("where" -> first table, "where2" -> joined table)
        off = 1 << 20

        def resolve_where(operand):
            q, c = locate(operand)
            return c + (off if q == name2 and name1 != name2 else 0)

        w1, w2, clause = [], [], []
        for pr in _preds(js_obj["where"], resolve_where):
            clause.append(pr)
            if pr[1] & PRED_OR:
                continue
            sides = {x[0] >= off for x in clause}
            if len(sides) > 1:
                raise Exception("a WHERE clause that ORs columns of both joined tables is not supported")
            if sides.pop():
                w2 += [(x[0] - off, x[1], x[2], x[3]) for x in clause]
            else:
                w1 += clause
            clause = []
        if w1:
            plan["where"] = w1
        if w2:
            plan["where2"] = w2
    if "groupby" in js_obj:
        gq, gc = locate(js_obj["groupby"]["value"])
        if gq != name2:
            raise Exception("JOIN ... GROUP BY must group on a column of the joined (second) table")
        s_cols, ops, out_names, out_alias = [], [], [], {}
        for dic in _as_list(js_obj["select"]):
            val = dic["value"]
            if isinstance(val, str):
                if locate(val) != (gq, gc):
                    raise Exception(f"{val} is not an aggregation function or the columns thats grouped on")
                if "name" in dic:
                    out_alias[dic["name"]] = 0
                continue            # the key is output column 0 anyway
            for agg_func, agg_val in FUNC_TO_FUT_EXT.items():
                if agg_func in val:
                    if val[agg_func] == "*":
                        s_cols.append(ca)
                    else:
                        q, c = locate(val[agg_func])
                        if q != name1:
                            raise Exception("aggregates after a JOIN must be over columns of the first table")
                        s_cols.append(c)
                    ops.append(agg_val)
                    out_names.append((agg_func, s_cols[-1]))
                    if "name" in dic:
                        out_alias[dic["name"]] = len(out_names)

        def resolve_output(operand):
            """HAVING / ORDER BY operand -> output column (0 = the group key, i = i-th aggregate)."""
            if isinstance(operand, dict):
                (f, c), = operand.items()
                c_idx = ca if c == "*" else locate(c)[1]
                for i, nm in enumerate(out_names):
                    if nm == (f, c_idx):
                        return i + 1
                raise Exception(f"{f}({c}) must appear in the select list to be used in HAVING / ORDER BY")
            if operand in out_alias:
                return out_alias[operand]
            if locate(operand) == (gq, gc):
                return 0
            raise Exception(f"{operand} is not an output column of the GROUP BY")

        if "having" in js_obj:
            plan["having"] = _preds(js_obj["having"], resolve_output)
        if orderby:
            plan["orderby"] = [(resolve_output(k["value"]), 1 if k.get("sort") == "desc" else 0) for k in orderby]
        plan.update({"g_col": gc, "select": s_cols, "groupbys": ops})
        return plan
    if "having" in js_obj:
        raise Exception("HAVING needs a GROUP BY clause")
    cols1, cols2 = [], []
    sel = js_obj["select"]
    if sel == "*":
        cols1, cols2 = list(range(len(schema[name1]))), list(range(len(schema[name2])))
    else:
        seen_second = False
        for dic in _as_list(sel):
            q, c = locate(dic["value"])
            if q == name1 and not seen_second:
                cols1.append(c)
            elif q == name2:
                seen_second = True
                cols2.append(c)
            else:
                raise Exception("select list after a JOIN must list first-table columns before second-table columns")
    if orderby:        # ORDER BY over the join's output columns (first-table columns, then second-table columns)
        out_cols = [(name1, c) for c in cols1] + [(name2, c) for c in cols2]
        keys = []
        for k in orderby:
            loc = locate(k["value"])
            if loc not in out_cols:
                raise Exception(f"{k['value']} must appear in the select list to be used in ORDER BY")
            keys.append((out_cols.index(loc), 1 if k.get("sort") == "desc" else 0))
        plan["orderby"] = keys
    plan.update({"select": cols1, "select2": cols2})
    return plan
