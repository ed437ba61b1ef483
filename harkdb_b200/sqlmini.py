"""A small SQL SELECT parser emitting ``moz_sql_parser``-shaped dicts.

The reference calls ``moz_sql_parser.parse`` (parse.py:6,27), an unpinned third-party package
(requirements.txt:3) that is not installed in this image.  ``parse(sql)`` below produces the same
JSON shapes for the clause set HarkDB's README lists (README.md:10-16) — the shapes parse.py relies
on are catalogued in SURVEY.md App. C:

    select a, b from t                 {"select": [{"value": "a"}, {"value": "b"}], "from": "t"}
    select a from t                    {"select": {"value": "a"}, "from": "t"}            (dict, not list)
    select * from t                    {"select": "*", "from": "t"}
    select distinct a, b from t        {"select_distinct": [{"value": "a"}, {"value": "b"}], "from": "t"}
    select max(c) from t               {"select": {"value": {"max": "c"}}, ...}
    ... where a > 4 and b <= 2.5       "where": {"and": [{"gt": ["a", 4]}, {"lte": ["b", 2.5]}]}
    ... group by a                     "groupby": {"value": "a"}
    ... having count(c) > 3            "having": {"gt": [{"count": "c"}, 3]}
    ... order by a, b desc             "orderby": [{"value": "a"}, {"value": "b", "sort": "desc"}]
    ... where a > 4 or not (b = 2)     "where": {"or": [{"gt": ["a", 4]}, {"not": {"eq": ["b", 2]}}]}
    ... where a between 1 and 5        "where": {"between": ["a", 1, 5]}          (not between: "not_between")
    ... where a in (1, 2, 3)           "where": {"in": ["a", [1, 2, 3]]}          (not in: "nin")
    from f join d on f.fk = d.pk       "from": ["f", {"join": "d", "on": {"eq": ["f.fk", "d.pk"]}}]
    ... limit 10                       "limit": 10
"""

from __future__ import annotations

import re
from typing import Any, List

KEYWORDS = {"select", "from", "where", "group", "by", "having", "order", "limit", "and", "or", "not", "as",
            "join", "inner", "on", "asc", "desc", "between", "in", "distinct"}
CMP = {"=": "eq", "==": "eq", "!=": "neq", "<>": "neq", ">": "gt", ">=": "gte", "<": "lt", "<=": "lte"}

_TOKEN = re.compile(r"""
    \s*(?:
      (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)
    | (?P<id>[A-Za-z_][A-Za-z_0-9]*(?:\.[A-Za-z_][A-Za-z_0-9]*)?)
    | (?P<op><=|>=|<>|!=|==|=|<|>)
    | (?P<punct>[(),*;+\-])
    )""", re.VERBOSE)


class SqlSyntaxError(Exception):
    pass


def _tokenize(sql: str) -> List[tuple]:
    toks, pos = [], 0
    sql = sql.strip()
    while pos < len(sql):
        m = _TOKEN.match(sql, pos)
        if not m or m.end() == pos:
            raise SqlSyntaxError(f"cannot tokenize at: {sql[pos:pos + 20]!r}")
        pos = m.end()
        if m.group("num") is not None:
            t = m.group("num")
            toks.append(("num", float(t) if any(ch in t for ch in ".eE") else int(t)))
        elif m.group("id") is not None:
            t = m.group("id")
            toks.append(("kw", t.lower()) if t.lower() in KEYWORDS else ("id", t))
        elif m.group("op") is not None:
            toks.append(("op", m.group("op")))
        else:
            toks.append(("punct", m.group("punct")))
    return toks


class _Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", None)

    def eat(self, kind=None, val=None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise SqlSyntaxError(f"expected {val or kind}, got {tok[1]!r}")
        self.i += 1
        return tok

    def at_kw(self, *vals):
        tok = self.peek()
        return tok[0] == "kw" and tok[1] in vals

    # operand := [-] number | func '(' (* | col) ')' | col
    def operand(self) -> Any:
        tok = self.peek()
        if tok == ("punct", "-") or tok == ("punct", "+"):
            self.eat()
            v = self.eat("num")[1]
            return -v if tok[1] == "-" else v
        if tok[0] == "num":
            return self.eat()[1]
        if tok[0] == "id":
            name = self.eat()[1]
            if self.peek() == ("punct", "("):
                self.eat()
                if self.peek() == ("punct", "*"):
                    self.eat()
                    arg = "*"
                else:
                    arg = self.eat("id")[1]
                self.eat("punct", ")")
                return {name.lower(): arg}
            return name
        raise SqlSyntaxError(f"unexpected {tok[1]!r}")

    def comparison(self) -> Any:
        if self.peek() == ("punct", "("):
            self.eat()
            e = self.or_expr()
            self.eat("punct", ")")
            return e
        lhs = self.operand()
        negated = False
        if self.at_kw("not") and self.peek(1)[0] == "kw" and self.peek(1)[1] in ("between", "in"):
            self.eat()
            negated = True
        if self.at_kw("between"):
            self.eat()
            lo = self.operand()
            self.eat("kw", "and")
            hi = self.operand()
            return {"not_between" if negated else "between": [lhs, lo, hi]}
        if self.at_kw("in"):
            self.eat()
            self.eat("punct", "(")
            vals = [self.operand()]
            while self.peek() == ("punct", ","):
                self.eat()
                vals.append(self.operand())
            self.eat("punct", ")")
            return {"nin" if negated else "in": [lhs, vals[0] if len(vals) == 1 else vals]}
        op = self.eat("op")[1]
        rhs = self.operand()
        return {CMP[op]: [lhs, rhs]}

    def not_expr(self) -> Any:
        if self.at_kw("not"):
            self.eat()
            return {"not": self.not_expr()}
        return self.comparison()

    def and_expr(self) -> Any:
        parts = [self.not_expr()]
        while self.at_kw("and"):
            self.eat()
            parts.append(self.not_expr())
        return parts[0] if len(parts) == 1 else {"and": parts}

    def or_expr(self) -> Any:
        parts = [self.and_expr()]
        while self.at_kw("or"):
            self.eat()
            parts.append(self.and_expr())
        return parts[0] if len(parts) == 1 else {"or": parts}

    def select_item(self) -> Any:
        item = {"value": self.operand()}
        if self.at_kw("as"):
            self.eat()
            item["name"] = self.eat("id")[1]
        elif self.peek()[0] == "id":
            item["name"] = self.eat()[1]
        return item

    def table_ref(self) -> Any:
        name = self.eat("id")[1]
        if self.at_kw("as"):
            self.eat()
            return {"value": name, "name": self.eat("id")[1]}
        if self.peek()[0] == "id":
            return {"value": name, "name": self.eat()[1]}
        return name

    def query(self) -> dict:
        out: dict = {}
        self.eat("kw", "select")
        sel_key = "select"
        if self.at_kw("distinct"):
            self.eat()
            sel_key = "select_distinct"
        if self.peek() == ("punct", "*"):
            self.eat()
            out[sel_key] = "*"
        else:
            items = [self.select_item()]
            while self.peek() == ("punct", ","):
                self.eat()
                items.append(self.select_item())
            out[sel_key] = items[0] if len(items) == 1 else items
        self.eat("kw", "from")
        frm: List[Any] = [self.table_ref()]
        while self.at_kw("join", "inner"):
            key = "join"
            if self.at_kw("inner"):
                self.eat()
                key = "inner join"
            self.eat("kw", "join")
            j = {key: self.table_ref()}
            self.eat("kw", "on")
            j["on"] = self.or_expr()
            frm.append(j)
        out["from"] = frm[0] if len(frm) == 1 else frm
        if self.at_kw("where"):
            self.eat()
            out["where"] = self.or_expr()
        if self.at_kw("group"):
            self.eat()
            self.eat("kw", "by")
            cols = [{"value": self.eat("id")[1]}]
            while self.peek() == ("punct", ","):
                self.eat()
                cols.append({"value": self.eat("id")[1]})
            out["groupby"] = cols[0] if len(cols) == 1 else cols
        if self.at_kw("having"):
            self.eat()
            out["having"] = self.or_expr()
        if self.at_kw("order"):
            self.eat()
            self.eat("kw", "by")
            keys = []
            while True:
                k = {"value": self.operand()}
                if self.at_kw("asc"):
                    self.eat()
                elif self.at_kw("desc"):
                    self.eat()
                    k["sort"] = "desc"
                keys.append(k)
                if self.peek() != ("punct", ","):
                    break
                self.eat()
            out["orderby"] = keys[0] if len(keys) == 1 else keys
        if self.at_kw("limit"):
            self.eat()
            out["limit"] = self.eat("num")[1]
        if self.peek() == ("punct", ";"):
            self.eat()
        if self.peek()[0] != "eof":
            raise SqlSyntaxError(f"unexpected trailing input: {self.peek()[1]!r}")
        return out


def parse(sql: str) -> dict:
    """SQL SELECT text -> moz_sql_parser-shaped dict."""
    return _Parser(_tokenize(sql)).query()
