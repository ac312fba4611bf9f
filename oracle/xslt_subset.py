"""A small XSLT 1.0 / XPath 1.0 processor: enough of the language to expand the reference's code templates.

TEST INFRASTRUCTURE.  The reference generates its per-model sources with ``xsltproc`` from
``FLAMEGPU/templates/*.xslt`` (reference Makefile: ``tools/common.mk:349-351``); this image has neither xsltproc nor
lxml, so the reference cannot produce its own ``src/dynamic/*`` here.  This module is the missing tool, written from
the XSLT 1.0 and XPath 1.0 recommendations -- it contains nothing of the reference.  ``oracle/build_ref.py`` drives it to
expand the reference's templates (read where they lie under /root/reference) into ``oracle/_ref/`` (git-ignored).

Supported (SURVEY.md section 8c lists what the eight templates use): ``stylesheet output include template(match|name)
param with-param variable for-each sort if choose/when/otherwise value-of text call-template apply-templates
copy-of message``; the full XPath 1.0 expression grammar over the child, parent, self, attribute, descendant,
descendant-or-self, ancestor, ancestor-or-self, following-sibling and preceding-sibling axes; the core function
library except ``id()`` and ``lang()``.  Output method is text.  Anything else raises ``XsltError`` naming the
construct, so a template that steps outside the subset fails loudly instead of expanding wrongly.
"""
from __future__ import annotations

import math
import os
import re
import xml.dom.minidom as minidom
from typing import Callable, Dict, List, Optional, Tuple

XSL_NS = "http://www.w3.org/1999/XSL/Transform"


class XsltError(Exception):
    pass


# --------------------------------------------------------------------------------------------------------------
# document model helpers (minidom nodes, annotated with a document-order index)
# --------------------------------------------------------------------------------------------------------------

def _index_document(doc) -> None:
    counter = [0]

    def walk(node):
        node._ord = counter[0]
        counter[0] += 1
        if node.nodeType == node.ELEMENT_NODE and node.attributes is not None:
            for i in range(node.attributes.length):
                a = node.attributes.item(i)
                a._ord = counter[0]
                a._owner = node
                counter[0] += 1
        for c in node.childNodes:
            walk(c)

    walk(doc)


def _parent(node):
    if node.nodeType == node.ATTRIBUTE_NODE:
        return getattr(node, "_owner", node.ownerElement)
    return node.parentNode


def string_value(node) -> str:
    t = node.nodeType
    if t in (node.TEXT_NODE, node.CDATA_SECTION_NODE):
        return node.data
    if t == node.ATTRIBUTE_NODE:
        return node.value
    if t in (node.COMMENT_NODE, node.PROCESSING_INSTRUCTION_NODE):
        return node.data
    out: List[str] = []

    def walk(n):
        for c in n.childNodes:
            if c.nodeType in (c.TEXT_NODE, c.CDATA_SECTION_NODE):
                out.append(c.data)
            elif c.nodeType == c.ELEMENT_NODE:
                walk(c)

    walk(node)
    return "".join(out)


class Fragment:
    """Result tree fragment of a text-output stylesheet: only its string value is observable."""

    def __init__(self, text: str):
        self.text = text


# --------------------------------------------------------------------------------------------------------------
# XPath values
# --------------------------------------------------------------------------------------------------------------

def to_string(v) -> str:
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        if math.isnan(v):
            return "NaN"
        if math.isinf(v):
            return "Infinity" if v > 0 else "-Infinity"
        if v == int(v) and abs(v) < 1e18:
            return str(int(v))
        s = repr(v)
        if "e" in s or "E" in s:
            s = ("%.17f" % v).rstrip("0").rstrip(".")
        return s
    if isinstance(v, list):
        return string_value(v[0]) if v else ""
    if isinstance(v, Fragment):
        return v.text
    raise XsltError(f"cannot convert {type(v)} to string")


_NUM_RE = re.compile(r"^\s*-?(\d+(\.\d*)?|\.\d+)\s*$")


def to_number(v) -> float:
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, float):
        return v
    s = to_string(v)
    return float(s) if _NUM_RE.match(s) else float("nan")


def to_boolean(v) -> bool:
    if isinstance(v, bool):
        return v
    if isinstance(v, float):
        return not (v == 0 or math.isnan(v))
    if isinstance(v, str):
        return len(v) > 0
    if isinstance(v, list):
        return len(v) > 0
    if isinstance(v, Fragment):
        return True
    raise XsltError(f"cannot convert {type(v)} to boolean")


def _xpath_round(x: float) -> float:
    if math.isnan(x) or math.isinf(x):
        return x
    return float(math.floor(x + 0.5))


# --------------------------------------------------------------------------------------------------------------
# XPath tokenizer + parser (recursive descent over the XPath 1.0 grammar) -> closures
# --------------------------------------------------------------------------------------------------------------

_TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<num>\d+(\.\d*)?|\.\d+)
  | (?P<str>"[^"]*"|'[^']*')
  | (?P<op>//|::|\.\.|!=|<=|>=|[/()\[\]@,|=<>+\-*.$])
  | (?P<name>[A-Za-z_][\w.\-]*(:[A-Za-z_*][\w.\-]*)?)
""", re.X)

_AXES = {"child", "parent", "self", "attribute", "descendant", "descendant-or-self", "ancestor", "ancestor-or-self",
         "following-sibling", "preceding-sibling"}
_NODE_TYPES = {"node", "text", "comment", "processing-instruction"}


class Context:
    __slots__ = ("node", "pos", "size", "vars", "current")

    def __init__(self, node, pos, size, variables, current=None):
        self.node, self.pos, self.size, self.vars = node, pos, size, variables
        self.current = current if current is not None else node


def _tokenize(expr: str) -> List[Tuple[str, str]]:
    toks: List[Tuple[str, str]] = []
    i = 0
    while i < len(expr):
        m = _TOKEN_RE.match(expr, i)
        if not m:
            raise XsltError(f"XPath: cannot tokenize {expr!r} at {i}")
        i = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        text = m.group(kind)
        # a name token may have swallowed a trailing '-' or '.' sequence that is really an operator: not in the
        # templates' vocabulary (names there never end in '-'), so it is not special-cased.
        toks.append((kind, text))
    # disambiguate '*' and operator names (XPath 1.0 section 3.7)
    out: List[Tuple[str, str]] = []
    for kind, text in toks:
        prev = out[-1] if out else None
        prev_is_operand = prev is not None and not (
            prev[0] == "op" and prev[1] in ("@", "::", "(", "[", ",", "/", "//", "|", "+", "-", "=", "!=", "<", "<=", ">", ">=", "$")
            or prev[0] == "operator")
        if kind == "op" and text == "*" and prev_is_operand:
            out.append(("operator", "*"))
        elif kind == "name" and text in ("and", "or", "mod", "div") and prev_is_operand:
            out.append(("operator", text))
        else:
            out.append((kind, text))
    return out


class _Parser:
    def __init__(self, expr: str, nsmap: Dict[str, str]):
        self.expr, self.ns = expr, nsmap
        self.toks = _tokenize(expr)
        self.i = 0

    # -- token helpers
    def peek(self, k=0):
        return self.toks[self.i + k] if self.i + k < len(self.toks) else (None, None)

    def take(self):
        t = self.peek()
        self.i += 1
        return t

    def accept(self, kind, text=None):
        k, t = self.peek()
        if k == kind and (text is None or t == text):
            self.i += 1
            return True
        return False

    def expect(self, kind, text):
        if not self.accept(kind, text):
            raise XsltError(f"XPath: expected {text!r} in {self.expr!r} near token {self.i}")

    def parse(self):
        f = self.or_expr()
        if self.i != len(self.toks):
            raise XsltError(f"XPath: trailing tokens in {self.expr!r}: {self.toks[self.i:]}")
        return f

    # -- expressions
    def or_expr(self):
        left = self.and_expr()
        while self.accept("operator", "or"):
            right = self.and_expr()
            left = (lambda a, b: lambda c: to_boolean(a(c)) or to_boolean(b(c)))(left, right)
        return left

    def and_expr(self):
        left = self.equality()
        while self.accept("operator", "and"):
            right = self.equality()
            left = (lambda a, b: lambda c: to_boolean(a(c)) and to_boolean(b(c)))(left, right)
        return left

    def equality(self):
        left = self.relational()
        while True:
            k, t = self.peek()
            if k == "op" and t in ("=", "!="):
                self.take()
                right = self.relational()
                left = (lambda a, b, op: lambda c: _compare(a(c), b(c), op))(left, right, t)
            else:
                return left

    def relational(self):
        left = self.additive()
        while True:
            k, t = self.peek()
            if k == "op" and t in ("<", "<=", ">", ">="):
                self.take()
                right = self.additive()
                left = (lambda a, b, op: lambda c: _compare(a(c), b(c), op))(left, right, t)
            else:
                return left

    def additive(self):
        left = self.multiplicative()
        while True:
            k, t = self.peek()
            if k == "op" and t in ("+", "-"):
                self.take()
                right = self.multiplicative()
                if t == "+":
                    left = (lambda a, b: lambda c: to_number(a(c)) + to_number(b(c)))(left, right)
                else:
                    left = (lambda a, b: lambda c: to_number(a(c)) - to_number(b(c)))(left, right)
            else:
                return left

    def multiplicative(self):
        left = self.unary()
        while True:
            k, t = self.peek()
            if k == "operator" and t in ("*", "div", "mod"):
                self.take()
                right = self.unary()
                left = (lambda a, b, op: lambda c: _arith(to_number(a(c)), to_number(b(c)), op))(left, right, t)
            else:
                return left

    def unary(self):
        if self.accept("op", "-"):
            inner = self.unary()
            return lambda c: -to_number(inner(c))
        return self.union()

    def union(self):
        left = self.path_expr()
        while self.accept("op", "|"):
            right = self.path_expr()
            left = (lambda a, b: lambda c: _doc_order(_nodeset(a(c)) + _nodeset(b(c))))(left, right)
        return left

    def path_expr(self):
        k, t = self.peek()
        is_filter = (k in ("num", "str") or (k == "op" and t in ("(", "$"))
                     or (k == "name" and self.peek(1) == ("op", "(") and t not in _NODE_TYPES))
        if not is_filter:
            return self.location_path()
        prim = self.primary()
        preds = []
        while self.peek() == ("op", "["):
            preds.append(self.predicate())
        if preds:
            prim = (lambda p, ps: lambda c: _apply_predicates(_nodeset(p(c)), ps, c))(prim, preds)
        k, t = self.peek()
        if k == "op" and t in ("/", "//"):
            self.take()
            steps = []
            if t == "//":
                steps.append(("descendant-or-self", ("type", "node"), []))
            steps += self.relative_steps()
            return (lambda p, st: lambda c: _walk(_nodeset(p(c)), st, c))(prim, steps)
        return prim

    def primary(self):
        k, t = self.take()
        if k == "num":
            v = float(t)
            return lambda c: v
        if k == "str":
            s = t[1:-1]
            return lambda c: s
        if k == "op" and t == "(":
            inner = self.or_expr()
            self.expect("op", ")")
            return inner
        if k == "op" and t == "$":
            _, name = self.take()

            def var(c, name=name):
                if name not in c.vars:
                    raise XsltError(f"XPath: unbound variable ${name} in {self.expr!r}")
                return c.vars[name]
            return var
        if k == "name":
            self.expect("op", "(")
            args = []
            if not self.accept("op", ")"):
                args.append(self.or_expr())
                while self.accept("op", ","):
                    args.append(self.or_expr())
                self.expect("op", ")")
            if t not in _FUNCTIONS:
                raise XsltError(f"XPath: function {t}() is outside the supported subset ({self.expr!r})")
            fn = _FUNCTIONS[t]
            return lambda c: fn(c, *[a for a in args])
        raise XsltError(f"XPath: unexpected token {t!r} in {self.expr!r}")

    def predicate(self):
        self.expect("op", "[")
        f = self.or_expr()
        self.expect("op", "]")
        return f

    def location_path(self):
        k, t = self.peek()
        if k == "op" and t == "/":
            self.take()
            k2, t2 = self.peek()
            starts_step = k2 == "name" or (k2 == "op" and t2 in ("*", "@", ".", ".."))
            steps = self.relative_steps() if starts_step else []
            return lambda c: _walk([_root(c.node)], steps, c)
        if k == "op" and t == "//":
            self.take()
            steps = [("descendant-or-self", ("type", "node"), [])] + self.relative_steps()
            return lambda c: _walk([_root(c.node)], steps, c)
        steps = self.relative_steps()
        return lambda c: _walk([c.node], steps, c)

    def relative_steps(self):
        steps = [self.step()]
        while True:
            k, t = self.peek()
            if k == "op" and t == "/":
                self.take()
                steps.append(self.step())
            elif k == "op" and t == "//":
                self.take()
                steps.append(("descendant-or-self", ("type", "node"), []))
                steps.append(self.step())
            else:
                return steps

    def step(self):
        if self.accept("op", "."):
            return ("self", ("type", "node"), [])
        if self.accept("op", ".."):
            return ("parent", ("type", "node"), [])
        axis = "child"
        if self.accept("op", "@"):
            axis = "attribute"
        else:
            k, t = self.peek()
            if k == "name" and self.peek(1) == ("op", "::"):
                if t not in _AXES:
                    raise XsltError(f"XPath: axis {t}:: is outside the supported subset ({self.expr!r})")
                axis = t
                self.take()
                self.take()
        k, t = self.take()
        if k == "op" and t == "*":
            test = ("name", None, None)
        elif k == "name":
            if t in _NODE_TYPES and self.peek() == ("op", "("):
                self.take()
                if self.peek()[0] == "str":
                    self.take()
                self.expect("op", ")")
                test = ("type", t)
            elif ":" in t:
                prefix, local = t.split(":", 1)
                if prefix not in self.ns:
                    raise XsltError(f"XPath: undeclared namespace prefix {prefix!r} in {self.expr!r}")
                test = ("name", self.ns[prefix], None if local == "*" else local)
            else:
                test = ("name", "", t)
        else:
            raise XsltError(f"XPath: bad node test {t!r} in {self.expr!r}")
        preds = []
        while self.peek() == ("op", "["):
            preds.append(self.predicate())
        return (axis, test, preds)


def _root(node):
    while _parent(node) is not None:
        node = _parent(node)
    return node


def _nodeset(v) -> list:
    if isinstance(v, list):
        return v
    raise XsltError("XPath: a node-set was required")


def _doc_order(nodes: list) -> list:
    seen, out = set(), []
    for n in sorted(nodes, key=lambda n: n._ord):
        if id(n) not in seen:
            seen.add(id(n))
            out.append(n)
    return out


def _axis_nodes(node, axis: str) -> list:
    """Nodes on `axis` from `node`, in axis order (reverse document order for the reverse axes)."""
    if axis == "child":
        return list(node.childNodes) if node.nodeType != node.ATTRIBUTE_NODE else []
    if axis == "self":
        return [node]
    if axis == "parent":
        p = _parent(node)
        return [p] if p is not None else []
    if axis == "attribute":
        if node.nodeType != node.ELEMENT_NODE or node.attributes is None:
            return []
        return [a for a in (node.attributes.item(i) for i in range(node.attributes.length))
                if a.prefix != "xmlns" and a.name != "xmlns"]
    if axis in ("descendant", "descendant-or-self"):
        out = [node] if axis == "descendant-or-self" else []

        def walk(n):
            for c in n.childNodes:
                out.append(c)
                walk(c)
        if node.nodeType != node.ATTRIBUTE_NODE:
            walk(node)
        return out
    if axis in ("ancestor", "ancestor-or-self"):
        out = [node] if axis == "ancestor-or-self" else []
        p = _parent(node)
        while p is not None:
            out.append(p)
            p = _parent(p)
        return out
    if axis == "following-sibling":
        out, s = [], node.nextSibling if node.nodeType != node.ATTRIBUTE_NODE else None
        while s is not None:
            out.append(s)
            s = s.nextSibling
        return out
    if axis == "preceding-sibling":
        out, s = [], node.previousSibling if node.nodeType != node.ATTRIBUTE_NODE else None
        while s is not None:
            out.append(s)
            s = s.previousSibling
        return out
    raise XsltError(f"XPath: axis {axis}")


def _test_node(node, axis: str, test) -> bool:
    if test[0] == "type":
        kind = test[1]
        if kind == "node":
            return True
        if kind == "text":
            return node.nodeType in (node.TEXT_NODE, node.CDATA_SECTION_NODE)
        if kind == "comment":
            return node.nodeType == node.COMMENT_NODE
        return node.nodeType == node.PROCESSING_INSTRUCTION_NODE
    principal = node.ATTRIBUTE_NODE if axis == "attribute" else node.ELEMENT_NODE
    if node.nodeType != principal:
        return False
    _, uri, local = test
    if uri is None:
        return True
    if (node.namespaceURI or "") != uri:
        return False
    return local is None or (node.localName or node.nodeName) == local


def _apply_predicates(nodes: list, preds: list, c: Context) -> list:
    for p in preds:
        size, kept = len(nodes), []
        for i, n in enumerate(nodes):
            v = p(Context(n, i + 1, size, c.vars, c.current))
            if (v == float(i + 1)) if isinstance(v, float) and not isinstance(v, bool) else to_boolean(v):
                kept.append(n)
        nodes = kept
    return nodes


def _walk(start: list, steps: list, c: Context) -> list:
    current = start
    for axis, test, preds in steps:
        found = []
        for n in current:
            cand = [m for m in _axis_nodes(n, axis) if _test_node(m, axis, test)]
            found.extend(_apply_predicates(cand, preds, c))
        current = _doc_order(found)
    return current


def _arith(a: float, b: float, op: str) -> float:
    if op == "*":
        return a * b
    if op == "div":
        if b == 0:
            if a == 0 or math.isnan(a):
                return float("nan")
            return math.copysign(float("inf"), a) * (math.copysign(1.0, b))
        return a / b
    if b == 0 or math.isnan(a) or math.isnan(b) or math.isinf(a):
        return float("nan")
    return math.fmod(a, b)


def _cmp_atoms(a, b, op: str) -> bool:
    if op in ("=", "!="):
        if isinstance(a, bool) or isinstance(b, bool):
            x, y = to_boolean(a), to_boolean(b)
        elif isinstance(a, float) or isinstance(b, float):
            x, y = to_number(a), to_number(b)
        else:
            x, y = to_string(a), to_string(b)
        return (x == y) if op == "=" else (x != y)
    x, y = to_number(a), to_number(b)
    return {"<": x < y, "<=": x <= y, ">": x > y, ">=": x >= y}[op]


def _compare(a, b, op: str) -> bool:
    """XPath 1.0 section 3.4: comparisons involving node-sets are existential."""
    a = a.text if isinstance(a, Fragment) else a
    b = b.text if isinstance(b, Fragment) else b
    la, lb = isinstance(a, list), isinstance(b, list)
    if la and lb:
        sb = [string_value(n) for n in b]
        return any(_cmp_atoms(string_value(n), s, op) for n in a for s in sb)
    if la or lb:
        nodes, other, flipped = (a, b, False) if la else (b, a, True)
        if isinstance(other, bool):
            return _cmp_atoms(to_boolean(nodes), other, op) if not flipped else _cmp_atoms(other, to_boolean(nodes), op)
        for n in nodes:
            s = string_value(n)
            if (_cmp_atoms(other, s, op) if flipped else _cmp_atoms(s, other, op)):
                return True
        return False
    return _cmp_atoms(a, b, op)


# -- core function library (arguments arrive as unevaluated closures) ------------------------------------------

def _f_string(c, *a):
    return to_string(a[0](c)) if a else string_value(c.node)


def _f_substring(c, s, start, length=None):
    s = to_string(s(c))
    st = _xpath_round(to_number(start(c)))
    if math.isnan(st):
        return ""
    if length is None:
        return "".join(ch for i, ch in enumerate(s, 1) if i >= st)
    ln = _xpath_round(to_number(length(c)))
    if math.isnan(ln):
        return ""
    return "".join(ch for i, ch in enumerate(s, 1) if st <= i < st + ln)


def _f_translate(c, s, frm, to):
    s, frm, to = to_string(s(c)), to_string(frm(c)), to_string(to(c))
    table = {}
    for i, ch in enumerate(frm):
        if ch not in table:
            table[ch] = to[i] if i < len(to) else None
    return "".join(table.get(ch, ch) or "" if ch in table else ch for ch in s)


def _f_name(c, *a, local=False):
    nodes = _nodeset(a[0](c)) if a else [c.node]
    if not nodes:
        return ""
    n = nodes[0]
    if n.nodeType in (n.ELEMENT_NODE, n.ATTRIBUTE_NODE):
        return (n.localName or n.nodeName) if local else n.nodeName
    return ""


def _f_namespace_uri(c, *a):
    nodes = _nodeset(a[0](c)) if a else [c.node]
    if not nodes or nodes[0].nodeType not in (nodes[0].ELEMENT_NODE, nodes[0].ATTRIBUTE_NODE):
        return ""
    return nodes[0].namespaceURI or ""


def _f_sum(c, ns):
    return float(sum(to_number(string_value(n)) for n in _nodeset(ns(c))))


_FUNCTIONS: Dict[str, Callable] = {
    "last": lambda c: float(c.size),
    "position": lambda c: float(c.pos),
    "count": lambda c, ns: float(len(_nodeset(ns(c)))),
    "local-name": lambda c, *a: _f_name(c, *a, local=True),
    "name": lambda c, *a: _f_name(c, *a),
    "namespace-uri": lambda c, *a: _f_namespace_uri(c, *a),
    "string": _f_string,
    "concat": lambda c, *a: "".join(to_string(x(c)) for x in a),
    "starts-with": lambda c, a, b: to_string(a(c)).startswith(to_string(b(c))),
    "contains": lambda c, a, b: to_string(b(c)) in to_string(a(c)),
    "substring-before": lambda c, a, b: (lambda s, t: s[:s.index(t)] if t in s else "")(to_string(a(c)), to_string(b(c))),
    "substring-after": lambda c, a, b: (lambda s, t: s[s.index(t) + len(t):] if t in s else "")(to_string(a(c)), to_string(b(c))),
    "substring": _f_substring,
    "string-length": lambda c, *a: float(len(to_string(a[0](c)) if a else string_value(c.node))),
    "normalize-space": lambda c, *a: " ".join((to_string(a[0](c)) if a else string_value(c.node)).split()),
    "translate": _f_translate,
    "boolean": lambda c, a: to_boolean(a(c)),
    "not": lambda c, a: not to_boolean(a(c)),
    "true": lambda c: True,
    "false": lambda c: False,
    "number": lambda c, *a: to_number(a[0](c)) if a else to_number(string_value(c.node)),
    "sum": _f_sum,
    "floor": lambda c, a: float(math.floor(x)) if not (math.isnan(x := to_number(a(c))) or math.isinf(x)) else x,
    "ceiling": lambda c, a: float(math.ceil(x)) if not (math.isnan(x := to_number(a(c))) or math.isinf(x)) else x,
    "round": lambda c, a: _xpath_round(to_number(a(c))),
    "current": lambda c: [c.current],
}


# --------------------------------------------------------------------------------------------------------------
# match patterns (template match="...")
# --------------------------------------------------------------------------------------------------------------

class _Pattern:
    """Union of location-path patterns restricted to child/attribute steps with '/' and '//' separators."""

    def __init__(self, text: str, nsmap: Dict[str, str]):
        self.text = text
        self.alternatives = []
        for alt in _split_top_level(text, "|"):
            alt = alt.strip()
            p = _Parser(alt, nsmap)
            absolute = False
            steps: List[Tuple[str, tuple]] = []       # (separator before the step, step)
            if p.peek() == ("op", "/") and len(p.toks) == 1:
                self.alternatives.append(("root", []))
                continue
            sep = None
            if p.accept("op", "//"):
                sep = "//"
            elif p.accept("op", "/"):
                absolute, sep = True, "/"
            while True:
                steps.append((sep, p.step()))
                if p.accept("op", "//"):
                    sep = "//"
                elif p.accept("op", "/"):
                    sep = "/"
                else:
                    break
            if p.i != len(p.toks):
                raise XsltError(f"pattern {text!r} is outside the supported subset")
            self.alternatives.append(("abs" if absolute else "rel", steps))
        self.priority = self._default_priority()

    def _default_priority(self) -> float:
        if len(self.alternatives) == 1 and self.alternatives[0][0] != "root":
            steps = self.alternatives[0][1]
            if len(steps) == 1 and steps[0][0] is None and not steps[0][1][2]:
                test = steps[0][1][1]
                if test[0] == "name":
                    if test[1] is None:
                        return -0.5
                    return 0.0 if test[2] is not None else -0.25
                return -0.5
        return 0.5

    def matches(self, node, c: Context) -> bool:
        for kind, steps in self.alternatives:
            if kind == "root":
                if node.nodeType == node.DOCUMENT_NODE:
                    return True
                continue
            if self._match_steps(node, steps, len(steps) - 1, kind == "abs", c):
                return True
        return False

    def _match_steps(self, node, steps, k, absolute, c) -> bool:
        sep, (axis, test, preds) = steps[k]
        if axis not in ("child", "attribute") or not _test_node(node, axis, test):
            return False
        parent = _parent(node)
        if preds:
            if parent is None:
                return False
            sibs = [m for m in _axis_nodes(parent, axis) if _test_node(m, axis, test)]
            if not any(m is node for m in _apply_predicates(sibs, preds, c)):
                return False
        if k == 0:
            if sep is None or sep == "//":
                return True
            return parent is not None and parent.nodeType == parent.DOCUMENT_NODE     # absolute '/'
        if parent is None:
            return False
        if sep == "/":
            return self._match_steps(parent, steps, k - 1, absolute, c)
        while parent is not None:
            if self._match_steps(parent, steps, k - 1, absolute, c):
                return True
            parent = _parent(parent)
        return False


def _split_top_level(text: str, sep: str) -> List[str]:
    out, depth, cur, quote = [], 0, [], None
    for ch in text:
        if quote:
            cur.append(ch)
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur))
    return out


# --------------------------------------------------------------------------------------------------------------
# the stylesheet
# --------------------------------------------------------------------------------------------------------------

class Stylesheet:
    def __init__(self, path: str):
        self.path = os.path.abspath(path)
        self.named: Dict[str, object] = {}
        self.matchers: List[Tuple[float, int, _Pattern, object]] = []
        self.globals: List[object] = []
        self._xpath_cache: Dict[Tuple[int, str], Callable] = {}
        self._nsmaps: Dict[int, Dict[str, str]] = {}
        self._serial = 0
        self.messages: List[str] = []
        self._load(self.path, set())

    # -- loading
    def _load(self, path: str, seen: set) -> None:
        if path in seen:
            raise XsltError(f"circular xsl:include of {path}")
        seen.add(path)
        doc = minidom.parse(path)
        root = doc.documentElement
        if root.namespaceURI != XSL_NS or root.localName not in ("stylesheet", "transform"):
            raise XsltError(f"{path}: not an XSLT stylesheet")
        for el in [n for n in root.childNodes if n.nodeType == n.ELEMENT_NODE]:
            if el.namespaceURI != XSL_NS:
                continue
            name = el.localName
            if name == "include" or name == "import":
                self._load(os.path.normpath(os.path.join(os.path.dirname(path), el.getAttribute("href"))), seen)
            elif name == "output":
                if el.getAttribute("method") not in ("text", ""):
                    raise XsltError(f"{path}: only the text output method is supported")
            elif name == "template":
                self._serial += 1
                if el.hasAttribute("name"):
                    self.named[el.getAttribute("name")] = el
                if el.hasAttribute("match"):
                    if el.hasAttribute("mode"):
                        raise XsltError("template modes are outside the supported subset")
                    pat = _Pattern(el.getAttribute("match"), self._nsmap(el))
                    prio = float(el.getAttribute("priority")) if el.hasAttribute("priority") else pat.priority
                    self.matchers.append((prio, self._serial, pat, el))
            elif name in ("variable", "param"):
                self.globals.append(el)
            elif name in ("strip-space", "preserve-space", "decimal-format", "namespace-alias", "attribute-set", "key"):
                raise XsltError(f"xsl:{name} is outside the supported subset")
        self.matchers.sort(key=lambda t: (t[0], t[1]), reverse=True)

    def _nsmap(self, el) -> Dict[str, str]:
        key = id(el)
        m = self._nsmaps.get(key)
        if m is None:
            m = {}
            chain = []
            n = el
            while n is not None and n.nodeType == n.ELEMENT_NODE:
                chain.append(n)
                n = n.parentNode
            for n in reversed(chain):
                for i in range(n.attributes.length):
                    a = n.attributes.item(i)
                    if a.prefix == "xmlns":
                        m[a.localName] = a.value
            self._nsmaps[key] = m
        return m

    def _xpath(self, el, attr: str) -> Callable:
        key = (id(el), attr)
        f = self._xpath_cache.get(key)
        if f is None:
            if not el.hasAttribute(attr):
                raise XsltError(f"xsl:{el.localName} without @{attr}")
            f = _Parser(el.getAttribute(attr), self._nsmap(el)).parse()
            self._xpath_cache[key] = f
        return f

    # -- running
    def transform(self, source_path: str, params: Optional[Dict[str, object]] = None) -> str:
        doc = minidom.parse(source_path)
        _index_document(doc)
        out: List[str] = []
        variables: Dict[str, object] = {}
        root_ctx = Context(doc, 1, 1, variables)
        for g in self.globals:
            name = g.getAttribute("name")
            if g.localName == "param" and params and name in params:
                variables[name] = params[name]
            else:
                variables[name] = self._binding_value(g, root_ctx)
        self._globals = variables
        self._apply([doc], root_ctx, out, {})
        return "".join(out)

    def _binding_value(self, el, c: Context):
        if el.hasAttribute("select"):
            return self._xpath(el, "select")(c)
        if not any(n.nodeType in (n.ELEMENT_NODE, n.TEXT_NODE, n.CDATA_SECTION_NODE) for n in el.childNodes):
            return ""                                       # XSLT 1.0 section 11.2: empty content, no select
        buf: List[str] = []
        self._run_children(el, c, buf)
        return Fragment("".join(buf))

    def _find_template(self, node, c: Context):
        for _, _, pat, el in self.matchers:
            if pat.matches(node, c):
                return el
        return None

    def _apply(self, nodes: list, c: Context, out: List[str], params: Dict[str, object]) -> None:
        size = len(nodes)
        for i, n in enumerate(nodes):
            ctx = Context(n, i + 1, size, c.vars)
            tmpl = self._find_template(n, ctx)
            if tmpl is not None:
                self._instantiate(tmpl, ctx, out, params)
            elif n.nodeType in (n.DOCUMENT_NODE, n.ELEMENT_NODE):       # built-in rules
                self._apply(list(n.childNodes), ctx, out, {})
            elif n.nodeType in (n.TEXT_NODE, n.CDATA_SECTION_NODE, n.ATTRIBUTE_NODE):
                out.append(string_value(n))

    def _instantiate(self, tmpl, c: Context, out: List[str], params: Dict[str, object]) -> None:
        scope = dict(self._globals)
        ctx = Context(c.node, c.pos, c.size, scope)
        self._run_children(tmpl, ctx, out, params)

    def _run_children(self, parent, c: Context, out: List[str], params: Optional[Dict[str, object]] = None) -> None:
        """Instantiate the template body under `parent`.  Variables bound here are visible to following siblings and
        their descendants only, so the body runs on a copy-on-bind scope."""
        scope = c.vars
        own_scope = False
        for node in parent.childNodes:
            t = node.nodeType
            if t in (node.TEXT_NODE, node.CDATA_SECTION_NODE):
                if node.data.strip(" \t\r\n") != "" :
                    out.append(node.data)
                continue
            if t != node.ELEMENT_NODE:
                continue
            if node.namespaceURI != XSL_NS:
                raise XsltError(f"literal result element <{node.nodeName}> in a text-output stylesheet")
            name = node.localName
            ctx = c if scope is c.vars else Context(c.node, c.pos, c.size, scope, c.current)
            if name in ("variable", "param"):
                if not own_scope:
                    scope = dict(scope)
                    own_scope = True
                    ctx = Context(c.node, c.pos, c.size, scope, c.current)
                vname = node.getAttribute("name")
                if name == "param" and params is not None and vname in params:
                    scope[vname] = params[vname]
                else:
                    scope[vname] = self._binding_value(node, ctx)
            else:
                self._instruction(name, node, ctx, out)

    def _instruction(self, name: str, el, c: Context, out: List[str]) -> None:
        if name == "value-of":
            out.append(to_string(self._xpath(el, "select")(c)))
        elif name == "text":
            out.append("".join(n.data for n in el.childNodes if n.nodeType in (n.TEXT_NODE, n.CDATA_SECTION_NODE)))
        elif name == "if":
            if to_boolean(self._xpath(el, "test")(c)):
                self._run_children(el, c, out)
        elif name == "choose":
            for branch in el.childNodes:
                if branch.nodeType != branch.ELEMENT_NODE:
                    continue
                if branch.localName == "when":
                    if to_boolean(self._xpath(branch, "test")(c)):
                        self._run_children(branch, c, out)
                        return
                elif branch.localName == "otherwise":
                    self._run_children(branch, c, out)
                    return
        elif name == "for-each":
            nodes = self._sorted(el, _nodeset(self._xpath(el, "select")(c)), c)
            size = len(nodes)
            for i, n in enumerate(nodes):
                self._run_children(el, Context(n, i + 1, size, c.vars), out)
        elif name == "call-template":
            tname = el.getAttribute("name")
            if tname not in self.named:
                raise XsltError(f"call-template: no template named {tname!r}")
            self._instantiate(self.named[tname], c, out, self._with_params(el, c))
        elif name == "apply-templates":
            if el.hasAttribute("mode"):
                raise XsltError("template modes are outside the supported subset")
            nodes = _nodeset(self._xpath(el, "select")(c)) if el.hasAttribute("select") else list(c.node.childNodes)
            self._apply(self._sorted(el, nodes, c), c, out, self._with_params(el, c))
        elif name == "copy-of":
            out.append(to_string(self._xpath(el, "select")(c)))
        elif name == "message":
            buf: List[str] = []
            self._run_children(el, c, buf)
            self.messages.append("".join(buf))
            if el.getAttribute("terminate") == "yes":
                raise XsltError("xsl:message terminate=yes: " + "".join(buf))
        elif name in ("sort", "with-param"):
            pass                                            # consumed by the parent instruction
        elif name == "comment" or name == "processing-instruction":
            pass                                            # not part of text output
        else:
            raise XsltError(f"xsl:{name} is outside the supported subset")

    def _with_params(self, el, c: Context) -> Dict[str, object]:
        return {w.getAttribute("name"): self._binding_value(w, c) for w in el.childNodes
                if w.nodeType == w.ELEMENT_NODE and w.namespaceURI == XSL_NS and w.localName == "with-param"}

    def _sorted(self, el, nodes: list, c: Context) -> list:
        keys = [s for s in el.childNodes if s.nodeType == s.ELEMENT_NODE and s.namespaceURI == XSL_NS and s.localName == "sort"]
        if not keys:
            return nodes
        size = len(nodes)
        decorated = list(enumerate(nodes))
        for s in reversed(keys):                            # stable sorts, least significant key first
            sel = self._xpath(s, "select") if s.hasAttribute("select") else (lambda cc: string_value(cc.node))
            numeric = s.getAttribute("data-type") == "number"
            descending = s.getAttribute("order") == "descending"

            def key(item, sel=sel, numeric=numeric):
                i, n = item
                v = sel(Context(n, i + 1, size, c.vars))
                if numeric:
                    x = to_number(v)
                    return (0, 0.0) if math.isnan(x) else (1, x)     # NaN sorts first in ascending order
                return to_string(v)
            decorated.sort(key=key, reverse=descending)     # reverse=True keeps equal items in document order too
        return [n for _, n in decorated]


def transform_file(stylesheet_path: str, source_path: str, output_path: Optional[str] = None) -> str:
    text = Stylesheet(stylesheet_path).transform(source_path)
    if output_path:
        os.makedirs(os.path.dirname(os.path.abspath(output_path)), exist_ok=True)
        with open(output_path, "w") as f:
            f.write(text)
    return text


if __name__ == "__main__":
    import sys
    if len(sys.argv) not in (3, 4):
        sys.exit("usage: xslt_subset.py stylesheet.xslt source.xml [output]")
    result = transform_file(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) == 4 else None)
    if len(sys.argv) == 3:
        sys.stdout.write(result)
