"""ORACLE: Python restatement of the reference's 0.xml reader.  TEST INFRASTRUCTURE ONLY.

Follows ``readInitialStates`` (FLAMEGPU/templates/io.xslt:211-621) character by
character: tags are matched by exact name, variable tags by *bare name across
all agent types* (io:458-459), values are parsed when the next ``<`` arrives
with C semantics (``(float)atof``, ``strtol(..., 0)``; cmn:168-221), an agent
is committed on ``</xagent>`` under the last ``<name>`` seen, unspecified
variables take ``defaultValue`` or 0, unknown tags are ignored, ``<!-- -->``
comments are skipped, reading stops at ``</states>``.
"""
from __future__ import annotations

import ctypes
import re
from typing import Dict, List, Tuple

import numpy as np

_libc = ctypes.CDLL("libc.so.6")
_libc.atof.restype = ctypes.c_double
_libc.atof.argtypes = [ctypes.c_char_p]
_libc.strtol.restype = ctypes.c_long
_libc.strtol.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
_libc.strtoul.restype = ctypes.c_ulong
_libc.strtoul.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]


def _parse(ctype: str, text: str):
    b = text.encode("latin-1", "replace")
    if ctype == "float":
        with np.errstate(over="ignore"):          # (float)atof("1e39") is +inf in C, silently
            return np.float32(_libc.atof(b))
    if ctype == "double":
        return np.float64(_libc.atof(b))
    if ctype == "int":
        return np.int32(np.int64(_libc.strtol(b, None, 0)).astype(np.int32))
    if ctype == "unsigned int":
        return np.uint32(np.uint64(_libc.strtoul(b, None, 0)).astype(np.uint32))
    raise ValueError(ctype)


def read_states(path: str, agents: List[Tuple[str, List[Tuple[str, str, float]]]], constants: List[Tuple[str, str]] = ()):
    """agents: [(agent name, [(var name, ctype, default[, arrayLength])])].  Returns
    ({agent: {var: array}}, {constant: value}); an array variable comes back as a (count, arrayLength) array."""
    data = open(path, "rb").read().decode("latin-1")
    agents = [(a, [(tuple(v) + (0,))[:4] for v in vs]) for a, vs in agents]
    lengths = {a: {v: int(l or 0) for v, _, _, l in vs} for a, vs in agents}
    agents = [(a, [v[:3] for v in vs]) for a, vs in agents]
    out = {a: {v: [] for v, _, _ in vs} for a, vs in agents}
    scalar_defaults = {a: {v: _parse(t, repr(d) if t in ("float", "double") else str(int(d))) for v, t, d in vs} for a, vs in agents}

    def _fresh():
        return {a: {v: (np.full(lengths[a][v], scalar_defaults[a][v]) if lengths[a][v] > 1 else scalar_defaults[a][v])
                    for v in scalar_defaults[a]} for a in scalar_defaults}
    defaults = _fresh()
    types = {a: {v: t for v, t, _ in vs} for a, vs in agents}
    cur = _fresh()
    env: Dict[str, object] = {}
    ctypes_env = dict(constants)
    buf: List[str] = []
    in_tag = in_comment = in_env = in_xagent = in_name = False
    open_var = ""
    agentname = ""
    for c in data:
        if in_comment:
            if c == "-":
                buf.append(c)
            elif c == ">" and len(buf) >= 2:
                in_comment = False
                buf = []
            else:
                buf = []
        elif c == ">":
            tag = "".join(buf)
            if tag == "/states":
                break
            if tag == "environment":
                in_env = True
            if tag == "/environment":
                in_env = False
            if tag == "name":
                in_name = True
            if tag == "/name":
                in_name = False
            if tag == "xagent":
                in_xagent = True
            if tag == "/xagent":
                if agentname in out:
                    for v in out[agentname]:
                        out[agentname][v].append(cur[agentname][v])
                cur = _fresh()
                in_xagent = False
            if tag.startswith("/"):
                if open_var == tag[1:]:
                    open_var = ""
            else:
                open_var = tag
            in_tag = False
            buf = []
        elif c == "<":
            text = "".join(buf)
            in_tag = True
            if in_name:
                agentname = text
            elif in_xagent:
                for a, vs in agents:
                    if open_var in types[a]:
                        if lengths[a][open_var] > 1:      # readArrayInput (io.xslt:69-93): strtok on ","
                            items = [t for t in text.split(",") if t != ""]
                            if len(items) > lengths[a][open_var]:
                                raise ValueError("Error: variable array %s->%s has too many items (%d), expected %d!"
                                                 % (a, open_var, lengths[a][open_var], lengths[a][open_var]))
                            for i, t in enumerate(items):
                                cur[a][open_var][i] = _parse(types[a][open_var], t)
                        else:
                            cur[a][open_var] = _parse(types[a][open_var], text)
            elif in_env:
                if open_var in ctypes_env:
                    env[open_var] = _parse(ctypes_env[open_var], text)
            buf = []
        elif in_tag:
            if len(buf) == 2 and c == "-" and buf[1] == "-" and buf[0] == "!":
                in_comment = True
                buf = []
            buf.append(c)
        else:
            buf.append(c)
    np_types = {"float": np.float32, "double": np.float64, "int": np.int32, "unsigned int": np.uint32}
    res = {a: {v: np.array(out[a][v], dtype=np_types[types[a][v]]) for v in out[a]} for a in out}
    return res, env


_INT_FORMATS = {"bool": "%d", "char": "%d", "unsigned char": "%u", "short": "%d", "unsigned short": "%u", "int": "%d",
                "unsigned int": "%u", "long long int": "%d", "unsigned long long int": "%u"}


def _format(ctype: str, value) -> str:
    """formatSpecifier (_common_templates.xslt:25-60): integers by type, anything else ``%f`` (C promotes a
    float argument to double, so ``%f`` of the float32 value is what ``sprintf`` prints)."""
    if ctype in _INT_FORMATS:
        return "%d" % int(value)
    return "%f" % float(value)


def write_states(path: str, iteration: int, agents, states, constants=()) -> None:
    """Python restatement of ``saveIterationData`` (io.xslt:124-200).

    agents:    [(agent name, [(state name, [(var name, ctype)])...])] in model order; ``states`` maps
               (agent, state) -> {var: array}.  constants: [(name, ctype, value-or-array)].
    """
    out = ["<states>\n<itno>%i</itno>\n<environment>\n" % iteration]
    for name, ctype, value in constants:
        vals = np.atleast_1d(value)
        out.append("\t<%s>%s</%s>\n" % (name, ",".join(_format(ctype, v) for v in vals), name))
    out.append("</environment>\n")
    for aname, st_list in agents:
        for sname, variables in st_list:
            cols = states[(aname, sname)]
            n = len(cols[variables[0][0]]) if variables else 0
            for i in range(n):
                out.append("<xagent>\n<name>%s</name>\n" % aname)
                for vname, ctype in variables:
                    vals = np.atleast_1d(cols[vname][i])      # an array variable: comma-separated (io.xslt:176-186)
                    out.append("<%s>%s</%s>\n" % (vname, ",".join(_format(ctype, v) for v in vals), vname))
                out.append("</xagent>\n")
    out.append("</states>\n")
    with open(path, "w") as fh:
        fh.write("".join(out))
