"""XMML model loader.

Parses a FLAME GPU 1.5 ``XMLModelFile.xml`` into plain dataclasses.  This
replaces the reference's XSLT front end (the element grammar is the one in
``FLAMEGPU/schemas/XMMLGPU.xsd:24-353``); only what the spatially partitioned
continuous-agent step needs is interpreted, everything else is kept as raw
text so that the generator can reject it with a clear message.

Namespaces follow the reference models: the default namespace is XMML, the
``gpu:`` prefix is XMMLGPU (``examples/*/src/model/XMLModelFile.xml:2``).
"""
from __future__ import annotations

import math
import struct
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional, Union

X = "{http://www.dcs.shef.ac.uk/~paul/XMML}"
G = "{http://www.dcs.shef.ac.uk/~paul/XMMLGPU}"

# scalar C types the hot path supports, with their byte width
SCALAR_TYPES = {
    "int": 4,
    "unsigned int": 4,
    "float": 4,
    "double": 8,
    "long long int": 8,
    "unsigned long long int": 8,
}
INT_TYPES = {"int", "unsigned int", "long long int", "unsigned long long int", "short", "unsigned short"}


class ModelError(ValueError):
    pass


def _f32(v: float) -> float:
    """Round a Python float to IEEE binary32 (the reference computes in float)."""
    return struct.unpack("f", struct.pack("f", v))[0]


def _text(node, tag, default=None):
    if node is None:
        return default
    child = node.find(tag)
    if child is None or child.text is None:
        return default
    return child.text.strip()


@dataclass
class Variable:
    name: str
    type: str
    default: Optional[str] = None
    array_length: int = 0  # 0 = scalar

    @property
    def size(self) -> int:
        if self.type not in SCALAR_TYPES:
            raise ModelError(f"variable '{self.name}': type '{self.type}' is not supported on the hot path")
        return SCALAR_TYPES[self.type]

    @property
    def is_int(self) -> bool:
        return self.type in INT_TYPES

    def default_literal(self) -> str:
        """C literal for the default value (`cmn` defaultInitialiser: defaultValue or 0)."""
        if self.default is not None and self.default != "":
            return self.default
        return "0"


@dataclass
class Condition:
    """Recursive <condition>/<gpu:globalCondition> tree (`_common_templates.xslt:270-301`)."""
    lhs: Union[str, "Condition"]  # C text, 'var:<name>' or nested
    op: str
    rhs: Union[str, "Condition"]

    def c_expr(self, var_fmt: str) -> str:
        """Emit the same fully parenthesised C expression the reference pastes.

        ``var_fmt`` formats an agent variable reference, e.g. ``'cur->{0}[index]'``.
        """
        def side(s):
            if isinstance(s, Condition):
                return s.c_expr(var_fmt)
            if s.startswith("var:"):
                return var_fmt.format(s[4:])
            return s
        return f"({side(self.lhs)}{self.op}{side(self.rhs)})"

    def variables(self) -> List[str]:
        out = []
        for s in (self.lhs, self.rhs):
            if isinstance(s, Condition):
                out += s.variables()
            elif s.startswith("var:"):
                out.append(s[4:])
        return out


@dataclass
class Function:
    name: str
    agent: str
    current_state: str
    next_state: str
    input_message: Optional[str] = None
    output_message: Optional[str] = None
    output_type: str = "single_message"
    xagent_output: Optional[str] = None
    xagent_output_state: Optional[str] = None
    reallocate: bool = True
    reallocate_literal: str = "true"
    rng: bool = False
    condition: Optional[Condition] = None
    global_condition: Optional[Condition] = None
    gc_max_iterations: int = 0
    gc_must_evaluate_to: bool = True


@dataclass
class Agent:
    name: str
    variables: List[Variable]
    functions: List[Function]
    states: List[str]
    initial_state: str
    type: str
    buffer_size: int

    def var(self, name: str) -> Variable:
        for v in self.variables:
            if v.name == name:
                return v
        raise KeyError(name)


@dataclass
class Message:
    name: str
    variables: List[Variable]
    buffer_size: int
    partitioning: str  # 'none' | 'spatial' | 'discrete' | 'graph'
    radius: float = 0.0
    bounds_min: tuple = (0.0, 0.0, 0.0)
    bounds_max: tuple = (0.0, 0.0, 0.0)
    # raw XML literals (the reference pastes them as `(float)<literal>`, sim:534-539)
    radius_literal: str = "0"
    min_literals: tuple = ("0", "0", "0")
    max_literals: tuple = ("0", "0", "0")

    def partition_dims(self):
        """`(int)round((max-min)/radius)` evaluated in float (`simulation.xslt:540-542`)."""
        r = _f32(self.radius)
        dims = []
        for lo, hi in zip(self.bounds_min, self.bounds_max):
            d = _f32(_f32(_f32(hi) - _f32(lo)) / r)
            dims.append(int(math.floor(d + 0.5)) if d >= 0 else -int(math.floor(-d + 0.5)))
        return tuple(dims)

    def grid_size(self) -> int:
        """Allocation size: XSLT `ceiling` in doubles (`header.xslt:102-105`)."""
        n = 1
        for lo, hi in zip(self.bounds_min, self.bounds_max):
            n *= int(math.ceil((hi - lo) / self.radius))
        return n

    def is_2d(self) -> bool:
        """2-D walk iff ceiling((zmax-zmin)/radius) == 1 (`FLAMEGPU_kernals.xslt:1052`)."""
        return int(math.ceil((self.bounds_max[2] - self.bounds_min[2]) / self.radius)) == 1

    def radix_bits(self) -> int:
        """`ceil(log(grid_size)/log(2))` (`simulation.xslt:282`)."""
        g = self.grid_size()
        return max(1, int(math.ceil(math.log(g) / math.log(2)))) if g > 1 else 1


@dataclass
class EnvConstant:
    name: str
    type: str
    default: Optional[str] = None
    array_length: int = 0


@dataclass
class Model:
    name: str
    agents: List[Agent]
    messages: List[Message]
    layers: List[List[str]]
    constants: List[EnvConstant] = field(default_factory=list)
    function_files: List[str] = field(default_factory=list)
    init_functions: List[str] = field(default_factory=list)
    step_functions: List[str] = field(default_factory=list)
    exit_functions: List[str] = field(default_factory=list)

    def agent(self, name: str) -> Agent:
        for a in self.agents:
            if a.name == name:
                return a
        raise KeyError(name)

    def message(self, name: str) -> Message:
        for m in self.messages:
            if m.name == name:
                return m
        raise KeyError(name)

    def function(self, name: str) -> Function:
        for a in self.agents:
            for f in a.functions:
                if f.name == name:
                    return f
        raise KeyError(name)

    def all_functions(self) -> List[Function]:
        return [f for a in self.agents for f in a.functions]

    @property
    def buffer_size_max(self) -> int:
        """`buffer_size_MAX` = largest agent bufferSize (`header.xslt:71-75`)."""
        return max(a.buffer_size for a in self.agents)


def _parse_variables(parent) -> List[Variable]:
    out = []
    if parent is None:
        return out
    for v in parent.findall(G + "variable"):
        al = _text(v, X + "arrayLength")
        out.append(Variable(
            name=_text(v, X + "name"),
            type=" ".join(_text(v, X + "type", "").split()),
            default=_text(v, X + "defaultValue"),
            array_length=int(al) if al else 0,
        ))
    return out


def _parse_condition(node) -> Condition:
    def side(s):
        val = s.find(X + "value")
        if val is not None:
            return (val.text or "").strip()
        av = s.find(X + "agentVariable")
        if av is not None:
            return "var:" + (av.text or "").strip()
        nested = s.find(X + "condition")
        if nested is not None:
            return _parse_condition(nested)
        raise ModelError("condition side has neither value, agentVariable nor condition")
    return Condition(lhs=side(node.find(X + "lhs")), op=_text(node, X + "operator"), rhs=side(node.find(X + "rhs")))


def parse_model(path: str) -> Model:
    root = ET.parse(path).getroot()
    if root.tag != G + "xmodel":
        raise ModelError(f"{path}: root element is {root.tag}, expected gpu:xmodel")

    env = root.find(G + "environment")
    constants, function_files, inits, steps, exits = [], [], [], [], []
    if env is not None:
        cs = env.find(G + "constants")
        if cs is not None:
            for v in cs.findall(G + "variable"):
                al = _text(v, X + "arrayLength")
                constants.append(EnvConstant(_text(v, X + "name"), " ".join(_text(v, X + "type", "").split()),
                                             _text(v, X + "defaultValue"), int(al) if al else 0))
        ff = env.find(G + "functionFiles")
        if ff is not None:
            function_files = [(f.text or "").strip() for f in ff.findall(X + "file")]
        for tag, lst in (("initFunctions", inits), ("stepFunctions", steps), ("exitFunctions", exits)):
            node = env.find(G + tag)
            if node is not None:
                for fn in node:
                    nm = _text(fn, G + "name")
                    if nm:
                        lst.append(nm)
        if env.find(G + "graphs") is not None:
            raise ModelError("static graphs (graph-edge messages) are outside the hot path")

    agents = []
    xagents = root.find(X + "xagents")
    for xa in (xagents.findall(G + "xagent") if xagents is not None else []):
        aname = _text(xa, X + "name")
        variables = _parse_variables(xa.find(X + "memory"))
        functions = []
        fnode = xa.find(X + "functions")
        for fn in (fnode.findall(G + "function") if fnode is not None else []):
            f = Function(name=_text(fn, X + "name"), agent=aname,
                         current_state=_text(fn, X + "currentState"), next_state=_text(fn, X + "nextState"))
            inp = fn.find(X + "inputs")
            if inp is not None and inp.find(G + "input") is not None:
                f.input_message = _text(inp.find(G + "input"), X + "messageName")
            outp = fn.find(X + "outputs")
            if outp is not None and outp.find(G + "output") is not None:
                o = outp.find(G + "output")
                f.output_message = _text(o, X + "messageName")
                f.output_type = _text(o, G + "type", "single_message")
            xo = fn.find(X + "xagentOutputs")
            if xo is not None and xo.find(G + "xagentOutput") is not None:
                o = xo.find(G + "xagentOutput")
                f.xagent_output = _text(o, X + "xagentName")
                f.xagent_output_state = _text(o, X + "state")
            cond = fn.find(X + "condition")
            if cond is not None:
                f.condition = _parse_condition(cond)
            gc = fn.find(G + "globalCondition")
            if gc is not None:
                f.global_condition = _parse_condition(gc)
                f.gc_max_iterations = int(_text(gc, G + "maxItterations", "0"))
                f.gc_must_evaluate_to = (_text(gc, G + "mustEvaluateTo", "true") == "true")
            lit = _text(fn, G + "reallocate", "true")
            f.reallocate_literal = lit
            f.reallocate = (lit == "true")
            f.rng = (_text(fn, G + "RNG", "false") == "true")
            functions.append(f)
        st = xa.find(X + "states")
        states = [_text(s, X + "name") for s in st.findall(G + "state")]
        agents.append(Agent(name=aname, variables=variables, functions=functions, states=states,
                            initial_state=_text(st, X + "initialState"),
                            type=_text(xa, G + "type", "continuous"),
                            buffer_size=int(_text(xa, G + "bufferSize"))))

    messages = []
    mnode = root.find(X + "messages")
    for m in (mnode.findall(G + "message") if mnode is not None else []):
        msg = Message(name=_text(m, X + "name"), variables=_parse_variables(m.find(X + "variables")),
                      buffer_size=int(_text(m, G + "bufferSize")), partitioning="none")
        sp = m.find(G + "partitioningSpatial")
        if sp is not None:
            msg.partitioning = "spatial"
            lits = {k: _text(sp, G + k) for k in ("radius", "xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}
            msg.radius_literal = lits["radius"]
            msg.min_literals = (lits["xmin"], lits["ymin"], lits["zmin"])
            msg.max_literals = (lits["xmax"], lits["ymax"], lits["zmax"])
            msg.radius = float(lits["radius"])
            msg.bounds_min = tuple(float(x) for x in msg.min_literals)
            msg.bounds_max = tuple(float(x) for x in msg.max_literals)
        elif m.find(G + "partitioningDiscrete") is not None:
            msg.partitioning = "discrete"
        elif m.find(G + "partitioningGraphEdge") is not None:
            msg.partitioning = "graph"
        messages.append(msg)

    layers = []
    lnode = root.find(X + "layers")
    for layer in (lnode.findall(X + "layer") if lnode is not None else []):
        layers.append([_text(lf, X + "name") for lf in layer.findall(G + "layerFunction")])

    model = Model(name=(_text(root, X + "name", "model") or "model").strip(), agents=agents, messages=messages,
                  layers=layers, constants=constants, function_files=function_files,
                  init_functions=inits, step_functions=steps, exit_functions=exits)
    validate(model)
    return model


def validate(model: Model) -> None:
    """Generator-time checks the reference raises as ``#error`` (`simulation.xslt:38-188`)
    plus the key/keyref integrity of `XMMLGPU.xsd:153-202`."""
    fnames = {f.name for f in model.all_functions()}
    for layer in model.layers:
        for fn in layer:
            if fn not in fnames:
                raise ModelError(f"layer references function '{fn}' which no agent defines")
    mnames = {m.name for m in model.messages}
    for a in model.agents:
        if a.type != "continuous":
            raise ModelError(f"agent '{a.name}': discrete agents are outside the hot path")
        if a.initial_state not in a.states:
            raise ModelError(f"agent '{a.name}': initialState '{a.initial_state}' is not a declared state")
        for f in a.functions:
            for s in (f.current_state, f.next_state):
                if s not in a.states:
                    raise ModelError(f"function '{f.name}': unknown state '{s}'")
            for mn in (f.input_message, f.output_message):
                if mn is not None and mn not in mnames:
                    raise ModelError(f"function '{f.name}': unknown message '{mn}'")
            if f.xagent_output is not None:
                tgt = model.agent(f.xagent_output)
                if f.xagent_output_state not in tgt.states:
                    raise ModelError(f"function '{f.name}': unknown xagentOutput state")
            if f.condition is not None and f.global_condition is not None:
                raise ModelError(f"function '{f.name}': both condition and globalCondition")
    for m in model.messages:
        if m.partitioning in ("discrete", "graph"):
            raise ModelError(f"message '{m.name}': {m.partitioning} partitioning is outside the hot path")
        for v in m.variables:
            if v.array_length:
                raise ModelError(f"message '{m.name}': array variables are not allowed in messages")
            _ = v.size
        if m.partitioning == "spatial":
            names = [v.name for v in m.variables]
            for ax in ("x", "y", "z"):
                if ax not in names:
                    raise ModelError(f"spatial message '{m.name}' needs float variables x, y, z")
            # radius must be a factor of the bounds (sim:73-99, epsilon 1e-10 in XSLT doubles)
            for lo, hi, ax in zip(m.bounds_min, m.bounds_max, "xyz"):
                q = (hi - lo) / m.radius
                if abs(q - round(q)) > 1e-6:
                    raise ModelError(f"message '{m.name}': radius is not a factor of the {ax} extent")
            dx, dy, dz = m.partition_dims()
            if dx < 3 or dy < 3:
                raise ModelError(f"message '{m.name}': partition grid needs at least 3 cells in x and y")
            if (dx, dy, dz) != tuple(int(math.ceil((hi - lo) / m.radius - 1e-9))
                                     for lo, hi in zip(m.bounds_min, m.bounds_max)):
                raise ModelError(f"message '{m.name}': float and XSLT partition dims disagree")
