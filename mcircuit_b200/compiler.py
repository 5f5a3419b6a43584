"""Netlist compiler (K6): descriptor graph -> flat, levelized, slot-allocated SoA program.

Replaces the front half of the reference's ``JIT.__init__`` and its helpers
(``/root/reference/core/simulator.py:16-74, 175-221``):

  flatten   iter_simulation_pins        (simulator.py:16-25)   -> ``Design.pins``
  connect   iter_simulation_connections (simulator.py:28-38)   -> ``Design.conn_*``
  schedule  iter_simulation_topology    (simulator.py:41-53)   -> ``Design.schedule``
  alias     _map_to_sources             (simulator.py:56-74)   -> ``Design.pin_net``
  lower     _translate_* / TRANSLATOR   (simulator.py:77-157)  -> 1-bit micro-ops
  levelize / slot-allocate / encode                             -> ``Program``

All passes are O(pins + connections); the reference's are O(nodes x connections)
(SURVEY Q13).  The *order* of evaluation is the reference's: the DFS post-order of every
composite's ``graph`` as produced by ``networkx.dfs_postorder_nodes`` (we call the same
function on the same graph object), composites expanded in place.

Execution model handed to the kernels
-------------------------------------
The reference executes one primitive after another, in place, over one global per net.  We
lower every primitive to 1-bit micro-ops that keep *exactly* that sequential, in-place
meaning (each signal has one writer; an op may read the signal it writes), then recover
parallelism:

  * forward read  (writer earlier in the sequence): RAW dependency, level(reader) >
    level(writer); the reader sees this step's value;
  * self read     (op reads the signal it writes): in place, sees last step's value;
  * back read     (writer later in the sequence): sees last step's value.  If the reader's
    level is below the writer's it simply reads the live slot before it is overwritten -
    the writer is pushed one level up when that is all it takes; otherwise the signal gets
    a shadow ``prev`` slot that the fused end-of-step *latch* level refreshes
    (``BUF cur -> prev``), and the late readers are pointed at the shadow.

A record is 16 bytes ``{op | in2<<8, in0, in1, out}`` (slot numbers); records are sorted by
level and ``level_off`` delimits levels.  Slots below ``n_persist`` live for the whole run
(pins that must stay observable, state, inputs, constants); slots at or above it are
scratch, recycled by liveness at level granularity.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import networkx as nx
import numpy as np

from .core.descriptors import (Adder, Clock, Composite, Constant, Counter, ExposedPin,
                               Gate, Not, Register)

# --------------------------------------------------------------------------------------
# opcodes (shared with csrc/mcb200.cu and include/mcb200.h - keep in sync)
# --------------------------------------------------------------------------------------
OP_NOP = 0
OP_BUF = 1    # out = a
OP_AND = 2    # out = a & b
OP_OR = 3     # out = a | b
OP_XOR = 4    # out = a ^ b
OP_ANDN = 5   # out = a & ~b
OP_MUX = 6    # out = (a & b) | (~a & c)
OP_AND3 = 7   # out = a & b & c
OP_OR3 = 8    # out = a | b | c
OP_XOR3 = 9   # out = a ^ b ^ c
OP_MAJ = 10   # out = (a & b) | (a & c) | (b & c)
OP_NEG = 0x80  # flag: complement the result
OP_NAMES = {OP_NOP: "nop", OP_BUF: "buf", OP_AND: "and", OP_OR: "or", OP_XOR: "xor",
            OP_ANDN: "andn", OP_MUX: "mux", OP_AND3: "and3", OP_OR3: "or3",
            OP_XOR3: "xor3", OP_MAJ: "maj"}
OP_ARITY = {OP_NOP: 0, OP_BUF: 1, OP_AND: 2, OP_OR: 2, OP_XOR: 2, OP_ANDN: 2, OP_MUX: 3,
            OP_AND3: 3, OP_OR3: 3, OP_XOR3: 3, OP_MAJ: 3}
IN2_LIMIT = 1 << 24   # in2 shares a word with the opcode

KIND_IN, KIND_OUT, KIND_INTERNAL = 0, 1, 2
_KIND_NAME = ("in", "out", "internal")


class CircuitError(ValueError):
    """Build-time validation failure (width mismatch, multi-driver net, wire loop)."""


class PinNotObservable(KeyError):
    """The pin exists but its value is not retained (``keep='ports'``)."""


# --------------------------------------------------------------------------------------
# per-descriptor pin tables
# --------------------------------------------------------------------------------------
class _PinTable:
    __slots__ = ("names", "widths", "kinds", "index")

    def __init__(self, desc):
        self.names: List[str] = []
        self.widths: List[int] = []
        self.kinds: List[int] = []
        for kind, gen in ((KIND_IN, desc.all_inputs()), (KIND_OUT, desc.all_outputs()),
                          (KIND_INTERNAL, desc.all_internals())):
            for name, width in gen:
                self.names.append(name)
                self.widths.append(int(width))
                self.kinds.append(kind)
        # first occurrence wins, like a path-keyed table built in iteration order would
        self.index: Dict[str, int] = {}
        for k, n in enumerate(self.names):
            self.index.setdefault(n, k)


class _Instance:
    """One instantiation of a Composite in the flattened hierarchy."""
    __slots__ = ("node", "path", "children")

    def __init__(self, node, path):
        self.node = node
        self.path = path
        self.children: Dict[str, object] = {}   # name -> leaf index (int) | _Instance


def _is_composite(desc) -> bool:
    # duck-typed so that descriptors made by the reference's own module also compile
    return hasattr(desc, "graph") and hasattr(desc, "connections")


def _type_name(desc) -> str:
    return type(desc).__name__


# --------------------------------------------------------------------------------------
# Design: the flattened, aliased circuit (still word-level, still in reference order)
# --------------------------------------------------------------------------------------
@dataclass
class Design:
    root: object
    map_pins: bool
    leaf_desc: List[object] = field(default_factory=list)
    leaf_path: List[str] = field(default_factory=list)       # '/a/b/' prefix of the leaf
    leaf_pin_base: List[int] = field(default_factory=list)
    leaf_table: List[_PinTable] = field(default_factory=list)
    pin_leaf: List[int] = field(default_factory=list)
    pin_width: List[int] = field(default_factory=list)
    pin_kind: List[int] = field(default_factory=list)
    top: Optional[_Instance] = None
    # connections in reference iteration order (simulator.py:28-38)
    conn_src: List[int] = field(default_factory=list)
    conn_dst: List[int] = field(default_factory=list)
    # schedule: ('emit', leaf) / ('propagate', conn index) in reference order (:41-53)
    schedule: List[Tuple[int, int]] = field(default_factory=list)
    pin_src: Optional[np.ndarray] = None      # ultimate source pin or -1 (_map_to_sources)
    pin_net: Optional[np.ndarray] = None      # pin -> net id
    net_width: List[int] = field(default_factory=list)
    net_sig: Optional[np.ndarray] = None      # net -> first signal id
    n_signals: int = 0

    # ---- names -------------------------------------------------------------------------
    def pin_path(self, pin: int) -> str:
        leaf = self.pin_leaf[pin]
        return self.leaf_path[leaf] + self.leaf_table[leaf].names[pin - self.leaf_pin_base[leaf]]

    def find_pin(self, path: str) -> int:
        """Pin index for an absolute pin path; KeyError if there is no such pin
        (reference: KeyError from ``_pin_map[path]``, simulator.py:278-283)."""
        if not isinstance(path, str) or not path.startswith("/"):
            raise KeyError(path)
        parts = path[1:].split("/")
        inst = self.top
        # longest-prefix walk through composite instances; the remainder is 'leaf/pin'
        i = 0
        while i < len(parts) - 2:
            child = inst.children.get(parts[i])
            if not isinstance(child, _Instance):
                raise KeyError(path)
            inst = child
            i += 1
        if len(parts) - i != 2:
            raise KeyError(path)
        leaf = inst.children.get(parts[i])
        if leaf is None or isinstance(leaf, _Instance):
            raise KeyError(path)
        k = self.leaf_table[leaf].index.get(parts[i + 1])
        if k is None:
            raise KeyError(path)
        return self.leaf_pin_base[leaf] + k

    def iter_pins(self):
        """(abs_path, width, kind) in the reference's order (simulator.py:16-25)."""
        for pin in range(len(self.pin_width)):
            yield self.pin_path(pin), self.pin_width[pin], _KIND_NAME[self.pin_kind[pin]]


_SCHED_EMIT, _SCHED_PROP = 0, 1


def flatten(root, map_pins: bool = True) -> Design:
    """Flatten + connect + schedule + alias.  Linear in pins + connections."""
    if not _is_composite(root):
        raise TypeError("root must be a Composite")
    d = Design(root=root, map_pins=bool(map_pins))
    tables: Dict[int, _PinTable] = {}
    keepalive = []          # id() keys stay valid while the objects are referenced

    def table_of(desc) -> _PinTable:
        t = tables.get(id(desc))
        if t is None:
            t = _PinTable(desc)
            tables[id(desc)] = t
            keepalive.append(desc)
        return t

    # ---- pass 1: leaves and pins in graph.nodes order, depth first (simulator.py:16-25) --
    def walk_pins(node, path) -> _Instance:
        inst = _Instance(node, path)
        get_child = node.get_child
        for name in node.graph.nodes:
            desc = get_child(name)
            if _is_composite(desc):
                inst.children[name] = walk_pins(desc, path + name + "/")
                continue
            leaf = len(d.leaf_desc)
            t = table_of(desc)
            d.leaf_desc.append(desc)
            d.leaf_path.append(path + name + "/")
            d.leaf_pin_base.append(len(d.pin_width))
            d.leaf_table.append(t)
            d.pin_leaf.extend([leaf] * len(t.names))
            d.pin_width.extend(t.widths)
            d.pin_kind.extend(t.kinds)
            inst.children[name] = leaf
        return inst

    d.top = walk_pins(root, "/")

    # ---- per-Composite-object caches: grouped connections and DFS post-order -----------
    grouped: Dict[int, Dict[str, list]] = {}
    postorder: Dict[int, list] = {}

    def groups_of(node):
        g = grouped.get(id(node))
        if g is None:
            g = {}
            for conn in node.connections:      # set order; stable within this process
                c0 = conn[0]
                key = c0 if "/" not in c0 else c0.split("/")[0]
                lst = g.get(key)
                if lst is None:
                    g[key] = [conn]
                else:
                    lst.append(conn)
            grouped[id(node)] = g
            keepalive.append(node)
        return g

    def post_of(node):
        p = postorder.get(id(node))
        if p is None:
            p = list(nx.dfs_postorder_nodes(node.graph))
            postorder[id(node)] = p
        return p

    leaf_table, leaf_pin_base = d.leaf_table, d.leaf_pin_base

    def resolve(inst: _Instance, cname: str, pname: str) -> int:
        leaf = inst.children.get(cname)                   # fast path: a direct leaf child
        if leaf.__class__ is int:
            k = leaf_table[leaf].index.get(pname)
            if k is None:
                raise KeyError(inst.path + cname + "/" + pname)
            return leaf_pin_base[leaf] + k
        here = inst
        parts = cname.split("/")
        for part in parts[:-1]:
            here = here.children.get(part)
            if not isinstance(here, _Instance):
                raise KeyError(inst.path + cname + "/" + pname)
        leaf = here.children.get(parts[-1])
        if leaf is None or isinstance(leaf, _Instance):
            raise KeyError(inst.path + cname + "/" + pname)
        k = d.leaf_table[leaf].index.get(pname)
        if k is None:
            raise KeyError(inst.path + cname + "/" + pname)
        return d.leaf_pin_base[leaf] + k

    # ---- pass 2: connections in the reference's enumeration order (simulator.py:28-38) --
    conn_first: Dict[int, Dict[str, int]] = {}     # id(instance) -> child name -> first conn index
    conn_src, conn_dst = d.conn_src, d.conn_dst

    def walk_conns(inst: _Instance):
        node = inst.node
        groups = groups_of(node)
        first = conn_first[id(inst)] = {}
        children = inst.children
        for name in node.graph.nodes:
            child = children[name]
            if child.__class__ is _Instance:
                walk_conns(child)
            grp = groups.get(name)
            if grp:
                first[name] = len(conn_src)
                for conn in grp:
                    conn_src.append(resolve(inst, conn[0], conn[1]))
                    conn_dst.append(resolve(inst, conn[2], conn[3]))

    walk_conns(d.top)

    # ---- pass 3: the schedule (simulator.py:41-53) ---------------------------------------
    sched = d.schedule

    want_props = not d.map_pins      # aliased globals make every propagate a self-copy (Q7)

    def walk_topo(inst: _Instance):
        node = inst.node
        groups = groups_of(node)
        first = conn_first[id(inst)]
        children = inst.children
        for name in post_of(node):
            child = children[name]
            if child.__class__ is _Instance:
                walk_topo(child)
            else:
                sched.append((_SCHED_EMIT, child))
            if want_props:
                grp = groups.get(name)
                if grp:
                    k0 = first[name]
                    for k in range(len(grp)):
                        sched.append((_SCHED_PROP, k0 + k))

    walk_topo(d.top)

    # ---- pass 4: aliasing (simulator.py:56-74) -------------------------------------------
    n_pins = len(d.pin_width)
    drv = [-1] * n_pins
    for s_, t in zip(conn_src, conn_dst):
        if drv[t] != -1 and drv[t] != s_:
            raise CircuitError(
                "pin %s is driven by both %s and %s; the reference keeps whichever the set "
                "iteration yields last (hash-seed dependent), so this is rejected"
                % (d.pin_path(t), d.pin_path(drv[t]), d.pin_path(s_)))
        drv[t] = s_
    state = bytearray(n_pins)                      # 0 unknown, 1 on the current chain, 2 done
    root_of = list(range(n_pins))
    for p in range(n_pins):
        if state[p] == 2:
            continue
        if drv[p] == -1:
            state[p] = 2
            continue
        chain = []
        cur = p
        while state[cur] == 0 and drv[cur] != -1:
            state[cur] = 1
            chain.append(cur)
            cur = drv[cur]
        if state[cur] == 1:
            raise CircuitError("wire loop through %s: the reference's _map_to_sources "
                               "(simulator.py:65-67) never terminates on it" % d.pin_path(cur))
        r = root_of[cur] if state[cur] == 2 else cur
        state[cur] = 2
        root_of[cur] = r
        for c in chain:
            root_of[c] = r
            state[c] = 2
    roots = np.asarray(root_of, dtype=np.int64)
    driven = np.asarray(drv, dtype=np.int64) != -1
    d.pin_src = np.where(driven, roots, -1)

    # ---- nets ----------------------------------------------------------------------------
    pw = np.asarray(d.pin_width, dtype=np.int64)
    if d.map_pins:
        # net ids in order of first appearance of their root among the pins
        uniq, first_idx, inverse = np.unique(roots, return_index=True, return_inverse=True)
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[np.argsort(first_idx, kind="stable")] = np.arange(len(uniq))
        pin_net = rank[inverse]
        net_root = np.empty(len(uniq), dtype=np.int64)
        net_root[rank] = uniq
        net_w = pw[net_root]
        bad = np.nonzero(pw != net_w[pin_net])[0]
        if len(bad):
            pb = int(bad[0])
            rb = int(roots[pb])
            raise CircuitError(
                "width mismatch: %s is %d bits but its source %s is %d bits (the "
                "reference fails in llvmlite with a type error here)"
                % (d.pin_path(pb), d.pin_width[pb], d.pin_path(rb), d.pin_width[rb]))
        d.pin_net = pin_net
        d.net_width = net_w.tolist()
    else:
        d.pin_net = np.arange(n_pins, dtype=np.int64)
        d.net_width = list(d.pin_width)
        for s_, t in zip(conn_src, conn_dst):
            if d.pin_width[s_] != d.pin_width[t]:
                raise CircuitError("width mismatch: %s (%d) -> %s (%d)" % (
                    d.pin_path(s_), d.pin_width[s_], d.pin_path(t), d.pin_width[t]))
    widths = np.asarray(d.net_width, dtype=np.int64)
    d.net_sig = np.concatenate(([0], np.cumsum(widths)))[:-1] if len(widths) else np.zeros(0, np.int64)
    d.n_signals = int(widths.sum())
    d._keepalive = keepalive
    return d


# --------------------------------------------------------------------------------------
# Lowering to 1-bit micro-ops (simulator.py:77-147)
# --------------------------------------------------------------------------------------
class _Ops:
    """Growing micro-op list; signals >= first_temp are compiler temporaries."""

    def __init__(self, n_signals: int, allow3: bool):
        self.code: List[int] = []
        self.a: List[int] = []
        self.b: List[int] = []
        self.c: List[int] = []
        self.out: List[int] = []
        self.is_gate: List[bool] = []     # counts toward gate-evals (not wire copies)
        self.n_sig = n_signals
        self.first_temp = n_signals
        self.allow3 = allow3
        self.dead = set()                 # pin signals orphaned by edge-detector sharing
        self.net_sig = None               # this compile's net -> first signal map (a COPY of the design's)
        self.shared_prev = set()          # prevclock signals that stand for a whole group of Registers / Counters

    def temp(self) -> int:
        t = self.n_sig
        self.n_sig += 1
        return t

    def emit(self, code, out, a=-1, b=-1, c=-1, gate=True):
        self.code.append(code)
        self.a.append(a)
        self.b.append(b)
        self.c.append(c)
        self.out.append(out)
        self.is_gate.append(gate)
        return out


_GATE2 = {0: OP_AND, 1: OP_OR, 2: OP_XOR}
_GATE3 = {0: OP_AND3, 1: OP_OR3, 2: OP_XOR3}


def _lower_gate(ops: _Ops, desc, ins: Sequence[Sequence[int]], out: Sequence[int]):
    """Left fold of ``op`` over in0..in{n-1}, optional final NOT (simulator.py:83-97).
    An ``op`` outside {AND, OR, XOR} leaves the first input untouched, as the reference's
    if/elif chain does."""
    n = len(ins)
    if n == 0:
        raise CircuitError("Gate with num_inputs=0 (the reference crashes storing None)")
    neg = OP_NEG if desc.negated else 0
    op2 = _GATE2.get(desc.op)
    op3 = _GATE3.get(desc.op)
    for bit in range(len(out)):
        srcs = [p[bit] for p in ins]
        if op2 is None or n == 1:
            ops.emit(OP_BUF | neg, out[bit], srcs[0])
            continue
        acc = srcs[0]
        i = 1
        while i < n:
            take3 = ops.allow3 and (n - i) >= 2
            last = (i + (2 if take3 else 1)) >= n
            dst = out[bit] if last else ops.temp()
            if take3:
                ops.emit(op3 | (neg if last else 0), dst, acc, srcs[i], srcs[i + 1])
                i += 2
            else:
                ops.emit(op2 | (neg if last else 0), dst, acc, srcs[i])
                i += 1
            acc = dst


def _lower_not(ops, desc, inp, out):
    for bit in range(len(out)):                      # simulator.py:77-80
        ops.emit(OP_BUF | OP_NEG, out[bit], inp[bit])


def _lower_clock(ops, desc, out):
    ops.emit(OP_BUF | OP_NEG, out[0], out[0])        # out = ~out, simulator.py:130-132


def _lower_counter(ops, desc, clock, out, prevclock, shared_rise=None):
    # rise = ~prevclock & clock, zero-extended: only bit 0 of `out` is ever muxed
    # (simulator.py:100-112); bits 1.. hold.  prevclock = clock, re-read last.
    if shared_rise is not None:
        ops.emit(OP_XOR, out[0], out[0], shared_rise)
        return shared_rise
    rise = ops.emit(OP_ANDN, ops.temp(), clock[0], prevclock[0])
    ops.emit(OP_XOR, out[0], out[0], rise)
    ops.emit(OP_BUF, prevclock[0], clock[0])
    return rise


def _lower_register(ops, desc, clock, data, out, prevclock, shared_rise=None):
    # Intended semantics of simulator.py:115-127 (the reference's emitter does not
    # compile): out = rise ? data : out ; prevclock = clock.
    rise = shared_rise if shared_rise is not None else ops.emit(OP_ANDN, ops.temp(), clock[0], prevclock[0])
    for bit in range(len(out)):
        if ops.allow3:
            ops.emit(OP_MUX, out[bit], rise, data[bit], out[bit])
        else:
            t1 = ops.emit(OP_AND, ops.temp(), rise, data[bit])
            t2 = ops.emit(OP_ANDN, ops.temp(), out[bit], rise)
            ops.emit(OP_OR, out[bit], t1, t2)
    if shared_rise is None:
        ops.emit(OP_BUF, prevclock[0], clock[0])
    return rise


def _prune_dead_temps(ops, start, live):
    """Drop the micro-ops emitted since ``start`` whose results (compiler temporaries) nobody reads:
    a textbook prefix network computes group-propagate terms that later levels never use."""
    live = set(live)
    keep = []
    for i in range(len(ops.code) - 1, start - 1, -1):
        o = ops.out[i]
        if o in live or o < ops.first_temp:
            keep.append(i)
            live.update(x for x in (ops.a[i], ops.b[i], ops.c[i]) if x >= 0)
    keep.reverse()
    for name in ("code", "a", "b", "c", "out", "is_gate"):
        lst = getattr(ops, name)
        lst[start:] = [lst[i] for i in keep]


def _prefix_sum(ops, a, b):
    """s1 = a + b mod 2^w as a parallel-prefix (Kogge-Stone) network: depth log2(w) + 2 instead of w.
    (g, p) = (a & b, a ^ b) per bit; groups combine as G = P_hi ? G_lo : G_hi (g and p are disjoint, so
    the select is exact), P = P_hi & P_lo; the carry into bit i is G of bits 0..i-1."""
    w = len(a)
    start = len(ops.code)
    p0 = [ops.emit(OP_XOR, ops.temp(), a[i], b[i]) for i in range(w)]
    G = [ops.emit(OP_AND, ops.temp(), a[i], b[i]) for i in range(w - 1)]       # the top bit's carry-out is not needed
    P = list(p0[:w - 1])
    d = 1
    while d < w - 1:
        nG, nP = list(G), list(P)
        for i in range(d, w - 1):
            if ops.allow3:
                nG[i] = ops.emit(OP_MUX, ops.temp(), P[i], G[i - d], G[i])
            else:
                t = ops.emit(OP_AND, ops.temp(), P[i], G[i - d])
                nG[i] = ops.emit(OP_OR, ops.temp(), G[i], t)
            nP[i] = ops.emit(OP_AND, ops.temp(), P[i], P[i - d])
        G, P = nG, nP
        d *= 2
    s1 = [p0[0]] + [ops.emit(OP_XOR, ops.temp(), p0[i], G[i - 1]) for i in range(1, w)]
    _prune_dead_temps(ops, start, s1)
    return s1


def _prefix_increment(ops, s1, k, dsts):
    """s2 = s1 + k (k one bit) with the carries k & s1[0] & ... & s1[i-1] as a prefix-AND network."""
    w = len(s1)
    A = list(s1[:w - 1])                      # A[i] = AND of s1[0..i] after the sweep
    d = 1
    while d < w - 1:
        nA = list(A)
        for i in range(d, w - 1):
            nA[i] = ops.emit(OP_AND, ops.temp(), A[i], A[i - d])
        A = nA
        d *= 2
    s2 = [ops.emit(OP_XOR, dsts[0], s1[0], k)]
    for i in range(1, w):
        c = ops.emit(OP_AND, ops.temp(), A[i - 1], k)
        s2.append(ops.emit(OP_XOR, dsts[i], s1[i], c))
    return s2


def _lower_adder(ops, desc, cin, a, b, cout, sum_, prefix=False):
    """sum = (a + b) + zext(cin) mod 2^w ; cout = sovf(a + b) | sovf(s1 + zext(cin))
    (simulator.py:135-147).  The reference loads a, b and cin before it stores
    anything, so when an output net is also one of this adder's inputs the results go to
    temporaries first.  ``prefix``: the two carry chains as parallel-prefix networks (word-level
    lowering of the primitive: same function, logarithmic depth)."""
    w = len(a)
    m = w - 1
    aliased = bool(set(sum_) & (set(a) | set(b) | set(cin))) or bool(
        set(cout) & (set(a) | set(b) | set(cin)))
    s1 = [0] * w
    carry = -1
    if prefix and w > 2:
        s1 = _prefix_sum(ops, a, b)
    for i in range(w if not (prefix and w > 2) else 0):
        if i == 0:
            s1[0] = ops.emit(OP_XOR, ops.temp(), a[0], b[0])
            if w > 1:
                carry = ops.emit(OP_AND, ops.temp(), a[0], b[0])
        else:
            if ops.allow3:
                s1[i] = ops.emit(OP_XOR3, ops.temp(), a[i], b[i], carry)
                if i < m:
                    carry = ops.emit(OP_MAJ, ops.temp(), a[i], b[i], carry)
            else:
                x = ops.emit(OP_XOR, ops.temp(), a[i], b[i])
                s1[i] = ops.emit(OP_XOR, ops.temp(), x, carry)
                if i < m:
                    g = ops.emit(OP_AND, ops.temp(), a[i], b[i])
                    p = ops.emit(OP_AND, ops.temp(), x, carry)
                    carry = ops.emit(OP_OR, ops.temp(), g, p)
    # signed overflow of a + b: operands agree in sign, result differs
    t1 = ops.emit(OP_XOR | OP_NEG, ops.temp(), a[m], b[m])
    t2 = ops.emit(OP_XOR, ops.temp(), a[m], s1[m])
    ovf1 = ops.emit(OP_AND, ops.temp(), t1, t2)
    # s2 = s1 + zext(cin)
    s2 = [0] * w
    k = cin[0]
    if prefix and w > 2:
        s2 = _prefix_increment(ops, s1, k, [ops.temp() if aliased else sum_[i] for i in range(w)])
    for i in range(w if not (prefix and w > 2) else 0):
        dst = ops.temp() if aliased else sum_[i]
        knext = -1
        if i < m:
            knext = ops.emit(OP_AND, ops.temp(), s1[i], k)
        s2[i] = ops.emit(OP_XOR, dst, s1[i], k)
        k = knext
    # signed overflow of s1 + zext(cin): for w > 1 the addend's sign bit is 0, so it
    # overflows iff s1 >= 0 and s2 < 0; for w == 1 zext(cin) *is* the sign bit.
    if w > 1:
        s2m = s2[m]
        ovf2 = ops.emit(OP_ANDN, ops.temp(), s2m, s1[m])
    else:
        ovf2 = ops.emit(OP_AND, ops.temp(), s1[0], cin[0])
    if aliased:
        co = ops.emit(OP_OR, ops.temp(), ovf1, ovf2)
        for i in range(w):
            ops.emit(OP_BUF, sum_[i], s2[i], gate=False)
        ops.emit(OP_BUF, cout[0], co, gate=False)
    else:
        ops.emit(OP_OR, cout[0], ovf1, ovf2)


def lower(design: Design, allow3: bool = True, share_edge_detectors: bool = False, adder_prefix: bool = False) -> _Ops:
    """``share_edge_detectors``: Registers / Counters that read the same clock net in the same
    epoch (all before, or all after, the op that writes it) compute the same ``rise`` and hold
    the same ``prevclock`` for ever, so one ANDN + one BUF serve the whole group and the
    members' ``prevclock`` pins become views of the leader's bit.  Exact unless a member's
    ``prevclock`` is poked individually (a poke then reaches the whole group) - hence opt-in."""
    ops = _Ops(design.n_signals, allow3)
    # sharing redirects the members' prevclock nets to the leader's bit: do that on a private
    # copy, so compiling one Design twice (or with and without sharing) gives the same result
    net_sig = ops.net_sig = design.net_sig.copy()
    pin_net = design.pin_net
    pin_width = design.pin_width
    # (clock signal, number of writes to it so far this step) -> (rise, prevclock): members that
    # read the clock net between different writes of it see different values and never share
    groups: Dict[Tuple[int, int], Tuple[int, int]] = {}
    written: Dict[int, int] = {}
    n_seen = 0
    pins_per_net = np.bincount(pin_net, minlength=len(design.net_width)) if share_edge_detectors else None

    def clocked(leaf, lower_fn, p, *mid):
        """Lower a Register / Counter, sharing the edge detector of its clock group."""
        if not share_edge_detectors:
            lower_fn(ops, design.leaf_desc[leaf], p["clock"], *mid, p["prevclock"])
            return
        prev_pin = design.leaf_pin_base[leaf] + design.leaf_table[leaf].index["prevclock"]
        alone = pins_per_net[pin_net[prev_pin]] == 1            # nothing else wired to prevclock
        key = (p["clock"][0], written.get(p["clock"][0], 0))
        hit = groups.get(key) if alone else None
        if hit is not None:
            rise, leader_prev = hit
            ops.shared_prev.add(int(leader_prev))
            ops.dead.add(p["prevclock"][0])
            net_sig[pin_net[prev_pin]] = leader_prev            # the pin now views the leader's bit
            lower_fn(ops, design.leaf_desc[leaf], p["clock"], *mid, range(leader_prev, leader_prev + 1),
                     shared_rise=rise)
        else:
            rise = lower_fn(ops, design.leaf_desc[leaf], p["clock"], *mid, p["prevclock"])
            if alone:
                groups[key] = (rise, p["prevclock"][0])

    def sig(pin: int) -> range:
        base = int(net_sig[pin_net[pin]])
        return range(base, base + pin_width[pin])

    def pins(leaf: int) -> Dict[str, range]:
        base = design.leaf_pin_base[leaf]
        return {n: sig(base + k) for n, k in design.leaf_table[leaf].index.items()}

    for kind, arg in design.schedule:
        if kind == _SCHED_PROP:
            if design.map_pins:
                continue                     # self-copy of an aliased global
            s, t = sig(design.conn_src[arg]), sig(design.conn_dst[arg])
            for x, y in zip(s, t):
                ops.emit(OP_BUF, y, x, gate=False)
            continue
        desc = design.leaf_desc[arg]
        tname = _type_name(desc)
        if tname == "Gate":
            base = design.leaf_pin_base[arg]
            n = desc.num_inputs
            ins = [sig(base + i) for i in range(n)]
            _lower_gate(ops, desc, ins, sig(base + n))
        elif tname == "Not":
            p = pins(arg)
            _lower_not(ops, desc, p["in"], p["out"])
        elif tname == "Clock":
            _lower_clock(ops, desc, pins(arg)["out"])
        elif tname == "Counter":
            p = pins(arg)
            clocked(arg, _lower_counter, p, p["out"])
        elif tname == "Register":
            p = pins(arg)
            clocked(arg, _lower_register, p, p["data"], p["out"])
        elif tname == "Adder":
            p = pins(arg)
            _lower_adder(ops, desc, p["cin"], p["a"], p["b"], p["cout"], p["sum"], prefix=adder_prefix)
        # ExposedPin, Constant and anything not in TRANSLATOR emit nothing (:214-221)
        if share_edge_detectors:
            for o in ops.out[n_seen:]:
                written[o] = written.get(o, 0) + 1
            n_seen = len(ops.out)
    return ops


# --------------------------------------------------------------------------------------
# Program
# --------------------------------------------------------------------------------------
@dataclass
class Program:
    design: Design
    records: np.ndarray            # uint32 [n_records, 4]: op|in2<<8, in0, in1, out (slots)
    level_off: np.ndarray          # int32 [n_levels + 1]; the last level is the latch level
    n_levels: int                  # including the latch level when has_latch
    has_latch: bool
    n_persist: int                 # slots [0, n_persist) live for the whole run
    n_scratch: int                 # slots [n_persist, n_persist + n_scratch) are recycled
    sig_slot: np.ndarray           # int64 [n_signals incl. temps]; -1 when never materialised
    sig_shadow: Dict[int, int]     # signal -> slot of its 'prev' shadow (latched)
    sig_kept: np.ndarray           # bool  [n_pin_signals]: value observable after a step
    n_gates: int                   # micro-ops per step that count as gate evaluations
    n_ops: int                     # all records per step incl. wire copies and latches
    gate_words_bytes: int          # algorithmic bytes per lane-word per step
    constants: List[Tuple[int, int, int]]   # (first signal, width, value) poked after build
    max_level_gates: int = 0
    records_hinted: Optional[np.ndarray] = None   # records + streaming hints (bit 31 of in0/in1)
    net_sig: Optional[np.ndarray] = None          # net -> first signal of THIS compile (edge-detector sharing redirects some)
    serial: bool = False                          # records are a serial stream (compile_serial), not levels
    shared_prev: Optional[set] = None             # prevclock signals shared by a group of clocked primitives

    @property
    def n_slots(self) -> int:
        return self.n_persist + self.n_scratch

    # -- pin access helpers -------------------------------------------------------------
    def pin_signals(self, path: str) -> Tuple[int, int]:
        pin = self.design.find_pin(path)
        net = int(self.design.pin_net[pin])
        net_sig = self.net_sig if self.net_sig is not None else self.design.net_sig
        return int(net_sig[net]), self.design.pin_width[pin]

    def pin_slots(self, path: str, need_value: bool = True) -> List[int]:
        base, width = self.pin_signals(path)
        if need_value and not bool(self.sig_kept[base:base + width].all()):
            raise PinNotObservable(
                "%s is an internal net whose storage is recycled (keep='ports'); build the "
                "engine with keep='all' or observe=[...] to read it" % path)
        return [int(s) for s in self.sig_slot[base:base + width]]

    def level_sizes(self) -> np.ndarray:
        return np.diff(self.level_off)


def compile_design(design: Design, keep="all", observe: Sequence[str] = (),
                   allow3: Optional[bool] = None, back_edges="auto",
                   share_edge_detectors: bool = False, sort_level_by_fanin: bool = False,
                   adder_prefix: bool = False) -> Program:
    """Lower, levelize, allocate slots and encode.  ``keep``: 'all' retains every pin net
    (full drop-in observability); 'ports' retains only what a later step or the caller can
    still need - undriven nets (inputs, constants), state, root-level ports and
    ``observe`` - and recycles the rest."""
    if keep not in ("all", "ports"):
        raise ValueError("keep must be 'all' or 'ports'")
    # how a read of last step's value is honoured when its reader would otherwise run at or
    # after the writer: 'order' always delays the writer, 'latch' always uses a shadow slot +
    # the end-of-step latch level, 'auto' delays the writer when that costs one level.
    if back_edges not in ("auto", "order", "latch"):
        raise ValueError("back_edges must be 'auto', 'order' or 'latch'")
    latch_slack = {"auto": 1, "order": 1 << 30, "latch": 0}[back_edges]
    if allow3 is None:
        allow3 = design.n_signals * 2 + 1024 < IN2_LIMIT
    ops = lower(design, allow3, share_edge_detectors, adder_prefix)
    n_ops = len(ops.code)
    n_sig = ops.n_sig
    n_pin_sig = ops.first_temp
    code, A, B, C, OUT = ops.code, ops.a, ops.b, ops.c, ops.out

    # ---- single writer per signal --------------------------------------------------------
    # The reference happily lets several primitives (or, with map_pins=False, a primitive and a
    # propagate) store to one global one after the other.  Keep that sequential meaning by
    # versioning: every write but the LAST of a step goes to a fresh temporary, readers between
    # two writes are pointed at the version that is live at their place in the sequence, and
    # the last write lands in the net's own (persistent) signal - which is also what readers
    # before the first write see: last step's final value.
    n_writers: Dict[int, int] = {}
    for o in OUT:
        n_writers[o] = n_writers.get(o, 0) + 1
    multi = {s_ for s_, k in n_writers.items() if k > 1}
    if multi:
        seen: Dict[int, int] = {}            # writes so far this step
        live: Dict[int, int] = {}            # signal -> version readers should use now
        for i in range(n_ops):
            for arr in (A, B, C):
                s_ = arr[i]
                if s_ in live:
                    arr[i] = live[s_]
            o = OUT[i]
            if o in multi:
                seen[o] = seen.get(o, 0) + 1
                if seen[o] < n_writers[o]:
                    v = n_sig
                    n_sig += 1
                    OUT[i] = v
                    live[o] = v
                else:
                    live.pop(o, None)        # the final write: readers use the net itself again
        ops.n_sig = n_sig
    writer = [-1] * n_sig
    for i in range(n_ops):
        o = OUT[i]
        if writer[o] != -1:
            raise AssertionError("net bit %s still has more than one driver" % _sig_name(design, o))
        writer[o] = i

    # ---- levels: RAW edges, and back reads settled by ordering or by a shadow ----------------
    # An op that reads the OLD value of a signal (its writer comes later in the sequence) only
    # has to run before that writer.  When the writer is reached we know the levels of all such
    # readers: if pushing the writer just above them costs at most `latch_slack` levels we do
    # that (no copy, no extra slot); otherwise the late readers are pointed at a shadow slot
    # that the end-of-step latch level refreshes.
    level = [0] * n_ops
    has_state_read = bytearray(n_sig)                   # read before written within a step
    old_readers: Dict[int, List[Tuple[int, int]]] = {}  # signal -> [(op, operand index)]
    shadow_sig: Dict[int, int] = {}
    operand = (A, B, C)
    for i in range(n_ops):
        lv = 0
        for k, s in enumerate((A[i], B[i], C[i])):
            if s < 0:
                continue
            w = writer[s]
            if w == -1:
                continue
            if w < i:
                lw = level[w] + 1
                if lw > lv:
                    lv = lw
            else:
                has_state_read[s] = 1
                if w > i:
                    old_readers.setdefault(s, []).append((i, k))
        pending = old_readers.pop(OUT[i], None)
        if pending:
            need = 1 + max(level[r] for r, _ in pending)
            if need > lv:
                if need - lv <= latch_slack:
                    lv = need                              # order: writer after its old readers
                else:
                    o = OUT[i]
                    sh = n_sig
                    n_sig += 1
                    shadow_sig[o] = sh
                    writer.append(-2)
                    for r, k in pending:
                        if level[r] >= lv:
                            operand[k][r] = sh             # latch: late readers see the shadow
        level[i] = lv
    n_eval_levels = (max(level) + 1) if n_ops else 0
    has_latch = bool(shadow_sig)
    latch_level = n_eval_levels
    for s, sh in shadow_sig.items():
        code.append(OP_BUF)
        A.append(s)
        B.append(-1)
        C.append(-1)
        OUT.append(sh)
        ops.is_gate.append(False)
        level.append(latch_level)
    n_total = len(code)
    n_levels = n_eval_levels + (1 if has_latch else 0)

    # ---- which signals are persistent ------------------------------------------------------
    persistent = bytearray(n_sig)
    kept = np.zeros(n_pin_sig, dtype=bool)
    dead = ops.dead
    for s in range(n_pin_sig):
        if s in dead:
            continue                                    # orphaned by edge-detector sharing
        if keep == "all" or writer[s] == -1 or has_state_read[s]:
            persistent[s] = 1
            kept[s] = True
    for s, sh in shadow_sig.items():
        persistent[s] = 1
        persistent[sh] = 1
        if s < n_pin_sig:
            kept[s] = True
    if keep == "ports":
        want = []
        top = design.top
        for name, child in top.children.items():
            if not isinstance(child, _Instance) and _type_name(design.leaf_desc[child]) == "ExposedPin":
                want.append(design.leaf_pin_base[child])
        for path in observe:
            want.append(design.find_pin(path))
        for pin in want:
            base = int(ops.net_sig[design.pin_net[pin]])
            for s in range(base, base + design.pin_width[pin]):
                persistent[s] = 1
                kept[s] = True
    # a state-read signal that is a compiler temp cannot exist (temps are SSA, forward only)

    # ---- order ops by (level, sequence) ------------------------------------------------------
    level_arr = np.asarray(level, dtype=np.int64)
    order = np.argsort(level_arr, kind="stable")
    counts = np.bincount(level_arr, minlength=n_levels) if n_total else np.zeros(0, np.int64)
    level_off = np.zeros(n_levels + 1, dtype=np.int64)
    np.cumsum(counts, out=level_off[1:])

    # ---- slots ------------------------------------------------------------------------------
    sig_slot = np.full(n_sig, -1, dtype=np.int64)
    n_persist = 0
    for s in range(n_sig):
        if persistent[s]:
            sig_slot[s] = n_persist
            n_persist += 1
    # last level at which each recyclable signal is read
    last_read = [-1] * n_sig
    for i in range(n_total):
        lv = level[i]
        for s in (A[i], B[i], C[i]):
            if s >= 0 and not persistent[s] and lv > last_read[s]:
                last_read[s] = lv
    free: List[int] = []
    release: Dict[int, List[int]] = {}
    n_scratch = 0
    slot_of = sig_slot  # alias
    order_list = order.tolist()
    pos = 0
    for lv in range(n_levels):
        rel = release.pop(lv, None)
        if rel:
            free.extend(rel)
        end = int(level_off[lv + 1])
        while pos < end:
            i = order_list[pos]
            pos += 1
            o = OUT[i]
            if persistent[o]:
                continue
            if free:
                slot = free.pop()
            else:
                slot = n_scratch
                n_scratch += 1
            slot_of[o] = n_persist + slot
            when = max(last_read[o], lv) + 1
            lst = release.get(when)
            if lst is None:
                release[when] = [slot]
            else:
                lst.append(slot)
    if n_persist + n_scratch >= (1 << 31):
        raise CircuitError("too many slots")
    if allow3 and n_persist + n_scratch >= IN2_LIMIT:
        # in2 would not fit beside the opcode: redo with 2-input micro-ops only
        return compile_design(design, keep=keep, observe=observe, allow3=False, back_edges=back_edges,
                              share_edge_detectors=share_edge_detectors,
                              sort_level_by_fanin=sort_level_by_fanin, adder_prefix=adder_prefix)

    # ---- optional: within a level, order records by the slots of their fan-in -------------------
    # Records of a level are independent, so their order is free.  Sorting by (in1 slot, in0
    # slot) makes the CTAs that run side by side read neighbouring rows (DRAM page locality)
    # and puts gates that share a fan-in row next to each other (the second read hits L2).
    if sort_level_by_fanin and n_total:
        b_all = np.asarray(B, dtype=np.int64)
        a_all = np.asarray(A, dtype=np.int64)
        sb_all = np.where(b_all >= 0, sig_slot[np.maximum(b_all, 0)], -1)
        sa_all = np.where(a_all >= 0, sig_slot[np.maximum(a_all, 0)], -1)
        order = np.lexsort((sa_all, sb_all, level_arr))

    # ---- encode ----------------------------------------------------------------------------
    code_a = np.asarray(code, dtype=np.int64)[order]
    a_a = np.asarray(A, dtype=np.int64)[order]
    b_a = np.asarray(B, dtype=np.int64)[order]
    c_a = np.asarray(C, dtype=np.int64)[order]
    o_a = np.asarray(OUT, dtype=np.int64)[order]

    def slots(x):
        r = np.zeros(len(x), dtype=np.int64)
        m = x >= 0
        r[m] = sig_slot[x[m]]
        return r

    sa, sb, sc, so = slots(a_a), slots(b_a), slots(c_a), slots(o_a)
    if n_total and (min(sa.min(), sb.min(), sc.min(), so.min()) < 0):
        raise AssertionError("operand without a slot")
    records = np.empty((n_total, 4), dtype=np.uint32)
    records[:, 0] = (code_a | (sc << 8)).astype(np.uint32)
    records[:, 1] = sa.astype(np.uint32)
    records[:, 2] = sb.astype(np.uint32)
    records[:, 3] = so.astype(np.uint32)

    # operand hints for the persistent level kernel (bit 31 of in0 / in1): "stream" when the
    # value was produced two or more levels back (or never, i.e. an input): it is not in L2
    # any more and should not push the previous level's rows out on its way through.
    wr = np.asarray(writer, dtype=np.int64)
    lvl_ord = level_arr[order]
    seq_ord = order

    def stream_hint(x):
        h = np.zeros(len(x), dtype=bool)
        m = x >= 0
        w = np.full(len(x), -1, dtype=np.int64)
        w[m] = wr[x[m]]
        never = m & (w < 0)
        fwd = m & (w >= 0) & (w < seq_ord)
        h[never] = True
        wl = np.zeros(len(x), dtype=np.int64)
        wl[fwd] = level_arr[w[fwd]]
        h[fwd] = (lvl_ord[fwd] - wl[fwd]) >= 2
        return h

    records_hinted = records.copy()
    records_hinted[:, 1] |= (stream_hint(a_a).astype(np.uint32) << np.uint32(31))
    records_hinted[:, 2] |= (stream_hint(b_a).astype(np.uint32) << np.uint32(31))

    is_gate = np.asarray(ops.is_gate, dtype=bool)
    n_gates = int(is_gate.sum())
    arity = np.zeros(256, dtype=np.int64)
    for op, ar in OP_ARITY.items():
        arity[op] = ar
        arity[op | OP_NEG] = ar
    # algorithmic bytes per lane-word: 8 B per fan-in read + 8 B for the write
    gw_bytes = int(((arity[np.asarray(code, dtype=np.int64)] + 1) * 8).sum()) if n_total else 0

    constants = []
    for leaf, desc in enumerate(design.leaf_desc):
        if _type_name(desc) == "Constant":
            pin = design.leaf_pin_base[leaf] + design.leaf_table[leaf].index["out"]
            base = int(ops.net_sig[design.pin_net[pin]])
            constants.append((base, design.pin_width[pin], int(desc.value)))

    return Program(
        net_sig=ops.net_sig, shared_prev=set(ops.shared_prev),
        design=design, records=records, level_off=level_off.astype(np.int32),
        n_levels=n_levels, has_latch=has_latch, n_persist=n_persist, n_scratch=n_scratch,
        sig_slot=sig_slot, sig_shadow={s: int(sig_slot[sh]) for s, sh in shadow_sig.items()},
        sig_kept=kept, n_gates=n_gates, n_ops=n_total, gate_words_bytes=gw_bytes,
        constants=constants, max_level_gates=int(counts.max()) if n_total else 0,
        records_hinted=records_hinted)


def schedule_batches(records: np.ndarray, level_off: np.ndarray, trash_slot: int, batch: int = 8):
    """Arrange a level-sorted record stream into batches of ``batch`` records for the resident
    kernel: every batch holds records of ONE base opcode from ONE level (records of a level
    are mutually independent, so the kernel may issue all loads of a batch before its first
    store), padded with records that read and write ``trash_slot`` - a private row nobody
    else uses - so the kernel needs no per-record opcode or NOP test.  The complement flag
    (bit 7) may differ inside a batch.  Returns (records_in_batches, n_batches)."""
    n = len(records)
    if n == 0:
        return np.zeros((0, 4), dtype=np.uint32), 0
    base = (records[:, 0] & 0x7F).astype(np.int64)
    n_levels = len(level_off) - 1
    level = np.repeat(np.arange(n_levels, dtype=np.int64), np.diff(level_off).astype(np.int64))
    live = base != OP_NOP
    idx = np.nonzero(live)[0]
    key = level[idx] * 128 + base[idx]
    order = idx[np.argsort(key, kind="stable")]           # by level, then opcode, then sequence
    key_sorted = level[order] * 128 + base[order]
    # group boundaries
    starts = np.nonzero(np.concatenate(([True], key_sorted[1:] != key_sorted[:-1])))[0]
    sizes = np.diff(np.concatenate((starts, [len(order)])))
    padded = (sizes + batch - 1) // batch * batch
    out_start = np.concatenate(([0], np.cumsum(padded)))[:-1]
    total = int(padded.sum())
    out = np.empty((total, 4), dtype=np.uint32)
    # padding: same opcode as the group, all operands on the trash slot
    grp_op = base[order[starts]].astype(np.uint32)
    pad_rec = np.empty((len(starts), 4), dtype=np.uint32)
    pad_rec[:, 0] = grp_op | (np.uint32(trash_slot) << np.uint32(8))
    pad_rec[:, 1] = trash_slot
    pad_rec[:, 2] = trash_slot
    pad_rec[:, 3] = trash_slot
    out[:] = np.repeat(pad_rec, padded, axis=0)
    grp_of = np.repeat(np.arange(len(starts)), sizes)
    pos_in_grp = np.arange(len(order)) - starts[grp_of]
    out[out_start[grp_of] + pos_in_grp] = records[order]
    return out, total // batch


def group_levels(records: np.ndarray, level_off: np.ndarray, trash_slot: int, batch: int = 4):
    """Same grouping as ``schedule_batches`` (per level, per base opcode, padded to a multiple
    of ``batch`` with trash-slot records) but also returns the level offsets of the grouped
    stream, for the level kernels that decode the opcode once per batch."""
    n_levels = len(level_off) - 1
    out, _ = schedule_batches(records, level_off, trash_slot, batch)
    base = (records[:, 0] & 0x7F).astype(np.int64)
    level = np.repeat(np.arange(n_levels, dtype=np.int64), np.diff(level_off).astype(np.int64))
    live = base != OP_NOP
    # padded size of every (level, opcode) group, summed per level
    key = level[live] * 128 + base[live]
    uniq, counts = np.unique(key, return_counts=True)
    padded = (counts + batch - 1) // batch * batch
    per_level = np.zeros(n_levels, dtype=np.int64)
    np.add.at(per_level, uniq // 128, padded)
    off = np.zeros(n_levels + 1, dtype=np.int64)
    np.cumsum(per_level, out=off[1:])
    assert int(off[-1]) == len(out)
    return out, off.astype(np.int32)


def _sig_name(design: Design, s: int) -> str:
    if s >= design.n_signals:
        return "<temp %d>" % s
    net = int(np.searchsorted(design.net_sig, s, side="right") - 1)
    bit = s - int(design.net_sig[net])
    for pin in range(len(design.pin_width)):
        if design.pin_net[pin] == net:
            return "%s[%d]" % (design.pin_path(pin), bit)
    return "<net %d bit %d>" % (net, bit)


def compile_circuit(root, map_pins: bool = True, keep="all", observe: Sequence[str] = (),
                    back_edges="auto", share_edge_detectors: bool = False,
                    sort_level_by_fanin: bool = False) -> Program:
    return compile_design(flatten(root, map_pins), keep=keep, observe=observe, back_edges=back_edges,
                          share_edge_detectors=share_edge_detectors, sort_level_by_fanin=sort_level_by_fanin)
