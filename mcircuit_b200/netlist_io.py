"""Project / netlist file format for the simulation path.

The reference sketches a JSON schema in ``serial.py:10-62`` (``type / width / num_inputs /
negated / op / children / connections / direction``) but the module does not import
(syntax error at ``serial.py:102``, ``desc.children`` does not exist) and the GUI pickles
instead.  This is a working loader/saver for that schema, extended to every descriptor of
``core/descriptors.py``; child order and connection order are preserved because the
evaluation order depends on them.

``save_program`` / ``load_program`` cache a *compiled* program (flat SoA arrays + the pin
table of the retained nets) so a large netlist does not pay the Python compile on every
run.
"""
from __future__ import annotations

import json
from typing import Dict

import numpy as np

from .core import descriptors as _own

FORMAT_VERSION = 1
_OPS = {0: "and", 1: "or", 2: "xor"}
_OPS_INV = {v: k for k, v in _OPS.items()}


def _desc_to_dict(desc, memo: Dict[int, int], defs: list):
    """Shared descriptor objects (the same Composite instanced under many names,
    examples/raw_counter.py:104-105) are stored once and referenced by index."""
    key = id(desc)
    if key in memo:
        return {"ref": memo[key]}
    t = type(desc).__name__
    d = {"type": t}
    if t == "Composite":
        children = []
        for name in desc.graph.nodes:                      # insertion order
            children.append([name, _desc_to_dict(desc.get_child(name), memo, defs)])
        d["children"] = children
        # edges carry the DFS order: (dst, src) in insertion order
        d["edges"] = [list(e) for e in desc.graph.edges]
        d["connections"] = sorted(list(c) for c in desc.connections)
    elif t == "ExposedPin":
        d.update(direction="in" if desc.direction == 0 else "out", width=desc.width)
    elif t == "Gate":
        d.update(op=_OPS.get(desc.op, desc.op), width=desc.width, num_inputs=desc.num_inputs,
                 negated=bool(desc.negated))
    elif t == "Constant":
        d.update(width=desc.width, value=int(desc.value))
    elif t in ("Not", "Register", "Counter", "Adder"):
        d.update(width=desc.width)
    elif t == "Clock":
        d.update(short=desc.short, long=desc.long)
    else:
        raise TypeError("cannot serialise %r" % (desc,))
    memo[key] = len(defs)
    defs.append(d)
    return {"ref": memo[key]}


def to_dict(root) -> dict:
    defs: list = []
    top = _desc_to_dict(root, {}, defs)
    return {"mcircuit_b200_netlist": FORMAT_VERSION, "defs": defs, "root": top["ref"]}


def from_dict(data: dict, M=_own):
    if data.get("mcircuit_b200_netlist") != FORMAT_VERSION:
        raise ValueError("not a netlist of format version %d" % FORMAT_VERSION)
    built: Dict[int, object] = {}

    def build(i):
        if i in built:
            return built[i]
        d = data["defs"][i]
        t = d["type"]
        if t == "Composite":
            c = M.Composite()
            for name, ref in d["children"]:
                c.add_child(name, build(ref["ref"]))
            # rebuild graph edges in their original order, then the translated connections
            for dst, src in d["edges"]:
                c.graph.add_edge(dst, src)
            for conn in d["connections"]:
                c.connections.add(tuple(conn))
            obj = c
        elif t == "ExposedPin":
            obj = M.ExposedPin(M.ExposedPin.IN if d["direction"] == "in" else M.ExposedPin.OUT, d["width"])
        elif t == "Gate":
            op = d["op"]
            obj = M.Gate(_OPS_INV.get(op, op), d["width"], d["num_inputs"], d["negated"])
        elif t == "Constant":
            obj = M.Constant(d["width"], d["value"])
        elif t == "Clock":
            obj = M.Clock()
            obj.short, obj.long = d.get("short", 1), d.get("long", 1)
        else:
            obj = getattr(M, t)(d["width"])
        built[i] = obj
        return obj

    return build(data["root"])


def save(root, path):
    with open(path, "w") as f:
        json.dump(to_dict(root), f, separators=(",", ":"))


def load(path, M=_own):
    with open(path) as f:
        return from_dict(json.load(f), M)


# ---- compiled-program cache ------------------------------------------------------------------
def save_program(program, path):
    """Flat binary (npz) form of a compiled Program: records, levels, slot geometry and the
    slot table of every retained pin, enough to run and to address pins by path."""
    d = program.design
    paths, bases, widths = [], [], []
    net_sig = program.net_sig if getattr(program, "net_sig", None) is not None else d.net_sig
    for pin in range(len(d.pin_width)):
        net = int(d.pin_net[pin])
        base = int(net_sig[net])
        w = d.pin_width[pin]
        if bool(program.sig_kept[base:base + w].all()):
            paths.append(d.pin_path(pin))
            bases.append(base)
            widths.append(w)
    ports = {"in": [], "out": []}
    for name, child in d.top.children.items():
        if isinstance(child, int) and type(d.leaf_desc[child]).__name__ == "ExposedPin":
            ports["in" if d.leaf_desc[child].direction == 0 else "out"].append("/%s/pin" % name)
    shadow = np.array(sorted(program.sig_shadow.items()), dtype=np.int64).reshape(-1, 2)
    hinted = program.records_hinted if program.records_hinted is not None else np.zeros((0, 4), dtype=np.uint32)
    info = getattr(program, "serial_info", None) or {}
    np.savez_compressed(
        path, version=FORMAT_VERSION, records=program.records, records_hinted=hinted,
        serial=bool(getattr(program, "serial", False)), n_ops=int(program.n_ops),
        max_ring_rows=int(info.get("max_ring_rows", 0)),
        level_off=program.level_off, n_persist=program.n_persist, n_scratch=program.n_scratch,
        sig_slot=program.sig_slot[:len(program.sig_kept)], sig_kept=program.sig_kept,
        n_gates=program.n_gates, gate_words_bytes=program.gate_words_bytes, has_latch=program.has_latch,
        constants=(np.array([[str(x) for x in c] for c in program.constants]) if program.constants
                   else np.zeros((0, 3), dtype="<U1")),
        shadow=shadow, pin_paths=np.array(paths), pin_base=np.array(bases, dtype=np.int64),
        pin_width=np.array(widths, dtype=np.int64), ports_in=np.array(ports["in"]), ports_out=np.array(ports["out"]))


class CachedProgram:
    """Duck-types ``compiler.Program`` for ``B200Engine(program=...)`` without a Design."""

    def __init__(self, z):
        from .compiler import PinNotObservable
        self._unobservable = PinNotObservable
        self.records = z["records"]
        self.records_hinted = z["records_hinted"] if len(z["records_hinted"]) else None
        # a stream program (mode='serial') must never be fed to the level kernels: the flag travels with the cache
        self.serial = bool(z["serial"]) if "serial" in z.files else False
        self.serial_info = {"max_ring_rows": int(z["max_ring_rows"])} if self.serial else None
        self.shared_prev = None
        self.level_off = z["level_off"]
        self.n_levels = len(self.level_off) - 1
        self.n_persist, self.n_scratch = int(z["n_persist"]), int(z["n_scratch"])
        self.sig_slot, self.sig_kept = z["sig_slot"], z["sig_kept"]
        self.n_gates, self.gate_words_bytes = int(z["n_gates"]), int(z["gate_words_bytes"])
        self.has_latch = bool(z["has_latch"])
        self.n_ops = int(z["n_ops"]) if "n_ops" in z.files else len(self.records)
        self.constants = [(int(a), int(b), int(c)) for a, b, c in z["constants"]]
        self.sig_shadow = {int(a): int(b) for a, b in z["shadow"]}
        self._pins = {str(p): (int(b), int(w)) for p, b, w in zip(z["pin_paths"], z["pin_base"], z["pin_width"])}
        self.ports_in = [str(p) for p in z["ports_in"]]
        self.ports_out = [str(p) for p in z["ports_out"]]
        self.design = None
        sizes = np.diff(self.level_off)
        self.max_level_gates = int(sizes.max()) if len(sizes) else 0

    @property
    def n_slots(self):
        return self.n_persist + self.n_scratch

    def pin_signals(self, path):
        return self._pins[path]                             # KeyError: unknown or not retained

    def pin_slots(self, path, need_value=True):
        base, width = self.pin_signals(path)
        return [int(s) for s in self.sig_slot[base:base + width]]


def load_program(path) -> CachedProgram:
    z = np.load(path, allow_pickle=False)
    if int(z["version"]) != FORMAT_VERSION:
        raise ValueError("program cache of another format version")
    return CachedProgram(z)
