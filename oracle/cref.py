"""ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/mcref.c, the plain-C
restatement of the reference's sequential, in-place, emit-order evaluation
(core/simulator.py:195-223), many lanes wide.  Used by tests/ as the fast checker at sizes
where the numpy restatement is too slow, and by bench.py as the CPU baseline ("port").

The emit list is derived with the same flattening as oracle/restate.py (which is pinned
against the reference's own generators); the per-primitive arithmetic is mcref.c's own.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import restate

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmcref.so")
_T = {"Not": 1, "Gate": 2, "Counter": 3, "Register": 4, "Clock": 5, "Adder": 6}
_PINS = {
    "Not": ("in", "out"), "Counter": ("clock", "out", "prevclock"),
    "Register": ("clock", "data", "out", "prevclock"), "Clock": ("out",),
    "Adder": ("cin", "a", "b", "cout", "sum"),
}
_lib = None


def build(force=False):
    src = os.path.join(HERE, "mcref.c")
    if force or not os.path.isfile(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], check=True,
                       capture_output=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        lib.mcref_run.restype = ctypes.c_int
        lib.mcref_run.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_void_p,
                                                          ctypes.c_int64, ctypes.c_int64,
                                                          ctypes.c_int64, ctypes.c_int]
        _lib = lib
    return _lib


class CRef:
    """Same surface as the reference's JIT (simulator.py:174-297) + lanes."""

    def __init__(self, root, burst_size, map_pins=True, lanes=64, threads=None):
        assert lanes % 64 == 0
        self.lib = _load()
        self.burst_size = int(burst_size)
        self.lanes, self.words = lanes, lanes // 64
        self.threads = threads or os.cpu_count() or 1
        pin_map = restate._sources(root) if map_pins else None
        self._pin_map = pin_map
        self._width, self._row = {}, {}
        n_rows = 0
        for path, width, _ in restate._pins(root):
            self._width[path] = width
            if map_pins and pin_map[path] is not None:
                continue
            self._row[path] = n_rows                      # one row per net bit (:185-191)
            n_rows += width
        self.n_rows = n_rows
        self.cb = 8                                       # mcref.c CB: columns per block
        self.n_blocks = (self.words + self.cb - 1) // self.cb
        # [block][row][CB]: one block's whole net file is contiguous (cache friendly)
        self.val = np.zeros((self.n_blocks, n_rows, self.cb), dtype=np.uint64)
        types, widths, auxs, pin_off, pins = [], [], [], [0], []
        constants = []
        n_gate_evals = 0
        for op, data in restate._topology(root):
            if op == "propagate":
                if map_pins:
                    continue                              # self-copy of an aliased global
                src, dst = data
                if self._width[src] != self._width[dst]:
                    raise TypeError("cannot store i%d to i%d*" % (self._width[src], self._width[dst]))
                types.append(7)
                widths.append(self._width[src])
                auxs.append(0)
                pins.extend((self._base(src), self._base(dst)))
                pin_off.append(len(pins))
                continue
            d, path = data
            t = type(d).__name__
            if t == "Constant":
                constants.append((d, path))
                continue
            if t not in _T:
                continue
            types.append(_T[t])
            widths.append(1 if t == "Clock" else d.width)
            if t == "Gate":
                if d.num_inputs < 1:
                    raise ValueError("Gate without inputs")
                auxs.append(d.num_inputs | ((d.op & 0xFF if d.op in (0, 1, 2) else 3) << 8) | (int(bool(d.negated)) << 16))
                pins.extend(self._base(path + "in%d" % i) for i in range(d.num_inputs))
                pins.append(self._base(path + "out"))
                n_gate_evals += d.width * max(1, d.num_inputs - 1)
            else:
                auxs.append(0)
                pins.extend(self._base(path + p) for p in _PINS[t])
            pin_off.append(len(pins))
        self.n_ops = len(types)
        self._type = np.asarray(types, dtype=np.int32)
        self._widths = np.asarray(widths, dtype=np.int32)
        self._aux = np.asarray(auxs, dtype=np.int32)
        self._pin_off = np.asarray(pin_off, dtype=np.int64)
        self._pins = np.asarray(pins, dtype=np.int64)
        for d, path in constants:                         # :275-276
            self.set_pin_state(path + "out", d.value)

    def _source(self, path):
        if self._pin_map is None:
            if path not in self._row:
                raise KeyError(path)
            return path
        src = self._pin_map[path]
        return path if src is None else src

    def _base(self, path):
        return self._row[self._source(path)]

    def run(self, steps, threads=None):
        rc = self.lib.mcref_run(self._type.ctypes.data, self._widths.ctypes.data, self._aux.ctypes.data,
                                self._pin_off.ctypes.data, self._pins.ctypes.data, self.n_ops,
                                self.val.ctypes.data, self.n_rows, self.n_blocks, int(steps),
                                int(threads or self.threads))
        if rc != 0:
            raise RuntimeError("mcref_run failed (%d)" % rc)

    def step(self):
        self.run(1)

    def burst(self):
        self.run(self.burst_size + 1)                     # :233-242

    def rows(self, first, count):
        """uint64[count, words] copy of net rows first..first+count."""
        v = self.val[:, first:first + count, :].transpose(1, 0, 2).reshape(count, -1)
        return np.ascontiguousarray(v[:, :self.words])

    def set_rows(self, first, planes):
        planes = np.asarray(planes, dtype=np.uint64)
        count = planes.shape[0]
        full = np.zeros((count, self.n_blocks * self.cb), dtype=np.uint64)
        full[:, :self.words] = planes.reshape(count, self.words)
        self.val[:, first:first + count, :] = full.reshape(count, self.n_blocks, self.cb).transpose(1, 0, 2)

    def planes(self, pin):
        return self.rows(self._base(pin), self._width[pin])

    def set_planes(self, pin, planes):
        self.set_rows(self._base(pin), np.asarray(planes, dtype=np.uint64).reshape(self._width[pin], self.words))

    def get_pin_state(self, pin, lanes=False):
        v = restate.unpack_lanes(self.planes(pin))
        return v if lanes else int(v[0])

    def set_pin_state(self, pin, value):
        w = self._width[pin]
        self._source(pin)
        if np.isscalar(value) or isinstance(value, int):
            value = np.full(self.lanes, int(value) & ((1 << w) - 1), dtype=np.uint64)
        self.set_planes(pin, restate.pack_lanes(np.asarray(value, dtype=np.uint64), w))
