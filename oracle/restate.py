"""ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path
(``mcircuit_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` legs may use it, and only as the checker.

A CPU restatement, in numpy, of the algorithm of the reference's simulation path
(matijakevic/mcircuit ``core/simulator.py``), generalised from one stimulus instance to
``lanes`` independent instances (one numpy element per lane, value-level: a w-bit pin is a
uint64 holding the w-bit integer, exactly what the reference's iW global holds).

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` runs this against the
reference's own LLVM JIT (through ``oracle/refshim.py``) whenever /root/reference is
present, and ``tests/golden/*.json`` (made by ``oracle/gen_golden.py`` from live JIT runs)
pin it everywhere else.  ``Register`` is the exception: the reference's emitter for it does
not compile (simulator.py:115-127), so its parity is "unpinned by the reference"; it is
checked against the JIT with the emitter patched to the evident intent (see gen_golden.py).

What follows what:
  _pins / _connections / _topology   simulator.py:16-25 / 28-38 / 41-53
  _sources                           simulator.py:56-74   (_map_to_sources)
  _EMIT[...]                         simulator.py:77-147  (_translate_*)
  RestatedJIT.__init__               simulator.py:175-221, 275-276
  get/set_pin_state, step, burst     simulator.py:278-297 (+ IR at :225-245)
"""
import numpy as np

_U64 = np.uint64
_ALL = _U64(0xFFFFFFFFFFFFFFFF)


def _is_composite(d):
    return hasattr(d, "graph") and hasattr(d, "connections")


def _tname(d):
    return type(d).__name__


def _mask(width):
    return _ALL if width >= 64 else _U64((1 << width) - 1)


# ---- simulator.py:16-25 ---------------------------------------------------------------
def _pins(node, path="/"):
    for name in node.graph.nodes:
        d = node.get_child(name)
        if _is_composite(d):
            yield from _pins(d, path + name + "/")
        else:
            here = path + name + "/"
            for p, w in d.all_inputs():
                yield here + p, w, "in"
            for p, w in d.all_outputs():
                yield here + p, w, "out"
            for p, w in d.all_internals():
                yield here + p, w, "internal"


def _by_source_child(node, cache):
    """connections of ``node`` keyed by the first path component of their source; the
    reference re-scans the whole set per child (simulator.py:35-38, :50-53) - same
    result, same relative order, linear time."""
    g = cache.get(id(node))
    if g is None:
        g = {}
        for c in node.connections:
            g.setdefault(c[0].split("/")[0], []).append(c)
        cache[id(node)] = g
    return g


# ---- simulator.py:28-38 ---------------------------------------------------------------
def _connections(node, path="/", cache=None):
    cache = {} if cache is None else cache
    groups = _by_source_child(node, cache)
    for name in node.graph.nodes:
        d = node.get_child(name)
        if _is_composite(d):
            yield from _connections(d, path + name + "/", cache)
        for c in groups.get(name, ()):
            yield path + c[0] + "/" + c[1], path + c[2] + "/" + c[3]


# ---- simulator.py:41-53 ---------------------------------------------------------------
def _topology(node, path="/", cache=None, order_cache=None):
    import networkx as nx
    cache = {} if cache is None else cache
    order_cache = {} if order_cache is None else order_cache
    groups = _by_source_child(node, cache)
    order = order_cache.get(id(node))
    if order is None:
        order = list(nx.dfs_postorder_nodes(node.graph))
        order_cache[id(node)] = order
    for name in order:
        d = node.get_child(name)
        if _is_composite(d):
            yield from _topology(d, path + name + "/", cache, order_cache)
        else:
            yield "emit", (d, path + name + "/")
        for c in groups.get(name, ()):
            yield "propagate", (path + c[0] + "/" + c[1], path + c[2] + "/" + c[3])


# ---- simulator.py:56-74 ---------------------------------------------------------------
def _sources(root):
    conns = {}
    for src, dst in _connections(root):
        conns[dst] = src
    traces = {}
    for pin, _, _ in _pins(root):
        cur = pin
        if cur not in conns:
            traces[pin] = None
            continue
        hops = 0
        while cur in conns:
            cur = conns[cur]
            hops += 1
            if hops > len(conns) + 1:
                raise ValueError("wire loop at %s" % pin)
        traces[pin] = cur
    return traces


# ---- signed helpers for sadd.with.overflow (simulator.py:142-147) ------------------------
def _sadd_overflow(x, y, width):
    """(sum mod 2^w, overflow flag) of the signed w-bit add of the unsigned-coded x, y."""
    m = (1 << width) - 1
    half = 1 << (width - 1)
    xs = [int(v) for v in x]
    ys = [int(v) for v in y]
    out = np.empty(len(xs), dtype=_U64)
    ovf = np.empty(len(xs), dtype=_U64)
    for i, (a, b) in enumerate(zip(xs, ys)):
        sa = a - (1 << width) if a >= half else a
        sb = b - (1 << width) if b >= half else b
        t = sa + sb
        ovf[i] = 1 if (t < -half or t > half - 1) else 0
        out[i] = t & m
    return out, ovf


def _sadd_overflow_fast(x, y, width):
    """Same as _sadd_overflow for width <= 62 with int64 arithmetic (no Python loop)."""
    half = np.int64(1 << (width - 1))
    full = np.int64(1 << width)
    xs = x.astype(np.int64)
    ys = y.astype(np.int64)
    xs = np.where(xs >= half, xs - full, xs)
    ys = np.where(ys >= half, ys - full, ys)
    t = xs + ys
    ovf = ((t < -half) | (t > half - 1)).astype(_U64)
    return (t & np.int64(full - 1)).astype(_U64), ovf


def _sadd(x, y, width):
    return _sadd_overflow_fast(x, y, width) if width <= 62 else _sadd_overflow(x, y, width)


class RestatedJIT:
    """Same constructor and methods as the reference's ``JIT`` (simulator.py:174-297),
    plus ``lanes``: every pin holds ``lanes`` independent values."""

    def __init__(self, root, burst_size, map_pins=True, lanes=1):
        self.root = root
        self.burst_size = int(burst_size)
        self.lanes = int(lanes)
        self._pin_map = _sources(root) if map_pins else None
        self._width = {}
        self._val = {}
        for path, width, _ in _pins(root):
            self._width[path] = width
            if map_pins and self._pin_map[path] is not None:
                continue
            self._val[path] = np.zeros(self.lanes, dtype=_U64)     # :185-191, init 0
        self._prog = []
        constants = []
        for op, data in _topology(root):
            if op == "propagate":
                src, dst = data
                s, t = self._source(src), self._source(dst)
                if self._width[src] != self._width[dst]:
                    raise TypeError("cannot store i%d to i%d*" % (self._width[src], self._width[dst]))
                self._prog.append((self._propagate, (s, t)))
            else:
                d, path = data
                fn = _EMIT.get(_tname(d))
                if fn is not None:
                    self._prog.append((fn, (self, d, path)))
                elif _tname(d) == "Constant":
                    constants.append(data)
        for d, path in constants:                                   # :275-276
            self.set_pin_state(path + "out", d.value)

    # simulator.py:278-283
    def _source(self, path):
        if self._pin_map is None:
            if path not in self._val:
                raise KeyError(path)
            return path
        src = self._pin_map[path]            # KeyError on unknown pin, like the reference
        return path if src is None else src

    def _g(self, path, pin):
        return self._val[self._source(path + pin)]

    def _set(self, path, pin, value):
        full = path + pin
        self._val[self._source(full)] = value & _mask(self._width[full])

    def _propagate(self, s, t):
        if s != t:
            self._val[t] = self._val[s].copy()

    # simulator.py:285-291 (values are masked to the pin width: documented deviation Q9)
    def get_pin_state(self, pin, lanes=False):
        v = self._val[self._source(pin)]
        return v.copy() if lanes else int(v[0])

    def set_pin_state(self, pin, value):
        src = self._source(pin)
        m = _mask(self._width[pin])
        if np.isscalar(value) or isinstance(value, int):
            self._val[src] = np.full(self.lanes, int(value) & int(m), dtype=_U64)
        else:
            v = np.asarray(value, dtype=_U64)
            if v.shape != (self.lanes,):
                raise ValueError("need %d lane values" % self.lanes)
            self._val[src] = v & m

    def step(self):
        for fn, args in self._prog:
            fn(*args)

    def burst(self):
        # the loop tests the pre-decrement counter: burst_size + 1 passes (:233-242)
        for _ in range(self.burst_size + 1):
            self.step()

    def run(self, steps):
        for _ in range(steps):
            self.step()


# ---- simulator.py:77-80 ---------------------------------------------------------------
def _emit_not(sim, d, path):
    sim._set(path, "out", ~sim._g(path, "in"))


# ---- simulator.py:83-97 ---------------------------------------------------------------
def _emit_gate(sim, d, path):
    res = None
    for i in range(d.num_inputs):
        v = sim._g(path, "in" + str(i))
        if res is None:
            res = v
        elif d.op == 0:
            res = res & v
        elif d.op == 1:
            res = res | v
        elif d.op == 2:
            res = res ^ v
    if d.negated:
        res = ~res
    sim._set(path, "out", res)


# ---- simulator.py:100-112 -------------------------------------------------------------
def _emit_counter(sim, d, path):
    v1 = sim._g(path, "prevclock")
    v2 = sim._g(path, "clock")
    v3 = (~v1 & v2) & _U64(1)                       # i1, then zext to iW: value 0 or 1
    out = sim._g(path, "out")
    m = _mask(d.width)
    nxt = (v3 & ((out + _U64(1)) & m)) | ((~v3 & m) & out)
    sim._set(path, "out", nxt)
    sim._set(path, "prevclock", sim._g(path, "clock"))


# ---- simulator.py:115-127 (intent; the reference's version does not compile) -------------
def _emit_register(sim, d, path):
    v1 = sim._g(path, "prevclock")
    v2 = sim._g(path, "clock")
    rise = (~v1 & v2) & _U64(1)
    sel = (_U64(0) - rise)                          # all-ones where rising
    nxt = (sel & sim._g(path, "data")) | (~sel & sim._g(path, "out"))
    sim._set(path, "out", nxt)
    sim._set(path, "prevclock", sim._g(path, "clock"))


# ---- simulator.py:130-132 -------------------------------------------------------------
def _emit_clock(sim, d, path):
    sim._set(path, "out", ~sim._g(path, "out"))


# ---- simulator.py:135-147 -------------------------------------------------------------
def _emit_adder(sim, d, path):
    a = sim._g(path, "a")
    b = sim._g(path, "b")
    cin = sim._g(path, "cin")                       # zext i1 -> iW keeps the value 0/1
    s1, o1 = _sadd(a, b, d.width)
    s2, o2 = _sadd(s1, cin & _U64(1), d.width)
    sim._set(path, "sum", s2)
    sim._set(path, "cout", o1 | o2)


_EMIT = {
    "Gate": _emit_gate, "Adder": _emit_adder, "Clock": _emit_clock, "Not": _emit_not,
    "Register": _emit_register, "Counter": _emit_counter,
}


# ---- bit-slicing helpers (lane l of word j  <->  lane index 64*j + l) ----------------------
def pack_lanes(values, width):
    """values: uint64[lanes] (lanes % 64 == 0) -> bit planes uint64[width, lanes // 64]."""
    values = np.asarray(values, dtype=_U64)
    lanes = values.shape[0]
    assert lanes % 64 == 0
    planes = np.zeros((width, lanes // 64), dtype=_U64)
    v = values.reshape(lanes // 64, 64)
    weights = (_U64(1) << np.arange(64, dtype=_U64))
    for bit in range(width):
        bits = (v >> _U64(bit)) & _U64(1)
        planes[bit] = (bits * weights).sum(axis=1, dtype=_U64)
    return planes


def unpack_lanes(planes):
    """bit planes uint64[width, words] -> values uint64[64 * words]."""
    planes = np.asarray(planes, dtype=_U64)
    width, words = planes.shape
    out = np.zeros(words * 64, dtype=_U64)
    shifts = np.arange(64, dtype=_U64)
    for bit in range(width):
        bits = (planes[bit][:, None] >> shifts[None, :]) & _U64(1)
        out |= bits.reshape(-1) << _U64(bit)
    return out


SEED = 0x6D63697263756974          # "mcircuit" 


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays (wraps mod 2^64)."""
    x = np.asarray(x, dtype=_U64)
    with np.errstate(over="ignore"):
        z = x + _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        return z ^ (z >> _U64(31))


def stimulus_word(stream, word_index, seed=SEED):
    """word(input stream s, lane word j) = splitmix64(seed ^ (s << 32) ^ j)  (SURVEY 8d)."""
    s = np.asarray(stream, dtype=_U64)
    j = np.asarray(word_index, dtype=_U64)
    return splitmix64(_U64(seed) ^ (s << _U64(32)) ^ j)
