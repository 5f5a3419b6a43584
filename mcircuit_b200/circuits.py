"""Synthetic circuits of the shapes named in BASELINE.json ``configs``.

Every builder takes the descriptor module ``M`` to build with, so the *same* construction
code can produce a circuit for this engine (``mcircuit_b200.core.descriptors``, default) and
for the unmodified reference (``/root/reference/core/descriptors.py``) when results are
pinned against its JIT.  Construction order (add_child / connect order) is part of a
circuit's meaning - the reference's evaluation order is a DFS over it - so the
builders are deterministic given their arguments.
"""
from __future__ import annotations

import random
from typing import List, Optional

from .core import descriptors as _own

TOPOLOGY_SEED = 20261019


# ---- C1: examples/counter.py:5-16 ----------------------------------------------------------
def counter_circuit(M=_own, width=16):
    s = M.Composite()
    s.add_child("clk", M.Clock())
    s.add_child("r", M.Counter(width))
    s.connect("clk", "out", "r", "clock")
    return s


# ---- examples/clock.py:4-9 -----------------------------------------------------------------
def inverter_clock(M=_own):
    s = M.Composite()
    s.add_child("not", M.Not())
    s.connect("not", "out", "not", "in")
    return s


# ---- the 5-gate full adder of examples/raw_counter.py:16-40 -----------------------------------
def full_adder(M=_own, width=1):
    ein, eout = M.ExposedPin(M.ExposedPin.IN, width), M.ExposedPin(M.ExposedPin.OUT, width)
    xor_, and_, or_ = M.Gate(M.Gate.XOR, width), M.Gate(M.Gate.AND, width), M.Gate(M.Gate.OR, width)
    fa = M.Composite()
    for name in ("a", "b", "cin"):
        fa.add_child(name, ein)
    for name in ("s", "cout"):
        fa.add_child(name, eout)
    fa.add_child("xor1", xor_)
    fa.add_child("xor2", xor_)
    fa.add_child("and1", and_)
    fa.add_child("and2", and_)
    fa.add_child("or1", or_)
    fa.connect("a", "", "xor1", "in0")
    fa.connect("b", "", "xor1", "in1")
    fa.connect("xor1", "out", "xor2", "in0")
    fa.connect("cin", "", "xor2", "in1")
    fa.connect("xor2", "out", "s", "")
    fa.connect("a", "", "and2", "in0")
    fa.connect("b", "", "and2", "in1")
    fa.connect("xor1", "out", "and1", "in0")
    fa.connect("cin", "", "and1", "in1")
    fa.connect("and1", "out", "or1", "in0")
    fa.connect("and2", "out", "or1", "in1")
    fa.connect("or1", "out", "cout", "")
    return fa


def half_adder(M=_own, width=1):
    ein, eout = M.ExposedPin(M.ExposedPin.IN, width), M.ExposedPin(M.ExposedPin.OUT, width)
    ha = M.Composite()
    ha.add_child("a", ein)
    ha.add_child("b", ein)
    ha.add_child("s", eout)
    ha.add_child("cout", eout)
    ha.add_child("xor1", M.Gate(M.Gate.XOR, width))
    ha.add_child("and1", M.Gate(M.Gate.AND, width))
    ha.connect("a", "", "xor1", "in0")
    ha.connect("b", "", "xor1", "in1")
    ha.connect("a", "", "and1", "in0")
    ha.connect("b", "", "and1", "in1")
    ha.connect("xor1", "out", "s", "")
    ha.connect("and1", "out", "cout", "")
    return ha


# ---- examples/raw_counter.py:42-121 -------------------------------------------------------------
def raw_counter(M=_own, bits=2, seed: Optional[int] = 0):
    """Gate-level ``bits``-bit counter: NOT-loop clock, master-slave flip-flops from SR/D
    latches, full adders; wired in an order shuffled by ``seed`` like the example."""
    rng = random.Random(seed)
    ein, eout = M.ExposedPin(M.ExposedPin.IN), M.ExposedPin(M.ExposedPin.OUT)
    nor = M.Gate(M.Gate.OR, negated=True)
    and_ = M.Gate(M.Gate.AND)
    not_ = M.Not()
    adder = full_adder(M)

    clock = M.Composite()
    clock.add_child("not", not_)
    clock.add_child("out", eout)
    clock.connect("not", "out", "not", "in")
    clock.connect("not", "out", "out", "")

    srlatch = M.Composite()
    srlatch.add_child("nor1", nor)
    srlatch.add_child("nor2", nor)
    srlatch.add_child("s", ein)
    srlatch.add_child("r", ein)
    srlatch.add_child("q", eout)
    srlatch.connect("nor2", "out", "nor1", "in1")
    srlatch.connect("nor1", "out", "nor2", "in0")
    srlatch.connect("s", "", "nor2", "in1")
    srlatch.connect("r", "", "nor1", "in0")
    srlatch.connect("nor1", "out", "q", "")

    dlatch = M.Composite()
    dlatch.add_child("sr", srlatch)
    dlatch.add_child("and1", and_)
    dlatch.add_child("and2", and_)
    dlatch.add_child("not", not_)
    dlatch.add_child("d", ein)
    dlatch.add_child("clk", ein)
    dlatch.add_child("q", eout)
    dlatch.connect("d", "", "not", "in")
    dlatch.connect("not", "out", "and1", "in0")
    dlatch.connect("clk", "", "and1", "in1")
    dlatch.connect("clk", "", "and2", "in0")
    dlatch.connect("d", "", "and2", "in1")
    dlatch.connect("and1", "out", "sr", "r")
    dlatch.connect("and2", "out", "sr", "s")
    dlatch.connect("sr", "q", "q", "")

    latch = M.Composite()
    latch.add_child("clk", ein)
    latch.add_child("d", ein)
    latch.add_child("q", eout)
    latch.add_child("not", not_)
    latch.add_child("dl1", dlatch)
    latch.add_child("dl2", dlatch)
    latch.connect("clk", "", "dl2", "clk")
    latch.connect("clk", "", "not", "in")
    latch.connect("not", "out", "dl1", "clk")
    latch.connect("dl1", "q", "dl2", "d")
    latch.connect("d", "", "dl1", "d")
    latch.connect("dl2", "q", "q", "")

    main = M.Composite()
    main.add_child("clk", clock)
    for i in range(1, bits + 1):
        main.add_child(f"b{i}", latch)
        main.add_child(f"a{i}", adder)
    order = list(range(1, bits + 1))
    rng.shuffle(order)
    for i in order:
        main.connect("clk", "out", f"b{i}", "clk")
        main.connect(f"b{i}", "q", f"a{i}", "a")
    rng.shuffle(order)
    for i in order:
        if i < bits:
            main.connect(f"a{i}", "cout", f"a{i + 1}", "cin")
        main.connect(f"a{i}", "s", f"b{i}", "d")
    return main


# ---- C4, all-reference variant: LFSR + pipeline out of gate-level flip-flops -------------------
def gate_level_flipflop(M=_own):
    """The NOT-loop clock and the master-slave D flip-flop of examples/raw_counter.py:42-92
    (SR latch -> D latch -> two D latches on opposite clock phases), as composites."""
    ein, eout = M.ExposedPin(M.ExposedPin.IN), M.ExposedPin(M.ExposedPin.OUT)
    nor, and_, not_ = M.Gate(M.Gate.OR, negated=True), M.Gate(M.Gate.AND), M.Not()
    clock = M.Composite()
    clock.add_child("not", not_)
    clock.add_child("out", eout)
    clock.connect("not", "out", "not", "in")
    clock.connect("not", "out", "out", "")
    sr = M.Composite()
    for name, d in (("nor1", nor), ("nor2", nor), ("s", ein), ("r", ein), ("q", eout)):
        sr.add_child(name, d)
    sr.connect("nor2", "out", "nor1", "in1")
    sr.connect("nor1", "out", "nor2", "in0")
    sr.connect("s", "", "nor2", "in1")
    sr.connect("r", "", "nor1", "in0")
    sr.connect("nor1", "out", "q", "")
    dl = M.Composite()
    for name, d in (("sr", sr), ("and1", and_), ("and2", and_), ("not", not_), ("d", ein), ("clk", ein), ("q", eout)):
        dl.add_child(name, d)
    dl.connect("d", "", "not", "in")
    dl.connect("not", "out", "and1", "in0")
    dl.connect("clk", "", "and1", "in1")
    dl.connect("clk", "", "and2", "in0")
    dl.connect("d", "", "and2", "in1")
    dl.connect("and1", "out", "sr", "r")
    dl.connect("and2", "out", "sr", "s")
    dl.connect("sr", "q", "q", "")
    ff = M.Composite()
    for name, d in (("clk", ein), ("d", ein), ("q", eout), ("not", not_), ("dl1", dl), ("dl2", dl)):
        ff.add_child(name, d)
    ff.connect("clk", "", "dl2", "clk")
    ff.connect("clk", "", "not", "in")
    ff.connect("not", "out", "dl1", "clk")
    ff.connect("dl1", "q", "dl2", "d")
    ff.connect("d", "", "dl1", "d")
    ff.connect("dl2", "q", "q", "")
    return clock, ff


def gate_level_lfsr(M=_own, n_bits=8, taps=(7, 5, 4, 3), stages=4):
    """C4's shape with storage the UNMODIFIED reference can run (its Register emitter does not
    compile): a Fibonacci LFSR of ``n_bits`` gate-level flip-flops (XNOR feedback, so the
    all-zero reset state is not a fixed point) and a ``stages``-deep flip-flop pipeline fed by
    XOR(previous stage, lfsr q0).  Ports: /q{i}/pin, /p{k}/pin."""
    clock, ff = gate_level_flipflop(M)
    eout = M.ExposedPin(M.ExposedPin.OUT)
    top = M.Composite()
    top.add_child("clk", clock)
    for i in range(n_bits):
        top.add_child(f"f{i}", ff)
        top.add_child(f"q{i}", eout)
    top.add_child("fb", M.Gate(M.Gate.XOR, 1, len(taps), True))
    for k in range(stages):
        top.add_child(f"s{k}", ff)
        top.add_child(f"x{k}", M.Gate(M.Gate.XOR))
        top.add_child(f"p{k}", eout)
    for i in range(n_bits):
        top.connect("clk", "out", f"f{i}", "clk")
        top.connect(f"f{i - 1}" if i else "fb", "q" if i else "out", f"f{i}", "d")
        top.connect(f"f{i}", "q", f"q{i}", "")
    for j, t in enumerate(taps):
        top.connect(f"f{t}", "q", "fb", f"in{j}")
    for k in range(stages):
        top.connect("clk", "out", f"s{k}", "clk")
        top.connect(f"s{k - 1}" if k else f"f{n_bits - 1}", "q", f"x{k}", "in0")
        top.connect("f0", "q", f"x{k}", "in1")
        top.connect(f"x{k}", "out", f"s{k}", "d")
        top.connect(f"s{k}", "q", f"p{k}", "")
    return top


# ---- C2: n-bit ripple-carry adder of 5-gate full adders -------------------------------------------
def ripple_adder(M=_own, nbits=32, width=1):
    """Ports: /a{i}/pin, /b{i}/pin, /cin/pin -> /s{i}/pin, /cout/pin.  ``width=64`` is the
    reference's own 64-lanes-per-gate trick (every pin carries 64 packed lanes)."""
    ein, eout = M.ExposedPin(M.ExposedPin.IN, width), M.ExposedPin(M.ExposedPin.OUT, width)
    fa = full_adder(M, width)
    top = M.Composite()
    top.add_child("cin", ein)
    for i in range(nbits):
        top.add_child(f"a{i}", ein)
        top.add_child(f"b{i}", ein)
    for i in range(nbits):
        top.add_child(f"s{i}", eout)
    top.add_child("cout", eout)
    for i in range(nbits):
        top.add_child(f"fa{i}", fa)
    for i in range(nbits):
        top.connect(f"a{i}", "", f"fa{i}", "a")
        top.connect(f"b{i}", "", f"fa{i}", "b")
        if i == 0:
            top.connect("cin", "", "fa0", "cin")
        else:
            top.connect(f"fa{i - 1}", "cout", f"fa{i}", "cin")
        top.connect(f"fa{i}", "s", f"s{i}", "")
    top.connect(f"fa{nbits - 1}", "cout", "cout", "")
    return top


# ---- C3: n x n array multiplier -----------------------------------------------------------------------
def array_multiplier(M=_own, n=64, width=1):
    """pp[i][j] = AND(a[i], b[j]); row j is added into the running sum with full adders (half
    adders where a term is absent).  Ports: /a{i}/pin, /b{j}/pin -> /p{k}/pin, k < 2n."""
    ein, eout = M.ExposedPin(M.ExposedPin.IN, width), M.ExposedPin(M.ExposedPin.OUT, width)
    and_ = M.Gate(M.Gate.AND, width)
    fa, ha = full_adder(M, width), half_adder(M, width)
    top = M.Composite()
    for i in range(n):
        top.add_child(f"a{i}", ein)
    for j in range(n):
        top.add_child(f"b{j}", ein)
    for k in range(2 * n):
        top.add_child(f"p{k}", eout)
    for j in range(n):
        for i in range(n):
            top.add_child(f"pp{j}_{i}", and_)
            top.connect(f"a{i}", "", f"pp{j}_{i}", "in0")
            top.connect(f"b{j}", "", f"pp{j}_{i}", "in1")
    # running sum: acc[k] = (child, pin) that currently holds bit k (k >= j at row j)
    acc = {i: (f"pp0_{i}", "out") for i in range(n)}
    top.connect("pp0_0", "out", "p0", "")
    for j in range(1, n):
        carry = None
        for i in range(n):
            k = i + j
            terms = [(f"pp{j}_{i}", "out")]
            if k in acc:
                terms.append(acc[k])
            if carry is not None:
                terms.append(carry)
            name = f"add{j}_{i}"
            if len(terms) == 3:
                top.add_child(name, fa)
                for pin, (c, p) in zip(("a", "b", "cin"), terms):
                    top.connect(c, p, name, pin)
                acc[k], carry = (name, "s"), (name, "cout")
            elif len(terms) == 2:
                top.add_child(name, ha)
                for pin, (c, p) in zip(("a", "b"), terms):
                    top.connect(c, p, name, pin)
                acc[k], carry = (name, "s"), (name, "cout")
            else:
                acc[k], carry = terms[0], None
        if carry is not None:
            acc[n + j] = carry
        c, p = acc[j]
        top.connect(c, p, f"p{j}", "")
    for k in range(n, 2 * n):
        if k in acc:
            c, p = acc[k]
            top.connect(c, p, f"p{k}", "")
    return top


# ---- C4: LFSRs + register pipeline -----------------------------------------------------------------------
def lfsr_pipeline(M=_own, n_lfsr=256, lfsr_bits=32, stages=1000, taps=(31, 21, 0), couple=True):
    """``n_lfsr`` feedback shift registers of ``lfsr_bits`` ``Register(1)`` each (feedback = XOR of the tap
    outputs into stage 0), one ``Clock``; a pipeline of ``stages`` ``Register(1)``: stage k data =
    XOR(stage k-1 out, row[k % n_lfsr].q0).

    The reference evaluates primitives one after another in place, sources first, and a ``Register`` is
    not double-buffered (simulator.py:115-127), so a directly chained register row FALLS THROUGH on the
    rising edge: q1..q31 all take q0's old value, and a feedback of an even number of taps is then 0 -
    an uncoupled row with four taps is all-zero after three steps, which would make every long-run
    comparison a comparison of zeros.  ``couple=True`` (default) therefore adds one more feedback input,
    q0 of the NEXT row (the last row takes row 0's), and the default has three own taps: row n evolves as
    q0 <- q0 ^ q0[next row], a 256-cell linear automaton per lane that never settles (checked for 2 x 10^5
    steps by the C4 tests).  Same micro-op count as the uncoupled four-tap row.
    Seeds go in through ``set_pin_state('/l{n}_q{b}/out', ...)``; observation ports /tap{n}/pin (q0 of each
    row) and /pipe_out/pin."""
    top = M.Composite()
    clk = M.Clock()
    reg = M.Register(1)
    eout = M.ExposedPin(M.ExposedPin.OUT)
    top.add_child("clk", clk)
    taps = [t for t in taps if t < lfsr_bits]
    n_fb = max(len(taps) + (1 if couple and n_lfsr > 1 else 0), 1)
    for n in range(n_lfsr):
        for b in range(lfsr_bits):
            top.add_child(f"l{n}_q{b}", reg)
        top.add_child(f"l{n}_fb", M.Gate(M.Gate.XOR, 1, n_fb))
        top.add_child(f"tap{n}", eout)
    for k in range(stages):
        top.add_child(f"p{k}", reg)
        top.add_child(f"px{k}", M.Gate(M.Gate.XOR, 1, 2))
    top.add_child("pipe_out", eout)
    for n in range(n_lfsr):
        for b in range(lfsr_bits):
            top.connect("clk", "out", f"l{n}_q{b}", "clock")
            if b > 0:
                top.connect(f"l{n}_q{b - 1}", "out", f"l{n}_q{b}", "data")
        for i, t in enumerate(taps):
            top.connect(f"l{n}_q{t}", "out", f"l{n}_fb", f"in{i}")
        if couple and n_lfsr > 1:
            top.connect(f"l{(n + 1) % n_lfsr}_q0", "out", f"l{n}_fb", f"in{len(taps)}")
        top.connect(f"l{n}_fb", "out", f"l{n}_q0", "data")
        top.connect(f"l{n}_q0", "out", f"tap{n}", "")
    for k in range(stages):
        top.connect("clk", "out", f"p{k}", "clock")
        src = f"p{k - 1}" if k > 0 else f"l{(n_lfsr - 1) % n_lfsr}_q{lfsr_bits - 1}"
        top.connect(src, "out", f"px{k}", "in0")
        top.connect(f"l{k % n_lfsr}_q0", "out", f"px{k}", "in1")
        top.connect(f"px{k}", "out", f"p{k}", "data")
    if stages:
        top.connect(f"p{stages - 1}", "out", "pipe_out", "")
    return top


# ---- C5: random levelized DAG ---------------------------------------------------------------------------------
def random_dag(M=_own, n_gates=1_000_000, depth=200, n_inputs=4096, seed=TOPOLOGY_SEED, width=1,
               in1_window: Optional[int] = None):
    """``depth`` levels x ``n_gates // depth`` two-input gates.  op uniform in {AND, OR,
    XOR}, negated Bernoulli(1/2); in0 uniform from the previous level, in1 uniform from any
    earlier level or the primary inputs (long-lived values).  Outputs = the last level.
    Ports: /i{k}/pin -> /o{k}/pin.  ``in1_window=w`` (measurement only) draws in1 from the last
    ``w`` levels instead, which keeps the live set small enough to stay in L2."""
    rng = random.Random(seed)
    per = n_gates // depth
    ein, eout = M.ExposedPin(M.ExposedPin.IN, width), M.ExposedPin(M.ExposedPin.OUT, width)
    kinds = [M.Gate(op, width, 2, neg) for op in (M.Gate.AND, M.Gate.OR, M.Gate.XOR) for neg in (False, True)]
    top = M.Composite()
    add_child, connect = top.add_child, top.connect
    names: List[str] = []
    for k in range(n_inputs):
        add_child(f"i{k}", ein)
        names.append(f"i{k}")
    prev_lo, prev_hi = 0, n_inputs
    for lv in range(depth):
        lo = len(names)
        for k in range(per):
            g = f"g{lv}_{k}"
            add_child(g, kinds[rng.randrange(6)])
            names.append(g)
        for k in range(per):
            g = names[lo + k]
            s0 = names[rng.randrange(prev_lo, prev_hi)]
            s1 = names[rng.randrange(0 if in1_window is None else max(0, lo - in1_window * per), lo)]
            connect(s0, "out" if s0[0] == "g" else "", g, "in0")
            connect(s1, "out" if s1[0] == "g" else "", g, "in1")
        prev_lo, prev_hi = lo, lo + per
    for k in range(per):
        add_child(f"o{k}", eout)
        connect(names[prev_lo + k], "out", f"o{k}", "")
    return top


# ---- fuzzing: small random circuits with feedback, hierarchy and every primitive -------------------------------
def random_circuit(M=_own, seed=0, n_nodes=24, max_width=4, feedback=True, hierarchy=True,
                   primitives=("Gate", "Not", "Constant", "Clock", "Counter", "Adder"), depth=0):
    """A random, always-valid circuit: every input pin gets exactly one driver of the right
    width (or stays undriven), outputs fan out freely, edges may point backwards when
    ``feedback``.  Returns (composite, probe pin paths)."""
    rng = random.Random(seed)
    top = M.Composite()
    outs = {}     # width -> list of (child, pin)   sources available for connecting
    ins = []      # (child, pin, width)             sinks still undriven
    probes = []
    widths = list(range(1, max_width + 1)) + [1, 1]

    def offer(child, pin, w, probe):
        outs.setdefault(w, []).append((child, pin))
        probes.append(probe)

    for i in range(n_nodes):
        name = f"n{i}"
        kind = rng.choice(primitives + (("Sub",) if hierarchy and depth < 2 else ()))
        w = rng.choice(widths)
        if kind == "Gate":
            ni = rng.choice((1, 2, 2, 2, 3, 4, 5))
            top.add_child(name, M.Gate(rng.choice((0, 1, 2)), w, ni, rng.random() < 0.5))
            for k in range(ni):
                ins.append((name, f"in{k}", w))
            offer(name, "out", w, f"/{name}/out")
        elif kind == "Not":
            top.add_child(name, M.Not(w))
            ins.append((name, "in", w))
            offer(name, "out", w, f"/{name}/out")
        elif kind == "Constant":
            top.add_child(name, M.Constant(w, rng.getrandbits(w)))
            offer(name, "out", w, f"/{name}/out")
        elif kind == "Clock":
            top.add_child(name, M.Clock())
            offer(name, "out", 1, f"/{name}/out")
        elif kind == "Counter":
            top.add_child(name, M.Counter(w))
            ins.append((name, "clock", 1))
            offer(name, "out", w, f"/{name}/out")
            probes.append(f"/{name}/prevclock")
        elif kind == "Register":
            top.add_child(name, M.Register(w))
            ins.append((name, "clock", 1))
            ins.append((name, "data", w))
            offer(name, "out", w, f"/{name}/out")
            probes.append(f"/{name}/prevclock")
        elif kind == "Adder":
            top.add_child(name, M.Adder(w))
            ins.append((name, "cin", 1))
            ins.append((name, "a", w))
            ins.append((name, "b", w))
            offer(name, "sum", w, f"/{name}/sum")
            offer(name, "cout", 1, f"/{name}/cout")
        elif kind == "Sub":
            sub, sub_probes = random_circuit(M, rng.getrandbits(30), max(3, n_nodes // 3), max_width,
                                             feedback, hierarchy, primitives, depth + 1)
            # give the sub-circuit ports
            n_in, n_out = rng.randrange(0, 3), rng.randrange(1, 3)
            sub_nodes = list(sub.graph.nodes)
            for k in range(n_in):
                pw = rng.choice(widths)
                sub.add_child(f"pi{k}", M.ExposedPin(M.ExposedPin.IN, pw))
                # drive an undriven inner input of that width if there is one
                for c in sub_nodes:
                    d = sub.get_child(c)
                    if hasattr(d, "graph") or type(d).__name__ == "ExposedPin":
                        continue
                    hit = [p for p, iw in d.all_inputs() if iw == pw and not any(
                        cc[2] == c and cc[3] == p for cc in sub.connections)]
                    if hit:
                        sub.connect(f"pi{k}", "", c, hit[0])
                        break
                ins.append((name, f"pi{k}", pw))
            for k in range(n_out):
                cands = [(c, p, ow) for c in sub_nodes for p, ow in sub.get_child(c).all_outputs()
                         if not hasattr(sub.get_child(c), "graph") and type(sub.get_child(c)).__name__ != "ExposedPin"]
                if not cands:
                    break
                c, p, ow = rng.choice(cands)
                sub.add_child(f"po{k}", M.ExposedPin(M.ExposedPin.OUT, ow))
                sub.connect(c, p, f"po{k}", "")
                outs.setdefault(ow, []).append((name, f"po{k}"))
                probes.append(f"/{name}/po{k}/pin")
            top.add_child(name, sub)
            probes.extend(f"/{name}{p}" for p in sub_probes)
    # top-level ports
    for k in range(2):
        w = rng.choice(widths)
        top.add_child(f"in{k}", M.ExposedPin(M.ExposedPin.IN, w))
        offer(f"in{k}", "", w, f"/in{k}/pin")
    order = {n: i for i, n in enumerate(top.graph.nodes)}
    rng.shuffle(ins)
    for child, pin, w in ins:
        cands = outs.get(w, [])
        if not feedback:
            cands = [c for c in cands if order[c[0]] < order[child]]
        if not cands or rng.random() < 0.1:
            continue                            # leave undriven
        sc, sp = rng.choice(cands)
        top.connect(sc, sp, child, pin)
    k = 0
    for w, lst in sorted(outs.items()):
        if rng.random() < 0.7:
            sc, sp = rng.choice(lst)
            top.add_child(f"out{k}", M.ExposedPin(M.ExposedPin.OUT, w))
            top.connect(sc, sp, f"out{k}", "")
            probes.append(f"/out{k}/pin")
            k += 1
    return top, probes


# ---- GUI-shaped circuits: what the reference GUI holds before Schematic.reconstruct (diagram.py:88-119) ----
# Elements are LIBRARY entries (app.py:575-589: Input/Output pins and the seven gates, named
# ``<kind>_<n>`` by element_factory, app.py:561-571); nets are the groups of pins joined by wires.
def gui_full_adder(M=_own):
    """1-bit full adder drawn with library gates, one unused NAND left on the sheet."""
    ein, eout = M.ExposedPin(M.ExposedPin.IN), M.ExposedPin(M.ExposedPin.OUT)
    elements = [("input_1", ein), ("input_2", ein), ("input_3", ein), ("xor_4", M.Gate(M.Gate.XOR)),
                ("xor_5", M.Gate(M.Gate.XOR)), ("and_6", M.Gate(M.Gate.AND)), ("and_7", M.Gate(M.Gate.AND)),
                ("or_8", M.Gate(M.Gate.OR)), ("output_9", eout), ("output_10", eout),
                ("nand_11", M.Gate(M.Gate.AND, negated=True))]
    nets = [
        [("input_1", "pin"), ("xor_4", "in0"), ("and_7", "in0")],
        [("input_2", "pin"), ("xor_4", "in1"), ("and_7", "in1")],
        [("input_3", "pin"), ("xor_5", "in1"), ("and_6", "in1")],
        [("xor_4", "out"), ("xor_5", "in0"), ("and_6", "in0")],
        [("xor_5", "out"), ("output_9", "pin")],
        [("and_6", "out"), ("or_8", "in0")],
        [("and_7", "out"), ("or_8", "in1")],
        [("or_8", "out"), ("output_10", "pin")],
    ]
    return elements, nets


def gui_latch_and_ring(M=_own):
    """Feedback drawn in the GUI: a cross-coupled NOR latch (set / reset inputs), a three-inverter
    ring oscillator, and an XNOR of the two on an output - every value depends on the order in
    which ``reconstruct`` adds children and edges."""
    ein, eout = M.ExposedPin(M.ExposedPin.IN), M.ExposedPin(M.ExposedPin.OUT)
    elements = [("input_1", ein), ("input_2", ein), ("nor_3", M.Gate(M.Gate.OR, negated=True)),
                ("nor_4", M.Gate(M.Gate.OR, negated=True)), ("not_5", M.Not()), ("not_6", M.Not()), ("not_7", M.Not()),
                ("xnor_8", M.Gate(M.Gate.XOR, negated=True)), ("output_9", eout), ("output_10", eout), ("output_11", eout)]
    nets = [
        [("input_1", "pin"), ("nor_3", "in0")],                     # reset
        [("input_2", "pin"), ("nor_4", "in1")],                     # set
        [("nor_3", "out"), ("nor_4", "in0"), ("output_9", "pin"), ("xnor_8", "in0")],     # q
        [("nor_4", "out"), ("nor_3", "in1"), ("output_10", "pin")],                       # not q
        [("not_5", "out"), ("not_6", "in")],
        [("not_6", "out"), ("not_7", "in")],
        [("not_7", "out"), ("not_5", "in"), ("xnor_8", "in1")],
        [("xnor_8", "out"), ("output_11", "pin")],
    ]
    return elements, nets


def gui_circuit(M=_own, name="full_adder"):
    """The Composite ``Schematic.reconstruct`` builds for one of the GUI-shaped sheets above."""
    from . import schematic
    elements, nets = {"full_adder": gui_full_adder, "latch_and_ring": gui_latch_and_ring}[name](M)
    return schematic.reconstruct(elements, nets, M=M)
