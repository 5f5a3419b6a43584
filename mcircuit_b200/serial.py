"""Serial programs (K3 v3): the reference's ``step()`` as ONE in-place instruction stream, fed by TMA.

The reference executes its primitives one after another, in emit order, over one global per
net (``/root/reference/core/simulator.py:195-223``).  Lanes never interact, so a GPU thread
that owns (a slice of) one lane column can do exactly the same thing: walk the lowered micro-ops
in the reference's own order, in place.  No levels, no shadow slots, no latch pass - feedback
reads see last step's value simply because their writer has not run yet, as in the reference.

What makes that fast on a B200 (``csrc/mcb200.cu::k_run_stream``):

  * **producer / consumer CTA** - one producer warp walks the stream ahead of four consumer warps and
    stages, per block of up to 32 records, the block's control line and every state row the block
    reads into a shared-memory ring: one ``cp.async.bulk`` for the control line and one
    ``cp.async.bulk.tensor.2d`` (tensor-map TMA) per *run* of consecutive slots, completion on
    mbarriers; consumers never wait for HBM/L2 on the dependent chain;
  * **register forwarding** - an operand that is the result of the record just before it is
    taken from the accumulator register (``ACC``), and a result all of whose readers are served
    that way (or from the stash) is never stored;
  * **stash** - a result read again within ``STASH_WINDOW`` micro-ops is also written to one of
    ``STASH_N`` per-thread shared-memory words and read back from there (``STASH``): a row the
    stream wrote a moment ago cannot come through the ring (the producer may have fetched it before
    the store), and going back to L2 for it would put a memory round trip on the chain;
  * **segments** - a block is a list of segments: generic records (decoded one by one), or a pattern
    the kernel has straight-line code for: ``MUX`` = a chain of enabled registers
    ``out = G ? d : out`` (every stage but the first takes the previous stage's NEW value - the
    reference's in-place order), ``XM`` = register-pipeline stages ``out = G ? (acc ^ q) : out``;
  * **event skipping** - ``MUX(G, d, out) -> out`` and ``out ^= G`` are the identity in every
    lane where ``G`` is 0.  A maximal run of such records on one enable ``G`` (the shared
    ``rise`` of a clock: every ``Register`` and ``Counter``, simulator.py:100-127), together with
    the unobservable temporaries only they consume, becomes a *guarded region*: a CTA whose
    lanes all have ``G == 0`` jumps over it (consumers vote, the producer follows).  A ``Clock``
    toggles every step (Q5), so half of all steps of a clocked design are skipped, exactly.

Which rows may come through the ring.  The producer fetches block ``b`` as soon as the consumers
have finished block ``b - G`` (``G <= GB_MAX`` ring groups), and a consumer's ordinary stores are
visible to TMA reads only after a ``fence.proxy.async`` - executed at the start of every
``FENCE_PERIOD``-th block and at every guard (where producer and consumers meet anyway).  So a row is
``RING`` only if a guard lies between its last store and the read, or if that store is at least
``RING_DISTANCE`` micro-ops back (32 per block x (GB_MAX + FENCE_PERIOD + 1) blocks); otherwise it is
``STASH``, ``ACC`` or ``LATE`` (an ordinary load at the point of use, always correct for a thread's
own column).

Stream: one 816-byte control line per block = 204 words (what the producer needs comes first):
  m0 = ring rows | flags << 8 (1 = ends in a GUARD segment, 2 = FENCE first) | segments << 16 | runs << 24
  m1 = blocks to jump when the guard's enable is zero everywhere;  m2 = records;  m3 = 0
  run[16] = first slot | (first ring row | log2(rows) << 5) << 24     (2^k consecutive slots = one TMA box)
  seg[24] = kind(2) | count << 2 | first record << 8 | first ring row << 13 | (stash entry + 1) << 18 | head is ACC << 24
            | q by record << 25
            (count: units of a pattern; a pattern segment ends at a unit whose result also goes to the stash)
  slot[32] = slot of ring row j
  32 records of four words:
    w0 = a field | op(4) << 24 | neg << 28 | store << 29
    w1 = b field | kind_a(3) << 24 | kind_b(3) << 27
    w2 = c field | kind_c(3) << 24
    w3 = out slot | (stash entry + 1) << 24
  field = ring row (RING), slot (LATE), stash entry (STASH).
Pattern segments use canonical rows: MUX unit u reads ring row ``first + u`` and stores to that row's slot;
XM stage u of n reads rows ``first + u`` (q) and ``first + n + u`` (old value, and the slot it stores to); an XM
segment flagged "q by record" takes each stage's q from wherever its XOR record says (stash, late load) and has
only the n old-value rows, ``first + u``.
"""
from __future__ import annotations

from bisect import bisect_left, bisect_right
from typing import Dict, List, Sequence

import numpy as np

from . import compiler
from .compiler import OP_ARITY, OP_MUX, OP_NEG, OP_XOR, CircuitError, Program, _Instance, _type_name

OP_GUARD = 11
K_NONE, K_RING, K_LATE, K_ACC, K_GREG, K_STASH = 0, 1, 2, 3, 4, 5
SEG_GENERIC, SEG_MUX, SEG_XM, SEG_GUARD = 0, 1, 2, 3
BLOCK = 32                 # records per block
RING_ROWS = 32             # rows one block can stage
MAX_SEG = 24
MAX_RUNS = 16
W_RUN, W_SEG, W_SLOT, W_REC = 4, 4 + MAX_RUNS, 4 + MAX_RUNS + MAX_SEG, 4 + MAX_RUNS + MAX_SEG + RING_ROWS
BLOCK_WORDS = W_REC + 4 * BLOCK
GB_MAX = 6                 # ring groups (blocks in flight) the kernel uses at most
FENCE_PERIOD = 2           # every block b with b % FENCE_PERIOD == 0 starts with fence.proxy.async
RING_DISTANCE = BLOCK * (GB_MAX + FENCE_PERIOD + 1)
STASH_N = 48               # per-thread shared-memory forwarding words
STASH_WINDOW = RING_DISTANCE   # micro-ops between a stashed result and its last reader, at most: beyond it the ring can stage the row
SLOT_LIMIT = 1 << 24
F_NEG, F_STORE = 1 << 4, 1 << 5
M_GUARD, M_FENCE = 1 << 8, 1 << 9
_COMMUTATIVE2 = (2, 3, 4)   # AND, OR, XOR
PAT_MUX, PAT_XM = SEG_MUX, SEG_XM
MIN_REGION = 32            # a guard costs a block boundary and a CTA-wide vote; shorter runs are not worth one


def compile_serial(design, keep="all", observe: Sequence[str] = (), share_edge_detectors: bool = False,
                   guard: bool = True, patterns: bool = True, stash: bool = True, ring: bool = True,
                   adder_prefix: bool = False) -> Program:
    """Lower ``design`` to a serial program.  ``keep`` / ``observe`` as in
    ``compiler.compile_design``; the result is a ``Program`` with ``serial=True`` whose
    ``records`` are the control lines described in the module docstring (``uint32[n_blocks, 204]``).
    ``guard`` / ``patterns`` / ``stash`` / ``ring`` switch the optimisations off one by one (tests, sweeps);
    the results are the same."""
    if keep not in ("all", "ports"):
        raise ValueError("keep must be 'all' or 'ports'")
    ops = compiler.lower(design, True, share_edge_detectors, adder_prefix)
    code, A, B, C, OUT = ops.code, ops.a, ops.b, ops.c, ops.out
    n_ops = len(code)
    n_sig = ops.n_sig
    n_pin_sig = ops.first_temp
    operands = (A, B, C)

    # ---- who writes, who reads, what is state -----------------------------------------------------
    n_writers = [0] * n_sig
    for o in OUT:
        n_writers[o] += 1
    state_read = bytearray(n_sig)            # read before its (first) write of the step
    written = bytearray(n_sig)
    last_read = [-1] * n_sig
    n_reads = [0] * n_sig
    first_write = [n_ops] * n_sig
    last_write = [-1] * n_sig
    for i in range(n_ops):
        for arr in operands:
            s = arr[i]
            if s < 0:
                continue
            if not written[s] and n_writers[s]:
                state_read[s] = 1
            last_read[s] = i
            n_reads[s] += 1
        o = OUT[i]
        written[o] = 1
        if first_write[o] == n_ops:
            first_write[o] = i
        last_write[o] = i

    persistent = bytearray(n_sig)
    kept = np.zeros(n_pin_sig, dtype=bool)
    for s in range(n_pin_sig):
        if s in ops.dead:
            continue
        if keep == "all" or n_writers[s] == 0 or state_read[s]:
            persistent[s] = 1
            kept[s] = True
    if keep == "ports":
        want = []
        for name, child in design.top.children.items():
            if not isinstance(child, _Instance) and _type_name(design.leaf_desc[child]) == "ExposedPin":
                want.append(design.leaf_pin_base[child])
        for path in observe:
            want.append(design.find_pin(path))
        for pin in want:
            base = int(ops.net_sig[design.pin_net[pin]])
            for s in range(base, base + design.pin_width[pin]):
                persistent[s] = 1
                kept[s] = True

    # ---- guarded regions ---------------------------------------------------------------------------------
    region_of = [-1] * n_ops                 # op -> region index
    regions: List[List[int]] = []            # [begin, end, guard signal]
    if guard and n_ops:
        def identity_when_zero(i, g):
            op = code[i] & 0x7F
            if code[i] & OP_NEG:
                return False
            if op == OP_MUX:
                return A[i] == g and C[i] == OUT[i] and B[i] != g
            if op == OP_XOR:
                return (A[i] == OUT[i] and B[i] == g) or (B[i] == OUT[i] and A[i] == g)
            return False

        count: Dict[int, int] = {}
        for i in range(n_ops):
            op = code[i] & 0x7F
            if op == OP_MUX and C[i] == OUT[i]:
                count[A[i]] = count.get(A[i], 0) + 1
            elif op == OP_XOR:
                for g, self_ in ((A[i], B[i]), (B[i], A[i])):
                    if self_ == OUT[i] and g != OUT[i]:
                        count[g] = count.get(g, 0) + 1
        for g, cnt in sorted(count.items(), key=lambda kv: -kv[1]):
            if cnt < MIN_REGION or n_writers[g] > 1:
                continue
            i = (first_write[g] + 1) if n_writers[g] else 0
            while i < n_ops:
                if region_of[i] != -1 or not identity_when_zero(i, g):
                    i += 1
                    continue
                j = i
                while j < n_ops and region_of[j] == -1 and OUT[j] != g and (
                        identity_when_zero(j, g) or
                        (not persistent[OUT[j]] and n_writers[OUT[j]] == 1 and g not in (A[j], B[j], C[j]))):
                    j += 1
                # temporaries written inside must die inside; cut the region in front of the first that does not
                end = j
                changed = True
                while changed:
                    changed = False
                    for k in range(i, end):
                        if not identity_when_zero(k, g) and last_read[OUT[k]] >= end:
                            end = k
                            changed = True
                            break
                # drop trailing producers nobody inside consumes any more
                while end > i and not identity_when_zero(end - 1, g):
                    end -= 1
                if end - i >= MIN_REGION:
                    idx = len(regions)
                    regions.append([i, end, g])
                    for k in range(i, end):
                        region_of[k] = idx
                    i = end
                else:
                    i = max(end, i + 1)


    region_start = {r[0]: k for k, r in enumerate(regions)}
    region_end = {r[1]: k for k, r in enumerate(regions)}
    after_region = set(region_end)           # ops right behind a region: the accumulator may be stale there

    # ---- who serves each read: ACC, GREG, STASH or memory ----------------------------------------------------
    writer_now = [-1] * n_sig
    served = [[None, None, None] for _ in range(n_ops)]       # K_ACC / K_GREG / ("stash", writer) / None = memory
    want_stash: Dict[int, int] = {}          # writer op -> last op that wants its result from the stash
    guard_reader: Dict[int, int] = {}        # region -> writer op whose result the GUARD record wants from the stash
    for i in range(n_ops):
        if i in region_start:                # the GUARD record reads the enable just before op i
            k = region_start[i]
            w = writer_now[regions[k][2]]
            if stash and w >= 0 and region_of[w] == -1 and i - w <= STASH_WINDOW:
                guard_reader[k] = w
                want_stash[w] = max(want_stash.get(w, -1), i)
        arity = OP_ARITY[code[i] & 0x7F]
        reg = region_of[i]
        for k, s in enumerate((A[i], B[i], C[i])[:arity]):
            w = writer_now[s]
            if reg >= 0 and s == regions[reg][2]:
                served[i][k] = K_GREG
            elif i > 0 and s == OUT[i - 1] and i not in after_region:
                served[i][k] = K_ACC
            elif (stash and w >= 0 and i - w <= STASH_WINDOW and (region_of[w] == -1 or region_of[w] == reg)):
                served[i][k] = ("stash", w)
                want_stash[w] = max(want_stash.get(w, -1), i)
        writer_now[OUT[i]] = i
    # entries: a writer keeps its entry until its last stash reader; when all are busy its readers go to memory
    stash_of: Dict[int, int] = {}
    free_entries = list(range(STASH_N - 1, -1, -1))
    expiry: Dict[int, List[int]] = {}
    for i in range(n_ops):
        for e in expiry.pop(i, ()):
            free_entries.append(e)
        if i in want_stash and free_entries:
            e = free_entries.pop()
            stash_of[i] = e
            expiry.setdefault(want_stash[i] + 1, []).append(e)
    for i in range(n_ops):
        for k in range(3):
            sv = served[i][k]
            if isinstance(sv, tuple) and sv[1] not in stash_of:
                served[i][k] = None
    for k in list(guard_reader):
        if guard_reader[k] not in stash_of:
            del guard_reader[k]

    # ---- store elision: a transient result all of whose readers take it from ACC or the stash ---------------
    mem_reads = [0] * n_sig                  # reads of a signal that go to memory
    for i in range(n_ops):
        arity = OP_ARITY[code[i] & 0x7F]
        for k, s in enumerate((A[i], B[i], C[i])[:arity]):
            if served[i][k] is None:
                mem_reads[s] += 1
    for k, (b, e, g) in enumerate(regions):
        if k not in guard_reader:
            mem_reads[g] += 1
    elide = [False] * n_ops
    for i in range(n_ops):
        o = OUT[i]
        if not persistent[o] and n_writers[o] == 1 and mem_reads[o] == 0:
            elide[i] = True

    # ---- slots ------------------------------------------------------------------------------------------
    sig_slot = np.full(n_sig, -1, dtype=np.int64)
    n_persist = 0
    for s in range(n_sig):
        if persistent[s]:
            sig_slot[s] = n_persist
            n_persist += 1
    free: List[int] = []
    release: Dict[int, List[int]] = {}
    n_scratch = 0
    for i in range(n_ops):
        rel = release.pop(i, None)
        if rel:
            free.extend(rel)
        o = OUT[i]
        if persistent[o] or elide[i] or sig_slot[o] >= 0:
            continue
        slot = free.pop() if free else n_scratch
        if slot == n_scratch:
            n_scratch += 1
        sig_slot[o] = n_persist + slot
        when = max(last_read[o], last_write[o], i) + 1
        release.setdefault(when, []).append(slot)
    if n_persist + n_scratch + 1 >= SLOT_LIMIT:
        raise CircuitError("too many slots (%d) for a serial program (24-bit slot fields)" % (n_persist + n_scratch))

    # ---- which memory reads may come through the ring (decided in op space: see the module docstring) -------
    store_ops: Dict[int, List[int]] = {}     # slot -> ops that store it
    for i in range(n_ops):
        if not elide[i]:
            store_ops.setdefault(int(sig_slot[OUT[i]]), []).append(i)
    guard_ops = sorted(region_start)         # a guard executes right in front of each of these ops

    def ring_ok(slot, i, for_guard=False):
        """May the read of ``slot`` by op i (for_guard: by the guard in front of op i) be staged ahead?"""
        if not ring:
            return False
        lst = store_ops.get(slot)
        if not lst:
            return True                      # never stored by the stream: inputs, constants
        j = bisect_left(lst, i)
        last = lst[j - 1] if j > 0 else lst[-1] - n_ops
        if i - last >= RING_DISTANCE:
            return True
        if guard_ops:                        # a guard (a fence and a rendezvous with the producer) in between
            k = (bisect_left if for_guard else bisect_right)(guard_ops, i)
            r = guard_ops[k - 1] if k > 0 else guard_ops[-1] - n_ops
            if last < r:
                return True
        return False

    # ---- pattern units ------------------------------------------------------------------------------------
    #   'M'  MUX(G, d, out) -> out, the enabled register (its old value comes through the ring)
    #   'X'  XOR(acc, row) whose only reader is the 'M' right behind it ('m'): a register pipeline stage
    cls = [""] * n_ops
    if patterns:
        for i in range(n_ops):
            k = region_of[i]
            if (k >= 0 and code[i] == OP_MUX and served[i][0] == K_GREG and C[i] == OUT[i] and served[i][2] is None
                    and served[i][1] != K_GREG and not elide[i] and ring_ok(int(sig_slot[OUT[i]]), i)):
                cls[i] = "M"
        for i in range(n_ops - 1):
            if (code[i] == OP_XOR and cls[i + 1] == "M" and region_of[i + 1] == region_of[i] and elide[i]
                    and i not in stash_of and served[i + 1][1] == K_ACC and n_reads[OUT[i]] == 1):
                sa, sb = served[i][0], served[i][1]
                if (sa == K_ACC) != (sb == K_ACC) and K_GREG not in (sa, sb):
                    cls[i] = "X"                     # the other operand: a ring row, a stash word or a late load
                    cls[i + 1] = "m"

    # ---- blocks: segments of records, ring rows, TMA runs ------------------------------------------------------
    n_acc = n_elided = n_late = n_ring = n_stash = n_greg = 0
    n_seg_kind = {SEG_GENERIC: 0, SEG_MUX: 0, SEG_XM: 0, SEG_GUARD: 0}
    n_units = {SEG_MUX: 0, SEG_XM: 0}

    def pieces(rows):
        """Power-of-two boxes of consecutive slots (one plane) that cover ``rows`` in order: [(first row, log2)]."""
        out, j = [], 0
        while j < len(rows):
            k = j + 1
            while k < len(rows) and rows[k] == rows[k - 1] + 1 and (rows[k] < n_persist) == (rows[j] < n_persist):
                k += 1
            n = k - j
            while n:
                lg = n.bit_length() - 1
                out.append((j, lg))
                j += 1 << lg
                n -= 1 << lg
        return out

    class Blk:
        def __init__(self):
            self.recs: List[List[int]] = []          # four words each
            self.segs: List[List[int]] = []          # [kind, count, first record, first row]
            self.rows: List[int] = []
            self.row_of: Dict[int, int] = {}         # generic rows, shared between records
            self.stored: set = set()
            self.guard = -1                          # region whose guard ends the block
            self.first_op = -1
            self.sealed = False                      # the last pattern segment ended in a stash write: do not extend it

        def fits(self, n_recs, new_rows, new_seg):
            return self.fits_rows(n_recs, self.rows + new_rows, new_seg)

        def fits_rows(self, n_recs, rows, new_seg):
            if len(self.recs) + n_recs > BLOCK or len(rows) > RING_ROWS:
                return False
            if new_seg and len(self.segs) >= MAX_SEG:
                return False
            return len(pieces(rows)) <= MAX_RUNS

    blocks: List[Blk] = []
    cur = Blk()
    guard_block: Dict[int, int] = {}
    region_first_block: Dict[int, int] = {}
    block_of_op = [0] * n_ops

    def close():
        nonlocal cur
        if cur.recs:
            blocks.append(cur)
            cur = Blk()

    def operand_plan(i, for_guard_slot=None):
        """[(kind, field-or-slot, is_memory)] for op i's operands given the current block (nothing is added yet)."""
        plan = []
        if for_guard_slot is not None:
            srcs = [for_guard_slot]
        else:
            arity = OP_ARITY[code[i] & 0x7F]
            srcs = (A[i], B[i], C[i])[:arity]
        for k, s in enumerate(srcs):
            sv = None if for_guard_slot is not None else served[i][k]
            if sv == K_GREG:
                plan.append([K_GREG, 0])
            elif sv == K_ACC:
                plan.append([K_ACC, 0])
            elif isinstance(sv, tuple):
                plan.append([K_STASH, stash_of[sv[1]]])
            else:
                slot = int(s) if for_guard_slot is not None else int(sig_slot[s])
                if slot < 0:
                    raise AssertionError("operand without a slot at op %d" % i)
                if slot not in cur.stored and ring_ok(slot, i, for_guard_slot is not None):
                    plan.append([K_RING, slot])
                else:
                    plan.append([K_LATE, slot])
        return plan

    def new_generic_rows(plan):
        rows = []
        for kind, f in plan:
            if kind == K_RING and f not in cur.row_of and f not in rows:
                rows.append(f)
        return rows

    def add_record(i, plan, seg_kind, new_seg):
        """Encode op i with the operand plan (RING entries still carry slots: resolved to rows here)."""
        nonlocal n_acc, n_elided, n_late, n_ring, n_stash, n_greg
        kinds, fields = [K_NONE] * 3, [0] * 3
        for k, (kind, f) in enumerate(plan):
            kinds[k] = kind
            if kind == K_RING:
                if seg_kind == SEG_GENERIC or f not in pattern_row:
                    if f not in cur.row_of:
                        cur.row_of[f] = len(cur.rows)
                        cur.rows.append(f)
                    fields[k] = cur.row_of[f]
                else:
                    fields[k] = pattern_row[f]
                n_ring += 1
            else:
                fields[k] = f
                n_late += kind == K_LATE
                n_acc += kind == K_ACC
                n_stash += kind == K_STASH
                n_greg += kind == K_GREG
        op = code[i]
        base = op & 0x7F
        if seg_kind == SEG_GENERIC and base in _COMMUTATIVE2 and kinds[1] == K_ACC and kinds[0] != K_ACC:
            kinds[0], kinds[1] = kinds[1], kinds[0]          # the accumulator operand goes first
            fields[0], fields[1] = fields[1], fields[0]
        flags = base | (F_NEG if op & OP_NEG else 0)
        out_slot = 0
        if not elide[i]:
            out_slot = int(sig_slot[OUT[i]])
            flags |= F_STORE
            cur.stored.add(out_slot)
        else:
            n_elided += 1
        stash_out = stash_of[i] + 1 if i in stash_of else 0
        if cur.first_op < 0:
            cur.first_op = i
        block_of_op[i] = len(blocks)
        cur.recs.append([fields[0] | (flags << 24), fields[1] | ((kinds[0] | (kinds[1] << 3)) << 24),
                         fields[2] | (kinds[2] << 24), out_slot | (stash_out << 24)])

    pattern_row: Dict[int, int] = {}
    i = 0
    while i < n_ops:
        if i in region_end:
            close()                                          # a jump lands on a block boundary
        if i in region_start:
            k = region_start[i]
            g_sig = regions[k][2]
            for attempt in (0, 1):
                if k in guard_reader:
                    plan = [[K_STASH, stash_of[guard_reader[k]]]]
                else:
                    plan = operand_plan(i, for_guard_slot=int(sig_slot[g_sig]))
                rows = new_generic_rows(plan)
                if cur.fits(1, rows, True):
                    break
                close()
            kind, f = plan[0]
            if kind == K_RING:
                if f not in cur.row_of:
                    cur.row_of[f] = len(cur.rows)
                    cur.rows.append(f)
                f = cur.row_of[f]
                n_ring += 1
            elif kind == K_LATE:
                n_late += 1
            else:
                n_stash += 1
            cur.segs.append([SEG_GUARD, 1, len(cur.recs), 0, 0, False])
            cur.recs.append([f | (OP_GUARD << 24), kind << 24, 0, 0])
            cur.guard = k
            n_seg_kind[SEG_GUARD] += 1
            if cur.first_op < 0:
                cur.first_op = i
            guard_block[k] = len(blocks)
            close()
            region_first_block[k] = len(blocks)
        c = cls[i]
        if c == "M":
            # one unit: 1 record, its old row; a head may bring a row of its own for d
            for attempt in (0, 1):
                plan = operand_plan(i)
                last_seg = cur.segs[-1] if cur.segs else None
                extend = (last_seg is not None and last_seg[0] == SEG_MUX and plan[1][0] == K_ACC and not cur.sealed and
                          last_seg[2] + last_seg[1] == len(cur.recs) and last_seg[3] + last_seg[1] == len(cur.rows))
                old_slot = int(sig_slot[OUT[i]])
                head_rows = [] if extend else new_generic_rows([plan[1]])
                if cur.fits(1, head_rows + [old_slot], not extend):
                    break
                close()
            pattern_row = {}
            if not extend:
                for f in head_rows:                          # the head's own data row goes in front of the segment
                    cur.row_of[f] = len(cur.rows)
                    cur.rows.append(f)
                cur.segs.append([SEG_MUX, 0, len(cur.recs), len(cur.rows), 0, False])
                cur.sealed = False
                n_seg_kind[SEG_MUX] += 1
            seg = cur.segs[-1]
            pattern_row = {old_slot: len(cur.rows)}
            cur.rows.append(old_slot)
            plan[2] = [K_RING, old_slot]
            if plan[1][0] == K_RING and plan[1][1] == old_slot:
                plan[1] = [K_LATE, old_slot]                 # MUX(G, out, out): degenerate, keep it simple
            add_record(i, [plan[0], plan[1], plan[2]], SEG_MUX, False)
            seg[1] += 1
            if i in stash_of:                                # the segment's last value goes to the stash: it ends here
                seg[4] = stash_of[i] + 1
                cur.sealed = True
            n_units[SEG_MUX] += 1
            i += 1
            continue
        if c == "X":
            for attempt in (0, 1):
                px = operand_plan(i)
                last_seg = cur.segs[-1] if cur.segs else None
                q_idx = 1 if px[0][0] == K_ACC else 0
                q_kind, q_field = px[q_idx]
                old_slot = int(sig_slot[OUT[i + 1]])
                ok_kinds = px[1 - q_idx][0] == K_ACC and q_kind in (K_RING, K_STASH, K_LATE) and old_slot not in cur.stored
                if not ok_kinds:
                    break
                by_record = q_kind != K_RING               # the stage's q does not come through the ring
                if by_record:
                    extend = (last_seg is not None and last_seg[0] == SEG_XM and last_seg[5] and not cur.sealed and
                              last_seg[2] + 2 * last_seg[1] == len(cur.recs) and last_seg[3] + last_seg[1] == len(cur.rows))
                    tentative = cur.rows + [old_slot]
                else:
                    extend = (last_seg is not None and last_seg[0] == SEG_XM and not last_seg[5] and not cur.sealed and
                              last_seg[2] + 2 * last_seg[1] == len(cur.recs) and last_seg[3] + 2 * last_seg[1] == len(cur.rows))
                    # canonical rows of a stage segment: all q rows, then all old rows (the olds are consecutive slots)
                    if extend:
                        at = last_seg[3] + last_seg[1]
                        tentative = cur.rows[:at] + [q_field] + cur.rows[at:] + [old_slot]
                    else:
                        tentative = cur.rows + [q_field, old_slot]
                if cur.fits_rows(2, tentative, not extend):
                    break
                close()
            if ok_kinds:
                if not extend:
                    cur.segs.append([SEG_XM, 0, len(cur.recs), len(cur.rows), 0, by_record])
                    cur.sealed = False
                    n_seg_kind[SEG_XM] += 1
                seg = cur.segs[-1]
                if by_record:
                    cur.rows.append(old_slot)
                    n_ring += 1
                    n_stash += q_kind == K_STASH
                    n_late += q_kind == K_LATE
                else:
                    cur.rows.insert(seg[3] + seg[1], q_field)
                    cur.rows.append(old_slot)
                    n_ring += 2
                    q_field = 0
                pattern_row = {}
                # records: XOR(acc, q); MUX(G, acc, old) -> out; ring rows are the segment's by convention
                block_of_op[i] = block_of_op[i + 1] = len(blocks)
                if cur.first_op < 0:
                    cur.first_op = i
                cur.recs.append([0 | (OP_XOR << 24), q_field | ((K_ACC | (q_kind << 3)) << 24), 0, 0])
                out_slot = old_slot
                stash_out = stash_of[i + 1] + 1 if (i + 1) in stash_of else 0
                cur.recs.append([0 | ((OP_MUX | F_STORE) << 24), 0 | ((K_GREG | (K_ACC << 3)) << 24),
                                 0 | (K_RING << 24), out_slot | (stash_out << 24)])
                cur.stored.add(out_slot)
                n_acc += 2
                n_greg += 1
                n_elided += 1
                seg[1] += 1
                if (i + 1) in stash_of:
                    seg[4] = stash_of[i + 1] + 1
                    cur.sealed = True
                n_units[SEG_XM] += 1
                i += 2
                continue
            cls[i] = cls[i + 1] = ""                          # fall through: generic records
            c = ""
        # generic record
        for attempt in (0, 1):
            plan = operand_plan(i)
            rows = new_generic_rows(plan)
            last_seg = cur.segs[-1] if cur.segs else None
            extend = last_seg is not None and last_seg[0] == SEG_GENERIC and last_seg[2] + last_seg[1] == len(cur.recs)
            if cur.fits(1, rows, not extend):
                break
            close()
        if not extend:
            cur.segs.append([SEG_GENERIC, 0, len(cur.recs), 0, 0, False])
            n_seg_kind[SEG_GENERIC] += 1
        pattern_row = {}
        add_record(i, plan, SEG_GENERIC, False)
        cur.segs[-1][1] += 1
        i += 1
    close()
    n_blocks = len(blocks)

    # ---- encode ----------------------------------------------------------------------------------------------
    stream = np.zeros((n_blocks, BLOCK_WORDS), dtype=np.uint32)
    max_rows = max_runs = 0
    for b, blk in enumerate(blocks):
        runs = pieces(blk.rows)
        assert len(blk.rows) <= RING_ROWS and len(runs) <= MAX_RUNS and len(blk.segs) <= MAX_SEG and len(blk.recs) <= BLOCK
        flags = (M_FENCE if b % FENCE_PERIOD == 0 else 0)
        if blk.guard >= 0:
            flags |= M_GUARD
            e = regions[blk.guard][1]
            last_block = block_of_op[e - 1]
            stream[b, 1] = last_block - b                    # blocks b+1 .. last_block are the region
        stream[b, 0] = len(blk.rows) | flags | (len(blk.segs) << 16) | (len(runs) << 24)
        stream[b, 2] = len(blk.recs)
        for k, (kind, count, first_rec, first_row, stash_out, by_record) in enumerate(blk.segs):
            head_acc = kind == SEG_MUX and ((blk.recs[first_rec][1] >> 27) & 7) == K_ACC
            stream[b, W_SEG + k] = (kind | (count << 2) | (first_rec << 8) | (first_row << 13) | (stash_out << 18) |
                                    (int(head_acc) << 24) | (int(bool(by_record)) << 25))
        for k, (first_row, lg) in enumerate(runs):
            stream[b, W_RUN + k] = blk.rows[first_row] | ((first_row | (lg << 5)) << 24)
        stream[b, W_SLOT:W_SLOT + len(blk.rows)] = blk.rows
        for u, rec in enumerate(blk.recs):
            stream[b, W_REC + 4 * u: W_REC + 4 * u + 4] = rec
        max_rows = max(max_rows, len(blk.rows))
        max_runs = max(max_runs, len(runs))

    is_gate = np.asarray(ops.is_gate, dtype=bool)
    arity_tab = np.zeros(256, dtype=np.int64)
    for op, ar in OP_ARITY.items():
        arity_tab[op] = ar
        arity_tab[op | OP_NEG] = ar
    gw_bytes = int(((arity_tab[np.asarray(code, dtype=np.int64)] + 1) * 8).sum()) if n_ops else 0

    constants = []
    for leaf, desc in enumerate(design.leaf_desc):
        if _type_name(desc) == "Constant":
            pin = design.leaf_pin_base[leaf] + design.leaf_table[leaf].index["out"]
            base = int(ops.net_sig[design.pin_net[pin]])
            constants.append((base, design.pin_width[pin], int(desc.value)))

    prog = Program(
        net_sig=ops.net_sig, shared_prev=set(ops.shared_prev), design=design, records=stream, level_off=np.array([0, n_blocks], dtype=np.int32),
        n_levels=1 if n_blocks else 0, has_latch=False, n_persist=n_persist, n_scratch=n_scratch,
        sig_slot=sig_slot, sig_shadow={}, sig_kept=kept, n_gates=int(is_gate.sum()), n_ops=n_ops,
        gate_words_bytes=gw_bytes, constants=constants, max_level_gates=n_ops, records_hinted=None, serial=True)
    guarded = sum(e - b for b, e, _ in regions)
    n_recs_total = sum(len(blk.recs) for blk in blocks)
    prog.serial_info = {"blocks": n_blocks, "ops": n_ops, "records": n_recs_total, "acc_operands": n_acc,
                        "stash_operands": n_stash, "ring_operands": n_ring, "late_operands": n_late,
                        "guard_register_operands": n_greg, "stores_elided": n_elided,
                        "stash_entries_used": len(set(stash_of.values())), "regions": len(regions), "guarded_ops": guarded,
                        "ring_rows_staged": int(sum(len(blk.rows) for blk in blocks)),
                        "tma_boxes": int(sum(len(pieces(blk.rows)) for blk in blocks)),
                        "max_ring_rows": max_rows, "max_runs": max_runs,
                        "segments": {"generic": n_seg_kind[SEG_GENERIC], "mux_chain": n_seg_kind[SEG_MUX],
                                     "xor_mux": n_seg_kind[SEG_XM], "guard": n_seg_kind[SEG_GUARD]},
                        "mux_chain_units": n_units[SEG_MUX], "xor_mux_units": n_units[SEG_XM]}
    return prog


def compile_circuit_serial(root, map_pins: bool = True, **kw) -> Program:
    return compile_serial(compiler.flatten(root, map_pins), **kw)


# ----------------------------------------------------------------------------------------------------
# decoding (tests, the CPU test double, debugging)
# ----------------------------------------------------------------------------------------------------
def decode_block(line):
    """One control line -> dict (see the module docstring)."""
    w = [int(x) for x in np.asarray(line, dtype=np.uint32)]
    n_rows, n_seg, n_runs, n_recs = w[0] & 0xFF, (w[0] >> 16) & 0xFF, w[0] >> 24, w[2]
    recs = []
    for u in range(n_recs):
        w0, w1, w2, w3 = w[W_REC + 4 * u: W_REC + 4 * u + 4]
        flags = w0 >> 24
        recs.append({"op": flags & 15, "neg": bool(flags & F_NEG), "store": bool(flags & F_STORE),
                     "fields": (w0 & 0xFFFFFF, w1 & 0xFFFFFF, w2 & 0xFFFFFF), "out": w3 & 0xFFFFFF,
                     "stash_out": (w3 >> 24) - 1, "kinds": ((w1 >> 24) & 7, (w1 >> 27) & 7, (w2 >> 24) & 7)})
    segs = [{"kind": x & 3, "count": (x >> 2) & 63, "first_rec": (x >> 8) & 31, "first_row": (x >> 13) & 31,
             "stash_out": ((x >> 18) & 63) - 1, "q_by_record": bool((x >> 25) & 1)} for x in w[W_SEG:W_SEG + n_seg]]
    runs = [{"slot": x & 0xFFFFFF, "first_row": (x >> 24) & 31, "rows": 1 << (x >> 29)} for x in w[W_RUN:W_RUN + n_runs]]
    return {"n_rows": n_rows, "guard": bool(w[0] & M_GUARD), "fence": bool(w[0] & M_FENCE), "span": w[1],
            "rows": w[W_SLOT:W_SLOT + n_rows], "segments": segs, "runs": runs, "records": recs}


def decode(stream: np.ndarray):
    for b in range(len(stream)):
        d = decode_block(stream[b])
        d["block"] = b
        yield d
