"""Serial programs (K3 v3): the reference's ``step()`` as ONE in-place instruction stream, fed by TMA.

The reference executes its primitives one after another, in emit order, over one global per
net (``/root/reference/core/simulator.py:195-223``).  Lanes never interact, so a GPU thread
that owns (a slice of) one lane column can do exactly the same thing: walk the lowered micro-ops
in the reference's own order, in place.  No levels, no shadow slots, no latch pass - feedback
reads see last step's value simply because their writer has not run yet, as in the reference.

What makes that fast on a B200 (``csrc/mcb200.cu::k_run_stream``):

  * **producer / consumer CTA** - one producer warp walks the stream ahead of four consumer warps and
    stages, per block of 8 records, the block's control line and every state row the block reads
    into a shared-memory ring with ``cp.async.bulk`` (TMA) + mbarriers; consumers never wait for
    HBM/L2 on the dependent chain, and up to ``GB_MAX`` blocks of rows are in flight per SM;
  * **register forwarding** - an operand that is the result of the record just before it is
    taken from the accumulator register (``ACC``), and a result all of whose readers are served
    that way (or from the stash) is never stored;
  * **stash** - a result read again within ``STASH_WINDOW`` micro-ops is also written to one of
    ``STASH_N`` per-thread shared-memory words and read back from there (``STASH``): a row the
    stream wrote a moment ago cannot come through the ring (the producer may have fetched it before
    the store), and going back to L2 for it would put a memory round trip on the chain;
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
``RING`` only if its last store lies before the latest such point that precedes ``b - GB_MAX``
(or before the last guard in front of ``b``); otherwise it is ``STASH``, ``ACC`` or ``LATE`` (an
ordinary load at the point of use, always correct for a thread's own column).

Stream: one 208-byte control line per block = 52 words:
  m0 = ring rows | flags << 8 (1 GUARD in record 7, 2 FENCE first, 4 a pattern block writes the stash) | pattern << 16 | units << 20
  m1 = blocks to jump when the guard's enable is zero everywhere;  m2 = records worth decoding;  m3 = 0
  f[16] = slots of the rows to stage, ring row j = f[j]
  8 records of four words:
    w0 = a field | op(4) << 24 | neg << 28 | store << 29
    w1 = b field | kind_a(3) << 24 | kind_b(3) << 27
    w2 = c field | kind_c(3) << 24
    w3 = out slot | (stash entry + 1) << 24
  field = ring row (RING), slot (LATE), stash entry (STASH).
Patterns (straight-line code in the kernel for the same records): 1 = up to 8 enabled registers
``MUX(G, acc | x, ring[u]) -> f[u]``; 2 = up to 4 pipeline stages ``XOR(acc, ring[2u])`` +
``MUX(G, acc, ring[2u+1]) -> f[2u+1]``; their rows are in this canonical order (a slot may be staged twice).
"""
from __future__ import annotations

from bisect import bisect_left
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import compiler
from .compiler import (OP_ARITY, OP_BUF, OP_MUX, OP_NEG, OP_NOP, OP_XOR, CircuitError, Program,
                       _Instance, _type_name)

OP_GUARD = 11
K_NONE, K_RING, K_LATE, K_ACC, K_GREG, K_STASH = 0, 1, 2, 3, 4, 5
BLOCK = 8
RING_ROWS = 16             # rows one block can stage
BLOCK_WORDS = 4 + RING_ROWS + 4 * BLOCK
GB_MAX = 24                # ring groups (blocks in flight) the kernel uses at most
FENCE_PERIOD = 16          # every block b with b % FENCE_PERIOD == 0 starts with fence.proxy.async
STASH_N = 32               # per-thread shared-memory forwarding words
STASH_WINDOW = 256         # micro-ops between a stashed result and its last reader, at most
SLOT_LIMIT = 1 << 24
F_NEG, F_STORE = 1 << 4, 1 << 5
M_GUARD, M_FENCE, M_STASH = 1 << 8, 1 << 9, 1 << 10
_COMMUTATIVE2 = (2, 3, 4)   # AND, OR, XOR
PAT_MUX, PAT_XM = 1, 2
MIN_REGION = 32            # a guard costs a block boundary and a CTA-wide vote; shorter runs are not worth one


def compile_serial(design, keep="all", observe: Sequence[str] = (), share_edge_detectors: bool = False,
                   guard: bool = True, patterns: bool = True, stash: bool = True, ring: bool = True) -> Program:
    """Lower ``design`` to a serial program.  ``keep`` / ``observe`` as in
    ``compiler.compile_design``; the result is a ``Program`` with ``serial=True`` whose
    ``records`` are the control lines described in the module docstring (``uint32[n_blocks, 52]``).
    ``guard`` / ``patterns`` / ``stash`` / ``ring`` switch the optimisations off one by one (tests, sweeps);
    the results are the same."""
    if keep not in ("all", "ports"):
        raise ValueError("keep must be 'all' or 'ports'")
    ops = compiler.lower(design, True, share_edge_detectors)
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
    first_read = [n_ops] * n_sig
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
            if first_read[s] == n_ops:
                first_read[s] = i
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

    # ---- store elision: a transient result whose only readers are the very next record ------------
    def next_only(i):
        o = OUT[i]
        if persistent[o] or n_writers[o] != 1:
            return False
        if n_reads[o] == 0:
            return True                                    # dead value: computed, never stored
        return first_read[o] == i + 1 and last_read[o] == i + 1

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


    # ---- block patterns: runs the kernel has straight-line code for (see k_run_stream) ------------------
    #   'M'  MUX(G, d, out) -> out, the enabled register; d = the accumulator or a row
    #   'X'  XOR(acc, row) whose only reader is the 'M' right behind it (a register pipeline stage)
    cls = [""] * n_ops
    if patterns:
        for i in range(n_ops):
            k = region_of[i]
            if k >= 0 and code[i] == OP_MUX and A[i] == regions[k][2] and C[i] == OUT[i] and B[i] != regions[k][2]:
                cls[i] = "M"
        for i in range(1, n_ops - 1):
            k = region_of[i]
            if (k >= 0 and code[i] == OP_XOR and cls[i + 1] == "M" and region_of[i + 1] == k and B[i + 1] == OUT[i]
                    and next_only(i) and n_reads[OUT[i]] >= 1 and i != regions[k][0]
                    and ((A[i] == OUT[i - 1]) != (B[i] == OUT[i - 1])) and regions[k][2] not in (A[i], B[i])):
                cls[i] = "X"
                cls[i + 1] = "m"                     # second half of an X-M pair

    # a pattern block costs a block boundary on both sides: a lone unit is cheaper as generic records
    i = 0
    while i < n_ops:
        if cls[i] in ("M", "X"):
            kind = cls[i]
            j, units = i, 0
            while j < n_ops and cls[j] == kind and region_of[j] == region_of[i]:
                j += 2 if kind == "X" else 1
                units += 1
            if units < 2:
                for q in range(i, j):
                    cls[q] = ""
            i = j
        else:
            i += 1

    # ---- layout: real records, GUARD at the last position of a block, NOP padding ---------------------
    # entries: [kind, op index or guard signal, region index]
    REC, NOPR, GUARDR = 0, 1, 2
    layout: List[List[int]] = []

    def pad_to(mod_target):
        while len(layout) % BLOCK != mod_target:
            layout.append([NOPR, -1, -1])

    region_start = {r[0]: k for k, r in enumerate(regions)}
    region_end = {r[1]: k for k, r in enumerate(regions)}
    guard_pos: Dict[int, int] = {}
    after_region = set()                     # ops that must not take their operand from ACC
    run_kind, run_fill = "", 0               # pattern run being laid out, units in its current block
    for i in range(n_ops):
        if i in region_end:
            pad_to(0)
            after_region.add(i)
            run_kind = ""
        if i in region_start:
            k = region_start[i]
            pad_to(BLOCK - 1)
            guard_pos[k] = len(layout)
            layout.append([GUARDR, regions[k][2], k])
            run_kind = ""
        kind = {"M": "M", "X": "XM", "m": "XM"}.get(cls[i], "")
        if cls[i] != "m":
            cap = {"M": BLOCK, "XM": BLOCK // 2}.get(kind, 0)
            if kind != run_kind or (kind and run_fill >= cap):
                if kind or run_kind:
                    pad_to(0)                        # pattern blocks start on a boundary and end in NOPs
                run_kind, run_fill = kind, 0
            run_fill += 1
        layout.append([REC, i, region_of[i]])
    pad_to(0)
    n_rec = len(layout)
    n_blocks = n_rec // BLOCK
    op_pos = [0] * n_ops
    for p, (kind, i, _) in enumerate(layout):
        if kind == REC:
            op_pos[i] = p
    region_span = {}                         # blocks to jump: from the block after the guard's to the region's last
    for k, (b, e, g) in enumerate(regions):
        region_span[k] = op_pos[e - 1] // BLOCK - guard_pos[k] // BLOCK

    # ---- who serves each read: ACC, GREG, STASH or memory ----------------------------------------------------
    # reads[i] = [(operand index, signal, writer op of this step or -1)]
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
    guard_sigs = {g for _, _, g in regions}
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

    # ---- where every slot is stored (record positions), guards and fences: what the ring may carry -----------
    stores_at: Dict[int, List[int]] = {}
    for p, (kind, i, reg) in enumerate(layout):
        if kind == REC and not elide[i]:
            stores_at.setdefault(int(sig_slot[OUT[i]]), []).append(p)
    guard_blocks = sorted(pos // BLOCK for pos in guard_pos.values())

    def last_store_block(slot, p):
        """Block (possibly negative = earlier steps) of the last store to ``slot`` in front of record position p."""
        lst = stores_at.get(slot)
        if not lst:
            return None
        j = bisect_left(lst, p)
        if j > 0:
            return lst[j - 1] // BLOCK
        return lst[-1] // BLOCK - n_blocks

    def cover(b):
        """Stores of blocks < cover(b) are visible to the producer's fetch of block b (module docstring)."""
        t = b - GB_MAX
        cyc = t // n_blocks
        c = ((t - cyc * n_blocks) // FENCE_PERIOD) * FENCE_PERIOD + cyc * n_blocks
        if guard_blocks:
            j = bisect_left(guard_blocks, b)
            g = guard_blocks[j - 1] if j > 0 else guard_blocks[-1] - n_blocks
            c = max(c, g + 1)
        return c

    # ---- operand kinds, ring rows per block, encoding -----------------------------------------------------------
    stream = np.zeros((n_blocks, BLOCK_WORDS), dtype=np.uint32)
    n_acc = n_elided = n_late = n_ring = n_stash = n_greg = 0
    rows: List[int] = []
    row_of: Dict[int, int] = {}
    stored_in_block: set = set()
    for p, (kind, i, reg) in enumerate(layout):
        b, u = divmod(p, BLOCK)
        if u == 0:
            rows, row_of, stored_in_block = [], {}, set()
        base_w = 4 + RING_ROWS + 4 * u

        def memory_operand(slot):
            """(kind, field) for a read of ``slot`` at record position p."""
            nonlocal n_late, n_ring
            ok = ring and slot not in stored_in_block
            if ok:
                lsb = last_store_block(slot, p)
                ok = lsb is None or lsb < cover(b)
            if ok and slot not in row_of and len(rows) >= RING_ROWS:
                ok = False
            if not ok:
                n_late += 1
                return K_LATE, slot
            if slot not in row_of:
                row_of[slot] = len(rows)
                rows.append(slot)
            n_ring += 1
            return K_RING, row_of[slot]

        if kind == GUARDR:
            g, k = i, reg
            if k in guard_reader:
                ak, af = K_STASH, stash_of[guard_reader[k]]
                n_stash += 1
            else:
                ak, af = memory_operand(int(sig_slot[g]))
            stream[b, base_w] = af | (OP_GUARD << 24)
            stream[b, base_w + 1] = ak << 24
            stream[b, 0] |= M_GUARD
            stream[b, 1] = region_span[k]
        elif kind == REC:
            op = code[i]
            base = op & 0x7F
            arity = OP_ARITY[base]
            kinds, fields = [K_NONE] * 3, [0] * 3
            for k, s in enumerate((A[i], B[i], C[i])[:arity]):
                sv = served[i][k]
                if sv == K_GREG:
                    kinds[k] = K_GREG
                    n_greg += 1
                elif sv == K_ACC:
                    kinds[k] = K_ACC
                    n_acc += 1
                elif isinstance(sv, tuple):
                    kinds[k], fields[k] = K_STASH, stash_of[sv[1]]
                    n_stash += 1
                else:
                    slot = int(sig_slot[s])
                    if slot < 0:
                        raise AssertionError("operand without a slot at op %d" % i)
                    kinds[k], fields[k] = memory_operand(slot)
            if base in _COMMUTATIVE2 and kinds[1] == K_ACC and kinds[0] != K_ACC:
                kinds[0], kinds[1] = kinds[1], kinds[0]          # the accumulator operand goes first
                fields[0], fields[1] = fields[1], fields[0]
            flags = base | (F_NEG if op & OP_NEG else 0)
            out_slot = 0
            if not elide[i]:
                out_slot = int(sig_slot[OUT[i]])
                flags |= F_STORE
                stored_in_block.add(out_slot)
            else:
                n_elided += 1
            stash_out = stash_of[i] + 1 if i in stash_of else 0
            stream[b, base_w] = fields[0] | (flags << 24)
            stream[b, base_w + 1] = fields[1] | ((kinds[0] | (kinds[1] << 3)) << 24)
            stream[b, base_w + 2] = fields[2] | (kinds[2] << 24)
            stream[b, base_w + 3] = out_slot | (stash_out << 24)
        if u == BLOCK - 1:
            stream[b, 0] |= len(rows)
            stream[b, 4:4 + len(rows)] = rows
            if b % FENCE_PERIOD == 0:
                stream[b, 0] |= M_FENCE

    # ---- tag the blocks that match a pattern exactly (the hint only selects straight-line code for the
    # same records; a block that misses any precondition simply stays generic) ---------------------------
    n_pat = {PAT_MUX: 0, PAT_XM: 0}
    M24 = 0xFFFFFF
    for b in range(n_blocks):
        blk = [[int(x) for x in stream[b, 4 + RING_ROWS + 4 * u: 8 + RING_ROWS + 4 * u]] for u in range(BLOCK)]
        ops_ = [(w[0] >> 24) & 15 for w in blk]
        last = max([q + 1 for q in range(BLOCK) if ops_[q] != OP_NOP] or [0])
        stream[b, 2] = last
        n = 0
        while n < BLOCK and ops_[n] not in (OP_NOP, OP_GUARD):
            n += 1
        if not patterns or n == 0 or any(o != OP_NOP for o in ops_[n:]):
            continue

        def ka(w):
            return (w[1] >> 24) & 7

        def kb(w):
            return (w[1] >> 27) & 7

        def kc(w):
            return (w[2] >> 24) & 7

        def is_m(w):
            return (w[0] >> 24) == (OP_MUX | F_STORE) and ka(w) == K_GREG and kc(w) == K_RING

        def is_x(w):
            return ((w[0] >> 24) == OP_XOR and ka(w) == K_ACC and kb(w) == K_RING and kc(w) == K_NONE and
                    (w[3] >> 24) == 0)

        pat = 0
        rows_now = [int(x) for x in stream[b, 4:4 + (int(stream[b, 0]) & 0xFF)]]
        R0 = 4 + RING_ROWS
        if all(is_m(blk[q]) for q in range(n)) and all(kb(blk[q]) == K_ACC for q in range(1, n)) and kb(blk[0]) != K_NONE:
            # canonical rows: row u = old value of unit u (its slot is also the slot unit u stores to), then the
            # row the head takes its data from, if it takes it from the ring
            new_rows = [rows_now[blk[q][2] & M24] for q in range(n)]
            if all(new_rows[q] == (blk[q][3] & M24) for q in range(n)):
                for q in range(n):
                    stream[b, R0 + 4 * q + 2] = q | (K_RING << 24)
                if kb(blk[0]) == K_RING:
                    new_rows.append(rows_now[blk[0][1] & M24])
                    stream[b, R0 + 1] = n | (int(stream[b, R0 + 1]) & 0xFF000000)
                pat = PAT_MUX | (n << 4)                     # only the head of a chain may take its data from elsewhere
        elif n % 2 == 0 and all(is_x(blk[q]) and is_m(blk[q + 1]) and kb(blk[q + 1]) == K_ACC
                                for q in range(0, n, 2)):
            # canonical rows: 2u = the row XORed in, 2u + 1 = old value of the stage's register (= its store slot)
            new_rows = []
            for q in range(0, n, 2):
                new_rows += [rows_now[blk[q][1] & M24], rows_now[blk[q + 1][2] & M24]]
            if all(new_rows[q + 1] == (blk[q + 1][3] & M24) for q in range(0, n, 2)):
                for q in range(0, n, 2):
                    stream[b, R0 + 4 * q + 1] = q | (int(stream[b, R0 + 4 * q + 1]) & 0xFF000000)
                    stream[b, R0 + 4 * (q + 1) + 2] = (q + 1) | (K_RING << 24)
                pat = PAT_XM | ((n // 2) << 4)
        if pat:
            stream[b, 0] = (int(stream[b, 0]) & ~0xFF) | len(new_rows) | (pat << 16)
            stream[b, 4:4 + RING_ROWS] = 0
            stream[b, 4:4 + len(new_rows)] = new_rows
            if any(w[3] >> 24 for w in blk[:n]):
                stream[b, 0] |= M_STASH
            n_pat[pat & 15] += 1

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
        net_sig=ops.net_sig, design=design, records=stream, level_off=np.array([0, n_blocks], dtype=np.int32),
        n_levels=1 if n_blocks else 0, has_latch=False, n_persist=n_persist, n_scratch=n_scratch,
        sig_slot=sig_slot, sig_shadow={}, sig_kept=kept, n_gates=int(is_gate.sum()), n_ops=n_ops,
        gate_words_bytes=gw_bytes, constants=constants, max_level_gates=n_ops, records_hinted=None, serial=True)
    guarded = sum(e - b for b, e, _ in regions)
    prog.serial_info = {"blocks": n_blocks, "ops": n_ops, "acc_operands": n_acc, "stash_operands": n_stash,
                        "ring_operands": n_ring, "late_operands": n_late, "guard_register_operands": n_greg,
                        "stores_elided": n_elided, "stash_entries_used": len(set(stash_of.values())),
                        "regions": len(regions), "guarded_ops": guarded, "padding": n_rec - n_ops - len(regions),
                        "ring_rows_staged": int((stream[:, 0] & 0xFF).sum()) if n_blocks else 0,
                        "max_ring_rows": int((stream[:, 0] & 0xFF).max()) if n_blocks else 0,
                        "blocks_mux_chain": n_pat[PAT_MUX], "blocks_xor_mux": n_pat[PAT_XM]}
    return prog


def compile_circuit_serial(root, map_pins: bool = True, **kw) -> Program:
    return compile_serial(compiler.flatten(root, map_pins), **kw)


# ----------------------------------------------------------------------------------------------------
# decoding (tests, the CPU test double, debugging)
# ----------------------------------------------------------------------------------------------------
def decode_block(line):
    """One control line -> dict (see the module docstring)."""
    w = [int(x) for x in np.asarray(line, dtype=np.uint32)]
    recs = []
    for u in range(BLOCK):
        w0, w1, w2, w3 = w[4 + RING_ROWS + 4 * u: 8 + RING_ROWS + 4 * u]
        flags = w0 >> 24
        recs.append({"op": flags & 15, "neg": bool(flags & F_NEG), "store": bool(flags & F_STORE),
                     "fields": (w0 & 0xFFFFFF, w1 & 0xFFFFFF, w2 & 0xFFFFFF), "out": w3 & 0xFFFFFF,
                     "stash_out": (w3 >> 24) - 1, "kinds": ((w1 >> 24) & 7, (w1 >> 27) & 7, (w2 >> 24) & 7)})
    n_rows = w[0] & 0xFF
    return {"n_rows": n_rows, "guard": bool(w[0] & M_GUARD), "fence": bool(w[0] & M_FENCE), "stash": bool(w[0] & M_STASH),
            "pattern": (w[0] >> 16) & 15, "units": (w[0] >> 20) & 0xFF, "span": w[1], "n_recs": w[2],
            "rows": w[4:4 + n_rows], "records": recs}


def decode(stream: np.ndarray):
    for b in range(len(stream)):
        d = decode_block(stream[b])
        d["block"] = b
        yield d
