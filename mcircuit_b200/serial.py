"""Serial programs (K3 v2): the reference's ``step()`` as ONE in-place instruction stream.

The reference executes its primitives one after another, in emit order, over one global per
net (``/root/reference/core/simulator.py:195-223``).  Lanes never interact, so a GPU thread
that owns (a slice of) one lane column can do exactly the same thing: walk the lowered micro-ops
in the reference's own order, in place.  No levels, no shadow slots, no latch pass - feedback
reads see last step's value simply because their writer has not run yet, as in the reference.

What makes that fast on a B200 (``csrc/mcb200.cu::k_run_serial``):

  * **register forwarding** - an operand that is the result of the record just before it is
    taken from the accumulator register (``ACC``), and a result whose only reader is the next
    record is never stored: the long dependent chains of sequential circuits (register
    pipelines, shift registers, ripple carries) run in registers instead of one memory round
    trip per stage;
  * **blocks of 8** - the memory operands of 8 consecutive records are loaded up front
    (``EARLY``) unless an earlier record of the same block stores to that slot (``LATE``), so a
    thread keeps 8-16 independent loads in flight;
  * **L2 prefetch** - a record's free slot field carries the cold row that the stream will load
    ``prefetch_distance`` records later;
  * **event skipping** - ``MUX(G, d, out) -> out`` and ``out ^= G`` are the identity in every
    lane where ``G`` is 0.  A maximal run of such records on one enable ``G`` (the shared
    ``rise`` of a clock: every ``Register`` and ``Counter``, simulator.py:100-127), together with
    the unobservable temporaries only they consume, becomes a *guarded region*: a warp whose
    lanes all have ``G == 0`` jumps over it.  A ``Clock`` toggles every step (Q5), so half of
    all steps of a clocked design are skipped, exactly.

Record: 16 bytes, four words of ``slot(24) | meta(8) << 24``:
  w0 = a slot   | op(4) | neg << 4 | store << 5 | keep_acc << 6
  w1 = b slot   | kind_a(3) | kind_b(3) << 3 | prefetch_field(2) << 6
  w2 = c slot   | kind_c(3)
  w3 = out slot   (GUARD: number of records to jump); in the FIRST record of a block the top byte is the
       block pattern: id(4) | units << 4 - 1 = up to 8 enabled registers ``MUX(G, acc | row, out) -> out``,
       2 = up to 4 pipeline stages ``XOR(acc, row)`` + ``MUX(G, acc, out) -> out``; the kernel runs such a
       block as straight-line code instead of decoding record by record
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import compiler
from .compiler import (OP_ARITY, OP_BUF, OP_MUX, OP_NEG, OP_NOP, OP_XOR, CircuitError, Program,
                       _Instance, _type_name)

OP_GUARD = 11
K_NONE, K_EARLY, K_LATE, K_ACC, K_GREG = 0, 1, 2, 3, 4
BLOCK = 8
SLOT_LIMIT = 1 << 24
F_NEG, F_STORE, F_KEEP_ACC = 1 << 4, 1 << 5, 1 << 6
_COMMUTATIVE2 = (2, 3, 4)   # AND, OR, XOR
PAT_MUX, PAT_XM = 1, 2
MIN_REGION = 32            # a guard costs a block boundary; shorter runs are not worth one
COLD_DISTANCE = 512        # records since a row was last touched before a prefetch pays


def compile_serial(design, keep="all", observe: Sequence[str] = (), share_edge_detectors: bool = False,
                   guard: bool = True, prefetch_distance: int = 64, patterns: bool = True) -> Program:
    """Lower ``design`` to a serial program.  ``keep`` / ``observe`` as in
    ``compiler.compile_design``; the result is a ``Program`` with ``serial=True`` whose
    ``records`` are the stream described in the module docstring."""
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

    # ---- layout: real records, GUARD at the last position of a block, NOP padding ---------------------
    # entries: [kind, op index or guard signal, region index]
    REC, NOPR, GUARDR = 0, 1, 2
    layout: List[List[int]] = []

    def pad_to(mod_target):
        while len(layout) % BLOCK != mod_target:
            layout.append([NOPR, -1, -1])

    # ---- block patterns: runs the kernel has straight-line code for (see k_run_serial) ------------------
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
    op_pos = [0] * n_ops
    for p, (kind, i, _) in enumerate(layout):
        if kind == REC:
            op_pos[i] = p
    region_span = {}
    for k, (b, e, g) in enumerate(regions):
        first = guard_pos[k] + 1
        last = op_pos[e - 1]
        span = (last // BLOCK + 1) * BLOCK - first
        region_span[k] = span

    # ---- slots ------------------------------------------------------------------------------------------
    # a result nobody but the next record reads stays in the accumulator (unless that record sits
    # behind a region boundary, where the accumulator is stale when the region was skipped)
    elide = [False] * n_ops
    for i in range(n_ops):
        if next_only(i) and (n_reads[OUT[i]] == 0 or (i + 1) not in after_region):
            elide[i] = True
    sig_slot = np.full(n_sig, -1, dtype=np.int64)
    n_persist = 0
    for s in range(n_sig):
        if persistent[s]:
            sig_slot[s] = n_persist
            n_persist += 1
    free: List[int] = []
    release: Dict[int, List[int]] = {}
    n_scratch = 0
    needs_slot = bytearray(n_sig)
    for i in range(n_ops):
        o = OUT[i]
        if not persistent[o] and not elide[i]:
            needs_slot[o] = 1
    # the guard signal must be loadable
    for b, e, g in regions:
        if not persistent[g]:
            needs_slot[g] = 1
    for i in range(n_ops):
        rel = release.pop(i, None)
        if rel:
            free.extend(rel)
        o = OUT[i]
        if persistent[o] or not needs_slot[o] or sig_slot[o] >= 0:
            continue
        slot = free.pop() if free else n_scratch
        if slot == n_scratch:
            n_scratch += 1
        sig_slot[o] = n_persist + slot
        when = max(last_read[o], last_write[o], i) + 1
        release.setdefault(when, []).append(slot)
    # a needed signal that is elided at its writer (guard signal read via GUARD only): give it a slot
    for b, e, g in regions:
        if sig_slot[g] < 0:
            sig_slot[g] = n_persist + n_scratch
            n_scratch += 1
            for i in range(n_ops):
                if OUT[i] == g:
                    elide[i] = False
    if n_persist + n_scratch + 1 >= SLOT_LIMIT:
        raise CircuitError("too many slots (%d) for a serial program (24-bit slot fields)" % (n_persist + n_scratch))

    # ---- operand kinds, LATE hazards per block, encoding ---------------------------------------------------
    rec = np.zeros((n_rec, 4), dtype=np.uint32)
    stored_in_block: set = set()
    early_loads: List[tuple] = []            # (position, slot) of every EARLY load, for the prefetcher
    touch: List[tuple] = []                  # (position, slot) of every access, for the cold test
    free_field: List[int] = [0] * n_rec      # 1 = a, 2 = b, 3 = c usable for a prefetch slot
    n_acc = n_elided = n_late = 0
    for p, (kind, i, reg) in enumerate(layout):
        if p % BLOCK == 0:
            stored_in_block = set()
        if kind == NOPR:
            free_field[p] = 1
            continue
        if kind == GUARDR:
            g = i
            ak = K_LATE                      # the enable is re-read from its row (an L1 hit)
            slot = int(sig_slot[g])
            rec[p, 0] = (slot if ak == K_LATE else 0) | (OP_GUARD << 24)
            rec[p, 1] = ak << 24
            rec[p, 3] = region_span[reg]
            if ak == K_LATE:
                touch.append((p, slot))
            free_field[p] = 2
            continue
        op = code[i]
        base = op & 0x7F
        arity = OP_ARITY[base]
        srcs = (A[i], B[i], C[i])[:arity]
        kinds, slots = [K_NONE] * 3, [0] * 3
        for k, s in enumerate(srcs):
            if reg >= 0 and s == regions[reg][2]:
                kinds[k] = K_GREG
            elif i > 0 and s == OUT[i - 1] and i not in after_region:
                kinds[k] = K_ACC
                n_acc += 1
            else:
                slot = int(sig_slot[s])
                if slot < 0:
                    raise AssertionError("operand without a slot at op %d" % i)
                slots[k] = slot
                if slot in stored_in_block:
                    kinds[k] = K_LATE
                    n_late += 1
                else:
                    kinds[k] = K_EARLY
                    early_loads.append((p, slot))
                touch.append((p, slot))
        if base in _COMMUTATIVE2 and kinds[1] == K_ACC and kinds[0] != K_ACC:
            kinds[0], kinds[1] = kinds[1], kinds[0]          # the accumulator operand goes first
            slots[0], slots[1] = slots[1], slots[0]
        flags = base | (F_NEG if op & OP_NEG else 0)
        out_slot = 0
        if not elide[i]:
            out_slot = int(sig_slot[OUT[i]])
            flags |= F_STORE
            stored_in_block.add(out_slot)
            touch.append((p, out_slot))
        else:
            n_elided += 1
        rec[p, 0] = slots[0] | (flags << 24)
        rec[p, 1] = slots[1] | ((kinds[0] | (kinds[1] << 3)) << 24)
        rec[p, 2] = slots[2] | (kinds[2] << 24)
        rec[p, 3] = out_slot
        for k in range(3):
            if kinds[k] in (K_NONE, K_ACC, K_GREG):
                free_field[p] = k + 1
                break

    # ---- prefetch: put the cold rows of the record `prefetch_distance` ahead into free fields -----------
    n_prefetch = 0
    if prefetch_distance > 0 and n_rec > 2 * prefetch_distance:
        last_touch: Dict[int, int] = {}
        for p, slot in touch:                        # previous step: positions p - n_rec
            last_touch[slot] = p - n_rec
        cold: List[tuple] = []
        touch_by_pos: Dict[int, List[int]] = {}
        for p, slot in touch:
            touch_by_pos.setdefault(p, []).append(slot)
        early_by_pos: Dict[int, List[int]] = {}
        for p, slot in early_loads:
            early_by_pos.setdefault(p, []).append(slot)
        for p in range(n_rec):
            for slot in early_by_pos.get(p, ()):
                if p - last_touch.get(slot, -10 ** 9) > COLD_DISTANCE:
                    cold.append((p, slot))
            for slot in touch_by_pos.get(p, ()):
                last_touch[slot] = p
        for p, slot in cold:
            q = (p - prefetch_distance) % n_rec
            for _ in range(prefetch_distance // 2):          # first record from there on with a free field
                if free_field[q]:
                    break
                q = (q + 1) % n_rec
            f = free_field[q]
            if not f:
                continue
            rec[q, f - 1] = (int(rec[q, f - 1]) & 0xFF000000) | slot
            rec[q, 1] = int(rec[q, 1]) | (f << 30)
            free_field[q] = 0
            n_prefetch += 1

    # ---- tag the blocks that match a pattern exactly (the hint only selects straight-line code for the
    # same records; a block that misses any precondition simply stays generic) ---------------------------
    n_pat = {PAT_MUX: 0, PAT_XM: 0}
    if patterns:
        M24 = 0xFFFFFF
        for b0 in range(0, n_rec, BLOCK):
            blk = [[int(x) for x in rec[p]] for p in range(b0, b0 + BLOCK)]
            fl = [w[0] >> 24 for w in blk]
            last = max([q + 1 for q in range(BLOCK) if (fl[q] & 15) != OP_NOP] or [0])
            rec[b0, 3] = int(rec[b0, 3]) | (last << 28)      # generic blocks: records worth decoding
            n = 0
            while n < BLOCK and (fl[n] & 15) != OP_NOP:
                n += 1
            if n == 0 or any((f & 15) != OP_NOP for f in fl[n:]):
                continue
            if any((w[1] >> 30) > 1 for w in blk[:n]):
                continue                                     # a prefetch slot anywhere but field a

            def is_m(w):
                return ((w[0] >> 24) == (OP_MUX | F_STORE) and ((w[1] >> 24) & 7) == K_GREG and
                        ((w[1] >> 27) & 7) in (K_ACC, K_EARLY) and ((w[2] >> 24) & 7) == K_EARLY and
                        (w[2] & M24) == (w[3] & M24))

            def is_x(w):
                return ((w[0] >> 24) == OP_XOR and ((w[1] >> 24) & 7) == K_ACC and ((w[1] >> 27) & 7) == K_EARLY and
                        ((w[2] >> 24) & 7) == K_NONE)

            pat = 0
            if all(is_m(w) for w in blk[:n]) and all(((w[1] >> 27) & 7) == K_ACC for w in blk[1:n]):
                pat = PAT_MUX | (n << 4)                     # only the head of a chain takes its data from a row
            elif n % 2 == 0 and all(is_x(blk[q]) and is_m(blk[q + 1]) and ((blk[q + 1][1] >> 27) & 7) == K_ACC
                                    for q in range(0, n, 2)):
                pat = PAT_XM | ((n // 2) << 4)
            if pat:
                rec[b0, 3] = (int(rec[b0, 3]) & 0xFFFFFF) | (pat << 24)
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
        net_sig=ops.net_sig, design=design, records=rec, level_off=np.array([0, n_rec], dtype=np.int32),
        n_levels=1 if n_rec else 0, has_latch=False, n_persist=n_persist, n_scratch=n_scratch,
        sig_slot=sig_slot, sig_shadow={}, sig_kept=kept, n_gates=int(is_gate.sum()), n_ops=n_ops,
        gate_words_bytes=gw_bytes, constants=constants, max_level_gates=n_ops, records_hinted=None, serial=True)
    guarded = sum(e - b for b, e, _ in regions)
    prog.serial_info = {"records": n_rec, "ops": n_ops, "acc_operands": n_acc, "stores_elided": n_elided,
                        "late_loads": n_late, "prefetches": n_prefetch, "regions": len(regions),
                        "guarded_ops": guarded, "padding": n_rec - n_ops - len(regions),
                        "blocks": n_rec // BLOCK, "blocks_mux_chain": n_pat[PAT_MUX], "blocks_xor_mux": n_pat[PAT_XM]}
    return prog


def compile_circuit_serial(root, map_pins: bool = True, **kw) -> Program:
    return compile_serial(compiler.flatten(root, map_pins), **kw)


# ----------------------------------------------------------------------------------------------------
# decoding (tests, the CPU test double, debugging)
# ----------------------------------------------------------------------------------------------------
def decode(records: np.ndarray):
    """Yield dicts for every record of a serial stream."""
    r = np.asarray(records, dtype=np.uint32)
    for p in range(len(r)):
        w0, w1, w2, w3 = (int(x) for x in r[p])
        flags = w0 >> 24
        yield {"pos": p, "op": flags & 15, "neg": bool(flags & F_NEG), "store": bool(flags & F_STORE),
               "keep_acc": bool(flags & F_KEEP_ACC),
               "slots": (w0 & 0xFFFFFF, w1 & 0xFFFFFF, w2 & 0xFFFFFF), "out": w3 & 0xFFFFFF,
               "kinds": ((w1 >> 24) & 7, (w1 >> 27) & 7, (w2 >> 24) & 7), "prefetch_field": w1 >> 30}
