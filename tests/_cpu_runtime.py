"""TEST DOUBLE - a numpy stand-in for ``mcircuit_b200.runtime.CudaRuntime``.

Lives under tests/ and is only ever injected by tests (``B200Engine(..., runtime=CpuRuntime())``)
so that the host logic - compiler output, tiling, pin access, sharding, the gloo multi-process
path - can be exercised without a GPU.  The product never imports it and has no CPU path.
It interprets the *compiled* program (records, levels, slots) the way the kernels do, which
also makes it the checker of the compiler against oracle/restate.py.
"""
import numpy as np

U64 = np.uint64
ALL = U64(0xFFFFFFFFFFFFFFFF)


class _State:
    def __init__(self, persist, pstride, pcol, scratch, sstride, scol, n_persist, n_slots, n_words):
        self.persist, self.pstride, self.pcol = persist, pstride, pcol
        self.scratch, self.sstride, self.scol = scratch, sstride, scol
        self.n_persist, self.n_slots, self.n_words = n_persist, n_slots, n_words

    def read(self, slots):
        slots = np.asarray(slots, dtype=np.int64)
        out = np.empty((len(slots), self.n_words), dtype=U64)
        cols = np.arange(self.n_words, dtype=np.int64)
        p = slots < self.n_persist
        if p.any():
            idx = slots[p][:, None] * self.pstride + self.pcol + cols[None, :]
            out[p] = self.persist[idx]
        if (~p).any():
            idx = (slots[~p] - self.n_persist)[:, None] * self.sstride + self.scol + cols[None, :]
            out[~p] = self.scratch[idx]
        return out

    def write(self, slots, values):
        slots = np.asarray(slots, dtype=np.int64)
        cols = np.arange(self.n_words, dtype=np.int64)
        p = slots < self.n_persist
        if p.any():
            idx = slots[p][:, None] * self.pstride + self.pcol + cols[None, :]
            self.persist[idx] = values[p]
        if (~p).any():
            idx = (slots[~p] - self.n_persist)[:, None] * self.sstride + self.scol + cols[None, :]
            self.scratch[idx] = values[~p]


def _apply(op, a, b, c):
    base = op & 0x7F
    if base == 1:
        r = a
    elif base == 2:
        r = a & b
    elif base == 3:
        r = a | b
    elif base == 4:
        r = a ^ b
    elif base == 5:
        r = a & ~b
    elif base == 6:
        r = (a & b) | (~a & c)
    elif base == 7:
        r = a & b & c
    elif base == 8:
        r = a | b | c
    elif base == 9:
        r = a ^ b ^ c
    elif base == 10:
        r = (a & b) | (a & c) | (b & c)
    else:
        raise ValueError("bad opcode %d" % op)
    return ~r if op & 0x80 else r


def splitmix64(x):
    x = np.asarray(x, dtype=U64)
    with np.errstate(over="ignore"):
        z = x + U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> U64(27))) * U64(0x94D049BB133111EB)
        return z ^ (z >> U64(31))


class CpuRuntime:
    name = "cpu-test-double"

    def __init__(self, memory=1 << 30):
        self._memory = memory
        self._launches = 0
        self.sm_count = 148
        self.l2_bytes = 126 << 20
        self.total_mem = memory

    def free_bytes(self):
        return self._memory

    def zeros_words(self, n):
        return np.zeros(int(n), dtype=U64)

    def empty_words(self, n):
        return np.zeros(int(n), dtype=U64)

    def to_device(self, a):
        return np.array(a, copy=True)

    def to_host_u64(self, a):
        return np.array(a, copy=True).view(U64)

    def synchronize(self):
        pass

    def make_state(self, persist, persist_stride, scratch, scratch_stride, n_persist, n_slots,
                   n_words, col_offset=0, scratch_col_offset=0):
        return _State(persist, persist_stride, col_offset, scratch, scratch_stride,
                      scratch_col_offset, n_persist, n_slots, n_words)

    # ---- program execution: level by level, all records of a level "at once" -----------------
    def _level(self, recs, st):
        if len(recs) == 0:
            return
        self._launches += 1
        op = recs[:, 0] & 0xFF
        live = (op & 0x7F) != 0
        recs, op = recs[live], op[live]
        if len(recs) == 0:
            return
        a = st.read(recs[:, 1])
        b = st.read(recs[:, 2])
        c = st.read(recs[:, 0] >> 8)
        out = np.empty_like(a)
        for code in np.unique(op):
            m = op == code
            out[m] = _apply(int(code), a[m], b[m], c[m])
        st.write(recs[:, 3], out)

    def eval_levels(self, d_records, h_level_off, lo, hi, st):
        for lv in range(lo, hi):
            self._level(d_records[h_level_off[lv]:h_level_off[lv + 1]].astype(np.int64), st)

    def latch(self, d_records, n, st):
        self._level(d_records[:n].astype(np.int64), st)

    def run(self, d_records, h_level_off, n_levels, st, steps, mode):
        batch = ((mode >> 8) & 0xFF) or 4
        hinted = (mode >> 24) & 1
        mode &= 0xFF
        if hinted:
            d_records = d_records.astype(np.int64)
            d_records[:, 1] &= 0x7FFFFFFF
            d_records[:, 2] &= 0x7FFFFFFF
        if mode == 0:
            assert all(int(x) % batch == 0 for x in h_level_off), "resident stream must be padded"
            # batches of records, loads before stores - exactly the kernel's hazard model
            recs = d_records.astype(np.int64)
            for _ in range(int(steps)):
                for i in range(0, int(h_level_off[n_levels]), batch):
                    self._level(recs[i:i + batch], st)
        else:
            for _ in range(int(steps)):
                self.eval_levels(d_records, h_level_off, 0, n_levels, st)

    # what the emulated producer sees: "early" = it fetches a block as soon as the kernel may (GB_MAX
    # blocks ahead, stopped by guards) and sees only stores that a fence has published; "late" = it
    # fetches at the last moment and sees everything.  A correct stream gives the same result in both.
    serial_fetch_model = "early"
    serial_cta_cols = 128

    def run_serial(self, d_stream, n_blocks, st, steps, word_bytes=0, d_skipped=None):
        """The stream kernel's execution model (csrc/mcb200.cu::k_run_stream), CTA by CTA
        (``serial_cta_cols`` lane words each): a producer stages the ring rows of each block ahead of
        the consumers; consumers run the records with ACC / GREG / STASH / RING / LATE operands; a guard
        is voted on by the whole CTA."""
        from collections import deque
        from mcircuit_b200 import serial as S
        lines = np.asarray(d_stream).astype(np.int64).reshape(-1, S.BLOCK_WORDS)[:n_blocks]
        blocks = [S.decode_block(l) for l in lines]
        self._launches += 1
        early = self.serial_fetch_model == "early"
        for c0 in range(0, st.n_words, self.serial_cta_cols):
            n = min(self.serial_cta_cols, st.n_words - c0)
            sub = _State(st.persist, st.pstride, st.pcol + c0, st.scratch, st.sstride, st.scol + c0,
                         st.n_persist, st.n_slots, n)
            acc = np.zeros(n, dtype=U64)
            greg = np.zeros(n, dtype=U64)
            stash = np.zeros((S.STASH_N, n), dtype=U64)
            unpublished = {}                       # slot -> value the async proxy still sees

            def fetch_row(slot):
                if early and slot in unpublished:
                    return unpublished[slot].copy()
                return sub.read([slot])[0].copy()

            def store(slot, val):
                if early and slot not in unpublished:
                    unpublished[slot] = sub.read([slot])[0].copy()
                sub.write([slot], val[None, :])

            queue = deque()                        # (block index, rows) staged and not yet consumed
            prod = {"b": 0, "step": 0, "blocked": False}

            def stage(blk):
                """What the producer's TMA boxes put into the ring group: box k covers 2^lg consecutive slots."""
                ring = [None] * blk["n_rows"]
                for run in blk["runs"]:
                    for k in range(run["rows"]):
                        ring[run["first_row"] + k] = fetch_row(run["slot"] + k)
                assert all(r is not None for r in ring), "ring rows not covered by the block's TMA runs"
                assert [run["slot"] + k for run in blk["runs"] for k in range(run["rows"])] == list(blk["rows"])
                return ring

            def pump():
                while not prod["blocked"] and prod["step"] < steps and len(queue) < (S.GB_MAX if early else 1):
                    blk = blocks[prod["b"]]
                    queue.append((prod["b"], stage(blk)))
                    if blk["guard"]:
                        prod["blocked"] = True
                    else:
                        prod["b"] += 1
                        if prod["b"] >= len(blocks):
                            prod["b"], prod["step"] = 0, prod["step"] + 1

            for step in range(int(steps)):
                b = 0
                while b < len(blocks):
                    pump()
                    qb, ring = queue.popleft()
                    assert qb == b, "producer and consumer disagree on the block sequence"
                    blk = blocks[b]
                    if blk["fence"]:
                        unpublished.clear()

                    def operand(kind, field):
                        if kind == S.K_RING:
                            return ring[field]
                        if kind == S.K_LATE:
                            return sub.read([field])[0]
                        if kind == S.K_ACC:
                            return acc
                        if kind == S.K_GREG:
                            return greg
                        if kind == S.K_STASH:
                            return stash[field].copy()
                        return np.zeros(n, dtype=U64)

                    recs = blk["records"]
                    guard_rec = None
                    for seg in blk["segments"]:
                        f, n, r0 = seg["first_rec"], seg["count"], seg["first_row"]
                        if seg["kind"] == S.SEG_MUX:          # the kernel's conventions: ring row r0 + u, store slot = its slot
                            d = operand(recs[f]["kinds"][1], recs[f]["fields"][1])
                            for u in range(n):
                                d = (greg & d) | (~greg & ring[r0 + u])
                                store(blk["rows"][r0 + u], d)
                            acc = d
                            if seg["stash_out"] >= 0:         # only the segment's last value can go to the stash
                                stash[seg["stash_out"]] = acc
                        elif seg["kind"] == S.SEG_XM and seg["q_by_record"]:
                            for u in range(n):                # q from wherever the stage's XOR record says, old = row r0 + u
                                x = recs[f + 2 * u]
                                q = operand(x["kinds"][1], x["fields"][1])
                                acc = (greg & (acc ^ q)) | (~greg & ring[r0 + u])
                                store(blk["rows"][r0 + u], acc)
                            if seg["stash_out"] >= 0:
                                stash[seg["stash_out"]] = acc
                        elif seg["kind"] == S.SEG_XM:
                            for u in range(n):
                                acc = (greg & (acc ^ ring[r0 + u])) | (~greg & ring[r0 + n + u])
                                store(blk["rows"][r0 + n + u], acc)
                            if seg["stash_out"] >= 0:
                                stash[seg["stash_out"]] = acc
                        elif seg["kind"] == S.SEG_GUARD:
                            guard_rec = recs[f]
                        else:
                            for r in recs[f:f + n]:
                                vals = [operand(k, fl) for k, fl in zip(r["kinds"], r["fields"])]
                                res = _apply(int(r["op"] | (0x80 if r["neg"] else 0)), vals[0], vals[1], vals[2])
                                if r["store"]:
                                    store(r["out"], res)
                                if r["stash_out"] >= 0:
                                    stash[r["stash_out"]] = res
                                acc = res
                    skip = 0
                    if blk["guard"]:
                        g = guard_rec
                        greg = operand(g["kinds"][0], g["fields"][0]).copy()
                        unpublished.clear()                    # consumers fence before they vote
                        if not greg.any():
                            skip = blk["span"]
                            if d_skipped is not None and c0 == 0:
                                d_skipped[0] += U64(1)
                        nb = b + 1 + skip
                        prod["blocked"] = False
                        prod["b"] = nb
                        if prod["b"] >= len(blocks):
                            prod["b"], prod["step"] = 0, prod["step"] + 1
                    b += 1 + skip
            assert not queue or prod["step"] >= steps

    def eval_levels_grouped(self, d_records, h_level_off, lo, hi, st):
        recs = d_records.astype(np.int64)
        recs[:, 1] &= 0x7FFFFFFF
        recs[:, 2] &= 0x7FFFFFFF
        self.eval_levels(recs, h_level_off, lo, hi, st)

    def run_pipelined(self, d_records, d_level_off, h_level_off, n_levels, st, tile_words, in_flight,
                      scratch_plane_words, d_sync, steps):
        recs = d_records.astype(np.int64)
        recs[:, 1] &= 0x7FFFFFFF          # strip the streaming hints
        recs[:, 2] &= 0x7FFFFFFF
        assert np.array_equal(np.asarray(d_level_off), np.asarray(h_level_off))
        assert all(int(x) % 4 == 0 for x in h_level_off)
        n_tiles = (st.n_words + tile_words - 1) // tile_words
        for t0 in range(0, n_tiles, in_flight):
            for _ in range(int(steps)):
                for lv in range(n_levels):                      # tiles of a group advance level by level
                    for g in range(in_flight):
                        tile = t0 + g
                        if tile >= n_tiles:
                            continue
                        c0 = tile * tile_words
                        sub = _State(st.persist, st.pstride, st.pcol + c0, st.scratch, st.sstride,
                                     st.scol + g * scratch_plane_words, st.n_persist, st.n_slots,
                                     min(tile_words, st.n_words - c0))
                        self.eval_levels(recs, h_level_off, lv, lv + 1, sub)

    # ---- K4 / K5 ---------------------------------------------------------------------------------
    def signature(self, d_slots, n, st, col_offset, d_popcount, d_checksum):
        rows = st.read(d_slots[:n])
        cols = np.arange(st.n_words, dtype=U64) + U64(col_offset)
        with np.errstate(over="ignore"):
            if d_popcount is not None:
                pc = np.array([sum(bin(int(w)).count("1") for w in r) for r in rows], dtype=U64)
                d_popcount[:n] = pc
            if d_checksum is not None:
                d_checksum[:n] = (rows * (U64(2) * cols + U64(1))[None, :]).sum(axis=1, dtype=U64)
        self._launches += 1

    def toggle_accum(self, d_slots, n, st, d_shadow, shadow_stride, d_toggles):
        rows = st.read(d_slots[:n])
        for i in range(n):
            sh = d_shadow[i * shadow_stride: i * shadow_stride + st.n_words]
            d_toggles[i] += U64(sum(bin(int(w)).count("1") for w in (rows[i] ^ sh)))
            sh[:] = rows[i]
        self._launches += 1

    def stimulus(self, d_slots, d_streams, n, st, col_offset, seed):
        cols = np.arange(st.n_words, dtype=U64) + U64(col_offset)
        s = np.asarray(d_streams[:n]).astype(U64)
        vals = splitmix64(U64(seed) ^ (s[:, None] << U64(32)) ^ cols[None, :])
        st.write(d_slots[:n], vals)
        self._launches += 1

    def fill_rows(self, d_slots, d_values, n, st):
        vals = np.repeat(np.asarray(d_values[:n]).view(U64)[:, None], st.n_words, axis=1)
        st.write(d_slots[:n], vals)
        self._launches += 1

    def pack(self, d_values, width, d_slots, st):
        v = np.asarray(d_values).view(U64)[: st.n_words * 64].reshape(st.n_words, 64)
        weights = U64(1) << np.arange(64, dtype=U64)
        planes = np.empty((width, st.n_words), dtype=U64)
        for b in range(width):
            planes[b] = (((v >> U64(b)) & U64(1)) * weights).sum(axis=1, dtype=U64)
        st.write(d_slots[:width], planes)
        self._launches += 1

    def unpack(self, d_values, width, d_slots, st):
        planes = st.read(d_slots[:width])
        shifts = np.arange(64, dtype=U64)
        out = np.zeros(st.n_words * 64, dtype=U64)
        for b in range(width):
            out |= (((planes[b][:, None] >> shifts[None, :]) & U64(1)).reshape(-1)) << U64(b)
        d_values[: st.n_words * 64] = out
        self._launches += 1

    def capture_rows(self, d_slots, n, st, d_dst, dst_word_offset, dst_stride):
        rows = st.read(d_slots[:n])
        for i in range(n):
            o = dst_word_offset + i * dst_stride
            d_dst[o:o + st.n_words] = rows[i]
        self._launches += 1

    def set_tuning(self, key, value):
        return 0

    def launch_count(self, reset=False):
        n = self._launches
        if reset:
            self._launches = 0
        return n
