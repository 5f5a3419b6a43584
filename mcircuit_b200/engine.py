"""B200Engine - the Executor of the drop-in simulation path.

Mirrors the reference's ``JIT(root, burst_size, map_pins)`` (``core/simulator.py:174-297``):
``get_pin_state``, ``set_pin_state``, ``step``, ``burst`` with the reference's semantics
(pin paths, aliasing, ``burst`` = burst_size + 1 steps, constants visible before the first
step, zero-initialised nets), and adds the multi-lane surface the reference does not have:
``lanes=N`` independent stimulus instances evaluated at once, bit-sliced 64 per word.

Documented deviations from the reference:
  * values written with ``set_pin_state`` are masked to the pin width (reference writes 64
    raw bits over an iW global, simulator.py:289-291);
  * no ``out.txt`` side effect (simulator.py:266-267);
  * ``Register`` follows the evident intent of simulator.py:115-127, which does not compile
    in the reference;
  * multi-driver nets, wire loops and width mismatches raise at build time.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import compiler
from .compiler import PinNotObservable, Program
from .runtime import MODE_LEVELS, MODE_RESIDENT

SEED = 0x6D63697263756974      # "mcircuit"
_ALL = 0xFFFFFFFFFFFFFFFF
_ALIGN_WORDS = 16              # rows start on 128-byte boundaries
DEFAULT_TILE_WORDS = 512       # 4 KB rows: a level's rows (20 MB for C5) survive in L2 until the next level
DEFAULT_STREAMS = 2            # two tiles interleaved; best measured (profiles/r01_sweep_streams_graphs_honest.jsonl)


def _round_up(x, m):
    return (x + m - 1) // m * m


class Executor:
    """The reference's abstract interface (core/simulator.py:160-171)."""

    def get_pin_state(self, pin):
        raise NotImplementedError

    def set_pin_state(self, pin, value):
        raise NotImplementedError

    def step(self):
        raise NotImplementedError

    def burst(self):
        raise NotImplementedError


class B200Engine(Executor):
    def __init__(self, root, burst_size=500, map_pins=True, *, lanes=64, device=None,
                 keep="auto", observe: Sequence[str] = (), tile_words: Optional[int] = None,
                 mode="auto", runtime=None, lane_word_offset=0, program: Optional[Program] = None,
                 memory_budget: Optional[int] = None, back_edges="auto", resident_batch=0,
                 resident_word_bytes=0, in_flight=2, streams=None, level_hints=False,
                 share_edge_detectors=None, sort_level_by_fanin=False, serial_guard=True,
                 serial_word_bytes=0, serial_options: Optional[dict] = None, adder_prefix=False):
        if lanes < 1 or lanes % 64:
            raise ValueError("lanes must be a positive multiple of 64 (one 64-bit word = 64 lanes)")
        # Registers / Counters on one clock share one edge detector by default (29,089 -> 10,707 micro-ops on C4, and
        # what lets K3 skip idle clock phases); a poke of ONE member's prevclock un-shares the engine on the fly
        # (_unshare_edge_detectors), so the reference's per-primitive prevclock semantics are kept
        if share_edge_detectors is None:
            share_edge_detectors = program is None
        self._ctor_kwargs = dict(lanes=lanes, device=device, keep=keep, observe=observe, tile_words=tile_words, mode=mode,
                                 lane_word_offset=lane_word_offset, memory_budget=memory_budget, back_edges=back_edges,
                                 resident_batch=resident_batch, resident_word_bytes=resident_word_bytes, in_flight=in_flight,
                                 streams=streams, level_hints=level_hints, share_edge_detectors=share_edge_detectors,
                                 sort_level_by_fanin=sort_level_by_fanin, serial_guard=serial_guard,
                                 serial_word_bytes=serial_word_bytes, serial_options=serial_options, adder_prefix=adder_prefix)
        if runtime is None:
            from .runtime import CudaRuntime
            runtime = CudaRuntime(device)           # raises when there is no B200 / no .so
        self.rt = runtime
        self.root = root
        self.burst_size = int(burst_size)
        self.map_pins = bool(map_pins)
        self.lanes = int(lanes)
        self.words = self.lanes // 64
        self.lane_word_offset = int(lane_word_offset)   # global column of our column 0 (sharding)
        self.stride = _round_up(self.words, _ALIGN_WORDS)
        budget = memory_budget if memory_budget is not None else int(self.rt.free_bytes() * 0.80)

        # ---- compile ------------------------------------------------------------------------
        if program is None:
            design = compiler.flatten(root, map_pins)
            if keep == "auto":
                keep = "all" if (design.n_signals * 3 // 2) * self.stride * 8 <= budget // 2 else "ports"
            if mode == "serial":
                from . import serial
                program = serial.compile_serial(design, keep=keep, observe=observe,
                                                share_edge_detectors=share_edge_detectors,
                                                guard=serial_guard, adder_prefix=adder_prefix, **(serial_options or {}))
            else:
                program = compiler.compile_design(design, keep=keep, observe=observe, back_edges=back_edges,
                                                  share_edge_detectors=share_edge_detectors,
                                                  sort_level_by_fanin=sort_level_by_fanin, adder_prefix=adder_prefix)
                if mode == "auto" and (program.n_ops / max(1, program.n_levels)) * min(self.words, DEFAULT_TILE_WORDS) < (1 << 19):
                    # narrow levels (sequential logic, long dependent chains): per-level launches would be launch
                    # bound; run the reference's own emit order as one stream instead (K3: one launch for any
                    # number of steps, rows staged by TMA, idle clock phases skipped)
                    from . import serial
                    try:
                        stream = serial.compile_serial(design, keep=keep, observe=observe,
                                                       share_edge_detectors=share_edge_detectors, guard=serial_guard,
                                                       adder_prefix=adder_prefix, **(serial_options or {}))
                        # ... but only when the stream is made of what K3 is good at - enabled-register chains and
                        # pipeline stages (consecutive rows in few TMA boxes, straight-line code).  A stream of generic
                        # records with scattered operands costs a TMA box per operand and ~45 instructions per record:
                        # the resident kernel on the levelized batches is 1.5 - 5 x faster there
                        # (profiles/r02_modes_compare.jsonl)
                        info = stream.serial_info
                        n_rec = max(1, info["records"])
                        registers = info["mux_chain_units"] + 2 * info["xor_mux_units"] >= 0.5 * n_rec
                        # ... or a stream whose operands are mostly forwarded through registers / stash (gate-level
                        # flip-flops, ripple carries): few TMA boxes, hardly any late load
                        forwarded = info["tma_boxes"] <= 0.25 * n_rec and info["late_operands"] <= 0.05 * n_rec
                        if registers or forwarded:
                            program = stream
                    except compiler.CircuitError:
                        pass                                     # more than 2^24 slots: stay with the level program
        if getattr(program, "serial", False):
            if mode not in ("auto", "serial"):
                raise ValueError("a serial program runs in mode='serial' only")
            mode = "serial"
        elif mode == "serial":
            raise ValueError("mode='serial' needs a program made by serial.compile_serial")
        self.program = program
        self.keep = keep
        # one extra scratch row that padding records of the resident kernel read and write
        P, S = program.n_persist, program.n_scratch + 1
        self.trash_slot = P + S - 1

        self.in_flight = max(1, min(4, int(in_flight)))     # tiles in flight of the pipelined kernel
        # level mode: independent tiles alternate over this many CUDA streams (each with its own
        # scratch plane) so one tile's launch gaps and kernel tails are filled by another's work
        if streams is None:
            streams = DEFAULT_STREAMS
        self.n_streams = max(1, min(4, int(streams))) if hasattr(self.rt, "torch") else 1
        self._streams = None

        # ---- tiling ------------------------------------------------------------------------------
        persist_bytes = P * self.stride * 8
        avg_level = program.n_ops / max(1, program.n_levels)
        if mode == "auto":
            # per-level launches cost a launch gap per level (~3.3 us) or the level's bytes at HBM speed; one resident
            # launch walks batches of 8 records at one dependent memory round trip each (~0.75 us) or its bytes at a
            # lower rate.  Pick the smaller estimate (checked against profiles/r02_modes_compare.jsonl: a 20k-gate,
            # 40-level DAG wants levels even at 2^16 lanes, a 1.4k-gate multiplier or a gate-level LFSR wants resident)
            tw = float(min(self.words, DEFAULT_TILE_WORDS))
            t_levels = program.n_levels * max(3.3e-6, avg_level * tw * 24.0 / 6.0e12)
            t_resident = max((program.n_ops / 8.0 + program.n_levels) * 0.75e-6, program.n_ops * tw * 24.0 / 4.0e12)
            mode = "levels" if (t_levels <= t_resident or self.trash_slot >= compiler.IN2_LIMIT) else "resident"
        wide = mode in ("levels", "levels_grouped")
        if tile_words is None:
            fits = (P + S) * self.stride * 8 <= budget
            # a level launch should carry as many gate-words as a 5,000-gate level of a 512-word tile does
            # (the measured sweet spot on C5); narrow levels (C3: 64 gates) get proportionally wider tiles
            target = DEFAULT_TILE_WORDS
            if avg_level < 5000:
                target = _round_up(int(min(self.words, DEFAULT_TILE_WORDS * 5000 / max(avg_level, 1.0))), _ALIGN_WORDS)
            if fits and not (wide and self.n_streams > 1 and self.words > 2 * target):
                tile_words = self.words                      # everything resident, one tile
            else:
                # wide level sweeps run best as small tiles (L2-resident levels) alternating over 2 streams
                room = max(budget - persist_bytes, 0)
                planes = max(self.n_streams, self.in_flight)
                tile_words = max(_ALIGN_WORDS, min(target,
                                                   (room // max(1, S * 8 * planes)) // _ALIGN_WORDS * _ALIGN_WORDS))
        tile_words = int(min(tile_words, self.words))
        if tile_words < self.words:
            tile_words = max(_ALIGN_WORDS, tile_words // _ALIGN_WORDS * _ALIGN_WORDS)
        self.tile_words = tile_words
        self.n_tiles = (self.words + tile_words - 1) // tile_words
        if persist_bytes + S * _round_up(tile_words, _ALIGN_WORDS) * 8 * (max(self.n_streams, self.in_flight)
                                                                            if tile_words < self.words else 1) > budget:
            raise MemoryError(
                "state does not fit: %d persistent slots x %d words + %d scratch slots x %d words"
                % (P, self.stride, S, tile_words))

        # ---- execution mode -------------------------------------------------------------------
        if mode not in ("levels", "levels_grouped", "resident", "persistent", "serial"):
            raise ValueError("mode must be 'auto', 'levels', 'levels_grouped', 'resident', 'serial' or 'persistent'")
        self.mode = mode
        # padding records of the resident / grouped / pipelined streams carry the trash slot as
        # in2, next to the opcode; plain level launches have no padding and no such limit
        self.serial_word_bytes = int(serial_word_bytes)
        self._skipped = None        # device counter of guarded regions warp 0 jumped over (serial mode)
        if mode not in ("levels", "serial") and self.trash_slot >= compiler.IN2_LIMIT:
            raise compiler.CircuitError("too many slots (%d) for the %s kernel's padding records; use mode='levels'"
                                        % (self.trash_slot, mode))

        # ---- buffers (torch owns them) ---------------------------------------------------------------
        if self.n_tiles == 1:
            # one plane: scratch rows follow the persistent rows with the same stride
            self._plane = self.rt.zeros_words((P + S) * self.stride + _ALIGN_WORDS)
            self._scratch = None
            self.scratch_stride = self.stride
        else:
            self._plane = self.rt.zeros_words(P * self.stride + _ALIGN_WORDS)
            self.scratch_stride = _round_up(tile_words, _ALIGN_WORDS)
            planes = self.in_flight if mode == "persistent" else self.n_streams
            self._scratch = self.rt.zeros_words(planes * S * self.scratch_stride + _ALIGN_WORDS)
        # level mode streams old fan-in through L2 with evict-first (hint bit 31 of in0 / in1)
        self.level_hints = bool(level_hints) and getattr(program, "records_hinted", None) is not None
        self._d_records = self.rt.to_device(program.records_hinted if self.level_hints else program.records)
        self._h_level_off = np.ascontiguousarray(program.level_off, dtype=np.int32)
        self._resident = None      # (d_records in batches, h_level_off), built on first use
        # resident kernel: records per batch (4 or 8) and bytes of a lane word per thread (0 = auto)
        self.resident_batch = int(resident_batch) if resident_batch else 8
        self.resident_word_bytes = int(resident_word_bytes)
        self.resident_batches = None
        self._whole = self._state_for(0, self.words, whole=True)
        self._slot_cache: Dict[str, object] = {}
        self.steps_done = 0
        self._toggle = None

        for base, width, value in program.constants:           # simulator.py:275-276
            self._poke_signals(base, width, int(value))

    def close(self):
        """Drop THIS engine's cached step graphs (they hold raw pointers into its buffers);
        other engines' graphs stay."""
        clear = getattr(self.rt, "graph_cache_clear", None)
        recs = getattr(self, "_d_records", None)
        if clear is not None and recs is not None and hasattr(recs, "data_ptr"):
            clear(recs)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    # state views
    # ------------------------------------------------------------------------------------------
    def _state_for(self, col0, n_words, whole=False, plane=0):
        P, S = self.program.n_persist, self.program.n_scratch + 1
        if self.n_tiles == 1:
            return self.rt.make_state(self._plane, self.stride, self._plane, self.stride, P, P + S,
                                      n_words, col_offset=col0,
                                      scratch_col_offset=P * self.stride + col0)
        if whole:
            # persistent rows only (pin access, stimulus, signatures)
            return self.rt.make_state(self._plane, self.stride, None, 0, P, P, n_words, col_offset=col0)
        return self.rt.make_state(self._plane, self.stride, self._scratch, self.scratch_stride,
                                  P, P + S, n_words, col_offset=col0,
                                  scratch_col_offset=plane * S * self.scratch_stride)

    def _tiles(self):
        for t in range(self.n_tiles):
            c0 = t * self.tile_words
            yield c0, min(self.tile_words, self.words - c0)

    # ------------------------------------------------------------------------------------------
    # Executor interface (core/simulator.py:285-297)
    # ------------------------------------------------------------------------------------------
    def step(self):
        self.run(1)

    def burst(self):
        # the reference's loop tests the pre-decrement counter: burst_size + 1 passes (Q1)
        self.run(self.burst_size + 1)

    def tick(self):
        """One full clock period of a ``Clock`` = two steps (it toggles every step)."""
        self.run(2)

    def run(self, steps: int):
        """``steps`` evaluation passes with all state resident on the device."""
        steps = int(steps)
        if steps <= 0 or self.program.n_ops == 0:
            self.steps_done += max(steps, 0)
            return
        if self._toggle is not None:
            for _ in range(steps):
                self._run_raw(1)
                self._toggle_sample()
        else:
            self._run_raw(steps)
        self.steps_done += steps

    def _run_raw(self, steps):
        if self.mode == "persistent":
            self._run_pipelined(steps)
            return
        if hasattr(self.rt, "torch") and self.n_tiles > 1 and self.mode in ("levels", "levels_grouped"):
            torch = self.rt.torch
            main = torch.cuda.current_stream(self.rt.device)
            streams = self._compute_streams()
            for st in streams:
                st.wait_stream(main)
            for t, (c0, n) in enumerate(self._tiles()):
                k = t % len(streams)
                with torch.cuda.stream(streams[k]):
                    self._run_tile(c0, n, steps, plane=k)
            for st in streams:
                main.wait_stream(st)
            return
        for c0, n in self._tiles():
            self._run_tile(c0, n, steps)

    def _compute_streams(self):
        if self._streams is None:
            torch = self.rt.torch
            self._streams = [torch.cuda.Stream(self.rt.device) for _ in range(self.n_streams)]
        return self._streams

    def _grouped(self):
        """Grouped (per level, per opcode, padded to 4) hinted record stream on the device."""
        if getattr(self, "_grouped_stream", None) is None:
            rec, off = compiler.group_levels(self.program.records_hinted, self.program.level_off,
                                             self.trash_slot, 4)
            self._grouped_stream = (self.rt.to_device(rec), self.rt.to_device(off), off)
        return self._grouped_stream

    def _run_pipelined(self, steps):
        """K1pp: every tile through every level in one cooperative launch, ``in_flight`` tiles
        at a time (split-phase grid barriers)."""
        d_rec, d_off, h_off = self._grouped()
        if getattr(self, "_sync", None) is None:
            self._sync = self.rt.zeros_words(8)
        P, S = self.program.n_persist, self.program.n_scratch + 1
        if self.n_tiles == 1:
            st = self._state_for(0, self.words)
            self.rt.run_pipelined(d_rec, d_off, h_off, self.program.n_levels, st, _round_up(self.words, 2), 1,
                                  0, self._sync, steps)
        else:
            st = self.rt.make_state(self._plane, self.stride, self._scratch, self.scratch_stride, P, P + S,
                                    self.words)
            self.rt.run_pipelined(d_rec, d_off, h_off, self.program.n_levels, st, self.tile_words,
                                  self.in_flight, S * self.scratch_stride, self._sync, steps)

    def _run_tile(self, c0, n, steps, plane=0):
        """All ``steps`` passes of one lane tile (tiles are independent: lanes never interact)."""
        if self.mode == "serial":
            if self._skipped is None:
                self._skipped = self.rt.zeros_words(1)
            # a ring group only has to hold as many rows as the widest block stages
            rows = int(getattr(self.program, "serial_info", {}).get("max_ring_rows", 0))
            self.rt.run_serial(self._d_records, len(self.program.records), self._state_for(c0, n), steps,
                               self.serial_word_bytes | (max(rows, 1) << 8), self._skipped)
        elif self.mode == "levels_grouped":
            d_rec, _, h_off = self._grouped()
            st = self._state_for(c0, n, plane=plane)
            for _ in range(int(steps)):
                self.rt.eval_levels_grouped(d_rec, h_off, 0, self.program.n_levels, st)
        elif self.mode == "resident":
            if self._resident is None:
                rec, n_batches = compiler.schedule_batches(self.program.records, self.program.level_off,
                                                           self.trash_slot, self.resident_batch)
                off = np.array([0, len(rec)], dtype=np.int32)
                self._resident = (self.rt.to_device(rec) if len(rec) else None, off)
                self.resident_batches = n_batches
            d_rec, off = self._resident
            if d_rec is not None:
                self.rt.run(d_rec, off, 1, self._state_for(c0, n), steps,
                            MODE_RESIDENT | (self.resident_batch << 8) | (self.resident_word_bytes << 16))
        else:
            self.rt.run(self._d_records, self._h_level_off, self.program.n_levels,
                        self._state_for(c0, n, plane=plane), steps,
                        MODE_LEVELS | ((1 << 24) if self.level_hints else 0))

    def run_host(self, in_plan, h_in, out_plan, h_out, steps: int = 1, io_chunk_words: int = 4096):
        """Host buffers in, host buffers out, pipelined by lane tile: the stimulus rows of
        tile t+1 are uploaded (copy stream) while tile t is evaluated (compute stream) and
        the result rows of tile t-1 are downloaded (second copy stream).  ``h_in`` /
        ``h_out``: pinned int64 tensors [n_bits, words] in the row order of the io plans
        (bit planes, the layout of ``get_pin_planes``).  Returns after enqueueing; the
        caller's current stream is joined with the download stream."""
        torch = self.rt.torch
        dev = self.rt.device
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = self._copy_streams
        main = torch.cuda.current_stream(dev)
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        base = self._plane.data_ptr()
        pitch = self.stride * 8
        in_pitch, out_pitch = h_in.stride(0) * 8, h_out.stride(0) * 8
        tiles = list(self._tiles())
        # I/O moves in chunks of several compute tiles: a strided copy of 4 KB row segments (one
        # 512-word tile) is descriptor-bound on the DMA engines, 32 KB segments are not
        per_chunk = max(1, io_chunk_words // max(1, self.tile_words))
        chunks = [tiles[i:i + per_chunk] for i in range(0, len(tiles), per_chunk)]
        ready = []
        for chunk in chunks:
            c0 = chunk[0][0]
            n = sum(t[1] for t in chunk)
            for first, count, k in in_plan["runs"]:
                self.rt.copy2d(base + (first * self.stride + c0) * 8, pitch,
                               h_in.data_ptr() + k * in_pitch + c0 * 8, in_pitch, n * 8, count, True, s_in)
            ev = torch.cuda.Event()
            ev.record(s_in)
            ready.append(ev)
        multi = len(tiles) > 1 and self.mode in ("levels", "levels_grouped")
        comp = self._compute_streams() if multi else [main]
        if multi:
            for cs in comp:
                cs.wait_stream(main)
        t = 0
        for chunk, ev in zip(chunks, ready):
            used = set()
            for c0, n in chunk:
                cs = comp[t % len(comp)]
                if cs not in used:
                    cs.wait_event(ev)
                    used.add(cs)
                with torch.cuda.stream(cs):
                    self._run_tile(c0, n, steps, plane=(t % len(comp)) if multi else 0)
                t += 1
            for cs in used:
                done = torch.cuda.Event()
                done.record(cs)
                s_out.wait_event(done)
            c0 = chunk[0][0]
            n = sum(x[1] for x in chunk)
            for first, count, k in out_plan["runs"]:
                self.rt.copy2d(h_out.data_ptr() + k * out_pitch + c0 * 8, out_pitch,
                               base + (first * self.stride + c0) * 8, pitch, n * 8, count, False, s_out)
        if multi:
            for cs in comp:
                main.wait_stream(cs)
        main.wait_stream(s_out)
        self.steps_done += int(steps)

    # ---- pins --------------------------------------------------------------------------------------
    def _slots_of(self, pin: str, need_value=True):
        key = (pin, need_value)
        hit = self._slot_cache.get(key)
        if hit is None:
            base, width = self.program.pin_signals(pin)      # KeyError on unknown pin (Q10)
            slots = self.program.pin_slots(pin, need_value=need_value)
            shadows = [self.program.sig_shadow.get(base + b, -1) for b in range(width)]
            hit = (base, width, slots, shadows)
            self._slot_cache[key] = hit
        return hit

    def _poke_signals(self, base, width, value: int):
        slots, vals = [], []
        for b in range(width):
            s = int(self.program.sig_slot[base + b])
            if s < 0 or s >= self.program.n_persist:
                continue                 # recycled net: a poke could not survive anyway
            word = _ALL if (value >> b) & 1 else 0
            slots.append(s)
            vals.append(word)
            sh = self.program.sig_shadow.get(base + b)
            if sh is not None:
                slots.append(sh)
                vals.append(word)
        if slots:
            self.rt.fill_rows(self.rt.to_device(np.asarray(slots, dtype=np.int32)),
                              self.rt.to_device(np.asarray(vals, dtype=np.uint64)), len(slots), self._whole)

    def _before_poke(self, pin):
        """A poke of one member's ``prevclock`` cannot be expressed while a group shares that bit."""
        shared = getattr(self.program, "shared_prev", None)
        if shared and self.program.design is not None:
            base, _ = self.program.pin_signals(pin)            # KeyError on an unknown pin, as always
            if base in shared:
                self._unshare_edge_detectors()

    def set_pin_state(self, pin, value):
        """Scalar: every lane gets ``value`` (masked to the pin width).  Array of ``lanes``
        uint64: one value per lane."""
        self._before_poke(pin)
        base, width, slots, shadows = self._slots_of(pin, need_value=False)
        if np.isscalar(value) or isinstance(value, int):
            self._poke_signals(base, width, int(value) & ((1 << width) - 1))
            return
        if hasattr(value, "detach"):                            # a torch tensor, on any device
            value = value.detach().cpu().numpy()
        v = np.ascontiguousarray(np.asarray(value).astype(np.uint64, copy=False))
        if v.shape != (self.lanes,):
            raise ValueError("expected %d lane values, got shape %s" % (self.lanes, v.shape))
        d_vals = self.rt.to_device(v)
        for group in (slots, shadows):
            if any(s < 0 or s >= self.program.n_persist for s in group):
                if group is slots:
                    raise PinNotObservable("%s is not retained; cannot drive it per lane" % pin)
                continue
            self.rt.pack(d_vals, width, self.rt.to_device(np.asarray(group, dtype=np.int32)), self._whole)

    def get_pin_state(self, pin, lanes=False):
        """Scalar form returns lane 0 as a Python int (the reference's single instance);
        ``lanes=True`` returns all lanes as uint64[lanes]."""
        base, width, slots, _ = self._slots_of(pin)
        if lanes:
            d_vals = self.rt.empty_words(self.lanes)
            self.rt.unpack(d_vals, width, self.rt.to_device(np.asarray(slots, dtype=np.int32)), self._whole)
            return self.rt.to_host_u64(d_vals)
        idx = np.asarray(slots, dtype=np.int64) * self.stride
        words = self.rt.to_host_u64(self._plane[self.rt.to_device(idx)])
        value = 0
        for b in range(width):
            value |= (int(words[b]) & 1) << b
        return value

    # bit-plane access: rows of the slot-major state, no transposition
    def set_pin_planes(self, pin, planes):
        """planes: uint64[width, words] - bit b of every lane word, already bit-sliced."""
        self._before_poke(pin)
        base, width, slots, shadows = self._slots_of(pin, need_value=False)
        p = np.ascontiguousarray(np.asarray(planes, dtype=np.uint64)).reshape(width, self.words)
        d = self.rt.to_device(p)
        for group in (slots, shadows):
            for b, s in enumerate(group):
                if 0 <= s < self.program.n_persist:
                    self._plane[s * self.stride: s * self.stride + self.words] = d[b]

    def get_pin_planes(self, pin) -> np.ndarray:
        base, width, slots, _ = self._slots_of(pin)
        out = np.empty((width, self.words), dtype=np.uint64)
        for b, s in enumerate(slots):
            out[b] = self.rt.to_host_u64(self._plane[s * self.stride: s * self.stride + self.words])
        return out

    def sample_planes(self, pins: Sequence[str], cols) -> np.ndarray:
        """uint64[n_bits, len(cols)]: the lane words of every bit of ``pins`` at the lane-word columns
        ``cols`` only - a gather of n_bits x len(cols) words, however wide the rows are."""
        slots = np.asarray(self._flat_slots(pins), dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        flat = slots[:, None] * self.stride + cols[None, :]
        return self.rt.to_host_u64(self._plane[self.rt.to_device(flat.reshape(-1))]).reshape(len(slots), len(cols))

    # ---- stimulus / signatures --------------------------------------------------------------------------
    def input_pins(self) -> List[str]:
        """Root-level IN ports, in child order."""
        if self.program.design is None:                     # program loaded from a cache
            return list(self.program.ports_in)
        d = self.program.design
        out = []
        for name, child in d.top.children.items():
            if isinstance(child, int):
                desc = d.leaf_desc[child]
                if type(desc).__name__ == "ExposedPin" and desc.direction == 0:
                    out.append("/%s/pin" % name)
        return out

    def output_pins(self) -> List[str]:
        if self.program.design is None:
            return list(self.program.ports_out)
        d = self.program.design
        out = []
        for name, child in d.top.children.items():
            if isinstance(child, int):
                desc = d.leaf_desc[child]
                if type(desc).__name__ == "ExposedPin" and desc.direction == 1:
                    out.append("/%s/pin" % name)
        return out

    def _flat_slots(self, pins: Sequence[str]) -> List[int]:
        out: List[int] = []
        for p in pins:
            out.extend(self._slots_of(p)[2])
        return out

    def randomize_inputs(self, pins: Optional[Sequence[str]] = None, seed: int = SEED):
        """Counter-based stimulus straight into the bit-sliced layout (K5):
        word(stream s, column j) = splitmix64(seed ^ (s << 32) ^ (lane_word_offset + j)),
        stream s = running bit index over ``pins``.  Regenerable on any host or GPU."""
        pins = self.input_pins() if pins is None else list(pins)
        slots, streams = [], []
        s = 0
        for p in pins:
            base, width, sl, sh = self._slots_of(p, need_value=False)
            for b in range(width):
                slots.append(sl[b])
                streams.append(s)
                if sh[b] >= 0:
                    slots.append(sh[b])
                    streams.append(s)
                s += 1
        if slots:
            self.rt.stimulus(self.rt.to_device(np.asarray(slots, dtype=np.int32)),
                             self.rt.to_device(np.asarray(streams, dtype=np.int32)), len(slots),
                             self._whole, self.lane_word_offset, seed)
        return s

    def signature(self, pins: Optional[Sequence[str]] = None):
        """(popcount[n_bits], checksum[n_bits]) over all lanes of every bit of ``pins``
        (default: root OUT ports).  Sums - shards combine with an all-reduce SUM (K4)."""
        pins = self.output_pins() if pins is None else list(pins)
        slots = self._flat_slots(pins)
        n = len(slots)
        d_pc, d_cs = self.signature_device(slots)
        if n == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
        return self.rt.to_host_u64(d_pc), self.rt.to_host_u64(d_cs)

    def signature_device(self, slots: Sequence[int]):
        plan = self.signature_plan(slots=slots)
        self.signature_run(plan)
        n = plan["n"]
        return plan["buf"][:n], plan["buf"][plan["half"]:plan["half"] + n]

    def signature_plan(self, pins: Optional[Sequence[str]] = None, slots: Optional[Sequence[int]] = None):
        """Pre-stage the slot list and one device buffer [popcounts | checksums] so that a
        signature per step costs one memset + one kernel (+ one all-reduce when sharded)."""
        if slots is None:
            slots = self._flat_slots(self.output_pins() if pins is None else list(pins))
        n = len(slots)
        half = max(n, 1)
        return {"n": n, "half": half, "buf": self.rt.zeros_words(2 * half),
                "d_slots": self.rt.to_device(np.asarray(list(slots) or [0], dtype=np.int32))}

    def signature_run(self, plan):
        buf, n, half = plan["buf"], plan["n"], plan["half"]
        if n:
            self.rt.signature(plan["d_slots"], n, self._whole, self.lane_word_offset,
                              buf[:half], buf[half:])
        return buf

    # ---- bulk host <-> device I/O of whole pins (bit planes), coalescing adjacent rows --------
    def _row_runs(self, slots: Sequence[int]):
        runs, start, prev = [], None, None
        for k, s in enumerate(slots):
            if start is None:
                start, first_k = s, k
            elif s != prev + 1:
                runs.append((start, prev - start + 1, first_k))
                start, first_k = s, k
            prev = s
        if start is not None:
            runs.append((start, prev - start + 1, first_k))
        return runs

    def io_plan(self, pins: Sequence[str]):
        """Slots of all bits of ``pins`` (in order) + maximal runs of adjacent rows."""
        slots = []
        for p in pins:
            slots.extend(self._slots_of(p, need_value=False)[2])
        if any(s < 0 or s >= self.program.n_persist for s in slots):
            raise PinNotObservable("bulk I/O needs retained pins")
        return {"slots": slots, "runs": self._row_runs(slots), "n": len(slots)}

    def rows_view(self, first_slot, count):
        """Device view [count, words] of adjacent persistent rows (contiguous when
        stride == words)."""
        flat = self._plane[first_slot * self.stride:(first_slot + count) * self.stride]
        return flat.reshape(count, self.stride)[:, :self.words]

    def upload_rows(self, plan, host_rows, non_blocking=True):
        """host_rows: [n, words] int64/uint64 tensor or array (pinned for async)."""
        for first, count, k in plan["runs"]:
            self.rows_view(first, count).copy_(host_rows[k:k + count], non_blocking=non_blocking)

    def download_rows(self, plan, host_rows, non_blocking=True):
        for first, count, k in plan["runs"]:
            host_rows[k:k + count].copy_(self.rows_view(first, count), non_blocking=non_blocking)

    # ---- toggle coverage ---------------------------------------------------------------------------------
    def enable_toggle_counts(self, pins: Sequence[str]):
        """Count, per bit of ``pins``, how many lane bits change from one step to the next."""
        slots = self._flat_slots(pins)
        n = len(slots)
        self._toggle = {
            "pins": list(pins), "n": n,
            "d_slots": self.rt.to_device(np.asarray(slots, dtype=np.int32)),
            "shadow": self.rt.zeros_words(max(1, n * self.stride)),
            "counts": self.rt.zeros_words(max(1, n)),
        }
        # prime the shadow with the current values so the first step counts real toggles
        self._toggle_sample()
        self._toggle["counts"][:] = 0

    def _toggle_sample(self):
        t = self._toggle
        if t["n"]:
            self.rt.toggle_accum(t["d_slots"], t["n"], self._whole, t["shadow"], self.stride, t["counts"])

    def toggle_counts(self) -> np.ndarray:
        if self._toggle is None:
            raise RuntimeError("enable_toggle_counts() first")
        return self.rt.to_host_u64(self._toggle["counts"][: self._toggle["n"]])

    def toggle_counts_device(self):
        return self._toggle["counts"][: self._toggle["n"]]

    # ---- waveform capture ---------------------------------------------------------------------------------
    def capture(self, pins: Sequence[str], steps: int, every: int = 1) -> np.ndarray:
        """Run ``steps`` steps and sample every bit of ``pins`` after each ``every``-th step.
        Samples stay on the device until the end.  Returns uint64[samples, n_bits, words]."""
        slots = self._flat_slots(pins)
        n = len(slots)
        n_samples = steps // every
        d_slots = self.rt.to_device(np.asarray(slots or [0], dtype=np.int32))
        buf = self.rt.zeros_words(max(1, n_samples * n * self.stride))
        k = 0
        for _ in range(n_samples):
            self.run(every)
            if n:
                self.rt.capture_rows(d_slots, n, self._whole, buf, k * n * self.stride, self.stride)
            k += 1
        if steps - n_samples * every:
            self.run(steps - n_samples * every)
        host = self.rt.to_host_u64(buf)[: n_samples * n * self.stride]
        return host.reshape(n_samples, max(n, 1), self.stride)[:, :n, :self.words].copy() if n else \
            np.zeros((n_samples, 0, self.words), dtype=np.uint64)

    def pin_widths(self, pins: Sequence[str]) -> List[int]:
        return [self._slots_of(p)[1] for p in pins]

    # ---- checkpoint / resume -------------------------------------------------------------------------
    def _records_crc(self):
        return int(np.bitwise_xor.reduce(self.program.records.view(np.uint32).ravel().astype(np.uint64))) \
            if self.program.n_ops else 0

    def _canonical_slots(self):
        """int64[n_pin_signals]: for every bit of every net, in the DESIGN's own numbering (the same for every compile
        of one circuit), the persistent slot that holds it in THIS program, or -1.  With shared edge detectors the
        members' ``prevclock`` nets resolve to their leader's slot; a state net kept only through its 'prev'
        shadow (levels mode) resolves to the shadow."""
        prog = self.program
        d = prog.design
        P = prog.n_persist
        widths = np.asarray(d.net_width, dtype=np.int64)
        n_pin = int(widths.sum())
        nets = np.repeat(np.arange(len(widths), dtype=np.int64), widths)
        bits = np.arange(n_pin, dtype=np.int64) - np.asarray(d.net_sig, dtype=np.int64)[nets]
        net_sig = np.asarray(prog.net_sig if prog.net_sig is not None else d.net_sig, dtype=np.int64)
        sig = net_sig[nets] + bits
        slots = np.asarray(prog.sig_slot, dtype=np.int64)[sig]
        for s_, shadow in prog.sig_shadow.items():
            if 0 <= shadow < P:
                hit = (sig == s_) & ~((slots >= 0) & (slots < P))
                slots[hit] = shadow
        slots[(slots < 0) | (slots >= P)] = -1
        return slots, sig

    def state_dict(self):
        """Everything that survives a step: the persistent rows (all lanes), keyed by net bit, and the step count.
        Scratch rows are dead between steps, so they are not saved.  Keyed by net, a checkpoint written under one
        slot layout (levels / resident / serial, with or without shared edge detectors) resumes under another
        engine of the same circuit."""
        prog = self.program
        P = prog.n_persist
        rows = self.rt.to_host_u64(self._plane[:P * self.stride]).reshape(P, self.stride)[:, :self.words]
        state = {"persist": np.ascontiguousarray(rows), "steps_done": self.steps_done, "lanes": self.lanes,
                 "n_persist": P, "records_crc": self._records_crc()}
        if prog.design is not None:
            state["net_bit_slots"] = self._canonical_slots()[0]
        return state

    def load_state_dict(self, state):
        prog = self.program
        P = prog.n_persist
        if state["lanes"] != self.lanes:
            raise ValueError("checkpoint is for another engine geometry")
        full = np.zeros((P, self.stride), dtype=np.uint64)
        if state["n_persist"] == P and state.get("records_crc") == self._records_crc():
            full[:, :self.words] = state["persist"]              # same program: slot for slot
        else:
            # another program of the same circuit: move the rows net bit by net bit
            if "net_bit_slots" not in state or prog.design is None:
                raise ValueError("checkpoint is for another netlist")
            mine, sig = self._canonical_slots()
            theirs = np.asarray(state["net_bit_slots"], dtype=np.int64)
            if len(theirs) != len(mine):
                raise ValueError("checkpoint is for another netlist")
            need = np.nonzero(mine >= 0)[0]
            missing = need[theirs[need] < 0]
            if len(missing):
                raise ValueError("checkpoint lacks %d net bits this engine keeps (e.g. bit %d): it was written with "
                                 "another keep= / observe= selection" % (len(missing), int(missing[0])))
            full[mine[need], :self.words] = state["persist"][theirs[need]]
            for s_, shadow in prog.sig_shadow.items():           # 'prev' shadows hold the end-of-step value
                if 0 <= shadow < P:
                    hit = np.nonzero((sig == s_) & (theirs >= 0))[0]
                    if len(hit):
                        full[shadow, :self.words] = state["persist"][theirs[hit[0]]]
        self._plane[:P * self.stride] = self.rt.to_device(full.reshape(-1))
        self.steps_done = int(state["steps_done"])

    def _unshare_edge_detectors(self):
        """Rebuild this engine WITHOUT shared edge detectors, keeping all state: called when a caller pokes the
        ``prevclock`` pin of one Register / Counter of a sharing group, which the shared bit cannot represent
        (in the reference every primitive has its own ``prevclock`` global, simulator.py:100-127)."""
        if self._toggle is not None:                             # its shadow rows are laid out for the shared program
            raise compiler.CircuitError("toggle counting is on and one member's prevclock of a shared group is being "
                                        "poked: build the engine with share_edge_detectors=False")
        kw = dict(self._ctor_kwargs)
        kw["share_edge_detectors"] = False
        if self.mode == "serial" or kw.get("mode") == "serial":
            kw["mode"] = "serial"
        new = B200Engine(self.root, self.burst_size, self.map_pins, runtime=self.rt, **kw)
        new.load_state_dict(self.state_dict())
        self.close()
        self.__dict__.update(new.__dict__)
        new._d_records = None                                    # its graphs and buffers are ours now

    def regions_skipped(self) -> int:
        """Serial mode: how many guarded regions the first warp jumped over so far (a sample of
        what every warp does when all lanes share the clock)."""
        return int(self.rt.to_host_u64(self._skipped)[0]) if self._skipped is not None else 0

    # ---- accounting ----------------------------------------------------------------------------------------
    def synchronize(self):
        self.rt.synchronize()

    @property
    def gates_per_step(self) -> int:
        return self.program.n_gates

    def gate_evals(self, steps=1) -> int:
        """lanes x gates x steps (BASELINE.json metric)."""
        return self.lanes * self.program.n_gates * int(steps)

    def algorithmic_bytes(self, steps=1) -> int:
        """8 B per fan-in word read + 8 B per output word written, per record."""
        return self.program.gate_words_bytes * self.words * int(steps)
