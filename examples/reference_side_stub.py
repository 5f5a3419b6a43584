"""The reference-side binding of INTEGRATION.md section 2 as a runnable file: a minimal
``Executor`` that a maintainer could paste into the reference's ``core/simulator.py`` next to
``class JIT`` (core/simulator.py:174).  It talks to ``libmcb200.so`` through ctypes only; the
netlist compiler (pure Python) comes from ``mcircuit_b200.compiler``.  The full-featured
engine is ``mcircuit_b200/engine.py``; ``tests/test_gpu_parity.py`` runs this stub against
the golden vectors harvested from the reference JIT.
"""
import ctypes
import os

import numpy as np
import torch

from mcircuit_b200.compiler import compile_circuit

_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcircuit_b200", "libmcb200.so")


class _McbState(ctypes.Structure):                           # include/mcb200.h: mcb_state
    _fields_ = [("d_persist", ctypes.c_void_p), ("persist_stride", ctypes.c_int64),
                ("d_scratch", ctypes.c_void_p), ("scratch_stride", ctypes.c_int64),
                ("n_persist", ctypes.c_int32), ("n_slots", ctypes.c_int32),
                ("n_words", ctypes.c_int64)]


class Executor:                                              # core/simulator.py:160-171
    def get_pin_state(self, pin):
        raise NotImplementedError

    def set_pin_state(self, pin, value):
        raise NotImplementedError

    def step(self):
        raise NotImplementedError

    def burst(self):
        raise NotImplementedError


class B200(Executor):                                        # replaces JIT, core/simulator.py:174-297
    _lib = None

    def __init__(self, root, burst_size, map_pins):          # core/simulator.py:175
        if B200._lib is None:
            lib = ctypes.CDLL(_LIB)
            lib.mcb_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_McbState),
                                    ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
            lib.mcb_fill_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.POINTER(_McbState), ctypes.c_void_p]
            lib.mcb_last_error.restype = ctypes.c_char_p
            B200._lib = lib
        self.burst_size = burst_size
        p = self.prog = compile_circuit(root, map_pins)      # replaces core/simulator.py:180-221
        words, stride = 1, 16                                # one 64-lane word, 128-byte rows
        n = p.n_persist + p.n_scratch
        self.plane = torch.zeros(n * stride + 16, dtype=torch.int64, device="cuda")
        self.recs = torch.from_numpy(np.ascontiguousarray(p.records).view(np.int32)).cuda()
        self.off = np.ascontiguousarray(p.level_off, dtype=np.int32)
        base = self.plane.data_ptr()
        self.st = _McbState(base, stride, base + 8 * p.n_persist * stride, stride, p.n_persist, n, words)
        for first, width, value in p.constants:              # core/simulator.py:275-276
            self._poke(first, width, value)

    def _check(self, rc):
        if rc:
            raise RuntimeError(self._lib.mcb_last_error().decode())

    def _poke(self, first_signal, width, value):             # set_pin_state, core/simulator.py:289-291
        slots = self.prog.sig_slot[first_signal:first_signal + width].astype(np.int32)
        vals = np.array([-(value >> b & 1) for b in range(width)], dtype=np.int64)
        d_slots, d_vals = torch.from_numpy(slots).cuda(), torch.from_numpy(vals).cuda()
        self._check(self._lib.mcb_fill_rows(d_slots.data_ptr(), d_vals.data_ptr(), width,
                                            ctypes.byref(self.st), None))
        torch.cuda.synchronize()                             # d_slots / d_vals die with this frame

    def set_pin_state(self, pin, value):
        self._poke(*self.prog.pin_signals(pin), value)       # KeyError on an unknown pin

    def get_pin_state(self, pin):                            # core/simulator.py:285-287
        slots = self.prog.pin_slots(pin)
        words = self.plane[torch.tensor(slots, device="cuda") * self.st.persist_stride].cpu()
        return sum((int(w) & 1) << b for b, w in enumerate(words))

    def _run(self, steps):
        self._check(self._lib.mcb_run(self.recs.data_ptr(), self.off.ctypes.data, self.prog.n_levels,
                                      ctypes.byref(self.st), steps, 1, None))

    def step(self):                                          # core/simulator.py:293-294
        self._run(1)

    def burst(self):                                         # core/simulator.py:296-297 (+1: :233-242)
        self._run(self.burst_size + 1)
