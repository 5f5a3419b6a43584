"""ctypes binding of libmcb200.so (include/mcb200.h) + device buffer ownership.

PyTorch is used for exactly three things: allocating device memory (``torch.zeros`` on a
CUDA device), moving bytes between host and device, and exposing the current stream.  All
compute goes through the C ABI into the hand-written sm_100a kernels.  There is no CPU
fallback: constructing ``CudaRuntime`` without the built library or without a B200 raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_int, c_int32, c_int64, c_uint32, c_uint64,
                    c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcb200.so")

MODE_RESIDENT = 0
MODE_LEVELS = 1
MODE_PERSISTENT = 2
RESIDENT_UNROLL = 4


class McbState(Structure):
    _fields_ = [
        ("d_persist", c_void_p), ("persist_stride", c_int64),
        ("d_scratch", c_void_p), ("scratch_stride", c_int64),
        ("n_persist", c_int32), ("n_slots", c_int32), ("n_words", c_int64),
    ]


class EngineUnavailable(RuntimeError):
    pass


_lib = None

# every symbol include/mcb200.h declares: (name, restype, argtypes)
_PROTOTYPES = [
    ("mcb_version", c_int, []),
    ("mcb_last_error", c_char_p, []),
    ("mcb_device_info", c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                POINTER(c_int64), POINTER(c_int64)]),
    ("mcb_eval_levels", c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(McbState), c_void_p]),
    ("mcb_latch", c_int, [c_void_p, c_int, POINTER(McbState), c_void_p]),
    ("mcb_run", c_int, [c_void_p, c_void_p, c_int, POINTER(McbState), c_int64, c_int, c_void_p]),
    ("mcb_run_serial", c_int, [c_void_p, c_int, POINTER(McbState), c_int64, c_int, c_void_p, c_void_p]),
    ("mcb_eval_levels_grouped", c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(McbState), c_void_p]),
    ("mcb_run_pipelined", c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(McbState), c_int64, c_int,
                                  c_int64, c_void_p, c_int64, c_void_p]),
    ("mcb_signature", c_int, [c_void_p, c_int, POINTER(McbState), c_int64, c_void_p, c_void_p,
                              c_void_p]),
    ("mcb_toggle_accum", c_int, [c_void_p, c_int, POINTER(McbState), c_void_p, c_int64, c_void_p,
                                 c_void_p]),
    ("mcb_stimulus", c_int, [c_void_p, c_void_p, c_int, POINTER(McbState), c_int64, c_uint64,
                             c_void_p]),
    ("mcb_fill_rows", c_int, [c_void_p, c_void_p, c_int, POINTER(McbState), c_void_p]),
    ("mcb_pack", c_int, [c_void_p, c_int, c_void_p, POINTER(McbState), c_void_p]),
    ("mcb_unpack", c_int, [c_void_p, c_int, c_void_p, POINTER(McbState), c_void_p]),
    ("mcb_capture_rows", c_int, [c_void_p, c_int, POINTER(McbState), c_void_p, c_int64, c_void_p]),
    ("mcb_copy2d", c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    ("mcb_graph_cache_clear", c_int, []),
    ("mcb_graph_cache_clear_owner", c_int, [c_void_p]),
    ("mcb_graph_cache_size", c_int, []),
    ("mcb_reset_tuning", c_int, []),
    ("mcb_set_tuning", c_int, [c_char_p, c_int]),
    ("mcb_launch_count", c_int64, [c_int]),
]
EXPORTED_SYMBOLS = [p[0] for p in _PROTOTYPES]


def load_library(path: str = LIB_PATH):
    """dlopen the engine and bind every prototype.  Raises EngineUnavailable when the
    shared object has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(path):
        raise EngineUnavailable(
            "%s not found: build it with __graft_entry__.build() (nvcc, sm_100a). "
            "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, restype, argtypes in _PROTOTYPES:
        fn = getattr(lib, name)       # AttributeError if a declared symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mcb_version() != 1:
        raise EngineUnavailable("libmcb200.so version %d, expected 1" % lib.mcb_version())
    _lib = lib
    return lib


class CudaRuntime:
    """One device + one stream.  Buffers are torch tensors; kernels are ours."""

    name = "cuda"

    def __init__(self, device=None):
        self.lib = load_library()
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise EngineUnavailable("no CUDA device visible; the engine has no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        sms, maj, mnr = c_int(), c_int(), c_int()
        mem, l2 = c_int64(), c_int64()
        with torch.cuda.device(self.device):
            rc = self.lib.mcb_device_info(index, sms, maj, mnr, mem, l2)
        if rc != 0:
            raise EngineUnavailable(self._err())
        self.sm_count, self.cc = sms.value, (maj.value, mnr.value)
        self.total_mem, self.l2_bytes = mem.value, l2.value

    # ---- plumbing ----------------------------------------------------------------------
    def _err(self):
        return (self.lib.mcb_last_error() or b"").decode("utf-8", "replace")

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("libmcb200: %s (code %d)" % (self._err(), rc))

    def _stream(self):
        return c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def free_bytes(self):
        with self.torch.cuda.device(self.device):
            free, _ = self.torch.cuda.mem_get_info()
        return int(free)

    def zeros_words(self, n):
        return self.torch.zeros(int(n), dtype=self.torch.int64, device=self.device)

    def empty_words(self, n):
        return self.torch.empty(int(n), dtype=self.torch.int64, device=self.device)

    def to_device(self, array: np.ndarray):
        a = np.ascontiguousarray(array)
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        elif a.dtype == np.uint64:
            a = a.view(np.int64)
        return self.torch.from_numpy(a).to(self.device)

    def to_host_u64(self, tensor) -> np.ndarray:
        return tensor.detach().cpu().numpy().view(np.uint64)

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    def ptr(self, tensor, word_offset=0):
        return tensor.data_ptr() + 8 * int(word_offset)

    def make_state(self, persist, persist_stride, scratch, scratch_stride, n_persist, n_slots,
                   n_words, col_offset=0, scratch_col_offset=0) -> McbState:
        st = McbState()
        st.d_persist = self.ptr(persist, col_offset) if persist is not None else None
        st.persist_stride = int(persist_stride)
        st.d_scratch = self.ptr(scratch, scratch_col_offset) if scratch is not None else None
        st.scratch_stride = int(scratch_stride)
        st.n_persist = int(n_persist)
        st.n_slots = int(n_slots)
        st.n_words = int(n_words)
        return st

    # ---- kernels -------------------------------------------------------------------------
    def eval_levels(self, d_records, h_level_off: np.ndarray, lo, hi, st: McbState):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_eval_levels(
                d_records.data_ptr(), h_level_off.ctypes.data, int(lo), int(hi),
                ctypes.byref(st), self._stream()))

    def latch(self, d_records, n, st: McbState):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_latch(d_records.data_ptr(), int(n), ctypes.byref(st), self._stream()))

    def run(self, d_records, h_level_off: np.ndarray, n_levels, st: McbState, steps, mode):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_run(
                d_records.data_ptr(), h_level_off.ctypes.data, int(n_levels), ctypes.byref(st),
                int(steps), int(mode), self._stream()))

    def run_serial(self, d_records, n_records, st: McbState, steps, word_bytes=0, d_skipped=None):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_run_serial(
                d_records.data_ptr(), int(n_records), ctypes.byref(st), int(steps), int(word_bytes),
                d_skipped.data_ptr() if d_skipped is not None else None, self._stream()))

    def eval_levels_grouped(self, d_records, h_level_off: np.ndarray, lo, hi, st: McbState):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_eval_levels_grouped(
                d_records.data_ptr(), h_level_off.ctypes.data, int(lo), int(hi), ctypes.byref(st),
                self._stream()))

    def run_pipelined(self, d_records, d_level_off, h_level_off: np.ndarray, n_levels, st: McbState,
                      tile_words, in_flight, scratch_plane_words, d_sync, steps):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_run_pipelined(
                d_records.data_ptr(), d_level_off.data_ptr(), h_level_off.ctypes.data, int(n_levels),
                ctypes.byref(st), int(tile_words), int(in_flight), int(scratch_plane_words),
                d_sync.data_ptr(), int(steps), self._stream()))

    def signature(self, d_slots, n, st: McbState, col_offset, d_popcount, d_checksum):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_signature(
                d_slots.data_ptr(), int(n), ctypes.byref(st), int(col_offset),
                d_popcount.data_ptr() if d_popcount is not None else None,
                d_checksum.data_ptr() if d_checksum is not None else None, self._stream()))

    def toggle_accum(self, d_slots, n, st: McbState, d_shadow, shadow_stride, d_toggles):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_toggle_accum(
                d_slots.data_ptr(), int(n), ctypes.byref(st), d_shadow.data_ptr(),
                int(shadow_stride), d_toggles.data_ptr(), self._stream()))

    def stimulus(self, d_slots, d_streams, n, st: McbState, col_offset, seed):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_stimulus(
                d_slots.data_ptr(), d_streams.data_ptr(), int(n), ctypes.byref(st),
                int(col_offset), c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), self._stream()))

    def fill_rows(self, d_slots, d_values, n, st: McbState):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_fill_rows(
                d_slots.data_ptr(), d_values.data_ptr(), int(n), ctypes.byref(st), self._stream()))

    def pack(self, d_values, width, d_slots, st: McbState):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_pack(
                d_values.data_ptr(), int(width), d_slots.data_ptr(), ctypes.byref(st), self._stream()))

    def unpack(self, d_values, width, d_slots, st: McbState):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_unpack(
                d_values.data_ptr(), int(width), d_slots.data_ptr(), ctypes.byref(st), self._stream()))

    def capture_rows(self, d_slots, n, st: McbState, d_dst, dst_word_offset, dst_stride):
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_capture_rows(d_slots.data_ptr(), int(n), ctypes.byref(st),
                                                  c_void_p(self.ptr(d_dst, dst_word_offset)), int(dst_stride),
                                                  self._stream()))

    def copy2d(self, dst_ptr, dst_pitch, src_ptr, src_pitch, width_bytes, rows, h2d: bool, stream=None):
        st = c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        with self.torch.cuda.device(self.device):
            self._check(self.lib.mcb_copy2d(c_void_p(dst_ptr), int(dst_pitch), c_void_p(src_ptr),
                                            int(src_pitch), int(width_bytes), int(rows),
                                            1 if h2d else 2, st))

    def set_tuning(self, key: str, value: int) -> int:
        return self.lib.mcb_set_tuning(key.encode(), int(value))

    def graph_cache_clear(self, owner=None) -> int:
        """Drop the cached step graphs of one engine (``owner`` = its record tensor) or all."""
        if owner is None:
            return int(self.lib.mcb_graph_cache_clear())
        return int(self.lib.mcb_graph_cache_clear_owner(owner.data_ptr()))

    def graph_cache_size(self) -> int:
        return int(self.lib.mcb_graph_cache_size())

    def reset_tuning(self):
        self.lib.mcb_reset_tuning()

    def launch_count(self, reset=False) -> int:
        return int(self.lib.mcb_launch_count(1 if reset else 0))
