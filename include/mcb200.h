/*
 * mcb200.h - C ABI of the B200-native gate-level evaluation engine (libmcb200.so).
 *
 * Drop-in boundary for the simulation path of matijakevic/mcircuit.  The reference has no
 * FFI of its own: its Python `JIT` object obtains two raw function pointers from LLVM MCJIT
 * and calls them through ctypes (`CFUNCTYPE(None)`), and peeks/pokes net storage through
 * `get_global_value_address` + a `c_ulonglong` cast.  Each entry point below names the
 * reference interface it replaces as `core/simulator.py:<line>`.
 *
 * Conventions
 *   - plain C types only; every `d_` pointer is a DEVICE pointer owned by the caller (the
 *     Python host keeps them alive as torch tensors); `h_` pointers are host memory;
 *   - the library allocates no persistent device memory; the only host-side state it keeps
 *     is the cache of instantiated CUDA graphs of `mcb_run` mode 1, which remember the
 *     pointers they were captured with (`mcb_graph_cache_clear_owner`);
 *   - every call returns 0 on success or a negative code; `mcb_last_error()` returns the
 *     message of the calling thread's last failure; no C++ exception crosses the boundary;
 *   - all calls are asynchronous enqueues on `stream` (a `cudaStream_t` passed as `void*`,
 *     NULL = the legacy default stream) unless stated otherwise.
 *
 * Data layout: one 64-bit word carries one wire for 64 stimulus lanes.
 * State is slot-major: `word(slot, column)`.  Slots `< n_persist` live in the persistent
 * plane `d_persist[slot * persist_stride + column]`; slots `>= n_persist` live in the
 * scratch plane `d_scratch[(slot - n_persist) * scratch_stride + column]` and are recycled
 * by liveness.  A netlist record is 16 bytes, `uint32 {op | in2 << 8, in0, in1, out}`;
 * records are sorted by level and `level_off[l] .. level_off[l+1]` delimits level `l`.
 */
#ifndef MCB200_H
#define MCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_VERSION 1

/* opcodes (low 7 bits of record word 0); bit 7 complements the result */
enum {
    MCB_OP_NOP = 0,  MCB_OP_BUF = 1,  MCB_OP_AND = 2,  MCB_OP_OR = 3,   MCB_OP_XOR = 4,
    MCB_OP_ANDN = 5, MCB_OP_MUX = 6,  MCB_OP_AND3 = 7, MCB_OP_OR3 = 8,  MCB_OP_XOR3 = 9,
    MCB_OP_MAJ = 10, MCB_OP_NEG = 0x80
};

/* error codes */
enum {
    MCB_OK = 0, MCB_ERR_ARG = -1, MCB_ERR_CUDA = -2, MCB_ERR_NO_DEVICE = -3,
    MCB_ERR_ARCH = -4, MCB_ERR_STATE = -5
};

/* where the lane words of a tile live */
typedef struct mcb_state {
    uint64_t *d_persist;      /* first column of this tile, slot 0                      */
    int64_t   persist_stride; /* words between consecutive persistent slots             */
    uint64_t *d_scratch;      /* first column of this tile's scratch, slot n_persist    */
    int64_t   scratch_stride; /* words between consecutive scratch slots                */
    int32_t   n_persist;      /* slots below this number are persistent                 */
    int32_t   n_slots;        /* n_persist + scratch slots (bounds checking only)       */
    int64_t   n_words;        /* columns (64-lane words) in this tile                   */
} mcb_state;

int         mcb_version(void);
const char *mcb_last_error(void);
/* sm count, compute capability, memory; fails with MCB_ERR_ARCH unless the device is sm_100 */
int mcb_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                    int64_t *total_mem, int64_t *l2_bytes);

/* K1 - levelized evaluation.  Replaces the body of the JIT'd `step()` for every primitive
 * (core/simulator.py:195-223 driving :77-147): evaluates levels [level_begin, level_end)
 * of the netlist on one lane tile, one launch per level.                                 */
int mcb_eval_levels(const uint32_t *d_records, const int32_t *h_level_off,
                    int level_begin, int level_end, const mcb_state *st, void *stream);

/* K2 - fused end-of-step latch.  Replaces the implicit "a fan-in whose driver is emitted
 * later sees the previous step's value" of the reference's in-place globals
 * (core/simulator.py:41-53, :185-191): runs `n` BUF records (cur -> prev shadow).  The
 * compiler also appends these as the last level, so mcb_eval_levels/mcb_run already
 * include them; this entry exists for callers that drive levels themselves.             */
int mcb_latch(const uint32_t *d_latch_records, int n, const mcb_state *st, void *stream);

/* K3 - resident multi-step run.  Replaces `burst()` (core/simulator.py:225-245, :296-297)
 * without its off-by-one: runs exactly `steps` steps; the Python `burst()` passes
 * burst_size + 1.  mode 0: one persistent kernel, each thread owns (a slice of) a lane column
 * and walks the whole record stream, which the host has arranged in batches of B
 * independent records of ONE base opcode (padded with records on a private trash slot; every
 * load of a batch may be issued before its first store); `h_level_off` then only has to
 * delimit segments that are multiples of B.
 * Bits 8..15 of `mode` carry B (0 = MCB_RESIDENT_UNROLL; 4 or 8), bits 16..23 the bytes
 * of a lane word one thread owns (0 = automatic; 8, 4, 2 or 1).  mode 1: per-level launches
 * captured once in a CUDA graph (cached per netlist and tile) and replayed `steps` times; bit
 * 24 of `mode` says the records carry operand hints (bit 31 of in0 / in1 = "produced two or
 * more levels back": that fan-in is loaded with an L2 evict-first policy).                               */
#define MCB_RESIDENT_UNROLL 4
int mcb_run(const uint32_t *d_records, const int32_t *h_level_off, int n_levels,
            const mcb_state *st, int64_t steps, int mode, void *stream);

/* K3 - streamed serial run.  Replaces `step()` / `burst()` (core/simulator.py:195-245, :293-297)
 * in the reference's own execution model: every micro-op in emit order, in place, `steps` times,
 * each consumer thread on its own (slice of a) lane column, ONE launch for any number of steps.
 * `d_stream` (16-byte aligned) holds `n_blocks` control lines of 204 words (mcircuit_b200/serial.py):
 * 4 meta words, 16 TMA runs, 24 segment words, 32 row slots, 32 records of four
 * `field(24) | meta(8) << 24` words.  A producer warp per CTA stages each block's control line
 * (cp.async.bulk) and its state rows (one cp.async.bulk.tensor.2d per run of 2^k consecutive slots;
 * the tensor maps of the state planes are encoded and cached inside the library) into a
 * shared-memory ring, handed to four consumer warps through mbarriers; operands come from the ring,
 * from per-thread stash words in shared memory, from the accumulator (the previous record's
 * result), from the guard register, or - for rows stored too recently to be staged - from an
 * ordinary load; a GUARD segment lets a CTA whose enable word is zero in every lane jump over its
 * region.  State rows and `d_stream` must be 16-byte aligned, row strides multiples of 16 bytes.
 * `word_bytes`: low byte = bytes of a lane word one thread owns (0 = 8; 8, 4, 2 or 1); bits 8.. =
 * rows a ring group must hold (0 = 32, the format's maximum).
 * `d_skipped`: optional device counter, += the guarded regions CTA 0 jumped over.             */
int mcb_run_serial(const uint32_t *d_stream, int n_blocks, const mcb_state *st, int64_t steps,
                   int word_bytes, uint64_t *d_skipped, void *stream);

/* K1g - mcb_eval_levels on the GROUPED stream: per level the records are sorted by base opcode
 * and every opcode group is padded to a multiple of 4 records (padding reads and writes a
 * trash slot), so a thread decodes one opcode per batch and the batch body is branch-free.   */
int mcb_eval_levels_grouped(const uint32_t *d_records, const int32_t *h_level_off,
                            int level_begin, int level_end, const mcb_state *st, void *stream);

/* K1pp - every lane tile of `st` (all `st->n_words` columns, `tile_words` at a time) through
 * every level in ONE cooperative launch, `in_flight` (1..4) tiles at a time with split-phase
 * grid barriers, on the grouped stream with operand streaming hints (bit 31 of in0 / in1 =
 * "producer is two or more levels back: stream it through L2"; slots are the low 31 bits).  `st->d_scratch` is the
 * scratch plane of in-flight slot 0, the others follow at `scratch_plane_words`; `d_sync`
 * is 64 bytes of caller-owned device memory (zeroed by the call): one 64-bit monotonic
 * arrival counter per in-flight slot, so the barrier cannot wrap however many steps one
 * launch runs.                                                                              */
int mcb_run_pipelined(const uint32_t *d_records, const int32_t *d_level_off,
                      const int32_t *h_level_off, int n_levels, const mcb_state *st,
                      int64_t tile_words, int in_flight, int64_t scratch_plane_words,
                      uint64_t *d_sync, int64_t steps, void *stream);

/* K4 - signatures and toggle counts (new capability; no reference counterpart - the
 * reference only peeks one pin at a time, core/simulator.py:285-287).  For each of `n`
 * persistent slots: d_popcount[i] = number of set lane bits, d_checksum[i] = sum over
 * columns j of word * (2 * (col_offset + j) + 1) mod 2^64.  Both are sums, so per-GPU
 * vectors combine with ncclSum.  Columns >= n_words are ignored.  With fewer rows than the
 * GPU needs CTAs a row is cut into column chunks whose partial sums meet by atomicAdd (the
 * result words are zeroed on the stream first); same for the toggle counts.               */
int mcb_signature(const int32_t *d_slots, int n, const mcb_state *st, int64_t col_offset,
                  uint64_t *d_popcount, uint64_t *d_checksum, void *stream);
/* d_toggles[i] += popcount(word ^ shadow); shadow = word, per slot and column            */
int mcb_toggle_accum(const int32_t *d_slots, int n, const mcb_state *st,
                     uint64_t *d_shadow, int64_t shadow_stride, uint64_t *d_toggles,
                     void *stream);

/* K5 - stimulus and packing.  Generalise `set_pin_state` / `get_pin_state`
 * (core/simulator.py:285-291) from one lane to all lanes.                                 */
/* word(slot i, column j) = splitmix64(seed ^ (streams[i] << 32) ^ (col_offset + j))       */
int mcb_stimulus(const int32_t *d_slots, const int32_t *d_streams, int n, const mcb_state *st,
                 int64_t col_offset, uint64_t seed, void *stream);
/* every column of slot i = d_values[i] (0 / ~0 broadcast for scalar pokes, constants)     */
int mcb_fill_rows(const int32_t *d_slots, const uint64_t *d_values, int n,
                  const mcb_state *st, void *stream);
/* values-per-lane -> bit planes: lane l of d_values (n_lanes = 64 * n_words values of
 * `width` bits) goes to bit (l % 64) of column (l / 64) of slots d_slots[0..width)        */
int mcb_pack(const uint64_t *d_values, int width, const int32_t *d_slots,
             const mcb_state *st, void *stream);
int mcb_unpack(uint64_t *d_values, int width, const int32_t *d_slots,
               const mcb_state *st, void *stream);

/* Waveform capture (observability after the path; the reference can only peek one pin of one
 * instance, core/simulator.py:285-287): d_dst[i * dst_stride + j] = word(d_slots[i], j).     */
int mcb_capture_rows(const int32_t *d_slots, int n, const mcb_state *st, uint64_t *d_dst,
                     int64_t dst_stride, void *stream);

/* Strided host<->device copy of `rows` row segments of `width_bytes` (pinned host memory for
 * true asynchrony): the bulk form of set/get_pin_state for whole lane tiles, so a tile's
 * stimulus upload and result download overlap the evaluation of its neighbours.
 * kind 1 = host to device, 2 = device to host.                                           */
int mcb_copy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch,
               int64_t width_bytes, int64_t rows, int kind, void *stream);

/* mcb_run mode 1 keeps one instantiated CUDA graph per (netlist, tile geometry, tuning), at
 * most 8,192 (least recently used quarter evicted beyond that).  An engine that is going away
 * calls mcb_graph_cache_clear_owner(its d_records) before freeing the buffers its graphs point
 * into - other engines' graphs stay; mcb_graph_cache_clear drops everything.  Both return the
 * number of graphs dropped.  Graphs are launched under the cache lock, so a concurrent clear
 * cannot destroy an exec that is about to be launched.                                      */
int mcb_graph_cache_clear(void);
int mcb_graph_cache_clear_owner(const uint32_t *d_records);
int mcb_graph_cache_size(void);

/* tuning knobs; returns the previous value.  Keys: level_block, level_unroll, level_ctas_per_sm,
 * level_cache, level_pdl (programmatic dependent launch between levels), level_tma (1 = K1t: rows
 * staged in shared memory by cp.async.bulk, tiles of <= 512 lane words), resident_*, persist_*,
 * serial_word_bytes, stream_groups (ring groups of K3, 0 = as many as fit, <= 8),
 * stream_smem_kb (shared memory K3 may take per CTA, 0 = the device maximum)                */
int mcb_set_tuning(const char *key, int value);
/* every knob back to the shipped default                                                   */
int mcb_reset_tuning(void);
/* number of kernel launches (graph kernel nodes included) this library has issued in this
 * PROCESS since the last reset - one atomic counter shared by all threads                  */
int64_t mcb_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
