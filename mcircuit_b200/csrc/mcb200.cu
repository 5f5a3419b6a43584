// mcb200.cu - hand-written sm_100a kernels + the C ABI of include/mcb200.h.
//
// Bit-sliced gate evaluation: one 64-bit word = one wire x 64 stimulus lanes, so every
// AND/OR/XOR/NOT/MUX of the reference's step function (core/simulator.py:77-147) is one
// bitwise instruction on a lane word.  State is slot-major (word(slot, column)); a level of
// the netlist is a set of independent records {op|in2<<8, in0, in1, out}.  This is integer,
// HBM-bound work: no tensor cores.  See DESIGN.md for the roofline of each kernel.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#include "../../include/mcb200.h"

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(MCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                               \
    } while (0)

#define LAUNCH_CHECK(name)                                                                 \
    do {                                                                                   \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess)                                                            \
            return fail(MCB_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// ------------------------------------------------------------------------------------------
// tuning knobs
// ------------------------------------------------------------------------------------------
struct Tuning {
    int level_block = 128;      // threads per CTA of the level kernel (128 measured best)
    int level_unroll = 2;       // records in flight per thread (1, 2, 4 or 8); 2 measured best
    int level_ctas_per_sm = 0;  // 0: one record batch per CTA; n: persistent, grid ~ sms * n
    int level_cache = 1;        // 0: default loads, 1: L1::no_allocate, 2: + L2 hints (hinted records!)
    int level_pdl = 1;          // programmatic dependent launch between consecutive levels (+2.5 % measured)
    int resident_block = 128;   // threads per CTA of the resident kernel
    int resident_tma = 1;       // stage the record stream with cp.async.bulk + mbarrier
    int resident_smem = 1;      // keep state in shared memory when it fits
    int resident_word_bytes = 0;  // bytes of a lane column per thread (0: auto, 8, 4, 2, 1)
    int resident_prefetch = 1;  // prefetch the fan-in rows two batches ahead
    int graph_min_steps = 1;    // mode 1: replay a cached CUDA graph when steps >= this (and a real stream)
    int persist_unroll = 2;     // records per thread batch of the persistent level kernel
    int persist_hints = 1;      // honour the operand streaming hints (bit 31)
    int persist_ctas_per_sm = 1;  // 0: as many as fit
    int persist_block = 1024;   // threads per CTA of the persistent level kernel
    int serial_block = 128;     // threads per CTA of the serial kernel (warps are independent)
    int serial_word_bytes = 0;  // bytes of a lane column per thread (0: auto, 8, 4, 2, 1)
    int level_tma = 0;          // 1: the level kernel stages rows in shared memory with TMA bulk copies (K1t)
    int stream_groups = 0;      // ring groups of the stream kernel (0: as many as fit, at most 24)
    int stream_smem_kb = 0;     // shared memory the stream kernel may take per CTA (0: the device maximum)
};
static Tuning g_tune;
static int g_sm_count = 0;

// cached step graphs of mcb_run mode 1 (host-side handles only; they remember pointers)
// A key starts with the record-stream pointer, which is what identifies the owning engine
// (mcb_graph_cache_clear_owner).  Entries are launched UNDER the mutex, so an engine that is
// being dropped on another thread cannot destroy an exec between lookup and launch.
struct GraphEntry { cudaGraphExec_t exec; long long nodes; unsigned long long last_use; };
static std::mutex g_graph_mu;
static std::unordered_map<std::string, GraphEntry> g_graphs;
static unsigned long long g_graph_tick = 0;
static const size_t GRAPH_CACHE_CAP = 8192;
static void clear_graphs_locked() {
    for (auto &kv : g_graphs) cudaGraphExecDestroy(kv.second.exec);
    g_graphs.clear();
}
static void evict_graphs_locked() {
    // drop the least recently used quarter (never everything: live engines keep their graphs)
    std::vector<unsigned long long> ticks;
    ticks.reserve(g_graphs.size());
    for (auto &kv : g_graphs) ticks.push_back(kv.second.last_use);
    std::nth_element(ticks.begin(), ticks.begin() + ticks.size() / 4, ticks.end());
    const unsigned long long cut = ticks[ticks.size() / 4];
    for (auto it = g_graphs.begin(); it != g_graphs.end();) {
        if (it->second.last_use <= cut) { cudaGraphExecDestroy(it->second.exec); it = g_graphs.erase(it); }
        else ++it;
    }
}

static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        g_sm_count = p.multiProcessorCount;
    }
    return g_sm_count;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
struct View {
    uint64_t *persist;
    long long pstride;
    uint64_t *scratch_biased;   // scratch - n_persist * sstride, so both planes index by slot
    long long sstride;
    uint32_t n_persist;
    long long n_words;
};

static View make_view(const mcb_state *st) {
    View v;
    v.persist = st->d_persist;
    v.pstride = st->persist_stride;
    v.sstride = st->scratch_stride;
    v.scratch_biased = st->d_scratch ? st->d_scratch - (long long)st->n_persist * st->scratch_stride
                                     : nullptr;
    v.n_persist = (uint32_t)st->n_persist;
    v.n_words = st->n_words;
    return v;
}

__device__ __forceinline__ uint64_t *row(const View &v, uint32_t slot) {
    return slot < v.n_persist ? v.persist + (long long)slot * v.pstride
                              : v.scratch_biased + (long long)slot * v.sstride;
}

__device__ __forceinline__ uint64_t apply_op(uint32_t op, uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    switch (op & 0x7f) {
        case MCB_OP_BUF:  r = a; break;
        case MCB_OP_AND:  r = a & b; break;
        case MCB_OP_OR:   r = a | b; break;
        case MCB_OP_XOR:  r = a ^ b; break;
        case MCB_OP_ANDN: r = a & ~b; break;
        case MCB_OP_MUX:  r = (a & b) | (~a & c); break;
        case MCB_OP_AND3: r = a & b & c; break;
        case MCB_OP_OR3:  r = a | b | c; break;
        case MCB_OP_XOR3: r = a ^ b ^ c; break;
        case MCB_OP_MAJ:  r = (a & b) | (a & c) | (b & c); break;
        default:          r = 0; break;
    }
    return (op & MCB_OP_NEG) ? ~r : r;
}

template <int OP, typename W>
__device__ __forceinline__ W op_eval(W a, W b, W c) {
    if (OP == MCB_OP_BUF) return a;
    if (OP == MCB_OP_AND) return a & b;
    if (OP == MCB_OP_OR) return a | b;
    if (OP == MCB_OP_XOR) return a ^ b;
    if (OP == MCB_OP_ANDN) return a & (W)~b;
    if (OP == MCB_OP_MUX) return (a & b) | ((W)~a & c);
    if (OP == MCB_OP_AND3) return a & b & c;
    if (OP == MCB_OP_OR3) return a | b | c;
    if (OP == MCB_OP_XOR3) return a ^ b ^ c;
    return (a & b) | (a & c) | (b & c);                  // MCB_OP_MAJ
}

template <int CACHE>
__device__ __forceinline__ ulonglong2 ld_state2(const uint64_t *p) {
    ulonglong2 r;
    if (CACHE == 1) {
        asm volatile("ld.global.L1::no_allocate.v2.u64 {%0, %1}, [%2];"
                     : "=l"(r.x), "=l"(r.y) : "l"(p));
    } else {
        asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    }
    return r;
}

__device__ __forceinline__ void st_state2(uint64_t *p, ulonglong2 v) {
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(v.x), "l"(v.y) : "memory");
}

__device__ __forceinline__ uint64_t make_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ ulonglong2 ld_state2_hint(const uint64_t *p, uint64_t policy) {
    ulonglong2 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                 : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void st_state2_hint(uint64_t *p, ulonglong2 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;"
                 :: "l"(p), "l"(v.x), "l"(v.y), "l"(policy) : "memory");
}

#define SLOT_MASK 0x7fffffffu

// ------------------------------------------------------------------------------------------
// K1: one level.  Thread = one 16-byte column pair; grid.x covers the tile's columns,
// grid.y strides over the level's records U at a time.  All 2U (3U) row loads of a batch are
// issued before the first use, so each thread keeps U*32..48 bytes in flight; a warp reads
// 512 contiguous bytes of each fan-in row and the record itself is one broadcast 16-byte
// load (uniform across the CTA).
// ------------------------------------------------------------------------------------------
template <int U, int CACHE, int PDL>
__global__ void __launch_bounds__(256)
k_eval_level(const uint4 *__restrict__ recs, int n_recs, View v, long long n_vec) {
    // PDL: launched with programmatic stream serialization, so this grid may start while the
    // previous level is still draining.  Let the NEXT level do the same right away, fetch what
    // does not depend on the previous level (the records), then wait for its rows.
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (col >= n_vec) return;
    const long long off = col * 2;
    // CACHE == 2: the record stream carries operand hints (bit 31 of in0 / in1 = produced two or
    // more levels back): such fan-in streams through L2 with evict-first so that it does not
    // push out the previous level's rows, which the next level is about to read
    uint64_t pol_stream = 0, pol_keep = 0;
    if (CACHE == 2) { pol_stream = make_policy_evict_first(); pol_keep = make_policy_evict_last(); }
    for (int g0 = blockIdx.y * U; g0 < n_recs; g0 += gridDim.y * U) {
        uint4 r[U];
        ulonglong2 a[U], b[U], c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            r[u] = (g0 + u < n_recs) ? __ldg(recs + g0 + u) : make_uint4(0, 0, 0, 0);
        }
        if (PDL && g0 == (int)(blockIdx.y * U)) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t op = r[u].x & 0x7f;
            a[u] = b[u] = c[u] = make_ulonglong2(0, 0);
            if (op != MCB_OP_NOP) {
                if (CACHE == 2) {
                    a[u] = ld_state2_hint(row(v, r[u].y & SLOT_MASK) + off, (r[u].y >> 31) ? pol_stream : pol_keep);
                    if (op != MCB_OP_BUF)
                        b[u] = ld_state2_hint(row(v, r[u].z & SLOT_MASK) + off, (r[u].z >> 31) ? pol_stream : pol_keep);
                    if (op >= MCB_OP_MUX) c[u] = ld_state2<1>(row(v, r[u].x >> 8) + off);
                    r[u].y &= SLOT_MASK;
                    r[u].z &= SLOT_MASK;
                } else {
                    a[u] = ld_state2<CACHE>(row(v, r[u].y) + off);
                    if (op != MCB_OP_BUF) b[u] = ld_state2<CACHE>(row(v, r[u].z) + off);
                    if (op >= MCB_OP_MUX) c[u] = ld_state2<CACHE>(row(v, r[u].x >> 8) + off);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t op = r[u].x & 0xff;
            if ((op & 0x7f) != MCB_OP_NOP) {
                ulonglong2 o;
                o.x = apply_op(op, a[u].x, b[u].x, c[u].x);
                o.y = apply_op(op, a[u].y, b[u].y, c[u].y);
                st_state2(row(v, r[u].w) + off, o);
            }
        }
    }
}

template <int U, int CACHE>
static void launch_level_t(const uint4 *recs, int n, const View &v, cudaStream_t s) {
    const long long n_vec = (v.n_words + 1) / 2;
    int block = g_tune.level_block;
    if (block > 256) block = 256;
    if (block < 32) block = 32;
    while (block > 32 && block / 2 >= n_vec) block /= 2;
    const long long gx = (n_vec + block - 1) / block;
    const long long max_gy = (n + U - 1) / U;
    long long gy = max_gy;
    if (g_tune.level_ctas_per_sm > 0) {
        const long long want = (long long)sm_count() * g_tune.level_ctas_per_sm * (256 / block);
        gy = (want + gx - 1) / gx;
        if (gy > max_gy) gy = max_gy;
    }
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy, 1);
    if (g_tune.level_pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_eval_level<U, CACHE, 1>, recs, n, v, n_vec);
    } else {
        k_eval_level<U, CACHE, 0><<<grid, block, 0, s>>>(recs, n, v, n_vec);
    }
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------
// K1t: one level with the per-level signal state staged in shared memory by TMA bulk copies.
// CTA = 1 producer warp + 4 consumer warps, persistent over the level's records (record i of CTA c
// = c + i * gridDim.x).  Producer: per record, the 16-byte record goes into the stage header and
// the first two fan-in rows of the lane tile (<= 4 KB each) are one cp.async.bulk each into the
// stage, completion on the stage's `full` mbarrier.  Consumers: 128 threads x 32 bytes cover a row;
// LOP3 on four words out of shared memory, the result straight to global memory.  Twelve stages,
// two CTAs per SM: ~190 KB of rows in flight per SM without a single LDG on the record -> address
// -> row chain that stalls K1.  Measured 1.7 x SLOWER than K1 on C5 (profiles/r02_k1t_vs_k1.md):
// opt-in (`level_tma` tuning).
// ------------------------------------------------------------------------------------------
#define LT_STAGES 12
#define LT_ROW_BYTES 4096
#define LT_HEAD_BYTES 1024
struct LtSmem {
    uint64_t full[LT_STAGES];
    uint64_t empty[LT_STAGES];
    uint4 hdr[LT_STAGES];
};

template <int PDL>
__global__ void __launch_bounds__(160, 2)
k_eval_level_tma(const uint4 *__restrict__ recs, int n_recs, View v, uint32_t row_bytes) {
    extern __shared__ __align__(128) unsigned char lt_smem[];
    LtSmem *sh = reinterpret_cast<LtSmem *>(lt_smem);
    unsigned char *in_base = lt_smem + LT_HEAD_BYTES;
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < LT_STAGES; ++s) {
            mbar_init(&sh->full[s], 1);
            mbar_init(&sh->empty[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, stride = gridDim.x;
    if (warp == 4) {
        // ---- producer: the two first fan-in rows of every record, one bulk copy each ----
        uint32_t s = 0, ph = 1;
        bool waited = !PDL;
        for (int i = first; i < n_recs; i += stride) {
            const uint4 r = __ldg(recs + i);
            if (!waited) {                              // the rows (not the records) depend on the previous level
                asm volatile("griddepcontrol.wait;" ::: "memory");
                waited = true;
            }
            mbar_wait(&sh->empty[s], ph);
            if (lane == 0) {
                const uint32_t op = r.x & 0x7fu;
                const uint32_t n_rows = op == MCB_OP_NOP ? 0u : (op == MCB_OP_BUF ? 1u : 2u);
                sh->hdr[s] = r;
                mbar_expect_tx(&sh->full[s], n_rows * row_bytes);
                unsigned char *dst = in_base + s * 2 * LT_ROW_BYTES;
                if (n_rows >= 1) bulk_g2s(dst, row(v, r.y), row_bytes, &sh->full[s]);
                if (n_rows >= 2) bulk_g2s(dst + LT_ROW_BYTES, row(v, r.z), row_bytes, &sh->full[s]);
            }
            if (++s == LT_STAGES) { s = 0; ph ^= 1u; }
        }
        if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
        return;
    }
    // ---- consumers: LOP3 out of shared memory, the result straight to global memory (32 bytes per thread) ----
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");      // before the first store of this level
    const uint32_t my = tid * 32u;
    const bool has0 = my < row_bytes, has1 = my + 16u < row_bytes;
    uint32_t s = 0, ph = 0;
    for (int i = first; i < n_recs; i += stride) {
        mbar_wait(&sh->full[s], ph);
        const uint4 r = sh->hdr[s];
        const uint32_t op = r.x & 0xffu, base = op & 0x7fu;
        const unsigned char *in = in_base + s * 2 * LT_ROW_BYTES + my;
        if (base != MCB_OP_NOP && has0) {
            ulonglong2 a0 = *reinterpret_cast<const ulonglong2 *>(in), a1 = make_ulonglong2(0, 0);
            ulonglong2 b0 = make_ulonglong2(0, 0), b1 = b0, c0 = b0, c1 = b0;
            if (has1) a1 = *reinterpret_cast<const ulonglong2 *>(in + 16);
            if (base != MCB_OP_BUF) {
                b0 = *reinterpret_cast<const ulonglong2 *>(in + LT_ROW_BYTES);
                if (has1) b1 = *reinterpret_cast<const ulonglong2 *>(in + LT_ROW_BYTES + 16);
            }
            if (base >= MCB_OP_MUX) {                   // the third operand of a 3-input micro-op: an ordinary load
                const unsigned char *cr = reinterpret_cast<const unsigned char *>(row(v, r.x >> 8)) + my;
                c0 = ld_state2<1>(reinterpret_cast<const uint64_t *>(cr));
                if (has1) c1 = ld_state2<1>(reinterpret_cast<const uint64_t *>(cr + 16));
            }
            ulonglong2 o0, o1;
            o0.x = apply_op(op, a0.x, b0.x, c0.x);
            o0.y = apply_op(op, a0.y, b0.y, c0.y);
            o1.x = apply_op(op, a1.x, b1.x, c1.x);
            o1.y = apply_op(op, a1.y, b1.y, c1.y);
            unsigned char *out = reinterpret_cast<unsigned char *>(row(v, r.w)) + my;
            st_state2(reinterpret_cast<uint64_t *>(out), o0);
            if (has1) st_state2(reinterpret_cast<uint64_t *>(out + 16), o1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->empty[s]);      // this warp is done with the stage's fan-in rows
        if (++s == LT_STAGES) { s = 0; ph ^= 1u; }
    }
}

static int launch_level_tma(const uint4 *recs, int n, const View &v, cudaStream_t s) {
    const uint32_t row_bytes = (uint32_t)((v.n_words + 1) / 2 * 2 * 8);
    const size_t smem = LT_HEAD_BYTES + (size_t)LT_STAGES * 2 * LT_ROW_BYTES;
    int grid = sm_count() * 2;
    if (grid > n) grid = n;
    static std::atomic<bool> attr_set[2][64];                // per device
    const int pdl = g_tune.level_pdl ? 1 : 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!attr_set[pdl][dev & 63].load(std::memory_order_acquire)) {
        if (pdl) CUDA_TRY(cudaFuncSetAttribute(k_eval_level_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CUDA_TRY(cudaFuncSetAttribute(k_eval_level_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[pdl][dev & 63].store(true, std::memory_order_release);
    }
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(160);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_eval_level_tma<1>, recs, n, v, row_bytes);
    } else {
        k_eval_level_tma<0><<<grid, 160, smem, s>>>(recs, n, v, row_bytes);
    }
    LAUNCH_CHECK("k_eval_level_tma");
    return 0;
}

static int launch_level(const uint4 *recs, int n, const View &v, cudaStream_t s, bool hinted = false) {
    if (n <= 0) return 0;
    // K1t: rows staged through shared memory by TMA bulk copies (tiles of up to 512 lane words: one 4 KB copy per row)
    if (g_tune.level_tma && !hinted && v.n_words * 8 <= LT_ROW_BYTES && ((v.pstride * 8) & 15) == 0 &&
        (!v.scratch_biased || ((v.sstride * 8) & 15) == 0) && (reinterpret_cast<uintptr_t>(v.persist) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(v.scratch_biased) & 15) == 0)
        return launch_level_tma(recs, n, v, s);
    const int U = g_tune.level_unroll;
    // hinted records MUST go through the variant that masks bit 31 of the operands
    const int C = hinted ? 2 : (g_tune.level_cache >= 2 ? 1 : g_tune.level_cache);
#define MCB_LL(UU)                                                                         \
    do {                                                                                   \
        if (C >= 2) launch_level_t<UU, 2>(recs, n, v, s);                                  \
        else if (C == 1) launch_level_t<UU, 1>(recs, n, v, s);                             \
        else launch_level_t<UU, 0>(recs, n, v, s);                                         \
    } while (0)
    if (U >= 8) MCB_LL(8);
    else if (U >= 4) MCB_LL(4);
    else if (U >= 2) MCB_LL(2);
    else MCB_LL(1);
#undef MCB_LL
    LAUNCH_CHECK("k_eval_level");
    return 0;
}

// ------------------------------------------------------------------------------------------
// L2-resident small tiles (used by K1pp below): all levels of a lane tile inside one launch.
// With a small tile (<= 512 words) one level's output rows (5,000 x 4 KB = 20 MB for C5) are
// still in the 126 MB L2 when the next level reads them as fan-in, and recycled scratch
// slots are rewritten before their dirty lines are ever evicted, so DRAM traffic drops below
// the algorithmic 24 B per gate-word.  Per-level kernel launches cannot exploit this (a 10 us
// level pays 3 us of launch gap); a grid barrier costs about 1 us.  Operand bit 31 (set by
// the compiler when the producer is two or more levels back) selects an L2 evict-first
// policy for the streaming fan-in so it does not push the recent rows out.  Loads never
// allocate in L1 (it is not coherent across SMs within a launch).
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// K1g / K1pp: level kernels on the GROUPED record stream (per level, records sorted by base
// opcode and every opcode group padded to a multiple of 4 with records on the trash slot), so
// a thread decodes ONE opcode for its whole batch and the batch body is branch-free.
//
// K1pp ("pipelined persistent"): one cooperative launch walks ALL lane tiles, G tiles in flight.
// Every (tile slot g, level) phase ends with a split-phase grid barrier - arrive now, wait just
// before the same slot's next level - so while slot g's barrier settles the CTA is already
// streaming slot g+1's level: no drain, no launch gap, and tiles small enough (256-512 lane
// words) that a level's rows are still in L2 when the next level reads them.
// ------------------------------------------------------------------------------------------
template <int OP, int U, int HINTS>
__device__ __forceinline__ void level_batch_op(const uint4 (&r)[U], const View &v, long long off,
                                               uint64_t pol_stream, uint64_t pol_keep) {
    constexpr int ARITY = (OP == MCB_OP_BUF) ? 1 : (OP >= MCB_OP_MUX ? 3 : 2);
    ulonglong2 x[U], y[U], z[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        y[u] = z[u] = make_ulonglong2(0, 0);
        if (HINTS) {
            x[u] = ld_state2_hint(row(v, r[u].y & SLOT_MASK) + off, (r[u].y >> 31) ? pol_stream : pol_keep);
            if (ARITY >= 2)
                y[u] = ld_state2_hint(row(v, r[u].z & SLOT_MASK) + off, (r[u].z >> 31) ? pol_stream : pol_keep);
        } else {
            x[u] = ld_state2<1>(row(v, r[u].y & SLOT_MASK) + off);
            if (ARITY >= 2) y[u] = ld_state2<1>(row(v, r[u].z & SLOT_MASK) + off);
        }
        if (ARITY >= 3) z[u] = ld_state2<1>(row(v, r[u].x >> 8) + off);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t neg = 0ull - (uint64_t)((r[u].x >> 7) & 1u);
        ulonglong2 o;
        o.x = op_eval<OP, uint64_t>(x[u].x, y[u].x, z[u].x) ^ neg;
        o.y = op_eval<OP, uint64_t>(x[u].y, y[u].y, z[u].y) ^ neg;
        if (HINTS) st_state2_hint(row(v, r[u].w & SLOT_MASK) + off, o, pol_keep);
        else st_state2(row(v, r[u].w & SLOT_MASK) + off, o);
    }
}

template <int U, int HINTS>
__device__ __forceinline__ void level_batch(const uint4 *__restrict__ recs, const View &v, long long off,
                                            uint64_t pol_stream, uint64_t pol_keep) {
    uint4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) r[u] = __ldg(recs + u);
    switch (r[0].x & 0x7f) {
        case MCB_OP_BUF:  level_batch_op<MCB_OP_BUF, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_AND:  level_batch_op<MCB_OP_AND, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_OR:   level_batch_op<MCB_OP_OR, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_XOR:  level_batch_op<MCB_OP_XOR, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_ANDN: level_batch_op<MCB_OP_ANDN, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_MUX:  level_batch_op<MCB_OP_MUX, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_AND3: level_batch_op<MCB_OP_AND3, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_OR3:  level_batch_op<MCB_OP_OR3, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_XOR3: level_batch_op<MCB_OP_XOR3, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        case MCB_OP_MAJ:  level_batch_op<MCB_OP_MAJ, U, HINTS>(r, v, off, pol_stream, pol_keep); break;
        default: break;
    }
}

// K1g: one level, grouped stream; same thread mapping as k_eval_level
template <int U>
__global__ void __launch_bounds__(256)
k_eval_level_g(const uint4 *__restrict__ recs, int n_recs, View v, long long n_vec) {
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (col >= n_vec) return;
    for (int g0 = blockIdx.y * U; g0 < n_recs; g0 += gridDim.y * U)
        level_batch<U, 0>(recs + g0, v, col * 2, 0, 0);
}

struct PipeArgs {
    uint64_t *persist;          // column 0 of tile 0
    long long pstride;
    uint64_t *scratch;          // plane of in-flight slot 0
    long long sstride;          // words between scratch rows
    long long splane;           // words between the scratch planes of in-flight slots
    uint32_t n_persist;
    long long n_words;          // all columns
    int tile_words;
    int n_tiles;
    int n_levels;
    long long steps;
    unsigned long long *sync;   // G monotonic 64-bit arrival counters, zeroed by the host (never wrap)
};

template <int U, int G, int HINTS>
__global__ void __launch_bounds__(1024, 1)
k_eval_pipelined(const uint4 *__restrict__ recs, const int *__restrict__ level_off, PipeArgs a) {
    const uint64_t pol_stream = make_policy_evict_first();
    const uint64_t pol_keep = make_policy_evict_last();
    // 64-bit on purpose: arrivals x n_cta passes 2^32 after a few hundred steps of a C5-sized run
    unsigned long long arrivals[G];
#pragma unroll
    for (int g = 0; g < G; ++g) arrivals[g] = 0;
    const unsigned int n_cta = gridDim.x;
    for (int t0 = 0; t0 < a.n_tiles; t0 += G) {
        for (long long step = 0; step < a.steps; ++step) {
            for (int lv = 0; lv < a.n_levels; ++lv) {
                const int lo = level_off[lv];
                const int batches = (level_off[lv + 1] - lo) / U;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int tile = t0 + g;
                    if (tile >= a.n_tiles) continue;
                    const long long c0 = (long long)tile * a.tile_words;
                    long long ncols = a.n_words - c0;
                    if (ncols > a.tile_words) ncols = a.tile_words;
                    const unsigned nv = (unsigned)((ncols + 1) / 2);
                    View v;
                    v.persist = a.persist + c0;
                    v.pstride = a.pstride;
                    v.sstride = a.sstride;
                    v.scratch_biased = a.scratch + (long long)g * a.splane - (long long)a.n_persist * a.sstride;
                    v.n_persist = a.n_persist;
                    v.n_words = ncols;
                    // wait: every CTA has finished this slot's previous phase
                    if (threadIdx.x == 0) {
                        const unsigned long long target = arrivals[g] * (unsigned long long)n_cta;
                        volatile unsigned long long *c = a.sync + g;
                        while (*c < target) { }
                        __threadfence();
                    }
                    __syncthreads();
                    // my contiguous share of the phase's (batch, column pair) items
                    const unsigned long long work = (unsigned long long)batches * nv;
                    const unsigned long long w_lo = work * blockIdx.x / n_cta;
                    const unsigned long long w_hi = work * (blockIdx.x + 1) / n_cta;
                    for (unsigned long long w = w_lo + threadIdx.x; w < w_hi; w += blockDim.x) {
                        const unsigned b = (unsigned)(w / nv);
                        const long long off = (long long)(w - (unsigned long long)b * nv) * 2;
                        level_batch<U, HINTS>(recs + lo + (int)b * U, v, off, pol_stream, pol_keep);
                    }
                    // arrive: publish this CTA's rows of the phase
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        __threadfence();
                        atomicAdd(a.sync + g, 1ull);
                    }
                    arrivals[g] += 1;
                }
            }
        }
    }
}

template <int U, int G>
static int launch_pipelined_ug(const uint4 *recs, const int *d_level_off, PipeArgs a, cudaStream_t s) {
    int block = g_tune.persist_block;
    if (block > 1024) block = 1024;
    if (block < 32) block = 32;
    block = block / 32 * 32;
    const int H = g_tune.persist_hints ? 1 : 0;
    const void *fn = H ? (const void *)k_eval_pipelined<U, G, 1> : (const void *)k_eval_pipelined<U, G, 0>;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, 0));
    if (per_sm < 1) return fail(MCB_ERR_CUDA, "pipelined kernel does not fit on an SM");
    if (g_tune.persist_ctas_per_sm > 0 && per_sm > g_tune.persist_ctas_per_sm) per_sm = g_tune.persist_ctas_per_sm;
    const long long grid = (long long)sm_count() * per_sm;
    CUDA_TRY(cudaMemsetAsync(a.sync, 0, 64, s));
    void *args[] = {(void *)&recs, (void *)&d_level_off, (void *)&a};
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(block), args, 0, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ------------------------------------------------------------------------------------------
// K3 (mode 0): resident run.  Lanes never interact, so a thread that owns a lane column can
// walk the whole record stream for many steps with no barrier at all: "level" only bounds
// which records may be batched (the host pads every level to a multiple of RU records, so a
// batch never straddles a level and its loads may all be issued before its stores).  The
// record stream - the same for every thread - is staged through shared memory in chunks by
// one TMA bulk copy per chunk (cp.async.bulk + mbarrier, double buffered); state words are
// either in global memory (L1/L2-resident across steps when small) or, when the whole slot
// file of the CTA's columns fits, in shared memory for the entire run.
// ------------------------------------------------------------------------------------------
#define REC_CHUNK 512                      // records per staged chunk (8 KB)

// The word a thread owns is W = 8, 4, 2 or 1 bytes of a lane column: lanes never interact, so
// a 64-lane word can be split among up to 8 threads.  Deep sequential circuits are latency
// bound (one L2/HBM round trip per dependent batch), and narrower words multiply the number
// of independent threads without adding traffic; wide runs use W = 8.
// UNIFIED: scratch rows follow the persistent rows with the same stride (single-tile runs), so
// a row address is one 32x32+64 multiply-add on a per-thread base that already includes the
// thread's column; otherwise the plane is picked by comparing with n_persist.
template <typename W> __device__ __forceinline__ W ld_global_w(const W *p);
template <> __device__ __forceinline__ uint64_t ld_global_w<uint64_t>(const uint64_t *p) {
    uint64_t x; asm volatile("ld.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory"); return x;
}
template <> __device__ __forceinline__ uint32_t ld_global_w<uint32_t>(const uint32_t *p) {
    uint32_t x; asm volatile("ld.global.u32 %0, [%1];" : "=r"(x) : "l"(p) : "memory"); return x;
}
template <> __device__ __forceinline__ uint16_t ld_global_w<uint16_t>(const uint16_t *p) {
    uint16_t x; asm volatile("ld.global.u16 %0, [%1];" : "=h"(x) : "l"(p) : "memory"); return x;
}
template <> __device__ __forceinline__ uint8_t ld_global_w<uint8_t>(const uint8_t *p) {
    uint32_t x; asm volatile("ld.global.u8 %0, [%1];" : "=r"(x) : "l"(p) : "memory"); return (uint8_t)x;
}
template <typename W> __device__ __forceinline__ void st_global_w(W *p, W x);
template <> __device__ __forceinline__ void st_global_w<uint64_t>(uint64_t *p, uint64_t x) {
    asm volatile("st.global.u64 [%0], %1;" :: "l"(p), "l"(x) : "memory");
}
template <> __device__ __forceinline__ void st_global_w<uint32_t>(uint32_t *p, uint32_t x) {
    asm volatile("st.global.u32 [%0], %1;" :: "l"(p), "r"(x) : "memory");
}
template <> __device__ __forceinline__ void st_global_w<uint16_t>(uint16_t *p, uint16_t x) {
    asm volatile("st.global.u16 [%0], %1;" :: "l"(p), "h"(x) : "memory");
}
template <> __device__ __forceinline__ void st_global_w<uint8_t>(uint8_t *p, uint8_t x) {
    asm volatile("st.global.u8 [%0], %1;" :: "l"(p), "r"((uint32_t)x) : "memory");
}

template <typename W, bool UNIFIED>
struct GlobalCols {
    char *base_p;                           // persist plane + this thread's column
    char *base_s;                           // biased scratch plane + this thread's column
    uint32_t stride_p, stride_s;            // bytes between rows
    uint32_t n_persist;
    __device__ __forceinline__ GlobalCols(const View &v, long long col) {
        base_p = reinterpret_cast<char *>(v.persist) + col * (long long)sizeof(W);
        base_s = reinterpret_cast<char *>(v.scratch_biased) + col * (long long)sizeof(W);
        stride_p = (uint32_t)(v.pstride * 8);
        stride_s = (uint32_t)(v.sstride * 8);
        n_persist = v.n_persist;
        // opaque to the optimiser: otherwise it rematerialises "kernel parameter + column offset" at every use
        // (5-7 instructions per row address instead of one IMAD.WIDE on a base register)
        asm volatile("" : "+l"(base_p), "+l"(base_s), "+r"(stride_p), "+r"(stride_s));
    }
    // every row of this thread becomes one private dummy word (threads beyond the last lane column: their loads and
    // stores then need no predicate at all)
    __device__ __forceinline__ void park(void *dummy) {
        base_p = base_s = reinterpret_cast<char *>(dummy);
        stride_p = stride_s = 0;
    }
    __device__ __forceinline__ W *at(uint32_t slot) const {
        if (UNIFIED) return reinterpret_cast<W *>(base_p + (unsigned long long)slot * stride_p);
        return reinterpret_cast<W *>(slot < n_persist ? base_p + (unsigned long long)slot * stride_p
                                                      : base_s + (unsigned long long)slot * stride_s);
    }
    // explicit state space: the laundered base is a generic pointer to the compiler
    __device__ __forceinline__ W ld(uint32_t slot) const { return ld_global_w<W>(at(slot)); }
    __device__ __forceinline__ void st(uint32_t slot, W x) const { st_global_w<W>(at(slot), x); }
    // non-binding hint: safe even when an earlier batch still has to store to that row (a
    // thread's own store is ordered before its later load of the same address)
    __device__ __forceinline__ void prefetch(uint32_t slot) const {
        asm volatile("prefetch.global.L1 [%0];" :: "l"(at(slot)));
    }
};

// state in shared memory: word(slot, thread) at smem[slot * blockDim.x + threadIdx.x]
template <typename W>
struct SmemCols {
    W *base;
    int stride;
    __device__ __forceinline__ W ld(uint32_t slot) const { return base[slot * stride]; }
    __device__ __forceinline__ void st(uint32_t slot, W x) const { base[slot * stride] = x; }
    __device__ __forceinline__ void prefetch(uint32_t) const {}
};

template <typename W>
__device__ __forceinline__ W apply_op_w(uint32_t op, W a, W b, W c) {
    W r;
    switch (op & 0x7f) {
        case MCB_OP_BUF:  r = a; break;
        case MCB_OP_AND:  r = a & b; break;
        case MCB_OP_OR:   r = a | b; break;
        case MCB_OP_XOR:  r = a ^ b; break;
        case MCB_OP_ANDN: r = a & (W)~b; break;
        case MCB_OP_MUX:  r = (a & b) | ((W)~a & c); break;
        case MCB_OP_AND3: r = a & b & c; break;
        case MCB_OP_OR3:  r = a | b | c; break;
        case MCB_OP_XOR3: r = a ^ b ^ c; break;
        case MCB_OP_MAJ:  r = (a & b) | (a & c) | (b & c); break;
        default:          r = 0; break;
    }
    return (op & MCB_OP_NEG) ? (W)~r : r;
}

// One batch = U records of the SAME base opcode that the host proved independent (one level
// of the netlist, grouped by opcode, padded with records that read and write a private trash
// slot).  The opcode is decoded once per batch; inside, the U load / compute / store chains
// are straight-line and branch-free, so the compiler interleaves them and a single warp keeps
// U x 2..3 loads in flight.  Every load of a batch is issued before its first store.
template <int OP, typename W, int U, class Cols>
__device__ __forceinline__ void run_batch_op(const uint4 *r4, const Cols &cols) {
    constexpr int ARITY = (OP == MCB_OP_BUF) ? 1 : (OP >= MCB_OP_MUX ? 3 : 2);
    W a[U], b[U], c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint4 r = r4[u];
        a[u] = cols.ld(r.y);
        b[u] = (ARITY >= 2) ? cols.ld(r.z) : (W)0;
        c[u] = (ARITY >= 3) ? cols.ld(r.x >> 8) : (W)0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint4 r = r4[u];
        const W neg = (W)0 - (W)((r.x >> 7) & 1u);       // all ones when the result is complemented
        cols.st(r.w, (W)(op_eval<OP, W>(a[u], b[u], c[u]) ^ neg));
    }
}

// Deep, narrow circuits are bound by one memory round trip per dependent batch; prefetching
// the fan-in rows of the batch PF_DIST batches ahead turns that into L1 hits.
#define PF_DIST 2

template <typename W, int U, class Cols>
__device__ __forceinline__ void prefetch_batch(const uint4 *r4, const Cols &cols) {
    const bool three = (r4[0].x & 0x7f) >= MCB_OP_MUX;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint4 r = r4[u];
        cols.prefetch(r.y);
        cols.prefetch(r.z);
        if (three) cols.prefetch(r.x >> 8);
    }
}

template <typename W, int U, class Cols>
__device__ __forceinline__ void run_batch(const uint4 *r4, const Cols &cols) {
    switch (r4[0].x & 0x7f) {                            // uniform: one decode per batch
        case MCB_OP_BUF:  run_batch_op<MCB_OP_BUF, W, U>(r4, cols); break;
        case MCB_OP_AND:  run_batch_op<MCB_OP_AND, W, U>(r4, cols); break;
        case MCB_OP_OR:   run_batch_op<MCB_OP_OR, W, U>(r4, cols); break;
        case MCB_OP_XOR:  run_batch_op<MCB_OP_XOR, W, U>(r4, cols); break;
        case MCB_OP_ANDN: run_batch_op<MCB_OP_ANDN, W, U>(r4, cols); break;
        case MCB_OP_MUX:  run_batch_op<MCB_OP_MUX, W, U>(r4, cols); break;
        case MCB_OP_AND3: run_batch_op<MCB_OP_AND3, W, U>(r4, cols); break;
        case MCB_OP_OR3:  run_batch_op<MCB_OP_OR3, W, U>(r4, cols); break;
        case MCB_OP_XOR3: run_batch_op<MCB_OP_XOR3, W, U>(r4, cols); break;
        case MCB_OP_MAJ:  run_batch_op<MCB_OP_MAJ, W, U>(r4, cols); break;
        default: break;                                   // a batch of NOPs
    }
}

template <typename W, int U, bool TMA, bool SMEM_STATE, bool UNIFIED>
__global__ void __launch_bounds__(256)
k_run_resident(const uint4 *__restrict__ recs, int n_recs, View v, long long steps, int n_slots,
               int g_prefetch) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4 *rbuf = reinterpret_cast<uint4 *>(smem_raw);                  // [2][REC_CHUNK]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + 2 * REC_CHUNK * sizeof(uint4));
    W *sstate = reinterpret_cast<W *>(smem_raw + 2 * REC_CHUNK * sizeof(uint4) + 128);

    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = col < v.n_words * (long long)(8 / sizeof(W));
    const int tid = threadIdx.x;
    const int n_chunks = (n_recs + REC_CHUNK - 1) / REC_CHUNK;
    const long long total = steps * (long long)n_chunks;

    const GlobalCols<W, UNIFIED> gcols(v, col);
    SmemCols<W> scols{sstate + tid, (int)blockDim.x};

    if (SMEM_STATE) {
        // bring this CTA's columns of every slot on chip once (coalesced: thread = column)
        for (int s = 0; s < n_slots; ++s) scols.st(s, active ? gcols.ld(s) : (W)0);
    }
    if (TMA) {
        if (tid == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && total > 0) {
            const int cnt = min(REC_CHUNK, n_recs);
            mbar_expect_tx(&bars[0], cnt * 16);
            bulk_g2s(rbuf, recs, cnt * 16, &bars[0]);
        }
    }
    for (long long it = 0; it < total; ++it) {
        const int chunk = (int)(it % n_chunks);
        const int buf = (int)(it & 1);
        const int base = chunk * REC_CHUNK;
        const int cnt = min(REC_CHUNK, n_recs - base);
        if (TMA) {
            // prefetch the next chunk into the other buffer (free since the barrier below)
            if (tid == 0 && it + 1 < total) {
                const int nchunk = (int)((it + 1) % n_chunks);
                const int nbase = nchunk * REC_CHUNK;
                const int ncnt = min(REC_CHUNK, n_recs - nbase);
                mbar_expect_tx(&bars[buf ^ 1], ncnt * 16);
                bulk_g2s(rbuf + (buf ^ 1) * REC_CHUNK, recs + nbase, ncnt * 16, &bars[buf ^ 1]);
            }
            mbar_wait(&bars[buf], (uint32_t)((it >> 1) & 1));
        } else {
            for (int i = tid; i < cnt; i += blockDim.x) rbuf[buf * REC_CHUNK + i] = __ldg(recs + base + i);
            __syncthreads();
        }
        const uint4 *rb = rbuf + buf * REC_CHUNK;
        if (active) {
            const bool pf = !SMEM_STATE && g_prefetch;
            if (pf)
                for (int i = 0; i < PF_DIST * U && i < cnt; i += U) prefetch_batch<W, U>(rb + i, gcols);
            for (int i = 0; i < cnt; i += U) {
                if (pf && i + PF_DIST * U < cnt) prefetch_batch<W, U>(rb + i + PF_DIST * U, gcols);
                if (SMEM_STATE) run_batch<W, U>(rb + i, scols);
                else run_batch<W, U>(rb + i, gcols);
            }
        }
        __syncthreads();     // everyone is done with rbuf[buf] before it is refilled
    }
    if (SMEM_STATE) {
        if (active)
            for (int s = 0; s < n_slots; ++s) gcols.st(s, scols.ld(s));
    }
}

template <typename W, int U>
static int launch_resident_wu(const uint4 *recs, int n_recs, const View &v, long long steps,
                              int n_slots, cudaStream_t s) {
    const size_t fixed = 2 * REC_CHUNK * sizeof(uint4) + 128;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int max_smem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const long long cols = v.n_words * (long long)(8 / sizeof(W));
    int block = g_tune.resident_block;
    if (block > 256) block = 256;
    if (block < 32) block = 32;
    // shrink the CTA until the grid covers the SMs (few columns) ...
    while (block > 32 && (cols + block - 1) / block < sm_count()) block /= 2;
    bool smem_state = false;
    size_t smem = fixed;
    if (g_tune.resident_smem) {
        // ... and keep the whole slot file of the CTA's columns on chip when it fits
        for (int b = block; b >= 32; b /= 2) {
            size_t need = fixed + (size_t)n_slots * b * sizeof(W);
            if (need <= (size_t)max_smem) { smem_state = true; smem = need; block = b; break; }
        }
    }
    const long long grid = (cols + block - 1) / block;
    const bool tma = g_tune.resident_tma != 0;
    // one plane when the scratch rows simply continue the persistent rows
    const bool unified = v.pstride * 8 < (1ll << 32) && v.sstride * 8 < (1ll << 32) &&
                         (v.scratch_biased == nullptr ||
                          (v.scratch_biased == v.persist && v.sstride == v.pstride));
    if (v.pstride * 8 >= (1ll << 32) || v.sstride * 8 >= (1ll << 32))
        return fail(MCB_ERR_ARG, "row stride of 4 GB or more is not supported by the resident kernel");
#define MCB_LR(T, S, UF)                                                                   \
    do {                                                                                   \
        CUDA_TRY(cudaFuncSetAttribute(k_run_resident<W, U, T, S, UF>,                      \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_run_resident<W, U, T, S, UF><<<(unsigned)grid, block, smem, s>>>(recs, n_recs, v, steps, n_slots, \
                                                                         g_tune.resident_prefetch); \
    } while (0)
    if (unified) {
        if (tma && smem_state) MCB_LR(true, true, true);
        else if (tma) MCB_LR(true, false, true);
        else if (smem_state) MCB_LR(false, true, true);
        else MCB_LR(false, false, true);
    } else {
        if (tma && smem_state) MCB_LR(true, true, false);
        else if (tma) MCB_LR(true, false, false);
        else if (smem_state) MCB_LR(false, true, false);
        else MCB_LR(false, false, false);
    }
#undef MCB_LR
    LAUNCH_CHECK("k_run_resident");
    return 0;
}

template <typename W>
static int launch_resident_w(int unroll, const uint4 *recs, int n_recs, const View &v,
                             long long steps, int n_slots, cudaStream_t s) {
    if (unroll == 8) return launch_resident_wu<W, 8>(recs, n_recs, v, steps, n_slots, s);
    return launch_resident_wu<W, 4>(recs, n_recs, v, steps, n_slots, s);
}

static int launch_resident(const uint4 *recs, int n_recs, const View &v, long long steps,
                           int n_slots, int unroll, int word_bytes, cudaStream_t s) {
    if (n_recs <= 0 || steps <= 0 || v.n_words <= 0) return 0;
    if (unroll != 4 && unroll != 8)
        return fail(MCB_ERR_ARG, "resident batch size must be 4 or 8");
    if (n_recs % unroll != 0) return fail(MCB_ERR_ARG, "resident stream must be padded to %d", unroll);
    if (word_bytes == 0) word_bytes = g_tune.resident_word_bytes;
    if (word_bytes == 0) {
        // A narrower word multiplies the independent threads (latency hiding) but also the
        // instructions per byte, so split only until every scheduler has ~3 warps.
        const long long want = (long long)sm_count() * 32;   // at least a warp per SM
        word_bytes = 8;
        while (word_bytes > 1 && v.n_words * (8 / word_bytes) < want) word_bytes /= 2;
    }
    switch (word_bytes) {
        case 8: return launch_resident_w<uint64_t>(unroll, recs, n_recs, v, steps, n_slots, s);
        case 4: return launch_resident_w<uint32_t>(unroll, recs, n_recs, v, steps, n_slots, s);
        case 2: return launch_resident_w<uint16_t>(unroll, recs, n_recs, v, steps, n_slots, s);
        case 1: return launch_resident_w<uint8_t>(unroll, recs, n_recs, v, steps, n_slots, s);
        default: return fail(MCB_ERR_ARG, "word_bytes must be 1, 2, 4 or 8");
    }
}

// ------------------------------------------------------------------------------------------
// K3 v3: streamed serial run (k_run_stream).  A consumer thread owns (a slice of) one lane column
// and executes the reference's step() exactly as the reference does - every micro-op in emit
// order, in place - for `steps` steps.  See mcircuit_b200/serial.py for the stream format and for
// the rules that decide which rows may be staged ahead of time.
//
// CTA = 4 consumer warps (128 columns) + 1 producer warp.  The producer walks the stream ahead of
// the consumers: per block (up to 32 records) it stages the block's 816-byte control line
// (cp.async.bulk) and every state row the block reads - one cp.async.bulk.tensor.2d per run of
// 2^k consecutive slots, a [2^k rows x 128 columns] box of the state plane described by a tensor
// map - into one of `n_groups` shared-memory ring groups, completion on the group's `full`
// mbarrier; a group is reused when the four consumer warps have arrived on its `empty` mbarrier.
// Consumers wait on `full`, run the block's segments out of shared memory (operands: ring row,
// stash word, accumulator, guard register, or - for rows the stream stored too recently to stage -
// an ordinary load) and store results straight to global memory.  Nothing on the dependent chain
// waits for L2 or HBM.
//
// A GUARD segment (last of its block) is a CTA-wide vote: every consumer warp publishes whether any
// of its lanes has a non-zero enable, arrives on `guard_bar` and waits for the other warps; the
// producer waits on the same barrier.  When the enable is zero everywhere, everybody jumps over the
// region: MUX(0, d, out) -> out and out ^= 0 are the identity, so skipping is exact (event-driven
// evaluation of idle clock phases); a warp that was out-voted runs the region with G = 0, which is
// the same identity.  Consumers execute fence.proxy.async.global before they vote and at the start
// of every FENCE block, which is what makes their earlier stores visible to later TMA reads.
// ------------------------------------------------------------------------------------------
#define STR_BLOCK 32
#define STR_RING_ROWS 32
#define STR_MAX_SEG 24
#define STR_MAX_RUNS 16
#define STR_W_RUN 4                         // meta and runs first: one word per producer lane
#define STR_W_SEG 20
#define STR_W_SLOT 44
#define STR_W_REC 76
#define STR_WORDS 204                       // uint32 per control line
#define STR_CTRL_BYTES 816
#define STR_CTRL_PITCH 896                  // ring rows start 128-byte aligned
#define STR_STASH 48
#define STR_CONSUMER_WARPS 4
#define STR_NT (32 * STR_CONSUMER_WARPS)    // columns per CTA
#define STR_GB_MAX 6
#define STR_PD 4                            // control lines the producer keeps in registers ahead of its issue point
#define STR_HEAD_BYTES 1024                 // barriers and votes
#define STR_M24 0x00ffffffu
#define STR_BOX_KINDS 6                     // boxes of 1, 2, 4, 8, 16, 32 rows
enum { TK_NONE = 0, TK_RING = 1, TK_LATE = 2, TK_ACC = 3, TK_GREG = 4, TK_STASH = 5 };
enum { SEG_GENERIC = 0, SEG_MUX = 1, SEG_XM = 2, SEG_GUARD = 3 };
#define MCB_OP_GUARD 11

// one [rows x 128 columns] box of a state plane -> shared memory (row-major, dense)
__device__ __forceinline__ void tma_box_g2s(void *dst, const void *tmap, int col, int row, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        :: "r"(smem_u32(dst)), "l"(tmap), "r"(col), "r"(row), "r"(smem_u32(bar)) : "memory");
}

struct StreamSmem {
    uint64_t full[STR_GB_MAX];
    uint64_t empty[STR_GB_MAX];
    uint64_t guard_bar;
    uint32_t votes[2][STR_CONSUMER_WARPS];
};

template <typename W, bool UNIFIED>
__device__ __forceinline__ W stream_operand(uint32_t kind, uint32_t field, const W *rows, const W *stash,
                                            const GlobalCols<W, UNIFIED> &cols, W acc, W greg, bool active) {
    switch (kind) {
        case TK_RING:  return rows[field * STR_NT];
        case TK_LATE:  return active ? cols.ld(field) : (W)0;
        case TK_ACC:   return acc;
        case TK_GREG:  return greg;
        case TK_STASH: return stash[field * STR_NT];
        default:       return (W)0;
    }
}

// The same, branch-free for the operand kinds that live in registers or shared memory: one load from an address
// that is valid whatever the kind, then selects.  Generic records decode three operands each; as three jump tables
// they cost three indirect branches per record on a single warp.
template <typename W, bool UNIFIED>
__device__ __forceinline__ W stream_operand_bf(uint32_t kind, uint32_t field, const W *rows, const W *stash,
                                               const GlobalCols<W, UNIFIED> &cols, W acc, W greg, bool active) {
    const bool in_smem = kind == TK_RING || kind == TK_STASH;
    const W *p = (kind == TK_STASH ? stash : rows) + (in_smem ? field : 0u) * STR_NT;
    W x = *p;
    if (kind == TK_LATE) x = active ? cols.ld(field) : (W)0;        // rare; uniform
    x = kind == TK_ACC ? acc : x;
    x = kind == TK_GREG ? greg : x;
    x = kind == TK_NONE ? (W)0 : x;
    return x;
}

// N enabled registers in a row: out = G ? d : out.  `rows` points at the ring row of the first unit (row u
// holds the old value of unit u), `slots` at its slot (= the slot to store to); every stage takes the
// previous stage's NEW value (the reference's in-place order) - pure register forwarding.  Straight-line:
// the compiler hoists all shared-memory reads and address arithmetic above the dependent chain.
template <typename W, bool UNIFIED, int N>
__device__ __forceinline__ void stream_mux_chunk(const W *rows, const uint32_t *slots, const GlobalCols<W, UNIFIED> &cols,
                                                 W &d, W greg, bool active) {
    uint32_t out[N];
    W old[N];
#pragma unroll
    for (int u = 0; u < N; ++u) {
        out[u] = slots[u];
        old[u] = rows[u * STR_NT];
    }
#pragma unroll
    for (int u = 0; u < N; ++u) {
        d = (greg & d) | ((W)~greg & old[u]);
        if (active) cols.st(out[u], d);
    }
}

// N register-pipeline stages: x = acc ^ q; out = G ? x : out.  `q` / `old` point at the first stage's rows.
template <typename W, bool UNIFIED, int N>
__device__ __forceinline__ void stream_xm_chunk(const W *qrows, const W *orows, const uint32_t *slots,
                                                const GlobalCols<W, UNIFIED> &cols, W &d, W greg, bool active) {
    uint32_t out[N];
    W q[N], old[N];
#pragma unroll
    for (int u = 0; u < N; ++u) {
        out[u] = slots[u];
        q[u] = qrows[u * STR_NT];
        old[u] = orows[u * STR_NT];
    }
#pragma unroll
    for (int u = 0; u < N; ++u) {
        d = (greg & (d ^ q[u])) | ((W)~greg & old[u]);
        if (active) cols.st(out[u], d);
    }
}

template <typename W, bool UNIFIED>
__global__ void __launch_bounds__(32 * (STR_CONSUMER_WARPS + 1), 1)
k_run_stream(const uint32_t *__restrict__ stream, int n_blocks, View v, long long steps, long long n_cols,
             const unsigned char *__restrict__ tmaps, unsigned long long *__restrict__ skipped, int n_groups,
             int ring_rows, unsigned char *__restrict__ parking) {
    extern __shared__ __align__(128) unsigned char smem[];
    StreamSmem *sh = reinterpret_cast<StreamSmem *>(smem);
    W *stash_base = reinterpret_cast<W *>(smem + STR_HEAD_BYTES);
    const uint32_t row_pitch = STR_NT * sizeof(W);
    unsigned char *groups = smem + STR_HEAD_BYTES + STR_STASH * row_pitch;
    const uint32_t group_bytes = STR_CTRL_PITCH + ring_rows * row_pitch;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long c0 = (long long)blockIdx.x * STR_NT;

    if (tid == 0) {
        for (int g = 0; g < n_groups; ++g) {
            mbar_init(&sh->full[g], 1);
            mbar_init(&sh->empty[g], STR_CONSUMER_WARPS);
        }
        mbar_init(&sh->guard_bar, STR_CONSUMER_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == STR_CONSUMER_WARPS) {
        // ---------------- producer ----------------
        uint32_t g = 0, ph = 1, gph = 0, gidx = 0;   // ph: the parity that lets the first pass through a fresh barrier
        unsigned char *grp = groups;
        // The control words of the next STR_PD blocks live in registers (lane j holds word j of each
        // line: meta and runs): a line is read once and never in L1, so without this window every block
        // would pay an L2 round trip on the issue path.  A jump over a skipped region refills the window.
        uint32_t wq[STR_PD];
        int b = 0, pb = 0;                       // block being issued / next block to prefetch
        long long step = 0, pstep = 0;
        bool refill = true;
        const bool my_word = lane < STR_W_SEG;          // meta + runs
        while (step < steps) {
            if (refill) {
                pb = b;
                pstep = step;
#pragma unroll
                for (int k = 0; k < STR_PD; ++k) {
                    wq[k] = (pstep < steps && my_word) ? __ldg(stream + (size_t)pb * STR_WORDS + lane) : 0u;
                    if (++pb == n_blocks) { pb = 0; ++pstep; }
                }
                refill = false;
            }
#pragma unroll
            for (int k = 0; k < STR_PD; ++k) {
                if (step >= steps || refill) break;
                const uint32_t w = wq[k];
                wq[k] = (pstep < steps && my_word) ? __ldg(stream + (size_t)pb * STR_WORDS + lane) : 0u;
                if (++pb == n_blocks) { pb = 0; ++pstep; }
                const uint32_t m0 = __shfl_sync(0xffffffffu, w, 0);
                const uint32_t n_rows = m0 & 0xffu, n_runs = m0 >> 24;
                mbar_wait(&sh->empty[g], ph);
                if (lane == 0) {
                    mbar_expect_tx(&sh->full[g], STR_CTRL_BYTES + n_rows * row_pitch);
                    bulk_g2s(grp, stream + (size_t)b * STR_WORDS, STR_CTRL_BYTES, &sh->full[g]);
                }
                __syncwarp();
                if (lane >= STR_W_RUN && lane < STR_W_RUN + n_runs) {
                    // 2^lg consecutive slots of one plane: one box of its tensor map
                    uint32_t slot = w & STR_M24;
                    const uint32_t first_row = (w >> 24) & 31u, lg = w >> 29;
                    uint32_t plane = 0;
                    if (!UNIFIED && slot >= v.n_persist) { plane = 1; slot -= v.n_persist; }
                    tma_box_g2s(grp + STR_CTRL_PITCH + first_row * row_pitch,
                                tmaps + (size_t)(plane * STR_BOX_KINDS + lg) * 128, (int)c0, (int)slot, &sh->full[g]);
                }
                grp += group_bytes;
                if (++g == (uint32_t)n_groups) { g = 0; ph ^= 1u; grp = groups; }
                int nb = b + 1;
                if (m0 & 0x100u) {                       // guard: follow the consumers' vote
                    mbar_wait(&sh->guard_bar, gph);
                    gph ^= 1u;
                    const volatile uint32_t *vt = sh->votes[gidx];
                    const uint32_t any = vt[0] | vt[1] | vt[2] | vt[3];
                    gidx ^= 1u;
                    if (!any) {
                        nb += (int)__shfl_sync(0xffffffffu, w, 1);
                        refill = true;
                    }
                }
                b = nb;
                if (b >= n_blocks) { b = 0; ++step; }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const long long col = c0 + tid;
    const bool in_range = col < n_cols;
    GlobalCols<W, UNIFIED> cols(v, col);
    if (!in_range) cols.park(parking + tid * 8);     // only in the last CTA: garbage in, garbage out, into a private word
    constexpr bool active = true;                    // ... so no load or store below is predicated
    const W *stash = stash_base + tid;
    W *stash_w = stash_base + tid;
    W acc = 0, greg = 0;
    uint32_t g = 0, ph = 0, gph = 0, gidx = 0;
    unsigned char *grp = groups;
    unsigned long long n_skipped = 0;
    for (long long step = 0; step < steps; ++step) {
        int b = 0;
        while (b < n_blocks) {
            mbar_wait(&sh->full[g], ph);
            const uint32_t *ctl = reinterpret_cast<const uint32_t *>(grp);
            const uint4 *recs = reinterpret_cast<const uint4 *>(grp + STR_W_REC * 4);
            const W *rows = reinterpret_cast<const W *>(grp + STR_CTRL_PITCH) + tid;
            const uint4 m = *reinterpret_cast<const uint4 *>(ctl);
            if (m.x & 0x200u) fence_async_global();
            const uint32_t n_seg = (m.x >> 16) & 0xffu;
            uint32_t guard_rec = 0;
            uint32_t sw_next = ctl[STR_W_SEG];
            for (uint32_t k = 0; k < n_seg; ++k) {
                const uint32_t sw = sw_next;
                sw_next = ctl[STR_W_SEG + k + 1];        // the next segment's word, off this segment's critical path
                                                         // (one word past the table is the first row slot: harmless)
                const uint32_t kind = sw & 3u, f = (sw >> 8) & 31u, r0 = (sw >> 13) & 31u, so = (sw >> 18) & 63u;
                int n = (int)((sw >> 2) & 63u);
                if (kind == SEG_MUX) {
                    W d = acc;
                    if (!(sw & (1u << 24))) {                 // the head takes its data from somewhere else than the accumulator
                        const uint4 rh = recs[f];
                        d = stream_operand<W, UNIFIED>((rh.y >> 27) & 7u, rh.y & STR_M24, rows, stash, cols, acc, greg, active);
                    }
                    const W *rw = rows + r0 * STR_NT;
                    const uint32_t *sl = ctl + STR_W_SLOT + r0;
#define MCB_MUX_STEP(N) if (n >= N) { stream_mux_chunk<W, UNIFIED, N>(rw, sl, cols, d, greg, active); rw += N * STR_NT; sl += N; n -= N; }
                    MCB_MUX_STEP(16) MCB_MUX_STEP(16) MCB_MUX_STEP(8) MCB_MUX_STEP(4) MCB_MUX_STEP(2) MCB_MUX_STEP(1)
#undef MCB_MUX_STEP
                    acc = d;
                    if (so) stash_w[(so - 1) * STR_NT] = acc;      // a pattern segment ends where a value goes to the stash
                } else if (kind == SEG_XM && (sw & (1u << 25))) {
                    // stages whose q does not come through the ring: fetched per stage as its XOR record says
                    W d = acc;
                    const W *ow = rows + r0 * STR_NT;
                    const uint32_t *sl = ctl + STR_W_SLOT + r0;
                    for (int u = 0; u < n; ++u) {
                        const uint4 x = recs[f + 2 * u];
                        const W q = stream_operand_bf<W, UNIFIED>((x.y >> 27) & 7u, x.y & STR_M24, rows, stash, cols, acc, greg, active);
                        d = (greg & (d ^ q)) | ((W)~greg & ow[u * STR_NT]);
                        if (active) cols.st(sl[u], d);
                    }
                    acc = d;
                    if (so) stash_w[(so - 1) * STR_NT] = acc;
                } else if (kind == SEG_XM) {
                    const W *qw = rows + r0 * STR_NT, *ow = qw + n * STR_NT;
                    const uint32_t *sl = ctl + STR_W_SLOT + r0 + n;
                    W d = acc;
#define MCB_XM_STEP(N) if (n >= N) { stream_xm_chunk<W, UNIFIED, N>(qw, ow, sl, cols, d, greg, active); \
                                     qw += N * STR_NT; ow += N * STR_NT; sl += N; n -= N; }
                    MCB_XM_STEP(8) MCB_XM_STEP(8) MCB_XM_STEP(4) MCB_XM_STEP(2) MCB_XM_STEP(1)
#undef MCB_XM_STEP
                    acc = d;
                    if (so) stash_w[(so - 1) * STR_NT] = acc;
                } else if (kind == SEG_GUARD) {
                    guard_rec = f;
                } else {
#pragma unroll 2
                    for (int u = 0; u < n; ++u) {
                        const uint4 r = recs[f + u];
                        const uint32_t flags = r.x >> 24;
                        // one dispatch per record (on the opcode); only the operands the opcode has are fetched
#define MCB_FA stream_operand_bf<W, UNIFIED>((r.y >> 24) & 7u, r.x & STR_M24, rows, stash, cols, acc, greg, active)
#define MCB_FB stream_operand_bf<W, UNIFIED>((r.y >> 27) & 7u, r.y & STR_M24, rows, stash, cols, acc, greg, active)
#define MCB_FC stream_operand_bf<W, UNIFIED>((r.z >> 24) & 7u, r.z & STR_M24, rows, stash, cols, acc, greg, active)
                        W res;
                        switch (flags & 15u) {
                            case MCB_OP_BUF:  res = MCB_FA; break;
                            case MCB_OP_AND:  { const W a = MCB_FA, b = MCB_FB; res = a & b; } break;
                            case MCB_OP_OR:   { const W a = MCB_FA, b = MCB_FB; res = a | b; } break;
                            case MCB_OP_XOR:  { const W a = MCB_FA, b = MCB_FB; res = a ^ b; } break;
                            case MCB_OP_ANDN: { const W a = MCB_FA, b = MCB_FB; res = a & (W)~b; } break;
                            case MCB_OP_MUX:  { const W a = MCB_FA, b = MCB_FB, c = MCB_FC; res = (a & b) | ((W)~a & c); } break;
                            case MCB_OP_AND3: { const W a = MCB_FA, b = MCB_FB, c = MCB_FC; res = a & b & c; } break;
                            case MCB_OP_OR3:  { const W a = MCB_FA, b = MCB_FB, c = MCB_FC; res = a | b | c; } break;
                            case MCB_OP_XOR3: { const W a = MCB_FA, b = MCB_FB, c = MCB_FC; res = a ^ b ^ c; } break;
                            case MCB_OP_MAJ:  { const W a = MCB_FA, b = MCB_FB, c = MCB_FC; res = (a & b) | (a & c) | (b & c); } break;
                            default:          res = 0; break;
                        }
#undef MCB_FA
#undef MCB_FB
#undef MCB_FC
                        if (flags & 16u) res = (W)~res;
                        if ((flags & 32u) && active) cols.st(r.w & STR_M24, res);
                        const uint32_t so = r.w >> 24;
                        if (so) stash_w[(so - 1) * STR_NT] = res;
                        acc = res;
                    }
                }
            }
            int skip = 0;
            if (m.x & 0x100u) {
                const uint4 r = recs[guard_rec];
                greg = stream_operand<W, UNIFIED>((r.y >> 24) & 7u, r.x & STR_M24, rows, stash, cols, acc, greg, active);
                fence_async_global();                    // everything stored so far becomes visible to later TMA reads
                const uint32_t any = __any_sync(0xffffffffu, in_range && greg != 0);
                __syncwarp();
                if (lane == 0) {
                    sh->votes[gidx][warp] = any;
                    mbar_arrive(&sh->guard_bar);
                }
                mbar_wait(&sh->guard_bar, gph);
                gph ^= 1u;
                const volatile uint32_t *vt = sh->votes[gidx];
                const uint32_t all = vt[0] | vt[1] | vt[2] | vt[3];
                gidx ^= 1u;
                if (!all) {
                    skip = (int)m.y;
                    n_skipped += 1;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh->empty[g]);
            grp += group_bytes;
            if (++g == (uint32_t)n_groups) { g = 0; ph ^= 1u; grp = groups; }
            b += 1 + skip;
        }
    }
    if (skipped && n_skipped && blockIdx.x == 0 && tid == 0) atomicAdd(skipped, n_skipped);
}

// ---- tensor maps of the state planes: [rows = slots][columns of W bytes], boxes of 2^k rows x 128 columns ----
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);
static PFN_tensorMapEncodeTiled tensor_map_encoder() {
    static PFN_tensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(p);
    }
    return fn;
}

struct TmapEntry { unsigned char *d_maps; unsigned long long last_use; };
static std::mutex g_tmap_mu;
static std::unordered_map<std::string, TmapEntry> g_tmaps;
static unsigned long long g_tmap_tick = 0;

// 2 planes x 6 box heights, 128 bytes each, in device memory; cached per geometry
static int stream_tensor_maps(const View &v, long long n_slots, int word_bytes, long long ext_cols, bool unified,
                              const unsigned char **out) {
    // the last geometry used (the common case: the same engine stepping): no formatting, no hashing
    struct Last { const void *p, *s; long long ps, ss, ns, ec; unsigned np; int wb; const unsigned char *maps; };
    static thread_local Last last = {nullptr, nullptr, 0, 0, 0, 0, 0, 0, nullptr};
    static thread_local unsigned long long last_epoch = 0;
    static std::atomic<unsigned long long> epoch{1};         // bumped whenever entries are evicted
    if (last.maps && last_epoch == epoch.load(std::memory_order_acquire) && last.p == v.persist &&
        last.s == v.scratch_biased && last.ps == v.pstride && last.ss == v.sstride && last.ns == n_slots &&
        last.ec == ext_cols && last.np == v.n_persist && last.wb == word_bytes) {
        *out = last.maps;
        return 0;
    }
    char key[256];
    snprintf(key, sizeof key, "%p|%lld|%p|%lld|%u|%lld|%d|%lld", (void *)v.persist, v.pstride, (void *)v.scratch_biased,
             v.sstride, v.n_persist, n_slots, word_bytes, ext_cols);
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto remember = [&](const unsigned char *maps) {
        last = Last{v.persist, v.scratch_biased, v.pstride, v.sstride, n_slots, ext_cols, v.n_persist, word_bytes, maps};
        last_epoch = epoch.load(std::memory_order_acquire);
    };
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) {
        it->second.last_use = ++g_tmap_tick;
        *out = it->second.d_maps;
        remember(*out);
        return 0;
    }
    PFN_tensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return fail(MCB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    alignas(64) CUtensorMap maps[2 * STR_BOX_KINDS];
    memset(maps, 0, sizeof maps);
    const CUtensorMapDataType dt = word_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64
                                 : word_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                                 : word_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    for (int plane = 0; plane < 2; ++plane) {
        void *base;
        long long rows, stride_words;
        if (plane == 0) {
            base = v.persist;
            rows = unified ? n_slots : (long long)v.n_persist;
            stride_words = v.pstride;
        } else {
            if (unified || !v.scratch_biased || n_slots <= (long long)v.n_persist) continue;
            base = v.scratch_biased + (long long)v.n_persist * v.sstride;
            rows = n_slots - (long long)v.n_persist;
            stride_words = v.sstride;
        }
        if (rows < 1) rows = 1;
        for (int lg = 0; lg < STR_BOX_KINDS; ++lg) {
            const cuuint64_t dims[2] = {(cuuint64_t)ext_cols, (cuuint64_t)rows};
            const cuuint64_t strides[1] = {(cuuint64_t)stride_words * 8ull};
            const cuuint32_t box[2] = {STR_NT, 1u << lg};
            const cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&maps[plane * STR_BOX_KINDS + lg], dt, 2, base, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return fail(MCB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): plane %d, %lld rows x %lld columns of %d bytes, "
                            "row stride %lld bytes, box %d rows", (int)r, plane, rows, ext_cols, word_bytes,
                            stride_words * 8, 1 << lg);
        }
    }
    if (g_tmaps.size() >= 256) {                    // drop the least recently used half
        std::vector<std::pair<unsigned long long, std::string>> order;
        for (auto &kv : g_tmaps) order.push_back({kv.second.last_use, kv.first});
        std::sort(order.begin(), order.end());
        for (size_t i = 0; i < order.size() / 2; ++i) {
            cudaFree(g_tmaps[order[i].second].d_maps);
            g_tmaps.erase(order[i].second);
        }
        epoch.fetch_add(1, std::memory_order_acq_rel);
    }
    unsigned char *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof maps));
    CUDA_TRY(cudaMemcpy(d, maps, sizeof maps, cudaMemcpyHostToDevice));
    g_tmaps[key] = TmapEntry{d, ++g_tmap_tick};
    *out = d;
    remember(d);
    return 0;
}

template <typename W>
static int launch_stream_w(const uint32_t *stream, int n_blocks, int ring_rows, const View &v, long long n_slots,
                           long long steps, unsigned long long *d_skipped, cudaStream_t s) {
    const long long per_word = 8 / (long long)sizeof(W);
    const long long cols = v.n_words * per_word;
    const long long ext_cols = (v.n_words + 15) / 16 * 16 * per_word;         // rows are padded to 128 bytes
    const long long grid = (cols + STR_NT - 1) / STR_NT;
    if (grid > 0x7fffffffll) return fail(MCB_ERR_ARG, "too many lane columns for one launch");
    if ((reinterpret_cast<uintptr_t>(v.persist) & 15) || ((v.pstride * 8) & 15) ||
        (v.scratch_biased && ((reinterpret_cast<uintptr_t>(v.scratch_biased) & 15) || ((v.sstride * 8) & 15))))
        return fail(MCB_ERR_ARG, "state rows must be 16-byte aligned for TMA");
    const bool unified = v.scratch_biased == nullptr ||
                         (v.scratch_biased == v.persist && v.sstride == v.pstride);
    if (ring_rows < 1) ring_rows = 1;
    if (ring_rows > STR_RING_ROWS) ring_rows = STR_RING_ROWS;
    const unsigned char *d_maps = nullptr;
    int rc = stream_tensor_maps(v, n_slots, (int)sizeof(W), ext_cols, unified, &d_maps);
    if (rc) return rc;
    const size_t row_pitch = STR_NT * sizeof(W);
    const size_t group_bytes = STR_CTRL_PITCH + (size_t)ring_rows * row_pitch;
    const size_t fixed = STR_HEAD_BYTES + STR_STASH * row_pitch;
    int dev = 0, max_smem = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    long long budget = g_tune.stream_smem_kb > 0 ? (long long)g_tune.stream_smem_kb * 1024 : (long long)max_smem;
    if (budget > max_smem) budget = max_smem;
    int groups = (int)((budget - (long long)fixed) / (long long)group_bytes);
    if (groups > STR_GB_MAX) groups = STR_GB_MAX;
    if (g_tune.stream_groups > 0 && groups > g_tune.stream_groups) groups = g_tune.stream_groups;
    if (groups < 2) return fail(MCB_ERR_ARG, "shared memory too small for the stream ring");
    const size_t smem = fixed + (size_t)groups * group_bytes;
    static std::atomic<unsigned char *> parking[64];          // per device: 8 bytes per consumer thread
    unsigned char *park = parking[dev & 63].load(std::memory_order_acquire);
    if (!park) {
        CUDA_TRY(cudaMalloc(&park, 4096));
        CUDA_TRY(cudaMemset(park, 0, 4096));
        unsigned char *expected = nullptr;
        if (!parking[dev & 63].compare_exchange_strong(expected, park)) {
            cudaFree(park);
            park = expected;
        }
    }
    auto kern = unified ? k_run_stream<W, true> : k_run_stream<W, false>;
    // the attribute is per device and per instantiation (W is a template parameter of this function)
    static std::atomic<size_t> smem_set[2][64];
    std::atomic<size_t> &done = smem_set[unified ? 1 : 0][dev & 63];
    if (done.load(std::memory_order_acquire) < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done.store(smem, std::memory_order_release);
    }
    kern<<<(unsigned)grid, 32 * (STR_CONSUMER_WARPS + 1), smem, s>>>(stream, n_blocks, v, steps, cols, d_maps, d_skipped,
                                                                     groups, ring_rows, park);
    LAUNCH_CHECK("k_run_stream");
    return 0;
}

// ------------------------------------------------------------------------------------------
// K4: signatures / toggles.  One CTA per slot; lane words are folded with popc, the 64-bit
// partial sums cross lanes by warp shuffle and cross warps through shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_sum(uint64_t x) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    return x;
}

__device__ __forceinline__ uint64_t block_sum(uint64_t x, uint64_t *sh) {
    x = warp_sum(x);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wid] = x;
    __syncthreads();
    uint64_t t = 0;
    if (wid == 0) {
        t = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0;
        t = warp_sum(t);
    }
    return t;   // valid in warp 0
}

__global__ void __launch_bounds__(256)
k_signature(const int *__restrict__ slots, int n, View v, long long col_offset,
            uint64_t *popcount, uint64_t *checksum, long long chunk_words) {
    // grid.y splits a row into column chunks when there are fewer rows than the GPU needs CTAs; the
    // partial sums of a row then meet in its (zeroed) result words by atomicAdd - sums mod 2^64 commute
    __shared__ uint64_t sh[8];
    const long long j0 = (long long)blockIdx.y * chunk_words;
    const long long j1 = j0 + chunk_words < v.n_words ? j0 + chunk_words : v.n_words;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t *r = row(v, (uint32_t)slots[i]);
        uint64_t pc = 0, cs = 0;
        long long j = j0 + 2ll * threadIdx.x;
        for (; j + 1 < j1; j += 2ll * blockDim.x) {              // rows are 16-byte aligned, chunks even
            const ulonglong2 w = *reinterpret_cast<const ulonglong2 *>(r + j);
            pc += (uint64_t)__popcll(w.x) + (uint64_t)__popcll(w.y);
            cs += w.x * (2ull * (uint64_t)(col_offset + j) + 1ull) + w.y * (2ull * (uint64_t)(col_offset + j + 1) + 1ull);
        }
        if (j < j1) {
            const uint64_t w = r[j];
            pc += (uint64_t)__popcll(w);
            cs += w * (2ull * (uint64_t)(col_offset + j) + 1ull);
        }
        pc = block_sum(pc, sh);
        cs = block_sum(cs, sh);
        if (threadIdx.x == 0) {
            if (gridDim.y == 1) {
                if (popcount) popcount[i] = pc;
                if (checksum) checksum[i] = cs;
            } else {
                if (popcount) atomicAdd(reinterpret_cast<unsigned long long *>(popcount + i), (unsigned long long)pc);
                if (checksum) atomicAdd(reinterpret_cast<unsigned long long *>(checksum + i), (unsigned long long)cs);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_toggle(const int *__restrict__ slots, int n, View v, uint64_t *shadow, long long shstride,
         uint64_t *toggles, long long chunk_words) {
    __shared__ uint64_t sh[8];
    const long long j0 = (long long)blockIdx.y * chunk_words;
    const long long j1 = j0 + chunk_words < v.n_words ? j0 + chunk_words : v.n_words;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t *r = row(v, (uint32_t)slots[i]);
        uint64_t *sd = shadow + (long long)i * shstride;
        uint64_t t = 0;
        for (long long j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
            const uint64_t w = r[j];
            t += (uint64_t)__popcll(w ^ sd[j]);
            sd[j] = w;
        }
        t = block_sum(t, sh);
        if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long *>(toggles + i), (unsigned long long)t);
    }
}

// ------------------------------------------------------------------------------------------
// K5: stimulus, fills, pack / unpack
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
k_stimulus(const int *__restrict__ slots, const int *__restrict__ streams, int n, View v,
           long long col_offset, uint64_t seed) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= v.n_words) return;
    for (int i = blockIdx.y; i < n; i += gridDim.y) {
        const uint64_t s = (uint64_t)(uint32_t)streams[i];
        row(v, (uint32_t)slots[i])[j] = splitmix64(seed ^ (s << 32) ^ (uint64_t)(col_offset + j));
    }
}

__global__ void __launch_bounds__(256)
k_fill_rows(const int *__restrict__ slots, const uint64_t *__restrict__ values, int n, View v) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= v.n_words) return;
    for (int i = blockIdx.y; i < n; i += gridDim.y) row(v, (uint32_t)slots[i])[j] = values[i];
}

// waveform capture: copy the current rows of `n` slots into a sample buffer
__global__ void __launch_bounds__(256)
k_capture(const int *__restrict__ slots, int n, View v, uint64_t *__restrict__ dst, long long dst_stride) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= v.n_words) return;
    for (int i = blockIdx.y; i < n; i += gridDim.y) dst[(long long)i * dst_stride + j] = row(v, (uint32_t)slots[i])[j];
}

// 64 lanes (two per thread of a warp) -> one word per bit plane, by ballot
__global__ void __launch_bounds__(256)
k_pack(const uint64_t *__restrict__ values, int width, const int *__restrict__ slots, View v) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= v.n_words) return;
    const uint64_t lo = values[warp * 64 + lane];
    const uint64_t hi = values[warp * 64 + 32 + lane];
    for (int b = 0; b < width; ++b) {
        const uint32_t wl = __ballot_sync(0xffffffffu, (lo >> b) & 1ull);
        const uint32_t wh = __ballot_sync(0xffffffffu, (hi >> b) & 1ull);
        if (lane == 0) row(v, (uint32_t)slots[b])[warp] = ((uint64_t)wh << 32) | wl;
    }
}

__global__ void __launch_bounds__(256)
k_unpack(uint64_t *__restrict__ values, int width, const int *__restrict__ slots, View v) {
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // lane index
    if (l >= v.n_words * 64) return;
    const long long j = l >> 6;
    const int bit = (int)(l & 63);
    uint64_t x = 0;
    for (int b = 0; b < width; ++b) x |= ((row(v, (uint32_t)slots[b])[j] >> bit) & 1ull) << b;
    values[l] = x;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
static int check_state(const mcb_state *st) {
    if (!st) return fail(MCB_ERR_ARG, "state is NULL");
    if (st->n_words < 0 || st->n_persist < 0 || st->n_slots < st->n_persist)
        return fail(MCB_ERR_ARG, "bad state geometry");
    if (st->n_persist > 0 && !st->d_persist) return fail(MCB_ERR_ARG, "d_persist is NULL");
    if (st->n_slots > st->n_persist && !st->d_scratch) return fail(MCB_ERR_ARG, "d_scratch is NULL");
    if ((st->persist_stride & 1) || (st->scratch_stride & 1))
        return fail(MCB_ERR_ARG, "strides must be even (16-byte column pairs)");
    if (((uintptr_t)st->d_persist & 15) || ((uintptr_t)st->d_scratch & 15))
        return fail(MCB_ERR_ARG, "state planes must be 16-byte aligned");
    if (st->n_persist > 0 && st->persist_stride < ((st->n_words + 1) & ~1ll))
        return fail(MCB_ERR_ARG, "persist_stride smaller than the tile");
    if (st->n_slots > st->n_persist && st->scratch_stride < ((st->n_words + 1) & ~1ll))
        return fail(MCB_ERR_ARG, "scratch_stride smaller than the tile");
    return 0;
}

extern "C" {

int mcb_version(void) { return MCB_VERSION; }

const char *mcb_last_error(void) { return g_err; }

int64_t mcb_launch_count(int reset) {
    long long v = reset ? g_launches.exchange(0) : g_launches.load();
    return (int64_t)v;
}

int mcb_set_tuning(const char *key, int value) {
    if (!key) return fail(MCB_ERR_ARG, "key is NULL");
    int *p = nullptr;
    if (!strcmp(key, "level_block")) p = &g_tune.level_block;
    else if (!strcmp(key, "level_unroll")) p = &g_tune.level_unroll;
    else if (!strcmp(key, "level_ctas_per_sm")) p = &g_tune.level_ctas_per_sm;
    else if (!strcmp(key, "level_cache")) p = &g_tune.level_cache;
    else if (!strcmp(key, "level_pdl")) p = &g_tune.level_pdl;
    else if (!strcmp(key, "resident_block")) p = &g_tune.resident_block;
    else if (!strcmp(key, "resident_tma")) p = &g_tune.resident_tma;
    else if (!strcmp(key, "resident_smem")) p = &g_tune.resident_smem;
    else if (!strcmp(key, "resident_word_bytes")) p = &g_tune.resident_word_bytes;
    else if (!strcmp(key, "resident_prefetch")) p = &g_tune.resident_prefetch;
    else if (!strcmp(key, "graph_min_steps")) p = &g_tune.graph_min_steps;
    else if (!strcmp(key, "persist_unroll")) p = &g_tune.persist_unroll;
    else if (!strcmp(key, "persist_hints")) p = &g_tune.persist_hints;
    else if (!strcmp(key, "persist_ctas_per_sm")) p = &g_tune.persist_ctas_per_sm;
    else if (!strcmp(key, "persist_block")) p = &g_tune.persist_block;
    else if (!strcmp(key, "serial_block")) p = &g_tune.serial_block;
    else if (!strcmp(key, "serial_word_bytes")) p = &g_tune.serial_word_bytes;
    else if (!strcmp(key, "level_tma")) p = &g_tune.level_tma;
    else if (!strcmp(key, "stream_groups")) p = &g_tune.stream_groups;
    else if (!strcmp(key, "stream_smem_kb")) p = &g_tune.stream_smem_kb;
    else return fail(MCB_ERR_ARG, "unknown tuning key %s", key);
    int old = *p;
    *p = value;
    return old;
}

int mcb_reset_tuning(void) {
    g_tune = Tuning();
    return 0;
}

int mcb_device_info(int device, int *sms, int *cc_major, int *cc_minor, int64_t *total_mem,
                    int64_t *l2_bytes) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(MCB_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(MCB_ERR_ARG, "device %d out of range", device);
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = (int64_t)p.totalGlobalMem;
    if (l2_bytes) *l2_bytes = (int64_t)p.l2CacheSize;
    if (p.major != 10)
        return fail(MCB_ERR_ARCH, "device %d is sm_%d%d; this library carries sm_100a code only",
                    device, p.major, p.minor);
    return 0;
}

static int eval_levels_impl(const uint32_t *d_records, const int32_t *h_level_off, int level_begin,
                            int level_end, const mcb_state *st, void *stream, bool hinted) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!d_records || !h_level_off || level_begin < 0 || level_end < level_begin)
        return fail(MCB_ERR_ARG, "bad netlist arguments");
    const View v = make_view(st);
    const uint4 *recs = reinterpret_cast<const uint4 *>(d_records);
    for (int l = level_begin; l < level_end; ++l) {
        const int lo = h_level_off[l], hi = h_level_off[l + 1];
        rc = launch_level(recs + lo, hi - lo, v, (cudaStream_t)stream, hinted);
        if (rc) return rc;
    }
    return 0;
}

int mcb_eval_levels(const uint32_t *d_records, const int32_t *h_level_off, int level_begin,
                    int level_end, const mcb_state *st, void *stream) {
    return eval_levels_impl(d_records, h_level_off, level_begin, level_end, st, stream, false);
}

int mcb_latch(const uint32_t *d_latch_records, int n, const mcb_state *st, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n < 0 || (n > 0 && !d_latch_records)) return fail(MCB_ERR_ARG, "bad latch list");
    return launch_level(reinterpret_cast<const uint4 *>(d_latch_records), n, make_view(st),
                        (cudaStream_t)stream);
}

int mcb_run(const uint32_t *d_records, const int32_t *h_level_off, int n_levels,
            const mcb_state *st, int64_t steps, int mode, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!d_records || !h_level_off || n_levels < 0 || steps < 0)
        return fail(MCB_ERR_ARG, "bad run arguments");
    if (steps == 0 || n_levels == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const View v = make_view(st);
    const uint4 *recs = reinterpret_cast<const uint4 *>(d_records);
    const int unroll = (mode >> 8) & 0xff, word_bytes = (mode >> 16) & 0xff;
    const bool hinted = (mode >> 24) & 1;      // mode 1: the records carry operand hints (bit 31)
    mode &= 0xff;
    if (mode == 0) {
        const int u = unroll ? unroll : MCB_RESIDENT_UNROLL;
        for (int l = 0; l <= n_levels; ++l)
            if (h_level_off[l] % u)
                return fail(MCB_ERR_ARG, "mode 0 needs every segment padded to %d records", u);
        return launch_resident(recs, h_level_off[n_levels], v, steps, st->n_slots, u, word_bytes, s);
    }
    if (mode != 1) return fail(MCB_ERR_ARG, "unknown mode %d", mode);
    if (steps < g_tune.graph_min_steps || s == nullptr) {
        for (int64_t t = 0; t < steps; ++t) {
            rc = eval_levels_impl(d_records, h_level_off, 0, n_levels, st, stream, hinted);
            if (rc) return rc;
        }
        return 0;
    }
    // One step = one CUDA graph (a kernel node per level), captured once per (netlist, tile
    // geometry, tuning) and kept: replaying it costs one launch call per step and the levels
    // follow each other on the device without waiting for the host.
    std::string key;
    {
        uint64_t h = 1469598103934665603ull;                       // FNV-1a of the level table
        for (int l = 0; l <= n_levels; ++l) { h ^= (uint64_t)(uint32_t)h_level_off[l]; h *= 1099511628211ull; }
        const int tune[6] = {g_tune.level_block, g_tune.level_unroll, g_tune.level_ctas_per_sm, g_tune.level_cache,
                             g_tune.level_pdl, g_tune.level_tma};
        int dev = 0;
        cudaGetDevice(&dev);
        key.append((const char *)&d_records, sizeof(d_records));
        key.append((const char *)&h, sizeof(h));
        key.append((const char *)&n_levels, sizeof(n_levels));
        key.append((const char *)st, sizeof(*st));
        key.append((const char *)tune, sizeof(tune));
        key.append((const char *)&dev, sizeof(dev));
        key.push_back(hinted ? 'h' : 'p');
    }
    // capture (first use) and every launch happen under the cache mutex
    std::lock_guard<std::mutex> lock(g_graph_mu);
    auto it = g_graphs.find(key);
    if (it == g_graphs.end()) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const long long before = g_launches.load();
        CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        rc = eval_levels_impl(d_records, h_level_off, 0, n_levels, st, stream, hinted);
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        const long long per_step = g_launches.load() - before;   // nodes captured, nothing ran yet
        g_launches.store(before);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(MCB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail(MCB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
        if (g_graphs.size() >= GRAPH_CACHE_CAP) evict_graphs_locked();
        it = g_graphs.emplace(key, GraphEntry{exec, per_step, 0}).first;
    }
    it->second.last_use = ++g_graph_tick;
    for (int64_t t = 0; t < steps; ++t) {
        cudaError_t le = cudaGraphLaunch(it->second.exec, s);
        if (le != cudaSuccess) return fail(MCB_ERR_CUDA, "graph launch failed: %s", cudaGetErrorString(le));
        g_launches.fetch_add(it->second.nodes, std::memory_order_relaxed);
    }
    return 0;
}

int mcb_run_serial(const uint32_t *d_stream, int n_blocks, const mcb_state *st, int64_t steps,
                   int word_bytes, uint64_t *d_skipped, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n_blocks < 0 || (n_blocks > 0 && !d_stream) || steps < 0)
        return fail(MCB_ERR_ARG, "bad serial-run arguments");
    if (n_blocks == 0 || steps == 0 || st->n_words == 0) return 0;
    if (st->n_slots > (1 << 24)) return fail(MCB_ERR_ARG, "serial records address at most 2^24 slots");
    if (reinterpret_cast<uintptr_t>(d_stream) & 15) return fail(MCB_ERR_ARG, "the stream must be 16-byte aligned");
    const View v = make_view(st);
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *sk = reinterpret_cast<unsigned long long *>(d_skipped);
    int ring_rows = word_bytes >> 8;                    // optional: rows a group must hold (0: the format's maximum)
    word_bytes &= 0xff;
    if (word_bytes == 0) word_bytes = g_tune.serial_word_bytes;
    if (word_bytes == 0) word_bytes = 8;
    if (ring_rows <= 0) ring_rows = STR_RING_ROWS;
    switch (word_bytes) {
        case 8: return launch_stream_w<uint64_t>(d_stream, n_blocks, ring_rows, v, st->n_slots, steps, sk, s);
        case 4: return launch_stream_w<uint32_t>(d_stream, n_blocks, ring_rows, v, st->n_slots, steps, sk, s);
        case 2: return launch_stream_w<uint16_t>(d_stream, n_blocks, ring_rows, v, st->n_slots, steps, sk, s);
        case 1: return launch_stream_w<uint8_t>(d_stream, n_blocks, ring_rows, v, st->n_slots, steps, sk, s);
        default: return fail(MCB_ERR_ARG, "word_bytes must be 1, 2, 4 or 8");
    }
}

int mcb_graph_cache_clear(void) {
    std::lock_guard<std::mutex> lock(g_graph_mu);
    const int n = (int)g_graphs.size();
    clear_graphs_locked();
    return n;
}

int mcb_graph_cache_clear_owner(const uint32_t *d_records) {
    std::lock_guard<std::mutex> lock(g_graph_mu);
    int n = 0;
    for (auto it = g_graphs.begin(); it != g_graphs.end();) {
        const void *owner = nullptr;
        memcpy(&owner, it->first.data(), sizeof(owner));
        if (owner == (const void *)d_records) { cudaGraphExecDestroy(it->second.exec); it = g_graphs.erase(it); ++n; }
        else ++it;
    }
    return n;
}

int mcb_graph_cache_size(void) {
    std::lock_guard<std::mutex> lock(g_graph_mu);
    return (int)g_graphs.size();
}

int mcb_run_pipelined(const uint32_t *d_records, const int32_t *d_level_off,
                      const int32_t *h_level_off, int n_levels, const mcb_state *st,
                      int64_t tile_words, int in_flight, int64_t scratch_plane_words,
                      uint64_t *d_sync, int64_t steps, void *stream) {
    if (!st) return fail(MCB_ERR_ARG, "state is NULL");
    mcb_state one = *st;                     // geometry of one tile for the stride checks
    if (one.n_words > tile_words) one.n_words = tile_words;
    int rc = check_state(&one);
    if (rc) return rc;
    if (st->n_persist > 0 && st->persist_stride < ((st->n_words + 1) & ~1ll))
        return fail(MCB_ERR_ARG, "persist_stride smaller than the lane extent");
    if (!d_records || !d_level_off || !h_level_off || !d_sync || n_levels < 0 || steps < 0 ||
        tile_words < 2 || (tile_words & 1) || in_flight < 1 || in_flight > 4)
        return fail(MCB_ERR_ARG, "bad pipelined-run arguments");
    if (steps == 0 || n_levels == 0 || st->n_words == 0) return 0;
    const int U = g_tune.persist_unroll >= 4 ? 4 : (g_tune.persist_unroll >= 2 ? 2 : 1);
    for (int l = 0; l <= n_levels; ++l)
        if (h_level_off[l] % 4) return fail(MCB_ERR_ARG, "grouped stream: levels must be multiples of 4");
    PipeArgs a;
    a.persist = st->d_persist;
    a.pstride = st->persist_stride;
    a.scratch = st->d_scratch;
    a.sstride = st->scratch_stride;
    a.splane = scratch_plane_words;
    a.n_persist = (uint32_t)st->n_persist;
    a.n_words = st->n_words;
    a.tile_words = (int)tile_words;
    a.n_tiles = (int)((st->n_words + tile_words - 1) / tile_words);
    a.n_levels = n_levels;
    a.steps = steps;
    a.sync = reinterpret_cast<unsigned long long *>(d_sync);
    if (a.n_tiles > 1 && (st->n_slots > st->n_persist) && scratch_plane_words <= 0)
        return fail(MCB_ERR_ARG, "scratch_plane_words must be given when several tiles are in flight");
    const uint4 *recs = reinterpret_cast<const uint4 *>(d_records);
    cudaStream_t s = (cudaStream_t)stream;
#define MCB_PP(UU)                                                                         \
    switch (in_flight) {                                                                   \
        case 1: return launch_pipelined_ug<UU, 1>(recs, d_level_off, a, s);                 \
        case 2: return launch_pipelined_ug<UU, 2>(recs, d_level_off, a, s);                 \
        case 3: return launch_pipelined_ug<UU, 3>(recs, d_level_off, a, s);                 \
        default: return launch_pipelined_ug<UU, 4>(recs, d_level_off, a, s);                \
    }
    if (U >= 2) { MCB_PP(2) }
    MCB_PP(1)
#undef MCB_PP
}

int mcb_eval_levels_grouped(const uint32_t *d_records, const int32_t *h_level_off, int level_begin,
                            int level_end, const mcb_state *st, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (!d_records || !h_level_off || level_begin < 0 || level_end < level_begin)
        return fail(MCB_ERR_ARG, "bad netlist arguments");
    const View v = make_view(st);
    const uint4 *recs = reinterpret_cast<const uint4 *>(d_records);
    const long long n_vec = (v.n_words + 1) / 2;
    int block = g_tune.level_block;
    if (block > 256) block = 256;
    if (block < 32) block = 32;
    while (block > 32 && block / 2 >= n_vec) block /= 2;
    const long long gx = (n_vec + block - 1) / block;
    const int U = g_tune.level_unroll >= 4 ? 4 : (g_tune.level_unroll >= 2 ? 2 : 1);
    for (int l = level_begin; l < level_end; ++l) {
        const int lo = h_level_off[l], n = h_level_off[l + 1] - lo;
        if (n <= 0) continue;
        if ((lo | n) % 4) return fail(MCB_ERR_ARG, "grouped stream: levels must be multiples of 4");
        long long gy = n / U;
        if (gy > 65535) gy = 65535;
        dim3 grid((unsigned)gx, (unsigned)gy, 1);
        if (U == 4) k_eval_level_g<4><<<grid, block, 0, (cudaStream_t)stream>>>(recs + lo, n, v, n_vec);
        else if (U == 2) k_eval_level_g<2><<<grid, block, 0, (cudaStream_t)stream>>>(recs + lo, n, v, n_vec);
        else k_eval_level_g<1><<<grid, block, 0, (cudaStream_t)stream>>>(recs + lo, n, v, n_vec);
        LAUNCH_CHECK("k_eval_level_g");
    }
    return 0;
}

int mcb_copy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch,
               int64_t width_bytes, int64_t rows, int kind, void *stream) {
    if (!dst || !src || width_bytes < 0 || rows < 0 || dst_pitch < width_bytes || src_pitch < width_bytes)
        return fail(MCB_ERR_ARG, "bad copy2d arguments");
    if (kind != 1 && kind != 2) return fail(MCB_ERR_ARG, "kind must be 1 (H2D) or 2 (D2H)");
    if (width_bytes == 0 || rows == 0) return 0;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width_bytes,
                               (size_t)rows, kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
    return 0;
}

// K4 grids: one CTA per row when there are enough rows, else rows are cut into even column chunks (>= 2,048 words)
static void row_reduce_grid(int n, long long n_words, dim3 *grid, long long *chunk_words) {
    const int want = sm_count() * 8;
    int gx = n < want ? n : want;
    long long chunks = 1;
    if (n < want) {
        chunks = want / (n > 0 ? n : 1);
        const long long max_chunks = (n_words + 2047) / 2048;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        if (chunks > 65535) chunks = 65535;
    }
    long long cw = (n_words + chunks - 1) / chunks;
    cw = (cw + 1) / 2 * 2;
    chunks = (n_words + cw - 1) / cw;
    *grid = dim3((unsigned)gx, (unsigned)chunks, 1);
    *chunk_words = cw;
}

int mcb_signature(const int32_t *d_slots, int n, const mcb_state *st, int64_t col_offset,
                  uint64_t *d_popcount, uint64_t *d_checksum, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n < 0 || (n > 0 && !d_slots)) return fail(MCB_ERR_ARG, "bad slot list");
    if (n == 0 || st->n_words == 0) return 0;
    dim3 grid;
    long long chunk_words;
    row_reduce_grid(n, st->n_words, &grid, &chunk_words);
    if (grid.y > 1) {                                    // partial sums meet by atomicAdd
        if (d_popcount) CUDA_TRY(cudaMemsetAsync(d_popcount, 0, (size_t)n * 8, (cudaStream_t)stream));
        if (d_checksum) CUDA_TRY(cudaMemsetAsync(d_checksum, 0, (size_t)n * 8, (cudaStream_t)stream));
    }
    k_signature<<<grid, 256, 0, (cudaStream_t)stream>>>(d_slots, n, make_view(st), col_offset,
                                                         d_popcount, d_checksum, chunk_words);
    LAUNCH_CHECK("k_signature");
    return 0;
}

int mcb_toggle_accum(const int32_t *d_slots, int n, const mcb_state *st, uint64_t *d_shadow,
                     int64_t shadow_stride, uint64_t *d_toggles, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_slots || !d_shadow || !d_toggles)) || shadow_stride < st->n_words)
        return fail(MCB_ERR_ARG, "bad toggle arguments");
    if (n == 0 || st->n_words == 0) return 0;
    dim3 grid;
    long long chunk_words;
    row_reduce_grid(n, st->n_words, &grid, &chunk_words);
    k_toggle<<<grid, 256, 0, (cudaStream_t)stream>>>(d_slots, n, make_view(st), d_shadow,
                                                      shadow_stride, d_toggles, chunk_words);
    LAUNCH_CHECK("k_toggle");
    return 0;
}

static dim3 row_grid(long long n_words, int n) {
    long long gx = (n_words + 255) / 256;
    long long gy = n;
    long long cap = (long long)sm_count() * 16 / (gx > 0 ? gx : 1);
    if (cap < 1) cap = 1;
    if (gy > cap) gy = cap;
    if (gy > 65535) gy = 65535;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}

int mcb_stimulus(const int32_t *d_slots, const int32_t *d_streams, int n, const mcb_state *st,
                 int64_t col_offset, uint64_t seed, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_slots || !d_streams))) return fail(MCB_ERR_ARG, "bad stimulus list");
    if (n == 0 || st->n_words == 0) return 0;
    k_stimulus<<<row_grid(st->n_words, n), 256, 0, (cudaStream_t)stream>>>(
        d_slots, d_streams, n, make_view(st), col_offset, seed);
    LAUNCH_CHECK("k_stimulus");
    return 0;
}

int mcb_fill_rows(const int32_t *d_slots, const uint64_t *d_values, int n, const mcb_state *st,
                  void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_slots || !d_values))) return fail(MCB_ERR_ARG, "bad fill list");
    if (n == 0 || st->n_words == 0) return 0;
    k_fill_rows<<<row_grid(st->n_words, n), 256, 0, (cudaStream_t)stream>>>(d_slots, d_values, n,
                                                                              make_view(st));
    LAUNCH_CHECK("k_fill_rows");
    return 0;
}

int mcb_capture_rows(const int32_t *d_slots, int n, const mcb_state *st, uint64_t *d_dst,
                     int64_t dst_stride, void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_slots || !d_dst)) || dst_stride < st->n_words)
        return fail(MCB_ERR_ARG, "bad capture arguments");
    if (n == 0 || st->n_words == 0) return 0;
    k_capture<<<row_grid(st->n_words, n), 256, 0, (cudaStream_t)stream>>>(d_slots, n, make_view(st), d_dst,
                                                                            dst_stride);
    LAUNCH_CHECK("k_capture");
    return 0;
}

int mcb_pack(const uint64_t *d_values, int width, const int32_t *d_slots, const mcb_state *st,
             void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (width < 1 || width > 64 || !d_values || !d_slots) return fail(MCB_ERR_ARG, "bad pack arguments");
    if (st->n_words == 0) return 0;
    const long long threads = st->n_words * 32;
    k_pack<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_values, width, d_slots,
                                                                                make_view(st));
    LAUNCH_CHECK("k_pack");
    return 0;
}

int mcb_unpack(uint64_t *d_values, int width, const int32_t *d_slots, const mcb_state *st,
               void *stream) {
    int rc = check_state(st);
    if (rc) return rc;
    if (width < 1 || width > 64 || !d_values || !d_slots) return fail(MCB_ERR_ARG, "bad unpack arguments");
    if (st->n_words == 0) return 0;
    const long long threads = st->n_words * 64;
    k_unpack<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_values, width, d_slots,
                                                                                  make_view(st));
    LAUNCH_CHECK("k_unpack");
    return 0;
}

}  // extern "C"
