/*
 * mcref.c - ORACLE, TEST INFRASTRUCTURE ONLY (see oracle/restate.py for the rules: never
 * linked, loaded or called by the product path).
 *
 * Plain-C restatement of the reference's execution model (matijakevic/mcircuit
 * core/simulator.py): ONE sequential pass, in the reference's emit order, over one storage
 * cell per net, updated in place - generalised from one stimulus instance to many by making
 * every cell a row of 64-lane words (bit b of a w-bit net is row net_base + b).  No
 * levelization, no slot recycling, no shadows: this is the reference's algorithm, not the
 * engine's, which is what makes it a checker.  It is also the CPU baseline ("port") that
 * bench.py times on the host cores, threaded over lane columns.
 *
 * Per-primitive semantics follow core/simulator.py:
 *   NOT      :77-80     GATE   :83-97     COUNTER :100-112   REGISTER :115-127 (intent)
 *   CLOCK    :130-132   ADDER  :135-147   PROPAGATE :210-213 (map_pins=False only)
 * The emit list is produced by oracle/cref.py from the same flattening as restate.py.
 *
 * Parity: pinned against tests/golden (reference JIT outputs) by tests/test_cref.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { T_NOT = 1, T_GATE = 2, T_COUNTER = 3, T_REGISTER = 4, T_CLOCK = 5, T_ADDER = 6, T_PROP = 7 };

#define CB 8 /* columns per block: one cache line of lane words */
#define MAXW 64

typedef struct {
    const int32_t *type;   /* [n_ops]                                                        */
    const int32_t *width;  /* [n_ops]                                                        */
    const int32_t *aux;    /* [n_ops] GATE: num_inputs | op << 8 | negated << 16              */
    const int64_t *pin_off;/* [n_ops + 1] into pins                                           */
    const int64_t *pins;   /* first row of each pin's net, canonical pin order per primitive  */
    int64_t n_ops;
    uint64_t *val;         /* [n_blocks][n_rows][CB]: a block of CB lane columns is contiguous, */
    int64_t n_rows;        /* so one block's whole net file is a dense, cache-friendly array    */
    int64_t blk0, blk1;    /* this worker's column blocks                                     */
    int64_t steps;
} job_t;

#define ROW(j, r) ((j)->val + (int64_t)(r) * CB)   /* within the current block (c = block base) */

/* the hot loop for a pure-gate netlist: kept tiny so the compiler vectorises it */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void gate2(uint64_t *restrict o, const uint64_t *a, const uint64_t *b, int op, int neg, int n) {
    uint64_t m = neg ? ~0ull : 0ull;
    if (op == 0) for (int i = 0; i < n; ++i) o[i] = (a[i] & b[i]) ^ m;
    else if (op == 1) for (int i = 0; i < n; ++i) o[i] = (a[i] | b[i]) ^ m;
    else if (op == 2) for (int i = 0; i < n; ++i) o[i] = (a[i] ^ b[i]) ^ m;
    else for (int i = 0; i < n; ++i) o[i] = a[i] ^ m;       /* unknown op: first input (:267-274) */
}

static void do_gate(const job_t *j, int64_t k, int64_t c, int n) {
    const int w = j->width[k], nin = j->aux[k] & 0xff, op = (j->aux[k] >> 8) & 0xff,
              neg = (j->aux[k] >> 16) & 1;
    const int64_t *p = j->pins + j->pin_off[k];           /* in0..in{n-1}, out */
    for (int b = 0; b < w; ++b) {
        uint64_t *o = ROW(j, p[nin] + b) + c;
        if (nin == 2) {
            gate2(o, ROW(j, p[0] + b) + c, ROW(j, p[1] + b) + c, op, neg, n);
            continue;
        }
        uint64_t res[CB];
        memcpy(res, ROW(j, p[0] + b) + c, n * sizeof(uint64_t));
        for (int i = 1; i < nin; ++i) {
            const uint64_t *v = ROW(j, p[i] + b) + c;
            if (op == 0) for (int x = 0; x < n; ++x) res[x] &= v[x];
            else if (op == 1) for (int x = 0; x < n; ++x) res[x] |= v[x];
            else if (op == 2) for (int x = 0; x < n; ++x) res[x] ^= v[x];
        }
        for (int x = 0; x < n; ++x) o[x] = neg ? ~res[x] : res[x];
    }
}

static void do_not(const job_t *j, int64_t k, int64_t c, int n) {
    const int64_t *p = j->pins + j->pin_off[k];           /* in, out */
    for (int b = 0; b < j->width[k]; ++b) {
        const uint64_t *a = ROW(j, p[0] + b) + c;
        uint64_t *o = ROW(j, p[1] + b) + c;
        for (int x = 0; x < n; ++x) o[x] = ~a[x];
    }
}

static void do_clock(const job_t *j, int64_t k, int64_t c, int n) {
    uint64_t *o = ROW(j, j->pins[j->pin_off[k]]) + c;     /* out */
    for (int x = 0; x < n; ++x) o[x] = ~o[x];
}

/* out = (zext(rise) & (out + 1)) | (~zext(rise) & out), literally, bit-sliced: the mask has
 * only bit 0 = rise (a zero-extended i1), so only bit 0 of out can change. */
static void do_counter(const job_t *j, int64_t k, int64_t c, int n) {
    const int w = j->width[k];
    const int64_t *p = j->pins + j->pin_off[k];           /* clock, out, prevclock */
    for (int x = 0; x < n; ++x) {
        uint64_t *clk = ROW(j, p[0]) + c + x, *prev = ROW(j, p[2]) + c + x;
        const uint64_t rise = ~*prev & *clk;
        uint64_t out[MAXW], inc[MAXW], carry = ~0ull;     /* + 1 */
        for (int b = 0; b < w; ++b) {
            out[b] = ROW(j, p[1] + b)[c + x];
            inc[b] = out[b] ^ carry;
            carry &= out[b];
        }
        for (int b = 0; b < w; ++b) {
            const uint64_t mask = (b == 0) ? rise : 0ull;
            ROW(j, p[1] + b)[c + x] = (mask & inc[b]) | (~mask & out[b]);
        }
        *prev = *clk;                                     /* clock re-read after the store */
    }
}

/* intent of simulator.py:115-127: the rising edge sign-extended to a full mask */
static void do_register(const job_t *j, int64_t k, int64_t c, int n) {
    const int w = j->width[k];
    const int64_t *p = j->pins + j->pin_off[k];           /* clock, data, out, prevclock */
    for (int x = 0; x < n; ++x) {
        uint64_t *clk = ROW(j, p[0]) + c + x, *prev = ROW(j, p[3]) + c + x;
        const uint64_t rise = ~*prev & *clk;
        uint64_t d[MAXW];
        for (int b = 0; b < w; ++b) d[b] = ROW(j, p[1] + b)[c + x];
        for (int b = 0; b < w; ++b) {
            uint64_t *o = &ROW(j, p[2] + b)[c + x];
            *o = (rise & d[b]) | (~rise & *o);
        }
        *prev = *clk;
    }
}

/* s1 = a + b ; s2 = s1 + zext(cin) ; cout = sovf(a, b) | sovf(s1, zext(cin)).  Signed overflow
 * of an add = carry into the sign bit XOR carry out of it (the textbook definition of what
 * llvm.sadd.with.overflow reports). */
static void do_adder(const job_t *j, int64_t k, int64_t c, int n) {
    const int w = j->width[k];
    const int64_t *p = j->pins + j->pin_off[k];           /* cin, a, b, cout, sum */
    for (int x = 0; x < n; ++x) {
        uint64_t a[MAXW], b[MAXW], s1[MAXW], s2[MAXW];
        const uint64_t cin = ROW(j, p[0])[c + x];
        for (int i = 0; i < w; ++i) { a[i] = ROW(j, p[1] + i)[c + x]; b[i] = ROW(j, p[2] + i)[c + x]; }
        uint64_t carry = 0, into_msb = 0;
        for (int i = 0; i < w; ++i) {
            if (i == w - 1) into_msb = carry;
            s1[i] = a[i] ^ b[i] ^ carry;
            carry = (a[i] & b[i]) | (a[i] & carry) | (b[i] & carry);
        }
        const uint64_t ovf1 = into_msb ^ carry;
        carry = 0; into_msb = 0;
        for (int i = 0; i < w; ++i) {
            const uint64_t y = (i == 0) ? cin : 0ull;       /* zext(cin) */
            if (i == w - 1) into_msb = carry;
            s2[i] = s1[i] ^ y ^ carry;
            carry = (s1[i] & y) | (s1[i] & carry) | (y & carry);
        }
        const uint64_t ovf2 = into_msb ^ carry;
        for (int i = 0; i < w; ++i) ROW(j, p[4] + i)[c + x] = s2[i];
        ROW(j, p[3])[c + x] = ovf1 | ovf2;
    }
}

static void do_prop(const job_t *j, int64_t k, int64_t c, int n) {
    const int64_t *p = j->pins + j->pin_off[k];           /* src, dst */
    if (p[0] == p[1]) return;
    for (int b = 0; b < j->width[k]; ++b)
        memcpy(ROW(j, p[1] + b) + c, ROW(j, p[0] + b) + c, n * sizeof(uint64_t));
}

static void *worker(void *arg) {
    const job_t *j = (const job_t *)arg;
    /* a block of columns goes through ALL steps before the next block: lanes are independent */
    for (int64_t blk = j->blk0; blk < j->blk1; ++blk) {
        const int64_t c = blk * j->n_rows * CB;            /* offset of this block's row 0 */
        const int n = CB;
        for (int64_t s = 0; s < j->steps; ++s) {
            for (int64_t k = 0; k < j->n_ops; ++k) {
                switch (j->type[k]) {
                    case T_GATE: do_gate(j, k, c, n); break;
                    case T_NOT: do_not(j, k, c, n); break;
                    case T_CLOCK: do_clock(j, k, c, n); break;
                    case T_COUNTER: do_counter(j, k, c, n); break;
                    case T_REGISTER: do_register(j, k, c, n); break;
                    case T_ADDER: do_adder(j, k, c, n); break;
                    case T_PROP: do_prop(j, k, c, n); break;
                    default: break;
                }
            }
        }
    }
    return NULL;
}

/* Run `steps` passes over `n_blocks` blocks of CB columns with `threads` workers.
 * val is [n_blocks][n_rows][CB].  Returns 0. */
int mcref_run(const int32_t *type, const int32_t *width, const int32_t *aux, const int64_t *pin_off,
              const int64_t *pins, int64_t n_ops, uint64_t *val, int64_t n_rows, int64_t n_blocks,
              int64_t steps, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    for (int64_t k = 0; k < n_ops; ++k)
        if (width[k] < 1 || width[k] > MAXW) return -1;
    int64_t blocks = n_blocks;
    if (threads > blocks) threads = (int)(blocks > 0 ? blocks : 1);
    job_t jobs[256];
    pthread_t tid[256];
    int64_t per = (blocks + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        job_t *j = &jobs[t];
        j->type = type; j->width = width; j->aux = aux; j->pin_off = pin_off; j->pins = pins;
        j->n_ops = n_ops; j->val = val; j->n_rows = n_rows; j->steps = steps;
        j->blk0 = t * per;
        j->blk1 = (t + 1) * per;
        if (j->blk0 > blocks) j->blk0 = blocks;
        if (j->blk1 > blocks) j->blk1 = blocks;
    }
    if (threads == 1) { worker(&jobs[0]); return 0; }
    for (int t = 0; t < threads; ++t) pthread_create(&tid[t], NULL, worker, &jobs[t]);
    for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
    return 0;
}

int mcref_version(void) { return 1; }
int mcref_block_columns(void) { return CB; }
