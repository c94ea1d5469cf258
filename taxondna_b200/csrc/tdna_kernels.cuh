// Kernels of the pairwise-distance engine (sm_100a).  See DESIGN.md for the roofline of each.
#pragma once
#include "tdna_device.cuh"
#include "../../include/taxondna_cuda.h"

namespace tdna {

static_assert(sizeof(tdna_row_result) == 120, "tdna_row_result must be 120 bytes without padding");

// ------------------------------------------------------------------------------------------
// pack: normalised symbols -> bit-planes.  HBM-bound, one pass.
// thread <-> (word k, row) with row fastest, so the 32-byte symbol reads are one sector each and
// the plane writes are contiguous within a 64-row panel.
// ------------------------------------------------------------------------------------------
struct PlaneBits {
    uint32_t b[6];
};
__constant__ uint16_t c_symlut[256];  // symbol -> SymBit set; 0x8000 = not in the alphabet

// rows [row0, row0 + rows) (a multiple of the 64-row panel; rows past n are written as empty): the whole alignment, or one chunk of
// a pipelined upload.  L clamps the row lengths: sites past the alignment's width do not exist.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ sym, const int32_t *__restrict__ len,
                                                   int64_t n, int64_t row0, int64_t rows, int64_t sym_stride, int W, int P, int L, PlaneBits bits,
                                                   uint32_t *__restrict__ planes, int *__restrict__ err) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * W) return;
    int64_t row = row0 + idx % rows;
    int k = (int)(idx / rows);
    uint32_t w[6] = {0, 0, 0, 0, 0, 0};
    if (row < n) {
        int l = min(len[row], L);
        const uint8_t *src = sym + row * sym_stride + 32 * k;  // rows keep the caller's stride: no alignment assumed
        bool bad = false;
#pragma unroll
        for (int b = 0; b < 32; b++) {
            uint32_t cls = 0;  // beyond the row's own length == '?' (Sequence.java:1326-1327)
            if (32 * k + b < l) cls = c_symlut[__ldg(src + b)];
            bad |= (cls & 0x8000u) != 0;
#pragma unroll
            for (int p = 0; p < 6; p++) w[p] |= ((cls & bits.b[p]) ? 1u : 0u) << b;
        }
        if (bad) atomicOr(err, 1);
    }
    for (int p = 0; p < P; p++) planes[plane_index(row, p, k, P, W)] = w[p];
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t *__restrict__ p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA unit, 1-D) primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// producer-side wait: the lane is far ahead of the consumers by design, so it sleeps between
// probes instead of burning the issue slots of the SM sub-partition it shares with a consumer warp
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(400);
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// count + distance tile kernel (the hot kernel).
//
// CTA = 8 consumer warps (16 ty x 16 tx threads) + 1 producer warp.  Tile = 128 rows x TN columns
// of the pair matrix, TN = 16*CN; a consumer thread owns 8 rows x CN columns, register-blocked, two
// 16-bit counts per 32-bit accumulator.  The producer warp's elected lane stages K-chunks of the
// operand planes into shared memory with 1-D bulk async copies (TMA unit; one copy per (panel,
// plane)) that complete on a "full" mbarrier; consumer warps release a stage by arriving on its
// "empty" mbarrier.  There is no CTA-wide barrier in steady state, so warps drift apart and the
// fp64 epilogue of one warp overlaps the integer main loop of the others.  Persistent grid over a
// flattened (tile, K-chunk) sequence.  Integer-issue bound (POPC on the XU pipe): DESIGN.md.
// ------------------------------------------------------------------------------------------
constexpr int TM = 128;
constexpr int KC = 8;     // words (256 sites) per K-chunk
constexpr int NST = 3;    // pipeline stages
constexpr int CONSUMER_WARPS = 8;
constexpr int TILE_THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int PREFIX_SMEM = 512;  // row-tile prefix entries cached in shared memory

template <int KM, int CN> struct TileCfg {
    static constexpr int P = ModeTraits<KM>::P;
    static constexpr int NACC = ModeTraits<KM>::NACC;
    static constexpr int TN = 16 * CN;
    static constexpr int XW = P * (TM / PR) * KC * PR;  // words per stage, x side
    static constexpr int YW = P * (TN / PR) * KC * PR;
    static constexpr int STAGE_WORDS = XW + YW;
    static constexpr size_t SMEM = (size_t)NST * STAGE_WORDS * 4 + 2 * NST * 8 + PREFIX_SMEM * 8 + 64;
    // 2 CTAs per SM when the accumulators are small enough (more warps to hide the epilogue's latency)
    static constexpr int MIN_CTAS = (CN * NACC <= 4) ? 2 : 1;
};

// Streaming Best-Match state (one pass over the triangle, no matrix): see "stream epilogue" below.
struct StreamParams {
    uint32_t *thr;               // [n] per-row candidate threshold, 16.16 fixed point of (needed distance + 2e-5); 65536 = pass all
    unsigned long long *mC;      // [n] min d (fp64 bits) over valid named conspecifics; 0 when the row has no conspecific in the list
    unsigned long long *mPar;    // [2n] min d over valid named columns with even / odd species id
    int *inval;                  // [n] # j != q with shared < minOverlap
    int *flags;                  // [n] TDNA_ROW_*
    int *cnt;                    // [n] candidates emitted for the row
    int32_t *cq, *cj;            // candidate records (row, column, distance), unordered
    double *cd;
    unsigned long long *ncand;   // records emitted so far
    long long cap;               // capacity of cq/cj/cd
    int *overflow;               // set when a record did not fit
    const int32_t *species;
    // what this pass produces (TDNA_WANT_*): Best-Match candidates, PairwiseDistribution classes, Cluster edges
    int want;
    const int32_t *genus, *epi, *full;
    const int4 *panel_range;     // per 64-row panel: (min, max) named species id, (min, max) genus id
    float *cls[2];               // TDNA_PD_INTRA / TDNA_PD_INTER distances as (float)d, unordered (may be null: count only)
    long long cls_cap[2];
    unsigned long long *ncls;    // [2] pushes so far
    // histogram mode of a class (tdna_class_histogram / TDNA_WANT_HIST_*): instead of one float per push, the pushes are COUNTED per
    // distinct fp64 distance in an open-addressing table (key = the bits of d, ~0 = empty) -- O(distinct values) memory however
    // many pairs the class has (10^6 rows: 5 * 10^8 congeneric pairs, a few thousand distinct distances)
    unsigned long long *hkey[2], *hcnt[2];
    uint32_t hmask;              // table size - 1 (a power of two)
    int *hover;                  // set when a table is full
    uint32_t edge_tfix;          // 16.16 fixed point of the edge threshold, rounded up
    double edge_t;
    int32_t *ei, *ej;            // edge records (may be null: count only)
    double *ed;
    long long edge_cap;
    unsigned long long *nedge;
    // count kernel (b) defers its slow path: the epilogue only queues (row, column, raw accumulator) of the pairs that pass
    // the seeded thresholds; tc::resolve_*_kernel turn the queue into candidate records after the pass
    // one 16-byte record per entry (x = row, bit 31: a deferred class pair; y = column, bit 31: a class-only entry; z = raw accumulator;
    // z, w = the exact counts once resolve_min_kernel has left them for resolve_emit_kernel): a lane's entries are contiguous, so its
    // stores fill whole 32-byte sectors (separate 4-byte arrays made every push a partial-sector write: 1.1 GB of L2 write sectors and
    // 4.9 GB of DRAM fills for a 190 MB queue, profiles/r2n_*)
    uint4 *q;
    unsigned long long *qn;
    long long qcap;
};
// (the streaming passes take their StreamParams as a __grid_constant__ kernel parameter: it sits in the launch's own constant bank,
// so two alignments / contexts on one device never share state)

enum { SCHED_RECT = 0, SCHED_SYMMETRIC = 1, SCHED_STREAM = 2 };
enum { EPI_MATRIX = 0, EPI_MATRIX_COUNTS = 1, EPI_STREAM = 2 /* Best-Match candidates only */, EPI_STREAM_X = 3 /* + classes / edges */ };

struct TileParams {
    const uint32_t *planes;
    int W;
    int64_t n;          // sequences
    int min_overlap;
    int64_t row_begin;  // first row of the band (multiple of 128)
    int64_t row_end;    // one past the last row of the band
    int symmetric;      // band == whole matrix: only tiles touching j >= i are computed, both halves written
    int sched;          // SCHED_*: how a tile number maps to (row tile, column tile)
    int first_col_extra;  // SCHED_STREAM off-diagonal phase: column tiles start TM/TN past the diagonal block
    int64_t t_offset, t_stride;  // this launch computes global tiles t_offset + k * t_stride (multi-GPU: rank + k * world)
    int64_t num_tiles;  // tiles of THIS launch
    const int64_t *tile_prefix;  // [row tiles + 1]; symmetric mode only
    int nrt;                     // row tiles
    int nct;                     // column tiles
    double *out;                 // band matrix, row-major, leading dimension ld
    int64_t ld;
    tdna_counts *counts;         // optional, same shape, ld = n
};

// local tile number of this launch -> (row tile, column tile)
template <int TN> __device__ __forceinline__ void tile_coords(const TileParams &p, const int64_t *prefix, int64_t t, int &ti, int &tj) {
    t = p.t_offset + t * p.t_stride;
    if (p.sched == SCHED_RECT) {
        ti = (int)(t / p.nct);
        tj = (int)(t % p.nct);
    } else {
        int lo = 0, hi = p.nrt;  // largest ti with prefix[ti] <= t
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (prefix[mid] <= t) lo = mid; else hi = mid;
        }
        ti = lo;
        tj = (int)(t - prefix[lo]) + ti * (TM / TN) + p.first_col_extra;
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- matrix epilogue: counts -> fp64 distance -> band matrix (registers only) ----------------
template <int KM, int CN, bool WRITE_COUNTS>
__device__ __forceinline__ void matrix_epilogue(const TileParams &p, uint32_t (&acc)[ModeTraits<KM>::NACC][8][CN], int ti, int tj, int tx, int ty) {
    constexpr int NACC = ModeTraits<KM>::NACC, TN = 16 * CN;
    const int64_t i0 = p.row_begin + (int64_t)ti * TM + ty * 8, j0 = (int64_t)tj * TN;
#pragma unroll
    for (int r = 0; r < 8; r++) {  // fully unrolled: acc[][][] must stay in registers
        const int64_t i = i0 + r;
        const bool row_ok = i < p.row_end;
        double dv[CN];
#pragma unroll
        for (int e = 0; e < CN; e++) {
            uint32_t a1 = 0;
            if constexpr (NACC == 2) a1 = acc[1][r][e];
            Counts cn = unpack_counts<KM>(acc[0][r][e], a1);
            dv[e] = distance_from_counts<KM>(cn, p.min_overlap);
            if constexpr (WRITE_COUNTS) {
                const int64_t j = j0 + (e >> 2) * PR + 4 * tx + (e & 3);
                if (row_ok && j < p.n) {
                    tdna_counts o = {cn.shared, cn.c1, cn.c2, cn.c3};
                    p.counts[(i - p.row_begin) * p.n + j] = o;
                }
            }
        }
        if (!row_ok) continue;
#pragma unroll
        for (int h = 0; h < CN / 4; h++) {
            const int64_t jb = j0 + h * PR + 4 * tx;
            double *dst = p.out + (i - p.row_begin) * p.ld + jb;
            if (jb + 3 < p.n) {  // ld and jb are multiples of 4 doubles: 32-byte aligned
                reinterpret_cast<double2 *>(dst)[0] = make_double2(dv[4 * h + 0], dv[4 * h + 1]);
                reinterpret_cast<double2 *>(dst)[1] = make_double2(dv[4 * h + 2], dv[4 * h + 3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (jb + e < p.n) dst[e] = dv[4 * h + e];
            }
            if (p.symmetric) {  // mirror (band == whole matrix, row_begin == 0)
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (jb + e < p.n && jb + e > i) p.out[(jb + e) * p.ld + i] = dv[4 * h + e];
            }
        }
    }
}

// ---- stream epilogue: no matrix -----------------------------------------------------------------
// Every unordered pair i < j is seen exactly once (triangle tiles).  A pair becomes a CANDIDATE of
// row i (and, separately, of row j) when its distance is not above that row's threshold
//     thr[q] >= max(min d over named conspecifics, min d over named even-species columns,
//                   min d over named odd-species columns) + 2e-5,
// which bounds every distance tdna_row_result refers to (best <= con/allo/other <= that max: two
// named columns of different parity are of different species, so one of them is not of species(best))
// plus the near-tie window.  Thresholds only ever decrease, so the candidate list of a row is a
// superset of {j : d(q,j) <= final threshold} and stream_finalize_kernel can run the exact two-sweep
// row summary on the list alone.  The per-pair test is integer only (mism << 16 <= thr * den, a
// conservative statement of d <= thr: for K2P, d >= (ts+tv)/n); fp64 happens on candidates only.
#define TDNA_THR_ALL 65536u
#define TDNA_WANT_SEED 0x100  // internal: the diagonal-block pass only tightens the per-row minima / thresholds, it emits nothing
#define TDNA_INF_BITS 0x7ff0000000000000ull

__device__ __forceinline__ void stream_emit_side(const StreamParams &sp, int q, int o, double d, uint32_t lhs, uint32_t den) {
    if (!(sp.want & TDNA_WANT_BEST_MATCH)) return;
    const uint32_t fresh = *reinterpret_cast<volatile uint32_t *>(sp.thr + q);  // other CTAs tighten it all the time
    if (lhs > fresh * den) return;
    if (!(sp.want & TDNA_WANT_SEED)) {
        const unsigned long long pos = atomicAdd(sp.ncand, 1ull);
        if ((long long)pos >= sp.cap) { *sp.overflow = 1; return; }
        sp.cq[pos] = q;
        sp.cj[pos] = o;
        sp.cd[pos] = d;
        atomicAdd(sp.cnt + q, 1);
    }
    const int spo = sp.species[o];
    if (spo < 0) return;  // unnamed columns bound nothing
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    if (spo == sp.species[q]) atomicMin(sp.mC + q, bits);
    atomicMin(sp.mPar + 2 * (int64_t)q + (spo & 1), bits);
    const unsigned long long mc = *reinterpret_cast<volatile unsigned long long *>(sp.mC + q);
    const unsigned long long me = *reinterpret_cast<volatile unsigned long long *>(sp.mPar + 2 * (int64_t)q);
    const unsigned long long mo = *reinterpret_cast<volatile unsigned long long *>(sp.mPar + 2 * (int64_t)q + 1);
    const unsigned long long m = max(mc, max(me, mo));
    if (m < TDNA_INF_BITS) {
        const double T = __longlong_as_double((long long)m) + 2e-5;
        if (T < 1.0) atomicMin(sp.thr + q, (uint32_t)(T * 65536.0) + 2u);
    }
}

template <int KM> __device__ __noinline__ void stream_emit(const StreamParams &sp, int i, int j, uint32_t a0, uint32_t a1, uint32_t lhs, uint32_t den, int min_overlap) {
    Counts cn = unpack_counts<KM>(a0, a1);
    const double d = distance_from_counts<KM>(cn, min_overlap);
    if ((sp.want & TDNA_WANT_EDGES) && d >= 0 && d <= sp.edge_t) {  // Cluster.java:797-804 (NaN fails both tests)
        const unsigned long long pos = atomicAdd(sp.nedge, 1ull);
        if (sp.ei && (long long)pos < sp.edge_cap) { sp.ei[pos] = i; sp.ej[pos] = j; sp.ed[pos] = d; }
    }
    if (!(d < __longlong_as_double((long long)TDNA_INF_BITS))) {  // NaN / +Inf: the host resolves these rows from tdna_row_distances
        if (sp.want & TDNA_WANT_SEED) return;
        atomicOr(sp.flags + i, TDNA_ROW_NONFINITE);
        atomicOr(sp.flags + j, TDNA_ROW_NONFINITE);
        return;
    }
    stream_emit_side(sp, i, j, d, lhs, den);
    stream_emit_side(sp, j, i, d, lhs, den);
}

// The same slow path for callers that enter it CONVERGED (every lane of the warp brings one pair, `active` lanes
// have one): both sides of the pair are handled together and independent memory operations are issued back to back,
// so a warp pays ~3 memory latencies per batch of up to 32 candidates instead of ~6 per candidate.
template <int KM> __device__ __forceinline__ void stream_emit_converged(const StreamParams &sp, bool active, int i, int j, uint32_t a0, uint32_t a1, uint32_t lhs, uint32_t den,
                                                                      int min_overlap) {
    const bool seed = (sp.want & TDNA_WANT_SEED) != 0, want_bm = (sp.want & TDNA_WANT_BEST_MATCH) != 0;
    double d = 0.0;
    uint32_t fri = 0, frj = 0;
    int spi = -1, spj = -1;
    if (active) {  // loads first, the fp64 division overlaps them
        if (want_bm) {
            fri = *reinterpret_cast<volatile uint32_t *>(sp.thr + i);
            frj = *reinterpret_cast<volatile uint32_t *>(sp.thr + j);
            spi = __ldg(sp.species + i);
            spj = __ldg(sp.species + j);
        }
        Counts cn = unpack_counts<KM>(a0, a1);
        d = distance_from_counts<KM>(cn, min_overlap);
        if ((sp.want & TDNA_WANT_EDGES) && d >= 0 && d <= sp.edge_t) {
            const unsigned long long pos = atomicAdd(sp.nedge, 1ull);
            if (sp.ei && (long long)pos < sp.edge_cap) { sp.ei[pos] = i; sp.ej[pos] = j; sp.ed[pos] = d; }
        }
        if (!(d < __longlong_as_double((long long)TDNA_INF_BITS))) {
            if (!seed) { atomicOr(sp.flags + i, TDNA_ROW_NONFINITE); atomicOr(sp.flags + j, TDNA_ROW_NONFINITE); }
            active = false;
        }
    }
    const bool ei = active && want_bm && lhs <= fri * den, ej = active && want_bm && lhs <= frj * den;
    if (!(ei || ej)) return;
    if (!seed) {
        const unsigned long long pos = atomicAdd(sp.ncand, (unsigned long long)(ei + ej));
        if ((long long)(pos + ei + ej) > sp.cap) { *sp.overflow = 1; }
        else {
            unsigned long long w = pos;
            if (ei) { sp.cq[w] = i; sp.cj[w] = j; sp.cd[w] = d; atomicAdd(sp.cnt + i, 1); w++; }
            if (ej) { sp.cq[w] = j; sp.cj[w] = i; sp.cd[w] = d; atomicAdd(sp.cnt + j, 1); }
        }
    }
    // minima (no return value: RED) and the thresholds derived from them
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    const bool ui = ei && spj >= 0, uj = ej && spi >= 0;  // unnamed columns bound nothing
    if (ui) { if (spj == spi) atomicMin(sp.mC + i, bits); atomicMin(sp.mPar + 2 * (int64_t)i + (spj & 1), bits); }
    if (uj) { if (spi == spj) atomicMin(sp.mC + j, bits); atomicMin(sp.mPar + 2 * (int64_t)j + (spi & 1), bits); }
    unsigned long long m_i[3] = {0, 0, 0}, m_j[3] = {0, 0, 0};
    if (ui) {
        m_i[0] = *reinterpret_cast<volatile unsigned long long *>(sp.mC + i);
        m_i[1] = *reinterpret_cast<volatile unsigned long long *>(sp.mPar + 2 * (int64_t)i);
        m_i[2] = *reinterpret_cast<volatile unsigned long long *>(sp.mPar + 2 * (int64_t)i + 1);
    }
    if (uj) {
        m_j[0] = *reinterpret_cast<volatile unsigned long long *>(sp.mC + j);
        m_j[1] = *reinterpret_cast<volatile unsigned long long *>(sp.mPar + 2 * (int64_t)j);
        m_j[2] = *reinterpret_cast<volatile unsigned long long *>(sp.mPar + 2 * (int64_t)j + 1);
    }
    if (ui) {  // fold the own value in: the RED above need not be visible to the loads yet
        if (spj == spi) m_i[0] = min(m_i[0], bits);
        m_i[1 + (spj & 1)] = min(m_i[1 + (spj & 1)], bits);
        const unsigned long long m = max(m_i[0], max(m_i[1], m_i[2]));
        if (m < TDNA_INF_BITS) {
            const double T = __longlong_as_double((long long)m) + 2e-5;
            if (T < 1.0) atomicMin(sp.thr + i, (uint32_t)(T * 65536.0) + 2u);
        }
    }
    if (uj) {
        if (spi == spj) m_j[0] = min(m_j[0], bits);
        m_j[1 + (spi & 1)] = min(m_j[1 + (spi & 1)], bits);
        const unsigned long long m = max(m_j[0], max(m_j[1], m_j[2]));
        if (m < TDNA_INF_BITS) {
            const double T = __longlong_as_double((long long)m) + 2e-5;
            if (T < 1.0) atomicMin(sp.thr + j, (uint32_t)(T * 65536.0) + 2u);
        }
    }
}

#define TDNA_HIST_EMPTY 0xffffffffffffffffull
__device__ __forceinline__ void hist_add(unsigned long long *keys, unsigned long long *cnts, uint32_t mask, int *over, double d, int times) {
    unsigned long long key = (unsigned long long)__double_as_longlong(d);
    if (d != d) key = 0x7ff8000000000000ull;  // one NaN
    // A class has hundreds of millions of pushes and a few dozen distinct distances (10^6 barcodes: 4.95 * 10^8 congeneric pairs, 73
    // values): the lanes that arrive here together and carry the same distance send ONE reduction between them (un-aggregated, the
    // same-address reductions alone cost 0.43 s of that 1.0 s pass).
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    const int total = __reduce_add_sync(peers, times);
    if ((int)(threadIdx.x & 31) != __ffs((int)peers) - 1) return;
    uint32_t h = (uint32_t)(key >> 32) * 0x9e3779b1u ^ (uint32_t)key * 0x85ebca6bu;
    h ^= h >> 15;
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const uint32_t slot = (h + probe) & mask;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(keys + slot);  // the common case: the value is there already
        if (cur == TDNA_HIST_EMPTY) cur = atomicCAS(keys + slot, TDNA_HIST_EMPTY, key);
        if (cur == TDNA_HIST_EMPTY || cur == key) {
            atomicAdd(cnts + slot, (unsigned long long)total);  // no return value: a reduction at the L2
            return;
        }
    }
    *over = 1;
}
// Shared-memory pre-aggregation in front of the global tables.  Hundreds of millions of pushes meet a few dozen addresses: even one
// reduction per warp and distinct value leaves every SM hammering the same L2 lines (measured on 10^6 rows: 0.5 ms per tile that
// holds class pairs).  A small table in shared memory (key: the bits of d, class in bit 63 -- distances are never negative) takes the
// pushes of a tile (tile kernel: one table per epilogue warp) or of a block's share of the queue (resolve kernel) and is flushed
// with one global reduction per distinct value.  false: the table is full -- the caller pushes to the global table directly.
__device__ __forceinline__ bool shist_add(unsigned long long *keys, unsigned int *cnts, uint32_t mask, int k, double d, int times) {
    unsigned long long key = (unsigned long long)__double_as_longlong(d);
    if (d != d) key = 0x7ff8000000000000ull;
    key |= (unsigned long long)k << 63;
    uint32_t h = (uint32_t)(key >> 32) * 0x9e3779b1u ^ (uint32_t)key * 0x85ebca6bu;
    h ^= h >> 15;
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const uint32_t slot = (h + probe) & mask;
        unsigned long long cur = keys[slot];
        if (cur == TDNA_HIST_EMPTY) cur = atomicCAS(keys + slot, TDNA_HIST_EMPTY, key);
        if (cur == TDNA_HIST_EMPTY || cur == key) { atomicAdd(cnts + slot, (unsigned int)times); return true; }
    }
    return false;
}
// every thread that calls this owns the slots slot0, slot0 + stride, ...; the table is empty afterwards
__device__ __forceinline__ void shist_flush(const StreamParams &sp, unsigned long long *keys, unsigned int *cnts, uint32_t size, uint32_t slot0, uint32_t stride) {
    for (uint32_t s = slot0; s < size; s += stride) {
        const unsigned long long key = keys[s];
        if (key == TDNA_HIST_EMPTY) continue;
        const int k = (int)(key >> 63);
        const unsigned int n = cnts[s];
        keys[s] = TDNA_HIST_EMPTY;
        cnts[s] = 0;
        // (no warp aggregation here: the slots of one table hold distinct values)
        const unsigned long long gk = key & 0x7fffffffffffffffull;
        uint32_t h = (uint32_t)(gk >> 32) * 0x9e3779b1u ^ (uint32_t)gk * 0x85ebca6bu;
        h ^= h >> 15;
        bool done = false;
        for (uint32_t probe = 0; probe <= sp.hmask && !done; probe++) {
            const uint32_t slot = (h + probe) & sp.hmask;
            unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(sp.hkey[k] + slot);
            if (cur == TDNA_HIST_EMPTY) cur = atomicCAS(sp.hkey[k] + slot, TDNA_HIST_EMPTY, gk);
            if (cur == TDNA_HIST_EMPTY || cur == gk) { atomicAdd(sp.hcnt[k] + slot, (unsigned long long)n); done = true; }
        }
        if (!done) *sp.hover = 1;
    }
}

// the pushes of one class pair: into the class's histogram when it has one, else into its list
__device__ __forceinline__ void class_push(const StreamParams &sp, int k, double d, int times) {
    if (sp.hkey[k]) { hist_add(sp.hkey[k], sp.hcnt[k], sp.hmask, sp.hover, d, times); return; }
    const float f = (float)d;
    for (int t = 0; t < times; t++) {
        const unsigned long long pos = atomicAdd(sp.ncls + k, 1ull);
        if (sp.cls[k] && (long long)pos < sp.cls_cap[k]) sp.cls[k][pos] = f;
    }
}

// PairwiseDistribution classes of one valid pair i < j (PairwiseDistribution.java:161-203): pushes (float)d
template <int KM> __device__ __noinline__ void class_emit(const StreamParams &sp, int i, int j, uint32_t a0, uint32_t a1, int min_overlap) {
    int times[2] = {0, 0};
    if (sp.want & TDNA_WANT_INTRA) {  // _addIntra: conspecific, both named; once per pair, twice when the full names are equal (:172)
        const int si = sp.species[i];
        if (si >= 0 && si == sp.species[j]) times[0] = 1 + (sp.full[i] == sp.full[j] ? 1 : 0);
    }
    if (sp.want & TDNA_WANT_INTER) {  // _addInter: same genus, epithets differ; exactly one direction has epi[q] < epi[j] (:191-198)
        if (sp.genus[i] == sp.genus[j] && sp.epi[i] != sp.epi[j]) times[1] = 1;
    }
    if (!(times[0] | times[1])) return;
    Counts cn = unpack_counts<KM>(a0, a1);
    const double d = distance_from_counts<KM>(cn, min_overlap);  // the pair is valid: never -1
#pragma unroll
    for (int k = 0; k < 2; k++)
        if (times[k]) class_push(sp, k, d, times[k]);
}

template <int KM, int CN, bool EXTRAS>
__device__ __forceinline__ void stream_epilogue(const TileParams &p, const StreamParams &sp, uint32_t (&acc)[ModeTraits<KM>::NACC][8][CN], int ti, int tj, int tx, int ty) {
    constexpr int NACC = ModeTraits<KM>::NACC, TN = 16 * CN;
    const int n = (int)p.n;
    const int i0 = ti * TM + ty * 8, j0 = tj * TN;
    uint32_t tfi[8], tfj[CN];
#pragma unroll
    for (int r = 0; r < 8; r++) tfi[r] = (i0 + r < n) ? __ldcg(sp.thr + i0 + r) : 0u;
#pragma unroll
    for (int e = 0; e < CN; e++) {
        const int j = j0 + (e >> 2) * PR + 4 * tx + (e & 3);
        tfj[e] = (j < n) ? __ldcg(sp.thr + j) : 0u;
    }
    const uint32_t edge_tf = (EXTRAS && (sp.want & TDNA_WANT_EDGES)) ? sp.edge_tfix : 0u;  // 0: only mism == 0 pairs pass, re-checked in the slow path
    if (EXTRAS && !(sp.want & TDNA_WANT_BEST_MATCH)) {
#pragma unroll
        for (int r = 0; r < 8; r++) tfi[r] = 0u;
#pragma unroll
        for (int e = 0; e < CN; e++) tfj[e] = 0u;
    }
    const bool seed = (sp.want & TDNA_WANT_SEED) != 0;  // diagonal-block pre-pass: thresholds only
    uint32_t bad_r = 0, bad_c = 0;  // invalid pairs per row / column of this thread, one nibble each (<= 8)
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int i = i0 + r;
#pragma unroll
        for (int e = 0; e < CN; e++) {
            const int j = j0 + (e >> 2) * PR + 4 * tx + (e & 3);
            const uint32_t a0 = acc[0][r][e];
            const uint32_t s = a0 >> 16, c1 = a0 & 0xffffu;
            if (j <= i || j >= n) {
                if (j == i && i < n && (int)s >= p.min_overlap && !seed) atomicOr(sp.flags + i, TDNA_ROW_SELF_VALID);
                continue;
            }
            if ((int)s < p.min_overlap) {  // -1.0 in the reference: not a candidate of anything
                if (!seed) {
                    bad_r += 1u << (4 * r);
                    bad_c += 1u << (4 * e);
                }
                continue;
            }
            uint32_t a1 = 0, mism, den;
            bool force = false;
            if constexpr (KM == KM_K2P) {
                a1 = acc[1][r][e];
                const uint32_t ts = a1 & 0xffffu, tv = a1 >> 16;
                mism = ts + tv;  // d_K2P >= P + Q
                den = c1;
                force = (2 * ts + tv >= c1) || (2 * tv >= c1);  // w1 <= 0 or w2 <= 0 or n == 0: NaN / Inf must be flagged
            } else if constexpr (KM == KM_TRANS) {
                mism = c1;
                den = s;
            } else {
                mism = s - c1;
                den = s;
            }
            const uint32_t lhs = mism << 16;
            bool pass = lhs <= tfi[r] * den || lhs <= tfj[e] * den || force;
            if constexpr (EXTRAS) pass |= lhs <= edge_tf * den;
            if (pass) stream_emit<KM>(sp, i, j, a0, a1, lhs, den, p.min_overlap);
        }
    }
    if (EXTRAS && (sp.want & (TDNA_WANT_INTRA | TDNA_WANT_INTER))) {
        // class pairs only live in tiles whose row and column panels share a species or a genus id range
        // (a handful of near-diagonal tiles when the list is sorted by name)
        bool may = false;
        for (int a = 0; a < TM / PR; a++)
            for (int b = 0; b < TN / PR; b++) {
                const int pi = ti * (TM / PR) + a, pj = tj * (TN / PR) + b;
                if ((int64_t)pi * PR >= n || (int64_t)pj * PR >= n) continue;
                const int4 ri = sp.panel_range[pi], rj = sp.panel_range[pj];
                may |= (ri.x <= rj.y && rj.x <= ri.y) || (ri.z <= rj.w && rj.z <= ri.w);
            }
        if (may) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int i = i0 + r;
#pragma unroll
                for (int e = 0; e < CN; e++) {
                    const int j = j0 + (e >> 2) * PR + 4 * tx + (e & 3);
                    const uint32_t a0 = acc[0][r][e];
                    if (j <= i || j >= n || (int)(a0 >> 16) < p.min_overlap) continue;
                    const bool hit = ((sp.want & TDNA_WANT_INTRA) && sp.species[i] == sp.species[j]) ||
                                     ((sp.want & TDNA_WANT_INTER) && sp.genus[i] == sp.genus[j]);
                    if (!hit) continue;
                    uint32_t a1 = 0;
                    if constexpr (NACC == 2) a1 = acc[1][r][e];
                    class_emit<KM>(sp, i, j, a0, a1, p.min_overlap);
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, (bad_r | bad_c) != 0)) {  // rare on clean data
#pragma unroll
        for (int r = 0; r < 8; r++) {
            int v = (int)((bad_r >> (4 * r)) & 15u);
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);  // the 16 tx lanes of this row
            if (tx == 0 && v) atomicAdd(sp.inval + i0 + r, v);
        }
#pragma unroll
        for (int e = 0; e < CN; e++) {
            int v = (int)((bad_c >> (4 * e)) & 15u);
            v += __shfl_xor_sync(0xffffffffu, v, 16);  // the two ty rows of this warp
            if ((threadIdx.x & 16) == 0 && v) atomicAdd(sp.inval + j0 + (e >> 2) * PR + 4 * tx + (e & 3), v);
        }
    }
}

template <int KM, int CN, int EPI>
__global__ void __launch_bounds__(TILE_THREADS, TileCfg<KM, CN>::MIN_CTAS) count_tile_kernel(const __grid_constant__ TileParams p, const __grid_constant__ StreamParams sp) {
    using Cfg = TileCfg<KM, CN>;
    constexpr int P = Cfg::P, NACC = Cfg::NACC, TN = Cfg::TN;
    extern __shared__ __align__(128) uint32_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NST * Cfg::STAGE_WORDS);
    uint64_t *empty = full + NST;
    int64_t *s_prefix = reinterpret_cast<int64_t *>(empty + NST);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int nchunks = (p.W + KC - 1) / KC;
    // tiles of this CTA: blockIdx.x, +gridDim.x, ...
    const int64_t my_tiles = (p.num_tiles > blockIdx.x) ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * nchunks;

    const bool prefix_in_smem = p.sched != SCHED_RECT && p.nrt + 1 <= PREFIX_SMEM;
    if (prefix_in_smem)
        for (int i = tid; i <= p.nrt; i += TILE_THREADS) s_prefix[i] = p.tile_prefix[i];
    const int64_t *prefix = prefix_in_smem ? s_prefix : p.tile_prefix;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ===== producer warp: one elected lane feeds the ring =====
        if ((tid & 31) == 0) {
            for (int64_t it = 0; it < total; it++) {
                const int slot = (int)(it % NST);
                if (it >= NST) mbar_wait_backoff(&empty[slot], (uint32_t)(((it / NST) - 1) & 1));
                const int64_t t = blockIdx.x + (it / nchunks) * gridDim.x;
                const int c = (int)(it % nchunks);
                int ti, tj;
                tile_coords<TN>(p, prefix, t, ti, tj);
                const int k0 = c * KC, kc = min(KC, p.W - k0);
                uint32_t *xs = smem + slot * Cfg::STAGE_WORDS, *ys = xs + Cfg::XW;
                const uint32_t bytes = (uint32_t)kc * PR * 4;
                mbar_expect_tx(&full[slot], bytes * P * (TM / PR + TN / PR));
                const int64_t xpanel = (p.row_begin + (int64_t)ti * TM) / PR, ypanel = ((int64_t)tj * TN) / PR;
#pragma unroll
                for (int h = 0; h < TM / PR; h++)
                    for (int pl = 0; pl < P; pl++)
                        bulk_g2s(xs + ((pl * (TM / PR) + h) * KC) * PR, p.planes + (((xpanel + h) * P + pl) * p.W + k0) * PR, bytes, &full[slot]);
#pragma unroll
                for (int h = 0; h < TN / PR; h++)
                    for (int pl = 0; pl < P; pl++)
                        bulk_g2s(ys + ((pl * (TN / PR) + h) * KC) * PR, p.planes + (((ypanel + h) * P + pl) * p.W + k0) * PR, bytes, &full[slot]);
            }
        }
        return;
    }

    // ===== consumer warps =====
    const int tx = tid & 15, ty = tid >> 4;
    uint32_t acc[NACC][8][CN];
    const int xh = ty >> 3, xr = (ty & 7) * 8;  // panel half and first row within it

    for (int64_t it = 0; it < total; it++) {
        const int c = (int)(it % nchunks);
        const int slot = (int)(it % NST);
        if (c == 0) {
#pragma unroll
            for (int a = 0; a < NACC; a++)
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int cc = 0; cc < CN; cc++) acc[a][r][cc] = 0;
        }
        mbar_wait(&full[slot], (uint32_t)((it / NST) & 1));
        const uint32_t *xs = smem + slot * Cfg::STAGE_WORDS, *ys = xs + Cfg::XW;
        const int kc = min(KC, p.W - c * KC);
        for (int k = 0; k < kc; k++) {
            uint32_t x[P][8], y[P][CN];
#pragma unroll
            for (int pl = 0; pl < P; pl++) {
                const uint4 *xp = reinterpret_cast<const uint4 *>(xs + ((pl * (TM / PR) + xh) * KC + k) * PR + xr);
                uint4 a = xp[0], b = xp[1];
                x[pl][0] = a.x; x[pl][1] = a.y; x[pl][2] = a.z; x[pl][3] = a.w;
                x[pl][4] = b.x; x[pl][5] = b.y; x[pl][6] = b.z; x[pl][7] = b.w;
#pragma unroll
                for (int h = 0; h < CN / 4; h++) {
                    uint4 v = *reinterpret_cast<const uint4 *>(ys + ((pl * (TN / PR) + h) * KC + k) * PR + 4 * tx);
                    y[pl][4 * h + 0] = v.x; y[pl][4 * h + 1] = v.y; y[pl][4 * h + 2] = v.z; y[pl][4 * h + 3] = v.w;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int cc = 0; cc < CN; cc++) {
                    uint32_t xv[P], yv[P];
#pragma unroll
                    for (int pl = 0; pl < P; pl++) { xv[pl] = x[pl][r]; yv[pl] = y[pl][cc]; }
                    uint32_t a1 = 0;
                    if constexpr (NACC == 2) {
                        accumulate_word<KM>(xv, yv, acc[0][r][cc], acc[1][r][cc]);
                    } else {
                        accumulate_word<KM>(xv, yv, acc[0][r][cc], a1);
                    }
                }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[slot]);  // this warp is done reading the stage
        if (c == nchunks - 1) {
            const int64_t t = blockIdx.x + (it / nchunks) * gridDim.x;
            int ti, tj;
            tile_coords<TN>(p, prefix, t, ti, tj);
            if constexpr (EPI == EPI_STREAM || EPI == EPI_STREAM_X) stream_epilogue<KM, CN, EPI == EPI_STREAM_X>(p, sp, acc, ti, tj, tx, ty);
            else matrix_epilogue<KM, CN, EPI == EPI_MATRIX_COUNTS>(p, acc, ti, tj, tx, ty);
        }
    }
}

// ------------------------------------------------------------------------------------------
// one query row straight from the planes (SortedSequenceList.sortAgainst's N getPairwise calls,
// tdna_pair with n_cols == 1).  thread <-> column j; the query's words are broadcast loads.
// ------------------------------------------------------------------------------------------
template <int KM>
__global__ void __launch_bounds__(256) row_from_planes_kernel(const uint32_t *__restrict__ planes, int W, int64_t q, int64_t j_begin,
                                                              int64_t j_end, int min_overlap, double *__restrict__ d_out,
                                                              int32_t *__restrict__ shared_out, tdna_counts *__restrict__ counts_out) {
    int64_t j = j_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= j_end) return;
    Counts c = pair_counts_global<KM>(planes, q, j, W);
    if (d_out) d_out[j - j_begin] = distance_from_counts<KM>(c, min_overlap);
    if (shared_out) shared_out[j - j_begin] = c.shared;
    if (counts_out) {
        tdna_counts o = {c.shared, c.c1, c.c2, c.c3};
        counts_out[j - j_begin] = o;
    }
}

// a sequence that is not a row of the alignment against every row (QuerySequence.java:125-145 -> SortedSequenceList.sortAgainst with
// a foreign query): the query's planes are packed like a one-row alignment (row 0 of qplanes); thread <-> column j
template <int KM>
__global__ void __launch_bounds__(256) query_from_planes_kernel(const uint32_t *__restrict__ qplanes, const uint32_t *__restrict__ planes, int W,
                                                                int64_t n, int min_overlap, double *__restrict__ d_out,
                                                                int32_t *__restrict__ shared_out, tdna_counts *__restrict__ counts_out) {
    constexpr int P = ModeTraits<KM>::P;
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t a0 = 0, a1 = 0;
    for (int k = 0; k < W; k++) {
        uint32_t x[P], y[P];
#pragma unroll
        for (int p = 0; p < P; p++) {
            x[p] = __ldg(qplanes + plane_index(0, p, k, P, W));  // the same word in every lane: a broadcast load
            y[p] = planes[plane_index(j, p, k, P, W)];
        }
        accumulate_word<KM>(x, y, a0, a1);
    }
    Counts c = unpack_counts<KM>(a0, a1);
    if (d_out) d_out[j] = distance_from_counts<KM>(c, min_overlap);
    if (shared_out) shared_out[j] = c.shared;
    if (counts_out) {
        tdna_counts o = {c.shared, c.c1, c.c2, c.c3};
        counts_out[j] = o;
    }
}

// an explicit list of pairs straight from the planes (tdna_pairs): thread <-> pair
template <int KM>
__global__ void __launch_bounds__(256) pairs_from_planes_kernel(const uint32_t *__restrict__ planes, int W, const int32_t *__restrict__ ii,
                                                                const int32_t *__restrict__ jj, int64_t m, int min_overlap,
                                                                double *__restrict__ d_out, tdna_counts *__restrict__ counts_out) {
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    Counts c = pair_counts_global<KM>(planes, ii[k], jj[k], W);
    if (d_out) d_out[k] = distance_from_counts<KM>(c, min_overlap);
    if (counts_out) {
        tdna_counts o = {c.shared, c.c1, c.c2, c.c3};
        counts_out[k] = o;
    }
}

// ------------------------------------------------------------------------------------------
// per-query summary over one band of the distance matrix (HBM/L2-bound: three sweeps of the
// query's row, the 2nd and 3rd from L2).  One CTA per query row.
// ------------------------------------------------------------------------------------------
struct Key {
    unsigned long long hi, lo;  // hi = bits of d (valid finite d >= 0: bit order == numeric order); lo = rank<<32 | j
};
#define TDNA_KEY_NONE 0xffffffffffffffffull
__device__ __forceinline__ bool key_less(const Key &a, const Key &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ Key key_min(const Key &a, const Key &b) { return key_less(b, a) ? b : a; }
__device__ __forceinline__ Key key_shfl_xor(const Key &k, int m) {
    Key r;
    r.hi = __shfl_xor_sync(0xffffffffu, k.hi, m);
    r.lo = __shfl_xor_sync(0xffffffffu, k.lo, m);
    return r;
}
// all threads get the block-wide min; scratch: >= 8 Keys
__device__ __forceinline__ Key block_key_min(Key k, Key *scratch) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) k = key_min(k, key_shfl_xor(k, m));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = k;
    __syncthreads();
    Key r = scratch[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = key_min(r, scratch[w]);
    return r;
}
__device__ __forceinline__ int block_sum(int v, int *scratch) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += scratch[w];
    return r;
}
__device__ __forceinline__ int block_min_i(int v, int *scratch) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, m));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = scratch[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = min(r, scratch[w]);
    return r;
}
__device__ __forceinline__ int block_max_i(int v, int *scratch) { return -block_min_i(-v, scratch); }

#define TDNA_NEAR 2e-5  // twice Settings' 1e-5 quantum (Settings.java:48): see DESIGN.md "near ties"

__device__ __forceinline__ bool is_near(double d, double ref) {
    double diff = fabs(d - ref);
    return d != ref && diff < TDNA_NEAR;
}

// (key, species) pairs: the two smallest keys among NAMED columns with DISTINCT species.  Mergeable,
// so one sweep yields both the best named match and the first entry of another species (the end of
// BlockAnalysis' leading run) without knowing species(best) in advance.
struct Top2 {
    Key k1, k2;
    int s1, s2;
};
__device__ __forceinline__ void top2_insert(Top2 &t, const Key &k, int s) {
    if (k.hi == TDNA_KEY_NONE) return;
    if (s == t.s1) { if (key_less(k, t.k1)) t.k1 = k; }
    else if (key_less(k, t.k1)) { t.k2 = t.k1; t.s2 = t.s1; t.k1 = k; t.s1 = s; }
    else if (s == t.s2) { if (key_less(k, t.k2)) t.k2 = k; }
    else if (key_less(k, t.k2)) { t.k2 = k; t.s2 = s; }
}
__device__ __forceinline__ Top2 top2_shfl_xor(const Top2 &t, int m) {
    Top2 r;
    r.k1 = key_shfl_xor(t.k1, m);
    r.k2 = key_shfl_xor(t.k2, m);
    r.s1 = __shfl_xor_sync(0xffffffffu, t.s1, m);
    r.s2 = __shfl_xor_sync(0xffffffffu, t.s2, m);
    return r;
}

constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_UNROLL = 4;  // independent 16-byte loads in flight per thread (64 KB per SM)

struct RowAgg {  // everything sweep 1 produces
    Key kb, kc, ka;
    Top2 t2;
    int nvalid, nonfinite;
};
__device__ __forceinline__ void agg_merge(RowAgg &a, const RowAgg &o) {
    a.kb = key_min(a.kb, o.kb);
    a.kc = key_min(a.kc, o.kc);
    a.ka = key_min(a.ka, o.ka);
    top2_insert(a.t2, o.t2.k1, o.t2.s1);
    top2_insert(a.t2, o.t2.k2, o.t2.s2);
    a.nvalid += o.nvalid;
    a.nonfinite |= o.nonfinite;
}
__device__ __forceinline__ RowAgg agg_shfl_xor(const RowAgg &a, int m) {
    RowAgg r;
    r.kb = key_shfl_xor(a.kb, m);
    r.kc = key_shfl_xor(a.kc, m);
    r.ka = key_shfl_xor(a.ka, m);
    r.t2 = top2_shfl_xor(a.t2, m);
    r.nvalid = __shfl_xor_sync(0xffffffffu, a.nvalid, m);
    r.nonfinite = __shfl_xor_sync(0xffffffffu, a.nonfinite, m);
    return r;
}

// ---- the two sweeps of a query's row, element by element (shared by the dense row kernel and the
// ---- candidate-list finaliser so that both state the same semantics) ---------------------------
__device__ __forceinline__ RowAgg agg_identity() {
    Key none = {TDNA_KEY_NONE, TDNA_KEY_NONE};
    RowAgg g = {none, none, none, {none, none, -2, -2}, 0, 0};
    return g;
}

// sweep 1: arg-min keys.  (j, d, spj) is one column of query q's row; the caller skipped j == q and d < 0.
__device__ __forceinline__ void sweep1_elem(RowAgg &g, int spq, bool qnamed, int64_t j, double d, int spj) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    g.nvalid++;
    if (!(d < INF)) { g.nonfinite = 1; if (d != d) return; }
    const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    const bool consp = qnamed && spj == spq;  // SortedSequenceComparator: conspecific-to-query first (:236-255)
    // fast reject: larger than every running minimum this column could still replace (NONE = all ones)
    unsigned long long thr = g.kb.hi;
    if (spj >= 0) {
        thr = max(thr, g.t2.k2.hi);
        if (qnamed) thr = max(thr, consp ? g.kc.hi : g.ka.hi);
    }
    if (bits > thr) return;
    Key k = {bits, ((unsigned long long)(consp ? 0 : 1) << 32) | (unsigned long long)j};
    g.kb = key_min(g.kb, k);
    if (spj >= 0) {
        top2_insert(g.t2, k, spj);
        if (qnamed) {
            Key k2 = {bits, (unsigned long long)j};
            if (consp) g.kc = key_min(g.kc, k2); else g.ka = key_min(g.ka, k2);
        }
    }
}

struct RowRefs {  // what sweep 2 measures against
    Key kb, kc, ka, ko;
    bool has_b, has_c, has_a, has_o;
    int jb, spb;
    double db, dc, da, dother, dmax;
};
__device__ __forceinline__ RowRefs make_refs(const RowAgg &g, const int32_t *__restrict__ species) {
    RowRefs f;
    f.kb = g.kb; f.kc = g.kc; f.ka = g.ka;
    f.has_b = f.kb.hi != TDNA_KEY_NONE; f.has_c = f.kc.hi != TDNA_KEY_NONE; f.has_a = f.ka.hi != TDNA_KEY_NONE;
    f.jb = f.has_b ? (int)(f.kb.lo & 0xffffffffu) : -1;
    f.db = f.has_b ? __longlong_as_double((long long)f.kb.hi) : -1.0;
    f.dc = f.has_c ? __longlong_as_double((long long)f.kc.hi) : -1.0;
    f.da = f.has_a ? __longlong_as_double((long long)f.ka.hi) : -1.0;
    f.spb = f.has_b ? species[f.jb] : -1;
    // first named entry of a species other than species(best): if best is named it IS t2.k1
    f.ko = (f.has_b && f.spb >= 0) ? g.t2.k2 : g.t2.k1;
    f.has_o = f.has_b && f.ko.hi != TDNA_KEY_NONE;
    f.dother = f.has_o ? __longlong_as_double((long long)f.ko.hi) : -1.0;
    // nothing beyond the largest reference distance plus the near-tie window is ever counted -- unless no
    // other species exists at all: then every column of species(best) belongs to the leading run
    f.dmax = f.has_o ? fmax(fmax(f.db, f.dc), fmax(f.da, f.dother)) + TDNA_NEAR : __longlong_as_double(0x7ff0000000000000LL);
    return f;
}

struct Sweep2 {  // tie class, near-tie windows, leading same-species run
    int v[10];   // ties, tie_unnamed, near_b, near_c, near_a, near_o, un_c, un_a, un_o, block
    int smin, smax;
};
__device__ __forceinline__ Sweep2 sweep2_identity() {
    Sweep2 s;
#pragma unroll
    for (int i = 0; i < 10; i++) s.v[i] = 0;
    s.smin = 0x7fffffff;
    s.smax = -1;
    return s;
}
__device__ __forceinline__ void sweep2_elem(Sweep2 &s, const RowRefs &f, int spq, bool qnamed, int64_t j, double d, int spj) {
    if (!(d <= f.dmax)) return;
    if (d == f.db && j != f.jb) {
        s.v[0]++;
        if (spj < 0) s.v[1]++; else { s.smin = min(s.smin, spj); s.smax = max(s.smax, spj); }
    }
    s.v[2] += is_near(d, f.db);
    if (f.has_c) { s.v[3] += is_near(d, f.dc); s.v[6] += (spj < 0 && d == f.dc); }
    if (f.has_a) { s.v[4] += is_near(d, f.da); s.v[7] += (spj < 0 && d == f.da); }
    if (f.has_o) { s.v[5] += is_near(d, f.dother); s.v[8] += (spj < 0 && d == f.dother); }
    if (j != f.jb && spj >= 0 && spj == f.spb) {
        const bool consp = qnamed && spj == spq;
        Key k = {(unsigned long long)__double_as_longlong(d), ((unsigned long long)(consp ? 0 : 1) << 32) | (unsigned long long)j};
        if (!f.has_o || key_less(k, f.ko)) s.v[9]++;
    }
}
__device__ __forceinline__ void sweep2_warp_reduce(Sweep2 &s, int first = 16) {  // first = half the lanes that hold one row
#pragma unroll
    for (int m = first; m >= 1; m >>= 1) {
#pragma unroll
        for (int i = 0; i < 10; i++) s.v[i] += __shfl_xor_sync(0xffffffffu, s.v[i], m);
        s.smin = min(s.smin, __shfl_xor_sync(0xffffffffu, s.smin, m));
        s.smax = max(s.smax, __shfl_xor_sync(0xffffffffu, s.smax, m));
    }
}
__device__ __forceinline__ tdna_row_result make_row_result(const RowRefs &f, const Sweep2 &s, int nvalid, int flags) {
    tdna_row_result r;
    r.d_best = f.db; r.d_con = f.dc; r.d_allo = f.da; r.d_other = f.dother;
    r.best = f.jb;
    r.con = f.has_c ? (int)(f.kc.lo & 0xffffffffu) : -1;
    r.allo = f.has_a ? (int)(f.ka.lo & 0xffffffffu) : -1;
    r.other = f.has_o ? (int)(f.ko.lo & 0xffffffffu) : -1;
    r.shared_con = 0; r.shared_allo = 0;  // filled by fill_shared_kernel
    r.tie_count = s.v[0]; r.tie_unnamed = s.v[1]; r.tie_species_min = s.smin; r.tie_species_max = s.smax;
    r.near_best = s.v[2]; r.near_con = s.v[3]; r.near_allo = s.v[4]; r.near_other = s.v[5];
    r.unnamed_at_con = s.v[6]; r.unnamed_at_allo = s.v[7]; r.unnamed_at_other = s.v[8];
    r.block_count = s.v[9];
    r.n_valid = nvalid;
    r.flags = flags;
    r.reserved = 0;
    r.reserved2 = 0;
    return r;
}

// Dense variant: two sweeps per query row of a materialised band, the second one out of L2: the grid
// is persistent with ONE 512-thread CTA per SM so that the rows in flight (148 x N x 8 B) stay
// L2-resident between the sweeps.
__global__ void __launch_bounds__(RS_THREADS, 1) row_summary_kernel(const double *__restrict__ mat, int64_t ld, int64_t n, int64_t row_begin,
                                                                   int64_t rows, const int32_t *__restrict__ species,
                                                                   tdna_row_result *__restrict__ out) {
    __shared__ RowAgg s_agg[RS_WARPS];
    __shared__ RowAgg s_final;
    __shared__ int s_cnt[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t stride = (int64_t)RS_THREADS * 2;

    for (int64_t qb = blockIdx.x; qb < rows; qb += gridDim.x) {
        const int64_t q = row_begin + qb;
        const double *row = mat + qb * ld;
        const int spq = species[q];
        const bool qnamed = spq >= 0;

        // ---- sweep 1 (HBM): arg-min keys -------------------------------------------------------
        RowAgg g = agg_identity();
        for (int64_t base = (int64_t)threadIdx.x * 2; base < n; base += stride * RS_UNROLL) {
            double2 v[RS_UNROLL];
            int2 sp2[RS_UNROLL];
#pragma unroll
            for (int u = 0; u < RS_UNROLL; u++) {
                const int64_t j0 = base + u * stride;
                if (j0 < n) {  // ld is a multiple of 16: the pair is in bounds and 16-byte aligned
                    v[u] = *reinterpret_cast<const double2 *>(row + j0);
                    sp2[u] = (j0 + 1 < n) ? *reinterpret_cast<const int2 *>(species + j0) : make_int2(species[j0], -1);
                }
            }
#pragma unroll
            for (int u = 0; u < RS_UNROLL; u++) {
                const int64_t j0 = base + u * stride;
                if (j0 >= n) continue;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int64_t j = j0 + e;
                    const double d = e ? v[u].y : v[u].x;
                    const int spj = e ? sp2[u].y : sp2[u].x;
                    if (j >= n || j == q || d < 0) continue;
                    sweep1_elem(g, spq, qnamed, j, d, spj);
                }
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) { RowAgg o = agg_shfl_xor(g, m); agg_merge(g, o); }
        __syncthreads();  // the previous row's readers of the shared arrays are done
        if (lane == 0) s_agg[warp] = g;
        if (threadIdx.x < 16) s_cnt[threadIdx.x] = (threadIdx.x == 10) ? 0x7fffffff : (threadIdx.x == 11 ? -1 : 0);
        __syncthreads();
        if (warp == 0) {  // one warp folds the per-warp partials
            g = lane < RS_WARPS ? s_agg[lane] : agg_identity();
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) { RowAgg o = agg_shfl_xor(g, m); agg_merge(g, o); }
            if (lane == 0) s_final = g;
        }
        __syncthreads();
        g = s_final;
        const RowRefs f = make_refs(g, species);

        // ---- sweep 2 (L2): tie class, near-tie windows, leading same-species run ----------------
        if (f.has_b) {
            Sweep2 s2 = sweep2_identity();
            for (int64_t base = (int64_t)threadIdx.x * 2; base < n; base += stride * RS_UNROLL) {
                double2 v[RS_UNROLL];
                int2 sp2[RS_UNROLL];
#pragma unroll
                for (int u = 0; u < RS_UNROLL; u++) {
                    const int64_t j0 = base + u * stride;
                    if (j0 < n) {
                        v[u] = *reinterpret_cast<const double2 *>(row + j0);
                        sp2[u] = (j0 + 1 < n) ? *reinterpret_cast<const int2 *>(species + j0) : make_int2(species[j0], -1);
                    }
                }
#pragma unroll
                for (int u = 0; u < RS_UNROLL; u++) {
                    const int64_t j0 = base + u * stride;
                    if (j0 >= n) continue;
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int64_t j = j0 + e;
                        const double d = e ? v[u].y : v[u].x;
                        const int spj = e ? sp2[u].y : sp2[u].x;
                        if (j >= n || j == q || d < 0) continue;
                        sweep2_elem(s2, f, spq, qnamed, j, d, spj);
                    }
                }
            }
            sweep2_warp_reduce(s2);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 10; i++)
                    if (s2.v[i]) atomicAdd(&s_cnt[i], s2.v[i]);
                if (s2.smin != 0x7fffffff) atomicMin(&s_cnt[10], s2.smin);
                if (s2.smax >= 0) atomicMax(&s_cnt[11], s2.smax);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            Sweep2 s2;
#pragma unroll
            for (int i = 0; i < 10; i++) s2.v[i] = s_cnt[i];
            s2.smin = s_cnt[10];
            s2.smax = s_cnt[11];
            double dself = row[q];
            out[q] = make_row_result(f, s2, g.nvalid, (!(dself < 0) ? TDNA_ROW_SELF_VALID : 0) | (g.nonfinite ? TDNA_ROW_NONFINITE : 0));
        }
    }
}

// ------------------------------------------------------------------------------------------
// label-derived tables (tdna_alignment_set_labels), built on the device: no host round trip
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t label_hash(int32_t v) {
    uint32_t x = (uint32_t)v * 0x9e3779b1u;
    return x ^ (x >> 15);
}
// open-addressing table of the named species ids with their member counts; keys start at -1 (0xff fill), counts at 0
__global__ void __launch_bounds__(256) species_count_kernel(const int32_t *__restrict__ species, int64_t n, int32_t *__restrict__ keys,
                                                            int32_t *__restrict__ counts, uint32_t mask) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int32_t v = species[q];
    if (v < 0) return;
    for (uint32_t h = label_hash(v) & mask;; h = (h + 1) & mask) {
        const int32_t old = atomicCAS(keys + h, -1, v);
        if (old == -1 || old == v) { atomicAdd(counts + h, 1); return; }
    }
}
// needc[q] != 0: q is named and another row carries the same species id
__global__ void __launch_bounds__(256) needc_kernel(const int32_t *__restrict__ species, int64_t n, const int32_t *__restrict__ keys,
                                                    const int32_t *__restrict__ counts, uint32_t mask, uint8_t *__restrict__ needc) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int32_t v = species[q];
    uint8_t r = 0;
    if (v >= 0)
        for (uint32_t h = label_hash(v) & mask;; h = (h + 1) & mask) {
            const int32_t k = keys[h];
            if (k == v) { r = counts[h] > 1; break; }
            if (k == -1) break;
        }
    needc[q] = r;
}
// per 64-row panel: (min, max) named species id, (min, max) genus id; one warp per panel
__global__ void __launch_bounds__(32) panel_range_kernel(const int32_t *__restrict__ species, const int32_t *__restrict__ genus, int64_t n,
                                                         int4 *__restrict__ out) {
    const int64_t pnl = blockIdx.x;
    int4 r = make_int4(0x7fffffff, -1, 0x7fffffff, (int)0x80000000);
    for (int x = threadIdx.x; x < PR; x += 32) {
        const int64_t i = pnl * PR + x;
        if (i >= n) continue;
        const int sv = species[i], gv = genus[i];
        if (sv >= 0) { r.x = min(r.x, sv); r.y = max(r.y, sv); }
        r.z = min(r.z, gv);
        r.w = max(r.w, gv);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        r.x = min(r.x, __shfl_xor_sync(0xffffffffu, r.x, m));
        r.y = max(r.y, __shfl_xor_sync(0xffffffffu, r.y, m));
        r.z = min(r.z, __shfl_xor_sync(0xffffffffu, r.z, m));
        r.w = max(r.w, __shfl_xor_sync(0xffffffffu, r.w, m));
    }
    if (threadIdx.x == 0) out[pnl] = r;
}

// ------------------------------------------------------------------------------------------
// streaming Best Match, after the tile pass: state init, candidate CSR, exact per-row summaries
// ------------------------------------------------------------------------------------------
// needc[q] != 0: q is named and another row of its species exists (labels only; set by the host side)
__global__ void __launch_bounds__(256) stream_init_kernel(StreamParams sp, int64_t n, const uint8_t *__restrict__ needc) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) { *sp.ncand = 0ull; *sp.overflow = 0; sp.ncls[0] = 0ull; sp.ncls[1] = 0ull; *sp.nedge = 0ull; *sp.qn = 0ull; }
    if (q >= n) return;
    sp.thr[q] = TDNA_THR_ALL;
    sp.mC[q] = needc[q] ? TDNA_INF_BITS : 0ull;
    sp.mPar[2 * q] = TDNA_INF_BITS;
    sp.mPar[2 * q + 1] = TDNA_INF_BITS;
    sp.inval[q] = 0;
    sp.flags[q] = 0;
    sp.cnt[q] = 0;
}

// thresholds from the (possibly all-reduced) minima: between the diagonal phase and the bulk phase
__global__ void __launch_bounds__(256) stream_thr_kernel(StreamParams sp, int64_t n) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const unsigned long long m = max(sp.mC[q], max(sp.mPar[2 * q], sp.mPar[2 * q + 1]));
    uint32_t tf = TDNA_THR_ALL;
    if (m < TDNA_INF_BITS) {
        const double T = __longlong_as_double((long long)m) + 2e-5;
        if (T < 1.0) tf = (uint32_t)(T * 65536.0) + 2u;
    }
    sp.thr[q] = min(sp.thr[q], tf);
}

// exclusive prefix sum of cnt[0..n) into ptr[0..n], three launches: per-block sums (SCAN_BLOCK items per CTA), a one-CTA scan of
// the block sums, per-block scan + offset.  (One CTA over all n took 48 us at n = 50,000 and a millisecond at 10^6 rows.)
constexpr int SCAN_BLOCK = 4096, SCAN_PER_THREAD = SCAN_BLOCK / 256;
__global__ void __launch_bounds__(256) scan_block_sums_kernel(const int *__restrict__ cnt, int64_t n, long long *__restrict__ bsum) {
    __shared__ long long s_w[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
    long long sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {  // coalesced: consecutive threads, consecutive items
        const int64_t i = base + k * 256 + threadIdx.x;
        if (i < n) sum += cnt[i];
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        bsum[blockIdx.x] = t;
    }
}
// one CTA: exclusive scan of the nb block sums in place; total -> ptr_n[0]
__global__ void __launch_bounds__(1024) scan_block_offsets_kernel(long long *__restrict__ bsum, int64_t nb, long long *__restrict__ ptr_n) {
    __shared__ long long s_part[1024];
    const int t = threadIdx.x;
    const int64_t per = (nb + 1023) / 1024, b = t * per, e = min(nb, b + per);
    long long sum = 0;
    for (int64_t i = b; i < e; i++) sum += bsum[i];
    s_part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele over the 1024 partials
        long long v = t >= off ? s_part[t - off] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    long long run = s_part[t] - sum;
    for (int64_t i = b; i < e; i++) { const long long v = bsum[i]; bsum[i] = run; run += v; }
    if (t == 1023) *ptr_n = s_part[1023];
}
// thread t of a block owns items [t * 16, t * 16 + 16) of the block's 4096
__global__ void __launch_bounds__(256) scan_apply_kernel(const int *__restrict__ cnt, int64_t n, const long long *__restrict__ boff,
                                                         long long *__restrict__ ptr) {
    __shared__ long long s_w[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD];
    long long sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD / 4; k++) {  // 16-byte loads; cnt is 16-byte aligned and base a multiple of 16 items
        int4 q = make_int4(0, 0, 0, 0);
        const int64_t i = base + 4 * k;
        if (i + 3 < n) q = *reinterpret_cast<const int4 *>(cnt + i);
        else {
            if (i < n) q.x = cnt[i];
            if (i + 1 < n) q.y = cnt[i + 1];
            if (i + 2 < n) q.z = cnt[i + 2];
        }
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
        sum += (long long)q.x + q.y + q.z + q.w;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    long long run = boff[blockIdx.x] + incl - sum;
    for (int k = 0; k < w; k++) run += s_w[k];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        if (base + k < n) ptr[base + k] = run;
        run += v[k];
    }
}

// scatter the unordered records into per-row lists; cursor[] starts at zero
__global__ void __launch_bounds__(256) stream_scatter_kernel(const int32_t *__restrict__ cq, const int32_t *__restrict__ cj,
                                                             const double *__restrict__ cd, long long ncand,
                                                             const long long *__restrict__ ptr, int *__restrict__ cursor,
                                                             int32_t *__restrict__ lj, double *__restrict__ ld) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < ncand; r += (long long)gridDim.x * blockDim.x) {
        const int q = cq[r];
        const long long slot = ptr[q] + atomicAdd(cursor + q, 1);
        lj[slot] = cj[r];
        ld[slot] = cd[r];
    }
}
__global__ void __launch_bounds__(256) count_rows_kernel(const int32_t *__restrict__ cq, long long ncand, int *__restrict__ cnt) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < ncand; r += (long long)gridDim.x * blockDim.x)
        atomicAdd(cnt + cq[r], 1);
}

// ---- packed exchange format (tdna_best_match_part_packed / _merge_gathered) ----------------------------------
__global__ void __launch_bounds__(256) pack_records_kernel(const int32_t *__restrict__ cq, const int32_t *__restrict__ cj, const double *__restrict__ cd,
                                                           long long ncand, long long *__restrict__ rec) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < ncand; r += (long long)gridDim.x * blockDim.x) {
        rec[2 * r] = ((long long)cq[r] << 32) | (long long)(uint32_t)cj[r];
        rec[2 * r + 1] = __double_as_longlong(cd[r]);
    }
}
__global__ void __launch_bounds__(256) pack_state_kernel(const int *__restrict__ inval, const int *__restrict__ flags, int64_t n, long long ncand,
                                                         long long *__restrict__ state, long long *__restrict__ count) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) *count = ncand;
    if (q >= n) return;
    const int f = flags[q];
    state[q] = (long long)(uint32_t)inval[q] | ((long long)(f & TDNA_ROW_SELF_VALID ? 1 : 0) << 32) | ((long long)(f & TDNA_ROW_NONFINITE ? 1 : 0) << 48);
}
__global__ void __launch_bounds__(256) unpack_state_kernel(const long long *__restrict__ state, int64_t n, int *__restrict__ inval, int *__restrict__ flags) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const long long s = state[q];
    inval[q] = (int)(s & 0xffffffffll);
    flags[q] = (((s >> 32) & 0xffff) ? TDNA_ROW_SELF_VALID : 0) | (((s >> 48) & 0xffff) ? TDNA_ROW_NONFINITE : 0);
}
// routed exchange: the records bucketed by the part that OWNS their row (row / rows_per_part); segment g at rec + 2 * g * seg_stride.
// Positions come from shared-memory counters per block and ONE global atomic per (block, destination).
constexpr int ROUTE_MAX_PARTS = 64;
__global__ void __launch_bounds__(256) route_records_kernel(const int32_t *__restrict__ cq, const int32_t *__restrict__ cj, const double *__restrict__ cd,
                                                            long long ncand, int64_t rows_per_part, int nparts, long long seg_stride,
                                                            long long *__restrict__ rec, unsigned long long *__restrict__ counts, int *__restrict__ overflow) {
    __shared__ int s_cnt[ROUTE_MAX_PARTS];
    __shared__ long long s_base[ROUTE_MAX_PARTS];
    const long long chunks = (ncand + 255) / 256;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        if (threadIdx.x < nparts) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const long long r = ch * 256 + threadIdx.x;
        int dest = -1, local = 0, q = 0;
        if (r < ncand) {
            q = cq[r];
            dest = (int)(q / rows_per_part);
            local = atomicAdd(&s_cnt[dest], 1);
        }
        __syncthreads();
        if (threadIdx.x < nparts && s_cnt[threadIdx.x] > 0)
            s_base[threadIdx.x] = (long long)atomicAdd(counts + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
        if (dest >= 0) {
            const long long pos = s_base[dest] + local;
            if (pos < seg_stride) {
                long long *o = rec + 2 * ((long long)dest * seg_stride + pos);
                o[0] = ((long long)q << 32) | (long long)(uint32_t)cj[r];
                o[1] = __double_as_longlong(cd[r]);
            } else *overflow = 1;
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) pack_state_only_kernel(const int *__restrict__ inval, const int *__restrict__ flags, int64_t n, long long *__restrict__ state) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int f = flags[q];
    state[q] = (long long)(uint32_t)inval[q] | ((long long)(f & TDNA_ROW_SELF_VALID ? 1 : 0) << 32) | ((long long)(f & TDNA_ROW_NONFINITE ? 1 : 0) << 48);
}
// gathered segments: segment g holds seg_count[g] records at rec + 2 * g * seg_stride
__global__ void __launch_bounds__(256) count_rows_seg_kernel(const long long *__restrict__ rec, int nseg, long long seg_stride,
                                                             const long long *__restrict__ seg_count, int *__restrict__ cnt) {
    const long long total = (long long)nseg * seg_stride;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
        if (r % seg_stride >= seg_count[r / seg_stride]) continue;
        atomicAdd(cnt + (int)(rec[2 * r] >> 32), 1);
    }
}
__global__ void __launch_bounds__(256) scatter_seg_kernel(const long long *__restrict__ rec, int nseg, long long seg_stride,
                                                          const long long *__restrict__ seg_count, const long long *__restrict__ ptr,
                                                          int *__restrict__ cursor, int32_t *__restrict__ lj, double *__restrict__ ld) {
    const long long total = (long long)nseg * seg_stride;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
        if (r % seg_stride >= seg_count[r / seg_stride]) continue;
        const long long key = rec[2 * r];
        const int q = (int)(key >> 32);
        const long long slot = ptr[q] + atomicAdd(cursor + q, 1);
        lj[slot] = (int32_t)(key & 0xffffffffll);
        ld[slot] = __longlong_as_double(rec[2 * r + 1]);
    }
}

// ---- in-library multi-GPU exchange (tdna_comm_init): nothing here needs a host round trip ---------------------------------
// What rank s sends to rank g is ONE segment of segw 64-bit words:
//   [0]            the record count (the atomic cursor while routing), [1] unused
//   [2 ..)         seg_stride 16-byte records (row << 32 | column, fp64 bits of the distance) of rows that g owns
//   then           rpp words: s's per-row state (invalid-pair count | self-valid << 32 | non-finite << 48) of g's rows
//   then           MULTI_STATE_EXTRA status words of s (the same in every segment s sends)
// so one grouped send / recv moves the records, the state and the status; the receiver sums state and status over the segments.
// The record count of the pass is read from device memory: the whole step is enqueued without waiting for the GPU.
constexpr int MULTI_STATE_EXTRA = 8;  // status: 0 segment overflow, 1 candidate / queue overflow, 2..3 class pushes, 4 edges, 5 candidates
__host__ __device__ inline long long multi_seg_words(long long seg_stride, long long rpp) { return 2 * (seg_stride + 1) + rpp + MULTI_STATE_EXTRA; }

__global__ void __launch_bounds__(64) multi_zero_kernel(long long *__restrict__ send, int nparts, long long segw, long long *__restrict__ status) {
    if (threadIdx.x < nparts) { send[threadIdx.x * segw] = 0; send[threadIdx.x * segw + 1] = 0; }
    if (threadIdx.x == 0) status[0] = 0;
}
__global__ void __launch_bounds__(256) route_records_dev_kernel(StreamParams sp, int64_t rows_per_part, int nparts, long long seg_stride, long long segw,
                                                                long long *__restrict__ send, long long *__restrict__ seg_overflow) {
    __shared__ int s_cnt[ROUTE_MAX_PARTS];
    __shared__ long long s_base[ROUTE_MAX_PARTS];
    if (*sp.overflow) return;  // the record list has holes: the status word tells every rank to retry
    const long long ncand = min((long long)*sp.ncand, sp.cap);
    const long long chunks = (ncand + 255) / 256;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        if (threadIdx.x < nparts) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const long long r = ch * 256 + threadIdx.x;
        int dest = -1, local = 0, q = 0;
        if (r < ncand) {
            q = sp.cq[r];
            dest = (int)(q / rows_per_part);
            local = atomicAdd(&s_cnt[dest], 1);
        }
        __syncthreads();
        if (threadIdx.x < nparts && s_cnt[threadIdx.x] > 0)
            s_base[threadIdx.x] = (long long)atomicAdd(reinterpret_cast<unsigned long long *>(send + threadIdx.x * segw), (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
        if (dest >= 0) {
            const long long pos = s_base[dest] + local;
            if (pos < seg_stride) {
                long long *o = send + dest * segw + 2 + 2 * pos;
                o[0] = ((long long)q << 32) | (long long)(uint32_t)sp.cj[r];
                o[1] = __double_as_longlong(sp.cd[r]);
            } else *seg_overflow = 1;
        }
        __syncthreads();
    }
}
// the state slices and the status words of every segment; status[0] (segment overflow) was set by the routing kernel.  send == nullptr:
// a pass without Best Match rows exchanges nothing -- only the status block (all-reduced by the caller) is written
__global__ void __launch_bounds__(256) multi_pack_kernel(StreamParams sp, int64_t n, int64_t rpp, int nparts, long long seg_stride, long long segw,
                                                         long long *__restrict__ send, long long *__restrict__ status) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long st[MULTI_STATE_EXTRA] = {status[0], *sp.overflow ? 1 : 0, (long long)sp.ncls[0], (long long)sp.ncls[1], (long long)*sp.nedge, (long long)*sp.ncand, 0, 0};
    if (q < MULTI_STATE_EXTRA) status[q] = st[q];
    if (!send) return;
    if (q < (int64_t)nparts * MULTI_STATE_EXTRA) send[(q / MULTI_STATE_EXTRA) * segw + 2 * (seg_stride + 1) + rpp + q % MULTI_STATE_EXTRA] = st[q % MULTI_STATE_EXTRA];
    if (q >= (int64_t)nparts * rpp) return;
    long long v = 0;
    if (q < n) {
        const int f = sp.flags[q];
        v = (long long)(uint32_t)sp.inval[q] | ((long long)(f & TDNA_ROW_SELF_VALID ? 1 : 0) << 32) | ((long long)(f & TDNA_ROW_NONFINITE ? 1 : 0) << 48);
    }
    send[(q / rpp) * segw + 2 * (seg_stride + 1) + q % rpp] = v;
}
// receiver: the own rows' state summed over the segments, list counters cleared, the status summed
__global__ void __launch_bounds__(256) multi_prep_kernel(const long long *__restrict__ recv, int nparts, long long seg_stride, long long segw, int64_t rpp,
                                                         int64_t q0, int64_t q1, int *__restrict__ inval, int *__restrict__ flags, int *__restrict__ cnt,
                                                         int *__restrict__ cursor, long long *__restrict__ status) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < MULTI_STATE_EXTRA) {
        long long t = 0;
        for (int g = 0; g < nparts; g++) t += recv[g * segw + 2 * (seg_stride + 1) + rpp + r];
        status[r] = t;
    }
    if (q0 + r >= q1) return;
    long long sum = 0;
    for (int g = 0; g < nparts; g++) sum += recv[g * segw + 2 * (seg_stride + 1) + r];
    inval[q0 + r] = (int)(sum & 0xffffffffll);
    flags[q0 + r] = (((sum >> 32) & 0xffff) ? TDNA_ROW_SELF_VALID : 0) | (((sum >> 48) & 0xffff) ? TDNA_ROW_NONFINITE : 0);
    cnt[q0 + r] = 0;
    cursor[q0 + r] = 0;
}
__global__ void __launch_bounds__(256) count_rows_hseg_kernel(const long long *__restrict__ recv, int nseg, long long seg_stride, long long segw, int *__restrict__ cnt) {
    const long long total = (long long)nseg * seg_stride;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
        const long long g = r / seg_stride, k = r % seg_stride;
        const long long *seg = recv + g * segw;
        if (k >= min(seg[0], seg_stride)) continue;
        atomicAdd(cnt + (int)(seg[2 + 2 * k] >> 32), 1);
    }
}
__global__ void __launch_bounds__(256) scatter_hseg_kernel(const long long *__restrict__ recv, int nseg, long long seg_stride, long long segw,
                                                           const long long *__restrict__ ptr, int *__restrict__ cursor, int32_t *__restrict__ lj,
                                                           double *__restrict__ ld) {
    const long long total = (long long)nseg * seg_stride;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
        const long long g = r / seg_stride, k = r % seg_stride;
        const long long *seg = recv + g * segw;
        if (k >= min(seg[0], seg_stride)) continue;
        const long long key = seg[2 + 2 * k];
        const int q = (int)(key >> 32);
        const long long slot = ptr[q] + atomicAdd(cursor + q, 1);
        lj[slot] = (int32_t)(key & 0xffffffffll);
        ld[slot] = __longlong_as_double(seg[3 + 2 * k]);
    }
}
// one CTA: exclusive scan of cnt[0..m) into ptr[0..m] (a rank's own rows: a few thousand)
__global__ void __launch_bounds__(1024) scan_small_kernel(const int *__restrict__ cnt, int64_t m, long long *__restrict__ ptr) {
    __shared__ long long s_part[1024];
    const int t = threadIdx.x;
    const int64_t per = (m + 1023) / 1024, b = t * per, e = min(m, b + per);
    long long sum = 0;
    for (int64_t i = b; i < e; i++) sum += cnt[i];
    s_part[t] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        long long v = t >= off ? s_part[t - off] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    long long run = s_part[t] - sum;
    for (int64_t i = b; i < e; i++) { ptr[i] = run; run += cnt[i]; }
    if (t == 1023) ptr[m] = s_part[1023];
}
// histogram table -> dense (d, count) entries; *n counts every non-empty slot (the caller stores the first cap)
__global__ void __launch_bounds__(256) hist_compact_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ cnts, uint32_t size,
                                                           tdna_hist_entry *__restrict__ out, long long cap, unsigned long long *__restrict__ n,
                                                           unsigned long long *__restrict__ pushes) {
    unsigned long long local = 0;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < size; s += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[s];
        if (k == TDNA_HIST_EMPTY) continue;
        const unsigned long long pos = atomicAdd(n, 1ull);
        if (out && (long long)pos < cap) { out[pos].d = __longlong_as_double((long long)k); out[pos].count = (int64_t)cnts[s]; }
        local += cnts[s];
    }
    if (local) atomicAdd(pushes, local);
}
__global__ void hist_meta_kernel(const int *__restrict__ over, const unsigned long long *__restrict__ n, long long *__restrict__ dst) {
    if (threadIdx.x == 0) { dst[0] = *over ? 1 : 0; dst[1] = (long long)*n; }
}
// entries of other ranks' histograms into this rank's table (multi-GPU merge)
__global__ void __launch_bounds__(256) hist_insert_kernel(const tdna_hist_entry *__restrict__ in, long long m, unsigned long long *__restrict__ keys,
                                                          unsigned long long *__restrict__ cnts, uint32_t mask, int *__restrict__ over) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x) {
        if (in[r].count <= 0) continue;
        unsigned long long key = (unsigned long long)__double_as_longlong(in[r].d);
        uint32_t h = (uint32_t)(key >> 32) * 0x9e3779b1u ^ (uint32_t)key * 0x85ebca6bu;
        h ^= h >> 15;
        bool done = false;
        for (uint32_t probe = 0; probe <= mask && !done; probe++) {
            const uint32_t slot = (h + probe) & mask;
            unsigned long long cur = atomicCAS(keys + slot, TDNA_HIST_EMPTY, key);
            if (cur == TDNA_HIST_EMPTY || cur == key) { atomicAdd(cnts + slot, (unsigned long long)in[r].count); done = true; }
        }
        if (!done) *over = 1;
    }
}

// rows past the end of the list in the last rank's slice: zero records, so that the all-gathered buffer is deterministic
__global__ void __launch_bounds__(256) zero_rows_kernel(tdna_row_result *__restrict__ res, int64_t q0, int64_t q1) {
    const int64_t words = (q1 - q0) * (int64_t)(sizeof(tdna_row_result) / 8);
    long long *p = reinterpret_cast<long long *>(res + q0);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < words; k += (int64_t)gridDim.x * blockDim.x) p[k] = 0;
}

// G lanes per query row (32 / G rows per warp): the two sweeps over the row's candidate list (complete below its threshold).
// Lists average ~15 entries on barcode data, so eight lanes per row keep the lanes busy and the butterflies short; rows of large
// species carry hundreds of entries and get a whole warp (the host picks G from the mean list length).
template <int G>
__global__ void __launch_bounds__(256) stream_finalize_kernel(const long long *__restrict__ ptr, const int32_t *__restrict__ lj,
                                                              const double *__restrict__ ld, const int *__restrict__ inval,
                                                              const int *__restrict__ flags, const int32_t *__restrict__ species,
                                                              int64_t n, tdna_row_result *__restrict__ out, int64_t q_begin, int64_t q_end) {
    const int gl = threadIdx.x & (G - 1);
    const int64_t q = q_begin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;  // rows [q_begin, q_end): a part's own rows, or all
    const bool live = q < q_end;  // dead groups run along with empty lists: the shuffles below are warp-wide
    const int64_t qq = live ? q : q_begin;
    const int spq = species[qq];
    const bool qnamed = spq >= 0;
    const long long b = live ? ptr[qq] : 0, e = live ? ptr[qq + 1] : 0;
    RowAgg g = agg_identity();
    for (long long k = b + gl; k < e; k += G) {
        const int j = lj[k];
        sweep1_elem(g, spq, qnamed, j, ld[k], species[j]);
    }
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) { RowAgg o = agg_shfl_xor(g, m); agg_merge(g, o); }
    const RowRefs f = make_refs(g, species);
    Sweep2 s2 = sweep2_identity();
    if (f.has_b) {
        for (long long k = b + gl; k < e; k += G) {
            const int j = lj[k];
            sweep2_elem(s2, f, spq, qnamed, j, ld[k], species[j]);
        }
    }
    sweep2_warp_reduce(s2, G / 2);
    // every valid column is counted through inval[] (the list holds only the near ones)
    if (live && gl == 0) out[q] = make_row_result(f, s2, (int)(n - 1) - inval[q], flags[q]);
}

// getSharedLength(query, first_con / first_allo) for the listing (BestMatch.java:372,386)
template <int KM>
__global__ void fill_shared_kernel(const uint32_t *__restrict__ planes, int W, int64_t row_begin, int64_t row_end,
                                   tdna_row_result *__restrict__ res) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t q = row_begin + (idx >> 1);
    if (q >= row_end) return;
    int which = (int)(idx & 1);
    int j = which ? res[q].allo : res[q].con;
    int sh = 0;
    if (j >= 0) sh = pair_counts_global<KM>(planes, q, j, W).shared;
    if (which) res[q].shared_allo = sh; else res[q].shared_con = sh;
}

// ------------------------------------------------------------------------------------------
// PairwiseDistribution classes and Cluster edges: compaction sweeps over a band (HBM-bound)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) class_distances_kernel(const double *__restrict__ mat, int64_t ld, int64_t n, int64_t row_begin,
                                                              int64_t rows, int klass, const int32_t *__restrict__ species,
                                                              const int32_t *__restrict__ genus, const int32_t *__restrict__ epi,
                                                              const int32_t *__restrict__ full, float *__restrict__ out, int64_t cap,
                                                              unsigned long long *__restrict__ counter, unsigned long long *__restrict__ hkey,
                                                              unsigned long long *__restrict__ hcnt, uint32_t hmask, int *__restrict__ hover) {
    const int64_t qb = blockIdx.x;  // grid.x = band rows, grid.y = column slices
    if (qb >= rows) return;
    const int64_t q = row_begin + qb;
    const int spq = species[q], gq = genus[q], eq = epi[q], fq = full[q];
    for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.y * blockDim.x) {
        if (j == q) continue;
        bool in;
        if (klass == TDNA_PD_INTRA)  // _addIntra (PairwiseDistribution.java:161-177)
            in = spq >= 0 && species[j] == spq && !(full[j] < fq);
        else                         // _addInter (:183-203)
            in = genus[j] == gq && epi[j] != eq && eq < epi[j];
        if (!in) continue;
        const double dv = mat[qb * ld + j];
        float f = (float)dv;
        if (f == -1.0f) continue;  // distances_push drops -1 (:79)
        if (hkey) { hist_add(hkey, hcnt, hmask, hover, dv, 1); continue; }
        unsigned long long pos = atomicAdd(counter, 1ull);
        if (out && (int64_t)pos < cap) out[pos] = f;
    }
}

__global__ void __launch_bounds__(256) edges_kernel(const double *__restrict__ mat, int64_t ld, int64_t n, int64_t row_begin, int64_t rows,
                                                    double lo, double t, int32_t *__restrict__ ii, int32_t *__restrict__ jj, double *__restrict__ dd,
                                                    int64_t cap, unsigned long long *__restrict__ counter) {
    const int64_t qb = blockIdx.x;
    if (qb >= rows) return;
    const int64_t q = row_begin + qb;
    for (int64_t j = q + 1 + (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.y * blockDim.x) {
        double d = mat[qb * ld + j];
        if (d >= lo && d <= t) {  // lo = 0: Cluster.java:797-804; any lo: FindDistances.java:227 (NaN fails both tests, -1 is a value)
            unsigned long long pos = atomicAdd(counter, 1ull);
            if (ii && (int64_t)pos < cap) { ii[pos] = (int32_t)q; jj[pos] = (int32_t)j; dd[pos] = d; }
        }
    }
}

// ---- per-row extremes among a row's congeners (ExtremePairwise.java:99-170, DistanceAnalysis.java:183-230) -----------------------
// Both modules sort the whole list against every query (N x sortAgainst) and then only look at the query's conspecifics and at its
// congeneric, allospecific entries.  One CTA per query row: the rows of the query's genus (member lists by genus, ascending list
// position) are evaluated straight from the bit-planes, the valid named entries of the two classes are sorted in shared memory by
// (class, distance, list position) -- the order SortedSequenceList gives them when its comparator is consistent -- and one thread
// reads the record off the sorted entries.  O(sum of genus sizes squared) pair evaluations instead of O(N^2).
constexpr int EXT_CAP = 4096;      // entries per query (conspecifics + congeners); larger genera are left to the host
constexpr int EXT_THREADS = 128;
template <int KM>
__global__ void __launch_bounds__(EXT_THREADS) row_extremes_kernel(const uint32_t *__restrict__ planes, int W, int64_t n, int min_overlap,
                                                                   const int32_t *__restrict__ species, const int32_t *__restrict__ genus,
                                                                   const int32_t *__restrict__ gslot, const int64_t *__restrict__ gstart,
                                                                   const int32_t *__restrict__ gmem, tdna_extreme_result *__restrict__ out) {
    extern __shared__ unsigned long long ext_smem[];
    unsigned long long *key = ext_smem;                                   // [EXT_CAP] (class << 63) | fp64 bits of d
    int32_t *col = reinterpret_cast<int32_t *>(ext_smem + EXT_CAP);       // [EXT_CAP]
    __shared__ int s_count, s_flags;
    const int64_t q = blockIdx.x;
    if (q >= n) return;
    tdna_extreme_result r = {};
    r.d_max_intra = r.d_min_inter = -1.0;
    r.max_intra = r.min_inter = -1;
    const int spq = species[q];
    if (spq < 0) {  // "Unable to identify species name" (ExtremePairwise.java:109-112)
        if (threadIdx.x == 0) { r.flags = TDNA_EXT_UNNAMED; out[q] = r; }
        return;
    }
    if (threadIdx.x == 0) { s_count = 0; s_flags = 0; }
    __syncthreads();
    const int64_t m0 = gstart[gslot[q]], m1 = gstart[gslot[q] + 1];
    int myflags = 0;
    if (m1 - m0 > EXT_CAP) myflags |= TDNA_EXT_TOO_MANY;
    else
        for (int64_t m = m0 + threadIdx.x; m < m1; m += EXT_THREADS) {
            const int j = gmem[m];
            const Counts c = pair_counts_global<KM>(planes, q, j, W);
            const double d = distance_from_counts<KM>(c, min_overlap);
            if (j == q) { if (!(d < 0.0)) myflags |= TDNA_ROW_SELF_VALID; continue; }
            const int spj = species[j];
            if (spj < 0 || d < 0.0) continue;  // nameless entries are skipped by both modules; -1 never enters the sorted list
            if (!(d < __longlong_as_double((long long)TDNA_INF_BITS))) { myflags |= TDNA_ROW_NONFINITE; continue; }
            const int slot = atomicAdd(&s_count, 1);
            key[slot] = (unsigned long long)__double_as_longlong(d) | (spj == spq ? 0ull : 1ull << 63);
            col[slot] = j;
        }
    if (myflags) atomicOr(&s_flags, myflags);
    __syncthreads();
    const int cnt = s_count;
    int padded = 1;
    while (padded < cnt) padded <<= 1;
    for (int k = cnt + threadIdx.x; k < padded; k += EXT_THREADS) { key[k] = ~0ull; col[k] = 0x7fffffff; }
    __syncthreads();
    for (int size = 2; size <= padded; size <<= 1)  // bitonic sort by (key, column)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (padded >> 1); t += EXT_THREADS) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long ka = key[lo], kb = key[hi];
                const int ca = col[lo], cb = col[hi];
                const bool gt = ka > kb || (ka == kb && ca > cb);
                if (gt == up) { key[lo] = kb; key[hi] = ka; col[lo] = cb; col[hi] = ca; }
            }
            __syncthreads();
        }
    if (threadIdx.x != 0) return;
    r.flags = s_flags;
    const double near = 2e-5;  // twice Settings' quantum: inside it the comparator's ties are not transitive
    int k = 0;
    for (; k < cnt && !(key[k] >> 63); k++) {}  // [0, k): conspecifics, ascending
    r.n_intra = k;
    r.n_inter = cnt - k;
    if (k > 0) {  // the LAST valid conspecific of the sorted list: the largest distance, and of equal distances the last in list order
        const double dmax = __longlong_as_double((long long)key[k - 1]);
        r.d_max_intra = dmax;
        r.max_intra = col[k - 1];
        for (int x = k - 2; x >= 0; x--) {
            const double d = __longlong_as_double((long long)key[x]);
            if (!(dmax - d < near)) break;
            if (d != dmax) r.near_intra++;
        }
        r.shared_intra = pair_counts_global<KM>(planes, q, r.max_intra, W).shared;
    }
    if (cnt > k) {  // the FIRST valid congeneric, allospecific entry; the sum in the order the sorted list holds them
        const double dmin = __longlong_as_double((long long)(key[k] & 0x7fffffffffffffffull));
        r.d_min_inter = dmin;
        r.min_inter = col[k];
        double sum = 0.0, prev = dmin;
        for (int x = k; x < cnt; x++) {
            const double d = __longlong_as_double((long long)(key[x] & 0x7fffffffffffffffull));
            sum = __dadd_rn(sum, d);
            if (d != dmin && d - dmin < near) r.near_inter++;
            if (d != prev && d - prev < near) r.near_order++;
            prev = d;
        }
        r.sum_inter = sum;
        r.shared_inter = pair_counts_global<KM>(planes, q, r.min_inter, W).shared;
    }
    out[q] = r;
}

// SpeciesDetail.getSequencesWithValidConspecificsCount's inner test (SpeciesDetail.java:78-95):
// shared >= minOverlap  <=>  the distance is not the -1.0 sentinel.
__global__ void __launch_bounds__(256) valid_consp_kernel(const double *__restrict__ mat, int64_t ld, int64_t n, int64_t row_begin,
                                                          const int32_t *__restrict__ species, uint8_t *__restrict__ has_valid) {
    __shared__ int iscratch[8];
    const int64_t q = row_begin + blockIdx.x;
    const int spq = species[q];
    int any = 0;
    if (spq >= 0)
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x)
            if (j != q && species[j] == spq && mat[(int64_t)blockIdx.x * ld + j] != -1.0) any = 1;
    any = block_sum(any, iscratch);
    if (threadIdx.x == 0) has_valid[q] = any ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// integer-pipe micro-benchmarks (roofline denominators for the count kernel)
// ------------------------------------------------------------------------------------------
template <int WHICH> __global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t b = seed ^ 0x5bd1e995u, c = seed * 7u + 3u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (WHICH == 0) a[i] = __popc(a[i]) + b;          // POPC (+ one IADD that keeps the value alive)
                else a[i] = (a[i] & b) | (a[i] ^ c);              // one LOP3 per result
            }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace tdna
