// Count kernel (b): the per-pair counts as a dense int8 contraction of one-hot matrices on the 5th-gen
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).  p-distance with ambiguity codes
// allowed, for alignments whose symbols are all in {A C G T - _ ?} ("clean": the alignment carries no
// IUPAC ambiguity letter; anything else runs on the popc kernel).
//
// One int32 accumulator per pair carries BOTH counts, from FIVE K-dimensions per site (one per valid symbol
// A C G T '-'):   a-side (row x):  u = b*1 + a*e_x     b-side (column y):  v = d*1 + c*e_y     ('_', '?': zero vector)
// with integers |entries| <= 127 chosen such that  a*c = -(W-1)  and  a*d + b*c + 5*b*d = W  for a power of two W > L
// (256, 1024 and 4096 have solutions), so that
//     u . v = W - (W-1)*[x == y]   for valid x, y, 0 otherwise, and summed over the sites
//     D(x,y) = W*shared - (W-1)*ident = ident + W*(shared - ident),   ident = D & (W-1), mism = D >> log2 W.
// shared = sites where both are a base or '-' (Sequence.getSharedLength, Sequence.java:1429-1455); ident = sites
// where both are the same base or both '-' (Sequence.identical, :1262-1302, restricted to the clean alphabet).
// D < 2^31 for L < 4096.  (The search for a, b, c, d is find_code() in tdna_api.cu.)
//
// Operand memory: per 128-row (a) / 256-row (b) panel and per K-chunk of 128 bytes, one contiguous
// block already in the shared-memory image tcgen05 wants (K-major, no swizzle: 8-row x 16-byte core
// matrices, [k16][row group][row][16 B]); a block is ONE cp.async.bulk.  CTA = 1 TMA producer warp,
// 1 MMA issuer warp (one elected lane), 4 epilogue warps (TMEM -> registers -> the same integer
// candidate filter as the popc kernel's stream epilogue).  Two accumulator stages of 256 columns.
#pragma once
#include <cuda.h>

#include "tdna_kernels.cuh"

namespace tdna {
namespace tc {

constexpr int M = 128, N = 256, KB = 128;          // tile rows, tile columns, K bytes per row per stage
constexpr int NSTAGE = 4;
constexpr int A_STAGE = M * KB, B_STAGE = N * KB;  // 16 KB + 32 KB
constexpr int STAGE = A_STAGE + B_STAGE;
constexpr int EPI_WARPS = 16;                      // four per TMEM lane quarter
constexpr int EPI_COLS = N / (EPI_WARPS / 4);      // columns per epilogue warp
constexpr int THREADS = 64 + EPI_WARPS * 32;       // warp 0 producer, warp 1 MMA issuer, warps 2.. epilogue
constexpr int DIMS = 5;                            // K-dimensions per site
constexpr size_t SMEM = (size_t)NSTAGE * STAGE + 256 /*barriers*/ + 16 * 64 * 4 /*column thresholds, per epilogue warp*/ + 1024 /*alignment*/;

struct Params {
    const uint8_t *A;   // [a-panel][chunk][16 KB]
    const uint8_t *B;   // [b-panel][chunk][32 KB]
    int nchunks;        // K-chunks per row
    int n;
    int min_overlap;
    uint32_t W;         // the accumulator's radix, a power of two
    int wshift;         // log2(W)
    int nrt;            // 128-row tiles
    int nct;            // 256-column tiles
    int band;           // 0: seed pass, tile t = (row tile t, column tile t / 2); else row tiles per band of the whole triangle
    int nbands;
    const int64_t *prefix;  // bulk phase: [nbands + 1] tile slots before band b
    int64_t num_tiles, t_offset, t_stride;
    tdna_counts *counts;    // verification mode: n x n counts (row-major), else null
    int extras;             // stream mode: classes / edges wanted (c_sp.want beyond Best Match)
    int hints;              // L2 cache-policy hints on the operand loads (experiments: TDNA_TC_HINTS)
    int piece;              // bytes per 1-D bulk copy (divides 16 KB); 0 = tensor-map (2-D tiled TMA) loads
};

// Tile slot -> (row tile, column tile); false = an empty slot.  The bulk phase walks BANDS of `band` row tiles,
// column tile outer / row tile inner, so that the ~148 tiles in flight at any time share ~16 a-panels and ~9
// b-panels: every operand block is fetched from HBM about once per band and otherwise hits L2.
__device__ __forceinline__ bool tile_coords(const Params &p, const int64_t *prefix, int64_t t, int &ti, int &tj) {
    t = p.t_offset + t * p.t_stride;
    if (p.band == 0) {
        ti = (int)t;
        tj = ti >> 1;
        return ti < p.nrt && tj < p.nct;
    }
    int lo = 0, hi = p.nbands;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid;
    }
    const int64_t r = t - prefix[lo];
    const int ti0 = lo * p.band;
    ti = ti0 + (int)(r % p.band);
    tj = (ti0 >> 1) + (int)(r / p.band);
    return ti < p.nrt && tj < p.nct && tj >= (ti >> 1);
}

// ---- operand encoding -------------------------------------------------------------------------
struct EncodeValues {
    int a, b, c, d;
};
// one thread = one 16-byte unit of one row; unit index order (panel, chunk, k16, row group, row) = memory order
template <bool BSIDE>
__global__ void __launch_bounds__(256) encode_kernel(const uint8_t *__restrict__ sym, const int32_t *__restrict__ len, int64_t n, int64_t npad,
                                                     int64_t sym_stride, int L, int nchunks, EncodeValues v, uint8_t *__restrict__ out,
                                                     int *__restrict__ not_clean) {
    constexpr int ROWS = BSIDE ? N : M;
    const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t units = npad * nchunks * (KB / 16);
    if (unit >= units) return;
    const int r = (int)(unit % 8);
    const int rg = (int)((unit / 8) % (ROWS / 8));
    const int k16 = (int)((unit / (8 * (ROWS / 8))) % (KB / 16));
    const int64_t pc = unit / (8 * (ROWS / 8) * (KB / 16));
    const int c = (int)(pc % nchunks);
    const int64_t panel = pc / nchunks;
    const int64_t row = panel * ROWS + rg * 8 + r;
    uint32_t w[4] = {0, 0, 0, 0};
    if (row < n) {
        const int l = len[row];
        const uint8_t *src = sym + row * sym_stride;
        const int kk0 = c * KB + k16 * 16;
        bool bad = false;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int kk = kk0 + b, site = kk / DIMS, dim = kk % DIMS;
            int val = 0;
            if (site < L && site < l) {
                const uint8_t ch = __ldg(src + site);
                int code;  // 0..4 = A C G T '-', 5 = '_' / '?', 6 = anything else
                switch (ch) {
                    case 'A': code = 0; break;
                    case 'C': code = 1; break;
                    case 'G': code = 2; break;
                    case 'T': code = 3; break;
                    case '-': code = 4; break;
                    case '_': case '?': code = 5; break;
                    default: code = 6; break;
                }
                bad |= code == 6;
                if (code < 5) val = BSIDE ? v.d + (dim == code ? v.c : 0) : v.b + (dim == code ? v.a : 0);
            }
            w[b >> 2] |= (uint32_t)(val & 0xff) << (8 * (b & 3));
        }
        if (bad) atomicOr(not_clean, 1);
    }
    reinterpret_cast<uint4 *>(out)[unit] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- tcgen05 / TMEM primitives ---------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // K-major, SWIZZLE_NONE: start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48)
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::i8: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major both, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {  // arrives on bar when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 2-D tiled TMA load (UTMALDG): box (256 uint32, rows) of the [.., 1 KB] view of an operand array
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// same with an L2 cache-policy hint (the a-panels of a band are re-read for every column tile: keep them)
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_hint(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) { mbar_wait_backoff(bar, parity); }

// ---- the tile kernel ---------------------------------------------------------------------------
template <bool COUNTS_MODE>
__global__ void __launch_bounds__(THREADS, 1) tile_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NSTAGE * STAGE);
    uint64_t *empty = full + NSTAGE;
    uint64_t *tfull = empty + NSTAGE;   // [2] accumulator stage ready for the epilogue
    uint64_t *tempty = tfull + 2;       // [2] accumulator stage drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    uint32_t *col_thr = reinterpret_cast<uint32_t *>(smem + (size_t)NSTAGE * STAGE + 256);  // [EPI_WARPS][EPI_COLS]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t my_tiles = (p.num_tiles > blockIdx.x) ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 256-column accumulator stages)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int64_t it = 0;
            for (int64_t k = 0; k < my_tiles; k++) {
                int ti, tj;
                if (!tile_coords(p, p.prefix, blockIdx.x + k * gridDim.x, ti, tj)) continue;
                const uint8_t *ga = p.A + (size_t)ti * p.nchunks * A_STAGE;
                const uint8_t *gb = p.B + (size_t)tj * p.nchunks * B_STAGE;
                for (int c = 0; c < p.nchunks; c++, it++) {
                    const int slot = (int)(it % NSTAGE);
                    if (it >= NSTAGE) mbar_wait(&empty[slot], (uint32_t)(((it / NSTAGE) - 1) & 1));  // the ring IS the bottleneck here: no back-off
                    uint8_t *sa = smem + (size_t)slot * STAGE, *sb = sa + A_STAGE;
                    mbar_expect_tx(&full[slot], STAGE);
                    if (p.piece == 0) {  // one block = 16 (a) / 32 (b) rows of 1 KB
                        if (p.hints) {
                            tma_load_2d_hint(sa, &mapA, 0, (int)(((int64_t)ti * p.nchunks + c) * (A_STAGE / 1024)), &full[slot], p.hints == 2 ? L2_EVICT_NORMAL : L2_EVICT_LAST);
                            tma_load_2d_hint(sb, &mapB, 0, (int)(((int64_t)tj * p.nchunks + c) * (B_STAGE / 1024)), &full[slot], p.hints == 3 ? L2_EVICT_FIRST : (p.hints == 2 ? L2_EVICT_LAST : L2_EVICT_NORMAL));
                        } else {
                            tma_load_2d(sa, &mapA, 0, (int)(((int64_t)ti * p.nchunks + c) * (A_STAGE / 1024)), &full[slot]);
                            tma_load_2d(sb, &mapB, 0, (int)(((int64_t)tj * p.nchunks + c) * (B_STAGE / 1024)), &full[slot]);
                        }
                        continue;
                    }
                    for (int o = 0; o < A_STAGE; o += p.piece) bulk_g2s(sa + o, ga + (size_t)c * A_STAGE + o, p.piece, &full[slot]);
                    for (int o = 0; o < B_STAGE; o += p.piece) bulk_g2s(sb + o, gb + (size_t)c * B_STAGE + o, p.piece, &full[slot]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int64_t it = 0, k = 0;  // k counts the non-empty tiles of this CTA
            for (int64_t kk = 0; kk < my_tiles; kk++) {
                int ti_, tj_;
                if (!tile_coords(p, p.prefix, blockIdx.x + kk * gridDim.x, ti_, tj_)) continue;
                const int as = (int)(k & 1);
                if (k >= 2) mbar_wait_sleep(&tempty[as], (uint32_t)(((k >> 1) - 1) & 1));
                fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * N);
                for (int c = 0; c < p.nchunks; c++, it++) {
                    const int slot = (int)(it % NSTAGE);
                    mbar_wait(&full[slot], (uint32_t)((it / NSTAGE) & 1));
                    fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)slot * STAGE), sb = sa + A_STAGE;
#pragma unroll
                    for (int s = 0; s < KB / 32; s++) {  // one MMA = 32 K-bytes = two 16-byte K-halves
                        const uint64_t ad = smem_desc(sa + s * 2 * (M * 16), M * 16, 128);
                        const uint64_t bd = smem_desc(sb + s * 2 * (N * 16), N * 16, 128);
                        mma_i8(d_tmem, ad, bd, (c | s) != 0 ? 1u : 0u);
                    }
                    mma_commit(&empty[slot]);  // the stage may be refilled once these MMAs have read it
                }
                mma_commit(&tfull[as]);        // accumulator complete
                k++;
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps: TMEM lane quarter = warp % 4, column group = (warp - 2) / 4 =====
        // Sixteen warps (four per scheduler): the per-pair test is a short dependent integer chain, so latency is hidden
        // by warps; the flags of 32 pairs are computed branch-free (~9 instructions per pair) before the rare slow path.
        const int q = warp & 3, g = (warp - 2) >> 2;
        uint32_t *my_thr = col_thr + (warp - 2) * EPI_COLS;  // this warp's column thresholds (warp-private: no CTA barrier)
        const StreamParams &sp = c_sp;
        const uint32_t wm1 = p.W - 1u;
        int64_t k = 0;
        for (int64_t kk = 0; kk < my_tiles; kk++, k++) {
            int ti, tj;
            if (!tile_coords(p, p.prefix, blockIdx.x + kk * gridDim.x, ti, tj)) { k--; continue; }
            const int as = (int)(k & 1);
            const int i = ti * M + q * 32 + lane, j0 = tj * N;
            const bool seed = !COUNTS_MODE && (sp.want & TDNA_WANT_SEED) != 0;  // diagonal-block pre-pass: thresholds only
            const bool want_bm = !COUNTS_MODE && (sp.want & TDNA_WANT_BEST_MATCH);
            const uint32_t edge_tf = (!COUNTS_MODE && p.extras && (sp.want & TDNA_WANT_EDGES)) ? sp.edge_tfix : 0u;
            const bool classes = !COUNTS_MODE && !seed && p.extras && (sp.want & (TDNA_WANT_INTRA | TDNA_WANT_INTER));
            uint32_t tf_row = edge_tf;
            if constexpr (!COUNTS_MODE) {
                __syncwarp();  // the previous tile's reads of my_thr are done
#pragma unroll
                for (int x = 0; x < EPI_COLS / 32; x++) {
                    const int j = j0 + g * EPI_COLS + x * 32 + lane;
                    my_thr[x * 32 + lane] = (want_bm && j < p.n) ? __ldcg(sp.thr + j) : 0u;
                }
                if (want_bm && i < p.n) tf_row = max(tf_row, __ldcg(sp.thr + i));
                __syncwarp();
            }
            // every pair of the tile is a real pair i < j < n: no range tests in the inner loop
            const bool interior = (int64_t)tj * N > (int64_t)ti * M + M - 1 && (int64_t)tj * N + N <= p.n;
            mbar_wait(&tfull[as], (uint32_t)((k >> 1) & 1));
            fence_after();
            int bad_row = 0;
#pragma unroll
            for (int cbi = 0; cbi < EPI_COLS / 32; cbi++) {
                const int cb = g * (EPI_COLS / 32) + cbi;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * N + cb * 32), v);
                if (cbi == EPI_COLS / 32 - 1) {  // this warp's columns of the stage are in registers: hand them back
                    fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                const int jb = j0 + cb * 32;
                if constexpr (COUNTS_MODE) {
#pragma unroll
                    for (int e = 0; e < 32; e++) {
                        const uint32_t mism = v[e] >> p.wshift, ident = v[e] & wm1;
                        if (i < p.n && jb + e < p.n) {
                            tdna_counts o = {(int)(ident + mism), (int)ident, 0, 0};
                            p.counts[(size_t)i * p.n + jb + e] = o;
                        }
                    }
                } else {
                    uint32_t pmask = 0, bmask = 0, smask = 0;  // slow path wanted / pair below minOverlap / valid self pair
                    const uint4 *ct4 = reinterpret_cast<const uint4 *>(my_thr + cbi * 32);
                    if (interior) {
#pragma unroll
                        for (int e4 = 0; e4 < 8; e4++) {
                            const uint4 ct = ct4[e4];
                            const uint32_t cts[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int e = e4 * 4 + u;
                                const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1;  // s = ident + mism
                                const bool ok = (int)s >= p.min_overlap;
                                const bool pass = (mism << 16) <= max(tf_row, cts[u]) * s;
                                if (ok && pass) pmask |= 1u << e;
                                if (!ok) bmask |= 1u << e;
                            }
                        }
                    } else {
#pragma unroll
                        for (int e4 = 0; e4 < 8; e4++) {
                            const uint4 ct = ct4[e4];
                            const uint32_t cts[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int e = e4 * 4 + u, j = jb + e;
                                const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1;
                                const bool inr = (j > i) & (j < p.n);
                                const bool ok = (int)s >= p.min_overlap;
                                const bool pass = (mism << 16) <= max(tf_row, cts[u]) * s;
                                if (inr && ok && pass) pmask |= 1u << e;
                                if (inr && !ok) bmask |= 1u << e;
                                if (j == i && i < p.n && ok) smask |= 1u << e;
                            }
                        }
                    }
                    if (smask && !seed) atomicOr(sp.flags + i, TDNA_ROW_SELF_VALID);
                    if (bmask && !seed) {
                        bad_row += __popc(bmask);
#pragma unroll
                        for (int e = 0; e < 32; e++)
                            if (bmask & (1u << e)) atomicAdd(sp.inval + jb + e, 1);
                    }
                    if (pmask) {
#pragma unroll
                        for (int e = 0; e < 32; e++)
                            if (pmask & (1u << e)) {
                                const uint32_t mism = v[e] >> p.wshift, ident = v[e] & wm1, s = ident + mism;
                                stream_emit<KM_P_AMBIG>(i, jb + e, ident | (s << 16), 0u, mism << 16, s, p.min_overlap);
                            }
                    }
                    if (classes) {
#pragma unroll
                        for (int e = 0; e < 32; e++) {
                            const int j = jb + e;
                            const uint32_t mism = v[e] >> p.wshift, ident = v[e] & wm1, s = ident + mism;
                            if (j <= i || j >= p.n || (int)s < p.min_overlap) continue;
                            const bool hit = ((sp.want & TDNA_WANT_INTRA) && sp.species[i] == sp.species[j]) ||
                                             ((sp.want & TDNA_WANT_INTER) && sp.genus[i] == sp.genus[j]);
                            if (hit) class_emit<KM_P_AMBIG>(i, j, ident | (s << 16), 0u, p.min_overlap);
                        }
                    }
                }
            }
            if constexpr (!COUNTS_MODE) {
                if (bad_row) atomicAdd(sp.inval + i, bad_row);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace tc
}  // namespace tdna
