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
constexpr int NSTAGE = 4;                          // 4 x 48 KB (9 x 24 KB measured slower: small TMA boxes)
constexpr int A_STAGE = M * KB, B_STAGE = N * KB;  // 16 KB + 32 KB
constexpr int STAGE = A_STAGE + B_STAGE;
constexpr int EPI_WARPS = 16;                      // four per TMEM lane quarter
constexpr int EPI_COLS = N / (EPI_WARPS / 4);      // columns per epilogue warp
constexpr int THREADS = 64 + EPI_WARPS * 32;       // warp 0 producer, warp 1 MMA issuer, warps 2.. epilogue
constexpr int MAX_DIMS = 6;                        // K-dimensions per site: 4 or 5 (one-hot A C G T [-]), 6 (transversions only)
constexpr int WTAB = 64;                           // slots of an epilogue warp's private class-histogram table (tdna_kernels.cuh: shist_add)
constexpr size_t SMEM = (size_t)NSTAGE * STAGE + 256 /*barriers*/ + 2 * 16 * 64 * 4 /*column thresholds + raw bounds, per epilogue warp*/ +
                        (size_t)EPI_WARPS * WTAB * 12 /*class-histogram tables*/ + 1024 /*alignment*/;

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
    int64_t num_tiles, t_offset, t_stride;  // this launch's k-th work item is global item ((k / unit) * t_stride + t_offset) * unit + k % unit
    int64_t k_base;         // first work item of this launch in the part's own list (a long pass goes out in segments)
    // a REGION launch (the pass pipelined with the upload, tdna_analyze_host): the launch's items are reg_n runs of consecutive slot
    // pairs -- per band, the columns of the row chunk that has just arrived; item l lies in run k with reg_prefix[k] <= l <
    // reg_prefix[k + 1] and is slot pair reg_start[k] + l - reg_prefix[k]
    const int64_t *reg_start, *reg_prefix;
    int reg_n;
    int unit;               // work items per scheduling unit (parts take units cyclically: balanced in work AND in candidates)
    tdna_counts *counts;    // verification mode: n x n counts (row-major), else null
    int extras;             // stream mode: classes / edges wanted (StreamParams::want beyond Best Match)
    const uint8_t *panel_full;  // [128-row panel] 1 = every row of the panel is valid at every site: pairs of two such panels share L sites
    int L;
    int k2p;                // the accumulator carries (n, ts + tv) of the K2P model; exact counts come from the bit-planes
    int trans;              // transversions only: mism = transversions, ident = shared - transversions (exact); d = transv / shared
    const uint32_t *planes; // K2P: the popc kernel's planes, for the exact counts of queued / class pairs
    int W32;
    int a_rows, b_rows;     // rows of the tensor-map view per a / b stage block (view row = inner box width)
    int hints;              // L2 cache-policy hints on the operand loads (experiments: TDNA_TC_HINTS)
    int piece;              // bytes per 1-D bulk copy (divides 16 KB); 0 = tensor-map (2-D tiled TMA) loads
    // IUPAC ambiguity letters are encoded as "no contribution"; amb[i] = how many of them row i
    // has.  The accumulator then bounds the pair from both sides -- mism >= mism', shared' <= shared <= shared' + amb[i] + amb[j] --
    // which is all the candidate filter needs; exact counts of the queued pairs come from the bit-planes (like K2P's).
    const int32_t *amb;        // nullptr: the alignment has no such letter
    const uint8_t *panel_amb;  // [128-row panel] 1 = some row of the panel has one
    // seed pass: the PAIRS of row tiles to visit, (even row tile, column tile) each -- the diagonal block of every pair plus, where the
    // labels say a row would otherwise see no conspecific or no species of the other id parity, the neighbouring block that has one
    const int2 *seed_list;
    int seed_pairs;
    int class_only;         // the pass wants class distances and nothing per row: bit 0 = intra, bit 1 = inter.  Tiles whose row and
    const int4 *panel_range;  // column panels share no species / genus id range hold no such pair and are skipped by every role
    // (min, max) species id and (min, max) genus id per 128-row tile panel and per 256-column tile panel: two loads per tile decide
    // whether a tile can hold a class pair (the 64-row panels' ranges cost 16 loads per thread and tile -- measurable on 10^6 rows)
    const int4 *range_rows, *range_cols;
    int64_t col_reach;      // >= 0: the slot numbering of this launch stops col_reach rows past each band (class-only launches)
    int64_t class_reach;    // no class pair (i, j > i) has j - i above this (from the labels; n when unknown): tiles further from the
                            // diagonal are ruled out by arithmetic alone, before any range is loaded
};

// Tile slot -> (row tile, column tile); false = an empty slot.  The bulk phase walks BANDS of `band` row tiles (64 by
// default), column tile outer / row tile inner, so that the ~148 tiles in flight at any time share the band's a-panels
// (L2-resident for the whole band) and two or three b-panels: every b-panel is fetched from HBM once per band.
__device__ __forceinline__ bool slot_coords(const Params &p, const int64_t *prefix, int64_t t, int &ti, int &tj) {
    if (p.band == 0) {
        if (p.seed_list) {
            if ((t >> 1) >= p.seed_pairs) return false;
            const int2 e = __ldg(p.seed_list + (t >> 1));
            ti = e.x + (int)(t & 1);
            tj = e.y;
            return ti < p.nrt && tj < p.nct;
        }
        ti = (int)t;
        tj = ti >> 1;
        return ti < p.nrt && tj < p.nct;
    }
    int lo = 0, hi = p.nbands;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid;
    }
    // Slots of a band, column tile by column tile: column c of the band (tile tj = tjmin + c) meets the band's first min(band, 2c + 2)
    // row tiles (the rest lie below the diagonal), so the first band / 2 columns form a staircase of c(c + 1) slots before column c
    // and every later column holds `band` slots.  Only real tiles are numbered: parts that take units of consecutive slots
    // cyclically get equal work (a rectangular numbering left the staircase's holes to some parts and not to others: 16 % imbalance
    // at eight parts).  Counts are even, so the slot pairs (2u, 2u + 1) of the two-CTA clusters stay adjacent row tiles of one column.
    const int64_t r = t - prefix[lo];
    const int ti0 = lo * p.band, tjmin = ti0 >> 1;
    // a class-only launch numbers, per band, only the columns a class pair can reach (col_reach rows past the band's last row)
    const int lim = p.col_reach >= 0 ? (int)min((int64_t)p.nct, ((int64_t)(ti0 + p.band) * M - 1 + p.col_reach) / N + 1) : p.nct;
    const int cols = max(lim - tjmin, 0), tri = min(cols, p.band >> 1);
    const int64_t T = (int64_t)tri * (tri + 1);
    if (r < T) {
        int c = (int)((sqrtf(4.0f * (float)r + 1.0f) - 1.0f) * 0.5f);
        while ((int64_t)c * (c + 1) > r) c--;
        while ((int64_t)(c + 1) * (c + 2) <= r) c++;
        ti = ti0 + (int)(r - (int64_t)c * (c + 1));
        tj = tjmin + c;
    } else {
        const int64_t r2 = r - T;
        ti = ti0 + (int)(r2 % p.band);
        tj = tjmin + tri + (int)(r2 / p.band);
    }
    return ti < p.nrt && tj < p.nct && tj >= (ti >> 1);
}

// k-th work item of this CTA (items = slots for CL = 1, PAIRS of slots (2u, 2u + 1) -- same column tile, adjacent row tiles -- for
// CL = 2); a part takes every t_stride-th UNIT of `unit` consecutive items.  Items past the end of the list are empty.  CL = 2: this
// CTA the one of its cluster rank; the pair is as valid as its even slot (an odd row tile past the end reads zero-filled operands).
// rows [r0, r0 + rows) as a union of the 64-row panels' (min, max) species / genus ids
__device__ __forceinline__ int4 range_union(const Params &p, int64_t r0, int rows) {
    int4 r = make_int4(0x7fffffff, -1, 0x7fffffff, (int)0x80000000);
    for (int64_t pn = r0 / PR; pn * PR < r0 + rows && pn * PR < p.n; pn++) {
        const int4 x = __ldg(p.panel_range + pn);
        r.x = min(r.x, x.x); r.y = max(r.y, x.y); r.z = min(r.z, x.z); r.w = max(r.w, x.w);
    }
    return r;
}
// a pass that only collects class distances: can the item (ROWS rows from row tile ti, column tile tj) hold an intra / inter pair?
__device__ __forceinline__ bool may_hold_class_pairs(const Params &p, int ti, int tj, int rows) {
    const int4 a = range_union(p, (int64_t)ti * M, rows), b = range_union(p, (int64_t)tj * N, N);
    return ((p.class_only & 1) && a.x <= b.y && b.x <= a.y) || ((p.class_only & 2) && a.z <= b.w && b.z <= a.w);
}
// CL = 4: clusters of FOUR CTAs on a block of 2 x 2 tiles (row tiles ti0, ti0 + 1; column tiles 2 sc, 2 sc + 1): rank = 2 r + c owns tile
// (ti0 + r, 2 sc + c), fetches HALF of its row tile's a-block for both CTAs of its row and half of its column tile's b-block for both
// CTAs of its column -- 24 KB of L2 reads per CTA and K-chunk instead of 32 KB.  (The pass is bound by the L2 -> SM fabric: ncu reads
// 12.2 TB/s of L2 reads on 50,000 x 658, the chip's measured cap.)  Quads are numbered per band: super column outer, row pair inner;
// prefix[b] = quads before band b.  Return: 0 = nothing to do (the same answer in all four CTAs), 1 = run, 2 = run the loads and the
// MMAs for the cluster's sake but this CTA's own tile lies outside the triangle / the list: no epilogue work.
template <int CL> __device__ __forceinline__ int my_tile(const Params &p, int64_t k, uint32_t rank, int &ti, int &tj) {
    if constexpr (CL == 4) {
        const int64_t l = p.k_base + (int64_t)(blockIdx.x >> 2) + k * (gridDim.x >> 2);
        const int64_t u = ((l / p.unit) * p.t_stride + p.t_offset) * p.unit + l % p.unit;
        int lo = 0, hi = p.nbands;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(p.prefix + mid) <= u) lo = mid; else hi = mid;
        }
        const int64_t q = u - __ldg(p.prefix + lo);
        const int rows2 = p.band >> 1, scmin = (lo * p.band) >> 2, nsc = (p.nct + 1) >> 1;
        const int sc = scmin + (int)(q / rows2), ti0 = lo * p.band + 2 * (int)(q % rows2);
        if (u >= __ldg(p.prefix + p.nbands) || sc >= nsc || ti0 >= p.nrt || min(2 * sc + 1, p.nct - 1) < (ti0 >> 1)) return 0;
        ti = ti0 + (int)(rank >> 1);
        tj = 2 * sc + (int)(rank & 1);
        return (ti < p.nrt && tj < p.nct && tj >= (ti >> 1)) ? 1 : 2;
    } else
    if constexpr (CL == 1) {
        const int64_t l = p.k_base + (int64_t)blockIdx.x + k * gridDim.x;
        if (!slot_coords(p, p.prefix, ((l / p.unit) * p.t_stride + p.t_offset) * p.unit + l % p.unit, ti, tj)) return false;
        return !p.class_only || ((int64_t)tj * N <= (int64_t)ti * M + (M - 1) + p.class_reach && may_hold_class_pairs(p, ti, tj, M));
    } else {
        const int64_t l = p.k_base + (int64_t)(blockIdx.x >> 1) + k * (gridDim.x >> 1);
        int64_t u;
        if (p.reg_n) {
            int lo = 0, hi = p.reg_n;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(p.reg_prefix + mid) <= l) lo = mid; else hi = mid;
            }
            u = __ldg(p.reg_start + lo) + (l - __ldg(p.reg_prefix + lo));
        } else u = ((l / p.unit) * p.t_stride + p.t_offset) * p.unit + l % p.unit;
        bool ok = slot_coords(p, p.prefix, 2 * u, ti, tj);
        if (ok && p.class_only) ok = (int64_t)tj * N <= (int64_t)ti * M + (2 * M - 1) + p.class_reach && may_hold_class_pairs(p, ti, tj, 2 * M);  // the same answer in both CTAs of the pair
        ti += (int)rank;
        return ok;
    }
}

// ---- operand encoding -------------------------------------------------------------------------
struct EncodeValues {
    int a, b, c, d;
    int gap_is_symbol;  // uncorrected p: '-' is the fifth symbol; K2P ignores gap sites (Sequence.java:1446-1471): no contribution
    // What the encoders read: per symbol CODE (0 .. 5; 5 = no contribution) the row-side and the column-side int8 vector over the
    // K-dimensions of a site.  One-hot codes (p, K2P): code = symbol, ta[code][dim] = b + a [dim == code], tb = d + c [dim == code].
    // Transversions only (trans = 1, six dimensions, W = r^2): code 0 purine {A G R}, 1 pyrimidine {C T Y}, 2 '-', 3 any other letter;
    //   row    purine (1 0 r 0 | 0 1)  pyrimidine (0 1 0 r | 0 1)  '-' (0 0 0 0 | 1 1)  other (0 0 0 0 | 0 1)
    //   column purine (1 0 0 r | 1 0)  pyrimidine (0 1 r 0 | 1 0)  '-' (0 0 0 0 | 0 1)  other (0 0 0 0 | 1 0)
    // so that u . v = [both in AGRCTY] + (W - 1) [purine x pyrimidine] + [both valid, one a '-'] = (shared - transv) + W * transv:
    // the accumulator decodes like uncorrected p's with ident := shared - transversions (countTransversions :1352-1371,
    // getSharedLength in that mode :1446-1451, 1472-1478) -- EXACT for every letter, no ambiguity-letter machinery.
    int trans;
    int8_t ta[6][8], tb[6][8];
    uint8_t breaks_full[6];  // a site holding this code keeps its 128-row panel from being "full" (shared == L for pairs of full panels)
};
// the symbol-class word of the pack kernel's table -> code
__device__ __forceinline__ uint32_t symbol_code(uint32_t cls, const EncodeValues &v) {
    if (v.trans) {
        if (cls & SB_U) return (cls & SB_PU) ? 0u : 1u;
        if (cls & SB_D) return 2u;
        if (cls & SB_LET) return 3u;
        return (cls & 0x8000u) ? 0x85u : 5u;
    }
    const uint32_t bases = cls & 15u;
    if (cls & SB_LET) return __popc(bases) == 1 ? (uint32_t)__ffs((int)bases) - 1u : 5u;  // an ambiguity letter contributes nothing here
    if (cls & SB_D) return v.gap_is_symbol ? 4u : 5u;
    return (cls & 0x8000u) ? 0x85u : 5u;
}
// one CTA = one operand block (panel x K-chunk): the ~27 sites the chunk covers are staged as symbol codes in shared
// memory (coalesced 27-byte runs per row), then every thread emits 16-byte units in memory order (coalesced uint4 stores)
constexpr int ENC_SITES = KB / 4 + 3;
template <bool BSIDE, int ROWS>
__global__ void __launch_bounds__(256) encode_kernel(const uint8_t *__restrict__ sym, const int32_t *__restrict__ len, int64_t n, int64_t sym_stride,
                                                     int L, int nchunks, int DIMS, EncodeValues v, uint8_t *__restrict__ out, int *__restrict__ not_clean,
                                                     uint8_t *__restrict__ panel_gappy) {
    __shared__ uint8_t codes[ROWS][ENC_SITES + 1];  // 0..4 = A C G T '-', 5 = no contribution ('_', '?', beyond the row; '-' when DIMS == 4)
    const int c = blockIdx.x;
    const int64_t panel = blockIdx.y;
    const int s0 = (c * KB) / DIMS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    bool bad = false, gappy = false;
    for (int row_l = warp; row_l < ROWS; row_l += 8) {  // one warp per row: a run of ENC_SITES consecutive bytes
        const int64_t row = panel * ROWS + row_l;
        const int row_len = row < n ? min(len[row], L) : 0;
        for (int s = lane; s < ENC_SITES; s += 32) {
            const int site = s0 + s;
            uint32_t code = 5;
            if (site < row_len) {
                code = symbol_code(c_symlut[sym[row * sym_stride + site]], v);  // the pack kernel's table: IUPAC mask in bits 0..3 (A C G T)
                if (code & 0x80u) { bad = true; code = 5; }
            }
            codes[row_l][s] = (uint8_t)code;
            // a site of the alignment of an existing row that contributes nothing: the panel is not "full"
            if (row < n && site < L && v.breaks_full[code]) gappy = true;
        }
    }
    if (bad) atomicOr(not_clean, 1);
    if constexpr (!BSIDE) {  // only the a-side launch tracks this (its panels are the 128-row panels the tile kernel asks about)
        if (gappy) panel_gappy[panel] = 1;
    }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)panel * nchunks + c) * (size_t)(ROWS * KB));
    for (int u = threadIdx.x; u < ROWS * (KB / 16); u += 256) {
        const int row_l = u % ROWS, k16 = u / ROWS;  // u = (k16 * (ROWS / 8) + rg) * 8 + r, row = rg * 8 + r
        uint32_t w[4] = {0, 0, 0, 0};
        int kk = c * KB + k16 * 16, site = kk / DIMS - s0, dim = kk % DIMS;  // one division per 16 bytes, then a running (site, dim)
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int code = codes[row_l][site];
            const int val = BSIDE ? v.tb[code][dim] : v.ta[code][dim];
            w[b >> 2] |= (uint32_t)(val & 0xff) << (8 * (b & 3));
            if (++dim == DIMS) { dim = 0; site++; }
        }
        dst[u] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Both operand images from ONE read of the symbols (the default, multicast-pair layout): one CTA = one K-chunk of one 256-row b-panel
// = the same chunk of two 128-row a-panels.  The symbol -> code table sits in shared memory (a divergent constant-memory index
// serialises), one warp stages a row's ~27 symbols, then every thread emits 16-byte units of the b-image and of the a-images.
__global__ void __launch_bounds__(256) encode_both_kernel(const uint8_t *__restrict__ sym, const int32_t *__restrict__ len, int64_t n, int64_t sym_stride,
                                                          int L, int nchunks, int DIMS, EncodeValues v, uint8_t *__restrict__ outA, uint8_t *__restrict__ outB,
                                                          int64_t a_panels, int *__restrict__ not_clean, uint8_t *__restrict__ panel_gappy, int64_t panel0) {
    __shared__ uint8_t codes[N][ENC_SITES + 1];
    __shared__ uint8_t s_code[256];  // the symbol's code (EncodeValues), 5 no contribution, 0x85 outside the alphabet
    __shared__ unsigned long long s_ta[6], s_tb[6];  // a code's operand bytes over the K-dimensions of a site, packed (dimension d in byte d)
    s_code[threadIdx.x] = (uint8_t)symbol_code(c_symlut[threadIdx.x], v);
    if (threadIdx.x < 6) {
        unsigned long long wa = 0, wb = 0;
        for (int d = 0; d < 8; d++) {
            wa |= (unsigned long long)(uint8_t)v.ta[threadIdx.x][d] << (8 * d);
            wb |= (unsigned long long)(uint8_t)v.tb[threadIdx.x][d] << (8 * d);
        }
        s_ta[threadIdx.x] = wa;
        s_tb[threadIdx.x] = wb;
    }
    __syncthreads();
    const int c = blockIdx.x;
    const int64_t panel = panel0 + blockIdx.y;  // 256-row panel (panel0: the first panel of a chunk of a pipelined upload)
    const int s0 = (c * KB) / DIMS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    bool bad = false;
    uint32_t gappy = 0;  // bit h: half h (a-panel 2 * panel + h) has a row with a site that contributes nothing
    for (int row_l = warp; row_l < N; row_l += 8) {
        const int64_t row = panel * N + row_l;
        const int row_len = row < n ? min(__ldg(len + row), L) : 0;
        const uint8_t *src = sym + row * sym_stride + s0;
#pragma unroll
        for (int s = lane; s < ENC_SITES; s += 32) {
            uint32_t code = 5;
            if (s0 + s < row_len) code = s_code[__ldg(src + s)];
            if (code & 0x80u) { bad = true; code = 5; }
            codes[row_l][s] = (uint8_t)code;
            if (row < n && s0 + s < L && v.breaks_full[code]) gappy |= 1u << (row_l >> 7);
        }
    }
    if (bad) atomicOr(not_clean, 1);
    if (gappy & 1u) panel_gappy[2 * panel] = 1;
    if (gappy & 2u) panel_gappy[2 * panel + 1] = 1;
    __syncthreads();
    auto unit = [&](int row_l, int k16, const unsigned long long *tab) {
        uint32_t w[4] = {0, 0, 0, 0};
        int kk = c * KB + k16 * 16, site = kk / DIMS - s0, dim = kk % DIMS;
        unsigned long long tw = tab[codes[row_l][site]] >> (8 * dim);  // one table word per SITE, a byte of it per K-byte
#pragma unroll
        for (int b = 0; b < 16; b++) {
            w[b >> 2] |= ((uint32_t)tw & 0xffu) << (8 * (b & 3));
            tw >>= 8;
            if (++dim == DIMS) { dim = 0; site++; tw = tab[codes[row_l][site]]; }
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    uint4 *dstB = reinterpret_cast<uint4 *>(outB + ((size_t)panel * nchunks + c) * (size_t)(N * KB));
    for (int u = threadIdx.x; u < N * (KB / 16); u += 256) dstB[u] = unit(u % N, u / N, s_tb);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (2 * panel + h >= a_panels) break;
        uint4 *dstA = reinterpret_cast<uint4 *>(outA + ((size_t)(2 * panel + h) * nchunks + c) * (size_t)(M * KB));
        for (int u = threadIdx.x; u < M * (KB / 16); u += 256) dstA[u] = unit(h * M + u % M, u / M, s_ta);
    }
}

// panel_gappy -> panel_full; panels beyond the last row are not full (their pairs are out of range anyway)
__global__ void __launch_bounds__(256) flip_flags_kernel(uint8_t *flags, int64_t k0, int64_t npan, int64_t real_panels) {
    int64_t k = k0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npan) flags[k] = (k < real_panels && flags[k] == 0) ? 1 : 0;
}

// per row: the number of sites (< min(len, L)) holding a letter the operands cannot express, i.e. any IUPAC ambiguity letter.
// (K2P: only R and Y are inside the model, Sequence.java:1498-1557, but every letter counts towards the K2P-mode shared length
// that the minOverlap gate tests, :1446-1471 -- one count bounds both.)
__global__ void __launch_bounds__(256) amb_count_kernel(const uint8_t *__restrict__ sym, const int32_t *__restrict__ len, int64_t n, int64_t sym_stride,
                                                        int L, int k2p, int32_t *__restrict__ amb, uint8_t *__restrict__ panel_amb, int *__restrict__ any, int64_t row0,
                                                        int gaps_dropped) {
    // gaps_dropped: the four-dimension code of uncorrected p leaves internal gaps out of the operands too (any[1] counts those sites:
    // a list with more than a sprinkling of them is re-encoded with the five-dimension code, where '-' is a symbol of its own)
    const int64_t row = row0 + (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = threadIdx.x & 31, row_len = min(len[row], L);
    int cnt = 0, gaps = 0, oh = 0;  // oh: sites the operands DO express (one unambiguous base; '-' where it is a symbol of its own)
    for (int site = lane; site < row_len; site += 32) {
        const uint32_t cls = c_symlut[sym[row * sym_stride + site]];
        if (cls & SB_LET) { if (__popc(cls & 15u) != 1) cnt++; else oh++; }
        if (cls & SB_D) gaps++;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
        gaps += __shfl_xor_sync(0xffffffffu, gaps, m);
        oh += __shfl_xor_sync(0xffffffffu, oh, m);
    }
    if (lane == 0) {
        const int dropped = gaps_dropped ? gaps : 0;  // internal gaps the operands leave out
        if (!gaps_dropped && !k2p) oh += gaps;        // ... or express as a symbol of their own
        amb[row] = (cnt + dropped) | (oh << 12);  // both below 4096 (L < 4096): AMB_MASK / AMB_SHIFT
        if (cnt + dropped) { panel_amb[row / M] = 1; *any = 1; }
        if (gaps) atomicAdd(any + 1, gaps);  // internal gap sites, whichever code is in use (the next list's code is chosen from this)
    }
}
constexpr uint32_t AMB_MASK = 0xfffu;
constexpr int AMB_SHIFT = 12;

// per tile panel (`per` consecutive 64-row panels): the union of the 64-row panels' (min, max) species / genus id ranges
__global__ void __launch_bounds__(256) tile_range_kernel(const int4 *__restrict__ panel_range, int64_t panels64, int per, int4 *__restrict__ out, int64_t nout) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nout) return;
    int4 r = make_int4(0x7fffffff, -1, 0x7fffffff, (int)0x80000000);
    for (int k = 0; k < per; k++) {
        const int64_t pn = t * per + k;
        if (pn >= panels64) break;
        const int4 x = panel_range[pn];
        r.x = min(r.x, x.x); r.y = max(r.y, x.y); r.z = min(r.z, x.z); r.w = max(r.w, x.w);
    }
    out[t] = r;
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
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {  // one accumulator column: lane r of the warp gets row r's value
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return v;
}
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

// seed pass, tiles with ambiguity letters: the exact distance of one pair from the bit-planes, as ordered bits (+Inf bits: no use)
__device__ __noinline__ unsigned long long exact_seed_bits(const Params &p, int i, int j) {
    if (j < 0) return TDNA_INF_BITS;
    double d;
    if (p.k2p) d = distance_from_counts<KM_K2P>(pair_counts_global<KM_K2P>(p.planes, i, j, p.W32), p.min_overlap);
    else d = distance_from_counts<KM_P_AMBIG>(pair_counts_global<KM_P_AMBIG>(p.planes, i, j, p.W32), p.min_overlap);
    if (!(d < __longlong_as_double((long long)TDNA_INF_BITS)) || d < 0.0) return TDNA_INF_BITS;
    return (unsigned long long)__double_as_longlong(d);
}

// the distance of a pair whose accumulator is exact: 1 - ident / shared (uncorrected p, Sequence.java:1711) or transversions / shared
// (transversions only, :1708; ident carries shared - transversions there)
__device__ __forceinline__ double exact_acc_distance(int trans, uint32_t ident, uint32_t s, int min_overlap) {
    if (trans) {
        Counts cn = {(int)s, (int)(s - ident), 0, 0};
        return distance_from_counts<KM_TRANS>(cn, min_overlap);
    }
    Counts cn = {(int)s, (int)ident, 0, 0};
    return distance_from_counts<KM_P_AMBIG>(cn, min_overlap);
}

// ---- the tile kernel ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d_mc(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t *bar, uint16_t cta_mask) {  // arrive on the same barrier of every CTA in the mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// cta_group::2: both CTAs of the pair load into their own shared memory, the bytes are credited to the LEADER's barrier
// (the barrier address with the peer bit cleared addresses CTA 0 of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(L2_EVICT_NORMAL)
                 : "memory");
}
constexpr uint32_t IDESC_PAIR = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * M) >> 4) << 24);
__device__ __forceinline__ void mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {  // M = 256 across the CTA pair
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC_PAIR), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar) {  // arrives on the same barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t *bar, uint32_t cta) {  // arrive on `bar` of CTA `cta` of the cluster
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CL = 2: clusters of two CTAs on the same column tile; each loads one half of the b-panel stage and multicasts it to both,
// so the pair reads 2 x 16 KB (a) + 32 KB (b) per K-chunk from L2 instead of 2 x 48 KB.
// GENERAL = false: uncorrected p on an alignment without ambiguity letters -- every accumulator is exact and none of the K2P /
// ambiguity-letter machinery is compiled in (it cost the hot loop 4 % through register pressure alone).
template <bool COUNTS_MODE, int CL, bool GENERAL>
__global__ void __launch_bounds__(THREADS, 1) tile_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                          const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ Params p,
                                                          const __grid_constant__ StreamParams sp) {
    // CL = 3: the pair runs ONE M = 256 MMA (cta_group::2): each CTA stages its own a-panel and its 128-row HALF of the
    // b-panel, i.e. 32 KB per stage instead of 48 KB -> a six-deep ring, and half the b-operand shared-memory reads
    constexpr bool PAIR = CL == 3;
    constexpr int B_ST = PAIR ? B_STAGE / 2 : B_STAGE, STG = A_STAGE + B_ST, NST = PAIR ? 6 : NSTAGE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NSTAGE * STAGE);
    uint64_t *empty = full + 8;
    uint64_t *tfull = empty + 8;        // [2] accumulator stage ready for the epilogue
    uint64_t *tempty = tfull + 2;       // [2] accumulator stage drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    uint32_t *col_thr = reinterpret_cast<uint32_t *>(smem + (size_t)NSTAGE * STAGE + 256);  // [EPI_WARPS][EPI_COLS]
    uint32_t *col_raw = col_thr + EPI_WARPS * EPI_COLS;  // [EPI_WARPS][EPI_COLS] tiles with ambiguity letters: the columns' bounds in accumulator units
    unsigned long long *wtab_keys = reinterpret_cast<unsigned long long *>(col_raw + EPI_WARPS * EPI_COLS);  // [EPI_WARPS][WTAB]
    unsigned int *wtab_cnts = reinterpret_cast<unsigned int *>(wtab_keys + EPI_WARPS * WTAB);                // [EPI_WARPS][WTAB]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank = 0;
    if constexpr (CL >= 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    constexpr int CSHIFT = CL == 1 ? 0 : (CL == 4 ? 2 : 1);  // CTAs per work item: 1, 2 or 4
    const int64_t units = gridDim.x >> CSHIFT, first = blockIdx.x >> CSHIFT;
    const int64_t my_tiles = (p.num_tiles > first) ? (p.num_tiles - first + units - 1) / units : 0;  // tiles (CL = 1) or tile pairs (CL = 2)

    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL == 4 ? 4 : (CL == 2 ? 2 : 1)); }
        for (int s = 0; s < 2; s++) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 256-column accumulator stages)
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    fence_before();
    __syncthreads();
    fence_after();
    if constexpr (CL >= 2) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int64_t it = 0;
            for (int64_t k = 0; k < my_tiles; k++) {
                int ti, tj;
                if (!my_tile<CL>(p, k, rank, ti, tj)) continue;
                const uint8_t *ga = p.A + (size_t)ti * p.nchunks * A_STAGE;
                const uint8_t *gb = p.B + (size_t)tj * p.nchunks * B_STAGE;
                for (int c = 0; c < p.nchunks; c++, it++) {
                    const int slot = (int)(it % NST);
                    if (it >= NST) mbar_wait(&empty[slot], (uint32_t)(((it / NST) - 1) & 1));  // the ring IS the bottleneck here: no back-off
                    uint8_t *sa = smem + (size_t)slot * STG, *sb = sa + A_STAGE;
                    if constexpr (PAIR) {  // both CTAs load (own a-panel, own half of the b-panel); the leader's barrier counts all 64 KB
                        if (rank == 0) mbar_expect_tx(&full[slot], 2 * STG);
                        tma_load_2d_pair(sa, &mapA, 0, (int)(((int64_t)ti * p.nchunks + c) * p.a_rows), &full[slot]);
                        tma_load_2d_pair(sb, &mapBh, 0, (int)(((int64_t)(2 * tj + (int)rank) * p.nchunks + c) * p.a_rows), &full[slot]);
                        continue;
                    }
                    mbar_expect_tx(&full[slot], STG);
                    if constexpr (CL == 4) {  // half of the row tile's a-block to the two CTAs of this row, half of the column tile's b-block to the two of this column
                        const uint32_t r = rank >> 1, cx = rank & 1;
                        tma_load_2d_mc(sa + cx * (A_STAGE / 2), &mapB /* CL = 4: the half-a-block view of the a-side */, 0,
                                       (int)(((int64_t)ti * p.nchunks + c) * p.a_rows + cx * (p.a_rows / 2)), &full[slot], (uint16_t)(3u << (2 * r)));
                        tma_load_2d_mc(sb + r * (B_STAGE / 2), &mapBh, 0, (int)(((int64_t)tj * p.nchunks + c) * p.b_rows + r * (p.b_rows / 2)),
                                       &full[slot], (uint16_t)(5u << cx));
                        continue;
                    }
                    if constexpr (CL == 2) {  // own a-panel; own half of the b-panel stage to both CTAs (same offsets, same barrier)
                        tma_load_2d(sa, &mapA, 0, (int)(((int64_t)ti * p.nchunks + c) * p.a_rows), &full[slot]);
                        tma_load_2d_mc(sb + rank * (B_STAGE / 2), &mapBh, 0, (int)(((int64_t)tj * p.nchunks + c) * p.b_rows + rank * (p.b_rows / 2)),
                                       &full[slot], (uint16_t)3);
                        continue;
                    }
                    if (p.piece == 0) {  // one block = 16 (a) / 32 (b) rows of 1 KB
                        if (p.hints) {
                            tma_load_2d_hint(sa, &mapA, 0, (int)(((int64_t)ti * p.nchunks + c) * p.a_rows), &full[slot], p.hints == 2 ? L2_EVICT_NORMAL : L2_EVICT_LAST);
                            tma_load_2d_hint(sb, &mapB, 0, (int)(((int64_t)tj * p.nchunks + c) * p.b_rows), &full[slot], p.hints == 3 ? L2_EVICT_FIRST : (p.hints == 2 ? L2_EVICT_LAST : L2_EVICT_NORMAL));
                        } else {
                            tma_load_2d(sa, &mapA, 0, (int)(((int64_t)ti * p.nchunks + c) * p.a_rows), &full[slot]);
                            tma_load_2d(sb, &mapB, 0, (int)(((int64_t)tj * p.nchunks + c) * p.b_rows), &full[slot]);
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
        // ===== MMA issuer (CL = 3: the leader CTA issues for the pair) =====
        if (lane == 0 && !(PAIR && rank != 0)) {
            int64_t it = 0, k = 0;  // k counts the non-empty tiles of this CTA
            for (int64_t kk = 0; kk < my_tiles; kk++) {
                int ti_, tj_;
                if (!my_tile<CL>(p, kk, rank, ti_, tj_)) continue;
                const int as = (int)(k & 1);
                if (k >= 2) mbar_wait_sleep(&tempty[as], (uint32_t)(((k >> 1) - 1) & 1));
                fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * N);
                for (int c = 0; c < p.nchunks; c++, it++) {
                    const int slot = (int)(it % NST);
                    mbar_wait(&full[slot], (uint32_t)((it / NST) & 1));
                    fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)slot * STG), sb = sa + A_STAGE;
#pragma unroll
                    for (int s = 0; s < KB / 32; s++) {  // one MMA = 32 K-bytes = two 16-byte K-halves
                        const uint64_t ad = smem_desc(sa + s * 2 * (M * 16), M * 16, 128);
                        if constexpr (PAIR) {  // each CTA holds 128 of the tile's 256 b-rows, laid out like an a-panel
                            const uint64_t bd = smem_desc(sb + s * 2 * (M * 16), M * 16, 128);
                            mma_i8_pair(d_tmem, ad, bd, (c | s) != 0 ? 1u : 0u);
                        } else {
                            const uint64_t bd = smem_desc(sb + s * 2 * (N * 16), N * 16, 128);
                            mma_i8(d_tmem, ad, bd, (c | s) != 0 ? 1u : 0u);
                        }
                    }
                    if constexpr (PAIR) mma_commit_pair(&empty[slot]);                     // the stage of BOTH CTAs may be refilled
                    else if constexpr (CL == 4) mma_commit_mc(&empty[slot], (uint16_t)15);  // every CTA of the quad waits for all four before it refills a stage
                    else if constexpr (CL == 2) mma_commit_mc(&empty[slot], (uint16_t)3);  // both CTAs write into this stage: tell both
                    else mma_commit(&empty[slot]);  // the stage may be refilled once these MMAs have read it
                }
                if constexpr (PAIR) mma_commit_pair(&tfull[as]);  // accumulator complete, in both CTAs' tensor memory
                else mma_commit(&tfull[as]);
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
        uint32_t *my_raw = col_raw + (warp - 2) * EPI_COLS;
        unsigned long long *my_keys = wtab_keys + (warp - 2) * WTAB;
        unsigned int *my_cnts = wtab_cnts + (warp - 2) * WTAB;
        for (int s = lane; s < WTAB; s += 32) { my_keys[s] = TDNA_HIST_EMPTY; my_cnts[s] = 0; }
        __syncwarp();
        bool wtab_used = false;
        const uint32_t wm1 = p.W - 1u;
        int64_t k = 0;
        for (int64_t kk = 0; kk < my_tiles; kk++, k++) {
            int ti, tj;
            const int tile_state = my_tile<CL>(p, kk, rank, ti, tj);
            if (!tile_state) { k--; continue; }
            const int as = (int)(k & 1);
            if (tile_state == 2) {  // CL = 4: this CTA's tile lies outside; its MMAs ran for the cluster's sake -- hand the accumulator back
                mbar_wait(&tfull[as], (uint32_t)((k >> 1) & 1));
                fence_after();
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
                continue;
            }
            const int i = ti * M + q * 32 + lane, j0 = tj * N;
            const bool seed = !COUNTS_MODE && (sp.want & TDNA_WANT_SEED) != 0;  // diagonal-block pre-pass: thresholds only
            const bool want_bm = !COUNTS_MODE && (sp.want & TDNA_WANT_BEST_MATCH);
            const uint32_t edge_tf = (!COUNTS_MODE && p.extras && (sp.want & TDNA_WANT_EDGES)) ? sp.edge_tfix : 0u;
            const bool k2p_force = GENERAL && p.k2p != 0;  // K2P: 2(ts + tv) >= n may mean w1 <= 0 or w2 <= 0 -> NaN / +Inf must reach the exact path
            bool classes = !COUNTS_MODE && !seed && p.extras && (sp.want & (TDNA_WANT_INTRA | TDNA_WANT_INTER));
            if (classes) classes = (int64_t)tj * N <= (int64_t)ti * M + (M - 1) + p.class_reach;
            if (classes) {  // class pairs only live in tiles whose row and column panels share a species or a genus id range
                const int4 ri = __ldg(p.range_rows + ti), rj = __ldg(p.range_cols + tj);
                classes = (ri.x <= rj.y && rj.x <= ri.y) || (ri.z <= rj.w && rj.z <= ri.w);
            }
            // every pair of the tile is a real pair i < j < n: no range tests in the inner loop
            const bool interior = (int64_t)tj * N > (int64_t)ti * M + M - 1 && (int64_t)tj * N + N <= p.n;
            // ... and every row / column is valid at all L sites: shared == L for every pair, so the test (mism << 16 <= thr * shared)
            // is mism <= (thr * L) >> 16, with the right-hand side staged per column and kept per row
            const bool full = !COUNTS_MODE && interior && !seed && p.L >= p.min_overlap && p.panel_full[ti] && p.panel_full[2 * tj] &&
                              p.panel_full[2 * tj + 1] && (int64_t)ti * M + M <= p.n;
            // some row or column of the tile has ambiguity letters: the accumulators bound the counts (see Params::amb)
            const bool tile_amb = GENERAL && !COUNTS_MODE && p.amb != nullptr && (p.panel_amb[ti] || p.panel_amb[2 * tj] || p.panel_amb[2 * tj + 1]);
            const uint32_t ambw_i = (GENERAL && !COUNTS_MODE && p.amb != nullptr && i < p.n) ? (uint32_t)__ldg(p.amb + i) : 0u;
            const uint32_t amb_i = ambw_i & AMB_MASK, oh_i = ambw_i >> AMB_SHIFT;  // ambiguity letters / expressible sites of this row
            const uint32_t amb_rmax = GENERAL ? __reduce_max_sync(0xffffffffu, amb_i) : 0u;
            uint32_t tf_row = edge_tf;
            uint32_t amb_cmax = 0;  // most ambiguity letters of any of this warp's columns
            if constexpr (!COUNTS_MODE) {
                __syncwarp();  // the previous tile's reads of my_thr are done
#pragma unroll
                for (int x = 0; x < EPI_COLS / 32; x++) {
                    const int j = j0 + g * EPI_COLS + x * 32 + lane;
                    if (seed) my_thr[x * 32 + lane] = (uint32_t)(j < p.n ? __ldg(sp.species + j) : -1);  // seed pass: the columns' species ids
                    else {
                        const uint32_t t = (want_bm && j < p.n) ? __ldcg(sp.thr + j) : 0u;  // <= 2^16
                        if (tile_amb) {
                            const uint32_t ajw = j < p.n ? (uint32_t)__ldg(p.amb + j) : 0u, aj = ajw & AMB_MASK;
                            my_thr[x * 32 + lane] = t | (aj << 20);  // amb[j] < 4096
                            // the pre-test of interior tiles (below): no pair of column j passes its column's threshold with more than
                            // (thr * S) >> 16 mismatches, S = oh[j] + amb[j] + max amb[i] >= shared' + amb[i] + amb[j] (the sites both rows
                            // express are among those column j expresses) -- stated on the raw accumulator
                            if constexpr (GENERAL) {
                                my_raw[x * 32 + lane] = ((uint32_t)(((unsigned long long)t * ((ajw >> AMB_SHIFT) + aj + amb_rmax)) >> 16) + 1u) << p.wshift;
                                amb_cmax = max(amb_cmax, aj);
                            }
                        }
                        else if (!full) my_thr[x * 32 + lane] = t;
                        else {
                            // shared == L everywhere: mism << 16 <= thr * L  <=>  mism <= (thr * L) >> 16 =: b.  The plain instantiation
                            // goes one step further and compares the RAW accumulator D = ident + W * mism (ident < W):
                            // mism <= b  <=>  D < (b + 1) * W -- no shift per pair
                            const uint32_t b = (uint32_t)(((unsigned long long)t * (uint32_t)p.L) >> 16);
                            my_thr[x * 32 + lane] = GENERAL ? b : (b + 1u) << p.wshift;
                        }
                    }
                }
                if (want_bm && i < p.n) tf_row = max(tf_row, __ldcg(sp.thr + i));
                __syncwarp();
            }
            const uint32_t brow = (uint32_t)(((unsigned long long)tf_row * (uint32_t)p.L) >> 16);
            const uint32_t brow_raw = (brow + 1u) << p.wshift;  // (b <= L < W: no overflow)
            // tiles with ambiguity letters: the same bound with L + amb[i] sites, and what K2P's forced pairs need (see the pre-test)
            if constexpr (GENERAL) amb_cmax = __reduce_max_sync(0xffffffffu, amb_cmax);
            const uint32_t brow_raw_amb = ((uint32_t)(((unsigned long long)tf_row * (oh_i + amb_i + amb_cmax)) >> 16) + 1u) << p.wshift;
            const uint32_t force_x2 = 2u * (amb_i + amb_cmax);
            const uint32_t half_l = k2p_force ? ((uint32_t)p.L + 1u) >> 1 : 0xffffffffu;  // 2 * mism >= L
            mbar_wait(&tfull[as], (uint32_t)((k >> 1) & 1));
            fence_after();
            int bad_row = 0;
            // seed pass: exact minima of this row over the columns of its diagonal block (both sides of the diagonal), by category
            unsigned long long seed_c = TDNA_INF_BITS, seed_p0 = TDNA_INF_BITS, seed_p1 = TDNA_INF_BITS;
            int seed_jc = -1, seed_j0 = -1, seed_j1 = -1;  // the columns those minima came from
            // plain instantiation: 1 - ident / shared is smallest where ident / shared is largest -- fractions compare exactly by integer
            // cross-multiplication (both below 2^12), so the fp64 division happens three times per row instead of once per pair
            uint32_t fc_i = 0, fc_s = 0, f0_i = 0, f0_s = 0, f1_i = 0, f1_s = 0;
            const int spi = (seed && i < p.n) ? __ldg(sp.species + i) : -1;
#pragma unroll 1
            for (int cbi = 0; cbi < EPI_COLS / 32; cbi++) {
                const int cb = g * (EPI_COLS / 32) + cbi;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * N + cb * 32), v);
                // exact class pairs are read from tensor memory a second time, column by column, in a ROLLED loop below: the stage stays ours
                // until then (the 32-fold unrolled loop this replaces carried a division and a table insert per copy: 3 * 10^9 warp
                // instructions and an instruction cache that never hit, 0.25 - 0.5 ms per tile with class pairs)
                const bool class_here = classes && !(GENERAL && (p.k2p || tile_amb));
                if (cbi == EPI_COLS / 32 - 1 && !class_here) {  // this warp's columns of the stage are in registers: hand them back
                    fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR) mbar_arrive_cta(&tempty[as], 0);  // the leader's MMA issuer waits for both CTAs' epilogues
                        else mbar_arrive(&tempty[as]);
                    }
                }
                const int jb = j0 + cb * 32;
#ifdef TDNA_TC_EXPERIMENTS
                if (p.hints == 8) continue;  // -DTDNA_TC_EXPERIMENTS, TDNA_TC_HINTS=8: no epilogue work at all (results are wrong) -- the MMA floor
#endif
                if constexpr (COUNTS_MODE) {
#pragma unroll
                    for (int e = 0; e < 32; e++) {
                        const uint32_t mism = v[e] >> p.wshift, ident = v[e] & wm1;
                        if (i < p.n && jb + e < p.n) {
                            tdna_counts o = {(int)(ident + mism), p.trans ? (int)mism : (int)ident, 0, 0};
                            p.counts[(size_t)i * p.n + jb + e] = o;
                        }
                    }
                } else if (seed) {
#pragma unroll
                    for (int e = 0; e < 32; e++) {  // fully unrolled: v[] must stay in registers
                        const int j = jb + e, spj = (int)my_thr[cbi * 32 + e];
                        const uint32_t mism = v[e] >> p.wshift, ident = v[e] & wm1, s = ident + mism;
                        if (j == i || j >= p.n || i >= p.n || spj < 0 || (int)s < p.min_overlap) continue;
                        if constexpr (!GENERAL) {
                            if (spj == spi && (fc_s == 0 || ident * fc_s > fc_i * s)) { fc_i = ident; fc_s = s; }
                            if (spj & 1) { if (f1_s == 0 || ident * f1_s > f1_i * s) { f1_i = ident; f1_s = s; } }
                            else if (f0_s == 0 || ident * f0_s > f0_i * s) { f0_i = ident; f0_s = s; }
                            continue;
                        }
                        // ambiguity letters: an UPPER bound of the distance is all a threshold needs (0 extra sites: the exact value)
                        const uint32_t aij = (GENERAL && p.amb != nullptr) ? amb_i + ((uint32_t)__ldg(p.amb + j) & AMB_MASK) : 0u;
                        double d;
                        if (GENERAL && p.k2p) {
                            // only n and ts + tv are known here: d_K2P <= -ln(1 - 2(P+Q)) / 2 (all differences transitions) bounds the
                            // minima from above, which is all a threshold needs; the exact minima come from the queue afterwards
                            if (mism + aij == 0) d = 0.0;
                            else {
                                const double w = dsub(1.0, ddiv(dmul(2.0, (double)(mism + aij)), (double)s));
                                if (!(w > 0.0)) continue;
                                d = dmul(-0.5, strict_log_call(w));
                            }
                        } else {
                            Counts cn = {(int)(s + aij), (int)ident, 0, 0};  // 1 - ident' / (shared' + amb) >= the pair's distance
                            d = distance_from_counts<KM_P_AMBIG>(cn, p.min_overlap);
                        }
                        if (!(d < __longlong_as_double((long long)TDNA_INF_BITS)) || d < 0.0) continue;
                        const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
                        if constexpr (GENERAL) {  // remember where each minimum came from (refined below in tiles with ambiguity letters)
                            if (spj == spi && bits < seed_c) { seed_c = bits; seed_jc = j; }
                            if (spj & 1) { if (bits < seed_p1) { seed_p1 = bits; seed_j1 = j; } }
                            else if (bits < seed_p0) { seed_p0 = bits; seed_j0 = j; }
                        } else {
                            if (spj == spi) seed_c = min(seed_c, bits);
                            if (spj & 1) seed_p1 = min(seed_p1, bits); else seed_p0 = min(seed_p0, bits);
                        }
                    }
                } else {
                    uint32_t pmask = 0, bmask = 0, smask = 0;  // slow path wanted / pair below minOverlap / valid self pair
                    const uint32_t ct_addr = smem_u32(my_thr + cbi * 32);  // explicit ld.shared (a generic LD otherwise)
                    auto lds4 = [](uint32_t addr) {
                        uint4 r;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
                        return r;
                    };
                    if (tile_amb) {
                        // the exact statement of the test for one pair: shared' <= shared <= sb, mism' <= mism
                        auto amb_pair = [&](int e, uint32_t ct) {
                            const int j = jb + e;
                            const uint32_t thr_j = ct & 0xfffffu, extra = amb_i + (ct >> 20);
                            const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1, sb = s + extra;
                            const bool inr = (j > i) & (j < p.n);
                            const bool sure = (int)s >= p.min_overlap, maybe = (int)sb >= p.min_overlap;
                            const bool pass = ((mism << 16) <= max(tf_row, thr_j) * sb) | (k2p_force & (2u * (mism + extra) >= s));
                            if (inr && maybe && (pass || !sure)) pmask |= 1u << e;  // an uncertain overlap is settled by the exact counts
                            if (inr && !maybe) bmask |= 1u << e;
                            if (j == i && i < p.n && (int)(s + amb_i) >= p.min_overlap) smask |= 1u << e;  // a row shares all its letters with itself
                        };
                        if (GENERAL && interior) {
                            // Interior tiles (every pair real, no self pair): a PRE-TEST on the raw accumulator D = ident' + W * mism' first --
                            // five instructions per pair instead of twenty-seven.  A pair can only matter to the exact test when
                            //   D < max(row bound, column bound)       (it may pass a threshold: see the column staging above), or
                            //   ident' < minOverlap                     (its overlap may be short or uncertain: shared' >= ident'), or
                            //   K2P: shared' < minOverlap, or 2 mism' + 2 (amb[i] + max amb[j]) >= shared'   (a forced pair),
                            // and only the columns some lane flags go through the exact test (a warp-uniform branch per column).
                            uint32_t m1 = 0;
                            const uint32_t cr_addr = smem_u32(my_raw + cbi * 32), mo_u = (uint32_t)max(p.min_overlap, 0);
#pragma unroll
                            for (int e4 = 0; e4 < 8; e4++) {
                                const uint4 cr = lds4(cr_addr + e4 * 16);
                                const uint32_t crs[4] = {cr.x, cr.y, cr.z, cr.w};
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    const int e = e4 * 4 + u;
                                    bool hit = v[e] < max(brow_raw_amb, crs[u]);
                                    if (k2p_force) {
                                        const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1;
                                        hit |= (s < mo_u) | (2u * mism + force_x2 >= s);
                                    } else hit |= (v[e] & wm1) < mo_u;
                                    if (hit) m1 |= 1u << e;
                                }
                            }
                            const uint32_t uni = __reduce_or_sync(0xffffffffu, m1);
                            if (uni) {
#pragma unroll
                                for (int e = 0; e < 32; e++)
                                    if (uni & (1u << e)) {
                                        uint32_t ct;
                                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ct) : "r"(ct_addr + e * 4));
                                        amb_pair(e, ct);
                                    }
                            }
                        } else {
#pragma unroll
                            for (int e4 = 0; e4 < 8; e4++) {
                                const uint4 ct = lds4(ct_addr + e4 * 16);
                                const uint32_t cts[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                                for (int u = 0; u < 4; u++) amb_pair(e4 * 4 + u, cts[u]);
                            }
                        }
                    } else if (full) {
#pragma unroll
                        for (int e4 = 0; e4 < 8; e4++) {
                            const uint4 ct = lds4(ct_addr + e4 * 16);
                            const uint32_t cts[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int e = e4 * 4 + u;
                                if constexpr (!GENERAL) {
                                    if (v[e] < max(brow_raw, cts[u])) pmask |= 1u << e;
                                } else {
                                    const uint32_t mism = v[e] >> p.wshift;
                                    if ((mism <= max(brow, cts[u])) | (mism >= half_l)) pmask |= 1u << e;
                                }
                            }
                        }
                    } else if (interior) {
#pragma unroll
                        for (int e4 = 0; e4 < 8; e4++) {
                            const uint4 ct = lds4(ct_addr + e4 * 16);
                            const uint32_t cts[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int e = e4 * 4 + u;
                                const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1;  // s = ident + mism
                                const bool ok = (int)s >= p.min_overlap;
                                const bool pass = ((mism << 16) <= max(tf_row, cts[u]) * s) | (k2p_force & (2u * mism >= s));
                                if (ok && pass) pmask |= 1u << e;
                                if (!ok) bmask |= 1u << e;
                            }
                        }
                    } else {
#pragma unroll
                        for (int e4 = 0; e4 < 8; e4++) {
                            const uint4 ct = lds4(ct_addr + e4 * 16);
                            const uint32_t cts[4] = {ct.x, ct.y, ct.z, ct.w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int e = e4 * 4 + u, j = jb + e;
                                const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1;
                                const bool inr = (j > i) & (j < p.n);
                                const bool ok = (int)s >= p.min_overlap;
                                const bool pass = ((mism << 16) <= max(tf_row, cts[u]) * s) | (k2p_force & (2u * mism >= s));
                                if (inr && ok && pass) pmask |= 1u << e;
                                if (inr && !ok) bmask |= 1u << e;
                                if (j == i && i < p.n && ok) smask |= 1u << e;
                            }
                        }
                    }
                    if (!want_bm && edge_tf == 0u) pmask = bmask = smask = 0;  // a class-only launch: no per-row output, no per-row state
                    if (smask && !seed) atomicOr(sp.flags + i, TDNA_ROW_SELF_VALID);
                    if (bmask && !seed) {
                        bad_row += __popc(bmask);
#pragma unroll
                        for (int e = 0; e < 32; e++)
                            if (bmask & (1u << e)) atomicAdd(sp.inval + jb + e, 1);
                    }
                    // class pairs whose exact counts must come from the bit-planes (K2P; tiles with ambiguity letters) are not computed
                    // here -- a 200-load chain per pair, one lane at a time, stalled the whole tile -- but queued with bit 31 of the
                    // row index set; resolve_emit_kernel emits them, one thread per entry
                    uint32_t cmask = 0;
                    const bool classes_deferred = classes && !class_here;
                    if (classes_deferred && i < p.n) {
                        const int spi_ = sp.species[i], gi_ = sp.genus[i];
#pragma unroll
                        for (int e = 0; e < 32; e++) {
                            const int j = jb + e;
                            if (j <= i || j >= p.n) continue;
                            const bool hit = ((sp.want & TDNA_WANT_INTRA) && spi_ == sp.species[j]) || ((sp.want & TDNA_WANT_INTER) && gi_ == sp.genus[j]);
                            if (!hit) continue;
                            const uint32_t mism = v[e] >> p.wshift, s = v[e] - mism * wm1;
                            if ((int)(s + (tile_amb ? amb_i + ((uint32_t)__ldg(p.amb + j) & AMB_MASK) : 0u)) >= p.min_overlap) cmask |= 1u << e;
                        }
                    }
                    pmask |= cmask;
                    // entries of a launch without per-row output serve the classes only (bit 31 of the column word): the launch that
                    // produces the rows has queued the same pairs for itself
                    const int conly = (!want_bm && edge_tf == 0u) ? (int)0x80000000 : 0;
                    if (!seed) {
                        // bulk pass: no fp64, no threshold traffic here -- the pairs that pass the seeded thresholds are queued
                        // (one warp-aggregated atomic) and resolved after the pass (resolve_*_kernel)
                        if (__any_sync(0xffffffffu, pmask != 0)) {
                            const int cnt = __popc(pmask);
                            int incl = cnt;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += t;
                            }
                            const int total = __shfl_sync(0xffffffffu, incl, 31);
                            unsigned long long base = 0;
                            if (lane == 0) base = atomicAdd(sp.qn, (unsigned long long)total);
                            base = __shfl_sync(0xffffffffu, base, 0);
                            unsigned long long w = base + (unsigned long long)(incl - cnt);
                            if ((long long)(base + total) <= sp.qcap) {
#pragma unroll
                                for (int e = 0; e < 32; e++)
                                    if (pmask & (1u << e)) { sp.q[w] = make_uint4((uint32_t)(i | (int)((cmask >> e) << 31)), (uint32_t)((jb + e) | conly), v[e], 0u); w++; }
                            } else if (lane == 0) *sp.overflow = 1;
                        }
                        pmask = 0;
                    }

                    if (class_here) {  // exact accumulators: one fp64 division per class pair, right here
                        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * N + cb * 32);
                        const int spi_ = i < p.n ? sp.species[i] : -1, gi_ = i < p.n ? sp.genus[i] : -1;
                        const bool hist_mode = sp.hkey[0] || sp.hkey[1];
#pragma unroll 1
                        for (int e = 0; e < 32; e++) {
                            __syncwarp();  // lanes that skipped out of the previous column's body rejoin here
                            const uint32_t val = tmem_ld1(tcol + (uint32_t)e);  // (warp-collective: every lane takes part)
                            const int j = jb + e;
                            const uint32_t mism = val >> p.wshift, ident = val & wm1, s = ident + mism;
                            if (j <= i || j >= p.n || (int)s < p.min_overlap) continue;
                            const bool hit = ((sp.want & TDNA_WANT_INTRA) && spi_ == sp.species[j]) || ((sp.want & TDNA_WANT_INTER) && gi_ == sp.genus[j]);
                            if (!hit) continue;
                            if (!hist_mode) {
                                if (p.trans) class_emit<KM_TRANS>(sp, i, j, mism | (s << 16), 0u, p.min_overlap);
                                else class_emit<KM_P_AMBIG>(sp, i, j, ident | (s << 16), 0u, p.min_overlap);
                                continue;
                            }
                            // histogram mode: the pushes of this tile go through the warp's shared-memory table first
                            int t0 = 0, t1 = 0;
                            if ((sp.want & TDNA_WANT_INTRA) && spi_ >= 0 && spi_ == sp.species[j]) t0 = 1 + (sp.full[i] == sp.full[j] ? 1 : 0);
                            if ((sp.want & TDNA_WANT_INTER) && gi_ == sp.genus[j] && sp.epi[i] != sp.epi[j]) t1 = 1;
                            if (!(t0 | t1)) continue;
                            const double d = exact_acc_distance(p.trans, ident, s, p.min_overlap);
                            if (t0) { if (sp.hkey[0] && shist_add(my_keys, my_cnts, WTAB - 1, 0, d, t0)) wtab_used = true; else class_push(sp, 0, d, t0); }
                            if (t1) { if (sp.hkey[1] && shist_add(my_keys, my_cnts, WTAB - 1, 1, d, t1)) wtab_used = true; else class_push(sp, 1, d, t1); }
                        }
                        if (cbi == EPI_COLS / 32 - 1) {  // now the stage can go back
                            fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if constexpr (PAIR) mbar_arrive_cta(&tempty[as], 0);
                                else mbar_arrive(&tempty[as]);
                            }
                        }
                    }
                }
            }
            if constexpr (!COUNTS_MODE) {
                if (__any_sync(0xffffffffu, wtab_used)) {  // this tile's class pushes: one global reduction per distinct value
                    __syncwarp();
                    shist_flush(sp, my_keys, my_cnts, WTAB, lane, 32);
                    __syncwarp();
                    wtab_used = false;
                }
                if (bad_row) atomicAdd(sp.inval + i, bad_row);
                if (!GENERAL && seed && i < p.n) {
                    auto frac_bits = [&](uint32_t ident, uint32_t sh) -> unsigned long long {
                        if (sh == 0) return TDNA_INF_BITS;
                        return (unsigned long long)__double_as_longlong(exact_acc_distance(p.trans, ident, sh, p.min_overlap));
                    };
                    seed_c = frac_bits(fc_i, fc_s);
                    seed_p0 = frac_bits(f0_i, f0_s);
                    seed_p1 = frac_bits(f1_i, f1_s);
                }
                if (GENERAL && seed && i < p.n && tile_amb) {
                    // ambiguity letters: the minima above are upper BOUNDS, loose by (amb[i] + amb[j]) / shared -- 0.1 at 4 % ambiguity letters,
                    // and the bulk would queue every congener.  The exact distance of the pair each bound came from (three pairs per row,
                    // from the bit-planes) is a valid and tight upper bound of the category's minimum.
                    seed_c = exact_seed_bits(p, i, seed_jc);
                    seed_p0 = exact_seed_bits(p, i, seed_j0);
                    seed_p1 = exact_seed_bits(p, i, seed_j1);
                }
                if (seed && i < p.n) {
                    if (seed_c < TDNA_INF_BITS) atomicMin(sp.mC + i, seed_c);
                    if (seed_p0 < TDNA_INF_BITS) atomicMin(sp.mPar + 2 * (int64_t)i, seed_p0);
                    if (seed_p1 < TDNA_INF_BITS) atomicMin(sp.mPar + 2 * (int64_t)i + 1, seed_p1);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if constexpr (CL >= 2) cluster_sync_all();  // no CTA leaves while its peer may still write into it
    if (warp == 1) {
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- after the bulk pass: the queue -> minima -> final thresholds -> candidate records --------------------------
// every queued pair is a pair i < j that is valid or (ambiguity letters) may be; value = raw accumulator
struct ResolveArgs {
    int wshift;
    uint32_t wm1;
    int min_overlap;
    const uint32_t *planes;  // non-null: exact counts of the pair from the bit-planes (K2P; alignments with ambiguity letters)
    int W32;
    int k2p;
    int trans;   // transversions only (exact accumulators: d = transv / shared)
    int cached;  // the exact counts of the rows' entries sit in qv / qx already (resolve_min_kernel put them there): no second walk of the planes
};
// exact counts <-> two queue words: (c1 | shared << 16, c2 | c3 << 16), every count <= L <= 65535
__device__ __forceinline__ double cached_distance(const ResolveArgs &ra, uint32_t w0, uint32_t w1, uint32_t &lhs, uint32_t &den) {
    Counts c = {(int)(w0 >> 16), (int)(w0 & 0xffffu), (int)(w1 & 0xffffu), (int)(w1 >> 16)};
    if (ra.k2p) {
        lhs = (uint32_t)(c.c2 + c.c3) << 16;
        den = (uint32_t)c.c1;
        return distance_from_counts<KM_K2P>(c, ra.min_overlap);
    }
    lhs = (uint32_t)(c.shared - c.c1) << 16;
    den = (uint32_t)c.shared;
    return distance_from_counts<KM_P_AMBIG>(c, ra.min_overlap);
}
__device__ __forceinline__ double queued_distance(const ResolveArgs &ra, int i, int j, uint32_t val, uint32_t &lhs, uint32_t &den, uint32_t *w0 = nullptr,
                                                  uint32_t *w1 = nullptr) {
    if (ra.planes) {
        const Counts c = ra.k2p ? pair_counts_global<KM_K2P>(ra.planes, i, j, ra.W32) : pair_counts_global<KM_P_AMBIG>(ra.planes, i, j, ra.W32);
        const uint32_t x0 = (uint32_t)c.c1 | ((uint32_t)c.shared << 16), x1 = (uint32_t)c.c2 | ((uint32_t)c.c3 << 16);
        if (w0) { *w0 = x0; *w1 = x1; }
        return cached_distance(ra, x0, x1, lhs, den);  // (p: -1.0 when the queued overlap was uncertain and is too short)
    }
    const uint32_t mism = val >> ra.wshift, ident = val & ra.wm1, s = ident + mism;
    lhs = mism << 16;
    den = s;
    return exact_acc_distance(ra.trans, ident, s, ra.min_overlap);
}
constexpr int RTAB = 1024;  // slots of the resolve kernel's per-block class-histogram table
// a queued class pair (bit 31 of the row index): PairwiseDistribution's pushes for it, with the exact distance d >= 0 (or NaN / +Inf)
__device__ __forceinline__ void class_emit_value(const StreamParams &sp, int i, int j, double d, unsigned long long *s_keys, unsigned int *s_cnts) {
    int times[2] = {0, 0};
    if (sp.want & TDNA_WANT_INTRA) {
        const int si = sp.species[i];
        if (si >= 0 && si == sp.species[j]) times[0] = 1 + (sp.full[i] == sp.full[j] ? 1 : 0);
    }
    if (sp.want & TDNA_WANT_INTER) {
        if (sp.genus[i] == sp.genus[j] && sp.epi[i] != sp.epi[j]) times[1] = 1;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (!times[k]) continue;
        if (s_keys && sp.hkey[k] && shist_add(s_keys, s_cnts, RTAB - 1, k, d, times[k])) continue;  // (RTAB is defined with the kernel below)
        class_push(sp, k, d, times[k]);
    }
}
__global__ void __launch_bounds__(256) resolve_min_kernel(StreamParams sp, ResolveArgs ra) {
    if (*sp.overflow) return;  // the queue has holes when a push did not fit: the caller falls back to the dense path
    const long long nq = min((long long)*sp.qn, sp.qcap);
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nq; r += (long long)gridDim.x * blockDim.x) {
        const uint4 ent = sp.q[r];
        if ((int)ent.y < 0) continue;  // a class-only entry
        const int i = (int)(ent.x & 0x7fffffffu), j = (int)ent.y;
        uint32_t lhs, den, w0 = 0, w1 = 0;
        const double d = queued_distance(ra, i, j, ent.z, lhs, den, &w0, &w1);
        if (ra.planes) sp.q[r] = make_uint4(ent.x, ent.y, w0, w1);  // the exact counts, for resolve_emit_kernel
        if (!(d < __longlong_as_double((long long)TDNA_INF_BITS)) || d < 0.0) continue;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
        const int spi = sp.species[i], spj = sp.species[j];
        if (spj >= 0 && lhs <= sp.thr[i] * den) { if (spj == spi) atomicMin(sp.mC + i, bits); atomicMin(sp.mPar + 2 * (int64_t)i + (spj & 1), bits); }
        if (spi >= 0 && lhs <= sp.thr[j] * den) { if (spi == spj) atomicMin(sp.mC + j, bits); atomicMin(sp.mPar + 2 * (int64_t)j + (spi & 1), bits); }
    }
}
__global__ void __launch_bounds__(256) resolve_emit_kernel(StreamParams sp, ResolveArgs ra) {
    __shared__ unsigned long long s_keys[RTAB];
    __shared__ unsigned int s_cnts[RTAB];
    const bool hist = sp.hkey[0] || sp.hkey[1];
    if (hist) {
        for (int s = threadIdx.x; s < RTAB; s += 256) { s_keys[s] = TDNA_HIST_EMPTY; s_cnts[s] = 0; }
        __syncthreads();
    }
    if (*sp.overflow) return;
    const long long nq = min((long long)*sp.qn, sp.qcap);
    const bool want_bm = (sp.want & TDNA_WANT_BEST_MATCH) != 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nq; r += (long long)gridDim.x * blockDim.x) {
        const uint4 ent = sp.q[r];
        const int qi_raw = (int)ent.x, qj_raw = (int)ent.y, i = qi_raw & 0x7fffffff, j = qj_raw & 0x7fffffff;
        uint32_t lhs, den;
        const double d = (ra.cached && qj_raw >= 0) ? cached_distance(ra, ent.z, ent.w, lhs, den) : queued_distance(ra, i, j, ent.z, lhs, den);
        if (qi_raw < 0 && !(d < 0.0)) class_emit_value(sp, i, j, d, hist ? s_keys : nullptr, s_cnts);  // a deferred class pair (valid: NaN / +Inf are kept, -1 is dropped)
        if (qj_raw < 0) continue;  // a class-only entry: the rows' launch queued this pair for itself
        if (d < 0.0) {  // queued with an uncertain overlap (ambiguity letters), and the exact overlap is below minOverlap: an invalid pair
            atomicAdd(sp.inval + i, 1);
            atomicAdd(sp.inval + j, 1);
            continue;
        }
        if ((sp.want & TDNA_WANT_EDGES) && d >= 0 && d <= sp.edge_t) {
            const unsigned long long pos = atomicAdd(sp.nedge, 1ull);
            if (sp.ei && (long long)pos < sp.edge_cap) { sp.ei[pos] = i; sp.ej[pos] = j; sp.ed[pos] = d; }
        }
        if (!(d < __longlong_as_double((long long)TDNA_INF_BITS))) {
            atomicOr(sp.flags + i, TDNA_ROW_NONFINITE);
            atomicOr(sp.flags + j, TDNA_ROW_NONFINITE);
            continue;
        }
        if (!want_bm) continue;
        const bool ei = lhs <= sp.thr[i] * den, ej = lhs <= sp.thr[j] * den;
        if (!(ei || ej)) continue;
        // one atomic per warp: positions from ballots over the lanes that got here together
        const unsigned amask = __activemask();
        const int lane = threadIdx.x & 31, mine = ei + ej;
        const unsigned bi = __ballot_sync(amask, ei), bj = __ballot_sync(amask, ej), lt = (1u << lane) - 1u;
        const int total = __popc(bi) + __popc(bj), below = __popc(bi & lt) + __popc(bj & lt);
        const int leader = __ffs((int)amask) - 1;
        unsigned long long pos = 0;
        if (lane == leader) pos = atomicAdd(sp.ncand, (unsigned long long)total);
        pos = __shfl_sync(amask, pos, leader) + (unsigned long long)below;
        if ((long long)(pos + mine) > sp.cap) { *sp.overflow = 1; continue; }
        unsigned long long w = pos;
        if (ei) { sp.cq[w] = i; sp.cj[w] = j; sp.cd[w] = d; atomicAdd(sp.cnt + i, 1); w++; }
        if (ej) { sp.cq[w] = j; sp.cj[w] = i; sp.cd[w] = d; atomicAdd(sp.cnt + j, 1); }
    }
    if (hist) {  // the block's class pushes: one global reduction per distinct value
        __syncthreads();
        shist_flush(sp, s_keys, s_cnts, RTAB, threadIdx.x, 256);
    }
}

}  // namespace tc
}  // namespace tdna
