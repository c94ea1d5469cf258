// C ABI of the engine (include/taxondna_cuda.h): context/alignment lifetime, band scheduling of the
// N x N pair space through HBM, kernel launches.  No CPU arithmetic path exists here by design.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "tdna_kernels.cuh"
#include "tdna_tc.cuh"
#include "tdna_fasta.cuh"
#include "tdna_comm.cuh"

using namespace tdna;

namespace {
thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(TDNA_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)
#define TRY(call)              \
    do {                       \
        int rc_ = (call);      \
        if (rc_ != TDNA_OK) return rc_; \
    } while (0)

int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
}  // namespace

// Per-alignment buffers come from a stream-ordered pool of the CONTEXT'S OWN (release threshold = never: creating and destroying an
// alignment per analysis costs microseconds, not cudaMalloc/cudaFree syncs).  The device's default pool -- which other users of
// cudaMallocAsync in the process share -- is left as it is.
#define DMALLOC(ptr, bytes) cudaMallocFromPoolAsync((void **)(ptr), (bytes), c->pool, c->stream)
#define DFREE(ptr) cudaFreeAsync((ptr), c->stream)

struct tdna_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;
    int64_t budget = 64ll << 30;
    int64_t launches = 0;
    int sm_count = 148;
    tdna_progress_fn progress = nullptr;
    void *progress_user = nullptr;
    volatile int cancel = 0;
    unsigned long long *d_counter = nullptr;
    // the N x N (or band x N) distance buffer is pooled per context: alignments come and go (one per
    // analysis in the host layer) but 10s of GB of cudaMalloc/cudaFree per call would dominate e2e.
    double *d_mat = nullptr;
    int64_t mat_cap = 0;              // doubles
    const void *mat_owner = nullptr;  // alignment whose whole-range matrix currently sits in d_mat
    // the most recent tile pass (tdna_last_tile_pass_ms): seed launch = [k0, s1], bulk launch = [b0, k1]; what lies between them
    // (the MIN all-reduce of the seeded minima when several GPUs share the pass) is not tile work
    cudaEvent_t ev_k0 = nullptr, ev_s1 = nullptr, ev_b0 = nullptr, ev_k1 = nullptr;
    bool ev_valid = false, ev_split = false;
    cudaEvent_t ev_ph[12] = {};       // phase marks of the most recent multi-GPU pass (tdna_multi_phase_ms)
    int n_ph = 0;
    cudaStream_t copy_stream = nullptr;  // chunked upload (tdna_alignment_create): the H2D copies; pack + encode follow on `stream`
    cudaEvent_t ev_copy[16] = {};
    Comm comm;                        // tdna_comm_init: NCCL communicator of this context (one rank per GPU)
    int64_t *h_status = nullptr;      // pinned: the status words a multi-GPU pass reads back at its one and only host sync
    int opt_best_match_path = TDNA_PATH_AUTO;
    int opt_tile_cn = 0;  // 0 = automatic
    int opt_count_kernel = TDNA_KERNEL_AUTO;
    uint32_t opt_hist_slots = 1u << 18;
    int64_t opt_segment_items = 0;  // 0 = automatic (segments of ~30 ms)
    // tdna_analyze_host: row chunks of the pass pipelined with the upload; 0 (default): not pipelined -- measured on 50,000 x 658:
    // 3.55 ms without, 3.57 - 3.78 ms with 3 - 8 chunks: the upload is 0.65 ms, a seed launch per chunk and the
    // encode kernels waiting for the persistent tile CTAs cost as much as the overlap gains
    int opt_pipeline_chunks = 0;
    int opt_unit_items = 0;  // work items per dealing unit of a multi-GPU tile pass; 0 = default (64 tile pairs)
    int opt_cluster = 2;  // CTAs per cluster of count kernel (b)'s bulk launch: 2 (pairs sharing the b-block) or 4 (2 x 2 tiles sharing a- and b-blocks)
    int max_quads = -1;   // resident clusters of four (cudaOccupancyMaxActiveClusters), queried once
    // records / queue entries per row the streaming buffers of NEW alignments start with: raised to what a pass on this context
    // needed (a host that creates an alignment per analysis would otherwise overflow and retry every time)
    double hint_cand_per_row = 64, hint_queue_per_row = 256;
    // what the previous list built on this context looked like to count kernel (b) (-1: none yet): a pipelined tdna_analyze_host
    // starts its region launches on that assumption instead of waiting for the first chunk's flags, and checks it at the end
    int hint_has_amb = -1, hint_dims5 = -1;
    cudaStream_t enc_stream = nullptr;   // pipelined tdna_analyze_host: pack + encode of the arriving chunks, beside the tile launches
    cudaEvent_t ev_enc[16] = {};
    double hc_t[12] = {};            // host wall-clock marks of the most recent tdna_analyze_host call (tdna_host_call_phase_ms)
    int64_t pipelined_passes = 0;   // tdna_analyze_host calls whose pass ran pipelined with the upload (tdna_pipelined_passes)
};

static inline void hc_mark(tdna_ctx *c, int k) {
    c->hc_t[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct tdna_aln {
    tdna_ctx *ctx = nullptr;
    int64_t n = 0, npad = 0;
    int L = 0, W = 0;
    int64_t sym_stride = 0;
    uint8_t *d_sym = nullptr;
    int32_t *d_len = nullptr;
    uint32_t *d_planes = nullptr;
    size_t planes_words = 0;
    int mode = TDNA_PDM_UNCORRECTED, min_overlap = 300, ambig = 1, kmode = KM_P_AMBIG, P = 0;
    bool packed = false;
    int32_t *d_species = nullptr, *d_genus = nullptr, *d_epi = nullptr, *d_full = nullptr;
    bool has_labels = false;
    int64_t row_begin = 0, row_end = 0;
    double *d_mat = nullptr;  // == ctx->d_mat once ensure_mat ran
    int64_t ld = 0;
    bool mat_valid = false;  // ctx->d_mat holds the whole row range of THIS alignment for the current config
    tdna_row_result *d_res = nullptr;
    int64_t *d_prefix = nullptr;
    int64_t prefix_cap = 0;
    void *d_scratch = nullptr;  // grow-only staging for host-destined compaction outputs
    size_t scratch_cap = 0;
    // streaming Best Match (tdna_kernels.cuh "stream epilogue")
    uint8_t *d_needc = nullptr;
    int4 *d_panel_range = nullptr, *d_range_rows = nullptr, *d_range_cols = nullptr;  // per 64-row panel; per 128-row / 256-column tile panel
    StreamParams sp = {};
    bool stream_alloc = false;
    void *d_stream_u64 = nullptr;  // ncand + overflow
    int64_t *d_sprefix[2] = {nullptr, nullptr};
    int64_t sprefix_tiles[2] = {0, 0};
    int sprefix_cn = 0;
    long long *d_ptr = nullptr;  // CSR row pointers [n + 1]
    long long *d_scan = nullptr; // block sums of the row-pointer scan
    std::vector<int32_t> h_species, h_genus;   // host copies of the species / genus ids (seed list of count kernel (b); class reach)
    int64_t class_reach = -1;                  // max j - i over the class pairs i < j the labels allow (-1: not computed yet)
    int64_t tcprefix_reach = -1, tcprefix_class_tiles = 0;  // d_tcprefix[0]: the slot numbering of class-only launches, built for this reach
    int2 *d_seedlist = nullptr;
    int seed_pairs = 0;
    bool seedlist_valid = false;
    long long need_cand = 0, need_queue = 0;  // demand of the last pass that overflowed
    bool stream_broken = false;               // a buffer could not be regrown: the streaming state is unusable
    int *d_cursor = nullptr;
    int32_t *d_lj = nullptr;
    double *d_ld = nullptr;
    long long list_cap = 0;
    long long last_ncand = 0, last_ncls[2] = {0, 0}, last_nedge = 0, last_queue = 0;
    int last_kernel = 0;  // TDNA_KERNEL_* of the most recent streaming pass
    bool avoid_tensor = false;  // AUTO: kernel (b)'s queue overflowed on this alignment (static thresholds too loose): use kernel (a)
    // count kernel (b): one-hot int8 operands (tdna_tc.cuh)
    uint8_t *d_tcA = nullptr, *d_tcB = nullptr, *d_tcfull = nullptr, *d_tcpanel_amb = nullptr;
    int *d_tcflag = nullptr;       // [0] a symbol outside the alphabet met while encoding, [1] some IUPAC ambiguity letter (or dropped gap),
                                   // [2] internal gap sites the four-dimension code of uncorrected p left out
    bool tc_dims5 = false;         // uncorrected p: the list has too many internal gaps for the four-dimension code
    tc::EncodeValues tc_values = {};
    int32_t *d_tcamb = nullptr;  // per row: sites holding a letter the operands cannot express (nullptr state: see tc_has_amb)
    bool tc_has_amb = false;
    int tc_state = 0;  // 0 = not built, 1 = ready, -1 = not eligible
    int tc_nchunks = 0, tc_kmode = -1, tc_dims = 5;
    uint32_t tc_W = 0;
    CUtensorMap tc_mapAh;  // CL = 4: half an a-block
    CUtensorMap tc_mapA, tc_mapB, tc_mapBh;  // [.., 256 x uint32] views of d_tcA / d_tcB, boxes of one operand block (Bh: half a b-block)
    bool tc_maps = false, tc_pair = false;
    int tc_inner = 256;
    int64_t *d_tcprefix[2] = {nullptr, nullptr};
    int64_t tcprefix_tiles[2] = {0, 0};
    int64_t *d_tcprefix4 = nullptr;  // CL = 4: quads (2 x 2 tiles) before each band
    int64_t tcprefix4_total = 0;
    // in-library multi-GPU pass (tdna_comm_init): exchange segments, reduced per-row state, the all-gathered row records
    long long *x_send = nullptr, *x_recv = nullptr, *x_state = nullptr;
    int64_t x_seg = 0;        // records per exchange segment (same on every rank: derived from n, the rank count and x_scale only)
    int x_scale = 1;          // doubled on every rank alike when some segment overflowed (the status word is all-reduced)
    int x_nranks = 0;
    int64_t res_rows = 0;     // records d_res holds (n, or nranks * rows_per_part once a multi-GPU pass ran)
    long long local_ncls[2] = {0, 0}, local_nedge = 0;  // this rank's share of the class distances / edges of the last pass
    // class histograms (tdna_class_histogram): open-addressing tables keyed by the fp64 bits of the distance
    unsigned long long *d_hkey[2] = {nullptr, nullptr}, *d_hcnt[2] = {nullptr, nullptr};
    uint32_t hist_size = 1u << 18;   // slots per table (a power of two; grown x4 when a pass fills one)
    uint32_t hist_alloc[2] = {0, 0};
    unsigned long long *d_hctr = nullptr;  // [0] table-full flag (int), [1] entries, [2] pushes (compaction counters)
    tdna_hist_entry *d_hent = nullptr;     // staging of compacted entries
    size_t hent_cap = 0;
    // tdna_row_extremes: member lists by genus (slot of each row's genus, start of each slot, rows ascending), built from the labels
    int32_t *d_gslot = nullptr, *d_gmem = nullptr;
    int64_t *d_gstart = nullptr;
    bool genus_lists_valid = false;
    std::vector<int2> h_seedlist;  // host copy of the seed list (the pipelined pass re-orders it by row chunk)
    // the pass pipelined with the upload (tdna_analyze_host, pipeline_*): per-launch overrides consulted by launch_tc_cl
    const int64_t *pl_reg_start = nullptr, *pl_reg_prefix = nullptr;  // device: the region's runs of slot pairs
    int pl_reg_n = 0;
    int64_t pl_reg_items = 0;
    const int2 *pl_seed_list = nullptr;
    int pl_seed_pairs = -1;  // -1: no override
    int pl_general = -1;     // -1: the instantiation follows tc_has_amb; 0 / 1: decided from the first chunk's flags
    void *d_plbuf = nullptr;
};

namespace {

void build_lut(uint16_t *lut) {
    for (int i = 0; i < 256; i++) lut[i] = 0x8000;
    auto set = [&](char ch, int mask) {  // mask: A=1 C=2 T=4 G=8 (Sequence.getint)
        uint32_t b = SB_V | SB_LET;
        if (mask & 1) b |= SB_A;
        if (mask & 2) b |= SB_C;
        if (mask & 4) b |= SB_T;
        if (mask & 8) b |= SB_G;
        lut[(unsigned char)ch] = (uint16_t)b;
    };
    set('A', 1); set('C', 2); set('T', 4); set('G', 8);
    set('R', 9); set('Y', 6); set('K', 12); set('M', 3); set('S', 10); set('W', 5);
    set('B', 14); set('D', 13); set('H', 7); set('V', 11); set('N', 15);
    // purine {A,G,R} / pyrimidine {C,T,Y} classes and the K2P within-class code
    lut['A'] |= SB_U | SB_PU;
    lut['G'] |= SB_U | SB_PU | SB_C0;
    lut['R'] |= SB_U | SB_PU | SB_C1;
    lut['C'] |= SB_U;
    lut['T'] |= SB_U | SB_C0;
    lut['Y'] |= SB_U | SB_C1;
    lut['-'] = SB_D | SB_V;
    lut['_'] = 0;
    lut['?'] = 0;
}

template <int KM> PlaneBits plane_bits() {
    PlaneBits pb = {};
    for (int i = 0; i < ModeTraits<KM>::P; i++) pb.b[i] = ModeTraits<KM>::bits[i];
    return pb;
}

// mode -> plane set; (re)allocates the planes.  The pack kernel itself runs in pack() (whole alignment) or chunk by chunk (aln_build)
int pack_prepare(tdna_aln *a, PlaneBits *pb_out) {
    tdna_ctx *c = a->ctx;
    int km;
    if (a->mode == TDNA_PDM_K2P) km = KM_K2P;
    else if (a->mode == TDNA_PDM_TRANS_ONLY) km = KM_TRANS;
    else km = a->ambig ? KM_P_AMBIG : KM_P_NOAMBIG;
    PlaneBits pb;
    int P;
    switch (km) {
        case KM_P_AMBIG: pb = plane_bits<KM_P_AMBIG>(); P = ModeTraits<KM_P_AMBIG>::P; break;
        case KM_P_NOAMBIG: pb = plane_bits<KM_P_NOAMBIG>(); P = ModeTraits<KM_P_NOAMBIG>::P; break;
        case KM_K2P: pb = plane_bits<KM_K2P>(); P = ModeTraits<KM_K2P>::P; break;
        default: pb = plane_bits<KM_TRANS>(); P = ModeTraits<KM_TRANS>::P; break;
    }
    size_t words = (size_t)a->npad * P * a->W;
    if (words > a->planes_words) {
        if (a->d_planes) CU(DFREE(a->d_planes));
        a->d_planes = nullptr;
        CU(DMALLOC(&a->d_planes, words * 4));
        a->planes_words = words;
    }
    a->kmode = km;
    a->P = P;
    a->mat_valid = false;
    *pb_out = pb;
    return TDNA_OK;
}

int pack(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    PlaneBits pb;
    TRY(pack_prepare(a, &pb));
    const int P = a->P;
    int *d_err = reinterpret_cast<int *>(c->d_counter);
    CU(cudaMemsetAsync(d_err, 0, 8, c->stream));
    int64_t total = a->npad * a->W;
    pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(a->d_sym, a->d_len, a->n, 0, a->npad, a->sym_stride, a->W, P, a->L, pb,
                                                                       a->d_planes, d_err);
    c->launches++;
    CU(cudaGetLastError());
    int err = 0;
    CU(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (err) return fail(TDNA_E_SYMBOL, "a symbol outside the normalised alphabet ACGT RYKMSWBDHVN - _ ?");
    a->packed = true;
    return TDNA_OK;
}

// The StreamParams of the next launch (a __grid_constant__ kernel parameter: nothing is shared between alignments or contexts on one
// device).  The seed pass (diagonal blocks) only tightens thresholds and wants Best Match only.
static StreamParams pass_params(const tdna_aln *a, bool seed) {
    StreamParams sp = a->sp;
    if (seed) sp.want = TDNA_WANT_BEST_MATCH | TDNA_WANT_SEED;
    return sp;
}
// A long tile pass goes out in SEGMENTS of its work list (tens of milliseconds each).  Between segments of a single-GPU pass the
// host waits for the segment, honours tdna_cancel and reports progress (DelayCallback.delay, Common/DelayCallback.java:34-56) --
// a 10^6-row pass takes a second.  Passes shared by a communicator are enqueued whole (a cancel there is honoured between passes:
// every rank would have to take the same decision at the same segment).
static int segment_boundary(tdna_ctx *c, int64_t done, int64_t total) {
    if (c->comm.active()) return TDNA_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (c->cancel) { c->cancel = 0; return fail(TDNA_E_CANCELLED, "cancelled"); }
    if (c->progress) c->progress(done, total, c->progress_user);
    return TDNA_OK;
}
int ensure_tc(tdna_aln *a);
static int ensure_class_reach(tdna_aln *a);
static int ensure_class_prefix(tdna_aln *a);
template <bool COUNTS_MODE> int launch_tc(tdna_aln *a, int phase, int part, int nparts, tdna_counts *counts);

template <int KM, int CN, int EPI> int launch_kernel(tdna_ctx *c, const TileParams &p, const StreamParams &sp = StreamParams()) {
    using Cfg = TileCfg<KM, CN>;
    auto kern = count_tile_kernel<KM, CN, EPI>;
    static bool attr_set[64] = {};  // per device: function attributes live in the device's context (one process may drive several)
    if (!attr_set[c->device & 63]) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set[c->device & 63] = true;
    }
    int64_t slots = (int64_t)c->sm_count * Cfg::MIN_CTAS;
    int64_t grid = p.num_tiles < slots ? p.num_tiles : slots;
    if (grid < 1) return TDNA_OK;
    kern<<<(unsigned)grid, TILE_THREADS, Cfg::SMEM, c->stream>>>(p, sp);
    c->launches++;
    CU(cudaGetLastError());
    return TDNA_OK;
}

template <int KM, int CN, bool WC> int launch_tiles_t(tdna_aln *a, int64_t b0, int64_t b1, bool symmetric, double *out, int64_t ld, tdna_counts *counts) {
    using Cfg = TileCfg<KM, CN>;
    tdna_ctx *c = a->ctx;
    TileParams p = {};
    p.planes = a->d_planes;
    p.W = a->W;
    p.n = a->n;
    p.min_overlap = a->min_overlap;
    p.row_begin = b0;
    p.row_end = b1;
    p.symmetric = symmetric ? 1 : 0;
    p.sched = symmetric ? SCHED_SYMMETRIC : SCHED_RECT;
    p.first_col_extra = 0;
    p.t_offset = 0;
    p.t_stride = 1;
    p.nct = (int)((a->n + Cfg::TN - 1) / Cfg::TN);
    int nrt = (int)((b1 - b0 + TM - 1) / TM);
    p.nrt = nrt;
    p.tile_prefix = nullptr;
    if (symmetric) {
        std::vector<int64_t> prefix(nrt + 1);
        prefix[0] = 0;
        for (int t = 0; t < nrt; t++) {
            int64_t first = (int64_t)t * (TM / Cfg::TN);
            prefix[t + 1] = prefix[t] + (p.nct > first ? p.nct - first : 0);
        }
        if (a->prefix_cap < nrt + 1) {
            if (a->d_prefix) CU(DFREE(a->d_prefix));
            a->d_prefix = nullptr;
            CU(DMALLOC(&a->d_prefix, (size_t)(nrt + 1) * 8));
            a->prefix_cap = nrt + 1;
        }
        CU(cudaMemcpyAsync(a->d_prefix, prefix.data(), (size_t)(nrt + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));  // prefix is a stack-lifetime host buffer
        p.tile_prefix = a->d_prefix;
        p.num_tiles = prefix[nrt];
    } else {
        p.num_tiles = (int64_t)nrt * p.nct;
    }
    p.out = out;
    p.ld = ld;
    p.counts = counts;
    CU(cudaEventRecord(c->ev_k0, c->stream));
    TRY((launch_kernel<KM, CN, WC ? EPI_MATRIX_COUNTS : EPI_MATRIX>(c, p)));
    CU(cudaEventRecord(c->ev_k1, c->stream));
    c->ev_valid = true;
    c->ev_split = false;
    return TDNA_OK;
}

// Tile shape: 128 x 128 (one CTA per SM) when there are plenty of tiles; 128 x 64 with two CTAs per
// SM when the problem is small (tile quantisation) or the mode needs two accumulator words per pair.
template <int KM, bool WC> int launch_tiles_km(tdna_aln *a, int64_t b0, int64_t b1, bool symmetric, double *out, int64_t ld, tdna_counts *counts) {
    if constexpr (ModeTraits<KM>::NACC == 2) {
        return launch_tiles_t<KM, 4, WC>(a, b0, b1, symmetric, out, ld, counts);
    } else {
        int64_t nrt = (b1 - b0 + TM - 1) / TM, nct = (a->n + 127) / 128;
        int64_t tiles = symmetric ? nrt * (nct + 1) / 2 : nrt * nct;
        int cn = a->ctx->opt_tile_cn ? a->ctx->opt_tile_cn : (tiles < 8 * (int64_t)a->ctx->sm_count ? 4 : 8);
        if (cn == 4) return launch_tiles_t<KM, 4, WC>(a, b0, b1, symmetric, out, ld, counts);
        return launch_tiles_t<KM, 8, WC>(a, b0, b1, symmetric, out, ld, counts);
    }
}

int launch_tiles(tdna_aln *a, int64_t b0, int64_t b1, bool symmetric, double *out, int64_t ld, tdna_counts *counts) {
    switch (a->kmode) {
        case KM_P_AMBIG: return counts ? launch_tiles_km<KM_P_AMBIG, true>(a, b0, b1, symmetric, out, ld, counts) : launch_tiles_km<KM_P_AMBIG, false>(a, b0, b1, symmetric, out, ld, counts);
        case KM_P_NOAMBIG: return counts ? launch_tiles_km<KM_P_NOAMBIG, true>(a, b0, b1, symmetric, out, ld, counts) : launch_tiles_km<KM_P_NOAMBIG, false>(a, b0, b1, symmetric, out, ld, counts);
        case KM_K2P: return counts ? launch_tiles_km<KM_K2P, true>(a, b0, b1, symmetric, out, ld, counts) : launch_tiles_km<KM_K2P, false>(a, b0, b1, symmetric, out, ld, counts);
        default: return counts ? launch_tiles_km<KM_TRANS, true>(a, b0, b1, symmetric, out, ld, counts) : launch_tiles_km<KM_TRANS, false>(a, b0, b1, symmetric, out, ld, counts);
    }
}

// ---- streaming Best Match: the tile pass --------------------------------------------------------
// Triangle tiles in two launches: the diagonal blocks first (list neighbours -- conspecifics and
// congeners when the list is sorted by name -- so the per-row thresholds are tight before the bulk),
// then everything right of them, row-tile major.  Kernel (a): part k of m takes the k-th contiguous 1/m of the tile list; kernel (b): units of
// 64 tile pairs dealt cyclically (launch_tc_cl).
template <int CN> int stream_prefixes(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    constexpr int TN = 16 * CN, R = TM / TN;
    if (a->sprefix_cn == CN && a->d_sprefix[0]) return TDNA_OK;
    int nrt = (int)((a->n + TM - 1) / TM);
    int64_t nct = (a->n + TN - 1) / TN;
    std::vector<int64_t> pre[2];
    for (int ph = 0; ph < 2; ph++) {
        pre[ph].resize(nrt + 1);
        pre[ph][0] = 0;
        for (int t = 0; t < nrt; t++) {
            int64_t first = (int64_t)t * R;
            int64_t diag = nct - first < R ? nct - first : R;
            if (diag < 0) diag = 0;
            int64_t rest = nct - first - R;
            if (rest < 0) rest = 0;
            pre[ph][t + 1] = pre[ph][t] + (ph == 0 ? diag : diag + rest);  // seed pass: diagonal blocks; bulk pass: the whole triangle
        }
        if (a->d_sprefix[ph]) CU(DFREE(a->d_sprefix[ph]));
        a->d_sprefix[ph] = nullptr;
        CU(DMALLOC(&a->d_sprefix[ph], (size_t)(nrt + 1) * 8));
        CU(cudaMemcpyAsync(a->d_sprefix[ph], pre[ph].data(), (size_t)(nrt + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        a->sprefix_tiles[ph] = pre[ph][nrt];
    }
    CU(cudaStreamSynchronize(c->stream));
    a->sprefix_cn = CN;
    return TDNA_OK;
}

template <int KM, int CN> int launch_stream_t(tdna_aln *a, int part, int nparts, int phases) {
    using Cfg = TileCfg<KM, CN>;
    tdna_ctx *c = a->ctx;
    TRY(stream_prefixes<CN>(a));
    if (phases & 1) CU(cudaEventRecord(c->ev_k0, c->stream));
    for (int ph = 0; ph < 2; ph++) {
        if (!(phases & (1 << ph))) continue;
        const StreamParams sp = pass_params(a, ph == 0);
        TileParams p = {};
        p.planes = a->d_planes;
        p.W = a->W;
        p.n = a->n;
        p.min_overlap = a->min_overlap;
        p.row_begin = 0;
        p.row_end = a->n;
        p.symmetric = 0;
        p.sched = SCHED_STREAM;
        p.first_col_extra = 0;
        p.nct = (int)((a->n + Cfg::TN - 1) / Cfg::TN);
        p.nrt = (int)((a->n + TM - 1) / TM);
        p.tile_prefix = a->d_sprefix[ph];
        // part k of m: a CONTIGUOUS 1/m of the tile list (all tiles cost the same; neighbours in the list share panels in L2)
        int64_t total = a->sprefix_tiles[ph];
        int64_t t0 = total * part / nparts, t1 = total * (part + 1) / nparts;
        p.t_stride = 1;
        if (ph == 1) CU(cudaEventRecord(c->ev_b0, c->stream));
        // segments of ~30 ms (a 128 x 64 tile of 658 columns takes ~0.1 us of the whole GPU): see segment_boundary
        const int64_t seg = c->opt_segment_items > 0 ? c->opt_segment_items : std::max<int64_t>(1 << 12, ((int64_t)1 << 18) * 21 / std::max(a->W, 1));
        for (int64_t s0 = t0; s0 < t1; s0 += seg) {
            p.t_offset = s0;
            p.num_tiles = std::min(seg, t1 - s0);
            if (ph == 0 || a->sp.want == TDNA_WANT_BEST_MATCH) TRY((launch_kernel<KM, CN, EPI_STREAM>(c, p, sp)));
            else TRY((launch_kernel<KM, CN, EPI_STREAM_X>(c, p, sp)));
            if (s0 + seg < t1) TRY(segment_boundary(c, s0 + seg - t0, t1 - t0));
        }
        if (ph == 0) CU(cudaEventRecord(c->ev_s1, c->stream));
    }
    if (phases & 2) {
        CU(cudaEventRecord(c->ev_k1, c->stream));
        c->ev_valid = true;
        c->ev_split = true;
    }
    return TDNA_OK;
}

template <int KM> int launch_stream_km(tdna_aln *a, int part, int nparts, int phases) {
    if constexpr (ModeTraits<KM>::NACC == 2) {
        return launch_stream_t<KM, 4>(a, part, nparts, phases);
    } else {
        int cn = a->ctx->opt_tile_cn ? a->ctx->opt_tile_cn : 4;
        if (cn == 4) return launch_stream_t<KM, 4>(a, part, nparts, phases);
        return launch_stream_t<KM, 8>(a, part, nparts, phases);
    }
}

static int launch_stream_popc(tdna_aln *a, int part, int nparts, int phases) {
    switch (a->kmode) {
        case KM_P_AMBIG: return launch_stream_km<KM_P_AMBIG>(a, part, nparts, phases);
        case KM_P_NOAMBIG: return launch_stream_km<KM_P_NOAMBIG>(a, part, nparts, phases);
        case KM_K2P: return launch_stream_km<KM_K2P>(a, part, nparts, phases);
        default: return launch_stream_km<KM_TRANS>(a, part, nparts, phases);
    }
}

// kernel (b) after its bulk launch(es): the queued pairs -> minima -> final thresholds -> candidate records
static int tc_resolve_queue(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    int wshift = 0;
    while ((1u << wshift) < a->tc_W) wshift++;
    unsigned grid = (unsigned)c->sm_count * 8;
    tc::ResolveArgs ra = {wshift, a->tc_W - 1u, a->min_overlap, (a->kmode == KM_K2P || a->tc_has_amb) ? a->d_planes : nullptr, a->W,
                          a->kmode == KM_K2P ? 1 : 0, a->kmode == KM_TRANS ? 1 : 0, 0};
    if (a->sp.want & TDNA_WANT_BEST_MATCH) {
        ra.cached = ra.planes ? 1 : 0;  // resolve_min_kernel leaves the exact counts of the rows' entries in the queue
        tc::resolve_min_kernel<<<grid, 256, 0, c->stream>>>(a->sp, ra);
        stream_thr_kernel<<<(unsigned)((a->n + 255) / 256), 256, 0, c->stream>>>(a->sp, a->n);
        c->launches += 2;
    }
    tc::resolve_emit_kernel<<<grid, 256, 0, c->stream>>>(a->sp, ra);
    c->launches++;
    CU(cudaGetLastError());
    return TDNA_OK;
}

// phases: bit 0 = diagonal blocks, bit 1 = the rest of the triangle
int launch_stream(tdna_aln *a, int part, int nparts, int phases) {
    tdna_ctx *c = a->ctx;
    // AUTO: kernel (b) wherever it is eligible -- it is ~3x faster than kernel (a) there (profiles/)
    bool tensor = c->opt_count_kernel == TDNA_KERNEL_TENSOR || (c->opt_count_kernel == TDNA_KERNEL_AUTO && !a->avoid_tensor && ensure_tc(a) == 1);
    a->last_kernel = tensor ? TDNA_KERNEL_TENSOR : TDNA_KERNEL_POPC;
    if (tensor) {
        if (ensure_tc(a) != 1) return fail(TDNA_E_STATE, "count kernel (b) needs uncorrected p with ambiguity allowed, K2P or transversions only, and L < 4096");
        if (phases & 1) CU(cudaEventRecord(c->ev_k0, c->stream));
        for (int ph = 0; ph < 2; ph++)
            if (phases & (1 << ph)) {
                if (ph == 1) CU(cudaEventRecord(c->ev_b0, c->stream));
                const int want = a->sp.want, row_bits = want & (TDNA_WANT_BEST_MATCH | TDNA_WANT_EDGES), class_bits = want & (TDNA_WANT_INTRA | TDNA_WANT_INTER);
                bool split = false;
                if (ph == 1 && row_bits && class_bits) {  // worth it when the tiles that can hold class pairs are a small part of the triangle
                    TRY(ensure_class_reach(a));
                    if (a->class_reach >= 0 && a->class_reach < a->n / 2) {
                        TRY(ensure_class_prefix(a));
                        split = a->tcprefix_class_tiles * 8 <= a->tcprefix_tiles[1];
                    }
                }
                if (split) {
                    // per-row outputs and class outputs in two launches: the rows' launch runs the whole triangle with none of the class
                    // machinery switched on (a class request in the same launch cost every tile of 10^6 rows a third of its time), the
                    // classes' launch visits only the near-diagonal tiles that can hold a class pair (computed twice: a few per mille)
                    a->sp.want = row_bits;
                    int rc = launch_tc<false>(a, ph, part, nparts, nullptr);
                    a->sp.want = class_bits;
                    if (rc == TDNA_OK) rc = launch_tc<false>(a, ph, part, nparts, nullptr);
                    a->sp.want = want;
                    TRY(rc);
                } else {
                    TRY(launch_tc<false>(a, ph, part, nparts, nullptr));
                }
                if (ph == 0) CU(cudaEventRecord(c->ev_s1, c->stream));
            }
        if (phases & 2) {
            CU(cudaEventRecord(c->ev_k1, c->stream));
            c->ev_valid = true;
            c->ev_split = true;
            TRY(tc_resolve_queue(a));
        }
        return TDNA_OK;
    }
    return launch_stream_popc(a, part, nparts, phases);
}

// ---- count kernel (b): operands, launches ----------------------------------------------------------
// Experiment switches (TDNA_TC_CLUSTER / BAND / INNER / PIECE / HINTS / GRID / NOSKIP) exist only in -DTDNA_TC_EXPERIMENTS builds; the
// product reads no environment variable on its launch path.
static const char *exp_env(const char *name) {
#ifdef TDNA_TC_EXPERIMENTS
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}
static bool tc_pair_mode() {  // 1 = single CTAs, 2 = pairs sharing the b-panel by multicast (default), 3 = cta_group::2 MMAs
    const char *e = exp_env("TDNA_TC_CLUSTER");  // 2: measured 2.79 ms against 2.86 ms for cta_group::2 on the 50,000 x 658 pass
    return (e ? atoi(e) : 2) == 3;
}
static int tc_band() {  // row tiles per band of the bulk phase
    const char *e = exp_env("TDNA_TC_BAND");
    // 64 row tiles = 8,192 rows: the band's a-panels (27 MB at 658 columns) stay in L2 while the b-panels stream past once per band;
    // measured on 50,000 x 658: band 16 2.82 ms / 2.04 GB DRAM, 32 2.76 ms, 64 2.75 ms, 128 2.75 ms
    int v = e ? atoi(e) : 64;
    return v < 1 ? 1 : v;
}
#define TC_BAND tc_band()

// Integers for the `dims`-dimensional symbol code of tdna_tc.cuh: u = b*1 + a*e_x, v = d*1 + c*e_y with
// a*c = -(W-1), a*d + b*c + dims*b*d = W, every entry (b, a+b, d, c+d) within int8, smallest W > L.
bool find_code(int L, int dims, tc::EncodeValues &out, uint32_t &Wout) {
    for (int W = 256; W <= 4096; W *= 2)  // a power of two: the epilogue decodes with a shift and a mask
        for (int aa = -127; aa <= 127 && W > L; aa++) {
            if (aa == 0 || (W - 1) % (aa < 0 ? -aa : aa) != 0) continue;
            int cc = -(W - 1) / aa;
            if (cc < -127 || cc > 127) continue;
            for (int bb = -127; bb <= 127; bb++) {
                if (aa + bb < -127 || aa + bb > 127) continue;
                int den = aa + dims * bb, num = W - bb * cc;
                if (den == 0 || num % den != 0) continue;
                int dd = num / den;
                if (dd < -127 || dd > 127 || cc + dd < -127 || cc + dd > 127) continue;
                out.a = aa; out.b = bb; out.c = cc; out.d = dd;
                Wout = (uint32_t)W;
                return true;
            }
        }
    return false;
}

// The operand code of count kernel (b) for a mode and an alignment width: the K-dimensions per site, the radix W of the accumulator
// and the row- / column-side int8 vector of every symbol code (tdna_tc.cuh: EncodeValues).  Host arithmetic only.
static bool tc_code_for(int kmode, int L, bool dims5, tc::EncodeValues &v, uint32_t &W, int &dims_out) {
    if (kmode == KM_P_NOAMBIG || L >= 4096 || L < 1) return false;
    int dims = (kmode == KM_K2P || !dims5) ? 4 : 5;
    v = tc::EncodeValues();
    W = 0;
    if (kmode == KM_TRANS) {
        // transversions only: six dimensions with entries 0, 1 and r = sqrt(W) (tdna_tc.cuh: EncodeValues); every symbol is expressed
        // exactly, so there is no ambiguity-letter machinery and the plain instantiation always runs
        dims = 6;
        int r = 16;
        while (r * r <= L) r *= 2;  // W = 256, 1024, 4096: the smallest above L
        W = (uint32_t)(r * r);
        v.trans = 1;
        const int8_t rr = (int8_t)r;
        const int8_t ta[4][6] = {{1, 0, rr, 0, 0, 1}, {0, 1, 0, rr, 0, 1}, {0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 1}};
        const int8_t tb[4][6] = {{1, 0, 0, rr, 1, 0}, {0, 1, rr, 0, 1, 0}, {0, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 1, 0}};
        for (int code = 0; code < 4; code++)
            for (int d = 0; d < 6; d++) { v.ta[code][d] = ta[code][d]; v.tb[code][d] = tb[code][d]; }
        // shared == L for every pair of two rows needs every site of both in AGRCTY or '-' (any other letter is only "shared" with a gap)
        v.breaks_full[3] = v.breaks_full[4] = v.breaks_full[5] = 1;
    } else {
        if (!find_code(L, dims, v, W)) {
            dims = 5;
            if (!find_code(L, dims, v, W)) return false;
        }
        v.gap_is_symbol = (kmode == KM_K2P || dims == 4) ? 0 : 1;
        for (int code = 0; code < dims; code++)  // one-hot: row (b + a [dim == code]), column (d + c [dim == code])
            for (int d = 0; d < dims; d++) { v.ta[code][d] = (int8_t)(v.b + (d == code ? v.a : 0)); v.tb[code][d] = (int8_t)(v.d + (d == code ? v.c : 0)); }
        for (int code = dims; code < 6; code++) v.breaks_full[code] = 1;  // a site without contribution ('_', '?', past the row's end; '-' when it is not a symbol)
    }
    dims_out = dims;
    return true;
}

static void tc_release(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    for (void **p : {(void **)&a->d_tcA, (void **)&a->d_tcB, (void **)&a->d_tcfull, (void **)&a->d_tcamb, (void **)&a->d_tcpanel_amb, (void **)&a->d_tcprefix[1],
                     (void **)&a->d_tcflag, (void **)&a->d_tcprefix4})
        if (*p) { DFREE(*p); *p = nullptr; }
    a->tc_state = 0;
}

// Everything about the operands that does not depend on the symbols: eligibility, the integer code, the buffers, the tile-slot prefix,
// the tensor maps.  1 = ready to be encoded into, -1 = this alignment / configuration is not eligible.
static int tc_prepare(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    if (a->kmode == KM_P_NOAMBIG || a->L >= 4096) return -1;
    // K2P ignores gap sites: four symbol dimensions -- where a four-dimension code exists (W = 1024 only, i.e. L < 1024); longer K2P
    // alignments use the five-dimension code with the gap dimension left empty (25 % more K-bytes per site).
    // Uncorrected p: barcode alignments rarely have INTERNAL gaps (external ones are '_' already), so the operands start with four
    // dimensions as well -- 20 % fewer K-bytes, MMAs and operand traffic -- and an internal '-' is left out like an ambiguity letter
    // (counted per row; the pairs it touches are bounded by the accumulator and settled from the bit-planes).  A list with more than
    // a sprinkling of internal gaps (tc_dims5: decided from the count the encoder returns) gets the five-dimension code, where '-'
    // is a symbol of its own and every accumulator is exact again.
    int dims = 0;
    tc::EncodeValues v = {};
    uint32_t W = 0;
    if (!tc_code_for(a->kmode, a->L, a->tc_dims5, v, W, dims)) return -1;
    const int nchunks = (a->L * dims + tc::KB - 1) / tc::KB;
    const int64_t npa = round_up(a->n, tc::M), npb = round_up(a->n, tc::N);
    const size_t npan = (size_t)(npb / tc::M);
    bool ok = cudaSuccess == DMALLOC(&a->d_tcA, (size_t)npa * nchunks * tc::KB) && cudaSuccess == DMALLOC(&a->d_tcB, (size_t)npb * nchunks * tc::KB) &&
              cudaSuccess == DMALLOC(&a->d_tcfull, npan) && cudaSuccess == DMALLOC(&a->d_tcamb, (size_t)a->n * 4) &&
              cudaSuccess == DMALLOC(&a->d_tcpanel_amb, npan) && cudaSuccess == DMALLOC(&a->d_tcflag, 16);
    // bulk phase: bands of TC_BAND row tiles; only real tiles are numbered (tdna_tc.cuh: slot_coords): a staircase over the band's
    // first TC_BAND / 2 columns, then TC_BAND slots per column
    const int nrt = (int)(npa / tc::M);
    const int64_t nct = npb / tc::N;
    const int nbands = (nrt + TC_BAND - 1) / TC_BAND;
    std::vector<int64_t> pre(nbands + 1);
    pre[0] = 0;
    for (int b = 0; b < nbands; b++) {
        int64_t tjmin = ((int64_t)b * TC_BAND >> 1);
        int64_t cols = nct - tjmin > 0 ? nct - tjmin : 0;
        int64_t tri = std::min<int64_t>(cols, TC_BAND >> 1);
        pre[b + 1] = pre[b] + tri * (tri + 1) + (cols - tri) * TC_BAND;
    }
    ok = ok && cudaSuccess == DMALLOC(&a->d_tcprefix[1], (size_t)(nbands + 1) * 8);
    // CL = 4 (tdna_tc.cuh: my_tile<4>): per band, super columns of two column tiles from the band's diagonal on, TC_BAND / 2 row pairs each
    std::vector<int64_t> pre4(nbands + 1);
    pre4[0] = 0;
    for (int b = 0; b < nbands; b++) {
        const int64_t scmin = ((int64_t)b * TC_BAND) >> 2, nsc = (nct + 1) >> 1;
        pre4[b + 1] = pre4[b] + std::max<int64_t>(nsc - scmin, 0) * (TC_BAND >> 1);
    }
    ok = ok && cudaSuccess == DMALLOC(&a->d_tcprefix4, (size_t)(nbands + 1) * 8);
    if (ok) cudaMemcpyAsync(a->d_tcprefix4, pre4.data(), (size_t)(nbands + 1) * 8, cudaMemcpyHostToDevice, c->stream);
    a->tcprefix4_total = pre4[nbands];
    if (ok) {
        cudaMemsetAsync(a->d_tcfull, 0, npan, c->stream);
        cudaMemsetAsync(a->d_tcpanel_amb, 0, npan, c->stream);
        cudaMemsetAsync(a->d_tcflag, 0, 16, c->stream);
        cudaMemcpyAsync(a->d_tcprefix[1], pre.data(), (size_t)(nbands + 1) * 8, cudaMemcpyHostToDevice, c->stream);
        ok = cudaStreamSynchronize(c->stream) == cudaSuccess;  // pre is a stack-lifetime host buffer
    }
    if (!ok) {
        cudaGetLastError();
        tc_release(a);
        return -1;
    }
    a->tcprefix_tiles[0] = nrt;
    a->tcprefix_tiles[1] = pre[nbands];
    a->tc_nchunks = nchunks;
    a->tc_W = W;
    a->tc_kmode = a->kmode;
    a->tc_dims = dims;
    a->tc_values = v;
    a->tc_pair = tc_pair_mode() || c->opt_cluster == 3;
    {   // tensor maps for the tiled TMA loads (driver entry point through the runtime: no -lcuda)
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult qr;
        a->tc_maps = false;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) == cudaSuccess && fp && qr == cudaDriverEntryPointSuccess) {
            encode_fn enc = reinterpret_cast<encode_fn>(fp);
            int inner = 256;  // uint32 elements per view row
            if (const char *e = exp_env("TDNA_TC_INNER")) { int x = atoi(e); if (x == 32 || x == 64 || x == 128 || x == 256) inner = x; }
            a->tc_inner = inner;
            cuuint64_t rowb = (cuuint64_t)inner * 4;
            cuuint64_t dimsA[2] = {(cuuint64_t)inner, (cuuint64_t)((size_t)npa * nchunks * tc::KB / rowb)}, dimsB[2] = {(cuuint64_t)inner, (cuuint64_t)((size_t)npb * nchunks * tc::KB / rowb)};
            cuuint64_t strides[1] = {rowb};
            cuuint32_t boxA[2] = {(cuuint32_t)inner, (cuuint32_t)(tc::A_STAGE / rowb)}, boxB[2] = {(cuuint32_t)inner, (cuuint32_t)(tc::B_STAGE / rowb)}, es[2] = {1, 1};
            CUresult r1 = enc(&a->tc_mapA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, a->d_tcA, dimsA, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r2 = enc(&a->tc_mapB, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, a->d_tcB, dimsB, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r3 = enc(&a->tc_mapBh, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, a->d_tcB, dimsB, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint32_t boxAh[2] = {(cuuint32_t)inner, (cuuint32_t)(tc::A_STAGE / rowb / 2)};
            CUresult r4 = enc(&a->tc_mapAh, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, a->d_tcA, dimsA, strides, boxAh, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            a->tc_maps = (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS && r3 == CUDA_SUCCESS && r4 == CUDA_SUCCESS);
        }
        cudaGetLastError();
    }
    return 1;
}

// Encodes the 256-row panels [panel0, panel0 + panels) on `stream` (the whole alignment, or one chunk of a chunked upload): both
// operand images, the per-panel "full" flags, the per-row counts of ambiguity letters.  d_tcflag[0]: a symbol outside the alphabet,
// d_tcflag[1]: some ambiguity letter.
static void tc_encode_range(tdna_aln *a, cudaStream_t stream, int64_t panel0, int64_t panels) {
    tdna_ctx *c = a->ctx;
    if (panels <= 0) return;
    const int nchunks = a->tc_nchunks, dims = a->tc_dims;
    const int64_t npa = round_up(a->n, tc::M), npb = round_up(a->n, tc::N);
    const int64_t row0 = panel0 * tc::N, rows = std::min<int64_t>(a->n, (panel0 + panels) * tc::N) - row0;
    if (rows > 0 && a->kmode == KM_TRANS) cudaMemsetAsync(a->d_tcamb + row0, 0, (size_t)rows * 4, stream);  // every symbol is expressed exactly
    else if (rows > 0)
        tc::amb_count_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(a->d_sym, a->d_len, a->n, a->sym_stride, a->L, a->kmode == KM_K2P ? 1 : 0, a->d_tcamb,
                                                                            a->d_tcpanel_amb, a->d_tcflag + 1, row0,
                                                                            (a->kmode != KM_K2P && a->tc_dims == 4) ? 1 : 0);
    if (a->tc_pair) {  // cta_group::2 (experiments): the b-side in 128-row panels as well; whole alignment only
        tc::encode_kernel<false, tc::M><<<dim3((unsigned)nchunks, (unsigned)(npa / tc::M)), 256, 0, stream>>>(a->d_sym, a->d_len, a->n, a->sym_stride, a->L, nchunks, dims, a->tc_values, a->d_tcA, a->d_tcflag, a->d_tcfull);
        tc::encode_kernel<true, tc::M><<<dim3((unsigned)nchunks, (unsigned)(npb / tc::M)), 256, 0, stream>>>(a->d_sym, a->d_len, a->n, a->sym_stride, a->L, nchunks, dims, a->tc_values, a->d_tcB, a->d_tcflag, nullptr);
        c->launches++;
    } else {
        tc::encode_both_kernel<<<dim3((unsigned)nchunks, (unsigned)panels), 256, 0, stream>>>(a->d_sym, a->d_len, a->n, a->sym_stride, a->L, nchunks, dims, a->tc_values, a->d_tcA,
                                                                                            a->d_tcB, npa / tc::M, a->d_tcflag, a->d_tcfull, panel0);
    }
    const int64_t k0 = 2 * panel0, k1 = std::min<int64_t>(npb / tc::M, 2 * (panel0 + panels));
    tc::flip_flags_kernel<<<(unsigned)((k1 - k0 + 255) / 256), 256, 0, stream>>>(a->d_tcfull, k0, k1, (int64_t)((a->n + tc::M - 1) / tc::M));  // gappy -> full
    c->launches += 3;
}

// uncorrected p on the four-dimension code: do the internal gaps it leaves out outnumber a sprinkling (1 site in 2,000)?
static bool tc_too_gappy(const tdna_aln *a, int gap_sites) {
    return a->kmode == KM_P_AMBIG && a->tc_dims == 4 && !a->tc_dims5 && (int64_t)gap_sites * 2000 > (int64_t)a->n * a->L;
}

// 1 = ready, -1 = this alignment / configuration is not eligible
int ensure_tc(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    if (a->kmode == KM_P_NOAMBIG || a->L >= 4096) return -1;
    if (a->tc_state == 1 && a->tc_kmode != a->kmode) tc_release(a);  // the operands encode the mode's valid-symbol set: rebuild
    if (a->tc_state != 0) return a->tc_state;
    if (tc_prepare(a) != 1) { a->tc_state = -1; return -1; }
    tc_encode_range(a, c->stream, 0, round_up(a->n, tc::N) / tc::N);
    int flags3[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(flags3, a->d_tcflag, 12, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) flags3[0] = 1;
    if (flags3[0]) {  // a symbol outside the alphabet (the pack kernel rejects those first) or a launch failure: the popc kernel handles it
        tc_release(a);
        a->tc_state = -1;
        return -1;
    }
    if (tc_too_gappy(a, flags3[2])) {  // once per alignment: the same operands again with '-' as a fifth symbol
        tc_release(a);
        a->tc_dims5 = true;
        return ensure_tc(a);
    }
    a->tc_has_amb = flags3[1] != 0;
    a->tc_state = 1;
    if (a->kmode == KM_P_AMBIG) c->hint_dims5 = (int64_t)flags3[2] * 2000 > (int64_t)a->n * a->L ? 1 : 0;  // the next list's first code
    return 1;
}

// How far from the diagonal a class pair can lie: for every row, the last row of its species / of its genus (one O(n) walk over the
// labels, once per label set).  A list sorted by name keeps this at the size of its largest genus.
static int ensure_class_reach(tdna_aln *a) {
    if (a->class_reach >= 0 || (int64_t)a->h_species.size() != a->n || (int64_t)a->h_genus.size() != a->n) return TDNA_OK;
    std::unordered_map<int32_t, int64_t> last_s, last_g;
    for (int64_t i = 0; i < a->n; i++) { if (a->h_species[i] >= 0) last_s[a->h_species[i]] = i; last_g[a->h_genus[i]] = i; }
    int64_t reach = 0;
    for (int64_t i = 0; i < a->n; i++) {
        if (a->h_species[i] >= 0) reach = std::max(reach, last_s[a->h_species[i]] - i);
        reach = std::max(reach, last_g[a->h_genus[i]] - i);
    }
    a->class_reach = reach;
    return TDNA_OK;
}
// Class pairs stay within class_reach rows of the diagonal: a class-only launch numbers only those columns of each band (the full
// numbering made every role of every CTA walk -- and skip -- 1.5 * 10^7 slots on 10^6 rows).  d_tcprefix[0], tcprefix_class_tiles.
static int ensure_class_prefix(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    if (a->tcprefix_reach == a->class_reach && a->d_tcprefix[0]) return TDNA_OK;
    const int nrt = (int)(round_up(a->n, tc::M) / tc::M), nct = (int)(round_up(a->n, tc::N) / tc::N), nbands = (nrt + TC_BAND - 1) / TC_BAND;
    std::vector<int64_t> pre(nbands + 1);
    pre[0] = 0;
    for (int b = 0; b < nbands; b++) {
        const int64_t ti0 = (int64_t)b * TC_BAND, tjmin = ti0 >> 1;
        const int64_t lim = std::min<int64_t>(nct, ((ti0 + TC_BAND) * tc::M - 1 + a->class_reach) / tc::N + 1);
        const int64_t cols = std::max<int64_t>(lim - tjmin, 0), tri = std::min<int64_t>(cols, TC_BAND >> 1);
        pre[b + 1] = pre[b] + tri * (tri + 1) + (cols - tri) * TC_BAND;
    }
    if (!a->d_tcprefix[0]) CU(DMALLOC(&a->d_tcprefix[0], (size_t)(nbands + 1) * 8));
    CU(cudaMemcpyAsync(a->d_tcprefix[0], pre.data(), (size_t)(nbands + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    a->tcprefix_reach = a->class_reach;
    a->tcprefix_class_tiles = pre[nbands];
    return TDNA_OK;
}

// Seed pass of kernel (b): which 256-column blocks each pair of row tiles (= one 256-row block) visits.  Always its own diagonal
// block; plus, from the labels, the block holding the nearest row of a species whose id has the parity the diagonal block lacks
// (thresholds are max(min conspecific, min even-id species, min odd-id species): a block inside one large species would leave
// one of them unbounded and the bulk pass would queue every pair of its rows), and the neighbouring block when a row at the
// block's edge has its only conspecific neighbours across the edge.
static int build_seed_list(tdna_aln *a) {
    if (a->seedlist_valid) return TDNA_OK;
    tdna_ctx *c = a->ctx;
    const int64_t n = a->n;
    const std::vector<int32_t> &sp = a->h_species;
    if ((int64_t)sp.size() != n) return fail(TDNA_E_STATE, "labels are not set");
    const int64_t nblk = (n + tc::N - 1) / tc::N;
    // per block and id parity: does the block hold a named row of that parity; the nearest such row before / after the block
    // (two passes over the labels, O(blocks) memory: this runs once per alignment and is part of the host-buffer path)
    std::vector<int32_t> before[2], after[2];
    std::vector<uint8_t> has[2];
    for (int m = 0; m < 2; m++) { before[m].assign((size_t)nblk, -1); after[m].assign((size_t)nblk, -1); has[m].assign((size_t)nblk, 0); }
    {
        int32_t last[2] = {-1, -1};
        for (int64_t u = 0; u < nblk; u++) {
            before[0][u] = last[0];
            before[1][u] = last[1];
            const int64_t c0 = u * tc::N, c1 = std::min<int64_t>(n, c0 + tc::N);
            for (int64_t i = c0; i < c1; i++)
                if (sp[i] >= 0) { last[sp[i] & 1] = (int32_t)i; has[sp[i] & 1][u] = 1; }
        }
        int32_t nxt[2] = {-1, -1};
        for (int64_t u = nblk - 1; u >= 0; u--) {
            after[0][u] = nxt[0];
            after[1][u] = nxt[1];
            const int64_t c0 = u * tc::N, c1 = std::min<int64_t>(n, c0 + tc::N);
            for (int64_t i = c1 - 1; i >= c0; i--)
                if (sp[i] >= 0) nxt[sp[i] & 1] = (int32_t)i;
        }
    }
    std::vector<int2> list;
    list.reserve((size_t)nblk * 2);
    for (int64_t u = 0; u < nblk; u++) {
        const int64_t c0 = u * tc::N, c1 = std::min<int64_t>(n, c0 + tc::N);
        int64_t extra[4];
        int ne = 0;
        auto add = [&](int64_t row) {
            if (row < 0 || row >= n) return;
            const int64_t b = row / tc::N;
            if (b == u) return;
            for (int k = 0; k < ne; k++)
                if (extra[k] == b) return;
            if (ne < 4) extra[ne++] = b;
        };
        for (int m = 0; m < 2; m++) {
            if (has[m][u]) continue;  // the block has a species of this parity
            const int64_t l = before[m][u], r = after[m][u];
            if (l >= 0 && (r < 0 || c0 - l <= r - (c1 - 1))) add(l); else if (r >= 0) add(r);
        }
        // rows at the block's edges whose conspecific neighbours are all across the edge (lists sorted by name: adjacent rows)
        if (c0 > 0 && sp[c0] >= 0 && sp[c0 - 1] == sp[c0] && !(c0 + 1 < c1 && sp[c0 + 1] == sp[c0])) add(c0 - 1);
        if (c1 < n && sp[c1 - 1] >= 0 && sp[c1] == sp[c1 - 1] && !(c1 - 2 >= c0 && sp[c1 - 2] == sp[c1 - 1])) add(c1);
        list.push_back(make_int2((int)(2 * u), (int)u));
        for (int k = 0; k < ne; k++) list.push_back(make_int2((int)(2 * u), (int)extra[k]));
    }
    if (a->d_seedlist) CU(DFREE(a->d_seedlist));
    a->d_seedlist = nullptr;
    CU(DMALLOC(&a->d_seedlist, list.size() * sizeof(int2)));
    CU(cudaMemcpyAsync(a->d_seedlist, list.data(), list.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    a->seed_pairs = (int)list.size();
    a->h_seedlist = list;
    a->seedlist_valid = true;
    return TDNA_OK;
}

template <bool COUNTS_MODE, int CL> int launch_tc_cl(tdna_aln *a, int phase, int part, int nparts, tdna_counts *counts) {
    tdna_ctx *c = a->ctx;
    // the plain instantiation serves uncorrected p on alignments without ambiguity letters (exact accumulators everywhere)
    const bool has_amb = a->pl_general >= 0 ? a->pl_general != 0 : a->tc_has_amb;
    const bool general = !COUNTS_MODE && (a->kmode == KM_K2P || has_amb);
    auto kern = tc::tile_kernel<COUNTS_MODE, CL, false>;
    if constexpr (!COUNTS_MODE) {
        if (general) kern = tc::tile_kernel<COUNTS_MODE, CL, true>;
    }
    static bool attr_set[64] = {};  // per device (see launch_kernel)
    if (!attr_set[c->device & 63]) {
        CU(cudaFuncSetAttribute(tc::tile_kernel<COUNTS_MODE, CL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM));
        if constexpr (!COUNTS_MODE) CU(cudaFuncSetAttribute(tc::tile_kernel<COUNTS_MODE, CL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM));
        attr_set[c->device & 63] = true;
    }
    tc::Params p = {};
    p.A = a->d_tcA;
    p.B = a->d_tcB;
    p.nchunks = a->tc_nchunks;
    p.n = (int)a->n;
    p.min_overlap = a->min_overlap;
    p.W = a->tc_W;
    p.wshift = 0;
    while ((1u << p.wshift) < a->tc_W) p.wshift++;
    p.nrt = (int)(round_up(a->n, tc::M) / tc::M);
    p.nct = (int)(round_up(a->n, tc::N) / tc::N);
    p.band = phase ? TC_BAND : 0;
    p.nbands = (p.nrt + TC_BAND - 1) / TC_BAND;
    p.prefix = a->d_tcprefix[1];
    if (!COUNTS_MODE && phase == 0 && a->pl_seed_pairs >= 0) {  // the pipelined pass: the items of the chunk that has just arrived
        p.seed_list = a->pl_seed_list;
        p.seed_pairs = a->pl_seed_pairs;
    } else if (!COUNTS_MODE && phase == 0 && a->has_labels) {
        TRY(build_seed_list(a));
        p.seed_list = a->d_seedlist;
        p.seed_pairs = a->seed_pairs;
    }
    p.counts = counts;
    p.extras = a->sp.want != TDNA_WANT_BEST_MATCH;
    p.k2p = a->kmode == KM_K2P ? 1 : 0;
    p.trans = a->kmode == KM_TRANS ? 1 : 0;
    p.panel_full = a->d_tcfull;
    p.L = a->L;
    p.amb = has_amb ? a->d_tcamb : nullptr;
    p.panel_amb = a->d_tcpanel_amb;
    if (!COUNTS_MODE && a->d_panel_range && (a->sp.want & ~(TDNA_WANT_INTRA | TDNA_WANT_INTER)) == 0 && !exp_env("TDNA_TC_NOSKIP")) {
        p.class_only = ((a->sp.want & TDNA_WANT_INTRA) ? 1 : 0) | ((a->sp.want & TDNA_WANT_INTER) ? 2 : 0);
        p.panel_range = a->d_panel_range;
    }
    p.range_rows = a->d_range_rows;
    p.range_cols = a->d_range_cols;
    p.class_reach = a->n;
    if (p.extras || p.class_only) {
        TRY(ensure_class_reach(a));
        if (a->class_reach >= 0) p.class_reach = a->class_reach;
    }
    p.col_reach = -1;
    if (phase && p.class_only && a->class_reach >= 0 && a->class_reach < a->n / 2) {
        TRY(ensure_class_prefix(a));
        p.col_reach = a->class_reach;
        p.prefix = a->d_tcprefix[0];
    }
    int64_t total = p.seed_list ? 2 * (int64_t)p.seed_pairs : (p.col_reach >= 0 ? a->tcprefix_class_tiles : a->tcprefix_tiles[phase]);   // slots
    if (CL == 4) {  // quads of 2 x 2 tiles, numbered per band (only the rows' bulk launch comes here: launch_tc)
        total = a->tcprefix4_total;
        p.prefix = a->d_tcprefix4;
    } else if (CL >= 2) total = (total + 1) / 2;  // pairs of slots
    // part k of m takes every m-th UNIT of 128 consecutive slots (64 row tiles x 2 column tiles of one band: the panels a
    // unit shares stay in this GPU's L2); units are dealt cyclically so that work AND candidates (which cluster near the
    // diagonal) are balanced across the parts
    p.unit = CL == 4 ? 32 : (CL >= 2 ? 64 : 128);
    if (c->opt_unit_items > 0 && CL == 2) p.unit = c->opt_unit_items;
    if (phase == 0) p.unit = 1;  // the seed pass has a few hundred items in all: deal them one by one, or most parts get none
    int64_t units_total = (total + p.unit - 1) / p.unit;
    int64_t my_units = units_total > part ? (units_total - part + nparts - 1) / nparts : 0;
    p.t_offset = part;
    p.t_stride = nparts;
    p.num_tiles = my_units * p.unit;  // items past the end of the list are skipped by the kernel
    const bool region = !COUNTS_MODE && phase == 1 && CL == 2 && a->pl_reg_n > 0;
    if (region) {  // (one part only: tdna_analyze_host pipelines on a single GPU)
        p.reg_start = a->pl_reg_start;
        p.reg_prefix = a->pl_reg_prefix;
        p.reg_n = a->pl_reg_n;
        p.num_tiles = a->pl_reg_items;
    }
    p.planes = a->d_planes;
    p.W32 = a->W;
    {
        const char *e = exp_env("TDNA_TC_PIECE");
        int v = e ? atoi(e) : (a->tc_maps ? 0 : 16384);  // default: tiled TMA when the tensor maps exist
        p.piece = (v >= 1024 && 16384 % v == 0) ? v : (a->tc_maps ? 0 : 16384);
    }
    if (const char *e = exp_env("TDNA_TC_HINTS")) p.hints = atoi(e);
    p.a_rows = tc::A_STAGE / (a->tc_inner * 4);
    p.b_rows = tc::B_STAGE / (a->tc_inner * 4);
    int64_t units = CL == 4 ? c->sm_count / 4 : (CL >= 2 ? c->sm_count / 2 : c->sm_count);
    if (CL == 4) {  // clusters of four whole-SM CTAs must fit their GPCs: ask how many can be resident at once
        if (c->max_quads < 0) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3((unsigned)(c->sm_count / 4 * 4));
            q.blockDim = dim3(tc::THREADS);
            q.dynamicSmemBytes = tc::SMEM;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 4;
            qa[0].val.clusterDim.y = qa[0].val.clusterDim.z = 1;
            q.attrs = qa;
            q.numAttrs = 1;
            int nq = 0;
            if (cudaOccupancyMaxActiveClusters(&nq, kern, &q) != cudaSuccess || nq < 1) { cudaGetLastError(); nq = 0; }
            c->max_quads = nq;
        }
        if (c->max_quads < 1) return fail(TDNA_E_STATE, "clusters of four CTAs cannot be resident on this device");
        units = std::min<int64_t>(units, c->max_quads);
    }
    if (const char *e = exp_env("TDNA_TC_GRID")) { int v = atoi(e); if (v >= 1 && v < units) units = v; }
    const StreamParams sp = pass_params(a, !COUNTS_MODE && phase == 0);
    // segments of ~30 ms (an item of two 128 x 256 tiles of 658 columns takes ~0.13 us of the whole GPU): see segment_boundary
    const int64_t items = p.num_tiles;
    const int64_t seg = (COUNTS_MODE || phase == 0 || region) ? items : c->opt_segment_items > 0 ? c->opt_segment_items : std::max<int64_t>(units * 64, ((int64_t)1 << 18) * 26 / std::max(p.nchunks, 1) * (CL >= 2 ? 1 : 2) / (CL == 4 ? 2 : 1));
    for (int64_t k0 = 0; k0 < items; k0 += seg) {
        p.k_base = k0;
        p.num_tiles = std::min(seg, items - k0);
        int64_t grid = (p.num_tiles < units ? p.num_tiles : units) * (CL == 4 ? 4 : (CL >= 2 ? 2 : 1));
        if (grid < 1) break;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(tc::THREADS);
        cfg.dynamicSmemBytes = tc::SMEM;
        cfg.stream = c->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL == 4 ? 4 : (CL >= 2 ? 2 : 1);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, kern, a->tc_mapA, CL == 4 ? a->tc_mapAh : a->tc_mapB, a->tc_mapBh, p, sp));
        c->launches++;
        CU(cudaGetLastError());
        if (k0 + seg < items) TRY(segment_boundary(c, k0 + seg, items));
    }
    return TDNA_OK;
}

// pairs of CTAs sharing the b-panel by TMA multicast need the tensor maps; TDNA_TC_CLUSTER=1 switches them off (experiments)
template <bool COUNTS_MODE> int launch_tc(tdna_aln *a, int phase, int part, int nparts, tdna_counts *counts) {
    const char *e = exp_env("TDNA_TC_CLUSTER");
    int cl = e ? atoi(e) : 2;
    if (a->tc_pair) cl = 3;          // the b-side operands were encoded for cta_group::2
    else if (cl == 3) cl = 2;
    if (!a->tc_maps || (TC_BAND & 1)) { if (a->tc_pair) return fail(TDNA_E_STATE, "cta_group::2 needs the tensor maps"); cl = 1; }
    if (cl == 3) return launch_tc_cl<COUNTS_MODE, 3>(a, phase, part, nparts, counts);
    if constexpr (!COUNTS_MODE) {
        // clusters of four (TDNA_OPT_CLUSTER = 4): the bulk launch that produces per-row output; seed, class-only and region launches
        // keep the pairs
        const bool rows_launch = phase == 1 && (a->sp.want & (TDNA_WANT_BEST_MATCH | TDNA_WANT_EDGES)) != 0 && a->pl_reg_n == 0;
        if (cl == 2 && a->ctx->opt_cluster == 4 && rows_launch && (TC_BAND % 4) == 0) return launch_tc_cl<COUNTS_MODE, 4>(a, phase, part, nparts, counts);
    }
    if (cl == 2) return launch_tc_cl<COUNTS_MODE, 2>(a, phase, part, nparts, counts);
    return launch_tc_cl<COUNTS_MODE, 1>(a, phase, part, nparts, counts);
}

int ensure_packed(tdna_aln *a) {
    if (!a) return fail(TDNA_E_ARG, "null alignment");
    if (!a->packed) TRY(pack(a));
    return TDNA_OK;
}

// rows per band so that band x ld doubles fit the budget (multiple of 128, at least 128)
int64_t band_rows_for(tdna_aln *a) {
    int64_t rows = a->row_end - a->row_begin;
    int64_t maxb = a->ctx->budget / (a->ld * 8);
    maxb = maxb / TM * TM;
    if (maxb < TM) maxb = TM;
    return rows < maxb ? rows : maxb;
}

int ensure_mat(tdna_aln *a, int64_t rows) {
    tdna_ctx *c = a->ctx;
    int64_t need = round_up(rows, TM) * a->ld;
    if (need > c->mat_cap) {
        if (c->d_mat) CU(cudaFree(c->d_mat));
        c->d_mat = nullptr;
        c->mat_cap = 0;
        c->mat_owner = nullptr;
        CU(cudaMalloc(&c->d_mat, (size_t)need * 8));
        c->mat_cap = need;
    }
    if (c->mat_owner != a) a->mat_valid = false;
    a->d_mat = c->d_mat;
    return TDNA_OK;
}

// Runs the count kernel band by band over the row range and hands each finished band (device
// resident) to `consume(b0, b1)`.  When the whole matrix is one band and the row range is
// everything, only the upper-triangle tiles are computed and mirrored.
template <class F> int for_each_band(tdna_aln *a, F consume) {
    TRY(ensure_packed(a));
    tdna_ctx *c = a->ctx;
    int64_t rows = a->row_end - a->row_begin;
    if (rows <= 0) return TDNA_OK;
    if (a->row_begin % TM != 0) return fail(TDNA_E_ARG, "row_begin must be a multiple of 128");
    int64_t band = band_rows_for(a);
    TRY(ensure_mat(a, band));
    bool whole = (band == rows);
    for (int64_t b0 = a->row_begin; b0 < a->row_end; b0 += band) {
        int64_t b1 = b0 + band < a->row_end ? b0 + band : a->row_end;
        if (c->cancel) { c->cancel = 0; return fail(TDNA_E_CANCELLED, "cancelled"); }
        if (!(whole && a->mat_valid)) {
            bool symmetric = whole && a->row_begin == 0 && a->row_end == a->n;
            c->mat_owner = a;
            TRY(launch_tiles(a, b0, b1, symmetric, a->d_mat, a->ld, nullptr));
            a->mat_valid = whole;
        }
        TRY(consume(b0, b1));
        if (c->progress) {
            CU(cudaStreamSynchronize(c->stream));
            c->progress(b1 - a->row_begin, rows, c->progress_user);
        }
    }
    return TDNA_OK;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice;
}

int ensure_scratch(tdna_aln *a, size_t bytes) {
    if (bytes > a->scratch_cap) {
        tdna_ctx *c = a->ctx;
        if (a->d_scratch) CU(DFREE(a->d_scratch));
        a->d_scratch = nullptr;
        a->scratch_cap = 0;
        CU(DMALLOC(&a->d_scratch, bytes));
        a->scratch_cap = bytes;
    }
    return TDNA_OK;
}

// the per-row result records: n, or every rank's padded slice once a multi-GPU pass gathers them here
int ensure_res(tdna_aln *a, int64_t rows) {
    tdna_ctx *c = a->ctx;
    if (rows < a->n) rows = a->n;
    if (a->d_res && a->res_rows >= rows) return TDNA_OK;
    if (a->d_res) CU(DFREE(a->d_res));
    a->d_res = nullptr;
    a->res_rows = 0;
    CU(DMALLOC(&a->d_res, (size_t)rows * sizeof(tdna_row_result)));
    a->res_rows = rows;
    return TDNA_OK;
}

int need_labels(tdna_aln *a) {
    if (!a->has_labels) return fail(TDNA_E_STATE, "tdna_alignment_set_labels has not been called");
    return TDNA_OK;
}

template <int KM> int launch_row_extremes(tdna_aln *a, tdna_extreme_result *d_out) {
    tdna_ctx *c = a->ctx;
    constexpr int SMEM = EXT_CAP * 12;
    static bool attr_set[64] = {};
    if (!attr_set[c->device & 63]) {
        CU(cudaFuncSetAttribute(row_extremes_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set[c->device & 63] = true;
    }
    row_extremes_kernel<KM><<<(unsigned)a->n, EXT_THREADS, SMEM, c->stream>>>(a->d_planes, a->W, a->n, a->min_overlap, a->d_species, a->d_genus, a->d_gslot,
                                                                               a->d_gstart, a->d_gmem, d_out);
    c->launches++;
    CU(cudaGetLastError());
    return TDNA_OK;
}

}  // namespace

extern "C" {

int32_t tdna_abi_version(void) { return TDNA_ABI_VERSION; }
const char *tdna_last_error(void) { return g_err.c_str(); }

int32_t tdna_init(int32_t device, tdna_ctx **out) {
    if (!out) return fail(TDNA_E_ARG, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(TDNA_E_CUDA, std::string("no CUDA device: this library has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= count) return fail(TDNA_E_ARG, "device out of range");
    CU(cudaSetDevice(device));
    tdna_ctx *c = new tdna_ctx();
    c->device = device;
    auto body = [&]() -> int {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CU(cudaMalloc(&c->d_counter, 64));
        CU(cudaEventCreate(&c->ev_k0));
        CU(cudaEventCreate(&c->ev_k1));
        CU(cudaEventCreate(&c->ev_s1));
        CU(cudaEventCreate(&c->ev_b0));
        for (cudaEvent_t &e : c->ev_ph) CU(cudaEventCreate(&e));
        CU(cudaMallocHost(&c->h_status, 64 * sizeof(int64_t)));
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : c->ev_copy) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&c->enc_stream, cudaStreamNonBlocking));
        for (cudaEvent_t &e : c->ev_enc) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        CU(cudaMemPoolCreate(&c->pool, &props));
        unsigned long long keep = ~0ull;
        CU(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
        uint16_t lut[256];
        build_lut(lut);
        CU(cudaMemcpyToSymbol(c_symlut, lut, sizeof(lut)));  // the same table for every context of the process: idempotent
        return TDNA_OK;
    };
    const int rc = body();
    if (rc != TDNA_OK) {  // nothing half-built survives a failed init
        const std::string why = tdna_last_error();
        tdna_shutdown(c);
        return fail(rc, why);
    }
    *out = c;
    return TDNA_OK;
}

int32_t tdna_shutdown(tdna_ctx *c) {
    if (!c) return TDNA_OK;
    cudaSetDevice(c->device);
    if (c->d_counter) cudaFree(c->d_counter);
    if (c->d_mat) cudaFree(c->d_mat);
    if (c->comm.comm && c->comm.api) c->comm.api->CommDestroy(c->comm.comm);
    c->comm.comm = nullptr;
    if (c->h_status) cudaFreeHost(c->h_status);
    for (cudaEvent_t e : c->ev_copy) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (cudaEvent_t e : c->ev_enc) if (e) cudaEventDestroy(e);
    if (c->enc_stream) cudaStreamDestroy(c->enc_stream);
    for (cudaEvent_t e : {c->ev_k0, c->ev_k1, c->ev_s1, c->ev_b0})
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_ph)
        if (e) cudaEventDestroy(e);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->pool) cudaMemPoolDestroy(c->pool);
    delete c;
    return TDNA_OK;
}

// ---- communicator ------------------------------------------------------------------------------------------------------------
int32_t tdna_comm_unique_id(uint8_t *id) {
    if (!id) return fail(TDNA_E_ARG, "null argument");
    std::string why;
    const NcclApi *nc = nccl_api(&why);
    if (!nc) return fail(TDNA_E_STATE, why);
    static_assert(sizeof(ncclUniqueId) == TDNA_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    ncclResult_t r = nc->GetUniqueId(&u);
    if (r != ncclSuccess) return fail(TDNA_E_CUDA, std::string("ncclGetUniqueId: ") + nc->GetErrorString(r));
    memcpy(id, &u, sizeof(u));
    return TDNA_OK;
}

int32_t tdna_comm_init(tdna_ctx *c, const uint8_t *id, int32_t rank, int32_t nranks) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(TDNA_E_ARG, "bad argument");
    if (nranks > ROUTE_MAX_PARTS) return fail(TDNA_E_ARG, "at most 64 ranks");
    if (c->comm.comm) return fail(TDNA_E_STATE, "the context already has a communicator");
    std::string why;
    const NcclApi *nc = nccl_api(&why);
    if (!nc) return fail(TDNA_E_STATE, why);
    CU(cudaSetDevice(c->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    ncclResult_t r = nc->CommInitRank(&comm, nranks, u, rank);
    if (r != ncclSuccess) return fail(TDNA_E_CUDA, std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
    c->comm.api = nc;
    c->comm.comm = comm;
    c->comm.rank = rank;
    c->comm.nranks = nranks;
    c->comm.enabled = true;
    return TDNA_OK;
}

int32_t tdna_comm_destroy(tdna_ctx *c) {
    if (!c) return fail(TDNA_E_ARG, "null ctx");
    if (c->comm.comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        c->comm.api->CommDestroy(c->comm.comm);
    }
    c->comm = Comm();
    return TDNA_OK;
}

int32_t tdna_comm_info(tdna_ctx *c, int32_t *rank, int32_t *nranks) {
    if (!c) return fail(TDNA_E_ARG, "null ctx");
    if (rank) *rank = c->comm.comm ? c->comm.rank : 0;
    if (nranks) *nranks = c->comm.comm ? c->comm.nranks : 1;
    return TDNA_OK;
}

int32_t tdna_init_multi(const int32_t *devices, int32_t ndev, tdna_ctx **out) {
    if (!devices || !out || ndev < 1 || ndev > ROUTE_MAX_PARTS) return fail(TDNA_E_ARG, "bad argument");
    for (int k = 0; k < ndev; k++) out[k] = nullptr;
    auto undo = [&]() { for (int k = 0; k < ndev; k++) { if (out[k]) tdna_shutdown(out[k]); out[k] = nullptr; } };
    for (int k = 0; k < ndev; k++) {
        int rc = tdna_init(devices[k], &out[k]);
        if (rc != TDNA_OK) { undo(); return rc; }
    }
    if (ndev == 1) return TDNA_OK;
    std::string why;
    const NcclApi *nc = nccl_api(&why);
    if (!nc) { undo(); return fail(TDNA_E_STATE, why); }
    ncclUniqueId u;
    ncclResult_t r = nc->GetUniqueId(&u);
    if (r == ncclSuccess) r = nc->GroupStart();
    std::vector<ncclComm_t> comms((size_t)ndev, nullptr);
    for (int k = 0; k < ndev && r == ncclSuccess; k++) {  // one thread, several devices: the inits must sit in one NCCL group
        if (cudaSetDevice(devices[k]) != cudaSuccess) { r = ncclUnhandledCudaError; break; }
        r = nc->CommInitRank(&comms[k], ndev, u, k);
    }
    ncclResult_t r2 = nc->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { undo(); return fail(TDNA_E_CUDA, std::string("NCCL communicator setup: ") + nc->GetErrorString(r)); }
    for (int k = 0; k < ndev; k++) {
        out[k]->comm.api = nc;
        out[k]->comm.comm = comms[k];
        out[k]->comm.rank = k;
        out[k]->comm.nranks = ndev;
        out[k]->comm.enabled = true;
    }
    return TDNA_OK;
}

void *tdna_stream(tdna_ctx *c) { return c ? (void *)c->stream : nullptr; }
int32_t tdna_set_memory_budget(tdna_ctx *c, int64_t bytes) {
    if (!c || bytes < (1 << 20)) return fail(TDNA_E_ARG, "bad budget");
    c->budget = bytes;
    return TDNA_OK;
}
int64_t tdna_launch_count(tdna_ctx *c) { return c ? c->launches : 0; }

int32_t tdna_set_progress(tdna_ctx *c, tdna_progress_fn fn, void *user) {
    if (!c) return fail(TDNA_E_ARG, "null ctx");
    c->progress = fn;
    c->progress_user = user;
    return TDNA_OK;
}
int32_t tdna_cancel(tdna_ctx *c) {
    if (!c) return fail(TDNA_E_ARG, "null ctx");
    c->cancel = 1;
    return TDNA_OK;
}

// Device buffers of a new alignment, the symbols, the bit-planes of the default configuration.  Any failure leaves `a` to the
// caller, who destroys it (nothing leaks: destroy frees whatever was allocated so far).
//  - several GPUs, host symbols (the same list on every rank -- one JVM: the same array): each rank uploads 1/G of the rows over
//    its own PCIe link and the slices are all-gathered over NVLink, instead of G full uploads through the host at once;
//  - one GPU, host symbols, 8,192+ rows: the upload runs in row chunks on a copy stream while the context's stream packs each
//    chunk that has arrived into bit-planes AND encodes count kernel (b)'s operand panels for the default configuration (the GPU
//    would idle through the PCIe time otherwise): both are ready when the last chunk lands.
// hook (optional): called after each row chunk of a chunked upload has been packed and encoded (enqueued), with the chunk's index
// and rows; chunk_rows > 0 fixes the chunk size (the pipelined pass wants whole bands of row tiles)
using ChunkHook = std::function<int(int, int64_t, int64_t)>;
static int aln_build(tdna_ctx *c, tdna_aln *a, const uint8_t *symbols, const int32_t *lengths, const ChunkHook *hook = nullptr, int64_t chunk_rows = 0,
                     const std::function<int(tdna_aln *)> *pre = nullptr) {
    const int64_t n = a->n, row_stride = a->sym_stride;
    const int32_t L = a->L;
    const size_t sym_bytes = (size_t)(n - 1) * row_stride + L;
    const int G = c->comm.active() ? c->comm.nranks : 1;
    const bool host_sym = !is_device_ptr(symbols);
    const bool chunked = G == 1 && host_sym && n >= 8192 && (chunk_rows == 0 || (n + chunk_rows - 1) / chunk_rows <= 15);
    if (G > 1 && n >= 1024 * (int64_t)G && host_sym) {
        const NcclApi *nc = c->comm.api;
        const int64_t rs = (n + G - 1) / G;
        const size_t slice = (size_t)rs * row_stride;
        CU(DMALLOC(&a->d_sym, slice * G));
        const int64_t r0 = std::min<int64_t>(n, (int64_t)c->comm.rank * rs);
        const size_t off = (size_t)r0 * row_stride;
        const size_t mine = off < sym_bytes ? std::min(slice, sym_bytes - off) : 0;
        if (mine) CU(cudaMemcpyAsync(a->d_sym + off, symbols + off, mine, cudaMemcpyHostToDevice, c->stream));
        ncclResult_t r = nc->AllGather(a->d_sym + (size_t)c->comm.rank * slice, a->d_sym, slice, ncclInt8, c->comm.comm, c->stream);
        if (r != ncclSuccess) return fail(TDNA_E_CUDA, std::string("ncclAllGather: ") + nc->GetErrorString(r));
    } else {
        CU(DMALLOC(&a->d_sym, sym_bytes));
        if (!chunked) CU(cudaMemcpyAsync(a->d_sym, symbols, sym_bytes, cudaMemcpyDefault, c->stream));
    }
    CU(DMALLOC(&a->d_len, (size_t)n * 4));
    if (lengths) CU(cudaMemcpyAsync(a->d_len, lengths, (size_t)n * 4, cudaMemcpyDefault, c->stream));
    else {
        fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a->d_len, n, L);
        c->launches++;
    }
    if (!chunked) {
        CU(cudaStreamSynchronize(c->stream));  // the caller's buffers have been read
        if (pre) TRY((*pre)(a));
        return pack(a);
    }
    PlaneBits pb;
    TRY(pack_prepare(a, &pb));
    const bool tc_ok = c->opt_count_kernel != TDNA_KERNEL_POPC && tc_prepare(a) == 1;  // (synchronises the stream: buffers exist, lengths are in)
    if (!tc_ok) CU(cudaStreamSynchronize(c->stream));
    int *d_err = reinterpret_cast<int *>(c->d_counter);
    CU(cudaMemsetAsync(d_err, 0, 8, c->stream));
    hc_mark(c, 2);
    // labels and the plan of a pipelined pass (tdna_analyze_host) go first: their small pageable copies would otherwise queue behind
    // the symbols on the copy engine (measured: 0.6 ms)
    int pre_rc = TDNA_OK;
    if (pre) pre_rc = (*pre)(a);
    const int C = 8;
    const int64_t R = chunk_rows > 0 ? chunk_rows : round_up((n + C - 1) / C, tc::N);
    const int64_t panels_all = round_up(n, tc::N) / tc::N;
    cudaError_t e = cudaSuccess;
    int hook_rc = TDNA_OK;
    for (int ch = 0; ch < 15 && e == cudaSuccess; ch++) {  // every copy is enqueued first: none waits for a host decision below
        const int64_t r0 = ch * R, r1 = std::min<int64_t>(n, r0 + R);
        if (r0 >= n) break;
        const size_t off = (size_t)r0 * row_stride;
        const size_t bytes = r1 == n ? (size_t)(r1 - r0 - 1) * row_stride + L : (size_t)(r1 - r0) * row_stride;
        e = cudaMemcpyAsync(a->d_sym + off, symbols + off, bytes, cudaMemcpyHostToDevice, c->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_copy[ch], c->copy_stream);
    }
    hc_mark(c, 3);
    // A pipelined pass packs and encodes on a stream of its own, so that the small encode kernels run BESIDE the persistent tile
    // kernels of the earlier chunks instead of between them (0.5 ms on 50,000 x 658); the pass's stream waits for a chunk's operands.
    const bool side = hook != nullptr && tc_ok && pre_rc == TDNA_OK;
    cudaStream_t es = side ? c->enc_stream : c->stream;
    if (side) {  // buffers, lengths and the zeroed flags were prepared on the pass's stream
        e = cudaEventRecord(c->ev_enc[15], c->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(es, c->ev_enc[15], 0);
    }
    for (int ch = 0; ch < 15 && e == cudaSuccess && hook_rc == TDNA_OK && pre_rc == TDNA_OK; ch++) {
        const int64_t r0 = ch * R, r1 = std::min<int64_t>(n, r0 + R);
        if (r0 >= n) break;
        if (ch == 1) hc_mark(c, 4);
        e = cudaStreamWaitEvent(es, c->ev_copy[ch], 0);
        if (e != cudaSuccess) break;
        const int64_t prows = std::min<int64_t>(a->npad, r0 + R) - r0;  // whole 64-row panels (the last chunk runs to npad)
        const int64_t total = prows * a->W;
        pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, es>>>(a->d_sym, a->d_len, n, r0, prows, row_stride, a->W, a->P, L, pb, a->d_planes, d_err);
        c->launches++;
        if (tc_ok) tc_encode_range(a, es, r0 / tc::N, std::min<int64_t>(panels_all, (r0 + R) / tc::N) - r0 / tc::N);
        if (side) {
            e = cudaEventRecord(c->ev_enc[ch], es);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_enc[ch], 0);
            if (e != cudaSuccess) break;
        }
        if (hook && tc_ok) hook_rc = (*hook)(ch, r0, r1);
    }
    if (pre_rc != TDNA_OK) hook_rc = pre_rc;
    hc_mark(c, 5);
    int flags[6] = {0, 0, 0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags, d_err, 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && tc_ok) e = cudaMemcpyAsync(flags + 2, a->d_tcflag, 12, cudaMemcpyDeviceToHost, c->stream);
    if (side) cudaStreamSynchronize(es);
    cudaError_t e2 = cudaStreamSynchronize(c->copy_stream);  // whatever happened: no copy outlives the call
    cudaError_t e3 = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = e2 != cudaSuccess ? e2 : e3;
    hc_mark(c, 6);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(TDNA_E_CUDA, cudaGetErrorString(e));
    if (hook_rc != TDNA_OK) return hook_rc;
    if (flags[0] || flags[2]) return fail(TDNA_E_SYMBOL, "a symbol outside the normalised alphabet ACGT RYKMSWBDHVN - _ ?");
    a->packed = true;
    if (tc_ok) {
        if (tc_too_gappy(a, flags[4])) {  // rebuilt with the five-dimension code at first use (ensure_tc)
            tc_release(a);
            a->tc_dims5 = true;
            c->hint_dims5 = 1;
        } else {
            a->tc_has_amb = flags[3] != 0;
            a->tc_state = 1;
            c->hint_has_amb = a->tc_has_amb ? 1 : 0;
            // five dimensions are right for any list; four are cheaper when it has hardly any internal gap (tc_too_gappy's measure)
            if (a->kmode == KM_P_AMBIG) c->hint_dims5 = (int64_t)flags[4] * 2000 > (int64_t)a->n * a->L ? 1 : 0;
        }
    }
    return TDNA_OK;
}

static int aln_new(tdna_ctx *c, const uint8_t *symbols, int64_t n, int32_t L, int64_t row_stride, const int32_t *lengths, tdna_aln **out,
                   const std::function<int(tdna_aln *)> *pre = nullptr, const ChunkHook *hook = nullptr, int64_t chunk_rows = 0, const int *cfg = nullptr) {
    if (!c || !symbols || !out) return fail(TDNA_E_ARG, "null argument");
    if (n < 1 || n > 0x7fffffff) return fail(TDNA_E_ARG, "n out of range");
    if (L < 1 || L > 65535) return fail(TDNA_E_ARG, "L must be in [1, 65535] (two 16-bit counts share an accumulator)");
    if (row_stride < L) return fail(TDNA_E_ARG, "row_stride < L");
    if (lengths && !is_device_ptr(lengths)) {  // the header promises lengths[i] <= L; hold the caller to it where the host can see them
        for (int64_t i = 0; i < n; i++)
            if (lengths[i] < 0 || lengths[i] > L) return fail(TDNA_E_ARG, "lengths[i] must be in [0, L]");
    }
    CU(cudaSetDevice(c->device));
    tdna_aln *a = new tdna_aln();
    a->ctx = c;
    a->n = n;
    a->L = L;
    a->W = (L + 31) / 32;
    a->sym_stride = row_stride;  // one contiguous copy at PCIe speed; the pack kernel reads the caller's layout
    a->npad = round_up(n, TM);
    a->ld = round_up(n, 16);
    a->row_begin = 0;
    a->row_end = n;
    a->hist_size = c->opt_hist_slots;
    a->tc_dims5 = c->hint_dims5 == 1;  // uncorrected p: start on the code the previous list on this context ended up with
    if (cfg) { a->mode = cfg[0]; a->min_overlap = cfg[1]; a->ambig = cfg[2]; }  // tdna_analyze_host: planes and operands are built once, for the right mode
    int rc = aln_build(c, a, symbols, lengths, hook, chunk_rows, pre);
    if (rc != TDNA_OK) {
        std::string keep = g_err;
        tdna_alignment_destroy(a);
        g_err = keep;
        return rc;
    }
    *out = a;
    return TDNA_OK;
}

int32_t tdna_alignment_create(tdna_ctx *c, const uint8_t *symbols, int64_t n, int32_t L, int64_t row_stride, const int32_t *lengths,
                              tdna_aln **out) {
    return aln_new(c, symbols, n, L, row_stride, lengths, out);
}

// Create + labels + configuration + analysis in ONE call from host buffers: one downcall for a host that has just loaded or edited
// its list (the upload itself is chunked and overlapped with packing / operand encoding inside tdna_alignment_create's path).
// ---- the pass pipelined with the upload ---------------------------------------------------------------------------------------
// tdna_analyze_host on one GPU, Best Match rows from count kernel (b): the upload goes in row chunks of whole BANDS (64 row tiles);
// as soon as a chunk has been packed and encoded, the seed items whose blocks have all arrived run, the thresholds of the rows so
// far are rebuilt (they only ever decrease: every earlier test stays valid) and ONE region launch computes the tiles whose column
// tile lies in that chunk -- per band, a run of consecutive slots of the usual numbering.  The PCIe time hides behind the tiles of
// the earlier chunks.  Which instantiation runs (plain / GENERAL) is decided from the FIRST chunk's flags; if a later chunk
// contradicts it (its first ambiguity letter, too many internal gaps for the four-dimension code), or a buffer overflows, the
// result is dropped and the ordinary pass runs on the alignment, which is complete and resident by then.
static int set_labels_impl(tdna_aln *a, const int32_t *species_id, const int32_t *genus_id, const int32_t *epithet_rank, const int32_t *fullname_rank,
                           bool wait);
static int ensure_stream_state(tdna_aln *a);
static int stream_merge(tdna_aln *a, const int32_t *cq, const int32_t *cj, const double *cd, long long ncand, const int32_t *inval,
                        const int32_t *flags, bool counts_ready);
struct Pipeline {
    bool active = false;      // region launches are going out
    bool abandoned = false;   // decided against (or gave up): the ordinary pass follows the upload
    int64_t chunk_rows = 0;
    int bands_per_chunk = 1, nchunks = 0;
    std::vector<int64_t> seed_off;                 // [nchunks + 1] into the chunk-ordered seed list
    std::vector<int64_t> reg_off, reg_n, reg_items;  // per chunk: offset of its tables in d_plbuf (int64 words), runs, items
    int2 *d_seeds = nullptr;
    int64_t *d_tables = nullptr;
};

static int pipeline_plan(tdna_aln *a, Pipeline &pl) {
    tdna_ctx *c = a->ctx;
    const int band = TC_BAND;
    const int nrt = (int)(round_up(a->n, tc::M) / tc::M);
    const int64_t nct = round_up(a->n, tc::N) / tc::N;
    const int nbands = (nrt + band - 1) / band;
    const int maxc = std::max(1, c->opt_pipeline_chunks);
    pl.bands_per_chunk = (nbands + maxc - 1) / maxc;
    pl.chunk_rows = (int64_t)pl.bands_per_chunk * band * tc::M;
    pl.nchunks = (nbands + pl.bands_per_chunk - 1) / pl.bands_per_chunk;
    // the seed list, ordered by the chunk in which BOTH blocks of an item have arrived
    TRY(build_seed_list(a));
    const int64_t blocks_per_chunk = pl.chunk_rows / tc::N;
    std::vector<std::vector<int2>> by_chunk((size_t)pl.nchunks);
    for (const int2 &it : a->h_seedlist) {
        const int64_t blk = std::max<int64_t>(it.x / 2, it.y);
        by_chunk[(size_t)std::min<int64_t>(pl.nchunks - 1, blk / blocks_per_chunk)].push_back(it);
    }
    std::vector<int2> ordered;
    pl.seed_off.assign((size_t)pl.nchunks + 1, 0);
    for (int ch = 0; ch < pl.nchunks; ch++) {
        ordered.insert(ordered.end(), by_chunk[ch].begin(), by_chunk[ch].end());
        pl.seed_off[ch + 1] = (int64_t)ordered.size();
    }
    // the regions: for chunk ch (bands [B0, B1), column tiles [cj0, cj1)), one run of slot pairs per band b < B1
    std::vector<int64_t> pre((size_t)nbands + 1, 0);
    auto before = [&](int b, int64_t crel) {  // slots of band b before its column crel (relative to the band's first column)
        const int64_t tjmin = ((int64_t)b * band) >> 1, cols = std::max<int64_t>(nct - tjmin, 0), tri = std::min<int64_t>(cols, band >> 1);
        return crel <= tri ? crel * (crel + 1) : tri * (tri + 1) + (crel - tri) * band;
    };
    for (int b = 0; b < nbands; b++) {
        const int64_t tjmin = ((int64_t)b * band) >> 1, cols = std::max<int64_t>(nct - tjmin, 0);
        pre[b + 1] = pre[b] + before(b, cols);
    }
    std::vector<int64_t> tables;
    pl.reg_off.assign((size_t)pl.nchunks, 0);
    pl.reg_n.assign((size_t)pl.nchunks, 0);
    pl.reg_items.assign((size_t)pl.nchunks, 0);
    for (int ch = 0; ch < pl.nchunks; ch++) {
        const int B0 = ch * pl.bands_per_chunk, B1 = std::min(nbands, B0 + pl.bands_per_chunk);
        const int64_t cj0 = ((int64_t)B0 * band) >> 1, cj1 = std::min<int64_t>(nct, ((int64_t)B1 * band) >> 1);
        std::vector<int64_t> start, prefix(1, 0);
        for (int b = 0; b < B1; b++) {
            const int64_t tjmin = ((int64_t)b * band) >> 1;
            const int64_t lo = std::max(cj0, tjmin) - tjmin, hi = cj1 - tjmin;
            if (hi <= lo) continue;
            const int64_t s0 = pre[b] + before(b, lo), s1 = pre[b] + before(b, hi);
            if (s1 <= s0) continue;
            start.push_back(s0 / 2);  // slot counts are even: pairs (2u, 2u + 1) never straddle a run
            prefix.push_back(prefix.back() + (s1 - s0) / 2);
        }
        pl.reg_off[ch] = (int64_t)tables.size();
        pl.reg_n[ch] = (int64_t)start.size();
        pl.reg_items[ch] = prefix.back();
        tables.insert(tables.end(), start.begin(), start.end());
        tables.insert(tables.end(), prefix.begin(), prefix.end());
    }
    const size_t tbytes = tables.size() * 8, sbytes = ordered.size() * sizeof(int2);
    if (a->d_plbuf) { CU(DFREE(a->d_plbuf)); a->d_plbuf = nullptr; }
    CU(DMALLOC(&a->d_plbuf, tbytes + sbytes + 16));
    pl.d_tables = reinterpret_cast<int64_t *>(a->d_plbuf);
    pl.d_seeds = reinterpret_cast<int2 *>(reinterpret_cast<char *>(a->d_plbuf) + tbytes);
    CU(cudaMemcpyAsync(pl.d_tables, tables.data(), tbytes, cudaMemcpyHostToDevice, c->stream));
    if (sbytes) CU(cudaMemcpyAsync(pl.d_seeds, ordered.data(), sbytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // stack-lifetime host buffers
    return TDNA_OK;
}

static void pipeline_clear(tdna_aln *a) {
    a->pl_reg_start = a->pl_reg_prefix = nullptr;
    a->pl_reg_n = 0;
    a->pl_reg_items = 0;
    a->pl_seed_list = nullptr;
    a->pl_seed_pairs = -1;
    a->pl_general = -1;
}

static int pipeline_chunk(tdna_aln *a, Pipeline &pl, int ch, int64_t r0, int64_t r1) {
    tdna_ctx *c = a->ctx;
    if (pl.abandoned) return TDNA_OK;
    if (ch == 0) {
        // Which instantiation (plain / GENERAL) and which code (four / five dimensions) -- known for certain only when the LAST chunk
        // has been encoded.  The pass starts on what the previous list on this context turned out to be (a host re-analysing an
        // edited list: the same answer nearly always) and tdna_analyze_host checks the assumption once the upload is complete;
        // waiting for the first chunk's flags instead cost a host round trip of 0.1 ms before the first launch.
        if (c->hint_has_amb < 0 || (a->kmode == KM_P_AMBIG && c->hint_dims5 != 0)) { pl.abandoned = true; return TDNA_OK; }
        if (!a->tc_maps || a->tc_pair) { pl.abandoned = true; return TDNA_OK; }  // region launches exist for the multicast pairs only
        a->pl_general = c->hint_has_amb;
        TRY(ensure_stream_state(a));
        a->sp.species = a->d_species;
        a->sp.genus = a->d_genus;
        a->sp.epi = a->d_epi;
        a->sp.full = a->d_full;
        a->sp.panel_range = a->d_panel_range;
        a->sp.want = TDNA_WANT_BEST_MATCH;
        stream_init_kernel<<<(unsigned)((a->n + 255) / 256), 256, 0, c->stream>>>(a->sp, a->n, a->d_needc);
        c->launches++;
        CU(cudaGetLastError());
        a->last_kernel = TDNA_KERNEL_TENSOR;
        c->ev_valid = false;
        pl.active = true;
    }
    if (ch >= pl.nchunks) return TDNA_OK;
    a->pl_seed_list = pl.d_seeds + pl.seed_off[ch];
    a->pl_seed_pairs = (int)(pl.seed_off[ch + 1] - pl.seed_off[ch]);
    a->pl_reg_n = 0;
    if (a->pl_seed_pairs > 0) TRY(launch_tc<false>(a, 0, 0, 1, nullptr));
    const int64_t rows = std::min<int64_t>(a->n, r1);
    stream_thr_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, c->stream>>>(a->sp, rows);
    c->launches++;
    CU(cudaGetLastError());
    if (pl.reg_items[ch] > 0) {
        a->pl_reg_start = pl.d_tables + pl.reg_off[ch];
        a->pl_reg_prefix = a->pl_reg_start + pl.reg_n[ch];
        a->pl_reg_n = (int)pl.reg_n[ch];
        a->pl_reg_items = pl.reg_items[ch];
        TRY(launch_tc<false>(a, 1, 0, 1, nullptr));
        a->pl_reg_n = 0;
    }
    return TDNA_OK;
}

// after the upload: the queue -> records -> rows.  TDNA_E_TOO_LARGE: a buffer overflowed (the ordinary pass retries with larger ones)
static int pipeline_finish(tdna_aln *a, tdna_analysis *io) {
    tdna_ctx *c = a->ctx;
    a->pl_seed_pairs = -1;
    TRY(tc_resolve_queue(a));
    struct { unsigned long long ncand; int overflow; int pad; unsigned long long ncls[2]; unsigned long long nedge; unsigned long long qn; } h;
    CU(cudaMemcpyAsync(&h, a->d_stream_u64, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h.overflow) {
        a->need_cand = (long long)h.ncand;
        a->need_queue = (long long)h.qn;
        return TDNA_E_TOO_LARGE;
    }
    a->last_ncand = (long long)h.ncand;
    a->last_queue = (long long)h.qn;
    a->last_ncls[0] = a->last_ncls[1] = a->local_ncls[0] = a->local_ncls[1] = 0;
    a->last_nedge = a->local_nedge = 0;
    TRY(ensure_res(a, a->n));
    TRY(stream_merge(a, a->sp.cq, a->sp.cj, a->sp.cd, (long long)h.ncand, a->sp.inval, a->sp.flags, true));
    if (io->rows) CU(cudaMemcpyAsync(io->rows, a->d_res, (size_t)a->n * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    io->n_intra = io->n_inter = io->n_edges = io->n_intra_local = io->n_inter_local = io->n_edges_local = 0;
    io->n_hist_intra = io->n_hist_inter = io->hist_intra_pushes = io->hist_inter_pushes = 0;
    c->pipelined_passes++;
    return TDNA_OK;
}

// Create + labels + configuration + analysis in ONE call from host buffers: one downcall for a host that has just loaded or edited
// its list.  Configuration and labels go first (they are known before the symbols arrive: the planes and operands are built once,
// for the right mode), the upload is chunked and overlapped with packing / operand encoding, and -- Best Match rows on one GPU -- with
// the pass itself (above).
int32_t tdna_analyze_host(tdna_ctx *c, const tdna_host_list *in, tdna_analysis *io, tdna_aln **aln_out) {
    if (!c || !in || !io) return fail(TDNA_E_ARG, "null argument");
    if (aln_out) *aln_out = nullptr;
    if (in->mode < 0 || in->mode > 2) return fail(TDNA_E_ARG, "mode must be 0, 1 or 2");
    const bool labels = in->species_id && in->genus_id && in->epithet_rank && in->fullname_rank;
    const bool want_pipeline = labels && io->want == TDNA_WANT_BEST_MATCH && !c->comm.active() && c->opt_best_match_path != TDNA_PATH_DENSE &&
                               c->opt_count_kernel != TDNA_KERNEL_POPC && in->n >= 16384 && in->n <= (1 << 17) && in->L < 4096 &&
                               (in->mode == TDNA_PDM_K2P || (in->mode == TDNA_PDM_UNCORRECTED && in->ambiguous_allowed)) &&
                               in->symbols && !is_device_ptr(in->symbols) && !(TC_BAND & 1) && c->opt_pipeline_chunks > 0;
    for (double &x : c->hc_t) x = 0;
    hc_mark(c, 0);
    Pipeline pl;
    const int cfg[3] = {in->mode, in->min_overlap, in->ambiguous_allowed ? 1 : 0};
    const std::function<int(tdna_aln *)> pre = [&](tdna_aln *a) -> int {
        if (labels) TRY(set_labels_impl(a, in->species_id, in->genus_id, in->epithet_rank, in->fullname_rank, false));
        if (want_pipeline) TRY(pipeline_plan(a, pl));
        hc_mark(c, 1);
        return TDNA_OK;
    };
    tdna_aln *built = nullptr;
    const ChunkHook hook = [&](int ch, int64_t r0, int64_t r1) -> int {
        // (the alignment under construction: aln_build hands no pointer, the plan was made on it)
        return built ? pipeline_chunk(built, pl, ch, r0, r1) : TDNA_OK;
    };
    tdna_aln *a = nullptr;
    const std::function<int(tdna_aln *)> pre2 = [&](tdna_aln *x) -> int { built = x; return pre(x); };
    int64_t chunk_rows = 0;
    if (want_pipeline) {  // whole bands of row tiles, at most eight chunks (pipeline_plan makes the same choice)
        const int64_t band_rows = (int64_t)TC_BAND * tc::M, nbands = (in->n + band_rows - 1) / band_rows;
        const int64_t maxc = std::max(1, c->opt_pipeline_chunks);
        chunk_rows = ((nbands + maxc - 1) / maxc) * band_rows;
    }
    int rc = aln_new(c, in->symbols, in->n, in->L, in->row_stride, in->lengths, &a, &pre2, want_pipeline ? &hook : nullptr, chunk_rows, cfg);
    if (rc != TDNA_OK) return rc;
    bool done = false;
    if (pl.active) {
        // the first chunk's verdict must hold for the whole list: the operands were kept (not too gappy for their code) and no
        // ambiguity letter turned up after a plain start
        const bool consistent = a->tc_state == 1 && !(a->tc_has_amb && a->pl_general == 0);  // (GENERAL assumed, no letter found: still exact)
        if (consistent) {
            rc = pipeline_finish(a, io);
            done = rc == TDNA_OK;
            if (rc == TDNA_E_TOO_LARGE) rc = TDNA_OK;  // the ordinary pass below grows the buffers
        }
    }
    hc_mark(c, 7);
    pipeline_clear(a);
    if (rc == TDNA_OK) rc = tdna_set_config(a, in->mode, in->min_overlap, in->ambiguous_allowed);
    if (rc == TDNA_OK && !done) {
        rc = tdna_analyze(a, io);
    }
    hc_mark(c, 8);
    if (rc != TDNA_OK || !aln_out) {
        std::string keep = g_err;
        tdna_alignment_destroy(a);
        g_err = keep;
        hc_mark(c, 9);
        return rc;
    }
    *aln_out = a;
    hc_mark(c, 9);
    return TDNA_OK;
}

int32_t tdna_alignment_destroy(tdna_aln *a) {
    if (!a) return TDNA_OK;
    tdna_ctx *c = a->ctx;
    cudaSetDevice(c->device);
    if (c->mat_owner == a) c->mat_owner = nullptr;
    void *ptrs[] = {a->d_sym, a->d_len, a->d_planes, a->d_species, a->d_genus, a->d_epi, a->d_full, a->d_res, a->d_prefix, a->d_scratch,
                    a->d_needc, a->d_panel_range, a->d_range_rows, a->d_range_cols, a->sp.q, a->d_tcA, a->d_tcB, a->d_tcfull, a->d_tcamb, a->d_tcpanel_amb, a->d_tcflag, a->d_tcprefix[0], a->d_tcprefix[1], a->sp.thr, a->sp.mC, a->sp.inval, a->sp.flags, a->sp.cnt, a->sp.cq, a->sp.cj, a->sp.cd,
                    a->d_stream_u64, a->d_sprefix[0], a->d_sprefix[1], a->d_ptr, a->d_scan, a->d_seedlist, a->d_cursor, a->d_lj, a->d_ld,
                    a->x_send, a->x_recv, a->x_state, a->d_hkey[0], a->d_hkey[1], a->d_hcnt[0], a->d_hcnt[1], a->d_hctr, a->d_hent, a->d_gslot, a->d_gmem, a->d_gstart, a->d_plbuf, a->d_tcprefix4};
    for (void *p : ptrs)
        if (p) DFREE(p);  // stream-ordered: runs after everything queued on the context's stream
    delete a;
    return TDNA_OK;
}

int32_t tdna_alignment_set_labels(tdna_aln *a, const int32_t *species_id, const int32_t *genus_id, const int32_t *epithet_rank,
                                  const int32_t *fullname_rank) {
    return set_labels_impl(a, species_id, genus_id, epithet_rank, fullname_rank, true);
}
// wait = false (tdna_analyze_host): everything is enqueued and the call returns; the caller synchronises the stream before it
// returns to ITS caller (the label arrays may be pinned memory, which an asynchronous copy reads later)
static int set_labels_impl(tdna_aln *a, const int32_t *species_id, const int32_t *genus_id, const int32_t *epithet_rank, const int32_t *fullname_rank,
                           bool wait) {
    if (!a || !species_id || !genus_id || !epithet_rank || !fullname_rank) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    size_t bytes = (size_t)a->n * 4;
    int32_t **dst[] = {&a->d_species, &a->d_genus, &a->d_epi, &a->d_full};
    const int32_t *src[] = {species_id, genus_id, epithet_rank, fullname_rank};
    for (int i = 0; i < 4; i++) {
        if (!*dst[i]) CU(DMALLOC(dst[i], bytes));
        CU(cudaMemcpyAsync(*dst[i], src[i], bytes, cudaMemcpyDefault, c->stream));
    }
    // needc[q] (q is named and its species has another member: the streaming thresholds wait for a conspecific only where
    // one can exist) and the per-panel label ranges (tile-level test for class pairs), all on the device
    uint32_t hcap = 1024;
    while (hcap < 2 * (uint64_t)a->n) hcap <<= 1;
    int32_t *d_keys, *d_cnts;
    CU(DMALLOC(&d_keys, (size_t)hcap * 4));
    CU(DMALLOC(&d_cnts, (size_t)hcap * 4));
    CU(cudaMemsetAsync(d_keys, 0xff, (size_t)hcap * 4, c->stream));
    CU(cudaMemsetAsync(d_cnts, 0, (size_t)hcap * 4, c->stream));
    if (!a->d_needc) CU(DMALLOC(&a->d_needc, (size_t)a->n));
    size_t panels = (size_t)(a->npad / PR);
    if (!a->d_panel_range) CU(DMALLOC(&a->d_panel_range, panels * sizeof(int4)));
    unsigned gq = (unsigned)((a->n + 255) / 256);
    species_count_kernel<<<gq, 256, 0, c->stream>>>(a->d_species, a->n, d_keys, d_cnts, hcap - 1);
    needc_kernel<<<gq, 256, 0, c->stream>>>(a->d_species, a->n, d_keys, d_cnts, hcap - 1, a->d_needc);
    panel_range_kernel<<<(unsigned)panels, 32, 0, c->stream>>>(a->d_species, a->d_genus, a->n, a->d_panel_range);
    {   // the same per tile panel of count kernel (b)
        const int64_t nrt = round_up(a->n, tc::M) / tc::M, nct = round_up(a->n, tc::N) / tc::N;
        if (!a->d_range_rows) CU(DMALLOC(&a->d_range_rows, (size_t)nrt * sizeof(int4)));
        if (!a->d_range_cols) CU(DMALLOC(&a->d_range_cols, (size_t)nct * sizeof(int4)));
        tc::tile_range_kernel<<<(unsigned)((nrt + 255) / 256), 256, 0, c->stream>>>(a->d_panel_range, (int64_t)panels, tc::M / PR, a->d_range_rows, nrt);
        tc::tile_range_kernel<<<(unsigned)((nct + 255) / 256), 256, 0, c->stream>>>(a->d_panel_range, (int64_t)panels, tc::N / PR, a->d_range_cols, nct);
        c->launches += 2;
    }
    c->launches += 3;
    CU(cudaGetLastError());
    CU(DFREE(d_keys));
    CU(DFREE(d_cnts));
    // host copy of the species ids (seed list of count kernel (b)): straight from the caller's array when that is host memory
    if (!is_device_ptr(species_id)) a->h_species.assign(species_id, species_id + a->n);
    else {
        a->h_species.resize((size_t)a->n);
        CU(cudaMemcpyAsync(a->h_species.data(), a->d_species, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    if (!is_device_ptr(genus_id)) a->h_genus.assign(genus_id, genus_id + a->n);
    else {
        a->h_genus.resize((size_t)a->n);
        CU(cudaMemcpyAsync(a->h_genus.data(), a->d_genus, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    a->class_reach = -1;
    a->tcprefix_reach = -1;
    const bool device_labels = is_device_ptr(species_id) || is_device_ptr(genus_id);  // their host copies are still in flight
    if (wait || device_labels) CU(cudaStreamSynchronize(c->stream));  // do not return before the copies have read the caller's arrays
    a->has_labels = true;
    a->seedlist_valid = false;
    a->genus_lists_valid = false;
    return TDNA_OK;
}

int32_t tdna_set_config(tdna_aln *a, int32_t mode, int32_t min_overlap, int32_t ambiguous_allowed) {
    if (!a) return fail(TDNA_E_ARG, "null alignment");
    if (mode < 0 || mode > 2) return fail(TDNA_E_ARG, "mode must be 0, 1 or 2");
    CU(cudaSetDevice(a->ctx->device));
    bool repack = !a->packed || mode != a->mode || (mode == TDNA_PDM_UNCORRECTED && (ambiguous_allowed != 0) != (a->ambig != 0));
    a->mode = mode;
    a->min_overlap = min_overlap;
    a->ambig = ambiguous_allowed ? 1 : 0;
    a->mat_valid = false;
    if (repack) TRY(pack(a));
    if (a->tc_state == -1) a->tc_state = 0;  // eligibility depends on the mode: ask again
    return TDNA_OK;
}

int32_t tdna_set_row_range(tdna_aln *a, int64_t row_begin, int64_t row_end) {
    if (!a || row_begin < 0 || row_end > a->n || row_begin > row_end) return fail(TDNA_E_ARG, "bad row range");
    if (row_begin % TM != 0) return fail(TDNA_E_ARG, "row_begin must be a multiple of 128");
    a->row_begin = row_begin;
    a->row_end = row_end;
    a->mat_valid = false;
    return TDNA_OK;
}

static int row_query(tdna_aln *a, int64_t q, int64_t j0, int64_t j1, double *d_dev, int32_t *sh_dev, tdna_counts *cnt_dev) {
    tdna_ctx *c = a->ctx;
    unsigned grid = (unsigned)((j1 - j0 + 255) / 256);
    switch (a->kmode) {
        case KM_P_AMBIG: row_from_planes_kernel<KM_P_AMBIG><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, q, j0, j1, a->min_overlap, d_dev, sh_dev, cnt_dev); break;
        case KM_P_NOAMBIG: row_from_planes_kernel<KM_P_NOAMBIG><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, q, j0, j1, a->min_overlap, d_dev, sh_dev, cnt_dev); break;
        case KM_K2P: row_from_planes_kernel<KM_K2P><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, q, j0, j1, a->min_overlap, d_dev, sh_dev, cnt_dev); break;
        default: row_from_planes_kernel<KM_TRANS><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, q, j0, j1, a->min_overlap, d_dev, sh_dev, cnt_dev); break;
    }
    c->launches++;
    CU(cudaGetLastError());
    return TDNA_OK;
}

int32_t tdna_pair(tdna_aln *a, int64_t i, int64_t j, tdna_counts *counts, double *d) {
    TRY(ensure_packed(a));
    if (i < 0 || j < 0 || i >= a->n || j >= a->n) return fail(TDNA_E_ARG, "index out of range");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    struct Scratch { double d; tdna_counts c; } *dev;
    CU(DMALLOC(&dev, sizeof(Scratch)));
    int rc = row_query(a, i, j, j + 1, &dev->d, nullptr, &dev->c);
    if (rc == TDNA_OK) {
        Scratch h;
        cudaError_t e = cudaMemcpyAsync(&h, dev, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
        else {
            if (d) *d = h.d;
            if (counts) *counts = h.c;
        }
    }
    DFREE(dev);
    return rc;
}

int32_t tdna_pairs(tdna_aln *a, const int32_t *ii, const int32_t *jj, int64_t m, tdna_counts *counts, double *d) {
    TRY(ensure_packed(a));
    if (m < 0 || (m > 0 && (!ii || !jj))) return fail(TDNA_E_ARG, "bad argument");
    if (m == 0) return TDNA_OK;
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    const bool di = is_device_ptr(ii), dj = is_device_ptr(jj), dd = d && is_device_ptr(d), dc = counts && is_device_ptr(counts);
    int32_t *gi = const_cast<int32_t *>(ii), *gj = const_cast<int32_t *>(jj);
    double *gd = d;
    tdna_counts *gc = counts;
    int rc = TDNA_OK;
    cudaError_t e = cudaSuccess;
    if (!di) { e = DMALLOC(&gi, (size_t)m * 4); if (e == cudaSuccess) e = cudaMemcpyAsync(gi, ii, (size_t)m * 4, cudaMemcpyDefault, c->stream); }
    if (e == cudaSuccess && !dj) { e = DMALLOC(&gj, (size_t)m * 4); if (e == cudaSuccess) e = cudaMemcpyAsync(gj, jj, (size_t)m * 4, cudaMemcpyDefault, c->stream); }
    if (e == cudaSuccess && d && !dd) e = DMALLOC(&gd, (size_t)m * 8);
    if (e == cudaSuccess && counts && !dc) e = DMALLOC(&gc, (size_t)m * sizeof(tdna_counts));
    if (e == cudaSuccess) {
        // indices are checked on the host when they are host memory; device index arrays are the caller's responsibility
        if (!di) for (int64_t k = 0; k < m; k++) if (ii[k] < 0 || ii[k] >= a->n) { rc = fail(TDNA_E_ARG, "pair index out of range"); break; }
        if (rc == TDNA_OK && !dj) for (int64_t k = 0; k < m; k++) if (jj[k] < 0 || jj[k] >= a->n) { rc = fail(TDNA_E_ARG, "pair index out of range"); break; }
    }
    if (e == cudaSuccess && rc == TDNA_OK) {
        unsigned grid = (unsigned)((m + 255) / 256);
        switch (a->kmode) {
            case KM_P_AMBIG: pairs_from_planes_kernel<KM_P_AMBIG><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, gi, gj, m, a->min_overlap, gd, gc); break;
            case KM_P_NOAMBIG: pairs_from_planes_kernel<KM_P_NOAMBIG><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, gi, gj, m, a->min_overlap, gd, gc); break;
            case KM_K2P: pairs_from_planes_kernel<KM_K2P><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, gi, gj, m, a->min_overlap, gd, gc); break;
            default: pairs_from_planes_kernel<KM_TRANS><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, gi, gj, m, a->min_overlap, gd, gc); break;
        }
        c->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess && d && !dd) e = cudaMemcpyAsync(d, gd, (size_t)m * 8, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess && counts && !dc) e = cudaMemcpyAsync(counts, gc, (size_t)m * sizeof(tdna_counts), cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    if (!di && gi) DFREE(gi);
    if (!dj && gj) DFREE(gj);
    if (d && !dd && gd) DFREE(gd);
    if (counts && !dc && gc) DFREE(gc);
    if (e != cudaSuccess) return fail(TDNA_E_CUDA, cudaGetErrorString(e));
    return rc;
}

int32_t tdna_row_distances(tdna_aln *a, int64_t q, double *d, int32_t *shared) {
    TRY(ensure_packed(a));
    if (q < 0 || q >= a->n || !d) return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    double *dd;
    int32_t *ds = nullptr;
    CU(DMALLOC(&dd, (size_t)a->n * 8));
    if (shared) CU(DMALLOC(&ds, (size_t)a->n * 4));
    int rc = row_query(a, q, 0, a->n, dd, ds, nullptr);
    if (rc == TDNA_OK) {
        cudaError_t e = cudaMemcpyAsync(d, dd, (size_t)a->n * 8, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess && shared) e = cudaMemcpyAsync(shared, ds, (size_t)a->n * 4, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
    }
    DFREE(dd);
    if (ds) DFREE(ds);
    return rc;
}

int32_t tdna_query_distances(tdna_aln *a, const uint8_t *symbols, int32_t len, double *d, int32_t *shared, tdna_counts *counts) {
    TRY(ensure_packed(a));
    if (!symbols || len < 0 || !d) return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    // sites past the alignment's width meet no row (every comparison stops at the shorter sequence, Sequence.java:1417-1419)
    const int32_t ql = len < a->L ? len : a->L;
    const size_t sym_bytes = (size_t)a->W * 32, plane_words = (size_t)PR * a->P * a->W;
    uint8_t *q_sym = nullptr;
    int32_t *q_len = nullptr;
    uint32_t *q_planes = nullptr;
    double *dd = nullptr;
    int32_t *ds = nullptr;
    tdna_counts *dc = nullptr;
    int rc = TDNA_OK;
    auto cleanup = [&]() {
        for (void *p : {(void *)q_sym, (void *)q_len, (void *)q_planes, (void *)dd, (void *)ds, (void *)dc})
            if (p) DFREE(p);
    };
#define QCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return fail(TDNA_E_CUDA, cudaGetErrorString(e_)); } } while (0)
    QCU(DMALLOC(&q_sym, sym_bytes));
    QCU(DMALLOC(&q_len, 4));
    QCU(DMALLOC(&q_planes, plane_words * 4));
    QCU(DMALLOC(&dd, (size_t)a->n * 8));
    if (shared) QCU(DMALLOC(&ds, (size_t)a->n * 4));
    if (counts) QCU(DMALLOC(&dc, (size_t)a->n * sizeof(tdna_counts)));
    QCU(cudaMemsetAsync(q_sym, '?', sym_bytes, c->stream));
    if (ql > 0) QCU(cudaMemcpyAsync(q_sym, symbols, (size_t)ql, cudaMemcpyDefault, c->stream));
    QCU(cudaMemcpyAsync(q_len, &ql, 4, cudaMemcpyHostToDevice, c->stream));
    int *d_err = reinterpret_cast<int *>(c->d_counter);
    QCU(cudaMemsetAsync(d_err, 0, 8, c->stream));
    PlaneBits pb;
    switch (a->kmode) {
        case KM_P_AMBIG: pb = plane_bits<KM_P_AMBIG>(); break;
        case KM_P_NOAMBIG: pb = plane_bits<KM_P_NOAMBIG>(); break;
        case KM_K2P: pb = plane_bits<KM_K2P>(); break;
        default: pb = plane_bits<KM_TRANS>(); break;
    }
    const int64_t total = (int64_t)PR * a->W;
    pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(q_sym, q_len, 1, 0, PR, (int64_t)sym_bytes, a->W, a->P, a->L, pb, q_planes, d_err);
    const unsigned grid = (unsigned)((a->n + 255) / 256);
    switch (a->kmode) {
        case KM_P_AMBIG: query_from_planes_kernel<KM_P_AMBIG><<<grid, 256, 0, c->stream>>>(q_planes, a->d_planes, a->W, a->n, a->min_overlap, dd, ds, dc); break;
        case KM_P_NOAMBIG: query_from_planes_kernel<KM_P_NOAMBIG><<<grid, 256, 0, c->stream>>>(q_planes, a->d_planes, a->W, a->n, a->min_overlap, dd, ds, dc); break;
        case KM_K2P: query_from_planes_kernel<KM_K2P><<<grid, 256, 0, c->stream>>>(q_planes, a->d_planes, a->W, a->n, a->min_overlap, dd, ds, dc); break;
        default: query_from_planes_kernel<KM_TRANS><<<grid, 256, 0, c->stream>>>(q_planes, a->d_planes, a->W, a->n, a->min_overlap, dd, ds, dc); break;
    }
    c->launches += 2;
    QCU(cudaGetLastError());
    int err = 0;
    QCU(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
    QCU(cudaMemcpyAsync(d, dd, (size_t)a->n * 8, cudaMemcpyDefault, c->stream));
    if (shared) QCU(cudaMemcpyAsync(shared, ds, (size_t)a->n * 4, cudaMemcpyDefault, c->stream));
    if (counts) QCU(cudaMemcpyAsync(counts, dc, (size_t)a->n * sizeof(tdna_counts), cudaMemcpyDefault, c->stream));
    QCU(cudaStreamSynchronize(c->stream));
#undef QCU
    cleanup();
    if (err) return fail(TDNA_E_SYMBOL, "a symbol outside the normalised alphabet ACGT RYKMSWBDHVN - _ ?");
    return rc;
}

int32_t tdna_matrix(tdna_aln *a, double *d, tdna_counts *counts) {
    if (!a || !d) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    if (counts) {  // diagnostic path: rectangular bands with the integer tuples as well
        TRY(ensure_packed(a));
        int64_t band = band_rows_for(a);
        int64_t cb = c->budget / (a->n * 16) / TM * TM;
        if (cb < TM) cb = TM;
        if (band > cb) band = cb;
        TRY(ensure_mat(a, band));
        a->mat_valid = false;
        c->mat_owner = nullptr;
        tdna_counts *dc;
        CU(DMALLOC(&dc, (size_t)band * a->n * sizeof(tdna_counts)));
        int rc = TDNA_OK;
        for (int64_t b0 = a->row_begin; b0 < a->row_end && rc == TDNA_OK; b0 += band) {
            int64_t b1 = b0 + band < a->row_end ? b0 + band : a->row_end;
            rc = launch_tiles(a, b0, b1, false, a->d_mat, a->ld, dc);
            if (rc != TDNA_OK) break;
            cudaError_t e = cudaMemcpy2DAsync(d + (b0 - a->row_begin) * a->n, (size_t)a->n * 8, a->d_mat, (size_t)a->ld * 8, (size_t)a->n * 8,
                                              (size_t)(b1 - b0), cudaMemcpyDefault, c->stream);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(counts + (b0 - a->row_begin) * a->n, dc, (size_t)(b1 - b0) * a->n * sizeof(tdna_counts), cudaMemcpyDefault, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
        }
        DFREE(dc);
        return rc;
    }
    TRY(for_each_band(a, [&](int64_t b0, int64_t b1) -> int {
        CU(cudaMemcpy2DAsync(d + (b0 - a->row_begin) * a->n, (size_t)a->n * 8, a->d_mat, (size_t)a->ld * 8, (size_t)a->n * 8, (size_t)(b1 - b0),
                             cudaMemcpyDefault, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return TDNA_OK;
    }));
    return TDNA_OK;
}

int32_t tdna_matrix_compute(tdna_aln *a, const double **d_dev) {
    if (!a) return fail(TDNA_E_ARG, "null alignment");
    CU(cudaSetDevice(a->ctx->device));
    TRY(ensure_packed(a));
    if (band_rows_for(a) != a->row_end - a->row_begin) return fail(TDNA_E_TOO_LARGE, "matrix exceeds the memory budget");
    TRY(for_each_band(a, [&](int64_t, int64_t) -> int { return TDNA_OK; }));
    if (d_dev) *d_dev = a->d_mat;
    return TDNA_OK;
}

static int fill_shared(tdna_aln *a, int64_t r0, int64_t r1) {
    tdna_ctx *c = a->ctx;
    int64_t rows = r1 - r0;
    if (rows <= 0) return TDNA_OK;
    unsigned grid = (unsigned)((rows * 2 + 255) / 256);
    switch (a->kmode) {
        case KM_P_AMBIG: fill_shared_kernel<KM_P_AMBIG><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, r0, r1, a->d_res); break;
        case KM_P_NOAMBIG: fill_shared_kernel<KM_P_NOAMBIG><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, r0, r1, a->d_res); break;
        case KM_K2P: fill_shared_kernel<KM_K2P><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, r0, r1, a->d_res); break;
        default: fill_shared_kernel<KM_TRANS><<<grid, 256, 0, c->stream>>>(a->d_planes, a->W, r0, r1, a->d_res); break;
    }
    c->launches++;
    CU(cudaGetLastError());
    return TDNA_OK;
}

static int best_match_dense(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    TRY(for_each_band(a, [&](int64_t b0, int64_t b1) -> int {
        int64_t grid = (b1 - b0) < c->sm_count ? (b1 - b0) : c->sm_count;
        row_summary_kernel<<<(unsigned)grid, RS_THREADS, 0, c->stream>>>(a->d_mat, a->ld, a->n, b0, b1 - b0, a->d_species, a->d_res);
        c->launches++;
        CU(cudaGetLastError());
        return TDNA_OK;
    }));
    return fill_shared(a, a->row_begin, a->row_end);
}

static int ensure_stream_state(tdna_aln *a) {
    if (a->stream_broken) return fail(TDNA_E_TOO_LARGE, "the streaming buffers could not be regrown");  // the caller takes the dense path
    if (a->stream_alloc) return TDNA_OK;
    tdna_ctx *c = a->ctx;
    size_t n = (size_t)a->n;
    // (the hints of an earlier list never ask for more than 2^28 entries -- 4 GB -- beyond the default: a short, hostile list must not
    // size the buffers of a long one)
    long long cap = std::min<long long>((long long)((double)n * c->hint_cand_per_row), std::max<long long>((long long)n * 64, 1ll << 28));
    if (cap < (1ll << 22)) cap = 1ll << 22;
    if (cap * 16 > c->budget / 2) cap = c->budget / 32;
    StreamParams &sp = a->sp;
    CU(DMALLOC(&sp.thr, n * 4));
    CU(DMALLOC(&sp.mC, n * 24));  // mC[n] followed by mPar[2n]: one buffer, so that parts can all-reduce(min) it in one go
    sp.mPar = sp.mC + n;
    CU(DMALLOC(&sp.inval, n * 4));
    CU(DMALLOC(&sp.flags, n * 4));
    CU(DMALLOC(&sp.cnt, n * 4));
    CU(DMALLOC(&sp.cq, (size_t)cap * 4));
    CU(DMALLOC(&sp.cj, (size_t)cap * 4));
    CU(DMALLOC(&sp.cd, (size_t)cap * 8));
    // kernel (b)'s queue holds every pair that passes the SEEDED threshold of either of its rows (about four times the final
    // candidates): 256 per row, 16 bytes each
    long long qcap = std::min<long long>((long long)((double)n * c->hint_queue_per_row), std::max<long long>((long long)n * 256, 1ll << 28));
    if (qcap < (1ll << 22)) qcap = 1ll << 22;
    if (qcap * 16 > c->budget / 2) qcap = c->budget / 32;
    CU(DMALLOC(&sp.q, (size_t)qcap * 16));
    sp.qcap = qcap;
    CU(DMALLOC(&a->d_stream_u64, 64));  // ncand, overflow, ncls[2], nedge, qn
    sp.ncand = reinterpret_cast<unsigned long long *>(a->d_stream_u64);
    sp.overflow = reinterpret_cast<int *>(reinterpret_cast<char *>(a->d_stream_u64) + 8);
    sp.ncls = sp.ncand + 2;
    sp.nedge = sp.ncand + 4;
    sp.qn = sp.ncand + 5;
    sp.want = TDNA_WANT_BEST_MATCH;
    sp.cap = cap;
    sp.species = a->d_species;
    CU(DMALLOC(&a->d_ptr, (n + 1) * 8));
    CU(DMALLOC(&a->d_cursor, n * 4));
    a->stream_alloc = true;
    return TDNA_OK;
}

// After an overflow: buffers sized to what the failed pass asked for (+25 %), within a quarter of the memory budget.  Lists with
// hundreds of conspecifics per species need it -- every conspecific closer than the nearest other species is a candidate.
// false: nothing to grow, or no room (the caller falls back to the next path).
static bool grow_stream_buffers(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    StreamParams &sp = a->sp;
    long long cap = sp.cap, qcap = sp.qcap;
    if (a->need_cand > cap) cap = a->need_cand + a->need_cand / 4;
    if (a->need_queue > qcap) qcap = a->need_queue + a->need_queue / 4;
    a->need_cand = a->need_queue = 0;
    if (cap == sp.cap && qcap == sp.qcap) return false;
    if (cap * 16 + qcap * 16 > c->budget / 4) return false;
    if (a->n > 0) {  // the next alignment on this context starts there (bounded: 4,096 entries per row)
        c->hint_cand_per_row = std::min(4096.0, std::max(c->hint_cand_per_row, (double)cap / (double)a->n));
        c->hint_queue_per_row = std::min(4096.0, std::max(c->hint_queue_per_row, (double)qcap / (double)a->n));
    }
    if (cap != sp.cap) {
        DFREE(sp.cq); DFREE(sp.cj); DFREE(sp.cd);
        sp.cq = sp.cj = nullptr; sp.cd = nullptr;
        if (cudaSuccess != DMALLOC(&sp.cq, (size_t)cap * 4) || cudaSuccess != DMALLOC(&sp.cj, (size_t)cap * 4) || cudaSuccess != DMALLOC(&sp.cd, (size_t)cap * 8)) {
            cudaGetLastError();
            a->stream_broken = true;
            return false;
        }
        sp.cap = cap;
    }
    if (qcap != sp.qcap) {
        DFREE(sp.q);
        sp.q = nullptr;
        if (cudaSuccess != DMALLOC(&sp.q, (size_t)qcap * 16)) {
            cudaGetLastError();
            a->stream_broken = true;
            return false;
        }
        sp.qcap = qcap;
    }
    return true;
}

// seed: state init + the diagonal blocks of this part (thresholds from list neighbours)
static int stream_seed(tdna_aln *a, int part, int nparts) {
    tdna_ctx *c = a->ctx;
    TRY(ensure_packed(a));
    TRY(ensure_stream_state(a));
    a->sp.species = a->d_species;
    a->sp.genus = a->d_genus;
    a->sp.epi = a->d_epi;
    a->sp.full = a->d_full;
    a->sp.panel_range = a->d_panel_range;
    stream_init_kernel<<<(unsigned)((a->n + 255) / 256), 256, 0, c->stream>>>(a->sp, a->n, a->d_needc);
    c->launches++;
    CU(cudaGetLastError());
    return launch_stream(a, part, nparts, 1);
}

// bulk: thresholds from the minima (all-reduced by the caller when there are several parts), then the
// rest of the triangle.  rc TDNA_E_TOO_LARGE: the candidate buffer overflowed (thresholds never
// tightened: e.g. one species only, or a list order in which neighbours are unrelated)
static int stream_bulk_launch(tdna_aln *a, int part, int nparts) {  // enqueue only: no host round trip
    tdna_ctx *c = a->ctx;
    stream_thr_kernel<<<(unsigned)((a->n + 255) / 256), 256, 0, c->stream>>>(a->sp, a->n);
    c->launches++;
    CU(cudaGetLastError());
    return launch_stream(a, part, nparts, 2);
}
struct StreamCounters { unsigned long long ncand; int overflow; int pad; unsigned long long ncls[2]; unsigned long long nedge; unsigned long long qn; };
static int stream_bulk(tdna_aln *a, int part, int nparts, long long *ncand_out) {
    tdna_ctx *c = a->ctx;
    TRY(stream_bulk_launch(a, part, nparts));
    StreamCounters h;
    CU(cudaMemcpyAsync(&h, a->d_stream_u64, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h.overflow) {  // both counters kept counting past their buffers: they are the demand (grow_stream_buffers)
        a->need_cand = (long long)h.ncand;
        a->need_queue = (long long)h.qn;
        return fail(TDNA_E_TOO_LARGE, "streaming Best Match: candidate buffer overflow");
    }
    a->last_ncand = (long long)h.ncand;
    a->last_queue = (long long)h.qn;
    a->last_ncls[0] = a->local_ncls[0] = (long long)h.ncls[0];
    a->last_ncls[1] = a->local_ncls[1] = (long long)h.ncls[1];
    a->last_nedge = a->local_nedge = (long long)h.nedge;
    if (ncand_out) *ncand_out = (long long)h.ncand;
    return TDNA_OK;
}

static int stream_part(tdna_aln *a, int part, int nparts, long long *ncand_out) {
    TRY(stream_seed(a, part, nparts));
    return stream_bulk(a, part, nparts, ncand_out);
}


// CSR row pointers d_ptr[0..n] from the per-row record counts: three launches (tdna_kernels.cuh: scan_*_kernel)
static int row_pointers(tdna_aln *a, const int *cnt) {
    tdna_ctx *c = a->ctx;
    const int64_t nb = (a->n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (!a->d_scan) CU(DMALLOC(&a->d_scan, (size_t)nb * 8));
    scan_block_sums_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(cnt, a->n, a->d_scan);
    scan_block_offsets_kernel<<<1, 1024, 0, c->stream>>>(a->d_scan, nb, a->d_ptr + a->n);
    scan_apply_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(cnt, a->n, a->d_scan, a->d_ptr);
    c->launches += 3;
    CU(cudaGetLastError());
    return TDNA_OK;
}

// G lanes per row from the mean list length (tdna_kernels.cuh: stream_finalize_kernel)
static void launch_finalize(tdna_aln *a, const int *inval, const int *flags, int64_t q0, int64_t q1, long long records) {
    tdna_ctx *c = a->ctx;
    const int64_t rows = q1 - q0;
    if (rows <= 0) return;
    const long long mean = records / rows;
    if (mean < 24) stream_finalize_kernel<8><<<(unsigned)((rows * 8 + 255) / 256), 256, 0, c->stream>>>(a->d_ptr, a->d_lj, a->d_ld, inval, flags, a->d_species, a->n, a->d_res, q0, q1);
    else if (mean < 64) stream_finalize_kernel<16><<<(unsigned)((rows * 16 + 255) / 256), 256, 0, c->stream>>>(a->d_ptr, a->d_lj, a->d_ld, inval, flags, a->d_species, a->n, a->d_res, q0, q1);
    else stream_finalize_kernel<32><<<(unsigned)((rows * 32 + 255) / 256), 256, 0, c->stream>>>(a->d_ptr, a->d_lj, a->d_ld, inval, flags, a->d_species, a->n, a->d_res, q0, q1);
}

static int stream_merge(tdna_aln *a, const int32_t *cq, const int32_t *cj, const double *cd, long long ncand, const int32_t *inval,
                        const int32_t *flags, bool counts_ready) {
    tdna_ctx *c = a->ctx;
    TRY(ensure_stream_state(a));
    TRY(ensure_res(a, a->n));
    size_t n = (size_t)a->n;
    unsigned gr = (unsigned)((ncand + 255) / 256);
    if (gr > 65535u * 4) gr = 65535u * 4;
    if (gr < 1) gr = 1;
    int *cnt = a->sp.cnt;
    if (!counts_ready) {  // gathered records: recount per row
        CU(cudaMemsetAsync(cnt, 0, n * 4, c->stream));
        count_rows_kernel<<<gr, 256, 0, c->stream>>>(cq, ncand, cnt);
        c->launches++;
    }
    TRY(row_pointers(a, cnt));
    if (ncand > a->list_cap) {
        if (a->d_lj) CU(DFREE(a->d_lj));
        if (a->d_ld) CU(DFREE(a->d_ld));
        a->d_lj = nullptr;
        a->d_ld = nullptr;
        long long cap = ncand + ncand / 4 + 1024;
        CU(DMALLOC(&a->d_lj, (size_t)cap * 4));
        CU(DMALLOC(&a->d_ld, (size_t)cap * 8));
        a->list_cap = cap;
    }
    CU(cudaMemsetAsync(a->d_cursor, 0, n * 4, c->stream));
    if (ncand > 0) {
        stream_scatter_kernel<<<gr, 256, 0, c->stream>>>(cq, cj, cd, ncand, a->d_ptr, a->d_cursor, a->d_lj, a->d_ld);
        c->launches++;
    }
    launch_finalize(a, inval, flags, 0, a->n, ncand);
    c->launches++;
    CU(cudaGetLastError());
    return fill_shared(a, 0, a->n);
}


// ---- the same pass shared by the GPUs of a communicator (tdna_comm_init) ------------------------------------------------------
// Every rank: seed its share of the diagonal blocks -> ncclAllReduce(MIN) of the 3n per-row minima -> its share of the triangle ->
// queue -> candidate records -> records routed to the rank that OWNS the row (grouped ncclSend / ncclRecv of fixed-size segments,
// each carrying its own count) -> ncclAllReduce(SUM) of the per-row state and the status words -> exact two-sweep summaries of the
// rank's own rows -> ncclAllGather of the row records.  Everything is enqueued on the context's stream; the host waits ONCE, at
// the end, and reads a status block that is identical on every rank (it was all-reduced), so that retries with larger buffers and
// the dense fallback are taken by all ranks together.  Results are bit-identical to one GPU's: min / sum / or are order-free and
// every rank's record list is complete below the row's final threshold.
#define NC(call)                                                                                                          \
    do {                                                                                                                  \
        ncclResult_t r_ = (call);                                                                                         \
        if (r_ != ncclSuccess && r_ != ncclInProgress) return fail(TDNA_E_CUDA, std::string(#call) + ": " + nc->GetErrorString(r_)); \
    } while (0)

static int64_t multi_seg_records(const tdna_aln *a, int G) {
    int64_t seg = (48 * a->n) / ((int64_t)G * G);
    if (seg < (1 << 14)) seg = 1 << 14;
    return seg * a->x_scale;
}

static int ensure_multi_buffers(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    const int G = c->comm.nranks;
    const int64_t seg = multi_seg_records(a, G);
    if (a->x_send && a->x_seg == seg && a->x_nranks == G) return TDNA_OK;
    for (long long **p : {&a->x_send, &a->x_recv, &a->x_state})
        if (*p) { CU(DFREE(*p)); *p = nullptr; }
    const size_t bytes = (size_t)G * (size_t)multi_seg_words(seg, tdna_rows_per_part(a->n, G)) * 8;
    CU(DMALLOC(&a->x_send, bytes));
    CU(DMALLOC(&a->x_recv, bytes));
    CU(DMALLOC(&a->x_state, (size_t)MULTI_STATE_EXTRA * 8));
    a->x_seg = seg;
    a->x_nranks = G;
    return TDNA_OK;
}

// One attempt.  *status: [0] segment overflow, [1] candidate / queue overflow (summed over the ranks), [2..4] class pushes / edges.
static int multi_attempt(tdna_aln *a, bool want_bm, int64_t *status) {
    tdna_ctx *c = a->ctx;
    const NcclApi *nc = c->comm.api;
    const int G = c->comm.nranks, rank = c->comm.rank;
    const int64_t n = a->n, rpp = tdna_rows_per_part(n, G);
    const int64_t q0 = std::min<int64_t>(n, (int64_t)rank * rpp), q1 = std::min<int64_t>(n, q0 + rpp);
    TRY(ensure_multi_buffers(a));
    TRY(ensure_res(a, (int64_t)G * rpp));
    const int64_t seg = a->x_seg;
    const long long segw = multi_seg_words(seg, rpp);
    c->n_ph = 0;
#define MARK() do { if (c->n_ph < 12) cudaEventRecord(c->ev_ph[c->n_ph++], c->stream); } while (0)
    MARK();  // 0: start
    // Lists of up to 131,072 rows: EVERY rank seeds the whole diagonal (70 us at 50,000 rows) -- the minima are then the all-reduced
    // ones by construction and the MIN all-reduce (a 56 us latency at eight ranks) is not needed.  Larger lists share the seed pass.
    const bool seed_all = n <= (1 << 17);
    if (seed_all) TRY(stream_seed(a, 0, 1));
    else TRY(stream_seed(a, rank, G));
    MARK();  // 1: seeded
    if (!seed_all) NC(nc->AllReduce(a->sp.mC, a->sp.mC, (size_t)3 * n, ncclUint64, ncclMin, c->comm.comm, c->stream));
    MARK();  // 2: minima reduced
    TRY(stream_bulk_launch(a, rank, G));
    MARK();  // 3: bulk + queue resolved
    if (want_bm) {
        multi_zero_kernel<<<1, 64, 0, c->stream>>>(a->x_send, G, segw, a->x_state);
        route_records_dev_kernel<<<(unsigned)c->sm_count * 4, 256, 0, c->stream>>>(a->sp, rpp, G, seg, segw, a->x_send, a->x_state);
        multi_pack_kernel<<<(unsigned)(((int64_t)G * rpp + 255) / 256), 256, 0, c->stream>>>(a->sp, n, rpp, G, seg, segw, a->x_send, a->x_state);
        c->launches += 3;
        CU(cudaGetLastError());
        NC(nc->GroupStart());
        for (int g = 0; g < G; g++) {
            NC(nc->Send(a->x_send + (size_t)g * segw, (size_t)segw, ncclInt64, g, c->comm.comm, c->stream));
            NC(nc->Recv(a->x_recv + (size_t)g * segw, (size_t)segw, ncclInt64, g, c->comm.comm, c->stream));
        }
        NC(nc->GroupEnd());
        MARK();  // 4: records, state and status exchanged
        const long long total = (long long)G * seg;  // upper bound of the records received
        if (total > a->list_cap) {
            if (a->d_lj) CU(DFREE(a->d_lj));
            if (a->d_ld) CU(DFREE(a->d_ld));
            a->d_lj = nullptr;
            a->d_ld = nullptr;
            a->list_cap = 0;
            CU(DMALLOC(&a->d_lj, (size_t)total * 4));
            CU(DMALLOC(&a->d_ld, (size_t)total * 8));
            a->list_cap = total;
        }
        const unsigned gr = (unsigned)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
        multi_prep_kernel<<<(unsigned)((std::max<int64_t>(q1 - q0, MULTI_STATE_EXTRA) + 255) / 256), 256, 0, c->stream>>>(
            a->x_recv, G, seg, segw, rpp, q0, q1, a->sp.inval, a->sp.flags, a->sp.cnt, a->d_cursor, a->x_state);
        count_rows_hseg_kernel<<<gr, 256, 0, c->stream>>>(a->x_recv, G, seg, segw, a->sp.cnt);
        // row pointers of the own rows only: d_ptr[q0 .. q1]
        if (q1 - q0 <= (1 << 17)) {
            scan_small_kernel<<<1, 1024, 0, c->stream>>>(a->sp.cnt + q0, q1 - q0, a->d_ptr + q0);
            c->launches++;
        } else {
            const int64_t m = q1 - q0, nb = (m + SCAN_BLOCK - 1) / SCAN_BLOCK;
            if (!a->d_scan) CU(DMALLOC(&a->d_scan, (size_t)((a->n + SCAN_BLOCK - 1) / SCAN_BLOCK + 1) * 8));
            scan_block_sums_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(a->sp.cnt + q0, m, a->d_scan);
            scan_block_offsets_kernel<<<1, 1024, 0, c->stream>>>(a->d_scan, nb, a->d_ptr + q0 + m);
            scan_apply_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(a->sp.cnt + q0, m, a->d_scan, a->d_ptr + q0);
            c->launches += 3;
        }
        scatter_hseg_kernel<<<gr, 256, 0, c->stream>>>(a->x_recv, G, seg, segw, a->d_ptr, a->d_cursor, a->d_lj, a->d_ld);
        const long long guess = a->last_ncand > 0 ? a->last_ncand / std::max<int64_t>(n, 1) * (q1 - q0) : 15 * (long long)(q1 - q0);
        launch_finalize(a, a->sp.inval, a->sp.flags, q0, q1, guess);
        c->launches += 4;
        CU(cudaGetLastError());
        TRY(fill_shared(a, q0, q1));
        if (q1 - q0 < rpp) {
            zero_rows_kernel<<<64, 256, 0, c->stream>>>(a->d_res, q1, (int64_t)rank * rpp + rpp);
            c->launches++;
        }
        MARK();  // 5: own rows finalised
        // in place: this rank's slice already sits at its offset of the gathered buffer
        NC(nc->AllGather(a->d_res + (int64_t)rank * rpp, a->d_res, (size_t)rpp * sizeof(tdna_row_result), ncclInt8, c->comm.comm, c->stream));
        MARK();  // 6: rows gathered
    } else {  // a pass without Best Match rows exchanges nothing but its status
        multi_zero_kernel<<<1, 64, 0, c->stream>>>(a->x_send, 0, segw, a->x_state);
        multi_pack_kernel<<<1, 256, 0, c->stream>>>(a->sp, n, rpp, G, seg, segw, nullptr, a->x_state);
        c->launches += 2;
        NC(nc->AllReduce(a->x_state, a->x_state, (size_t)MULTI_STATE_EXTRA, ncclInt64, ncclSum, c->comm.comm, c->stream));
    }
#undef MARK
    // the one host synchronisation of the step: the status words summed over the ranks and this rank's own counters
    CU(cudaMemcpyAsync(c->h_status, a->x_state, MULTI_STATE_EXTRA * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_status + 16, a->d_stream_u64, sizeof(StreamCounters), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < MULTI_STATE_EXTRA; k++) status[k] = c->h_status[k];
    StreamCounters h;
    memcpy(&h, c->h_status + 16, sizeof(h));
    if (h.overflow) { a->need_cand = (long long)h.ncand; a->need_queue = (long long)h.qn; }
    a->local_ncls[0] = (long long)h.ncls[0];
    a->local_ncls[1] = (long long)h.ncls[1];
    a->local_nedge = (long long)h.nedge;
    a->last_ncls[0] = status[2];
    a->last_ncls[1] = status[3];
    a->last_nedge = status[4];
    a->last_ncand = status[5];
    a->last_queue = (long long)h.qn;
    return TDNA_OK;
}

static int best_match_dense(tdna_aln *a);
// TDNA_E_TOO_LARGE (on every rank alike): the streaming pass cannot serve this list; the caller runs the dense row-band fallback
static int multi_pass(tdna_aln *a, bool want_bm) {
    int64_t status[MULTI_STATE_EXTRA];
    for (int attempt = 0; attempt < 4; attempt++) {
        TRY(multi_attempt(a, want_bm, status));
        if (status[0] == 0 && status[1] == 0) return TDNA_OK;
        // some rank overflowed: every rank sees the same two words and retries together
        if (status[0]) a->x_scale *= 2;                                      // identical on every rank
        if (a->need_cand || a->need_queue) {                                   // this rank's own buffers
            if (!grow_stream_buffers(a) && a->last_kernel == TDNA_KERNEL_TENSOR && a->ctx->opt_count_kernel == TDNA_KERNEL_AUTO) a->avoid_tensor = true;
        }
    }
    return fail(TDNA_E_TOO_LARGE, "streaming Best Match: a candidate buffer or exchange segment overflowed on some rank");
}

// dense fallback on several GPUs: every rank sweeps its own band of query rows, the row records are all-gathered
static int multi_dense_rows(tdna_aln *a) {
    tdna_ctx *c = a->ctx;
    const NcclApi *nc = c->comm.api;
    const int G = c->comm.nranks, rank = c->comm.rank;
    const int64_t n = a->n, rpp = tdna_rows_per_part(n, G);
    const int64_t q0 = std::min<int64_t>(n, (int64_t)rank * rpp), q1 = std::min<int64_t>(n, q0 + rpp);
    TRY(ensure_res(a, (int64_t)G * rpp));
    const int64_t rb = a->row_begin, re = a->row_end;
    a->row_begin = q0;
    a->row_end = q1;
    a->mat_valid = false;
    int rc = q1 > q0 ? best_match_dense(a) : TDNA_OK;
    a->row_begin = rb;
    a->row_end = re;
    a->mat_valid = false;
    TRY(rc);
    if (q1 - q0 < rpp) {
        zero_rows_kernel<<<64, 256, 0, c->stream>>>(a->d_res, q1, (int64_t)rank * rpp + rpp);
        c->launches++;
    }
    NC(nc->AllGather(a->d_res + (int64_t)rank * rpp, a->d_res, (size_t)rpp * sizeof(tdna_row_result), ncclInt8, c->comm.comm, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

int32_t tdna_best_match_seed(tdna_aln *a, int32_t part, int32_t nparts, uint64_t **minima_dev) {
    if (!a || nparts < 1 || part < 0 || part >= nparts) return fail(TDNA_E_ARG, "bad argument");
    CU(cudaSetDevice(a->ctx->device));
    TRY(need_labels(a));
    TRY(stream_seed(a, part, nparts));
    CU(cudaStreamSynchronize(a->ctx->stream));
    if (minima_dev) *minima_dev = reinterpret_cast<uint64_t *>(a->sp.mC);
    return TDNA_OK;
}

int32_t tdna_best_match_part(tdna_aln *a, int32_t part, int32_t nparts, int32_t seeded, tdna_stream_part *out) {
    if (!a || !out || nparts < 1 || part < 0 || part >= nparts) return fail(TDNA_E_ARG, "bad argument");
    CU(cudaSetDevice(a->ctx->device));
    TRY(need_labels(a));
    if (seeded && !a->stream_alloc) return fail(TDNA_E_STATE, "tdna_best_match_seed has not run");
    long long ncand = 0;
    if (seeded) TRY(stream_bulk(a, part, nparts, &ncand));
    else TRY(stream_part(a, part, nparts, &ncand));
    out->cand_row = a->sp.cq;
    out->cand_col = a->sp.cj;
    out->cand_d = a->sp.cd;
    out->n_cand = ncand;
    out->invalid_count = a->sp.inval;
    out->flags = a->sp.flags;
    return TDNA_OK;
}

int32_t tdna_best_match_part_packed(tdna_aln *a, int32_t part, int32_t nparts, int32_t seeded, int64_t *rec, int64_t seg_stride, int64_t *count,
                                    int64_t *state) {
    if (!a || !rec || !count || !state || seg_stride < 1 || nparts < 1 || part < 0 || part >= nparts) return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    if (seeded && !a->stream_alloc) return fail(TDNA_E_STATE, "tdna_best_match_seed has not run");
    long long ncand = 0;
    if (seeded) TRY(stream_bulk(a, part, nparts, &ncand));
    else TRY(stream_part(a, part, nparts, &ncand));
    if (ncand > seg_stride) return fail(TDNA_E_TOO_LARGE, "the part's candidates do not fit the exchange segment");
    unsigned gr = (unsigned)((ncand + 255) / 256);
    if (gr < 1) gr = 1;
    if (gr > 65535u * 4) gr = 65535u * 4;
    pack_records_kernel<<<gr, 256, 0, c->stream>>>(a->sp.cq, a->sp.cj, a->sp.cd, ncand, reinterpret_cast<long long *>(rec));
    pack_state_kernel<<<(unsigned)((a->n + 255) / 256), 256, 0, c->stream>>>(a->sp.inval, a->sp.flags, a->n, ncand, reinterpret_cast<long long *>(state),
                                                                            reinterpret_cast<long long *>(count));
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));  // the caller hands the buffers to NCCL on another stream
    return TDNA_OK;
}

int32_t tdna_best_match_merge_gathered(tdna_aln *a, const int64_t *rec, int32_t nseg, int64_t seg_stride, const int64_t *seg_count, const int64_t *state,
                                       tdna_row_result *out) {
    if (!a || !rec || !seg_count || !state || nseg < 1 || seg_stride < 1) return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    TRY(ensure_packed(a));
    TRY(ensure_stream_state(a));
    TRY(ensure_res(a, a->n));
    size_t n = (size_t)a->n;
    const long long total = (long long)nseg * seg_stride;  // upper bound of the records: sizes the lists without a host round trip
    if (total > a->list_cap) {
        if (a->d_lj) CU(DFREE(a->d_lj));
        if (a->d_ld) CU(DFREE(a->d_ld));
        a->d_lj = nullptr;
        a->d_ld = nullptr;
        CU(DMALLOC(&a->d_lj, (size_t)total * 4));
        CU(DMALLOC(&a->d_ld, (size_t)total * 8));
        a->list_cap = total;
    }
    unsigned gr = (unsigned)((total + 255) / 256);
    if (gr > 65535u * 4) gr = 65535u * 4;
    const long long *lrec = reinterpret_cast<const long long *>(rec), *lcnt = reinterpret_cast<const long long *>(seg_count);
    unpack_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const long long *>(state), a->n, a->sp.inval, a->sp.flags);
    CU(cudaMemsetAsync(a->sp.cnt, 0, n * 4, c->stream));
    CU(cudaMemsetAsync(a->d_cursor, 0, n * 4, c->stream));
    count_rows_seg_kernel<<<gr, 256, 0, c->stream>>>(lrec, nseg, seg_stride, lcnt, a->sp.cnt);
    TRY(row_pointers(a, a->sp.cnt));
    scatter_seg_kernel<<<gr, 256, 0, c->stream>>>(lrec, nseg, seg_stride, lcnt, a->d_ptr, a->d_cursor, a->d_lj, a->d_ld);
    launch_finalize(a, a->sp.inval, a->sp.flags, 0, a->n, total / 2);  // the segments are sized with about 2x headroom
    c->launches += 4;
    CU(cudaGetLastError());
    TRY(fill_shared(a, 0, a->n));
    if (out) CU(cudaMemcpyAsync(out, a->d_res, n * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

// ---- routed exchange: records travel to the part that owns their row; every part finalises its own rows only ----------------
int64_t tdna_rows_per_part(int64_t n, int32_t nparts) {
    if (n < 1 || nparts < 1) return 0;
    return round_up((n + nparts - 1) / nparts, 128);
}

int32_t tdna_best_match_part_routed(tdna_aln *a, int32_t part, int32_t nparts, int32_t seeded, int64_t *rec, int64_t seg_stride, int64_t *counts,
                                    int64_t *state) {
    if (!a || !rec || !counts || !state || seg_stride < 1 || nparts < 1 || nparts > ROUTE_MAX_PARTS || part < 0 || part >= nparts)
        return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    if (seeded && !a->stream_alloc) return fail(TDNA_E_STATE, "tdna_best_match_seed has not run");
    long long ncand = 0;
    if (seeded) TRY(stream_bulk(a, part, nparts, &ncand));
    else TRY(stream_part(a, part, nparts, &ncand));
    int *d_over = reinterpret_cast<int *>(c->d_counter) + 4;
    CU(cudaMemsetAsync(d_over, 0, 4, c->stream));
    CU(cudaMemsetAsync(counts, 0, (size_t)nparts * 8, c->stream));
    unsigned gr = (unsigned)((ncand + 255) / 256);
    if (gr < 1) gr = 1;
    if (gr > (unsigned)c->sm_count * 16) gr = (unsigned)c->sm_count * 16;
    route_records_kernel<<<gr, 256, 0, c->stream>>>(a->sp.cq, a->sp.cj, a->sp.cd, ncand, tdna_rows_per_part(a->n, nparts), nparts, seg_stride,
                                                    reinterpret_cast<long long *>(rec), reinterpret_cast<unsigned long long *>(counts), d_over);
    pack_state_only_kernel<<<(unsigned)((a->n + 255) / 256), 256, 0, c->stream>>>(a->sp.inval, a->sp.flags, a->n, reinterpret_cast<long long *>(state));
    c->launches += 2;
    CU(cudaGetLastError());
    int over = 0;
    CU(cudaMemcpyAsync(&over, d_over, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // the caller hands the buffers to NCCL on another stream
    if (over) return fail(TDNA_E_TOO_LARGE, "a destination's candidates do not fit its exchange segment");
    return TDNA_OK;
}

int32_t tdna_best_match_merge_routed(tdna_aln *a, int32_t part, int32_t nparts, const int64_t *rec, int64_t seg_stride, const int64_t *seg_count,
                                     const int64_t *state, tdna_row_result *out_slice) {
    if (!a || !rec || !seg_count || !state || seg_stride < 1 || nparts < 1 || nparts > ROUTE_MAX_PARTS || part < 0 || part >= nparts)
        return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    TRY(ensure_packed(a));
    TRY(ensure_stream_state(a));
    TRY(ensure_res(a, a->n));
    const size_t n = (size_t)a->n;
    const int64_t rpp = tdna_rows_per_part(a->n, nparts);
    const int64_t q0 = std::min<int64_t>(a->n, (int64_t)part * rpp), q1 = std::min<int64_t>(a->n, q0 + rpp);
    const long long total = (long long)nparts * seg_stride;  // upper bound of the records received
    if (total > a->list_cap) {
        if (a->d_lj) CU(DFREE(a->d_lj));
        if (a->d_ld) CU(DFREE(a->d_ld));
        a->d_lj = nullptr;
        a->d_ld = nullptr;
        CU(DMALLOC(&a->d_lj, (size_t)total * 4));
        CU(DMALLOC(&a->d_ld, (size_t)total * 8));
        a->list_cap = total;
    }
    unsigned gr = (unsigned)((total + 255) / 256);
    if (gr > 65535u * 4) gr = 65535u * 4;
    const long long *lrec = reinterpret_cast<const long long *>(rec), *lcnt = reinterpret_cast<const long long *>(seg_count);
    unpack_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const long long *>(state), a->n, a->sp.inval, a->sp.flags);
    CU(cudaMemsetAsync(a->sp.cnt, 0, n * 4, c->stream));
    CU(cudaMemsetAsync(a->d_cursor, 0, n * 4, c->stream));
    count_rows_seg_kernel<<<gr, 256, 0, c->stream>>>(lrec, nparts, seg_stride, lcnt, a->sp.cnt);  // every record received is one of this part's rows
    TRY(row_pointers(a, a->sp.cnt));
    scatter_seg_kernel<<<gr, 256, 0, c->stream>>>(lrec, nparts, seg_stride, lcnt, a->d_ptr, a->d_cursor, a->d_lj, a->d_ld);
    launch_finalize(a, a->sp.inval, a->sp.flags, q0, q1, total / 3);  // the segments are sized with about 3x headroom
    c->launches += 4;
    CU(cudaGetLastError());
    TRY(fill_shared(a, q0, q1));
    if (out_slice) {  // rows_per_part records: this part's rows, zero padded past the last row of the list
        if (q1 > q0) CU(cudaMemcpyAsync(out_slice, a->d_res + q0, (size_t)(q1 - q0) * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
        if (q1 - q0 < rpp) {
            if (is_device_ptr(out_slice)) CU(cudaMemsetAsync(out_slice + (q1 - q0), 0, (size_t)(rpp - (q1 - q0)) * sizeof(tdna_row_result), c->stream));
            else { CU(cudaStreamSynchronize(c->stream)); memset(out_slice + (q1 - q0), 0, (size_t)(rpp - (q1 - q0)) * sizeof(tdna_row_result)); }
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

int32_t tdna_best_match_merge(tdna_aln *a, const int32_t *cand_row, const int32_t *cand_col, const double *cand_d, int64_t n_cand,
                              const int32_t *invalid_count, const int32_t *flags, tdna_row_result *out) {
    if (!a || !invalid_count || !flags || n_cand < 0 || (n_cand > 0 && (!cand_row || !cand_col || !cand_d)))
        return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    TRY(ensure_packed(a));
    TRY(stream_merge(a, cand_row, cand_col, cand_d, n_cand, invalid_count, flags, false));
    if (out) CU(cudaMemcpyAsync(out, a->d_res, (size_t)a->n * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

static int class_distances_dense(tdna_aln *a, int32_t klass, float *out, int64_t cap, int64_t *n_out, bool hist = false);
static int edges_dense(tdna_aln *a, double t, int32_t *ii, int32_t *jj, double *d, int64_t cap, int64_t *n_out, double lo = 0.0);

// ---- class histograms ------------------------------------------------------------------------------------------------------------
#define TDNA_E_HIST_FULL 100  // internal: a histogram table filled up -- grow it and repeat the pass

static int hist_prepare(tdna_aln *a, int k) {
    tdna_ctx *c = a->ctx;
    if (!a->d_hctr) CU(DMALLOC(&a->d_hctr, 64));
    if (a->hist_alloc[k] != a->hist_size) {
        if (a->d_hkey[k]) CU(DFREE(a->d_hkey[k]));
        if (a->d_hcnt[k]) CU(DFREE(a->d_hcnt[k]));
        a->d_hkey[k] = a->d_hcnt[k] = nullptr;
        a->hist_alloc[k] = 0;
        CU(DMALLOC(&a->d_hkey[k], (size_t)a->hist_size * 8));
        CU(DMALLOC(&a->d_hcnt[k], (size_t)a->hist_size * 8));
        a->hist_alloc[k] = a->hist_size;
    }
    CU(cudaMemsetAsync(a->d_hkey[k], 0xff, (size_t)a->hist_size * 8, c->stream));
    CU(cudaMemsetAsync(a->d_hcnt[k], 0, (size_t)a->hist_size * 8, c->stream));
    CU(cudaMemsetAsync(a->d_hctr, 0, 64, c->stream));
    return TDNA_OK;
}

static int hist_staging(tdna_aln *a, size_t entries) {
    tdna_ctx *c = a->ctx;
    if (entries <= a->hent_cap) return TDNA_OK;
    if (a->d_hent) CU(DFREE(a->d_hent));
    a->d_hent = nullptr;
    a->hent_cap = 0;
    CU(DMALLOC(&a->d_hent, entries * sizeof(tdna_hist_entry)));
    a->hent_cap = entries;
    return TDNA_OK;
}

// table k -> (out, cap); with a communicator the tables of all ranks are merged into this rank's first (every rank ends up with
// the complete histogram).  TDNA_E_HIST_FULL: some table filled up (the same verdict on every rank).
static int hist_finish(tdna_aln *a, int k, tdna_hist_entry *out, int64_t cap, int64_t *n_out, int64_t *pushes_out) {
    tdna_ctx *c = a->ctx;
    const bool multi = c->comm.active();
    const unsigned grid = (unsigned)c->sm_count * 4;
    struct { int over; int pad; unsigned long long n, pushes; } h;
    if (multi) {
        const NcclApi *nc = c->comm.api;
        const int G = c->comm.nranks;
        // this rank's entries, their number on every rank, the entries of every rank; the others' go into this rank's table
        TRY(hist_staging(a, (size_t)a->hist_size * (size_t)(G + 1)));
        tdna_hist_entry *mine = a->d_hent + (size_t)a->hist_size * G;
        CU(cudaMemsetAsync(a->d_hctr + 1, 0, 16, c->stream));
        hist_compact_kernel<<<grid, 256, 0, c->stream>>>(a->d_hkey[k], a->d_hcnt[k], a->hist_size, mine, (long long)a->hist_size, a->d_hctr + 1, a->d_hctr + 2);
        c->launches++;
        // the full flag and the entry count of every rank, all-gathered
        long long *d_pairs = nullptr;
        CU(DMALLOC(&d_pairs, (size_t)G * 16));
        hist_meta_kernel<<<1, 32, 0, c->stream>>>(reinterpret_cast<const int *>(a->d_hctr), a->d_hctr + 1, d_pairs + 2 * c->comm.rank);
        c->launches++;
        NC(nc->AllGather(d_pairs + 2 * c->comm.rank, d_pairs, 16, ncclInt8, c->comm.comm, c->stream));
        std::vector<long long> meta((size_t)2 * G);
        CU(cudaMemcpyAsync(meta.data(), d_pairs, (size_t)G * 16, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(DFREE(d_pairs));
        long long over = 0, maxn = 0;
        for (int g = 0; g < G; g++) { over += meta[2 * g]; maxn = std::max(maxn, meta[2 * g + 1]); }
        if (over) return TDNA_E_HIST_FULL;
        if (maxn > 0) {
            NC(nc->AllGather(mine, a->d_hent, (size_t)maxn * sizeof(tdna_hist_entry), ncclInt8, c->comm.comm, c->stream));
            for (int g = 0; g < G; g++) {
                if (g == c->comm.rank || meta[2 * g + 1] == 0) continue;
                hist_insert_kernel<<<grid, 256, 0, c->stream>>>(a->d_hent + (size_t)g * maxn, meta[2 * g + 1], a->d_hkey[k], a->d_hcnt[k], a->hist_size - 1,
                                                               reinterpret_cast<int *>(a->d_hctr));
                c->launches++;
            }
        }
    }
    tdna_hist_entry *dout = nullptr;
    const bool direct = out && cap > 0 && is_device_ptr(out);
    if (out && cap > 0) {
        if (direct) dout = out;
        else { TRY(hist_staging(a, (size_t)std::min<int64_t>(cap, a->hist_size))); dout = a->d_hent; }
    }
    const long long dcap = out && cap > 0 ? std::min<int64_t>(cap, a->hist_size) : 0;
    CU(cudaMemsetAsync(a->d_hctr + 1, 0, 16, c->stream));
    hist_compact_kernel<<<grid, 256, 0, c->stream>>>(a->d_hkey[k], a->d_hcnt[k], a->hist_size, dout, dcap, a->d_hctr + 1, a->d_hctr + 2);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&h, a->d_hctr, 24, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h.over) return TDNA_E_HIST_FULL;  // (single GPU; after a multi-GPU merge: the union did not fit -- the same on every rank)
    if (dout && !direct) {
        const long long m = std::min<long long>((long long)h.n, dcap);
        if (m > 0) CU(cudaMemcpyAsync(out, dout, (size_t)m * sizeof(tdna_hist_entry), cudaMemcpyDefault, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    *n_out = (int64_t)h.n;
    *pushes_out = (int64_t)h.pushes;
    return TDNA_OK;
}

// One streaming pass for any subset of {Best Match, intra, inter, edges}.  TDNA_E_TOO_LARGE: candidate overflow.
static int analyze_stream(tdna_aln *a, tdna_analysis *io) {
    tdna_ctx *c = a->ctx;
    TRY(ensure_packed(a));
    TRY(ensure_stream_state(a));
    StreamParams &sp = a->sp;
    // histogram requests are class requests whose pushes are counted per distinct distance instead of listed
    const bool hist[2] = {(io->want & TDNA_WANT_HIST_INTRA) != 0, (io->want & TDNA_WANT_HIST_INTER) != 0};
    const int want = (io->want & (TDNA_WANT_BEST_MATCH | TDNA_WANT_INTRA | TDNA_WANT_INTER | TDNA_WANT_EDGES)) | (hist[0] ? TDNA_WANT_INTRA : 0) |
                     (hist[1] ? TDNA_WANT_INTER : 0);
    sp.want = want;
    float *dcls[2] = {nullptr, nullptr};
    float *hcls[2] = {io->intra, io->inter};
    int64_t ccap[2] = {io->intra_cap, io->inter_cap};
    bool own_cls[2] = {false, false};
    int32_t *dei = nullptr, *dej = nullptr;
    double *ded = nullptr;
    bool own_edges = false;
    int rc = TDNA_OK;
    auto cleanup = [&]() {
        for (int k = 0; k < 2; k++)
            if (own_cls[k] && dcls[k]) DFREE(dcls[k]);
        if (own_edges) { if (dei) DFREE(dei); if (dej) DFREE(dej); if (ded) DFREE(ded); }
        sp.want = TDNA_WANT_BEST_MATCH;
        sp.cls[0] = sp.cls[1] = nullptr;
        sp.hkey[0] = sp.hkey[1] = sp.hcnt[0] = sp.hcnt[1] = nullptr;
        sp.ei = sp.ej = nullptr;
        sp.ed = nullptr;
    };
    sp.hkey[0] = sp.hkey[1] = sp.hcnt[0] = sp.hcnt[1] = nullptr;
    for (int k = 0; k < 2; k++) {
        sp.cls[k] = nullptr;
        sp.cls_cap[k] = 0;
        if (hist[k] || !(io->want & (k == 0 ? TDNA_WANT_INTRA : TDNA_WANT_INTER)) || !hcls[k] || ccap[k] <= 0) continue;
        if (is_device_ptr(hcls[k])) dcls[k] = hcls[k];
        else {
            cudaError_t e = DMALLOC(&dcls[k], (size_t)ccap[k] * 4);
            if (e != cudaSuccess) { cleanup(); return fail(TDNA_E_CUDA, cudaGetErrorString(e)); }
            own_cls[k] = true;
        }
        sp.cls[k] = dcls[k];
        sp.cls_cap[k] = ccap[k];
    }
    sp.ei = sp.ej = nullptr;
    sp.ed = nullptr;
    sp.edge_cap = 0;
    sp.edge_t = io->edge_threshold;
    sp.edge_tfix = 0;
    if (io->want & TDNA_WANT_EDGES) {
        double t = io->edge_threshold;
        sp.edge_tfix = !(t >= 0) ? 0u : (t >= 1.0 ? TDNA_THR_ALL : (uint32_t)(t * 65536.0) + 2u);
        if (io->edge_i && io->edge_j && io->edge_d && io->edge_cap > 0) {
            if (is_device_ptr(io->edge_i) && is_device_ptr(io->edge_j) && is_device_ptr(io->edge_d)) {
                dei = io->edge_i; dej = io->edge_j; ded = io->edge_d;
            } else {
                own_edges = true;
                cudaError_t e = DMALLOC(&dei, (size_t)io->edge_cap * 4);
                if (e == cudaSuccess) e = DMALLOC(&dej, (size_t)io->edge_cap * 4);
                if (e == cudaSuccess) e = DMALLOC(&ded, (size_t)io->edge_cap * 8);
                if (e != cudaSuccess) { cleanup(); return fail(TDNA_E_CUDA, cudaGetErrorString(e)); }
            }
            sp.ei = dei; sp.ej = dej; sp.ed = ded;
            sp.edge_cap = io->edge_cap;
        }
    }
    long long ncand = 0;
    const bool multi = c->comm.active();  // the GPUs of the communicator share the pass (every rank makes this same call)
    for (int hattempt = 0;; hattempt++) {
        for (int k = 0; k < 2 && rc == TDNA_OK; k++)
            if (hist[k]) {
                rc = hist_prepare(a, k);
                sp.hkey[k] = a->d_hkey[k];
                sp.hcnt[k] = a->d_hcnt[k];
                sp.hmask = a->hist_size - 1;
                sp.hover = reinterpret_cast<int *>(a->d_hctr);
            }
        if (rc != TDNA_OK) break;
        if (multi) rc = multi_pass(a, (want & TDNA_WANT_BEST_MATCH) != 0);
        else {
            rc = stream_part(a, 0, 1, &ncand);
            for (int attempt = 0; rc == TDNA_E_TOO_LARGE && attempt < 3 && grow_stream_buffers(a); attempt++) rc = stream_part(a, 0, 1, &ncand);
        }
        if (rc != TDNA_OK) break;
        if (hist[0]) rc = hist_finish(a, 0, io->hist_intra, io->hist_intra_cap, &io->n_hist_intra, &io->hist_intra_pushes);
        if (rc == TDNA_OK && hist[1]) rc = hist_finish(a, 1, io->hist_inter, io->hist_inter_cap, &io->n_hist_inter, &io->hist_inter_pushes);
        if (rc != TDNA_E_HIST_FULL) break;
        if (hattempt >= 3 || a->hist_size >= (1u << 26)) { rc = fail(TDNA_E_TOO_LARGE, "a class histogram has more than 2^26 distinct distances"); break; }
        a->hist_size *= 4;  // (every rank saw the same verdict: the tables grow alike)
        rc = TDNA_OK;
    }
    if (rc == TDNA_OK && (io->want & TDNA_WANT_BEST_MATCH)) {
        if (!multi) {
            rc = ensure_res(a, a->n);
            if (rc == TDNA_OK) rc = stream_merge(a, sp.cq, sp.cj, sp.cd, ncand, sp.inval, sp.flags, true);
        }
        if (rc == TDNA_OK && io->rows) {
            cudaError_t e = cudaMemcpyAsync(io->rows, a->d_res, (size_t)a->n * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream);
            if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
        }
    }
    if (rc == TDNA_OK) {
        io->n_intra = (io->want & TDNA_WANT_INTRA) ? a->last_ncls[0] : 0;
        io->n_inter = (io->want & TDNA_WANT_INTER) ? a->last_ncls[1] : 0;
        io->n_edges = (io->want & TDNA_WANT_EDGES) ? a->last_nedge : 0;
        io->n_intra_local = (io->want & TDNA_WANT_INTRA) ? a->local_ncls[0] : 0;
        io->n_inter_local = (io->want & TDNA_WANT_INTER) ? a->local_ncls[1] : 0;
        io->n_edges_local = (io->want & TDNA_WANT_EDGES) ? a->local_nedge : 0;
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 2 && e == cudaSuccess; k++)
            if (own_cls[k]) {
                int64_t m = a->local_ncls[k] < ccap[k] ? a->local_ncls[k] : ccap[k];
                if (m > 0) e = cudaMemcpyAsync(hcls[k], dcls[k], (size_t)m * 4, cudaMemcpyDefault, c->stream);
            }
        if (e == cudaSuccess && own_edges) {
            int64_t m = a->local_nedge < io->edge_cap ? a->local_nedge : io->edge_cap;
            if (m > 0) {
                e = cudaMemcpyAsync(io->edge_i, dei, (size_t)m * 4, cudaMemcpyDefault, c->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(io->edge_j, dej, (size_t)m * 4, cudaMemcpyDefault, c->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(io->edge_d, ded, (size_t)m * 8, cudaMemcpyDefault, c->stream);
            }
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

// class histograms from the materialised matrix (the rank's own row band when a communicator shares the request)
static int hist_dense(tdna_aln *a, tdna_analysis *io) {
    const bool hist[2] = {(io->want & TDNA_WANT_HIST_INTRA) != 0, (io->want & TDNA_WANT_HIST_INTER) != 0};
    if (!hist[0] && !hist[1]) return TDNA_OK;
    int rc = TDNA_OK;
    for (int hattempt = 0;; hattempt++) {
        int64_t dummy = 0;
        for (int k = 0; k < 2 && rc == TDNA_OK; k++)
            if (hist[k]) {
                rc = hist_prepare(a, k);
                if (rc == TDNA_OK && a->row_end > a->row_begin) rc = class_distances_dense(a, k, nullptr, 0, &dummy, true);
            }
        if (rc != TDNA_OK) return rc;
        if (hist[0]) rc = hist_finish(a, 0, io->hist_intra, io->hist_intra_cap, &io->n_hist_intra, &io->hist_intra_pushes);
        if (rc == TDNA_OK && hist[1]) rc = hist_finish(a, 1, io->hist_inter, io->hist_inter_cap, &io->n_hist_inter, &io->hist_inter_pushes);
        if (rc != TDNA_E_HIST_FULL) return rc;
        if (hattempt >= 3 || a->hist_size >= (1u << 26)) return fail(TDNA_E_TOO_LARGE, "a class histogram has more than 2^26 distinct distances");
        a->hist_size *= 4;
        rc = TDNA_OK;
    }
}

// several GPUs, dense fallback: Best Match rows from every rank's own band of query rows (all-gathered); class distances and
// edges of the rank's own band with the totals all-reduced.  Every rank takes this path together (multi_pass's status is shared).
static int analyze_dense_multi(tdna_aln *a, tdna_analysis *io) {
    tdna_ctx *c = a->ctx;
    const NcclApi *nc = c->comm.api;
    const int G = c->comm.nranks, rank = c->comm.rank;
    if (io->want & TDNA_WANT_BEST_MATCH) {
        TRY(multi_dense_rows(a));
        if (io->rows) CU(cudaMemcpyAsync(io->rows, a->d_res, (size_t)a->n * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    const int64_t rpp = tdna_rows_per_part(a->n, G);
    const int64_t q0 = std::min<int64_t>(a->n, (int64_t)rank * rpp), q1 = std::min<int64_t>(a->n, q0 + rpp);
    const int64_t rb = a->row_begin, re = a->row_end;
    a->row_begin = q0; a->row_end = q1; a->mat_valid = false;
    int64_t loc[3] = {0, 0, 0};
    int rc = TDNA_OK;
    if (q1 > q0) {
        if (rc == TDNA_OK && (io->want & TDNA_WANT_INTRA)) rc = class_distances_dense(a, TDNA_PD_INTRA, io->intra, io->intra_cap, &loc[0]);
        if (rc == TDNA_OK && (io->want & TDNA_WANT_INTER)) rc = class_distances_dense(a, TDNA_PD_INTER, io->inter, io->inter_cap, &loc[1]);
        if (rc == TDNA_OK && (io->want & TDNA_WANT_EDGES)) rc = edges_dense(a, io->edge_threshold, io->edge_i, io->edge_j, io->edge_d, io->edge_cap, &loc[2]);
    }
    if (rc == TDNA_OK) rc = hist_dense(a, io);  // (merged over the ranks inside hist_finish)
    a->row_begin = rb; a->row_end = re; a->mat_valid = false;
    TRY(rc);
    io->n_intra_local = loc[0]; io->n_inter_local = loc[1]; io->n_edges_local = loc[2];
    long long *d_tot = reinterpret_cast<long long *>(c->d_counter);
    CU(cudaMemcpyAsync(d_tot, loc, 24, cudaMemcpyHostToDevice, c->stream));
    NC(nc->AllReduce(d_tot, d_tot, 3, ncclInt64, ncclSum, c->comm.comm, c->stream));
    int64_t tot[3];
    CU(cudaMemcpyAsync(tot, d_tot, 24, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    io->n_intra = tot[0]; io->n_inter = tot[1]; io->n_edges = tot[2];
    return TDNA_OK;
}

static int analyze_dense(tdna_aln *a, tdna_analysis *io) {
    tdna_ctx *c = a->ctx;
    if (io->want & TDNA_WANT_BEST_MATCH) {
        TRY(ensure_res(a, a->n));
        TRY(best_match_dense(a));
        int64_t rows = a->row_end - a->row_begin;
        if (io->rows && rows > 0)
            CU(cudaMemcpyAsync(io->rows + a->row_begin, a->d_res + a->row_begin, (size_t)rows * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    io->n_intra = io->n_inter = io->n_edges = 0;
    if (io->want & TDNA_WANT_INTRA) TRY(class_distances_dense(a, TDNA_PD_INTRA, io->intra, io->intra_cap, &io->n_intra));
    if (io->want & TDNA_WANT_INTER) TRY(class_distances_dense(a, TDNA_PD_INTER, io->inter, io->inter_cap, &io->n_inter));
    if (io->want & TDNA_WANT_EDGES) TRY(edges_dense(a, io->edge_threshold, io->edge_i, io->edge_j, io->edge_d, io->edge_cap, &io->n_edges));
    io->n_intra_local = io->n_intra; io->n_inter_local = io->n_inter; io->n_edges_local = io->n_edges;
    TRY(hist_dense(a, io));
    return TDNA_OK;
}

int32_t tdna_analyze(tdna_aln *a, tdna_analysis *io) {
    if (!a || !io) return fail(TDNA_E_ARG, "null argument");
    if (io->want == 0 || (io->want & ~(TDNA_WANT_BEST_MATCH | TDNA_WANT_INTRA | TDNA_WANT_INTER | TDNA_WANT_EDGES | TDNA_WANT_HIST_INTRA | TDNA_WANT_HIST_INTER)))
        return fail(TDNA_E_ARG, "bad want mask");
    if (((io->want & TDNA_WANT_INTRA) && (io->want & TDNA_WANT_HIST_INTRA)) || ((io->want & TDNA_WANT_INTER) && (io->want & TDNA_WANT_HIST_INTER)))
        return fail(TDNA_E_ARG, "a class is either listed or histogrammed in one pass");
    io->n_hist_intra = io->n_hist_inter = io->hist_intra_pushes = io->hist_inter_pushes = 0;
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    if (io->want & (TDNA_WANT_BEST_MATCH | TDNA_WANT_INTRA | TDNA_WANT_INTER | TDNA_WANT_HIST_INTRA | TDNA_WANT_HIST_INTER)) TRY(need_labels(a));
    bool whole = a->row_begin == 0 && a->row_end == a->n;
    if (whole && c->comm.active()) {
        // the communicator's GPUs share the request; every decision below is the same on every rank (labels are part of the contract,
        // the streaming pass reports an all-reduced status)
        if (!a->has_labels) return fail(TDNA_E_STATE, "a multi-GPU pass needs the labels (tdna_alignment_set_labels)");
        if (c->opt_best_match_path != TDNA_PATH_DENSE) {
            int rc = analyze_stream(a, io);
            if (rc != TDNA_E_TOO_LARGE || c->opt_best_match_path == TDNA_PATH_STREAM) return rc;
        }
        return analyze_dense_multi(a, io);
    }
    // a materialised matrix of this very configuration is already resident (tdna_matrix_compute): sweeping it is cheaper than a pass
    bool have_matrix = a->mat_valid && c->mat_owner == a && !(io->want & TDNA_WANT_BEST_MATCH) && c->opt_best_match_path != TDNA_PATH_STREAM;
    if (whole && c->opt_best_match_path != TDNA_PATH_DENSE && a->has_labels && !have_matrix) {
        int rc = analyze_stream(a, io);
        if (rc == TDNA_E_TOO_LARGE && c->opt_count_kernel == TDNA_KERNEL_AUTO && a->last_kernel == TDNA_KERNEL_TENSOR) {
            // kernel (b) filters with the seeded thresholds only; kernel (a) tightens them while it runs: try it before the matrix
            a->avoid_tensor = true;
            rc = analyze_stream(a, io);
        }
        if (rc != TDNA_E_TOO_LARGE || c->opt_best_match_path == TDNA_PATH_STREAM) return rc;
    }
    return analyze_dense(a, io);
}

int32_t tdna_class_distances(tdna_aln *a, int32_t klass, float *out, int64_t cap, int64_t *n_out) {
    if (!a || !n_out) return fail(TDNA_E_ARG, "null argument");
    if (klass != TDNA_PD_INTRA && klass != TDNA_PD_INTER) return fail(TDNA_E_ARG, "bad class");
    tdna_analysis io = {};
    io.want = klass == TDNA_PD_INTRA ? TDNA_WANT_INTRA : TDNA_WANT_INTER;
    if (klass == TDNA_PD_INTRA) { io.intra = out; io.intra_cap = cap; } else { io.inter = out; io.inter_cap = cap; }
    TRY(tdna_analyze(a, &io));
    *n_out = klass == TDNA_PD_INTRA ? io.n_intra : io.n_inter;
    return TDNA_OK;
}

int32_t tdna_class_histogram(tdna_aln *a, int32_t klass, tdna_hist_entry *out, int64_t cap, int64_t *n_out, int64_t *pushes) {
    if (!a || !n_out) return fail(TDNA_E_ARG, "null argument");
    if (klass != TDNA_PD_INTRA && klass != TDNA_PD_INTER) return fail(TDNA_E_ARG, "bad class");
    tdna_analysis io = {};
    io.want = klass == TDNA_PD_INTRA ? TDNA_WANT_HIST_INTRA : TDNA_WANT_HIST_INTER;
    if (klass == TDNA_PD_INTRA) { io.hist_intra = out; io.hist_intra_cap = cap; } else { io.hist_inter = out; io.hist_inter_cap = cap; }
    TRY(tdna_analyze(a, &io));
    *n_out = klass == TDNA_PD_INTRA ? io.n_hist_intra : io.n_hist_inter;
    if (pushes) *pushes = klass == TDNA_PD_INTRA ? io.hist_intra_pushes : io.hist_inter_pushes;
    return TDNA_OK;
}

int32_t tdna_edges_below(tdna_aln *a, double t, int32_t *ii, int32_t *jj, double *d, int64_t cap, int64_t *n_out) {
    if (!a || !n_out) return fail(TDNA_E_ARG, "null argument");
    if (!a->has_labels) return edges_dense(a, t, ii, jj, d, cap, n_out);  // the streaming state needs the labels
    tdna_analysis io = {};
    io.want = TDNA_WANT_EDGES;
    io.edge_threshold = t;
    io.edge_i = ii; io.edge_j = jj; io.edge_d = d; io.edge_cap = cap;
    TRY(tdna_analyze(a, &io));
    *n_out = io.n_edges;
    return TDNA_OK;
}

int32_t tdna_pairs_between(tdna_aln *a, double lo, double hi, int32_t *ii, int32_t *jj, double *d, int64_t cap, int64_t *n_out) {
    if (!a || !n_out) return fail(TDNA_E_ARG, "null argument");
    return edges_dense(a, hi, ii, jj, d, cap, n_out, lo);  // row bands of the matrix (the lists these callers hold are small)
}

int32_t tdna_best_match_compute(tdna_aln *a, const tdna_row_result **out_dev) {
    if (!a) return fail(TDNA_E_ARG, "null alignment");
    tdna_analysis io = {};
    io.want = TDNA_WANT_BEST_MATCH;
    TRY(tdna_analyze(a, &io));
    if (out_dev) *out_dev = a->d_res;
    return TDNA_OK;
}

int32_t tdna_best_match(tdna_aln *a, tdna_row_result *out) {
    if (!out) return fail(TDNA_E_ARG, "out is null");
    TRY(tdna_best_match_compute(a, nullptr));
    tdna_ctx *c = a->ctx;
    int64_t rows = a->row_end - a->row_begin;
    if (rows > 0)
        CU(cudaMemcpyAsync(out + a->row_begin, a->d_res + a->row_begin, (size_t)rows * sizeof(tdna_row_result), cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

// hist: the pushes go into the class's histogram table (prepared by the caller) instead of a list
static int class_distances_dense(tdna_aln *a, int32_t klass, float *out, int64_t cap, int64_t *n_out, bool hist) {
    if (!a || !n_out) return fail(TDNA_E_ARG, "null argument");
    if (klass != TDNA_PD_INTRA && klass != TDNA_PD_INTER) return fail(TDNA_E_ARG, "bad class");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    CU(cudaMemsetAsync(c->d_counter, 0, 8, c->stream));
    unsigned long long *hk = hist ? a->d_hkey[klass] : nullptr, *hc = hist ? a->d_hcnt[klass] : nullptr;
    if (hist) { out = nullptr; cap = 0; }
    float *dout = nullptr;
    bool direct = false;
    if (out && cap > 0) {
        direct = is_device_ptr(out);  // a device destination is written by the kernel itself
        if (direct) dout = out;
        else {
            TRY(ensure_scratch(a, (size_t)cap * 4));
            dout = reinterpret_cast<float *>(a->d_scratch);
        }
    }
    int rc = for_each_band(a, [&](int64_t b0, int64_t b1) -> int {
        unsigned gy = (unsigned)((a->n + 256 * 8 - 1) / (256 * 8));
        if (gy > 65535) gy = 65535;
        dim3 grid((unsigned)(b1 - b0), gy);
        class_distances_kernel<<<grid, 256, 0, c->stream>>>(a->d_mat, a->ld, a->n, b0, b1 - b0, klass, a->d_species, a->d_genus, a->d_epi, a->d_full,
                                                           dout, cap, c->d_counter, hk, hc, a->hist_size - 1, reinterpret_cast<int *>(a->d_hctr));
        c->launches++;
        CU(cudaGetLastError());
        return TDNA_OK;
    });
    unsigned long long cnt = 0;
    if (rc == TDNA_OK) {
        cudaError_t e = cudaMemcpyAsync(&cnt, c->d_counter, 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess && dout && !direct) {
            int64_t m = (int64_t)cnt < cap ? (int64_t)cnt : cap;
            e = cudaMemcpyAsync(out, dout, (size_t)m * 4, cudaMemcpyDefault, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        }
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
    }
    *n_out = (int64_t)cnt;
    return rc;
}

static int edges_dense(tdna_aln *a, double t, int32_t *ii, int32_t *jj, double *d, int64_t cap, int64_t *n_out, double lo) {
    if (!a || !n_out) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->d_counter, 0, 8, c->stream));
    int32_t *di = nullptr, *dj = nullptr;
    double *dd = nullptr;
    bool want = ii && jj && d && cap > 0;
    if (want) {
        CU(DMALLOC(&di, (size_t)cap * 4));
        CU(DMALLOC(&dj, (size_t)cap * 4));
        CU(DMALLOC(&dd, (size_t)cap * 8));
    }
    int rc = for_each_band(a, [&](int64_t b0, int64_t b1) -> int {
        unsigned gy = (unsigned)((a->n + 256 * 8 - 1) / (256 * 8));
        if (gy > 65535) gy = 65535;
        dim3 grid((unsigned)(b1 - b0), gy);
        edges_kernel<<<grid, 256, 0, c->stream>>>(a->d_mat, a->ld, a->n, b0, b1 - b0, lo, t, di, dj, dd, cap, c->d_counter);
        c->launches++;
        CU(cudaGetLastError());
        return TDNA_OK;
    });
    unsigned long long cnt = 0;
    if (rc == TDNA_OK) {
        cudaError_t e = cudaMemcpyAsync(&cnt, c->d_counter, 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess && want) {
            int64_t m = (int64_t)cnt < cap ? (int64_t)cnt : cap;
            e = cudaMemcpy(ii, di, (size_t)m * 4, cudaMemcpyDefault);
            if (e == cudaSuccess) e = cudaMemcpy(jj, dj, (size_t)m * 4, cudaMemcpyDefault);
            if (e == cudaSuccess) e = cudaMemcpy(d, dd, (size_t)m * 8, cudaMemcpyDefault);
        }
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
    }
    if (di) DFREE(di);
    if (dj) DFREE(dj);
    if (dd) DFREE(dd);
    *n_out = (int64_t)cnt;
    return rc;
}

// ---- per-row extremes among the congeners (tdna_kernels.cuh: row_extremes_kernel) ---------------------------------------------
static int ensure_genus_lists(tdna_aln *a) {
    if (a->genus_lists_valid) return TDNA_OK;
    tdna_ctx *c = a->ctx;
    const int64_t n = a->n;
    if ((int64_t)a->h_genus.size() != n) return fail(TDNA_E_STATE, "labels are not set");
    std::unordered_map<int32_t, int32_t> slot_of;
    std::vector<int32_t> gslot((size_t)n);
    std::vector<int64_t> gstart;
    for (int64_t i = 0; i < n; i++) {
        auto it = slot_of.emplace(a->h_genus[i], (int32_t)slot_of.size()).first;
        gslot[i] = it->second;
    }
    gstart.assign(slot_of.size() + 1, 0);
    for (int64_t i = 0; i < n; i++) gstart[gslot[i] + 1]++;
    for (size_t g = 0; g < slot_of.size(); g++) gstart[g + 1] += gstart[g];
    std::vector<int64_t> cursor(gstart.begin(), gstart.end() - 1);
    std::vector<int32_t> gmem((size_t)n);
    for (int64_t i = 0; i < n; i++) gmem[cursor[gslot[i]]++] = (int32_t)i;  // ascending list positions within a genus
    if (!a->d_gslot) CU(DMALLOC(&a->d_gslot, (size_t)n * 4));
    if (!a->d_gmem) CU(DMALLOC(&a->d_gmem, (size_t)n * 4));
    if (a->d_gstart) { CU(DFREE(a->d_gstart)); a->d_gstart = nullptr; }
    CU(DMALLOC(&a->d_gstart, gstart.size() * 8));
    CU(cudaMemcpyAsync(a->d_gslot, gslot.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(a->d_gmem, gmem.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(a->d_gstart, gstart.data(), gstart.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // the vectors are stack-lifetime
    a->genus_lists_valid = true;
    return TDNA_OK;
}

int32_t tdna_row_extremes(tdna_aln *a, tdna_extreme_result *out) {
    if (!a || !out) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    TRY(ensure_packed(a));
    if (a->n == 0) return TDNA_OK;
    if (a->n > 0x7fffffff) return fail(TDNA_E_TOO_LARGE, "too many rows");
    TRY(ensure_genus_lists(a));
    const bool dev = is_device_ptr(out);
    tdna_extreme_result *d_out = out;
    if (!dev) {
        TRY(ensure_scratch(a, (size_t)a->n * sizeof(tdna_extreme_result)));
        d_out = reinterpret_cast<tdna_extreme_result *>(a->d_scratch);
    }
    switch (a->kmode) {
        case KM_P_AMBIG: TRY(launch_row_extremes<KM_P_AMBIG>(a, d_out)); break;
        case KM_P_NOAMBIG: TRY(launch_row_extremes<KM_P_NOAMBIG>(a, d_out)); break;
        case KM_K2P: TRY(launch_row_extremes<KM_K2P>(a, d_out)); break;
        default: TRY(launch_row_extremes<KM_TRANS>(a, d_out)); break;
    }
    if (!dev) CU(cudaMemcpyAsync(out, d_out, (size_t)a->n * sizeof(tdna_extreme_result), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

int32_t tdna_valid_conspecifics(tdna_aln *a, uint8_t *has_valid) {
    if (!a || !has_valid) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(need_labels(a));
    uint8_t *dv;
    CU(DMALLOC(&dv, (size_t)a->n));
    CU(cudaMemsetAsync(dv, 0, (size_t)a->n, c->stream));
    int rc = for_each_band(a, [&](int64_t b0, int64_t b1) -> int {
        valid_consp_kernel<<<(unsigned)(b1 - b0), 256, 0, c->stream>>>(a->d_mat, a->ld, a->n, b0, a->d_species, dv);
        c->launches++;
        CU(cudaGetLastError());
        return TDNA_OK;
    });
    if (rc == TDNA_OK) {
        int64_t rows = a->row_end - a->row_begin;
        cudaError_t e = cudaMemcpyAsync(has_valid + a->row_begin, dv + a->row_begin, (size_t)rows, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
    }
    DFREE(dv);
    return rc;
}

int32_t tdna_time_matrix_kernel(tdna_aln *a, int32_t reps, float *ms_per_launch) {
    if (!a || !ms_per_launch || reps < 1) return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(ensure_packed(a));
    int64_t rows = a->row_end - a->row_begin;
    if (band_rows_for(a) != rows) return fail(TDNA_E_TOO_LARGE, "matrix exceeds the memory budget");
    TRY(ensure_mat(a, rows));
    bool symmetric = a->row_begin == 0 && a->row_end == a->n;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    c->mat_owner = a;
    TRY(launch_tiles(a, a->row_begin, a->row_end, symmetric, a->d_mat, a->ld, nullptr));  // warm
    CU(cudaEventRecord(e0, c->stream));
    for (int r = 0; r < reps; r++) TRY(launch_tiles(a, a->row_begin, a->row_end, symmetric, a->d_mat, a->ld, nullptr));
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    a->mat_valid = true;
    *ms_per_launch = ms / reps;
    return TDNA_OK;
}

int32_t tdna_set_option(tdna_ctx *c, int32_t key, int64_t value) {
    if (!c) return fail(TDNA_E_ARG, "null ctx");
    switch (key) {
        case TDNA_OPT_BEST_MATCH_PATH:
            if (value < TDNA_PATH_AUTO || value > TDNA_PATH_STREAM) return fail(TDNA_E_ARG, "bad path");
            c->opt_best_match_path = (int)value;
            return TDNA_OK;
        case TDNA_OPT_COUNT_KERNEL:
            if (value < TDNA_KERNEL_AUTO || value > TDNA_KERNEL_TENSOR) return fail(TDNA_E_ARG, "bad kernel");
            c->opt_count_kernel = (int)value;
            return TDNA_OK;
        case TDNA_OPT_TILE_COLUMNS:
            if (value != 0 && value != 64 && value != 128) return fail(TDNA_E_ARG, "tile columns must be 0, 64 or 128");
            c->opt_tile_cn = (int)(value / 16);
            return TDNA_OK;
        case TDNA_OPT_HIST_SLOTS:
            if (value < 1024 || value > (1 << 26) || (value & (value - 1))) return fail(TDNA_E_ARG, "histogram slots: a power of two in [2^10, 2^26]");
            c->opt_hist_slots = (uint32_t)value;
            return TDNA_OK;
        case TDNA_OPT_MULTI:
            if (value < 0 || value > 2) return fail(TDNA_E_ARG, "TDNA_OPT_MULTI is 0, 1 or 2");
            c->comm.enabled = value != 0;
            c->comm.force = value == 2;
            return TDNA_OK;
        case TDNA_OPT_UNIT_ITEMS:
            if (value < 0 || value > 4096) return fail(TDNA_E_ARG, "unit items: 0 (default) to 4096");
            c->opt_unit_items = (int)value;
            return TDNA_OK;
        case TDNA_OPT_CLUSTER:
            if (value != 2 && value != 3 && value != 4) return fail(TDNA_E_ARG, "cluster mode: 2, 3 or 4");
            c->opt_cluster = (int)value;
            return TDNA_OK;
        case TDNA_OPT_PIPELINE_CHUNKS:
            if (value < 0 || value > 15) return fail(TDNA_E_ARG, "pipeline chunks: 0 (off) to 15");
            c->opt_pipeline_chunks = (int)value;
            return TDNA_OK;
        case TDNA_OPT_SEGMENT_ITEMS:
            if (value < 0) return fail(TDNA_E_ARG, "segment items: 0 (automatic) or a positive count");
            c->opt_segment_items = value;
            return TDNA_OK;
    }
    return fail(TDNA_E_ARG, "unknown option");
}

int32_t tdna_tensor_counts(tdna_aln *a, tdna_counts *out) {
    if (!a || !out) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    TRY(ensure_packed(a));
    if (a->kmode != KM_P_AMBIG && a->kmode != KM_TRANS) return fail(TDNA_E_STATE, "tdna_tensor_counts: uncorrected p or transversions only (for K2P the accumulator carries n and ts + tv)");
    if (ensure_tc(a) != 1) return fail(TDNA_E_STATE, "count kernel (b) needs uncorrected p with ambiguity allowed, K2P or transversions only, and L < 4096");
    if (a->tc_has_amb) return fail(TDNA_E_STATE, "tdna_tensor_counts: the alignment has IUPAC ambiguity letters -- kernel (b) only bounds such pairs (exact counts come from the bit-planes)");
    tdna_counts *dc = out;
    bool own = !is_device_ptr(out);
    size_t bytes = (size_t)a->n * a->n * sizeof(tdna_counts);
    if (own) CU(DMALLOC(&dc, bytes));
    CU(cudaMemsetAsync(dc, 0xff, bytes, c->stream));  // pairs the triangle does not cover stay -1
    int rc = launch_tc<true>(a, 1, 0, 1, dc);  // the bulk enumeration covers the whole triangle
    if (rc == TDNA_OK) {
        cudaError_t e = cudaSuccess;
        if (own) e = cudaMemcpyAsync(out, dc, bytes, cudaMemcpyDefault, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(TDNA_E_CUDA, cudaGetErrorString(e));
    }
    if (own) DFREE(dc);
    return rc;
}

// ---- FASTA on the device (tdna_fasta.cuh) ---------------------------------------------------------------------------------
struct tdna_fasta {
    tdna_ctx *ctx = nullptr;
    int64_t n = 0;
    int32_t width = 0;
    int64_t stride = 0;
    uint8_t *d_rows = nullptr;
    int *d_seq_length = nullptr, *d_flags = nullptr, *d_header_length = nullptr;
    long long *d_header_offset = nullptr;
};

// exclusive prefix sum of n ints into ptr[0..n] with the row-pointer kernels; scratch holds ceil(n / SCAN_BLOCK) sums
static int scan_ints(tdna_ctx *c, const int *cnt, int64_t n, long long *ptr, long long *scratch) {
    const int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scan_block_sums_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(cnt, n, scratch);
    scan_block_offsets_kernel<<<1, 1024, 0, c->stream>>>(scratch, nb, ptr + n);
    scan_apply_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(cnt, n, scratch, ptr);
    c->launches += 3;
    CU(cudaGetLastError());
    return TDNA_OK;
}

int32_t tdna_fasta_destroy(tdna_fasta *f) {
    if (!f) return TDNA_OK;
    tdna_ctx *c = f->ctx;
    cudaSetDevice(c->device);
    void *ptrs[] = {f->d_rows, f->d_seq_length, f->d_flags, f->d_header_length, f->d_header_offset};
    for (void *p : ptrs)
        if (p) DFREE(p);
    delete f;
    return TDNA_OK;
}

int32_t tdna_fasta_parse(tdna_ctx *c, const uint8_t *text, int64_t nbytes, tdna_fasta **out) {
    if (!c || !text || !out || nbytes < 0) return fail(TDNA_E_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    tdna_fasta *f = new tdna_fasta();
    f->ctx = c;
    *out = nullptr;
    if (nbytes == 0) { *out = f; return TDNA_OK; }
    // device scratch, all freed before returning
    uint8_t *d_text = nullptr;
    int *d_block_lines = nullptr, *d_kind = nullptr, *d_symbols = nullptr, *d_isname = nullptr, *d_max = nullptr;
    long long *d_block_first = nullptr, *d_scratch = nullptr, *d_line_start = nullptr, *d_rec_of_line = nullptr, *d_sym_before = nullptr, *d_rec_line = nullptr;
    int rc = TDNA_OK;
    auto cleanup = [&]() {
        void *ptrs[] = {d_block_lines, d_kind, d_symbols, d_isname, d_max, d_block_first, d_scratch, d_line_start, d_rec_of_line, d_sym_before, d_rec_line};
        for (void *p : ptrs)
            if (p) DFREE(p);
        if (d_text && d_text != text) DFREE(d_text);
    };
#define FCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); tdna_fasta_destroy(f); return fail(TDNA_E_CUDA, cudaGetErrorString(e_)); } } while (0)
    if (is_device_ptr(text)) d_text = const_cast<uint8_t *>(text);
    else {
        FCU(DMALLOC(&d_text, (size_t)nbytes));
        FCU(cudaMemcpyAsync(d_text, text, (size_t)nbytes, cudaMemcpyHostToDevice, c->stream));
    }
    const int64_t nblocks = (nbytes + fasta::BLOCK_BYTES - 1) / fasta::BLOCK_BYTES;
    FCU(DMALLOC(&d_block_lines, (size_t)nblocks * 4));
    FCU(DMALLOC(&d_block_first, (size_t)(nblocks + 1) * 8));
    FCU(DMALLOC(&d_scratch, (size_t)((std::max<int64_t>(nblocks, nbytes / 2 + 2) + SCAN_BLOCK - 1) / SCAN_BLOCK + 1) * 8));
    fasta::count_lines_kernel<<<(unsigned)nblocks, 256, 0, c->stream>>>(d_text, nbytes, d_block_lines);
    c->launches++;
    if ((rc = scan_ints(c, d_block_lines, nblocks, d_block_first, d_scratch)) != TDNA_OK) { cleanup(); tdna_fasta_destroy(f); return rc; }
    long long nlines = 0;
    FCU(cudaMemcpyAsync(&nlines, d_block_first + nblocks, 8, cudaMemcpyDeviceToHost, c->stream));
    FCU(cudaStreamSynchronize(c->stream));
    FCU(DMALLOC(&d_line_start, (size_t)nlines * 8));
    FCU(DMALLOC(&d_kind, (size_t)nlines * 4));
    FCU(DMALLOC(&d_symbols, (size_t)nlines * 4));
    FCU(DMALLOC(&d_isname, (size_t)nlines * 4));
    FCU(DMALLOC(&d_rec_of_line, (size_t)(nlines + 1) * 8));
    FCU(DMALLOC(&d_sym_before, (size_t)(nlines + 1) * 8));
    fasta::list_lines_kernel<<<(unsigned)nblocks, 32, 0, c->stream>>>(d_text, nbytes, d_block_first, d_line_start);
    fasta::classify_lines_kernel<<<(unsigned)((nlines + 255) / 256), 256, 0, c->stream>>>(d_text, nbytes, d_line_start, nlines, d_kind, d_symbols, d_isname);
    c->launches += 2;
    if ((rc = scan_ints(c, d_isname, nlines, d_rec_of_line, d_scratch)) != TDNA_OK || (rc = scan_ints(c, d_symbols, nlines, d_sym_before, d_scratch)) != TDNA_OK) {
        cleanup(); tdna_fasta_destroy(f); return rc;
    }
    long long nrec = 0;
    FCU(cudaMemcpyAsync(&nrec, d_rec_of_line + nlines, 8, cudaMemcpyDeviceToHost, c->stream));
    FCU(cudaStreamSynchronize(c->stream));
    f->n = nrec;
    if (nrec > 0) {
        FCU(DMALLOC(&d_rec_line, (size_t)nrec * 8));
        FCU(DMALLOC(&f->d_header_offset, (size_t)nrec * 8));
        FCU(DMALLOC(&f->d_header_length, (size_t)nrec * 4));
        FCU(DMALLOC(&f->d_seq_length, (size_t)nrec * 4));
        FCU(DMALLOC(&f->d_flags, (size_t)nrec * 4));
        FCU(DMALLOC(&d_max, 4));
        FCU(cudaMemsetAsync(d_max, 0, 4, c->stream));
        FCU(cudaMemsetAsync(f->d_flags, 0, (size_t)nrec * 4, c->stream));
        fasta::records_kernel<<<(unsigned)((nlines + 255) / 256), 256, 0, c->stream>>>(d_text, nbytes, d_line_start, nlines, d_kind, d_rec_of_line, d_sym_before, d_rec_line,
                                                                                    f->d_header_offset, f->d_header_length);
        fasta::record_lengths_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, c->stream>>>(d_rec_line, nrec, nlines, d_sym_before, f->d_seq_length, d_max);
        c->launches += 2;
        int width = 0;
        FCU(cudaMemcpyAsync(&width, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
        FCU(cudaStreamSynchronize(c->stream));
        f->width = width;
        f->stride = round_up(std::max(width, 1), 32);
        FCU(DMALLOC(&f->d_rows, (size_t)nrec * f->stride));
        FCU(cudaMemsetAsync(f->d_rows, '?', (size_t)nrec * f->stride, c->stream));
        fasta::scatter_symbols_kernel<<<(unsigned)((nlines * 32 + 255) / 256), 256, 0, c->stream>>>(d_text, nbytes, d_line_start, nlines, d_kind, d_rec_of_line, d_sym_before,
                                                                                                 d_rec_line, f->d_rows, f->stride, f->d_flags);
        fasta::external_gaps_kernel<<<(unsigned)((nrec * 32 + 255) / 256), 256, 0, c->stream>>>(f->d_rows, f->stride, (int)f->stride, f->d_seq_length, nrec, f->d_flags);
        c->launches += 2;
        FCU(cudaGetLastError());
    }
    FCU(cudaStreamSynchronize(c->stream));
#undef FCU
    cleanup();
    *out = f;
    return TDNA_OK;
}

int64_t tdna_fasta_count(const tdna_fasta *f) { return f ? f->n : 0; }
int32_t tdna_fasta_width(const tdna_fasta *f) { return f ? f->width : 0; }

int32_t tdna_fasta_records(tdna_fasta *f, int64_t *header_offset, int32_t *header_length, int32_t *seq_length, int32_t *flags) {
    if (!f) return fail(TDNA_E_ARG, "null argument");
    tdna_ctx *c = f->ctx;
    CU(cudaSetDevice(c->device));
    if (f->n == 0) return TDNA_OK;
    if (header_offset) CU(cudaMemcpyAsync(header_offset, f->d_header_offset, (size_t)f->n * 8, cudaMemcpyDefault, c->stream));
    if (header_length) CU(cudaMemcpyAsync(header_length, f->d_header_length, (size_t)f->n * 4, cudaMemcpyDefault, c->stream));
    if (seq_length) CU(cudaMemcpyAsync(seq_length, f->d_seq_length, (size_t)f->n * 4, cudaMemcpyDefault, c->stream));
    if (flags) CU(cudaMemcpyAsync(flags, f->d_flags, (size_t)f->n * 4, cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

int32_t tdna_fasta_symbols(tdna_fasta *f, uint8_t *out, int64_t row_stride) {
    if (!f || !out || row_stride < f->width) return fail(TDNA_E_ARG, "bad argument");
    tdna_ctx *c = f->ctx;
    CU(cudaSetDevice(c->device));
    if (f->n == 0 || f->width == 0) return TDNA_OK;
    CU(cudaMemcpy2DAsync(out, (size_t)row_stride, f->d_rows, (size_t)f->stride, (size_t)f->width, (size_t)f->n, cudaMemcpyDefault, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TDNA_OK;
}

int32_t tdna_alignment_create_from_fasta(tdna_ctx *c, tdna_fasta *f, tdna_aln **out) {
    if (!c || !f || !out) return fail(TDNA_E_ARG, "null argument");
    if (f->ctx != c) return fail(TDNA_E_ARG, "the records were parsed on another context");
    if (f->n < 1 || f->width < 1) return fail(TDNA_E_ARG, "no sequences");
    std::vector<int32_t> flags((size_t)f->n);
    TRY(tdna_fasta_records(f, nullptr, nullptr, nullptr, flags.data()));
    for (int64_t r = 0; r < f->n; r++)
        if (flags[r] & (fasta::FLAG_INVALID | fasta::FLAG_NEEDS_HOST))
            return fail(TDNA_E_SYMBOL, "record " + std::to_string(r) + ((flags[r] & fasta::FLAG_INVALID) ? ": a symbol outside the alphabet" : ": '[..]' / '(..)' groups need the host normaliser"));
    // device-to-device: the symbols never visit the host (rows keep the parser's stride; lengths are the records' own)
    return tdna_alignment_create(c, f->d_rows, f->n, f->width, f->stride, f->d_seq_length, out);
}

int32_t tdna_consensus(tdna_aln *a, const int32_t *members, const int64_t *offsets, int64_t nclusters, uint8_t *out, int64_t row_stride, int32_t *out_len) {
    if (!a || !members || !offsets || !out || !out_len || nclusters < 0 || row_stride < a->L) return fail(TDNA_E_ARG, "bad argument");
    if (nclusters == 0) return TDNA_OK;
    tdna_ctx *c = a->ctx;
    CU(cudaSetDevice(c->device));
    const int64_t total = offsets[nclusters];
    if (offsets[0] != 0 || total < 0) return fail(TDNA_E_ARG, "offsets must start at 0 and be host memory");
    for (int64_t k = 0; k < nclusters; k++)
        if (offsets[k + 1] <= offsets[k]) return fail(TDNA_E_ARG, "a bin has no member");
    int32_t *d_members = nullptr;
    long long *d_offsets = nullptr;
    uint8_t *d_out = nullptr;
    int *d_len = nullptr, *d_flags = nullptr;
    auto cleanup = [&]() {
        for (void *p : {(void *)d_members, (void *)d_offsets, (void *)d_out, (void *)d_len, (void *)d_flags})
            if (p) DFREE(p);
    };
#define KCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return fail(TDNA_E_CUDA, cudaGetErrorString(e_)); } } while (0)
    const int64_t stride = round_up(a->L, 32);
    KCU(DMALLOC(&d_members, (size_t)total * 4));
    KCU(DMALLOC(&d_offsets, (size_t)(nclusters + 1) * 8));
    KCU(DMALLOC(&d_out, (size_t)nclusters * stride));
    KCU(DMALLOC(&d_len, (size_t)nclusters * 4));
    KCU(DMALLOC(&d_flags, (size_t)nclusters * 4));
    std::vector<int32_t> hm(members, members + total);
    for (int32_t m : hm)
        if (m < 0 || m >= a->n) { cleanup(); return fail(TDNA_E_ARG, "member index out of range"); }
    KCU(cudaMemcpyAsync(d_members, hm.data(), (size_t)total * 4, cudaMemcpyHostToDevice, c->stream));
    KCU(cudaMemcpyAsync(d_offsets, offsets, (size_t)(nclusters + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    KCU(cudaMemsetAsync(d_out, '?', (size_t)nclusters * stride, c->stream));
    KCU(cudaMemsetAsync(d_flags, 0, (size_t)nclusters * 4, c->stream));
    fasta::consensus_kernel<<<(unsigned)nclusters, 256, 0, c->stream>>>(a->d_sym, a->d_len, a->sym_stride, a->L, d_members, d_offsets, d_out, stride, d_len);
    fasta::external_gaps_kernel<<<(unsigned)((nclusters * 32 + 255) / 256), 256, 0, c->stream>>>(d_out, stride, (int)stride, d_len, nclusters, d_flags);
    c->launches += 2;
    KCU(cudaGetLastError());
    KCU(cudaMemcpy2DAsync(out, (size_t)row_stride, d_out, (size_t)stride, (size_t)a->L, (size_t)nclusters, cudaMemcpyDefault, c->stream));
    KCU(cudaMemcpyAsync(out_len, d_len, (size_t)nclusters * 4, cudaMemcpyDefault, c->stream));
    KCU(cudaStreamSynchronize(c->stream));
#undef KCU
    cleanup();
    return TDNA_OK;
}

// The integer code of count kernel (b) for an alignment of L columns and `dims` symbol dimensions: code[0..3] = a, b, c, d and the
// power of two W.  Pure host arithmetic (no device needed): the CPU tests check its algebra.
int32_t tdna_tensor_code(int32_t L, int32_t dims, int32_t *code, uint32_t *W) {
    if (L < 1 || dims < 1 || !code || !W) return fail(TDNA_E_ARG, "bad argument");
    tc::EncodeValues v = {};
    uint32_t w = 0;
    if (!find_code(L, dims, v, w)) return fail(TDNA_E_STATE, "no code with int8 entries for this width and dimension count");
    code[0] = v.a; code[1] = v.b; code[2] = v.c; code[3] = v.d;
    *W = w;
    return TDNA_OK;
}

int32_t tdna_tensor_tables(int32_t mode, int32_t ambiguous_allowed, int32_t L, int32_t five_dims, int8_t *row_side, int8_t *col_side, int32_t *dims, uint32_t *W) {
    if (!row_side || !col_side || !dims || !W) return fail(TDNA_E_ARG, "null argument");
    const int kmode = mode == TDNA_PDM_K2P ? KM_K2P : (mode == TDNA_PDM_TRANS_ONLY ? KM_TRANS : (ambiguous_allowed ? KM_P_AMBIG : KM_P_NOAMBIG));
    tc::EncodeValues v;
    uint32_t w = 0;
    int d = 0;
    if (!tc_code_for(kmode, L, five_dims != 0, v, w, d)) return fail(TDNA_E_STATE, "count kernel (b) has no code for this mode / width");
    for (int code = 0; code < 6; code++)
        for (int k = 0; k < 8; k++) { row_side[code * 8 + k] = v.ta[code][k]; col_side[code * 8 + k] = v.tb[code][k]; }
    *dims = d;
    *W = w;
    return TDNA_OK;
}

int32_t tdna_last_count_kernel(tdna_aln *a) {
    if (!a) return 0;
    return a->last_kernel;
}

int32_t tdna_multi_phase_ms(tdna_ctx *c, float *ms, int32_t cap, int32_t *n_out) {
    if (!c || !ms || !n_out || cap < 0) return fail(TDNA_E_ARG, "bad argument");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    int k = 0;
    for (; k + 1 < c->n_ph && k < cap; k++) CU(cudaEventElapsedTime(&ms[k], c->ev_ph[k], c->ev_ph[k + 1]));
    *n_out = k;
    return TDNA_OK;
}

int64_t tdna_pipelined_passes(tdna_ctx *c) { return c ? c->pipelined_passes : 0; }

int32_t tdna_host_call_phase_ms(tdna_ctx *c, double *ms, int32_t cap, int32_t *n_out) {
    if (!c || !ms || !n_out) return fail(TDNA_E_ARG, "null argument");
    int k = 0;
    for (int i = 1; i < 10 && k < cap; i++, k++) ms[k] = (c->hc_t[i] > 0 && c->hc_t[0] > 0) ? c->hc_t[i] - c->hc_t[0] : -1.0;
    *n_out = k;
    return TDNA_OK;
}

int32_t tdna_last_pass_stats(tdna_aln *a, int64_t *candidates, int64_t *queued) {
    if (!a) return fail(TDNA_E_ARG, "null alignment");
    if (candidates) *candidates = a->last_ncand;
    if (queued) *queued = a->last_queue;
    return TDNA_OK;
}

int32_t tdna_tensor_dims(tdna_aln *a) {
    if (!a || a->tc_state != 1) return 0;
    return a->tc_dims;
}

int32_t tdna_last_tile_pass_ms(tdna_ctx *c, float *ms) {
    if (!c || !ms) return fail(TDNA_E_ARG, "null argument");
    if (!c->ev_valid) return fail(TDNA_E_STATE, "no tile pass has run yet");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev_k1));
    if (c->ev_split) {  // seed launch + bulk launch; the thresholds (and, on several GPUs, the all-reduce) between them are not tile work
        float seed_ms = 0, bulk_ms = 0;
        CU(cudaEventElapsedTime(&seed_ms, c->ev_k0, c->ev_s1));
        CU(cudaEventElapsedTime(&bulk_ms, c->ev_b0, c->ev_k1));
        *ms = seed_ms + bulk_ms;
        return TDNA_OK;
    }
    CU(cudaEventElapsedTime(ms, c->ev_k0, c->ev_k1));
    return TDNA_OK;
}

int32_t tdna_measure_int_peaks(tdna_ctx *c, double *popc_per_s, double *lop3_per_s) {
    if (!c) return fail(TDNA_E_ARG, "null ctx");
    CU(cudaSetDevice(c->device));
    const int blocks = c->sm_count * 8, iters = 4096;
    uint32_t *out;
    CU(cudaMalloc(&out, (size_t)blocks * 256 * 4));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double res[2];
    for (int which = 0; which < 2; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(e0, c->stream));
            if (which == 0) int_peak_kernel<0><<<blocks, 256, 0, c->stream>>>(out, iters, 12345u + rep);
            else int_peak_kernel<1><<<blocks, 256, 0, c->stream>>>(out, iters, 12345u + rep);
            c->launches++;
            CU(cudaEventRecord(e1, c->stream));
            CU(cudaEventSynchronize(e1));
            float ms;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        res[which] = (double)blocks * 256 * iters * 32.0 / (best * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (popc_per_s) *popc_per_s = res[0];
    if (lop3_per_s) *lop3_per_s = res[1];
    return TDNA_OK;
}

}  // extern "C"
