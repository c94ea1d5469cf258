/*
 * taxondna_cuda.h -- C ABI of the B200 (sm_100a) all-vs-all pairwise-distance engine that
 * replaces the bodies of TaxonDNA's Sequence.getPairwise family and the O(N^2) loops of its
 * consumers.  The reference is 100% Java and has no native interface for this path
 * (SURVEY.md section 8b), so each entry point below cites the *Java method(s)* whose body a host
 * layer (java/ via Panama FFM; taxondna_b200/host in this repo) replaces with a call to it.
 * Paths are relative to src/main/java/com/ggvaidya/TaxonDNA/ of the reference.
 *
 * Conventions
 *  - C99, no C++ types, no exceptions.  Every function returns int32_t: 0 = TDNA_OK, else a
 *    TDNA_E_* code; tdna_last_error() gives the text (owned by the library, valid until the next
 *    call on the same thread).
 *  - Buffers are caller-allocated and caller-owned; the library never keeps a caller pointer
 *    after the call returns.  Output pointers may be HOST or DEVICE addresses (the copy uses
 *    cudaMemcpyDefault / UVA), so a torch tensor's data_ptr() works as well as a Java
 *    MemorySegment.  Input pointers likewise.
 *  - One context per GPU and per process ("one process per GPU"); a context owns one CUDA
 *    stream.  Calls on one context are serialised by the caller (the reference serialises all
 *    analyses behind SequenceList's static lock, Common/DNA/SequenceList.java:403-480).
 *  - There is NO CPU fallback: without a CUDA device tdna_init fails with TDNA_E_CUDA.
 *  - Row index == position in the SequenceList order the host layer passed (SORT_BYNAME in the
 *    GUI, SpeciesIdentifier/SequencePanel.java:46,177-249).
 *  - Numerics: integer counts, uncorrected p and transversion distances are bit-exact (fp64, the reference's operation order, no
 *    FMA contraction).  K2P distances (Sequence.java:1559-1570) use a restatement of fdlibm's log (= StrictMath.log): bit-identical
 *    to that, and within 1 ulp of the log terms of any conforming JVM's Math.log (:1564), which HotSpot may intrinsify -- a last-bit
 *    difference can only matter where Settings.makeLongFromDouble truncates d * 10^5 within an ulp of an integer.
 */
#ifndef TAXONDNA_CUDA_H
#define TAXONDNA_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDNA_ABI_VERSION 2

enum {
    TDNA_OK = 0,
    TDNA_E_CUDA = 1,     /* CUDA runtime error (incl. "no device") */
    TDNA_E_ARG = 2,      /* bad argument */
    TDNA_E_SYMBOL = 3,   /* a symbol outside the normalised alphabet (Sequence.changeSequence would throw) */
    TDNA_E_STATE = 4,    /* call order (e.g. labels needed but not set) */
    TDNA_E_TOO_LARGE = 5,/* a device-resident result does not fit the memory budget */
    TDNA_E_CANCELLED = 6 /* tdna_cancel() was called (DelayAbortedException, Common/DelayCallback.java:48) */
};

/* Sequence.PDM_* (Common/DNA/Sequence.java:53-55) */
enum { TDNA_PDM_UNCORRECTED = 0, TDNA_PDM_K2P = 1, TDNA_PDM_TRANS_ONLY = 2 };

typedef struct tdna_ctx tdna_ctx;
typedef struct tdna_aln tdna_aln;

/* Integer counts of one pair; which are meaningful depends on the mode:
 *   shared : Sequence.getSharedLength (Sequence.java:1414-1485), mode-specific
 *   c1     : uncorrected -> countIdentical (:1322-1341); K2P -> n (:1514-1528); trans-only -> countTransversions (:1352-1371)
 *   c2, c3 : K2P -> nTransitions, nTransversions (:1530-1541); else 0 */
typedef struct {
    int32_t shared, c1, c2, c3;
} tdna_counts;

/* Per-query summary: everything BestMatch.run (SpeciesIdentifier/BestMatch.java:238-506) and
 * BlockAnalysis.run (SpeciesIdentifier/BlockAnalysis.java:238-350) read off the list that
 * SortedSequenceList.sortAgainst (Common/DNA/SortedSequenceList.java:78-128) builds for one
 * query, in the well-formed case, plus the evidence the host layer needs to
 * decide that the case IS well formed (near_* / unnamed_* / TDNA_ROW_NONFINITE); rows that are
 * not are re-resolved on the host from tdna_row_distances with the reference's exact sort.
 * "valid" = not (d < 0), j != q.  key(j) = (d, j conspecific-to-q ? 0 : 1, j)  -- the order
 * SortedSequenceComparator (SortedSequenceList.java:197-261) induces when it is consistent. */
typedef struct {
    double d_best;   /* distance of the arg-min key over valid j; -1.0 if none ("No match") */
    double d_con;    /* min over valid NAMED conspecific j (first_con, BestMatch.java:343-347); -1.0 if none */
    double d_allo;   /* min over valid NAMED allospecific j (first_allo, :348-352); -1.0 if none */
    double d_other;  /* min key over valid named j with species != species(best), j != best (end of the
                        BlockAnalysis block, BlockAnalysis.java:282-294); -1.0 if none */
    int32_t best, con, allo, other;   /* row indices of the above; -1 if none */
    int32_t shared_con, shared_allo;  /* getSharedLength(query, first_con/first_allo) (BestMatch.java:372,386) */
    int32_t tie_count;                /* # valid j != best with d == d_best exactly (count_bestMatches, :296-326) */
    int32_t tie_species_min, tie_species_max; /* species ids over the NAMED ties (excl. best); INT32_MAX / -1 if none */
    int32_t tie_unnamed;              /* # unnamed ties (the reference NPEs on these, BestMatch.java:319) */
    int32_t block_count;              /* # valid named j != best of species(best) with key(j) < key(other) */
    int32_t near_best, near_con, near_allo, near_other; /* # valid j with 0 < |d - d_ref| < 2e-5 */
    int32_t unnamed_at_con, unnamed_at_allo, unnamed_at_other; /* # unnamed valid j with d == d_ref */
    int32_t n_valid;                  /* # valid j != q */
    int32_t flags;                    /* TDNA_ROW_* */
    int32_t reserved, reserved2;      /* zero; the record is exactly 120 bytes with no padding */
} tdna_row_result;

enum {
    TDNA_ROW_SELF_VALID = 1, /* d(q,q) is valid, so q itself heads the sorted list */
    TDNA_ROW_NONFINITE = 2   /* some valid d(q,j) is NaN or +Inf (K2P saturation): host must resolve */
};

/* Pair classes of PairwiseDistribution (Common/DNA/PairwiseDistribution.java:41-42, 161-203) */
enum { TDNA_PD_INTRA = 0, TDNA_PD_INTER = 1 };

/* ---- lifetime ------------------------------------------------------------------------- */
int32_t tdna_abi_version(void);
/* Creates the context for CUDA device `device` (one per process in the torchrun layout).
 * Threading (the reference runs one analysis at a time under SequenceList's static lock, SequenceList.java:403-440): calls may come
 * from any thread; calls on ONE context are serialised by the caller; contexts on DIFFERENT devices are independent.  Use one
 * context per device and process: the streaming passes keep their parameters in per-device constant memory. */
int32_t tdna_init(int32_t device, tdna_ctx **out);
int32_t tdna_shutdown(tdna_ctx *ctx);
const char *tdna_last_error(void);
/* ---- several GPUs of one box -----------------------------------------------------------------------------------
 * The reference runs every analysis on one thread of one JVM (SpeciesIdentifier.java:668-681); here the O(N^2) pass of ONE analysis
 * is shared by up to 8 GPUs.  One context per GPU; the contexts form a communicator (NCCL over NVLink / NVSwitch, bound with dlopen
 * at the first tdna_comm_* call -- a single-GPU host never loads it).  From then on every whole-range tdna_analyze / tdna_best_match*
 * / tdna_class_distances / tdna_edges_below / tdna_class_histogram on an alignment is a COLLECTIVE call: every rank makes it, on an
 * alignment with the same content, labels and configuration, from its own thread (one JVM: one thread per GPU) or its own process
 * (torchrun: one process per GPU).  Inside the call the ranks share the triangle's tiles, all-reduce the seeded per-row minima,
 * route the candidate records to the rank that owns the row, and all-gather the row records -- all on the contexts' streams, with
 * one host synchronisation at the end.  Per-row records and histograms come back complete and bit-identical to one GPU's on every
 * rank; variable-length lists (class distances, edges) come back as this rank's share (n_*_local entries) with the totals in n_*.
 *   tdna_comm_unique_id : rank 0 makes the 128-byte id; the host hands it to every rank (a Java array, an MPI / torch broadcast)
 *   tdna_comm_init      : collective; the context joins as `rank` of `nranks`
 *   tdna_init_multi     : ONE process, one thread: creates ndev contexts (out[k] on devices[k], rank k) with their communicator */
#define TDNA_COMM_ID_BYTES 128
int32_t tdna_comm_unique_id(uint8_t *id128);
int32_t tdna_comm_init(tdna_ctx *ctx, const uint8_t *id128, int32_t rank, int32_t nranks);
int32_t tdna_comm_destroy(tdna_ctx *ctx);
int32_t tdna_comm_info(tdna_ctx *ctx, int32_t *rank, int32_t *nranks);
int32_t tdna_init_multi(const int32_t *devices, int32_t ndev, tdna_ctx **out);

/* The context's CUDA stream (a cudaStream_t) so that a caller can order its own work/events on it. */
void *tdna_stream(tdna_ctx *ctx);
/* Bytes of HBM the library may use for device-resident N x N results (default 64 GiB). */
int32_t tdna_set_memory_budget(tdna_ctx *ctx, int64_t bytes);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t tdna_launch_count(tdna_ctx *ctx);

/* ---- alignment ------------------------------------------------------------------------ */
/* symbols: n rows of `row_stride` bytes, alphabet after Sequence.changeSequence
 * (Common/DNA/Sequence.java:751-853): ACGT RYKMSWBDHVN - _ ? ; lengths[i] <= L is the row's own
 * length (shorter rows behave as '?'-padded: the reference loops to min(len1,len2),
 * Sequence.java:1326-1327; TDNA_E_ARG if a host lengths[i] is outside [0, L]); lengths may be NULL (all L).  The symbols are copied
 * to the device and packed there into bit-planes; nothing of the caller's memory is retained.  Host symbols of 8,192+ rows are
 * uploaded in row chunks while the chunks that have arrived are packed and (for the default configuration: uncorrected p,
 * ambiguity allowed) encoded into count kernel (b)'s operands, so the PCIe time hides that work; with a communicator each rank
 * uploads 1/G of the rows and the slices are all-gathered over NVLink. */
int32_t tdna_alignment_create(tdna_ctx *ctx, const uint8_t *symbols, int64_t n, int32_t L,
                              int64_t row_stride, const int32_t *lengths, tdna_aln **out);
int32_t tdna_alignment_destroy(tdna_aln *aln);
/* Labels per row, all computed by the host layer from the names:
 *   species_id   : id of getSpeciesName() (Sequence.java:446-450), -1 = null (unnamed)
 *   genus_id     : id of getGenusName() (:453), any int (the empty genus is an id too)
 *   epithet_rank : rank of getSpeciesNameOnly() (:458) under String.compareTo (equal strings equal rank)
 *   fullname_rank: rank of getFullName() under String.compareTo (equal strings equal rank)
 * Needed by tdna_best_match / tdna_class_distances only. */
int32_t tdna_alignment_set_labels(tdna_aln *aln, const int32_t *species_id, const int32_t *genus_id,
                                  const int32_t *epithet_rank, const int32_t *fullname_rank);
/* The three Sequence statics (Sequence.java:53-66, setters :86-136). Re-packs the planes. */
int32_t tdna_set_config(tdna_aln *aln, int32_t mode, int32_t min_overlap, int32_t ambiguous_allowed);
/* Multi-GPU: this context computes query rows [row_begin,row_end) only (all columns); results for
 * other rows are left untouched.  Default: all rows.  See DESIGN.md "Multi-GPU". */
int32_t tdna_set_row_range(tdna_aln *aln, int64_t row_begin, int64_t row_end);

/* ---- per-pair / per-row (Sequence.*, SortedSequenceList, QuerySequence) ----------------- */
/* Sequence.getPairwiseNoBuffer / getSharedLength / countIdentical / countTransversions /
 * getK2PDistance (Sequence.java:1322-1717): one pair, a one-warp launch. counts or d may be NULL. */
int32_t tdna_pair(tdna_aln *aln, int64_t i, int64_t j, tdna_counts *counts, double *d);
/* A batch of arbitrary pairs in one launch: (counts[k], d[k]) = pair (ii[k], jj[k]).  What PairwiseDistances
 * (Common/DNA/PairwiseDistances.java:128-185) and PairwiseExplorer need: the distances of an explicit pair list with
 * the pair identities kept.  ii / jj / counts / d may be host or device pointers; counts or d may be NULL. */
int32_t tdna_pairs(tdna_aln *aln, const int32_t *ii, const int32_t *jj, int64_t m, tdna_counts *counts, double *d);
/* d[j] = query.getPairwise(seq_j) for all j (what SortedSequenceList.sortAgainst evaluates,
 * SortedSequenceList.java:104-118); shared may be NULL. */
int32_t tdna_row_distances(tdna_aln *aln, int64_t q, double *d, int32_t *shared);

/* A sequence that is NOT a row of the alignment against every row: d[n] (and optionally shared[n], counts[n]).
 * `symbols` holds `len` bytes of the normalised alphabet (host or device memory); sites past the alignment's width are ignored
 * (every comparison stops at the shorter sequence, Sequence.java:1417-1419).  Replaces the N getPairwise calls of
 * SortedSequenceList.sortAgainst (SortedSequenceList.java:78-128) for a foreign query, as issued by
 * QuerySequence.query (QuerySequence.java:113-145). */
int32_t tdna_query_distances(tdna_aln *aln, const uint8_t *symbols, int32_t len, double *d, int32_t *shared, tdna_counts *counts);

/* ---- FASTA on the device (the format on the input side of the path) -------------------------------------
 * FastaFile.appendFromFile / makeSequence (Common/DNA/formats/FastaFile.java:144-289) and the alphabet part of
 * Sequence.changeSequence (Common/DNA/Sequence.java:751-840) as byte-parallel kernels: `text` (host or device, nbytes) ->
 * records.  Per record: where its header line (after '>') sits in `text` (names are parsed by the host layer as before), its
 * sequence length, flags; and the n x width symbol matrix in the normalised alphabet ('?' padded), resident on the device. */
typedef struct tdna_fasta tdna_fasta;
enum { TDNA_FASTA_URACIL = 1,      /* U converted to T (FastaFile.java:273-285 warns) */
       TDNA_FASTA_INVALID = 2,     /* a symbol isValid rejects (Sequence.java:927-965), or a '_' where a gap stroke meets it (:812-838) */
       TDNA_FASTA_NEEDS_HOST = 4   /* "[AG]" / "(AG)" group syntax (:758-796): normalise this record on the host */ };
int32_t tdna_fasta_parse(tdna_ctx *ctx, const uint8_t *text, int64_t nbytes, tdna_fasta **out);
int32_t tdna_fasta_destroy(tdna_fasta *f);
int64_t tdna_fasta_count(const tdna_fasta *f);
int32_t tdna_fasta_width(const tdna_fasta *f);
/* any of the four outputs may be NULL; n entries each */
int32_t tdna_fasta_records(tdna_fasta *f, int64_t *header_offset, int32_t *header_length, int32_t *seq_length, int32_t *flags);
int32_t tdna_fasta_symbols(tdna_fasta *f, uint8_t *out, int64_t row_stride);
/* an alignment straight from the parsed records: the symbols never visit the host.  TDNA_E_SYMBOL if a record is flagged
 * INVALID or NEEDS_HOST. */
int32_t tdna_alignment_create_from_fasta(tdna_ctx *ctx, tdna_fasta *f, tdna_aln **out);

/* ---- consensus of bins (the step after Cluster: Cluster.makeConsensusOfBin, SpeciesIdentifier/Cluster.java:123-138) -------------
 * The left fold of Sequence.getConsensus (Common/DNA/Sequence.java:1374-1398; consensus() :1008-1050) over each bin's members, as
 * the reference computes it with a re-normalisation after every step.  Bin k = rows members[offsets[k] .. offsets[k + 1]) (host
 * arrays, offsets[0] = 0).  out: nclusters rows of row_stride >= L bytes in the normalised alphabet ('?' padded), out_len[k] = the
 * longest member's length (host or device). */
int32_t tdna_consensus(tdna_aln *aln, const int32_t *members, const int64_t *offsets, int64_t nclusters, uint8_t *out, int64_t row_stride,
                       int32_t *out_len);

/* ---- full matrix (config C2; Cluster.java:605-626 prints one per cluster) ---------------- */
/* d: (row_end-row_begin) x n doubles, row-major, ld = n.  counts (optional): same shape. */
int32_t tdna_matrix(tdna_aln *aln, double *d, tdna_counts *counts);
/* Device-resident variant: computes into library memory (TDNA_E_TOO_LARGE if over budget);
 * *d_dev receives the device pointer (valid until the alignment/config changes). */
int32_t tdna_matrix_compute(tdna_aln *aln, const double **d_dev);

/* ---- Best Match / Best Close Match / All Species Barcodes -------------------------------- */
/* out: n records (only rows in the row range are written).  Streams row bands through HBM when
 * the N x N matrix exceeds the budget, so n = 1,000,000 works. */
int32_t tdna_best_match(tdna_aln *aln, tdna_row_result *out);
/* Same, results left on the device; *out_dev receives the device pointer to n records. */
int32_t tdna_best_match_compute(tdna_aln *aln, const tdna_row_result **out_dev);

/* ---- streaming Best Match, in parts (multi-GPU) ------------------------------------------------ */
/* tdna_best_match on the whole row range does NOT materialise the N x N matrix: one pass over the
 * triangle of tiles keeps, per query row, only the CANDIDATE columns whose distance is not above a
 * per-row bound that provably covers every distance tdna_row_result refers to (+ the near-tie
 * window), and the exact two-sweep summary then runs on those short lists (DESIGN.md "streaming
 * Best Match").  The two halves are exposed so that several GPUs can share one pass:
 *   part k of m takes the k-th of m equal, contiguous slices of the triangle's tile list; its candidates / per-row invalid-pair counts /
 *   per-row flags stay on the device (pointers in *out, valid until the next call on the alignment);
 *   the caller concatenates the candidate records of all parts (NCCL all_gather), sums the invalid
 *   counts and ORs the flags (NCCL all_reduce), and hands the union to tdna_best_match_merge. */
typedef struct {
    const int32_t *cand_row;      /* device, n_cand: query row q */
    const int32_t *cand_col;      /* device, n_cand: column j */
    const double *cand_d;         /* device, n_cand: d(q, j), valid and finite */
    int64_t n_cand;
    const int32_t *invalid_count; /* device, n: # j != q with getSharedLength < minOverlap seen by this part */
    const int32_t *flags;         /* device, n: TDNA_ROW_* seen by this part */
} tdna_stream_part;
/* Optional first half of a part: initialises the per-row state and runs the part's share of the
 * DIAGONAL tile blocks only, then exposes the per-row minima the candidate thresholds are derived
 * from -- 3n uint64 (fp64 bit patterns; n conspecific minima, then 2n even/odd-species minima) on
 * the device.  With several parts the caller all-reduces them with MIN (bit patterns of
 * non-negative doubles order like integers) so that every part starts the bulk of the triangle with
 * thresholds from ALL list neighbours, then calls tdna_best_match_part(..., seeded = 1, ...). */
int32_t tdna_best_match_seed(tdna_aln *aln, int32_t part, int32_t nparts, uint64_t **minima_dev);
/* seeded = 0: the whole part in one call.  seeded = 1: continue after tdna_best_match_seed. */
int32_t tdna_best_match_part(tdna_aln *aln, int32_t part, int32_t nparts, int32_t seeded, tdna_stream_part *out);
/* Inputs are DEVICE pointers (the part's own, or gathered); out: n records, host or device. */
int32_t tdna_best_match_merge(tdna_aln *aln, const int32_t *cand_row, const int32_t *cand_col, const double *cand_d,
                              int64_t n_cand, const int32_t *invalid_count, const int32_t *flags, tdna_row_result *out);

/* The same exchange without any host-side glue: the part writes PACKED outputs straight into caller-provided DEVICE
 * buffers that can be handed to NCCL as they are --
 *   rec   : capacity seg_stride records of two int64: (row << 32 | column), fp64 bits of d(row, column)
 *   count : one int64, the number of records written (TDNA_E_TOO_LARGE if it would exceed seg_stride)
 *   state : n int64, invalid_count | (self-valid count << 32) | (non-finite count << 48): SUM-reducible
 * and tdna_best_match_merge_gathered consumes the all_gathered records (nseg segments of seg_stride records, their
 * counts on the device) and the all_reduced state.  out: n records, host or device, or NULL (results stay on the device). */
int32_t tdna_best_match_part_packed(tdna_aln *aln, int32_t part, int32_t nparts, int32_t seeded, int64_t *rec, int64_t seg_stride,
                                    int64_t *count, int64_t *state);
int32_t tdna_best_match_merge_gathered(tdna_aln *aln, const int64_t *rec, int32_t nseg, int64_t seg_stride, const int64_t *seg_count,
                                       const int64_t *state, tdna_row_result *out);

/* Routed variant of the same exchange (row-sharded merge, what bench.py runs at N > 1): the records are bucketed by the part that
 * OWNS their row, owner(row) = row / tdna_rows_per_part(n, nparts).  `rec` holds nparts segments of seg_stride records (16 B each),
 * segment g at rec + 2 * g * seg_stride with counts[g] records; TDNA_E_TOO_LARGE if a segment is too small.  The caller exchanges
 * the segments and counts with an all-to-all (equal splits) and SUM-reduces `state` ... */
int64_t tdna_rows_per_part(int64_t n, int32_t nparts);
int32_t tdna_best_match_part_routed(tdna_aln *aln, int32_t part, int32_t nparts, int32_t seeded, int64_t *rec, int64_t seg_stride, int64_t *counts,
                                    int64_t *state);
/* ... then every part finalises ONLY the rows it owns from the segments it received (same layout: segment g = the records part g
 * sent) and writes tdna_rows_per_part() records (zero padded past the end of the list) to out_slice (host or device); an all-gather
 * of the slices is the whole result, bit-identical to one GPU's. */
int32_t tdna_best_match_merge_routed(tdna_aln *aln, int32_t part, int32_t nparts, const int64_t *rec, int64_t seg_stride, const int64_t *seg_count,
                                     const int64_t *state, tdna_row_result *out_slice);

/* ---- one pass, several analyses ------------------------------------------------------------------ */
/* SpeciesIdentifier runs Pairwise Summary, the threshold percentile, Best (Close) Match and Cluster on the
 * SAME list one after the other (PairwiseSummary.java:131-151 -> BestMatch.java:107-120 "Compute from
 * Pairwise Summary" -> BestMatch.run); each of them is O(N^2) getPairwise calls in the reference.  This
 * entry point produces any subset of their inputs from ONE streaming pass over the triangle. */
enum { TDNA_WANT_BEST_MATCH = 1, /* tdna_row_result per row (as tdna_best_match) */
       TDNA_WANT_INTRA = 2,      /* PairwiseDistribution PD_INTRA distances (as tdna_class_distances) */
       TDNA_WANT_INTER = 4,      /* PairwiseDistribution PD_INTER distances */
       TDNA_WANT_EDGES = 8,      /* pairs with 0 <= d <= edge_threshold (as tdna_edges_below) */
       TDNA_WANT_HIST_INTRA = 16,/* PD_INTRA as a histogram over the distinct distances (as tdna_class_histogram); excludes TDNA_WANT_INTRA */
       TDNA_WANT_HIST_INTER = 32 /* PD_INTER as a histogram; excludes TDNA_WANT_INTER */ };
/* One distinct distance of a class and how many times PairwiseDistribution pushes it (PairwiseDistribution.java:77-103, 161-203;
 * -1 is dropped, NaN is one entry).  The host layer sorts the entries by (float)d and reads counts, ranges and the threshold
 * percentile (PairwiseSummary.java:376-418) off the cumulative counts -- the list the reference sorts (4 B per pair: 2 GB for the
 * congeneric class of 10^6 barcodes) never exists. */
typedef struct {
    double d;
    int64_t count;
} tdna_hist_entry;
typedef struct {
    int32_t want;            /* in: TDNA_WANT_* bits */
    int32_t reserved;
    tdna_row_result *rows;   /* in: n records, host or device; NULL leaves them on the device (tdna_best_match_compute) */
    float *intra;            /* in: capacity intra_cap (host or device); NULL = count only */
    int64_t intra_cap;
    float *inter;
    int64_t inter_cap;
    double edge_threshold;
    int32_t *edge_i, *edge_j; /* in: capacity edge_cap; NULL = count only */
    double *edge_d;
    int64_t edge_cap;
    int64_t n_intra, n_inter, n_edges; /* out: totals (may exceed the capacities: only the first cap are stored) */
    int64_t n_intra_local, n_inter_local, n_edges_local; /* out: entries THIS rank produced (== the totals on one GPU); with a
                                                            communicator the lists hold this rank's share and the host concatenates */
    tdna_hist_entry *hist_intra, *hist_inter; /* in: capacity hist_*_cap entries (host or device), unsorted; NULL = count only */
    int64_t hist_intra_cap, hist_inter_cap;
    int64_t n_hist_intra, n_hist_inter;       /* out: distinct distances (may exceed the capacity: call again); with a communicator the
                                                 histograms are COMPLETE on every rank (merged over NVLink) */
    int64_t hist_intra_pushes, hist_inter_pushes; /* out: sum of the counts = PairwiseDistribution.countValidComparisons() */
} tdna_analysis;
int32_t tdna_analyze(tdna_aln *aln, tdna_analysis *io);

/* One analysis of a list that is not on the device yet, in ONE call: what a module does when its list has just been loaded or edited
 * (SpeciesIdentifier's modules rebuild everything on dataChanged(), UIExtension.java:40).  Same results as tdna_alignment_create +
 * tdna_alignment_set_labels + tdna_set_config + tdna_analyze, one downcall instead of four.  Nothing of the caller's memory is
 * referenced after the call returns.  *aln_out (optional) receives the resident alignment for further calls; with aln_out == NULL it
 * is destroyed. */
typedef struct {
    const uint8_t *symbols;   /* as tdna_alignment_create */
    int64_t n;
    int32_t L;
    int32_t mode;             /* TDNA_PDM_* */
    int64_t row_stride;
    const int32_t *lengths;   /* may be NULL */
    const int32_t *species_id, *genus_id, *epithet_rank, *fullname_rank; /* as tdna_alignment_set_labels; all NULL = no labels */
    int32_t min_overlap;
    int32_t ambiguous_allowed;
} tdna_host_list;
int32_t tdna_analyze_host(tdna_ctx *ctx, const tdna_host_list *in, tdna_analysis *io, tdna_aln **aln_out);

/* ---- options ------------------------------------------------------------------------------------ */
enum { TDNA_OPT_BEST_MATCH_PATH = 1, /* TDNA_PATH_*: how tdna_best_match computes a whole-range request */
       TDNA_OPT_TILE_COLUMNS = 2,    /* 0 = automatic, 64 or 128: column width of a popc count tile */
       TDNA_OPT_COUNT_KERNEL = 3,    /* TDNA_KERNEL_*: which count kernel the streaming pass uses */
       TDNA_OPT_HIST_SLOTS = 5,      /* first size of a class histogram's table for alignments created from now on (a power of two,
                                        default 2^18; a pass that fills a table repeats with four times the slots) */
       TDNA_OPT_UNIT_ITEMS = 9,      /* several GPUs: tile pairs per unit of the cyclic dealing of the triangle (0 = default) */
       TDNA_OPT_CLUSTER = 8,         /* count kernel (b), bulk launch: 2 = pairs of CTAs sharing a b-block by TMA multicast; 4 = clusters of
                                        2 x 2 tiles sharing a- and b-blocks (a quarter less L2 traffic) */
       TDNA_OPT_PIPELINE_CHUNKS = 7, /* tdna_analyze_host: at most this many row chunks for a pass PIPELINED with the upload (region
                                        launches per chunk; 0 = default = one pass after the upload, which measured faster) */
       TDNA_OPT_SEGMENT_ITEMS = 6,   /* work items (tiles of kernel (a), tile pairs of kernel (b)) per launch of a streaming pass; 0 =
                                        automatic (~30 ms of work).  Between launches a single-GPU pass honours tdna_cancel and calls
                                        the progress callback */
       TDNA_OPT_MULTI = 4            /* 1 (default): whole-range calls are shared by the communicator's ranks; 0: this GPU alone;
                                        2: the collective path even when the communicator has ONE rank (tests on a one-GPU box) */ };
enum { TDNA_KERNEL_AUTO = 0,   /* the faster eligible one (A/B on ncu evidence) */
       TDNA_KERNEL_POPC = 1,   /* (a) bit-planes + LOP3/POPC, any alphabet, any mode */
       TDNA_KERNEL_TENSOR = 2  /* (b) int8 contraction on tcgen05/TMEM: uncorrected p with ambiguity allowed, K2P, or transversions
                                  only; L < 4096.  In p and K2P, pairs that involve IUPAC ambiguity letters are only BOUNDED by it
                                  (enough for the candidate filter) and their exact counts come from the bit-planes; transversions
                                  only is exact for every letter.  TDNA_E_STATE if the configuration is not eligible */ };
enum { TDNA_PATH_AUTO = 0,   /* streaming; falls back to the dense path if the candidate buffer overflows */
       TDNA_PATH_DENSE = 1,  /* materialise row bands of the fp64 matrix, sweep them */
       TDNA_PATH_STREAM = 2  /* streaming only (TDNA_E_TOO_LARGE on overflow) */ };
int32_t tdna_set_option(tdna_ctx *ctx, int32_t key, int64_t value);

/* ---- Pairwise Summary / threshold percentile --------------------------------------------- */
/* All distances of one PairwiseDistribution class as (float)d, -1 dropped, UNSORTED multiset
 * (PairwiseDistribution.java:77-103, 161-203):
 *   INTRA: unordered named conspecific pairs, once per pair (twice when the two full names are equal, :172)
 *   INTER: same genus id, different epithet, once per pair (:191-198)
 * Two-call protocol: with out == NULL only *n_out is written (the count).  */
int32_t tdna_class_distances(tdna_aln *aln, int32_t klass, float *out, int64_t cap, int64_t *n_out);

/* The same class as a histogram (SURVEY.md 8b tdna_class_histogram): *n_out distinct distances, the first cap of them in out (host or
 * device, unsorted), *pushes = the number of distances the reference's list would hold.  O(distinct) memory on device and host. */
int32_t tdna_class_histogram(tdna_aln *aln, int32_t klass, tdna_hist_entry *out, int64_t cap, int64_t *n_out, int64_t *pushes);

/* ---- Cluster ---------------------------------------------------------------------------- */
/* All unordered pairs i<j with 0 <= d <= t (Cluster.java:797-804), unsorted.  Two-call protocol
 * as above (ii == NULL -> count only). */
int32_t tdna_edges_below(tdna_aln *aln, double t, int32_t *ii, int32_t *jj, double *d, int64_t cap,
                         int64_t *n_out);

/* ---- per-row extremes among the congeners ------------------------------------------------------------------
 * ExtremePairwise.run (SpeciesIdentifier/ExtremePairwise.java:99-170) and DistanceAnalysis.run (SpeciesIdentifier/
 * DistanceAnalysis.java:183-230) call sortAgainst for every query and then read only the query's conspecifics and its congeneric,
 * allospecific entries off the sorted list.  One launch evaluates, for every NAMED query row q, its pairs with the rows of the same
 * genus id (both named) and returns what the two modules read, in the well-formed case (the comparator is
 * consistent), plus the evidence the host layer needs to tell that it is (near_*, flags); other rows are re-resolved on the host
 * with the exact sort, as for tdna_row_result.
 *   conspecific: species id equal (>= 0), j != q.   congeneric, allospecific: genus id equal, species ids differ (both >= 0). */
typedef struct {
    double d_max_intra;  /* the LAST valid conspecific of the sorted list (ExtremePairwise.java:149-163): the largest distance; -1.0 if none */
    double d_min_inter;  /* the FIRST valid congeneric, allospecific entry (:121-146; DistanceAnalysis.java:203-219 "closest"); -1.0 if none */
    double sum_inter;    /* the congeneric, allospecific distances added up in sorted order (DistanceAnalysis.java:217, average()) */
    int32_t max_intra, min_inter;       /* their rows (of equal distances: the last / the first in list order); -1 if none */
    int32_t shared_intra, shared_inter; /* getOverlap(query, that row) (ExtremePairwise.java:178, 192) */
    int32_t n_intra, n_inter;           /* valid named entries of each class */
    int32_t near_intra, near_inter;     /* entries of the class with 0 < |d - d_ref| < 2e-5 */
    int32_t near_order;                 /* neighbours in the sorted congeneric entries with 0 < difference < 2e-5 (the sum's order) */
    int32_t flags;                      /* TDNA_ROW_SELF_VALID, TDNA_ROW_NONFINITE, TDNA_EXT_* */
} tdna_extreme_result;                  /* 64 bytes, no padding */
enum { TDNA_EXT_UNNAMED = 16,   /* the query has no species name: nothing else is filled in */
       TDNA_EXT_TOO_MANY = 32  /* the query's genus has more than 4,096 rows: resolve on the host */ };
/* out: n records, host or device.  Needs the labels. */
int32_t tdna_row_extremes(tdna_aln *aln, tdna_extreme_result *out);

/* ---- SequenceMatrix callers ------------------------------------------------------------------------------
 * SequenceMatrix compares the sequences of MANY small per-gene alignments: FindDistances (SequenceMatrix/FindDistances.java:196-246)
 * runs getPairwise on every (sequence of any gene, sequence of any gene) and keeps `from <= d <= to`; DisplayDistancesMode
 * (SequenceMatrix/DisplayDistancesMode.java:255-290, minOverlap forced to 1 :163) runs it on every (taxon's sequence of a gene, the
 * reference taxon's sequence of that gene).  Here all the genes' sequences form ONE ragged alignment (lengths[] per row: every
 * comparison stops at the shorter sequence, Sequence.java:1417-1419), so that either caller is one call: tdna_pairs (explicit pair
 * list) for DisplayDistancesMode, and this for FindDistances --
 * all unordered pairs i < j with lo <= d <= hi, d exactly as Sequence.getPairwise returns it (lo <= -1 admits the pairs below
 * minOverlap; NaN never matches), unsorted.  Two-call protocol as tdna_edges_below (ii == NULL -> count only). */
int32_t tdna_pairs_between(tdna_aln *aln, double lo, double hi, int32_t *ii, int32_t *jj, double *d, int64_t cap, int64_t *n_out);

/* ---- SpeciesDetail ----------------------------------------------------- */
/* has_valid[i] = 1 iff some OTHER row of the same species has getSharedLength >= minOverlap
 * (SpeciesDetail.getSequencesWithValidConspecificsCount, Common/DNA/SpeciesDetail.java:78-95). */
int32_t tdna_valid_conspecifics(tdna_aln *aln, uint8_t *has_valid);

/* ---- progress / cancel (DelayCallback, Common/DelayCallback.java:34-56) --------------------
 * The callback runs on the calling thread: between row bands of a matrix-shaped request, and between the launches a long streaming
 * pass is cut into (done / total in work items of the current launch sequence; a 10^6-row pass reports ~30 times).  tdna_cancel may
 * be called from any thread, or from inside the callback (DelayCallback.delay throwing DelayAbortedException): the call in flight
 * returns TDNA_E_CANCELLED at its next band / segment boundary and the alignment stays usable.  A pass shared by a communicator is
 * enqueued whole: there a cancel takes effect at the next call. */
typedef void (*tdna_progress_fn)(int64_t done, int64_t total, void *user);
int32_t tdna_set_progress(tdna_ctx *ctx, tdna_progress_fn fn, void *user);
int32_t tdna_cancel(tdna_ctx *ctx);

/* ---- measurement helpers (bench.py; not part of the drop-in surface) ---------------------- */
/* Launches only the count+distance kernel over the current row range into the resident matrix
 * `reps` times and returns the mean device time per launch in milliseconds (CUDA events on the
 * context's stream). */
int32_t tdna_time_matrix_kernel(tdna_aln *aln, int32_t reps, float *ms_per_launch);
/* TDNA_KERNEL_POPC / TDNA_KERNEL_TENSOR: which count kernel the alignment's most recent streaming pass ran on (0: none yet). */
/* The integer code count kernel (b) uses for L columns and `dims` symbol dimensions: code[0..3] = a, b, c, d with
 * a*c = -(W-1), a*d + b*c + dims*b*d = W, all operand entries within int8, W the smallest usable power of two above L -- so that a
 * row symbol x and a column symbol y contribute W - (W-1)[x == y] to the int32 accumulator.  Host arithmetic only (no device). */
int32_t tdna_tensor_code(int32_t L, int32_t dims, int32_t *code, uint32_t *W);
int32_t tdna_last_count_kernel(tdna_aln *aln);
/* The operand tables count kernel (b) encodes with for a mode and width (host arithmetic only): per symbol CODE (6 rows of 8 bytes) the
 * row-side and the column-side int8 vector over the *dims K-dimensions of a site, and the accumulator's radix W.  Codes -- one-hot
 * modes (uncorrected p, K2P): 0..3 = A C G T, 4 = '-' where it is a symbol of its own (five dimensions), others contribute nothing;
 * transversions only: 0 purine {A G R}, 1 pyrimidine {C T Y}, 2 '-', 3 any other letter.  row . column = ident + W * mism of the site
 * (transversions only: (shared - transv) + W * transv).  five_dims: uncorrected p with '-' as a fifth symbol. */
int32_t tdna_tensor_tables(int32_t mode, int32_t ambiguous_allowed, int32_t L, int32_t five_dims, int8_t *row_side, int8_t *col_side, int32_t *dims,
                           uint32_t *W);
/* K-dimensions per site of the operands count kernel (b) built for this alignment (5, or 4 when no row has an internal gap / in K2P
 * mode; 6 for transversions only); 0 if they have not been built.  The roofline's algorithmic work per pair is 2 * dims * L int8 operations. */
int32_t tdna_tensor_dims(tdna_aln *aln);
/* Sizes of the most recent streaming pass on this GPU: candidate records emitted (two per pair at most) and pairs count kernel (b)
 * queued for exact resolution (0 on count kernel (a)). */
int32_t tdna_last_pass_stats(tdna_aln *aln, int64_t *candidates, int64_t *queued);
/* How many tdna_analyze_host calls on this context ran their pass pipelined with the upload (Best Match rows on one GPU from host
 * symbols, 16,384 to 131,072 rows, count kernel (b)): region launches per row chunk instead of one pass after the upload. */
int64_t tdna_pipelined_passes(tdna_ctx *ctx);
/* Host wall clock of the context's most recent tdna_analyze_host call, in milliseconds since its entry, at: labels + plan done, buffers
 * ready, every upload chunk enqueued, the first chunk's verdict in, every chunk's kernels enqueued, the upload's final
 * synchronisation, the pass finished, the analysis returned, the call returned (-1: a mark the call did not pass). */
int32_t tdna_host_call_phase_ms(tdna_ctx *ctx, double *ms, int32_t cap, int32_t *n_out);
/* Device time between the phase marks of the context's most recent multi-GPU pass (CUDA events on its stream), in order: seed launch,
 * MIN all-reduce of the per-row minima, thresholds + bulk launch + queue resolution, routing + exchange of records / state / status,
 * merge of the own rows, all-gather of the row records.  *n_out phases are written (at most cap). */
int32_t tdna_multi_phase_ms(tdna_ctx *ctx, float *ms, int32_t cap, int32_t *n_out);
/* Device time (CUDA events on the context's stream) of the most recent count-tile pass -- the tile
 * kernel launch(es) of the last matrix band or streaming pass; waits for it to finish. */
int32_t tdna_last_tile_pass_ms(tdna_ctx *ctx, float *ms);
/* Verification of count kernel (b): (shared, ident) of every pair straight from the tcgen05 kernel, n x n
 * records row-major (host or device); TDNA_E_STATE if the alignment is not eligible for it. */
int32_t tdna_tensor_counts(tdna_aln *aln, tdna_counts *out);
/* Integer-pipe micro-benchmarks: results/s of dependent-free POPC and LOP3 streams over all SMs. */
int32_t tdna_measure_int_peaks(tdna_ctx *ctx, double *popc_per_s, double *lop3_per_s);

#ifdef __cplusplus
}
#endif
#endif /* TAXONDNA_CUDA_H */
