// Host layer above the C ABI: the reference's Java API for the distance path, re-hosted in C++
// (the reference is Java; this image has no JDK -- see INTEGRATION.md for the FFM binding the Java
// layer uses against the same C ABI).  Class and method names, argument meaning and error behaviour
// follow the Java classes cited at each declaration; paths are relative to
// src/main/java/com/ggvaidya/TaxonDNA/ of the reference.
//
// No distance arithmetic happens here: every count and every distance comes from
// libtaxondna_cuda.so.  What stays on the host is what the reference also does on the host around
// the arithmetic: names, list order, Settings' quantisation, the sort of one query's row when the
// streamed summary is not provably equivalent to it, and the report text.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/taxondna_cuda.h"

namespace taxondna {

// Common/DNA/SequenceException.java
struct SequenceException : std::runtime_error {
    using std::runtime_error::runtime_error;
};
// Common/DelayAbortedException.java
struct DelayAbortedException : std::runtime_error {
    DelayAbortedException() : std::runtime_error("aborted") {}
};
// stands for the NullPointerException the reference dies with in a few documented corners
struct NullPointerException : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct EngineError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Common/DelayCallback.java:34-56
struct DelayCallback {
    virtual ~DelayCallback() {}
    virtual void begin() {}
    virtual void delay(int done, int total) { (void)done; (void)total; }  // may throw DelayAbortedException
    virtual void end() {}
    virtual void addWarning(const std::string &) {}
};

// Common/DNA/Settings.java:69-97 (accurateTo = 100000)
struct Settings {
    static int64_t makeLongFromDouble(double d);
    static double makeDoubleFromLong(int64_t i);
    static double roundOff(double d);
    static double percentage(double x, double y);
    static bool identical(double d1, double d2);
};

// Double.toString / Float.toString
std::string javaDouble(double x);
std::string javaFloat(float x);

class SequenceList;
class DistanceEngine;

// Common/DNA/Sequence.java
class Sequence {
  public:
    static constexpr int PDM_UNCORRECTED = 0, PDM_K2P = 1, PDM_TRANS_ONLY = 2;  // :53-55
    // the three statics (:57-66) and their accessors (:86-136)
    static void setMinOverlap(int v);
    static int getMinOverlap();
    static bool areAmbiguousBasesAllowed();
    static void ambiguousBasesAllowed(bool now);
    static int getPairwiseDistanceMethod();
    static void setPairwiseDistanceMethod(int m);
    static void clearPairwiseCache();  // :1606-1611 (no cache exists here; kept for the API)

    Sequence(const std::string &name, const std::string &seq);  // :163-166; throws SequenceException
    struct Normalised {};                                        // tag: `seq` is already in the normalised alphabet (device parser)
    Sequence(const std::string &name, const std::string &seq, Normalised);
    void changeName(const std::string &name);                   // :668-742
    void changeSequence(const std::string &seq);                // :751-853

    const std::string &getFullName() const { return name_; }    // :228
    std::string getDisplayName() const;                          // :236-271
    std::optional<std::string> getSpeciesName() const;           // :446-450
    const std::string &getGenusName() const { return genus_; }
    const std::string &getSpeciesNameOnly() const { return species_; }
    const std::string &getSubspeciesName() const { return subspecies_; }
    const std::string &getFamilyName() const { return family_; }
    std::optional<std::string> getGI() const;
    bool getWarningFlag() const { return warning_; }
    int getLength() const { return (int)seq_.size(); }
    std::string getSequence() const;                             // '_' shown as '-' (:286)
    const std::string &getSequenceRaw() const { return seq_; }   // normalised alphabet
    uint64_t getId() const { return id_; }                       // stands for the UUID (:848-852)
    int compareByDisplayName(const Sequence &o) const;           // :878-887

    // the per-pair family (:1322-1717): each is one call into the CUDA library
    double getPairwise(const Sequence &o) const;
    double getPairwiseNoBuffer(const Sequence &o) const;
    int getSharedLength(const Sequence &o) const;
    int getOverlap(const Sequence &o) const { return getSharedLength(o); }
    int countIdentical(const Sequence &o) const;
    int countTransversions(const Sequence &o) const;
    double getK2PDistance(const Sequence &o) const;
    bool hasMinOverlap(const Sequence &o) const;

    // single-character statics (:927-1303)
    static bool isValid(char ch);
    static bool isAmbiguous(char ch);
    static bool isPurine(char ch);
    static bool isPyrimidine(char ch);
    static int getint(char ch);
    static char getcode(int v);

  private:
    friend class DistanceEngine;
    friend class SequenceList;
    tdna_counts pairCounts(const Sequence &o, double *d, int mode) const;
    std::string name_, genus_, species_, subspecies_, gi_, family_, seq_;
    bool warning_ = false;
    uint64_t id_ = 0;
    // where this sequence currently lives on the device (set by DistanceEngine)
    mutable const DistanceEngine *home_ = nullptr;
    mutable int64_t home_row_ = -1;
    mutable uint64_t home_gen_ = 0;
};
using SequencePtr = std::shared_ptr<Sequence>;

// Common/DNA/SequenceList.java (the parts the distance modules use)
class SequenceList {
  public:
    static constexpr int SORT_UNSORTED = 0, SORT_BYNAME = 1;  // :57-58
    SequenceList() {}
    explicit SequenceList(const std::vector<SequencePtr> &v) : seqs_(v) { modified(); }
    static SequenceList readFasta(const std::string &path);   // formats/FastaFile.java:144-289
    static SequenceList fromFastaText(const std::string &text);
    // the same list through the device parser (tdna_fasta_parse: record structure and alphabet normalisation as byte-parallel
    // kernels; names are parsed here from the header spans).  Files with "[AG]" groups or invalid symbols take the host path, which
    // reports them the reference's way.
    static SequenceList fromFastaTextOnDevice(const std::string &text);
    void add(const SequencePtr &s) { seqs_.push_back(s); modified(); }
    int count() const { return (int)seqs_.size(); }
    const SequencePtr &get(int i) const { return seqs_.at(i); }
    const std::vector<SequencePtr> &sequences() const { return seqs_; }
    int resort(int method);  // :337-380 (stable)
    void lock() {}           // :403-480: the reference serialises analyses; so does the caller here
    void unlock() {}
    void modified();         // :782-787
    // the device-side image of this list in its current order (built lazily, dropped by modified())
    DistanceEngine &engine() const;

  private:
    std::vector<SequencePtr> seqs_;
    int sortedBy_ = SORT_UNSORTED;
    mutable std::shared_ptr<DistanceEngine> engine_;
};

// The B200 piece of the host layer: one tdna_aln bound to one SequenceList order.
class DistanceEngine {
  public:
    explicit DistanceEngine(const SequenceList &list);
    ~DistanceEngine();
    DistanceEngine(const DistanceEngine &) = delete;
    int64_t n() const { return n_; }
    void syncConfig() const;  // pushes the Sequence statics if they changed
    tdna_counts pair(int64_t i, int64_t j, double *d) const;
    void rowDistances(int64_t q, std::vector<double> &d, std::vector<int32_t> *shared) const;
    // a sequence that is not in the list against every row: one launch (tdna_query_distances)
    void queryDistances(const Sequence &query, std::vector<double> &d, std::vector<int32_t> *shared) const;
    std::vector<double> matrix() const;  // n x n, row-major
    std::vector<tdna_row_result> bestMatch() const;
    std::vector<float> classDistances(int klass) const;
    // the class as (distinct fp64 distance, count) entries, unsorted; both classes come from ONE pass, the second request is served
    // from the first one's result
    const std::vector<tdna_hist_entry> &classHistogram(int klass) const;
    std::vector<double> pairs(const std::vector<int32_t> &ii, const std::vector<int32_t> &jj) const;  // one launch for a pair list
    void edgesBelow(double t, std::vector<int32_t> &ii, std::vector<int32_t> &jj, std::vector<double> &d) const;
    // unordered pairs i < j with lo <= getPairwise <= hi (tdna_pairs_between)
    void pairsBetween(double lo, double hi, std::vector<int32_t> &ii, std::vector<int32_t> &jj, std::vector<double> &d) const;
    std::vector<uint8_t> validConspecifics() const;
    std::vector<tdna_extreme_result> rowExtremes() const;  // one launch: tdna_row_extremes
    // consensus sequence (normalised alphabet) of every bin of list positions: one launch (tdna_consensus)
    std::vector<std::string> consensus(const std::vector<std::vector<int>> &bins) const;
    const std::vector<int32_t> &speciesId() const { return species_id_; }
    static tdna_ctx *context();  // process-wide, device from $TDNA_DEVICE (default 0)
    static int64_t launchCount();
    uint64_t generation() const { return gen_; }
    static bool isLive(const DistanceEngine *e, uint64_t gen);  // is `e` still the engine of generation `gen`?

  private:
    tdna_aln *aln_ = nullptr;
    int64_t n_ = 0;
    int64_t width_ = 0;  // columns of the alignment (the longest sequence)
    uint64_t gen_ = 0;
    std::vector<int32_t> species_id_;
    mutable int cfg_mode_ = -1, cfg_overlap_ = -1, cfg_ambig_ = -1;
    int64_t class_cap_[2] = {0, 0};              // label-derived upper bounds of |PD_INTRA|, |PD_INTER|
    mutable std::vector<float> class_cache_[2];  // both classes of the current configuration (one pass serves both)
    mutable bool class_cached_ = false;
    mutable std::vector<tdna_hist_entry> hist_cache_[2];
    mutable bool hist_cached_ = false;
};

// Common/DNA/SortedSequenceList.java:60-162 (+ SortedSequenceComparator :197-261)
class SortedSequenceList {
  public:
    explicit SortedSequenceList(const SequenceList &list) : original_(&list) {}
    void sortAgainst(const SequencePtr &query, DelayCallback *delay);  // exact: TimSort under the reference's comparator
    int count() const { return (int)sorted_.size(); }
    SequencePtr get(int x) const;  // null past the end (:141-144)
    SequencePtr getQuery() const { return query_; }
    double getDistance(int x) const;
    int position(int x) const { return sorted_.at(x); }  // list position of entry x (not in the Java API)
    // getSharedLength(query, entry x): the row kernel produces it with the distances (not in the Java API, which asks
    // Sequence.getOverlap pair by pair)
    int getOverlap(int x) const { return shared_.at(sorted_.at(x)); }

  private:
    const SequenceList *original_;
    SequencePtr query_;
    std::vector<int> sorted_;
    std::vector<double> dist_;
    std::vector<int32_t> shared_;
};

// java.util.TimSort restated (Arrays.sort(Object[], Comparator)); cmp(a, b) as Comparator.compare.
// Throws std::invalid_argument("Comparison method violates its general contract!") like the JDK.
void javaSort(std::vector<int> &a, const std::function<int(int, int)> &cmp);

// Common/DNA/PairwiseDistribution.java:41-402
// A sorted multiset of floats as runs of equal values: what PairwiseDistribution's float[] holds after Arrays.sort, without the
// array (the congeneric class of 10^6 barcodes has 5 * 10^8 distances and a few thousand distinct ones).  Built from the device's
// class histogram (tdna_class_histogram); index k means position k of the reference's sorted array.
class FloatRuns {
  public:
    FloatRuns() = default;
    void push_run(float v, int64_t count) {  // runs must arrive in the order Arrays.sort(float[]) gives
        if (count <= 0) return;
        v_.push_back(v);
        end_.push_back((end_.empty() ? 0 : end_.back()) + count);
    }
    int64_t size() const { return end_.empty() ? 0 : end_.back(); }
    bool empty() const { return end_.empty(); }
    size_t runs() const { return v_.size(); }
    float run_value(size_t r) const { return v_[r]; }
    int64_t run_count(size_t r) const { return end_[r] - (r ? end_[r - 1] : 0); }
    float operator[](int64_t k) const;  // the k-th smallest
    float front() const { return v_.front(); }
    float back() const { return v_.back(); }
    std::vector<float> expand() const;

  private:
    std::vector<float> v_;
    std::vector<int64_t> end_;  // entries up to and including run r
};

class PairwiseDistribution {
  public:
    static constexpr int PD_INTRA = 0, PD_INTER = 1;
    static constexpr char CUMUL_FORWARD = 'F', CUMUL_BACKWARD = 'B';
    PairwiseDistribution(const SequenceList &list, int type, DelayCallback *delay);
    int countSequences() const { return count_sequences_; }
    int countValidComparisons() const { return (int)d_.size(); }
    int getZero() const;
    int getOne() const;
    std::string getDistributionAsString(double from, double to, double interval, char cumulDirection) const;
    int getBetween(double from, double to) const;
    int getBetweenIncl(float from, float to) const;
    float getMaximumDistance() const;
    float getMinimumDistance() const;
    std::vector<float> getDistancesBetween(double from, double to) const;  // the reference's Vector, expanded (small lists)
    FloatRuns runsBetween(double from, double to) const;                   // the same selection as runs
    const FloatRuns &runs() const { return d_; }

  private:
    FloatRuns d_;  // in the order Arrays.sort(float[]) gives (NaN last)
    int count_sequences_ = 0;
};

// Common/DNA/PairwiseDistance.java:30-75 -- one kept pair
struct PairwiseDistance {
    SequencePtr seqA, seqB;
    double distance;
    const SequencePtr &getSequenceA() const { return seqA; }
    const SequencePtr &getSequenceB() const { return seqB; }
    double getDistance() const { return distance; }
    bool isMentioned(const Sequence &s) const;
};

// Common/DNA/PairwiseDistances.java:49-377: like PairwiseDistribution, but the pairs keep their identities, BOTH
// directions of a pair are pushed (the half-table test is commented out at :171) and a per-sequence mean is kept.
class PairwiseDistances {
  public:
    static constexpr int PD_INTRA = 0, PD_INTER = 1;
    PairwiseDistances(const SequenceList &list, int type, DelayCallback *delay);
    int countSequences() const { return count_sequences_; }
    int countValidComparisons() const { return (int)d_.size(); }
    int getZero() const;
    int getOne() const;
    int getBetween(double from, double to) const { return (int)getDistancesBetween(from, to - 0.000001).size(); }
    int getBetweenIncl(double from, double to) const { return (int)getDistancesBetween(from, to).size(); }
    double getMaximumDistance() const { return d_.empty() ? 0.0 : d_.back().distance; }
    double getMinimumDistance() const { return d_.empty() ? 0.0 : d_.front().distance; }
    std::vector<PairwiseDistance> getDistancesBetween(double from, double to) const;
    double getAverageDistance(const std::string &seqName) const;  // -1.0 when the name has no entry
    std::vector<std::string> getAveragedSequences() const;
    const std::vector<PairwiseDistance> &all() const { return d_; }

  private:
    std::vector<PairwiseDistance> d_;  // sorted by PairwiseDistance.compareTo through Collections.sort (TimSort)
    std::vector<std::pair<std::string, double>> averages_;  // Hashtable.put order; later puts of a name win
    int count_sequences_ = 0;
};

// SpeciesIdentifier/BestMatch.java:164-582 without the AWT panel: run() returns text_main's content.
class BestMatch {
  public:
    explicit BestMatch(const SequenceList &list) : list_(&list) {}
    void setThreshold(double percent) { threshold_percent_ = percent; }  // the text field, in %
    std::string run(DelayCallback *delay = nullptr);
    // counters of the last run (BestMatch.java:171-189)
    int best_match_correct = 0, best_match_ambiguous = 0, best_match_incorrect = 0;
    int best_close_match_correct = 0, best_close_match_ambiguous = 0, best_close_match_incorrect = 0, best_close_match_nomatch = 0;
    int count_no_matches = 0, count_zero_percent_matches = 0, count_allo_at_zero = 0, count_seqs_with_valid_conspecific_matches = 0;
    int count_sequences_without_species_names = 0;
    int rows_resolved_on_host = 0;  // queries whose streamed summary was not provably the sort's outcome
    bool force_exact = false;       // resolve every query with the exact sort (validation switch)

  private:
    const SequenceList *list_;
    double threshold_percent_ = 3.0;  // "03.000" (BestMatch.java:62)
};

// SpeciesIdentifier/BlockAnalysis.java:179-387 ("All Species Barcodes")
class BlockAnalysis {
  public:
    explicit BlockAnalysis(const SequenceList &list) : list_(&list) {}
    void setThreshold(double percent) { threshold_percent_ = percent; }
    std::string run(DelayCallback *delay = nullptr);
    int block_correct = 0, block_incorrect = 0, block_ambiguous = 0, block_nomatch = 0, block_noseq = 0, count_seq_without_species_name = 0;
    int rows_resolved_on_host = 0;
    bool force_exact = false;

  private:
    const SequenceList *list_;
    double threshold_percent_ = 3.0;
};

// SpeciesIdentifier/PairwiseSummary.java:109-485
class PairwiseSummary {
  public:
    explicit PairwiseSummary(const SequenceList &list) : list_(&list) {}
    std::string run(DelayCallback *delay = nullptr);
    float getFivePercentCutoff() const { return five_; }  // :53-55
    std::shared_ptr<PairwiseDistribution> getPD_intra() const { return intra_; }
    std::shared_ptr<PairwiseDistribution> getPD_inter() const { return inter_; }

  private:
    const SequenceList *list_;
    float five_ = 0;
    std::shared_ptr<PairwiseDistribution> intra_, inter_;
};

// SpeciesIdentifier/Cluster.java:741-867 (the clustering) and :188-260 (per-cluster statistics)
class Cluster {
  public:
    struct Stat {
        int size;
        double largest_pairwise;
        int valid_comparisons, valid_comparisons_over;
        double percentage;
    };
    explicit Cluster(const SequenceList &list) : list_(&list) {}
    void setThreshold(double percent) { max_pairwise_ = percent / 100.0; }
    void run(DelayCallback *delay = nullptr);
    const std::vector<std::vector<int>> &clusters() const { return clusters_; }  // list positions, reference order
    // Cluster.makeConsensusOfBin (:123-138) for every cluster: the raw (normalised-alphabet) consensus sequences
    std::vector<std::string> consensusSequences() const;
    const std::vector<Stat> &stats() const { return stats_; }

  private:
    const SequenceList *list_;
    double max_pairwise_ = 0.03;
    std::vector<std::vector<int>> clusters_;
    std::vector<Stat> stats_;
};

// SpeciesIdentifier/ExtremePairwise.java:72-210: per sequence, the largest conspecific and the smallest congeneric,
// allospecific distance.  run() returns text_matches' content.
class ExtremePairwise {
  public:
    struct Row {
        int largest_intra = -1, smallest_inter = -1;  // list positions; -1 = none; both -2 = no species name
    };
    explicit ExtremePairwise(const SequenceList &list) : list_(&list) {}
    std::string run(DelayCallback *delay = nullptr);
    const std::vector<Row> &rows() const { return rows_; }
    int rows_resolved_on_host = 0;  // queries whose device record was not provably the sort's outcome
    bool force_exact = false;       // resolve every query with the exact sort (validation switch)

  private:
    const SequenceList *list_;
    std::vector<Row> rows_;
};

// SpeciesIdentifier/QuerySequence.java:113-180: a pasted sequence against the list
class QuerySequence {
  public:
    explicit QuerySequence(const SequenceList &list) : list_(&list) {}
    // the entries of list_result, in order; throws SequenceException for a bad sequence (:150-155 reports it)
    std::vector<std::string> query(const std::string &string_sequence, DelayCallback *delay = nullptr);
    const std::vector<int> &positions() const { return positions_; }  // list position of every entry (map_index)
    const std::vector<double> &distances() const { return distances_; }

  private:
    const SequenceList *list_;
    std::vector<int> positions_;
    std::vector<double> distances_;
};

// SpeciesIdentifier/PairwiseExplorer.java:109-190: the pairs of a class in 0.5 % categories
class PairwiseExplorer {
  public:
    struct Category {
        std::string key;
        std::vector<PairwiseDistance> distances;
    };
    explicit PairwiseExplorer(const SequenceList &list) : list_(&list) {}
    std::vector<Category> run(int type, DelayCallback *delay = nullptr);  // PairwiseDistances::PD_INTRA / PD_INTER

  private:
    const SequenceList *list_;
    std::shared_ptr<PairwiseDistances> intra_, inter_;  // kept between runs like the reference's fields (:125-152)
};

// SpeciesIdentifier/DistanceAnalysis.java:161-320: per-genus means of the smallest / mean congeneric, allospecific distance
class DistanceAnalysis {
  public:
    explicit DistanceAnalysis(const SequenceList &list) : list_(&list) {}
    std::string run(DelayCallback *delay = nullptr);
    int rows_resolved_on_host = 0;
    bool force_exact = false;

  private:
    const SequenceList *list_;
};

// ---- SequenceMatrix's two distance callers: many small per-gene alignments in one launch --------------
// The (taxon, charset) -> Sequence grid they read through TableManager (SequenceMatrix/TableManager.java:254-371): charsets and
// taxa in display order; a missing cell is a taxon without a sequence for that gene.
class SequenceGrid {
  public:
    SequenceGrid(const std::vector<std::string> &charsets, const std::vector<std::string> &taxa) : charsets_(charsets), taxa_(taxa) {}
    void setSequence(const std::string &colName, const std::string &seqName, const SequencePtr &seq);
    void setCancelled(const std::string &colName, const std::string &seqName);  // TableManager.isSequenceCancelled
    const std::vector<std::string> &getCharsets() const { return charsets_; }
    const std::vector<std::string> &getSequenceNames() const { return taxa_; }
    SequencePtr getSequence(const std::string &colName, const std::string &seqName) const;  // null when absent
    bool isSequenceCancelled(const std::string &colName, const std::string &seqName) const;
    // TableManager.getSequenceListByColumn :359-371 (taxa in display order, absent cells skipped)
    std::vector<std::pair<std::string, SequencePtr>> getSequenceListByColumn(const std::string &colName) const;

  private:
    std::vector<std::string> charsets_, taxa_;
    std::vector<std::pair<std::pair<std::string, std::string>, SequencePtr>> cells_;
    std::vector<std::pair<std::string, std::string>> cancelled_;
};

// SequenceMatrix/FindDistances.java:123-289: every sequence of the chosen genes against every other one, the pairs with
// from <= getPairwise <= to in both directions, sorted by PairwiseDistance.compareTo.  All genes' sequences form one ragged
// alignment on the device; the O(N^2) loop of :200-244 is one tdna_pairs_between call.
class FindDistances {
  public:
    struct Result {
        int outer, inner;  // rows of the (column, taxon) enumeration
        double distance;
    };
    explicit FindDistances(const SequenceGrid &grid) : grid_(&grid) {}
    // from / to as fractions (the text fields hold percentages, :136-164); colName empty = "All genes".  Returns text_main's content.
    std::string displayResults(double from, double to, const std::string &colName = std::string(), DelayCallback *delay = nullptr);
    const std::vector<Result> &results() const { return results_; }

  private:
    const SequenceGrid *grid_;
    std::vector<Result> results_;
};

// SequenceMatrix/DisplayDistancesMode.java:163-398: per gene, every taxon's sequence against the reference taxon's; normalised per
// gene, taxa sorted by their mean normalised distance over the other genes, per-gene ranks.  The distances of all genes are one
// tdna_pairs launch.
class DisplayDistancesMode {
  public:
    static constexpr double DIST_ILLEGAL = -32.0, DIST_NO_COMPARE_SEQ = -64.0, DIST_SEQ_NA = -128.0, DIST_NO_OVERLAP = -256.0,
                            DIST_SEQ_ON_TOP = -1024.0, DIST_CANCELLED = -2048.0;  // :50-55
    explicit DisplayDistancesMode(const SequenceGrid &grid) : grid_(&grid) {}
    void setDistanceMethod(int pdm) { currentDistanceMethod_ = pdm; }  // the menu (:60-63; "Transversions only" is the default)
    void activateDisplay(const std::string &selected_colName);         // :157-171: minOverlap := 1, the menu's distance method
    void deactivateDisplay();                                          // :173-179: both statics restored
    // :217-398; returns the taxa in sorted order
    std::vector<std::string> getSortedSequences(const std::string &referenceTaxon);
    const std::vector<std::string> &columns() const { return columns_; }
    const std::vector<std::vector<double>> &distances() const { return distances_; }       // [column][taxon]
    const std::vector<std::vector<double>> &normDistances() const { return norm_distances_; }
    const std::vector<std::vector<int>> &ranks() const { return ranks_; }
    const std::vector<double> &scores() const { return scores_; }

  private:
    const SequenceGrid *grid_;
    int currentDistanceMethod_ = Sequence::PDM_TRANS_ONLY, oldOverlap_ = 0, oldDistanceMode_ = 0;
    std::string selected_colName_;
    std::vector<std::string> columns_;
    std::vector<std::vector<double>> distances_, norm_distances_;
    std::vector<std::vector<int>> ranks_;
    std::vector<double> scores_;
};

// Common/DNA/SpeciesDetails.java / SpeciesDetail.java:78-95
struct SpeciesDetails {
    explicit SpeciesDetails(const SequenceList &list);
    // name -> (sequence count, sequences with a valid conspecific)
    std::vector<std::string> names;
    std::vector<int> sequences_count, valid_conspecifics_count;
    int indexOf(const std::string &name) const;
};

}  // namespace taxondna
