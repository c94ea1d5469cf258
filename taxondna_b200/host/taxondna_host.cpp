// Implementation of the host layer (see taxondna_host.hpp).  Reference lines are cited per function.
#include "taxondna_host.hpp"

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <regex>
#include <sstream>
#include <unordered_map>

namespace taxondna {

// ------------------------------------------------------------------------------------------------
// Settings (Common/DNA/Settings.java:69-97)
// ------------------------------------------------------------------------------------------------
static int64_t javaD2L(double d) {  // JLS 5.1.3 narrowing: NaN -> 0, saturating
    if (d != d) return 0;
    if (d >= 9223372036854775807.0) return INT64_MAX;
    if (d <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)d;
}
int64_t Settings::makeLongFromDouble(double d) { return javaD2L(d * 100000); }
double Settings::makeDoubleFromLong(int64_t i) { return (double)i / (double)100000; }
double Settings::roundOff(double d) { return makeDoubleFromLong(makeLongFromDouble(d)); }
double Settings::percentage(double x, double y) {
    if (y == 0) return 0;
    return (double)(javaD2L(roundOff(x / y) * 100 * 100)) / 100;
}
bool Settings::identical(double d1, double d2) { return makeLongFromDouble(d1) == makeLongFromDouble(d2); }

// ------------------------------------------------------------------------------------------------
// Double.toString / Float.toString: shortest digits that round-trip (JDK >= 19), Java's layout
// ------------------------------------------------------------------------------------------------
template <class T> static std::string javaFloating(T x) {
    if (x != x) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    std::string sign = std::signbit(x) ? "-" : "";
    T ax = std::fabs(x);
    if (ax == 0) return sign + "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, ax, std::chars_format::scientific);
    std::string s(buf, r.ptr);  // d.ddddde[+-]XX, shortest
    size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    int exp10 = std::atoi(s.c_str() + e + 1);
    std::string digits;
    for (char ch : mant)
        if (ch != '.') digits.push_back(ch);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    int pointPos = exp10 + 1;  // value = 0.DIGITS * 10^pointPos
    if (ax >= (T)1e-3 && ax < (T)1e7) {
        std::string body;
        if (pointPos <= 0) body = "0." + std::string(-pointPos, '0') + digits;
        else if (pointPos >= (int)digits.size()) body = digits + std::string(pointPos - digits.size(), '0') + ".0";
        else body = digits.substr(0, pointPos) + "." + digits.substr(pointPos);
        return sign + body;
    }
    std::string m = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0");
    return sign + m + "E" + std::to_string(exp10);
}
std::string javaDouble(double x) { return javaFloating<double>(x); }
std::string javaFloat(float x) { return javaFloating<float>(x); }

// ------------------------------------------------------------------------------------------------
// Sequence: statics
// ------------------------------------------------------------------------------------------------
static std::atomic<int> g_method{Sequence::PDM_UNCORRECTED}, g_minOverlap{300};
static std::atomic<bool> g_ambig{true};
static std::atomic<uint64_t> g_nextId{1};

void Sequence::setMinOverlap(int v) { g_minOverlap = v; }
int Sequence::getMinOverlap() { return g_minOverlap; }
bool Sequence::areAmbiguousBasesAllowed() { return g_ambig; }
void Sequence::ambiguousBasesAllowed(bool now) { g_ambig = now; }
int Sequence::getPairwiseDistanceMethod() { return g_method; }
void Sequence::setPairwiseDistanceMethod(int m) { g_method = m; }
void Sequence::clearPairwiseCache() {}

// ------------------------------------------------------------------------------------------------
// Sequence: characters (Sequence.java:927-1237)
// ------------------------------------------------------------------------------------------------
bool Sequence::isValid(char ch) {
    if (ch > 'a' && ch < 'z') ch = (char)(ch - ('a' - 'A'));
    return std::strchr("ACTGRYKMSNBDHVW-?_", ch) != nullptr && ch != 0;
}
bool Sequence::isAmbiguous(char ch) {
    if (ch > 'a' && ch < 'z') ch = (char)(ch - ('a' - 'A'));
    return std::strchr("RYKMSNBDHVW", ch) != nullptr && ch != 0;
}
bool Sequence::isPurine(char ch) { return ch == 'R' || ch == 'A' || ch == 'G'; }
bool Sequence::isPyrimidine(char ch) { return ch == 'Y' || ch == 'T' || ch == 'C'; }
int Sequence::getint(char ch) {
    if (ch >= 'a' && ch <= 'z') ch = (char)(ch - ('a' - 'A'));
    int r = 0;
    if (std::strchr("ARMNDHVW", ch) && ch) r |= 1;
    if (std::strchr("CYMSNBHV", ch) && ch) r |= 2;
    if (std::strchr("TYKNBDHW", ch) && ch) r |= 4;
    if (std::strchr("GRKSNBDV", ch) && ch) r |= 8;
    return r;
}
char Sequence::getcode(int v) {
    static const char tbl[17] = "-ACMTWYHGRSVKDBN";
    return tbl[v & 15];
}

// ------------------------------------------------------------------------------------------------
// Sequence: names (Sequence.java:668-742, 236-271, 446-471)
// ------------------------------------------------------------------------------------------------
static std::string trim(const std::string &s) {  // String.trim(): chars <= ' '
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') a++;
    while (b > a && (unsigned char)s[b - 1] <= ' ') b--;
    return s.substr(a, b - a);
}

Sequence::Sequence(const std::string &name, const std::string &seq) {
    changeName(name);
    std::string up = seq;
    for (auto &ch : up)
        if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
    changeSequence(trim(up));
}

Sequence::Sequence(const std::string &name, const std::string &seq, Normalised) {
    changeName(name);
    seq_ = seq;
    id_ = g_nextId++;
}

void Sequence::changeName(const std::string &nameIn) {
    static const std::regex p3("([A-Z][a-z]+) ([a-z]+) ([a-z]+)\\b"), p2("([A-Z][a-z]+) ([a-z]+)\\b"),
        psp("([A-Z][a-z]+) ([a-z]+)\\.\\b"), pgi("gi\\|([0-9]+)[\\|:]"), pfam("\\(family:\\s*([A-Za-z]+)\\s*\\)");
    std::string name = nameIn;
    std::replace(name.begin(), name.end(), '\n', ' ');
    name = trim(name);
    name_ = name;
    genus_ = species_ = subspecies_ = gi_ = family_ = "";
    warning_ = false;
    std::smatch m;
    if (std::regex_search(name, m, p3)) {
        genus_ = m[1];
        species_ = m[2];
        subspecies_ = m[3];
    } else if (std::regex_search(name, m, p2)) {
        genus_ = m[1];
        species_ = m[2];
    } else if (std::regex_search(name, m, psp)) {
        genus_ = m[1];
        species_ = m[2];
        warning_ = true;
    }
    if (genus_.empty()) warning_ = true;
    if (std::regex_search(name, m, pgi)) gi_ = m[1];
    if (std::regex_search(name, m, pfam)) family_ = m[1];
}

void Sequence::changeSequence(const std::string &in) {
    std::string seq = in;
    for (auto &ch : seq)
        if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
    std::string buf;
    for (size_t x = 0; x < seq.size(); x++) {
        char ch = seq[x];
        if (ch == '(' || ch == '[') {  // :761-796
            int single = 0;
            for (size_t y = x + 1; y < seq.size(); y++, x++) {
                ch = seq[y];
                if (ch == '(' || ch == '[')
                    throw SequenceException(name_ + ": Character '" + ch + "' found unexpectedly while inside a ambiguous base");
                else if (ch == ')' || ch == ']') break;
                else {
                    if (!isValid(ch)) throw SequenceException(name_ + ": Invalid character '" + ch + "' found in sequence.");
                    single |= getint(ch);
                }
            }
            x++;
            buf.push_back(getcode(single));
        } else buf.push_back(ch);
    }
    int stopped = 0;
    for (size_t x = 0; x < buf.size(); x++) {  // :806-823
        if (buf[x] == '_') throw SequenceException(name_ + ": You may not use '_' in a sequence (index " + std::to_string(x) + ")");
        else if (buf[x] == '-') { stopped = (int)x; buf[x] = '_'; }
        else if (buf[x] == '?') continue;
        else break;
    }
    for (int x = (int)buf.size() - 1; x >= 0; x--) {  // :825-840
        if (buf[x] == '_') {
            if (stopped == x) break;
            throw SequenceException(name_ + ": You may not use '_' in a sequence (index " + std::to_string(x) + ")");
        } else if (buf[x] == '-') buf[x] = '_';
        else if (buf[x] == '?') continue;
        else break;
    }
    for (size_t x = 0; x < buf.size(); x++) {  // doSanityChecks :1727-1763
        if (buf[x] == 'U' || buf[x] == 'u') throw SequenceException(name_ + ": Uracil at index " + std::to_string(x));
        if (!isValid(buf[x])) throw SequenceException(name_ + ": An illegal base '" + buf[x] + "' at index " + std::to_string(x));
    }
    seq_ = buf;
    id_ = g_nextId++;  // a new UUID (:848-852): any device binding is stale now
    home_ = nullptr;
    home_row_ = -1;
}

std::optional<std::string> Sequence::getSpeciesName() const {
    if (species_.empty() || genus_.empty()) return std::nullopt;
    return genus_ + " " + species_;
}
std::optional<std::string> Sequence::getGI() const {
    if (gi_.empty()) return std::nullopt;
    return gi_;
}
std::string Sequence::getDisplayName() const {
    if (warning_) return "{" + name_.substr(0, std::min<size_t>(name_.size(), 80)) + "}";
    if (!getGI()) return name_;
    if (auto sp = getSpeciesName()) {
        std::string n = *sp;
        if (!subspecies_.empty()) n += " (" + subspecies_ + ")";
        if (!family_.empty()) n += " (Family " + family_ + ")";
        n += " (gi:" + gi_ + ")";
        return n;
    }
    return name_;
}
std::string Sequence::getSequence() const {
    std::string s = seq_;
    std::replace(s.begin(), s.end(), '_', '-');
    return s;
}
static int compareToIgnoreCase(const std::string &a, const std::string &b) {  // String.compareToIgnoreCase (ASCII)
    size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; i++) {
        unsigned char ca = a[i], cb = b[i];
        if (ca != cb) {
            int ua = std::toupper(ca), ub = std::toupper(cb);
            if (ua != ub) {
                int la = std::tolower(ua), lb = std::tolower(ub);
                if (la != lb) return la - lb;
            }
        }
    }
    return (int)a.size() - (int)b.size();
}
int Sequence::compareByDisplayName(const Sequence &o) const {
    if (o.warning_ && !warning_) return -1;
    if (!o.warning_ && warning_) return 1;
    return compareToIgnoreCase(getDisplayName(), o.getDisplayName());
}

// ------------------------------------------------------------------------------------------------
// the CUDA library
// ------------------------------------------------------------------------------------------------
static void ck(int rc) {
    if (rc != TDNA_OK) {
        if (rc == TDNA_E_CANCELLED) throw DelayAbortedException();
        throw EngineError(std::string("taxondna_cuda: ") + tdna_last_error());
    }
}
static tdna_ctx *g_ctx = nullptr;
static std::atomic<uint64_t> g_gen{1};
// live engines by generation: a Sequence remembers the engine it was last uploaded with; the engine may be gone by the time it asks
static std::mutex g_live_mutex;
static std::unordered_map<uint64_t, const DistanceEngine *> g_live;
bool DistanceEngine::isLive(const DistanceEngine *e, uint64_t gen) {
    std::lock_guard<std::mutex> lock(g_live_mutex);
    auto it = g_live.find(gen);
    return it != g_live.end() && it->second == e;
}
tdna_ctx *DistanceEngine::context() {
    if (!g_ctx) {
        const char *dev = std::getenv("TDNA_DEVICE");
        ck(tdna_init(dev ? std::atoi(dev) : 0, &g_ctx));  // throws without a GPU: there is no CPU path
    }
    return g_ctx;
}
int64_t DistanceEngine::launchCount() { return g_ctx ? tdna_launch_count(g_ctx) : 0; }

// builds the normalised, '?'-padded symbol matrix of a set of sequences
static void packSymbols(const std::vector<const Sequence *> &seqs, std::vector<uint8_t> &sym, std::vector<int32_t> &len, int &L) {
    L = 1;
    for (auto s : seqs) L = std::max(L, s->getLength());
    sym.assign((size_t)seqs.size() * L, (uint8_t)'?');
    len.resize(seqs.size());
    for (size_t i = 0; i < seqs.size(); i++) {
        const std::string &r = seqs[i]->getSequenceRaw();
        std::memcpy(&sym[i * (size_t)L], r.data(), r.size());
        len[i] = (int32_t)r.size();
    }
}

DistanceEngine::DistanceEngine(const SequenceList &list) {
    n_ = list.count();
    gen_ = g_gen++;
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live[gen_] = this; }
    if (n_ == 0) return;
    std::vector<const Sequence *> ptrs;
    for (auto &s : list.sequences()) ptrs.push_back(s.get());
    std::vector<uint8_t> sym;
    std::vector<int32_t> len;
    int L;
    packSymbols(ptrs, sym, len, L);
    ck(tdna_alignment_create(context(), sym.data(), n_, L, L, len.data(), &aln_));
    width_ = L;
    // labels: ids of species / genus strings, ranks of epithets and full names under String.compareTo
    std::map<std::string, int> spIds, genIds;
    std::vector<int32_t> genus(n_), epi(n_), full(n_);
    species_id_.resize(n_);
    std::vector<std::string> epis, fulls;
    for (auto s : ptrs) {
        epis.push_back(s->getSpeciesNameOnly());
        fulls.push_back(s->getFullName());
    }
    auto ranks = [](std::vector<std::string> v) {
        std::vector<std::string> u = v;
        std::sort(u.begin(), u.end());  // byte order == String.compareTo for ASCII / Latin-1 names
        u.erase(std::unique(u.begin(), u.end()), u.end());
        std::vector<int32_t> r(v.size());
        for (size_t i = 0; i < v.size(); i++) r[i] = (int32_t)(std::lower_bound(u.begin(), u.end(), v[i]) - u.begin());
        return r;
    };
    epi = ranks(epis);
    full = ranks(fulls);
    for (int64_t i = 0; i < n_; i++) {
        auto sp = ptrs[i]->getSpeciesName();
        species_id_[i] = sp ? spIds.emplace(*sp, (int)spIds.size()).first->second : -1;
        genus[i] = genIds.emplace(ptrs[i]->getGenusName(), (int)genIds.size()).first->second;
        ptrs[i]->home_ = this;
        ptrs[i]->home_row_ = i;
        ptrs[i]->home_gen_ = gen_;
    }
    ck(tdna_alignment_set_labels(aln_, species_id_.data(), genus.data(), epi.data(), full.data()));
    // upper bounds of the two PairwiseDistribution classes from the labels alone (ordered conspecific pairs cover the
    // double push of equal full names; all congeneric pairs cover the inter class): one pass then needs no sizing call
    std::map<int, int64_t> perSpecies, perGenus;
    for (int64_t i = 0; i < n_; i++) {
        if (species_id_[i] >= 0) perSpecies[species_id_[i]]++;
        perGenus[genus[i]]++;
    }
    class_cap_[0] = class_cap_[1] = 0;
    for (auto &kv : perSpecies) class_cap_[0] += kv.second * (kv.second - 1);
    for (auto &kv : perGenus) class_cap_[1] += kv.second * (kv.second - 1) / 2;
}
DistanceEngine::~DistanceEngine() {
    { std::lock_guard<std::mutex> lock(g_live_mutex); g_live.erase(gen_); }
    if (aln_) tdna_alignment_destroy(aln_);
}
void DistanceEngine::syncConfig() const {
    int m = g_method, o = g_minOverlap, a = g_ambig ? 1 : 0;
    if (m != cfg_mode_ || o != cfg_overlap_ || a != cfg_ambig_) {
        ck(tdna_set_config(aln_, m, o, a));
        cfg_mode_ = m;
        cfg_overlap_ = o;
        cfg_ambig_ = a;
        class_cached_ = false;
        hist_cached_ = false;
    }
}
tdna_counts DistanceEngine::pair(int64_t i, int64_t j, double *d) const {
    syncConfig();
    tdna_counts c;
    ck(tdna_pair(aln_, i, j, &c, d));
    return c;
}
void DistanceEngine::rowDistances(int64_t q, std::vector<double> &d, std::vector<int32_t> *shared) const {
    syncConfig();
    d.resize(n_);
    if (shared) shared->resize(n_);
    ck(tdna_row_distances(aln_, q, d.data(), shared ? shared->data() : nullptr));
}
void DistanceEngine::queryDistances(const Sequence &query, std::vector<double> &d, std::vector<int32_t> *shared) const {
    syncConfig();
    d.resize(n_);
    if (shared) shared->resize(n_);
    const std::string &raw = query.getSequenceRaw();
    ck(tdna_query_distances(aln_, reinterpret_cast<const uint8_t *>(raw.data()), (int32_t)raw.size(), d.data(), shared ? shared->data() : nullptr, nullptr));
}
std::vector<std::string> DistanceEngine::consensus(const std::vector<std::vector<int>> &bins) const {
    std::vector<std::string> out;
    if (bins.empty()) return out;
    std::vector<int32_t> members;
    std::vector<int64_t> offsets(1, 0);
    for (auto &b : bins) {
        for (int m : b) members.push_back(m);
        offsets.push_back((int64_t)members.size());
    }
    std::vector<uint8_t> buf((size_t)bins.size() * width_);
    std::vector<int32_t> lens(bins.size());
    ck(tdna_consensus(aln_, members.data(), offsets.data(), (int64_t)bins.size(), buf.data(), width_, lens.data()));
    for (size_t k = 0; k < bins.size(); k++) out.emplace_back(reinterpret_cast<const char *>(buf.data() + k * width_), (size_t)lens[k]);
    return out;
}
std::vector<double> DistanceEngine::matrix() const {
    syncConfig();
    std::vector<double> d((size_t)n_ * n_);
    ck(tdna_matrix(aln_, d.data(), nullptr));
    return d;
}
std::vector<tdna_row_result> DistanceEngine::bestMatch() const {
    syncConfig();
    std::vector<tdna_row_result> r(n_);
    ck(tdna_best_match(aln_, r.data()));
    return r;
}
// PairwiseSummary builds PD_INTRA and then PD_INTER (PairwiseSummary.java:131-151): both come from ONE pass over the
// pair space (tdna_analyze), sized from the labels, and the second request is served from the first one's result.
std::vector<float> DistanceEngine::classDistances(int klass) const {
    syncConfig();
    if (!class_cached_) {
        tdna_analysis io = {};
        io.want = TDNA_WANT_INTRA | TDNA_WANT_INTER;
        class_cache_[0].assign((size_t)std::max<int64_t>(class_cap_[0], 1), 0.f);
        class_cache_[1].assign((size_t)std::max<int64_t>(class_cap_[1], 1), 0.f);
        io.intra = class_cache_[0].data();
        io.intra_cap = class_cap_[0];
        io.inter = class_cache_[1].data();
        io.inter_cap = class_cap_[1];
        ck(tdna_analyze(aln_, &io));
        if (io.n_intra > class_cap_[0] || io.n_inter > class_cap_[1]) throw EngineError("class distances exceed their label-derived bound");
        class_cache_[0].resize((size_t)io.n_intra);
        class_cache_[1].resize((size_t)io.n_inter);
        class_cached_ = true;
    }
    return class_cache_[klass == TDNA_PD_INTRA ? 0 : 1];
}
const std::vector<tdna_hist_entry> &DistanceEngine::classHistogram(int klass) const {
    syncConfig();
    if (!hist_cached_) {
        int64_t cap = 1 << 16;
        for (;;) {
            tdna_analysis io = {};
            io.want = TDNA_WANT_HIST_INTRA | TDNA_WANT_HIST_INTER;
            hist_cache_[0].assign((size_t)cap, tdna_hist_entry{0.0, 0});
            hist_cache_[1].assign((size_t)cap, tdna_hist_entry{0.0, 0});
            io.hist_intra = hist_cache_[0].data();
            io.hist_inter = hist_cache_[1].data();
            io.hist_intra_cap = io.hist_inter_cap = cap;
            ck(tdna_analyze(aln_, &io));
            if (io.n_hist_intra <= cap && io.n_hist_inter <= cap) {
                hist_cache_[0].resize((size_t)io.n_hist_intra);
                hist_cache_[1].resize((size_t)io.n_hist_inter);
                break;
            }
            cap = std::max(io.n_hist_intra, io.n_hist_inter);
        }
        hist_cached_ = true;
    }
    return hist_cache_[klass == TDNA_PD_INTRA ? 0 : 1];
}
std::vector<double> DistanceEngine::pairs(const std::vector<int32_t> &ii, const std::vector<int32_t> &jj) const {
    syncConfig();
    std::vector<double> d(ii.size());
    if (!ii.empty()) ck(tdna_pairs(aln_, ii.data(), jj.data(), (int64_t)ii.size(), nullptr, d.data()));
    return d;
}
void DistanceEngine::edgesBelow(double t, std::vector<int32_t> &ii, std::vector<int32_t> &jj, std::vector<double> &d) const {
    syncConfig();
    int64_t cnt = 0;
    ck(tdna_edges_below(aln_, t, nullptr, nullptr, nullptr, 0, &cnt));
    ii.resize((size_t)std::max<int64_t>(cnt, 1));
    jj.resize(ii.size());
    d.resize(ii.size());
    int64_t got = 0;
    ck(tdna_edges_below(aln_, t, ii.data(), jj.data(), d.data(), cnt, &got));
    ii.resize((size_t)cnt);
    jj.resize((size_t)cnt);
    d.resize((size_t)cnt);
}
void DistanceEngine::pairsBetween(double lo, double hi, std::vector<int32_t> &ii, std::vector<int32_t> &jj, std::vector<double> &d) const {
    syncConfig();
    int64_t cnt = 0;
    ck(tdna_pairs_between(aln_, lo, hi, nullptr, nullptr, nullptr, 0, &cnt));
    ii.resize((size_t)std::max<int64_t>(cnt, 1));
    jj.resize(ii.size());
    d.resize(ii.size());
    int64_t got = 0;
    ck(tdna_pairs_between(aln_, lo, hi, ii.data(), jj.data(), d.data(), cnt, &got));
    ii.resize((size_t)cnt);
    jj.resize((size_t)cnt);
    d.resize((size_t)cnt);
}
std::vector<tdna_extreme_result> DistanceEngine::rowExtremes() const {
    syncConfig();
    std::vector<tdna_extreme_result> v((size_t)n_);
    if (n_ > 0) ck(tdna_row_extremes(aln_, v.data()));
    return v;
}
std::vector<uint8_t> DistanceEngine::validConspecifics() const {
    syncConfig();
    std::vector<uint8_t> v(n_);
    ck(tdna_valid_conspecifics(aln_, v.data()));
    return v;
}

// One pair.  If both sequences sit in the same live device alignment it is one tdna_pair call on it;
// free-standing sequences (the reference's own KATs build two and compare them, Sequence.java:1913-1955)
// get a transient two-row alignment.  Either way the arithmetic is the CUDA library's.
tdna_counts Sequence::pairCounts(const Sequence &o, double *d, int mode) const {
    if (home_ && home_ == o.home_ && o.home_gen_ == home_gen_ && mode == g_method && DistanceEngine::isLive(home_, home_gen_))
        return home_->pair(home_row_, o.home_row_, d);
    std::vector<uint8_t> sym;
    std::vector<int32_t> len;
    int L;
    packSymbols({this, &o}, sym, len, L);
    tdna_aln *a = nullptr;
    ck(tdna_alignment_create(DistanceEngine::context(), sym.data(), 2, L, L, len.data(), &a));
    tdna_counts c;
    int rc = tdna_set_config(a, mode, g_minOverlap, g_ambig ? 1 : 0);
    if (rc == TDNA_OK) rc = tdna_pair(a, 0, 1, &c, d);
    tdna_alignment_destroy(a);
    ck(rc);
    return c;
}
double Sequence::getPairwiseNoBuffer(const Sequence &o) const {
    double d;
    pairCounts(o, &d, g_method);
    return d;
}
// The cache (:1587-1672) is gone: whole analyses are single calls on the list's engine.  A caller that LOOPS over free-standing
// sequences pays an upload and a launch per pair here (~100 us, two orders of magnitude above a Hashtable hit): put them in a
// SequenceList and use its engine (pairs / rowDistances) instead.
double Sequence::getPairwise(const Sequence &o) const { return getPairwiseNoBuffer(o); }
int Sequence::getSharedLength(const Sequence &o) const { return pairCounts(o, nullptr, g_method).shared; }
int Sequence::countIdentical(const Sequence &o) const { return pairCounts(o, nullptr, PDM_UNCORRECTED).c1; }
int Sequence::countTransversions(const Sequence &o) const { return pairCounts(o, nullptr, PDM_TRANS_ONLY).c1; }
double Sequence::getK2PDistance(const Sequence &o) const {
    // getK2PDistance itself has no overlap gate (:1498-1573): evaluate with minOverlap 0
    std::vector<uint8_t> sym;
    std::vector<int32_t> len;
    int L;
    packSymbols({this, &o}, sym, len, L);
    tdna_aln *a = nullptr;
    ck(tdna_alignment_create(DistanceEngine::context(), sym.data(), 2, L, L, len.data(), &a));
    double d = 0;
    int rc = tdna_set_config(a, PDM_K2P, 0, g_ambig ? 1 : 0);
    if (rc == TDNA_OK) rc = tdna_pair(a, 0, 1, nullptr, &d);
    tdna_alignment_destroy(a);
    ck(rc);
    return d;
}
bool Sequence::hasMinOverlap(const Sequence &o) const { return getOverlap(o) >= g_minOverlap; }

// ------------------------------------------------------------------------------------------------
// SequenceList
// ------------------------------------------------------------------------------------------------
void SequenceList::modified() {
    sortedBy_ = SORT_UNSORTED;
    engine_.reset();
}
int SequenceList::resort(int method) {
    int prev = sortedBy_;
    if (method == SORT_BYNAME && sortedBy_ != SORT_BYNAME) {
        std::stable_sort(seqs_.begin(), seqs_.end(), [](const SequencePtr &a, const SequencePtr &b) { return a->compareByDisplayName(*b) < 0; });
        engine_.reset();
    }
    sortedBy_ = method;
    return prev;
}
DistanceEngine &SequenceList::engine() const {
    if (!engine_) engine_ = std::make_shared<DistanceEngine>(*this);
    return *engine_;
}
SequenceList SequenceList::fromFastaText(const std::string &text) {
    SequenceList out;
    std::string name, seq;
    auto flush = [&]() {
        if (name.empty()) return;
        std::string nm = name;
        std::replace(nm.begin(), nm.end(), '_', ' ');  // makeSequence (FastaFile.java:268-270)
        std::string s;
        for (char ch : seq) {
            if (std::isspace((unsigned char)ch)) continue;
            if (ch == 'U') ch = 'T';
            if (ch == 'u') ch = 't';
            s.push_back(ch);
        }
        out.seqs_.push_back(std::make_shared<Sequence>(nm, s));
    };
    std::istringstream in(text);
    std::string line;
    auto blank = [](const std::string &l) { return std::all_of(l.begin(), l.end(), [](char c) { return std::isspace((unsigned char)c); }); };
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (blank(line)) continue;
        size_t f = line.find_first_not_of(" \t");
        if (line[f] == '#') continue;
        if (line[0] == '>') {
            flush();
            std::string nm = line.substr(1);
            size_t a = nm.find_first_not_of(" \t\f\v");
            nm = a == std::string::npos ? "" : nm.substr(a);
            name = nm.empty() ? "No name specified in file" : nm;
            seq.clear();
        } else seq += trim(line);
    }
    flush();
    out.modified();
    return out;
}
SequenceList SequenceList::fromFastaTextOnDevice(const std::string &text) {
    tdna_fasta *f = nullptr;
    ck(tdna_fasta_parse(DistanceEngine::context(), reinterpret_cast<const uint8_t *>(text.data()), (int64_t)text.size(), &f));
    struct Guard { tdna_fasta *f; ~Guard() { tdna_fasta_destroy(f); } } guard{f};
    const int64_t n = tdna_fasta_count(f);
    const int32_t width = tdna_fasta_width(f);
    std::vector<int64_t> off((size_t)n);
    std::vector<int32_t> hlen((size_t)n), slen((size_t)n), flags((size_t)n);
    ck(tdna_fasta_records(f, off.data(), hlen.data(), slen.data(), flags.data()));
    for (int32_t fl : flags)
        if (fl & (TDNA_FASTA_INVALID | TDNA_FASTA_NEEDS_HOST)) return fromFastaText(text);  // the host path words the errors / expands the groups
    std::vector<uint8_t> sym((size_t)n * std::max(width, 1));
    if (n > 0 && width > 0) ck(tdna_fasta_symbols(f, sym.data(), width));
    SequenceList out;
    for (int64_t r = 0; r < n; r++) {
        std::string nm = text.substr((size_t)off[r], (size_t)hlen[r]);  // after '>' up to the line end: "^>\\s*(.+)\\s*$" keeps trailing blanks
        size_t a = nm.find_first_not_of(" \t\f\v\r\n");
        nm = a == std::string::npos ? "" : nm.substr(a);
        if (nm.empty()) nm = "No name specified in file";
        std::replace(nm.begin(), nm.end(), '_', ' ');  // makeSequence (FastaFile.java:268-270)
        out.seqs_.push_back(std::make_shared<Sequence>(nm, std::string(reinterpret_cast<const char *>(sym.data() + (size_t)r * width), (size_t)slen[r]),
                                                       Sequence::Normalised{}));
    }
    out.modified();
    return out;
}
SequenceList SequenceList::readFasta(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return fromFastaText(ss.str());
}

// ------------------------------------------------------------------------------------------------
// java.util.TimSort (see the header).  MIN_MERGE 32, MIN_GALLOP 7, post-JDK-8072909 mergeCollapse.
// ------------------------------------------------------------------------------------------------
namespace {
struct TimSort {
    using Cmp = std::function<int(int, int)>;
    std::vector<int> &a;
    const Cmp &c;
    int minGallop = 7;
    std::vector<int> runBase, runLen, tmp;
    TimSort(std::vector<int> &a_, const Cmp &c_) : a(a_), c(c_) {}

    static void binarySort(std::vector<int> &a, int lo, int hi, int start, const Cmp &c) {
        if (start == lo) start++;
        for (; start < hi; start++) {
            int pivot = a[start];
            int left = lo, right = start;
            while (left < right) {
                int mid = (int)(((unsigned)left + (unsigned)right) >> 1);
                if (c(pivot, a[mid]) < 0) right = mid; else left = mid + 1;
            }
            for (int k = start; k > left; k--) a[k] = a[k - 1];
            a[left] = pivot;
        }
    }
    static int countRunAndMakeAscending(std::vector<int> &a, int lo, int hi, const Cmp &c) {
        int runHi = lo + 1;
        if (runHi == hi) return 1;
        if (c(a[runHi++], a[lo]) < 0) {
            while (runHi < hi && c(a[runHi], a[runHi - 1]) < 0) runHi++;
            std::reverse(a.begin() + lo, a.begin() + runHi);
        } else {
            while (runHi < hi && c(a[runHi], a[runHi - 1]) >= 0) runHi++;
        }
        return runHi - lo;
    }
    static int minRunLength(int n) {
        int r = 0;
        while (n >= 32) { r |= (n & 1); n >>= 1; }
        return n + r;
    }
    static int gallopLeft(int key, const int *a, int base, int len, int hint, const Cmp &c) {
        int lastOfs = 0, ofs = 1;
        if (c(key, a[base + hint]) > 0) {
            int maxOfs = len - hint;
            while (ofs < maxOfs && c(key, a[base + hint + ofs]) > 0) { lastOfs = ofs; ofs = (ofs << 1) + 1; if (ofs <= 0) ofs = maxOfs; }
            if (ofs > maxOfs) ofs = maxOfs;
            lastOfs += hint; ofs += hint;
        } else {
            int maxOfs = hint + 1;
            while (ofs < maxOfs && c(key, a[base + hint - ofs]) <= 0) { lastOfs = ofs; ofs = (ofs << 1) + 1; if (ofs <= 0) ofs = maxOfs; }
            if (ofs > maxOfs) ofs = maxOfs;
            int t = lastOfs; lastOfs = hint - ofs; ofs = hint - t;
        }
        lastOfs++;
        while (lastOfs < ofs) {
            int m = lastOfs + ((ofs - lastOfs) >> 1);
            if (c(key, a[base + m]) > 0) lastOfs = m + 1; else ofs = m;
        }
        return ofs;
    }
    static int gallopRight(int key, const int *a, int base, int len, int hint, const Cmp &c) {
        int ofs = 1, lastOfs = 0;
        if (c(key, a[base + hint]) < 0) {
            int maxOfs = hint + 1;
            while (ofs < maxOfs && c(key, a[base + hint - ofs]) < 0) { lastOfs = ofs; ofs = (ofs << 1) + 1; if (ofs <= 0) ofs = maxOfs; }
            if (ofs > maxOfs) ofs = maxOfs;
            int t = lastOfs; lastOfs = hint - ofs; ofs = hint - t;
        } else {
            int maxOfs = len - hint;
            while (ofs < maxOfs && c(key, a[base + hint + ofs]) >= 0) { lastOfs = ofs; ofs = (ofs << 1) + 1; if (ofs <= 0) ofs = maxOfs; }
            if (ofs > maxOfs) ofs = maxOfs;
            lastOfs += hint; ofs += hint;
        }
        lastOfs++;
        while (lastOfs < ofs) {
            int m = lastOfs + ((ofs - lastOfs) >> 1);
            if (c(key, a[base + m]) < 0) ofs = m; else lastOfs = m + 1;
        }
        return ofs;
    }
    void mergeCollapse() {
        while (runLen.size() > 1) {
            int n = (int)runLen.size() - 2;
            if ((n > 0 && runLen[n - 1] <= runLen[n] + runLen[n + 1]) || (n > 1 && runLen[n - 2] <= runLen[n] + runLen[n - 1])) {
                if (runLen[n - 1] < runLen[n + 1]) n--;
            } else if (n < 0 || runLen[n] > runLen[n + 1]) break;
            mergeAt(n);
        }
    }
    void mergeForceCollapse() {
        while (runLen.size() > 1) {
            int n = (int)runLen.size() - 2;
            if (n > 0 && runLen[n - 1] < runLen[n + 1]) n--;
            mergeAt(n);
        }
    }
    void mergeAt(int i) {
        int base1 = runBase[i], len1 = runLen[i], base2 = runBase[i + 1], len2 = runLen[i + 1];
        runLen[i] = len1 + len2;
        runBase.erase(runBase.begin() + i + 1);
        runLen.erase(runLen.begin() + i + 1);
        int k = gallopRight(a[base2], a.data(), base1, len1, 0, c);
        base1 += k; len1 -= k;
        if (len1 == 0) return;
        len2 = gallopLeft(a[base1 + len1 - 1], a.data(), base2, len2, len2 - 1, c);
        if (len2 == 0) return;
        if (len1 <= len2) mergeLo(base1, len1, base2, len2); else mergeHi(base1, len1, base2, len2);
    }
    [[noreturn]] static void violated() { throw std::invalid_argument("Comparison method violates its general contract!"); }
    void mergeLo(int base1, int len1, int base2, int len2) {
        tmp.assign(a.begin() + base1, a.begin() + base1 + len1);
        int cursor1 = 0, cursor2 = base2, dest = base1;
        a[dest++] = a[cursor2++];
        if (--len2 == 0) { std::copy(tmp.begin() + cursor1, tmp.begin() + cursor1 + len1, a.begin() + dest); return; }
        if (len1 == 1) { std::copy(a.begin() + cursor2, a.begin() + cursor2 + len2, a.begin() + dest); a[dest + len2] = tmp[cursor1]; return; }
        int mg = minGallop;
        bool done = false;
        while (!done) {
            int count1 = 0, count2 = 0;
            do {
                if (c(a[cursor2], tmp[cursor1]) < 0) {
                    a[dest++] = a[cursor2++]; count2++; count1 = 0;
                    if (--len2 == 0) { done = true; break; }
                } else {
                    a[dest++] = tmp[cursor1++]; count1++; count2 = 0;
                    if (--len1 == 1) { done = true; break; }
                }
            } while ((count1 | count2) < mg);
            if (done) break;
            do {
                count1 = gallopRight(a[cursor2], tmp.data(), cursor1, len1, 0, c);
                if (count1 != 0) {
                    std::copy(tmp.begin() + cursor1, tmp.begin() + cursor1 + count1, a.begin() + dest);
                    dest += count1; cursor1 += count1; len1 -= count1;
                    if (len1 <= 1) { done = true; break; }
                }
                a[dest++] = a[cursor2++];
                if (--len2 == 0) { done = true; break; }
                count2 = gallopLeft(tmp[cursor1], a.data(), cursor2, len2, 0, c);
                if (count2 != 0) {
                    std::copy(a.begin() + cursor2, a.begin() + cursor2 + count2, a.begin() + dest);  // forward copy, dest < cursor2
                    dest += count2; cursor2 += count2; len2 -= count2;
                    if (len2 == 0) { done = true; break; }
                }
                a[dest++] = tmp[cursor1++];
                if (--len1 == 1) { done = true; break; }
                mg--;
            } while (count1 >= 7 || count2 >= 7);
            if (done) break;
            if (mg < 0) mg = 0;
            mg += 2;
        }
        minGallop = mg < 1 ? 1 : mg;
        if (len1 == 1) {
            std::copy(a.begin() + cursor2, a.begin() + cursor2 + len2, a.begin() + dest);
            a[dest + len2] = tmp[cursor1];
        } else if (len1 == 0) violated();
        else std::copy(tmp.begin() + cursor1, tmp.begin() + cursor1 + len1, a.begin() + dest);
    }
    void mergeHi(int base1, int len1, int base2, int len2) {
        tmp.assign(a.begin() + base2, a.begin() + base2 + len2);
        int cursor1 = base1 + len1 - 1, cursor2 = len2 - 1, dest = base2 + len2 - 1;
        a[dest--] = a[cursor1--];
        if (--len1 == 0) { std::copy(tmp.begin(), tmp.begin() + len2, a.begin() + dest - (len2 - 1)); return; }
        if (len2 == 1) {
            dest -= len1; cursor1 -= len1;
            std::copy_backward(a.begin() + cursor1 + 1, a.begin() + cursor1 + 1 + len1, a.begin() + dest + 1 + len1);
            a[dest] = tmp[cursor2];
            return;
        }
        int mg = minGallop;
        bool done = false;
        while (!done) {
            int count1 = 0, count2 = 0;
            do {
                if (c(tmp[cursor2], a[cursor1]) < 0) {
                    a[dest--] = a[cursor1--]; count1++; count2 = 0;
                    if (--len1 == 0) { done = true; break; }
                } else {
                    a[dest--] = tmp[cursor2--]; count2++; count1 = 0;
                    if (--len2 == 1) { done = true; break; }
                }
            } while ((count1 | count2) < mg);
            if (done) break;
            do {
                count1 = len1 - gallopRight(tmp[cursor2], a.data(), base1, len1, len1 - 1, c);
                if (count1 != 0) {
                    dest -= count1; cursor1 -= count1; len1 -= count1;
                    std::copy_backward(a.begin() + cursor1 + 1, a.begin() + cursor1 + 1 + count1, a.begin() + dest + 1 + count1);
                    if (len1 == 0) { done = true; break; }
                }
                a[dest--] = tmp[cursor2--];
                if (--len2 == 1) { done = true; break; }
                count2 = len2 - gallopLeft(a[cursor1], tmp.data(), 0, len2, len2 - 1, c);
                if (count2 != 0) {
                    dest -= count2; cursor2 -= count2; len2 -= count2;
                    std::copy(tmp.begin() + cursor2 + 1, tmp.begin() + cursor2 + 1 + count2, a.begin() + dest + 1);
                    if (len2 <= 1) { done = true; break; }
                }
                a[dest--] = a[cursor1--];
                if (--len1 == 0) { done = true; break; }
                mg--;
            } while (count1 >= 7 || count2 >= 7);
            if (done) break;
            if (mg < 0) mg = 0;
            mg += 2;
        }
        minGallop = mg < 1 ? 1 : mg;
        if (len2 == 1) {
            dest -= len1; cursor1 -= len1;
            std::copy_backward(a.begin() + cursor1 + 1, a.begin() + cursor1 + 1 + len1, a.begin() + dest + 1 + len1);
            a[dest] = tmp[cursor2];
        } else if (len2 == 0) violated();
        else std::copy(tmp.begin(), tmp.begin() + len2, a.begin() + dest - (len2 - 1));
    }
};
}  // namespace

void javaSort(std::vector<int> &a, const std::function<int(int, int)> &cmp) {
    int lo = 0, hi = (int)a.size(), n = hi;
    if (n < 2) return;
    if (n < 32) {
        int init = TimSort::countRunAndMakeAscending(a, lo, hi, cmp);
        TimSort::binarySort(a, lo, hi, lo + init, cmp);
        return;
    }
    TimSort ts(a, cmp);
    int minRun = TimSort::minRunLength(n);
    do {
        int runLen = TimSort::countRunAndMakeAscending(a, lo, hi, cmp);
        if (runLen < minRun) {
            int force = n <= minRun ? n : minRun;
            TimSort::binarySort(a, lo, lo + force, lo + runLen, cmp);
            runLen = force;
        }
        ts.runBase.push_back(lo);
        ts.runLen.push_back(runLen);
        ts.mergeCollapse();
        lo += runLen;
        n -= runLen;
    } while (n != 0);
    ts.mergeForceCollapse();
}

// ------------------------------------------------------------------------------------------------
// SortedSequenceList (SortedSequenceList.java:78-128, comparator :197-261)
// ------------------------------------------------------------------------------------------------
void SortedSequenceList::sortAgainst(const SequencePtr &query, DelayCallback *delay) {
    if (!original_ || !query) throw std::runtime_error("sortAgainst called on a SortedSequenceList which hasn't been set up properly!");
    query_ = query;
    const auto &seqs = original_->sequences();
    const int n = (int)seqs.size();
    int q = -1;
    for (int i = 0; i < n; i++)
        if (seqs[i]->getId() == query->getId()) { q = i; break; }
    if (delay) delay->begin();
    if (n == 0) { dist_.clear(); shared_.clear(); }
    else if (q >= 0) original_->engine().rowDistances(q, dist_, &shared_);  // one launch: the row the reference builds with N getPairwise calls
    else original_->engine().queryDistances(*query, dist_, &shared_);      // a query that is not in the list (QuerySequence): one launch too
    sorted_.clear();
    for (int i = 0; i < n; i++) {
        if (delay) delay->delay(i + 1, n);
        if (!(dist_[i] < 0)) sorted_.push_back(i);  // :114
    }
    std::vector<std::optional<std::string>> names(n);
    for (int i = 0; i < n; i++) names[i] = seqs[i]->getSpeciesName();
    auto qname = query->getSpeciesName();
    const std::vector<double> &d = dist_;
    javaSort(sorted_, [&](int o1, int o2) -> int {
        double d1 = d[o1], d2 = d[o2];
        if (d1 < 0) return +1;
        if (d2 < 0) return -1;
        int64_t diff = Settings::makeLongFromDouble(d1 - d2);
        if (diff == 0) {
            bool q1 = (o1 == q), q2 = (o2 == q);
            if (q1 && q2) return 0;
            else if (q1) return -1;
            else if (q2) return +1;
            if (!names[o1] || !names[o2]) return 0;
            bool c1 = qname && *names[o1] == *qname, c2 = qname && *names[o2] == *qname;
            if (c1 && c2) return 0;
            if (c1) return -1;
            else if (c2) return +1;
            return 0;
        }
        return (int)diff;  // (int)(long): low 32 bits (:260)
    });
    if (delay) delay->end();
}
SequencePtr SortedSequenceList::get(int x) const {
    if (x < 0 || x >= (int)sorted_.size()) return nullptr;
    return original_->get(sorted_[x]);
}
double SortedSequenceList::getDistance(int x) const {
    if (!query_) return -1;
    return dist_[sorted_.at(x)];
}

// ------------------------------------------------------------------------------------------------
// PairwiseDistribution (PairwiseDistribution.java:109-152, 161-203, 247-297, 357-396)
// ------------------------------------------------------------------------------------------------
PairwiseDistribution::PairwiseDistribution(const SequenceList &list, int type, DelayCallback *delay) {
    if (type != PD_INTRA && type != PD_INTER) throw std::runtime_error("Programmer Error in PairwiseDistribution");
    if (delay) delay->begin();
    count_sequences_ = list.count();
    if (list.count() > 0) {
        if (type == PD_INTRA) {
            // conspecificIterator walks the FIRST contiguous run of the species in BYNAME order
            // (SequenceList.java:614-658, 983-1030); the device class is "same species id", which is
            // the same thing iff conspecifics are contiguous.  Refuse rather than differ.
            const auto &sp = list.engine().speciesId();
            std::unordered_map<int, int> last;
            for (int i = 0; i < (int)sp.size(); i++) {
                if (sp[i] < 0) continue;
                auto it = last.find(sp[i]);
                if (it != last.end() && it->second != i - 1)
                    throw EngineError("conspecific sequences are not contiguous in list order (resort(SORT_BYNAME) first)");
                last[sp[i]] = i;
            }
        }
        // the class as a histogram over its distinct fp64 distances (tdna_class_histogram); distances_push stores (float) d (:77-103)
        const auto &h = list.engine().classHistogram(type == PD_INTRA ? TDNA_PD_INTRA : TDNA_PD_INTER);
        std::vector<std::pair<float, int64_t>> e;
        e.reserve(h.size());
        for (const auto &x : h) e.emplace_back((float)x.d, x.count);
        // Arrays.sort(float[]): numeric order, -0.0 < 0.0, NaN last
        auto less = [](float a, float b) {
            bool na = a != a, nb = b != b;
            if (na || nb) return !na && nb;
            if (a == b) return std::signbit(a) && !std::signbit(b);
            return a < b;
        };
        std::sort(e.begin(), e.end(), [&](const std::pair<float, int64_t> &a, const std::pair<float, int64_t> &b) { return less(a.first, b.first); });
        for (size_t k = 0; k < e.size();) {  // several doubles round to one float: one run
            size_t m = k;
            int64_t cnt = 0;
            while (m < e.size() && !less(e[k].first, e[m].first)) cnt += e[m++].second;
            d_.push_run(e[k].first, cnt);
            k = m;
        }
    }
    if (delay) delay->end();
}
// ------------------------------------------------------------------------------------------------
// PairwiseDistances (PairwiseDistances.java:49-377) / PairwiseDistance (PairwiseDistance.java:30-75)
// ------------------------------------------------------------------------------------------------
bool PairwiseDistance::isMentioned(const Sequence &s) const { return s.getId() == seqA->getId() || s.getId() == seqB->getId(); }

PairwiseDistances::PairwiseDistances(const SequenceList &list, int type, DelayCallback *delay) {
    if (type != PD_INTRA && type != PD_INTER) throw std::runtime_error("Programmer Error in PairwiseDistances: Please inform the programmer!");
    if (delay) delay->begin();
    const int n = list.count();
    const auto &seqs = list.sequences();
    // the pushes in the reference's order: (query, seq) for every query in list order
    std::vector<int32_t> qi, qj;
    std::vector<int> first_of_query(n + 1, 0);
    std::vector<bool> has_entry(n, false);
    if (n > 0) {
        const auto &sp = list.engine().speciesId();
        if (type == PD_INTRA) {  // conspecificIterator = the first contiguous run of the species (SequenceList.java:614-658): see PairwiseDistribution
            std::unordered_map<int, int> last, first;
            for (int i = 0; i < n; i++) {
                if (sp[i] < 0) continue;
                auto it = last.find(sp[i]);
                if (it != last.end() && it->second != i - 1)
                    throw EngineError("conspecific sequences are not contiguous in list order (resort(SORT_BYNAME) first)");
                if (it == last.end()) first[sp[i]] = i;
                last[sp[i]] = i;
            }
            for (int q = 0; q < n; q++) {
                first_of_query[q] = (int)qi.size();
                if (sp[q] < 0) continue;  // :133 no species name: no pushes and NO average entry
                has_entry[q] = true;
                for (int x = first[sp[q]]; x <= last[sp[q]]; x++)
                    if (x != q) { qi.push_back(q); qj.push_back(x); }
            }
        } else {
            std::vector<const std::string *> genus(n), epi(n);
            for (int i = 0; i < n; i++) { genus[i] = &seqs[i]->getGenusName(); epi[i] = &seqs[i]->getSpeciesNameOnly(); }
            // group by genus string so that the O(N^2) string compares of :165-181 become per-genus loops (same pushes, same order)
            std::map<std::string, std::vector<int>> byGenus;
            for (int i = 0; i < n; i++) byGenus[*genus[i]].push_back(i);
            for (int q = 0; q < n; q++) {
                first_of_query[q] = (int)qi.size();
                has_entry[q] = true;
                for (int x : byGenus[*genus[q]])  // ascending list positions == list order
                    if (x != q && *epi[x] != *epi[q]) { qi.push_back(q); qj.push_back(x); }
            }
        }
        first_of_query[n] = (int)qi.size();
        std::vector<double> d = list.engine().pairs(qi, qj);  // ONE launch instead of one getPairwise per pair
        std::vector<int> kept;
        for (int q = 0; q < n; q++) {
            if (delay) delay->delay(count_sequences_, n);
            count_sequences_++;
            if (!has_entry[q]) continue;
            double total = 0.0;
            int count = 0;
            for (int k = first_of_query[q]; k < first_of_query[q + 1]; k++) {
                if (!(d[k] < 0)) kept.push_back(k);          // distances_push (:64-68): NaN is kept
                if (d[k] > -1) { total += d[k]; count++; }   // :146-149 / :177-180
            }
            averages_.emplace_back(seqs[q]->getFullName(), total / count);  // 0.0 / 0 = NaN, as in Java
        }
        // Collections.sort(distances) with PairwiseDistance.compareTo = (int) makeLongFromDouble(d1 - d2) (:58-63)
        if (!kept.empty())
            javaSort(kept, [&](int a, int b) -> int { return (int)(int32_t)(uint32_t)(uint64_t)Settings::makeLongFromDouble(d[a] - d[b]); });
        d_.reserve(kept.size());
        for (int k : kept) d_.push_back(PairwiseDistance{seqs[qi[k]], seqs[qj[k]], d[k]});
    }
    if (delay) delay->end();
}
int PairwiseDistances::getZero() const {
    int c = 0;
    for (auto &p : d_) { if (Settings::identical(p.distance, 0.0)) c++; else break; }
    return c;
}
int PairwiseDistances::getOne() const {
    int c = 0;
    for (int x = (int)d_.size() - 1; x >= 0; x--) { if (Settings::identical(d_[x].distance, 1.0)) c++; else break; }
    return c;
}
std::vector<PairwiseDistance> PairwiseDistances::getDistancesBetween(double from, double to) const {
    std::vector<PairwiseDistance> v;
    bool count = false;
    for (auto &p : d_) {
        if (p.distance >= from) count = true;
        if (p.distance > to) break;
        if (count) v.push_back(p);
    }
    return v;
}
double PairwiseDistances::getAverageDistance(const std::string &seqName) const {
    for (auto it = averages_.rbegin(); it != averages_.rend(); ++it)
        if (it->first == seqName) return it->second;
    return -1.0;
}
std::vector<std::string> PairwiseDistances::getAveragedSequences() const {
    std::vector<std::string> out;
    std::set<std::string> seen;
    for (auto &kv : averages_)
        if (seen.insert(kv.first).second) out.push_back(kv.first);
    return out;
}

float FloatRuns::operator[](int64_t k) const {
    size_t r = (size_t)(std::upper_bound(end_.begin(), end_.end(), k) - end_.begin());
    return v_[std::min(r, v_.size() - 1)];
}
std::vector<float> FloatRuns::expand() const {
    std::vector<float> out;
    out.reserve((size_t)size());
    for (size_t r = 0; r < v_.size(); r++) out.insert(out.end(), (size_t)run_count(r), v_[r]);
    return out;
}

static bool pdIdentical(float x, float y) { return (y - x) < 0.0000001; }  // :393-396 (one-sided, float vs double literal)
int PairwiseDistribution::getZero() const {
    int64_t c = 0;
    for (size_t r = 0; r < d_.runs(); r++) {
        if (pdIdentical(d_.run_value(r), 0.0f)) c += d_.run_count(r); else break;
    }
    return (int)c;
}
int PairwiseDistribution::getOne() const {
    int64_t c = 0;
    for (size_t r = d_.runs(); r-- > 0;) {
        if (pdIdentical(d_.run_value(r), 1.0f)) c += d_.run_count(r); else break;
    }
    return (int)c;
}
// :357-373 -- one sweep over the sorted values with a flag (so NaNs at the end are taken iff the flag is still up): the same sweep
// over the runs
FloatRuns PairwiseDistribution::runsBetween(double from, double to) const {
    FloatRuns v;
    bool count = false;
    for (size_t r = 0; r < d_.runs(); r++) {
        const float x = d_.run_value(r);
        if ((double)x >= from) count = true;
        if ((double)x > to) count = false;
        if (count) v.push_run(x, d_.run_count(r));
    }
    return v;
}
std::vector<float> PairwiseDistribution::getDistancesBetween(double from, double to) const { return runsBetween(from, to).expand(); }
int PairwiseDistribution::getBetween(double from, double to) const { return (int)runsBetween(from, to - 0.000001).size(); }
int PairwiseDistribution::getBetweenIncl(float from, float to) const { return (int)runsBetween((double)from, (double)to).size(); }
float PairwiseDistribution::getMaximumDistance() const { return d_.empty() ? 0.0f : d_.back(); }
float PairwiseDistribution::getMinimumDistance() const { return d_.empty() ? 0.0f : d_.front(); }
std::string PairwiseDistribution::getDistributionAsString(double d_from, double d_to, double d_interval, char dir) const {
    bool forward = dir != CUMUL_BACKWARD;
    std::string str;
    auto entry = [&](const std::string &a, const std::string &b, const std::string &c, const std::string &d) { str += a + "\t" + b + "\t" + c + "\t" + d + "\t\n"; };
    auto pc = [](double x, double y) { return javaDouble(Settings::percentage(x, y)); };
    entry("Distances", "Freq.", "Perc.", "Cumulative");
    int count = (int)d_.size();
    int f_from = getBetweenIncl(0, (float)d_from);
    float cumul = ((float)f_from / count);
    if (!forward) cumul = 1.0f - cumul;
    entry("<= " + pc(d_from, 1.0) + "%", std::to_string(f_from), pc(f_from, count), pc(cumul, 1.0));
    for (double x = d_from + 0.000001; x < d_to; x += d_interval) {
        double d_next = x + d_interval;
        f_from = getBetweenIncl((float)x, (float)d_next);
        if (forward) cumul += ((float)f_from / count); else cumul -= ((float)f_from / count);
        entry(pc(x, 1) + "% to " + pc(d_next, 1) + "%", std::to_string(f_from), pc(f_from, count), pc(cumul, 1.0));
    }
    f_from = getBetweenIncl((float)(d_to + 0.000001), (float)1.0);
    if (forward) cumul += ((float)f_from / count); else cumul -= ((float)f_from / count);
    entry("> " + pc(d_to, 1.0) + "%", std::to_string(f_from), pc(f_from, count), pc(cumul, 1.0));
    return str;
}

// ------------------------------------------------------------------------------------------------
// SpeciesDetails
// ------------------------------------------------------------------------------------------------
SpeciesDetails::SpeciesDetails(const SequenceList &list) {
    if (list.count() == 0) return;
    auto valid = list.engine().validConspecifics();
    std::unordered_map<std::string, int> idx;
    for (int i = 0; i < list.count(); i++) {
        auto sp = list.get(i)->getSpeciesName();
        if (!sp) continue;
        auto it = idx.find(*sp);
        int k;
        if (it == idx.end()) {
            k = (int)names.size();
            idx[*sp] = k;
            names.push_back(*sp);
            sequences_count.push_back(0);
            valid_conspecifics_count.push_back(0);
        } else k = it->second;
        sequences_count[k]++;
        valid_conspecifics_count[k] += valid[i] ? 1 : 0;
    }
}
int SpeciesDetails::indexOf(const std::string &name) const {
    for (size_t i = 0; i < names.size(); i++)
        if (names[i] == name) return (int)i;
    return -1;
}

// ------------------------------------------------------------------------------------------------
// BestMatch.run (BestMatch.java:164-582)
// ------------------------------------------------------------------------------------------------
namespace {
// What the three loops of BestMatch.run read off the sorted list for one query.
struct QueryOutcome {
    bool has_best = false;
    int best = -1;
    double best_d = 0;
    int count_bestMatches = 0;
    bool clean_block = false, mixed_block = false;
    int first_con = -1, first_allo = -1;
    double con_d = 0, allo_d = 0;
    int con_shared = 0, allo_shared = 0;
};

// Is the streamed per-row summary provably what the reference's sort would yield?  (DESIGN.md "near ties")
bool summaryIsExact(const tdna_row_result &r, const std::vector<int32_t> &sp) {
    if (r.flags & TDNA_ROW_NONFINITE) return false;
    if (r.n_valid == 0) return true;
    if (!(r.flags & TDNA_ROW_SELF_VALID)) return false;
    if (r.near_best != 0) return false;
    if (!(r.d_best == 0.0 || r.d_best >= 2e-5)) return false;  // nothing may crowd the query's own 0
    if (r.tie_count > 0 && (r.tie_unnamed > 0 || sp[r.best] < 0)) return false;
    if (r.con >= 0 && (r.near_con != 0 || r.unnamed_at_con != 0)) return false;
    if (r.allo >= 0 && (r.near_allo != 0 || r.unnamed_at_allo != 0)) return false;
    return true;
}
}  // namespace

std::string BestMatch::run(DelayCallback *delay) {
    best_match_correct = best_match_ambiguous = best_match_incorrect = 0;
    best_close_match_correct = best_close_match_ambiguous = best_close_match_incorrect = best_close_match_nomatch = 0;
    count_no_matches = count_zero_percent_matches = count_allo_at_zero = count_seqs_with_valid_conspecific_matches = 0;
    count_sequences_without_species_names = 0;
    rows_resolved_on_host = 0;
    const SequenceList &set = *list_;
    const int total = set.count();
    double threshold = threshold_percent_;
    threshold /= 100;
    std::string listing = "Query\tMatch\tIdentification\n";
    if (delay) delay->begin();
    std::vector<tdna_row_result> rows;
    if (total > 0) rows = set.engine().bestMatch();  // ONE device pass replaces N sortAgainst calls
    const auto &sp = set.engine().speciesId();
    SortedSequenceList sset(set);
    auto pc = [](double x, double y) { return javaDouble(Settings::percentage(x, y)); };
    for (int x = 0; x < total; x++) {
        const SequencePtr &query = set.get(x);
        if (delay) delay->delay(x, total);
        auto qname = query->getSpeciesName();
        if (!qname) { count_sequences_without_species_names++; continue; }
        QueryOutcome o;
        const tdna_row_result &r = rows[x];
        if (!force_exact && summaryIsExact(r, sp)) {
            if (r.n_valid > 0) {
                o.has_best = true;
                o.best = r.best;
                o.best_d = r.d_best;
                o.count_bestMatches = r.tie_count;
                if (r.tie_count > 0) {
                    int spb = sp[r.best];
                    if (r.tie_species_min == r.tie_species_max) { o.clean_block = (r.tie_species_min == spb); o.mixed_block = !o.clean_block; }
                    else o.mixed_block = true;
                }
                o.first_con = r.con; o.con_d = r.d_con; o.con_shared = r.shared_con;
                o.first_allo = r.allo; o.allo_d = r.d_allo; o.allo_shared = r.shared_allo;
            }
        } else {  // the reference's own procedure on this query's row (exact TimSort order)
            rows_resolved_on_host++;
            sset.sortAgainst(query, nullptr);
            int cs = sset.count();
            SequencePtr bm = sset.get(1);
            if (bm && sset.getDistance(1) != -1) {
                o.has_best = true;
                o.best = sset.position(1);
                o.best_d = sset.getDistance(1);
                for (int y = 2; y < cs; y++) {
                    SequencePtr match = sset.get(y);
                    if (!Settings::identical(sset.getDistance(y), o.best_d)) break;
                    o.count_bestMatches++;
                    auto mn = match->getSpeciesName();
                    if (!mn) throw NullPointerException("BestMatch.java:319: a tied match has no species name");
                    if (bm->getSpeciesName() && *mn == *bm->getSpeciesName()) o.clean_block = true; else o.mixed_block = true;
                }
                std::vector<int32_t> shared;
                bool haveShared = false;
                std::vector<double> dd;
                for (int y = 1; y < cs; y++) {
                    SequencePtr match = sset.get(y);
                    auto mn = match->getSpeciesName();
                    if (!mn) continue;
                    if (sset.getDistance(y) >= 0) {
                        if (o.first_con < 0 && *mn == *qname) { o.first_con = sset.position(y); o.con_d = sset.getDistance(y); }
                        else if (o.first_allo < 0 && *mn != *qname) { o.first_allo = sset.position(y); o.allo_d = sset.getDistance(y); }
                        if (o.first_con >= 0 && o.first_allo >= 0) break;
                    }
                }
                if (o.first_con >= 0 || o.first_allo >= 0) {
                    set.engine().rowDistances(x, dd, &shared);
                    haveShared = true;
                }
                if (haveShared && o.first_con >= 0) o.con_shared = shared[o.first_con];
                if (haveShared && o.first_allo >= 0) o.allo_shared = shared[o.first_allo];
            }
        }
        listing += query->getDisplayName();
        if (!o.has_best) {
            listing += "\t\tNo match.\n";
            count_no_matches++;
            continue;
        }
        const SequencePtr &bestMatch = set.get(o.best);
        double bd = o.best_d;
        if (Settings::identical(bd, 0)) count_zero_percent_matches++;
        if (o.first_con >= 0) count_seqs_with_valid_conspecific_matches++;
        if (o.first_con < 0) listing += "\tNo conspecific in database\t---\t0";
        else listing += "\t" + set.get(o.first_con)->getDisplayName() + "\t" + pc(o.con_d, 1) + "\t" + std::to_string(o.con_shared);
        if (o.first_allo < 0) listing += "\tNo allospecific in database\t---\t0";
        else listing += "\t" + set.get(o.first_allo)->getDisplayName() + "\t" + pc(o.allo_d, 1) + "\t" + std::to_string(o.allo_shared);
        bool conspecific = bestMatch->getSpeciesName() && *bestMatch->getSpeciesName() == *qname;
        bool within = bd <= threshold;
        const char *tail = within ? " (within threshold)\n" : " (outside threshold)\n";
        if (!o.clean_block && !o.mixed_block) {
            listing += "\t" + bestMatch->getDisplayName();
            if (conspecific) {
                listing += "\tSuccessful match at " + pc(bd, 1) + "%";
                best_match_correct++;
                if (within) best_close_match_correct++; else best_close_match_nomatch++;
            } else {
                if (Settings::identical(bd, 0)) count_allo_at_zero++;
                listing += "\tIncorrect match at " + pc(bd, 1) + "%";
                best_match_incorrect++;
                if (within) best_close_match_incorrect++; else best_close_match_nomatch++;
            }
            listing += tail;
        } else if (o.clean_block && !o.mixed_block) {
            listing += "\t" + bestMatch->getDisplayName() + " and " + std::to_string(o.count_bestMatches) + " others";
            if (conspecific) {
                listing += "\tSuccessful match at " + pc(bd, 1) + "%";
                best_match_correct++;
                if (within) best_close_match_correct++; else best_close_match_nomatch++;
            } else {
                if (Settings::identical(bd, 0)) count_allo_at_zero++;
                listing += "\tIncorrect match at " + pc(bd, 1) + "%";
                best_match_incorrect++;
                if (within) best_close_match_incorrect++; else best_close_match_nomatch++;
            }
            listing += tail;
        } else {
            if (Settings::identical(bd, 0)) count_allo_at_zero++;
            listing += "\t" + bestMatch->getDisplayName() + " and " + std::to_string(o.count_bestMatches) +
                       " others from different species\tMultiple species found at " + pc(bd, 1) + "%, identification with certainty is impossible";
            best_match_ambiguous++;
            if (within) best_close_match_ambiguous++; else best_close_match_nomatch++;
            listing += tail;
        }
    }
    int valid = total - count_no_matches - count_sequences_without_species_names;
    auto S = [](int v) { return std::to_string(v); };
    std::string text = "Sequences:\t" + S(total) + "\n" +
        "Sequences without recognizable species names (ignored in all subsequent counts):\t" + S(count_sequences_without_species_names) +
        "\nSequences with atleast one matching sequence in the data set:\t" + S(valid) + "\n" +
        "Sequences with atleast one matching conspecific sequence in the data set:\t" + S(count_seqs_with_valid_conspecific_matches) +
        "\nSequences with a closest match at 0%:\t" + S(count_zero_percent_matches) +
        "\nAllospecific matches at 0%:\t" + S(count_allo_at_zero) + "\t(" + pc(count_allo_at_zero, count_zero_percent_matches) + "% of all matches at 0%)" +
        "\n\nCorrect identifications according to \"Best Match\":\t" + S(best_match_correct) + " (" + pc(best_match_correct, valid) + "%)" +
        "\nAmbiguous according to \"Best Match\":\t" + S(best_match_ambiguous) + " (" + pc(best_match_ambiguous, valid) + "%)" +
        "\nIncorrect identifications according to \"Best Match\":\t" + S(best_match_incorrect) + " (" + pc(best_match_incorrect, valid) + "%)" +
        "\n\nCorrect identifications according to \"Best Close Match\":\t" + S(best_close_match_correct) + " (" + pc(best_close_match_correct, valid) + "%)" +
        "\nAmbiguous according to \"Best Close Match\":\t" + S(best_close_match_ambiguous) + " (" + pc(best_close_match_ambiguous, valid) + "%)" +
        "\nIncorrect identifications according to \"Best Close Match\":\t" + S(best_close_match_incorrect) + " (" + pc(best_close_match_incorrect, valid) + "%)" +
        "\nSequences without any match closer than " + pc(threshold, 1) + "%:\t" + S(best_close_match_nomatch) + " (" + pc(best_close_match_nomatch, valid) + "%)" +
        "\n\n" + listing;
    if (delay) delay->end();
    return text;
}

// ------------------------------------------------------------------------------------------------
// BlockAnalysis.run (BlockAnalysis.java:179-387)
// ------------------------------------------------------------------------------------------------
std::string BlockAnalysis::run(DelayCallback *delay) {
    block_correct = block_incorrect = block_ambiguous = block_nomatch = block_noseq = count_seq_without_species_name = 0;
    rows_resolved_on_host = 0;
    const SequenceList &set = *list_;
    const int count_sequences = set.count();
    double threshold = threshold_percent_;
    threshold /= 100;
    std::string listing = "Query\tFound in a complete block?\n", nonames = "Sequences without an identifiable species name\n";
    SpeciesDetails sd(set);
    if (delay) delay->begin();
    std::vector<tdna_row_result> rows;
    if (count_sequences > 0) rows = set.engine().bestMatch();
    const auto &sp = set.engine().speciesId();
    SortedSequenceList sset(set);
    auto pc = [](double x, double y) { return javaDouble(Settings::percentage(x, y)); };
    for (int x = 0; x < count_sequences; x++) {
        const SequencePtr &query = set.get(x);
        auto name_query = query->getSpeciesName();
        if (!name_query) {
            count_seq_without_species_name++;
            nonames += query->getFullName() + "\n";
            continue;
        }
        if (delay) delay->delay(x, count_sequences);
        const tdna_row_result &r = rows[x];
        int sorted_count;   // sset.count()
        int first = -1;     // list position of sset.get(1)
        double first_d = 0;
        int block_size = 1;
        bool exact_ok = !force_exact && !(r.flags & TDNA_ROW_NONFINITE) &&
                        (r.n_valid == 0 ||
                         ((r.flags & TDNA_ROW_SELF_VALID) && r.near_best == 0 && (r.d_best == 0.0 || r.d_best >= 2e-5) &&
                          !(r.tie_count > 0 && (r.tie_unnamed > 0 || sp[r.best] < 0)) &&
                          (r.other < 0 || (r.near_other == 0 && r.unnamed_at_other == 0))));
        bool is_correct = false, is_incorrect = false, nomatches = false;
        std::optional<std::string> name_first;
        if (exact_ok) {
            sorted_count = r.n_valid + ((r.flags & TDNA_ROW_SELF_VALID) ? 1 : 0);
            if (sorted_count > 0) {
                if (r.n_valid == 0) throw NullPointerException("BlockAnalysis.java:275: only the query survived sortAgainst");
                first = r.best;
                first_d = r.d_best;
                name_first = set.get(first)->getSpeciesName();
                if (!(first_d > threshold)) block_size = name_first ? 1 + r.block_count : 1;
            }
        } else {
            rows_resolved_on_host++;
            sset.sortAgainst(query, nullptr);
            sorted_count = sset.count();
            if (sorted_count > 0) {
                if (sorted_count < 2) throw NullPointerException("BlockAnalysis.java:275: only the query survived sortAgainst");
                first = sset.position(1);
                first_d = sset.getDistance(1);
                name_first = sset.get(1)->getSpeciesName();
                if (!(first_d > threshold)) {
                    for (int y = 2; y < sorted_count; y++) {
                        auto mn = sset.get(y)->getSpeciesName();
                        if (!mn) continue;
                        if (!name_first || *mn != *name_first) break;
                        block_size++;
                    }
                }
            }
        }
        if (sorted_count == 0) block_noseq++;
        else if (first_d > threshold) nomatches = true;
        else {
            if (block_size <= 1) block_ambiguous++;
            else {
                int k = sd.indexOf(*name_first);
                int vc = sd.valid_conspecifics_count[k];
                if (block_size == vc) {
                    if (*name_first != *name_query) { block_incorrect++; is_incorrect = true; } else block_ambiguous++;
                } else if (block_size == vc - 1) {
                    if (*name_first == *name_query) { block_correct++; is_correct = true; } else block_ambiguous++;
                } else block_ambiguous++;
            }
        }
        listing += query->getFullName();
        if (is_correct) listing += "\tcorrect\t";
        else if (is_incorrect) listing += "\tincorrect\t";
        else if (nomatches) { listing += "\tno match\t"; block_nomatch++; }
        else listing += "\tambiguous\t";
        listing += "\n";
    }
    int valid = count_sequences - block_noseq;
    auto S = [](int v) { return std::to_string(v); };
    std::string text = "Sequences:\t" + S(count_sequences) + "\nSequences without valid species names:\t" + S(count_seq_without_species_name) +
        "\nSequences with atleast one sequence with an overlap of " + S(Sequence::getMinOverlap()) + " base pairs:\t" + S(valid) +
        "\n\nCorrect identifications according to \"All Species Barcodes\":\t" + S(block_correct) + " (" + pc(block_correct, valid) + "%)" +
        "\nAmbiguous according to \"All Species Barcodes\":\t" + S(block_ambiguous) + " (" + pc(block_ambiguous, valid) + "%)" +
        "\nIncorrect identifications according to \"All Species Barcodes\":\t" + S(block_incorrect) + " (" + pc(block_incorrect, valid) + "%)" +
        "\nSequences with no match closer than " + pc(threshold, 1) + "%:\t" + S(block_nomatch) + " (" + pc(block_nomatch, valid) + "%)" +
        "\n\n" + listing + "\n\n" + nonames;
    if (delay) delay->end();
    return text;
}

// ------------------------------------------------------------------------------------------------
// PairwiseSummary.run (PairwiseSummary.java:109-485)
// ------------------------------------------------------------------------------------------------
static void padln(std::string &main, const std::string &x) {
    main += x;
    if (x.size() < 30) main += std::string(30 - x.size(), ' ');
    main += "\n";
}
static void padln(std::string &main, const std::string &x, const std::string &y) {
    main += x;
    if (x.size() < 30) main += std::string(30 - x.size(), ' ');
    main += "\t" + y + "\n";
}

std::string PairwiseSummary::run(DelayCallback *delay) {
    const SequenceList &set = *list_;
    five_ = 0;
    if (set.count() < 2) return "Pairwise distributions cannot be determined for less than two sequences";
    intra_ = std::make_shared<PairwiseDistribution>(set, PairwiseDistribution::PD_INTRA, delay);
    inter_ = std::make_shared<PairwiseDistribution>(set, PairwiseDistribution::PD_INTER, delay);
    PairwiseDistribution &intra = *intra_, &inter = *inter_;
    std::map<std::string, int> species;
    for (auto &s : set.sequences())
        if (auto n = s->getSpeciesName()) species[*n] = 1;
    std::string str;
    std::string mo = std::to_string(Sequence::getMinOverlap());
    auto pc = [](double x, double y) { return javaDouble(Settings::percentage(x, y)); };
    padln(str, "COUNTS");
    padln(str, "No of sequences:", std::to_string(set.count()) + "\tsequences");
    padln(str, "No of species:", std::to_string(species.size()) + "\tspecies");
    padln(str, "Note that minimum overlap is set to " + mo + " bp: pairs of sequences with less that " + mo +
                   " bp overlap will not be counted as a pairwise distance for this analysis.");
    padln(str, "\nRANGES");
    padln(str, "WARNING: These calculations are new to Species Identifier as of June 14, 2012, and have not been comprehensively tested yet. Please use with caution!\n");
    // The reference walks the sorted list from one end while (float) x / size <= 0.05 (:204-216, :236-247, :388-404).  The quotient
    // is non-decreasing in x, so the last x that passes is found by bisection and the value read off the runs: same result, O(log n).
    auto passes = [](int64_t x, int64_t size) { return !((double)((float)x / (float)size) > 0.05); };
    auto lastPassing = [&](int64_t lo, int64_t hi, int64_t size) {  // largest x in [lo, hi] that passes, lo - 1 if none (passes is monotone)
        if (hi < lo || !passes(lo, size)) return lo - 1;
        while (lo < hi) {
            int64_t mid = lo + (hi - lo + 1) / 2;
            if (passes(mid, size)) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    auto trimIntra = [&](const FloatRuns &dl) {  // x = 1 .. size - 1: dist = dl[size - x]
        const int64_t size = dl.size(), x = lastPassing(1, size - 1, size);
        return x >= 1 ? dl[size - x] : dl[size - 1];
    };
    auto trimInter = [&](const FloatRuns &dl, float start) {  // x = 0 .. size - 1: dist = dl[x]
        const int64_t size = dl.size();
        if (size == 0) return start;
        const int64_t x = lastPassing(0, size - 1, size);
        return x >= 0 ? dl[x] : start;
    };
    if (intra.countValidComparisons() == 0) {
        padln(str, "There are no valid intraspecific pairwise distances (i.e. no two sequences have less than " + mo + " bp overlap).");
    } else {
        FloatRuns dl = intra.runsBetween(0, 1);
        if (dl.empty()) throw NullPointerException("PairwiseSummary.java:204: no intraspecific distance in [0,1]");
        float mx = trimIntra(dl);
        padln(str, "Intraspecific pairwise distances range between " + pc(intra.getMinimumDistance(), 1) + "% and " + pc(intra.getMaximumDistance(), 1) +
                       "% (contains " + pc(intra.getBetweenIncl(intra.getMinimumDistance(), intra.getMaximumDistance()), intra.countValidComparisons()) +
                       "% of all valid intraspecific pairwise distances)");
        padln(str, "  Removing the largest 5% of intraspecific distances reduces that range to " + pc(intra.getMinimumDistance(), 1) + "% and " + pc(mx, 1) +
                       "% (contains " + pc(intra.getBetweenIncl(intra.getMinimumDistance(), mx), intra.countValidComparisons()) +
                       "% of all valid intraspecific pairwise distances)");
    }
    if (inter.countValidComparisons() == 0) {
        padln(str, "There are no valid interspecific pairwise distances (i.e. no two sequences have less than " + mo + " bp overlap).");
    } else {
        FloatRuns dl = inter.runsBetween(0, 1);
        float mn = trimInter(dl, inter.getMinimumDistance());
        padln(str, "Interspecific, intrageneric pairwise distances range between " + pc(inter.getMinimumDistance(), 1) + "% and " + pc(inter.getMaximumDistance(), 1) +
                       "% (contains " + pc(inter.getBetweenIncl(inter.getMinimumDistance(), inter.getMaximumDistance()), inter.countValidComparisons()) +
                       "% of all valid interspecific, intrageneric pairwise distances)");
        padln(str, "  Removing the largest 5% of interspecific, intrageneric distances reduces that range to " + pc(inter.getMinimumDistance(), 1) + "% and " + pc(mn, 1) +
                       "% (contains " + pc(inter.getBetweenIncl(mn, inter.getMaximumDistance()), inter.countValidComparisons()) +
                       "% of all valid interspecific, intrageneric pairwise distances)");
    }
    padln(str, "\nPERCENTAGES");
    float min_inter = inter.getMinimumDistance(), max_intra = intra.getMaximumDistance();
    int count_comparisons = inter.countValidComparisons() + intra.countValidComparisons();
    float overlap = std::fabs(max_intra - min_inter);
    if (inter.countValidComparisons() == 0) {
        padln(str, "Total overlap:\tNo interspecific, intrageneric comparisons.\n              \tIntraspecific comparisons range from " + pc(intra.getMinimumDistance(), 1) +
                       "% to " + pc(intra.getMaximumDistance(), 1) + "% (total width: " + pc(intra.getMaximumDistance() - intra.getMinimumDistance(), 1) + "%)");
    } else {
        int within = intra.getBetweenIncl(min_inter, max_intra) + inter.getBetweenIncl(min_inter, max_intra);
        FloatRuns dl_intra = intra.runsBetween(0, 1), dl_inter = inter.runsBetween(0, 1);
        if (dl_intra.empty()) {
            padln(str, "Total overlap:\t No intraspecific distances present.");
            padln(str, "Overlap with 5% error margs on both ends:\t No intraspecific distances present.");
        } else if (dl_inter.empty()) {
            padln(str, "Total overlap:\t No interspecific distances present.");
            padln(str, "Overlap with 5% error margs on both ends:\t No interspecific distances present.");
        } else {
            padln(str, "Total overlap:\t" + pc(overlap, 1) + "% (from " + pc(min_inter, 1) + "% to " + pc(max_intra, 1) + "%, covering " + pc(within, count_comparisons) +
                           "% of all intra and interspecific but intrageneric sequences)");
            min_inter = trimInter(dl_inter, min_inter);
            {   // (:388-404: the same walk as trimIntra, starting from max_intra)
                const int64_t size = dl_intra.size(), x = lastPassing(1, size - 1, size);
                if (x >= 1) max_intra = dl_intra[size - x];
            }
            overlap = std::fabs(max_intra - min_inter);
            within = inter.getBetweenIncl(min_inter, max_intra) + intra.getBetweenIncl(min_inter, max_intra);
            five_ = (float)Settings::percentage(max_intra, 1);
            padln(str, "Overlap with 5% error margins on both ends:\t " + pc(overlap, 1) + "% (from " + pc(min_inter, 1) + "% to " + pc(max_intra, 1) + "%, covering " +
                           pc(within, count_comparisons) + "% of intra and interspecific but intrageneric sequences)");
            padln(str, "Five percent intraspecific cutoff:\t " + javaFloat(five_) + "%");
        }
    }
    padln(str, "\nDISTRIBUTION FOR ALL INTRASPECIFIC DISTANCES (FROM 0.0 TO 0.20)");
    padln(str, intra.getDistributionAsString(0.0, 0.20, 0.005, PairwiseDistribution::CUMUL_BACKWARD));
    padln(str, "\nDISTRIBUTION FOR ALL INTRASPECIFIC DISTANCES (FROM 0.0 TO 0.30) AT 1% intervals");
    padln(str, intra.getDistributionAsString(0.0, 0.30, 0.01, PairwiseDistribution::CUMUL_BACKWARD));
    padln(str, "\nDISTRIBUTION FOR ALL INTRASPECIFIC DISTANCES");
    padln(str, intra.getDistributionAsString(0.0, 1.0, 0.05, PairwiseDistribution::CUMUL_BACKWARD));
    padln(str, "\nDISTRIBUTION FOR ALL INTERSPECIFIC DISTANCES WITHIN A GENUS (FROM 0.0 to 0.20)");
    padln(str, inter.getDistributionAsString(0.0, 0.20, 0.005, PairwiseDistribution::CUMUL_FORWARD));
    padln(str, "\nDISTRIBUTION FOR ALL INTERSPECIFIC DISTANCES WITHIN A GENUS (FROM 0.0 TO 0.30) AT 1% intervals");
    padln(str, inter.getDistributionAsString(0.0, 0.30, 0.01, PairwiseDistribution::CUMUL_FORWARD));
    padln(str, "\nDISTRIBUTION FOR ALL INTERSPECIFIC DISTANCES WITHIN A GENUS");
    padln(str, inter.getDistributionAsString(0.0, 1.0, 0.05, PairwiseDistribution::CUMUL_FORWARD));
    return str;
}

// ------------------------------------------------------------------------------------------------
// Cluster.run (Cluster.java:741-867): the device yields the edge set {0 <= d <= t}; the sequential
// merge that fixes cluster and member ORDER is replayed on the host from the adjacency lists.
// ------------------------------------------------------------------------------------------------
void Cluster::run(DelayCallback *delay) {
    const SequenceList &set = *list_;
    const int n = set.count();
    clusters_.clear();
    stats_.clear();
    if (n == 0) return;
    if (delay) delay->begin();
    std::vector<int32_t> ii, jj;
    std::vector<double> dd;
    set.engine().edgesBelow(max_pairwise_, ii, jj, dd);
    std::vector<std::vector<int>> earlier(n);  // neighbours with a smaller list position
    for (size_t e = 0; e < ii.size(); e++) earlier[std::max(ii[e], jj[e])].push_back(std::min(ii[e], jj[e]));
    std::vector<std::shared_ptr<std::vector<int>>> cl;  // `clusters` Vector, in order
    std::vector<std::shared_ptr<std::vector<int>>> owner(n);
    for (int s = 0; s < n; s++) {
        if (delay) delay->delay(s, n);
        // clusters that contain a neighbour, visited from the last cluster to the first (:797-855)
        std::vector<std::vector<int> *> hit;
        for (int nb : earlier[s]) hit.push_back(owner[nb].get());
        std::sort(hit.begin(), hit.end());
        hit.erase(std::unique(hit.begin(), hit.end()), hit.end());
        if (hit.empty()) {
            auto v = std::make_shared<std::vector<int>>(1, s);
            cl.push_back(v);
            owner[s] = v;
            continue;
        }
        std::shared_ptr<std::vector<int>> acc;
        for (int ci = (int)cl.size() - 1; ci >= 0; ci--) {
            if (!std::binary_search(hit.begin(), hit.end(), cl[ci].get())) continue;
            if (!acc) {
                cl[ci]->push_back(s);
                acc = cl[ci];
                owner[s] = acc;
            } else {
                for (int m : *cl[ci]) { acc->push_back(m); owner[m] = acc; }
                cl.erase(cl.begin() + ci);
            }
        }
    }
    for (auto &c : cl) clusters_.push_back(*c);
    // per-cluster statistics (:188-260) from the full matrix rows of the members
    std::vector<double> row;
    for (auto &b : clusters_) {
        Stat st{(int)b.size(), 0.0, 0, 0, 0.0};
        for (int a : b) {
            if (b.size() > 1) set.engine().rowDistances(a, row, nullptr);
            for (int c : b) {
                if (set.get(a)->getFullName() == set.get(c)->getFullName()) continue;
                st.valid_comparisons++;
                double d = row[c];
                if (d < 0) continue;
                if (d > st.largest_pairwise) st.largest_pairwise = d;
                if (d > max_pairwise_) st.valid_comparisons_over++;
            }
        }
        if (st.valid_comparisons != 0) st.percentage = Settings::percentage(st.valid_comparisons_over, st.valid_comparisons);
        stats_.push_back(st);
    }
    if (delay) delay->end();
}

// ------------------------------------------------------------------------------------------------
// ExtremePairwise.run (ExtremePairwise.java:72-210)
// ------------------------------------------------------------------------------------------------
// One tdna_row_extremes launch gives every query's record (the conspecifics and congeners only: O(sum of genus sizes squared) pair
// evaluations); a query whose record is not provably what the sorted list would show -- near ties around an extreme, NaN, the
// query not heading its own list, a genus beyond the kernel's capacity, ambiguity codes switched off (d(q, q) need not be 0 then) --
// is resolved with the exact sort as before.
static bool extreme_record_decides(const tdna_extreme_result &e) {
    if (e.flags & (TDNA_ROW_NONFINITE | TDNA_EXT_TOO_MANY)) return false;
    if (!(e.flags & TDNA_ROW_SELF_VALID) && (e.n_intra + e.n_inter) > 0) return false;  // entry 0 would not be the query (:149 skips it)
    if (Sequence::getPairwiseDistanceMethod() == Sequence::PDM_UNCORRECTED && !Sequence::areAmbiguousBasesAllowed()) return false;
    return true;
}
std::string ExtremePairwise::run(DelayCallback *delay) {
    const SequenceList &list = *list_;
    rows_.clear();
    rows_resolved_on_host = 0;
    std::string results = "Sequence name\tLargest conspecific match\tDistance\tOverlap\tClosest congeneric, interspecific match\tDistance\tOverlap\n";
    if (delay) delay->begin();
    SortedSequenceList sorted(list);
    const int total_sequences = list.count();
    std::vector<tdna_extreme_result> ext;
    if (!force_exact && total_sequences > 0) ext = list.engine().rowExtremes();
    auto entry = [&](int x) {  // display name, distance as a percentage, overlap
        return sorted.get(x)->getDisplayName() + "\t" + javaDouble(Settings::percentage(sorted.getDistance(x), 1)) + "\t" +
               std::to_string(sorted.getOverlap(x)) + "\t";
    };
    auto entry_of = [&](int row, double d, int shared) {
        return list.get(row)->getDisplayName() + "\t" + javaDouble(Settings::percentage(d, 1)) + "\t" + std::to_string(shared) + "\t";
    };
    for (int count = 0; count < total_sequences; count++) {
        if (delay) delay->delay(count, total_sequences);
        const SequencePtr &seq = list.get(count);
        auto name = seq->getSpeciesName();
        if (!name) {  // :109-112
            results += seq->getFullName() + "\tUnable to identify species name\n";
            rows_.push_back(Row{-2, -2});
            continue;
        }
        if (!ext.empty() && extreme_record_decides(ext[count]) && ext[count].near_intra == 0 && ext[count].near_inter == 0) {
            const tdna_extreme_result &e = ext[count];
            results += seq->getDisplayName() + "\t";
            results += e.max_intra < 0 ? std::string("No matching conspecific sequence\tN/A\tN/A\t") : entry_of(e.max_intra, e.d_max_intra, e.shared_intra);
            results += e.min_inter < 0 ? std::string("No matching congeneric, interspecific sequence\tN/A\tN/A\t") : entry_of(e.min_inter, e.d_min_inter, e.shared_inter);
            results += '\n';
            rows_.push_back(Row{e.max_intra, e.min_inter});
            continue;
        }
        rows_resolved_on_host++;
        sorted.sortAgainst(seq, nullptr);
        int largest = -1, smallest = -1;  // entries of the sorted list
        const std::string &genus1 = seq->getGenusName();
        for (int x = 0; x < sorted.count(); x++) {  // :121-146
            SequencePtr seq2 = sorted.get(x);
            auto name2 = seq2->getSpeciesName();
            if (!name2) continue;
            const std::string &genus2 = seq2->getGenusName();
            if (!genus1.empty() && !genus2.empty() && genus1 == genus2 && *name != *name2) {
                if (sorted.getDistance(x) != -1) { smallest = x; break; }
            }
        }
        for (int x = sorted.count() - 1; x >= 1; x--) {  // :149-163: entry 0 is taken to be the query
            SequencePtr seq2 = sorted.get(x);
            auto name2 = seq2->getSpeciesName();
            if (!name2) continue;
            if (*name == *name2 && sorted.getDistance(x) != -1) { largest = x; break; }
        }
        results += seq->getDisplayName() + "\t";
        results += largest < 0 ? std::string("No matching conspecific sequence\tN/A\tN/A\t") : entry(largest);
        results += smallest < 0 ? std::string("No matching congeneric, interspecific sequence\tN/A\tN/A\t") : entry(smallest);
        results += '\n';
        rows_.push_back(Row{largest < 0 ? -1 : sorted.position(largest), smallest < 0 ? -1 : sorted.position(smallest)});
    }
    if (delay) delay->end();
    return results;
}

// ------------------------------------------------------------------------------------------------
// QuerySequence.query (QuerySequence.java:113-180)
// ------------------------------------------------------------------------------------------------
std::vector<std::string> QuerySequence::query(const std::string &string_sequence, DelayCallback *delay) {
    positions_.clear();
    distances_.clear();
    std::string buff;
    for (char ch : string_sequence)
        if (!std::isspace((unsigned char)ch)) buff += ch;  // :127-133
    auto q = std::make_shared<Sequence>("Query", buff);     // throws SequenceException
    SortedSequenceList sset(*list_);
    sset.sortAgainst(q, delay);
    std::vector<std::string> out;
    for (int x = 0; x < sset.count(); x++) {
        double distance = sset.getDistance(x);
        if (distance < 0) continue;
        char num[64];
        std::snprintf(num, sizeof num, "%.4f", distance * 100);  // MessageFormat "({0,number,##0.0000}%) {1}"
        out.push_back(std::string("(") + num + "%) " + sset.get(x)->getDisplayName());
        positions_.push_back(sset.position(x));
        distances_.push_back(distance);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// PairwiseExplorer.run (PairwiseExplorer.java:109-190)
// ------------------------------------------------------------------------------------------------
std::vector<PairwiseExplorer::Category> PairwiseExplorer::run(int type, DelayCallback *delay) {
    std::shared_ptr<PairwiseDistances> &current = type == PairwiseDistances::PD_INTRA ? intra_ : inter_;
    if (!current) current = std::make_shared<PairwiseDistances>(*list_, type, delay);
    std::vector<Category> cats;
    for (int x = -5; x < 1000; x += 5) {
        double from = (double)x / 1000 + 0.000001;
        double to = (double)(x + 5) / 1000;
        auto distances = current->getDistancesBetween(from, to);
        if (!distances.empty()) {
            std::string key = javaFloat((float)x / 10) + "% to " + javaFloat(((float)x + 5) / 10) + "% (" + std::to_string(distances.size()) + " matches)";
            if (x + 5 == 0) key = "Before 0% (" + std::to_string(distances.size()) + " matches)";
            cats.push_back(Category{key, std::move(distances)});
        }
    }
    cats.push_back(Category{"Averages", {}});
    return cats;
}

// ------------------------------------------------------------------------------------------------
// DistanceAnalysis.run (DistanceAnalysis.java:161-320)
// The reference walks Hashtables keyed by Sequence objects (identity hash: an order that differs from run to run) and by
// species name; the sums here are taken in list order / order of first appearance.  Only the last ulp of a mean depends on
// that order, and the means are printed through Settings.percentage (two decimals of a percentage).
// ------------------------------------------------------------------------------------------------
std::string DistanceAnalysis::run(DelayCallback *delay) {
    const SequenceList &list = *list_;
    const int n = list.count();
    for (int i = 0; i < n; i++)
        if (!list.get(i)->getSpeciesName()) throw NullPointerException("DistanceAnalysis.java:196/205: a sequence without a species name");
    auto average = [](const std::vector<double> &v) {  // :139-151
        double sum = 0.0;
        int count = 0;
        for (double val : v)
            if (val > -1) { sum += val; count++; }
        return sum / count;
    };
    if (delay) delay->begin();
    SortedSequenceList ssl(list);
    std::vector<std::string> species_order;
    std::set<std::string> species_seen;
    struct PerSeq { int q; double avg, closest; };
    std::vector<PerSeq> per_seq;
    rows_resolved_on_host = 0;
    std::vector<tdna_extreme_result> ext;  // one launch: every query's closest / summed congeneric, allospecific distance
    if (!force_exact && n > 0) ext = list.engine().rowExtremes();
    for (int current_seq = 0; current_seq < n; current_seq++) {
        if (delay) delay->delay(current_seq, n);
        const SequencePtr &query = list.get(current_seq);
        const std::string qname = *query->getSpeciesName();
        if (species_seen.insert(qname).second) species_order.push_back(qname);
        if (!ext.empty()) {
            const tdna_extreme_result &e = ext[current_seq];
            // the record decides when the sorted order of the congeneric entries is beyond doubt (it fixes "closest" and the order of
            // the sum): no near ties among them, no NaN
            if (!(e.flags & (TDNA_ROW_NONFINITE | TDNA_EXT_TOO_MANY)) && e.near_inter == 0 && e.near_order == 0) {
                if (e.n_inter > 0) per_seq.push_back(PerSeq{current_seq, e.sum_inter / e.n_inter, e.d_min_inter});
                continue;
            }
        }
        rows_resolved_on_host++;
        ssl.sortAgainst(query, nullptr);
        double closest = -1;
        std::vector<double> inters;
        for (int x = 0; x < ssl.count(); x++) {
            SequencePtr compare = ssl.get(x);
            if (*compare->getSpeciesName() == qname) continue;
            if (compare->getGenusName() == query->getGenusName()) {
                double dist = ssl.getDistance(x);
                if (closest < 0) closest = dist;
                inters.push_back(dist);
            }
        }
        if (closest != -1) per_seq.push_back(PerSeq{current_seq, average(inters), closest});
    }
    std::map<std::string, std::vector<double>> genera_all, genera_smallest;  // std::string order == String.compareTo for the names the parser accepts
    std::unordered_map<std::string, std::vector<size_t>> by_species;  // the entries of a species, in order (the reference rescans them all)
    for (size_t k = 0; k < per_seq.size(); k++) by_species[*list.get(per_seq[k].q)->getSpeciesName()].push_back(k);
    for (const std::string &sp : species_order) {
        std::optional<std::string> genus;
        std::vector<double> d_all, d_small;
        for (size_t k : by_species[sp]) {
            const PerSeq &e = per_seq[k];
            genus = list.get(e.q)->getGenusName();
            d_all.push_back(e.avg);
            d_small.push_back(e.closest);
        }
        if (!genus) genus = sp.substr(0, sp.find(' '));
        if (!d_all.empty()) {
            genera_all[*genus].push_back(average(d_all));
            genera_smallest[*genus].push_back(average(d_small));
        } else {
            genera_all[*genus].push_back(-1);
            genera_smallest[*genus].push_back(-1);
        }
    }
    std::string buff = "Genus\t\t\tAvg. Avg. Smallest Inter\t\tAvg. Avg. Avg. Inter\n";
    for (auto &kv : genera_all) {
        double avg_inter = average(kv.second), avg_smallest = average(genera_smallest[kv.first]);
        if (avg_inter < 0) buff += kv.first + "\t\t\tNo comparisions\t\tNo comparisions\n";
        else buff += kv.first + "\t\t\t" + javaDouble(Settings::percentage(avg_smallest, 1)) + "\t\t" + javaDouble(Settings::percentage(avg_inter, 1)) + "\n";
    }
    if (delay) delay->end();
    return buff;
}

std::vector<std::string> Cluster::consensusSequences() const {
    return list_->engine().consensus(clusters_);
}

// ------------------------------------------------------------------------------------------------
// SequenceMatrix: the grid, FindDistances.displayResults (FindDistances.java:123-289), DisplayDistancesMode.getSortedSequences
// (DisplayDistancesMode.java:217-398)
// ------------------------------------------------------------------------------------------------
void SequenceGrid::setSequence(const std::string &colName, const std::string &seqName, const SequencePtr &seq) {
    for (auto &c : cells_)
        if (c.first.first == colName && c.first.second == seqName) { c.second = seq; return; }
    cells_.push_back({{colName, seqName}, seq});
}
void SequenceGrid::setCancelled(const std::string &colName, const std::string &seqName) { cancelled_.push_back({colName, seqName}); }
SequencePtr SequenceGrid::getSequence(const std::string &colName, const std::string &seqName) const {
    for (auto &c : cells_)
        if (c.first.first == colName && c.first.second == seqName) return c.second;
    return nullptr;
}
bool SequenceGrid::isSequenceCancelled(const std::string &colName, const std::string &seqName) const {
    for (auto &c : cancelled_)
        if (c.first == colName && c.second == seqName) return true;
    return false;
}
std::vector<std::pair<std::string, SequencePtr>> SequenceGrid::getSequenceListByColumn(const std::string &colName) const {
    std::map<std::string, SequencePtr> byTaxon;
    for (auto &c : cells_)
        if (c.first.first == colName && c.second) byTaxon[c.first.second] = c.second;
    std::vector<std::pair<std::string, SequencePtr>> out;
    for (auto &t : taxa_) {
        auto it = byTaxon.find(t);
        if (it != byTaxon.end()) out.push_back({t, it->second});
    }
    return out;
}

std::string FindDistances::displayResults(double from, double to, const std::string &colName, DelayCallback *delay) {
    results_.clear();
    std::vector<std::string> cols;  // :168-176
    if (colName.empty()) cols = grid_->getCharsets(); else cols.push_back(colName);
    if (delay) delay->begin();
    // one row per (column, sequence) in the order of the reference's outer loops (:200-207): the whole grid is ONE ragged alignment
    std::vector<SequencePtr> rows;
    std::vector<int> col_of;
    for (size_t c = 0; c < cols.size(); c++)
        for (auto &ts : grid_->getSequenceListByColumn(cols[c])) { rows.push_back(ts.second); col_of.push_back((int)c); }
    std::vector<double> dist;
    if (!rows.empty()) {
        SequenceList flat(rows);
        std::vector<int32_t> ii, jj;
        std::vector<double> d;
        flat.engine().pairsBetween(from, to, ii, jj, d);  // the pairs of :213-229, each unordered pair once
        // both directions, in the order the four nested loops meet them: (outer row, inner row) ascending
        std::vector<Result> found;
        found.reserve(2 * ii.size());
        for (size_t k = 0; k < ii.size(); k++) { found.push_back({ii[k], jj[k], d[k]}); found.push_back({jj[k], ii[k], d[k]}); }
        std::sort(found.begin(), found.end(), [](const Result &a, const Result &b) { return a.outer != b.outer ? a.outer < b.outer : a.inner < b.inner; });
        std::vector<int> order(found.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
        if (!order.empty())  // Collections.sort(results) with PairwiseDistance.compareTo (:249; PairwiseDistance.java:58-63)
            javaSort(order, [&](int a, int b) -> int { return (int)(int32_t)(uint32_t)(uint64_t)Settings::makeLongFromDouble(found[a].distance - found[b].distance); });
        for (int k : order) results_.push_back(found[k]);
    }
    std::string buff;
    auto renamed = [&](int row) {  // :231-235: a copy renamed "<display name> (from <column>)"; getFullName() returns the trimmed name
        Sequence copy(*rows[row]);
        copy.changeName(copy.getDisplayName() + " (from " + cols[col_of[row]] + ")");
        return copy.getFullName();
    };
    for (auto &r : results_) buff += renamed(r.outer) + "\t" + renamed(r.inner) + "\t" + javaDouble(Settings::percentage(r.distance, 1)) + "%\n";
    std::string pdm_method = "(unknown)";
    const int id = Sequence::getPairwiseDistanceMethod();
    if (id == Sequence::PDM_UNCORRECTED) pdm_method = "uncorrected pairwise distances";
    else if (id == Sequence::PDM_K2P) pdm_method = "K2P distances";
    else if (id == Sequence::PDM_TRANS_ONLY) pdm_method = "transversions only";
    if (delay) delay->end();
    return std::to_string(results_.size()) + " sequence pairs found with distances between " + javaDouble(Settings::percentage(from, 1)) + "% and " +
           javaDouble(Settings::percentage(to, 1)) + "%.\nNote that distances were calculated using " + pdm_method +
           " and sequence overlaps of less than " + std::to_string(Sequence::getMinOverlap()) + " bp weren't counted.\n\n" + buff;
}

void DisplayDistancesMode::activateDisplay(const std::string &selected_colName) {
    oldOverlap_ = Sequence::getMinOverlap();
    oldDistanceMode_ = Sequence::getPairwiseDistanceMethod();
    Sequence::setMinOverlap(1);
    Sequence::setPairwiseDistanceMethod(currentDistanceMethod_);
    selected_colName_ = selected_colName;
}
void DisplayDistancesMode::deactivateDisplay() {
    Sequence::setMinOverlap(oldOverlap_);
    Sequence::setPairwiseDistanceMethod(oldDistanceMode_);
}
std::vector<std::string> DisplayDistancesMode::getSortedSequences(const std::string &seqName_top) {
    columns_ = grid_->getCharsets();  // getSortedColumns :196-213: sorted, the selected column first
    std::sort(columns_.begin(), columns_.end());
    {
        auto it = std::find(columns_.begin(), columns_.end(), selected_colName_);
        if (it != columns_.end()) { columns_.erase(it); columns_.insert(columns_.begin(), selected_colName_); }
        else if (!columns_.empty()) selected_colName_ = columns_[0];
    }
    std::vector<std::string> taxa = grid_->getSequenceNames();
    const size_t nx = columns_.size(), ny = taxa.size();
    // every (taxon, reference) comparison of every gene in ONE launch over one ragged alignment of all the cells
    std::vector<SequencePtr> rows;
    std::map<std::pair<size_t, size_t>, int> index;
    for (size_t x = 0; x < nx; x++)
        for (size_t y = 0; y < ny; y++)
            if (SequencePtr s = grid_->getSequence(columns_[x], taxa[y])) { index[{x, y}] = (int)rows.size(); rows.push_back(s); }
    size_t ytop = ny;
    for (size_t y = 0; y < ny; y++)
        if (taxa[y] == seqName_top) { ytop = y; break; }
    std::vector<int32_t> qi, qj;
    std::map<std::pair<size_t, size_t>, size_t> slot;
    for (size_t x = 0; x < nx; x++) {
        auto top = ytop < ny ? index.find({x, ytop}) : index.end();
        if (top == index.end()) continue;
        for (size_t y = 0; y < ny; y++) {
            if (taxa[y] == seqName_top) continue;
            auto it = index.find({x, y});
            if (it == index.end()) continue;
            slot[{x, y}] = qi.size();
            qi.push_back(it->second);
            qj.push_back(top->second);
        }
    }
    std::vector<double> d;
    if (!qi.empty()) {
        SequenceList flat(rows);
        d = flat.engine().pairs(qi, qj);
    }
    std::vector<std::vector<double>> dist(nx, std::vector<double>(ny, DIST_ILLEGAL)), norm(nx, std::vector<double>(ny, 0.0));
    std::vector<double> mx(nx, -1.0), mn(nx, +2.0);
    for (size_t x = 0; x < nx; x++)  // pass 1 :255-290
        for (size_t y = 0; y < ny; y++) {
            double v;
            const bool has_top = ytop < ny && index.count({x, ytop});
            if (!has_top) v = DIST_NO_COMPARE_SEQ;
            else if (taxa[y] == seqName_top) v = DIST_SEQ_ON_TOP;
            else if (!index.count({x, y})) v = grid_->isSequenceCancelled(columns_[x], taxa[y]) ? DIST_CANCELLED : DIST_SEQ_NA;
            else { v = d[slot[{x, y}]]; if (v < 0) v = DIST_NO_OVERLAP; }
            dist[x][y] = v;
            if (v >= 0) { if (v > mx[x]) mx[x] = v; if (v < mn[x]) mn[x] = v; }
        }
    for (size_t x = 0; x < nx; x++)  // pass 2 :293-300
        for (size_t y = 0; y < ny; y++) norm[x][y] = dist[x][y] >= 0 ? (dist[x][y] - mn[x]) / (mx[x] - mn[x]) : dist[x][y];
    std::vector<double> total(ny);
    for (size_t y = 0; y < ny; y++) {  // pass 3 :305-326
        double totalScore = 0.0;
        int count = 0;
        for (size_t x = 0; x < nx; x++) {
            if (columns_[x] == selected_colName_) continue;
            if (dist[x][y] >= 0) { totalScore += norm[x][y]; count++; }
        }
        total[y] = totalScore / count;  // 0.0 / 0 = NaN, as in Java
    }
    std::vector<int> order(ny);
    for (size_t y = 0; y < ny; y++) order[y] = (int)y;
    if (!order.empty())  // Collections.sort(scores), Score.compareTo :106-127
        javaSort(order, [&](int a, int b) -> int {
            if (taxa[a] == seqName_top) return -1;
            if (taxa[b] == seqName_top) return +1;
            return (int)(int32_t)(uint32_t)(uint64_t)(Settings::makeLongFromDouble(total[a]) - Settings::makeLongFromDouble(total[b]));
        });
    distances_.assign(nx, std::vector<double>(ny));
    norm_distances_.assign(nx, std::vector<double>(ny));
    ranks_.assign(nx, std::vector<int>(ny));
    scores_.clear();
    std::vector<std::string> sorted;
    for (size_t k = 0; k < ny; k++) {
        // sequencesList.indexOf(seqName) :345: the FIRST taxon of that name
        size_t old = (size_t)(std::find(taxa.begin(), taxa.end(), taxa[order[k]]) - taxa.begin());
        for (size_t x = 0; x < nx; x++) { distances_[x][k] = dist[x][old]; norm_distances_[x][k] = norm[x][old]; }
        sorted.push_back(taxa[order[k]]);
        scores_.push_back(total[order[k]]);
    }
    for (size_t x = 0; x < nx; x++) {  // :371-385
        std::vector<double> ll;
        for (size_t y = 0; y < ny; y++)
            if (distances_[x][y] >= 0) ll.push_back(distances_[x][y]);
        std::sort(ll.begin(), ll.end());
        for (size_t y = 0; y < ny; y++) {
            if (distances_[x][y] >= 0) ranks_[x][y] = (int)(std::find(ll.begin(), ll.end(), distances_[x][y]) - ll.begin());
            else ranks_[x][y] = (int)std::floor(distances_[x][y]);
        }
    }
    return sorted;
}

}  // namespace taxondna
