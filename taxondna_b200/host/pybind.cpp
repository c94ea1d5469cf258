// pybind11 view of the host layer so that the parity tests read like the reference's own tests
// (Sequence("1", "AAAA...").getSharedLength(seq2) == 10, BestMatch(list).run(), ...).
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "taxondna_host.hpp"

namespace py = pybind11;
using namespace taxondna;

PYBIND11_MODULE(_host, m) {
    m.doc() = "TaxonDNA distance-path host layer over libtaxondna_cuda.so";
    py::register_exception<SequenceException>(m, "SequenceException");
    py::register_exception<NullPointerException>(m, "NullPointerException");
    py::register_exception<EngineError>(m, "EngineError");
    py::register_exception<DelayAbortedException>(m, "DelayAbortedException");

    py::class_<Settings>(m, "Settings")
        .def_static("makeLongFromDouble", &Settings::makeLongFromDouble)
        .def_static("roundOff", &Settings::roundOff)
        .def_static("percentage", &Settings::percentage)
        .def_static("identical", &Settings::identical);
    m.def("javaDouble", &javaDouble);
    m.def("javaFloat", &javaFloat);
    m.def("javaSort", [](std::vector<int> a, const std::function<int(int, int)> &cmp) {
        javaSort(a, cmp);
        return a;
    });
    m.def("launchCount", &DistanceEngine::launchCount);

    py::class_<Sequence, SequencePtr>(m, "Sequence")
        .def(py::init<const std::string &, const std::string &>())
        .def_readonly_static("PDM_UNCORRECTED", &Sequence::PDM_UNCORRECTED)
        .def_readonly_static("PDM_K2P", &Sequence::PDM_K2P)
        .def_readonly_static("PDM_TRANS_ONLY", &Sequence::PDM_TRANS_ONLY)
        .def_static("setMinOverlap", &Sequence::setMinOverlap)
        .def_static("getMinOverlap", &Sequence::getMinOverlap)
        .def_static("ambiguousBasesAllowed", &Sequence::ambiguousBasesAllowed)
        .def_static("areAmbiguousBasesAllowed", &Sequence::areAmbiguousBasesAllowed)
        .def_static("setPairwiseDistanceMethod", &Sequence::setPairwiseDistanceMethod)
        .def_static("getPairwiseDistanceMethod", &Sequence::getPairwiseDistanceMethod)
        .def_static("clearPairwiseCache", &Sequence::clearPairwiseCache)
        .def_static("isValid", &Sequence::isValid)
        .def_static("isAmbiguous", &Sequence::isAmbiguous)
        .def_static("isPurine", &Sequence::isPurine)
        .def_static("isPyrimidine", &Sequence::isPyrimidine)
        .def("getFullName", &Sequence::getFullName)
        .def("getDisplayName", &Sequence::getDisplayName)
        .def("getSpeciesName", &Sequence::getSpeciesName)
        .def("getGenusName", &Sequence::getGenusName)
        .def("getSpeciesNameOnly", &Sequence::getSpeciesNameOnly)
        .def("getSubspeciesName", &Sequence::getSubspeciesName)
        .def("getGI", &Sequence::getGI)
        .def("getWarningFlag", &Sequence::getWarningFlag)
        .def("getLength", &Sequence::getLength)
        .def("getSequence", &Sequence::getSequence)
        .def("getSequenceRaw", &Sequence::getSequenceRaw)
        .def("getPairwise", &Sequence::getPairwise)
        .def("getPairwiseNoBuffer", &Sequence::getPairwiseNoBuffer)
        .def("getSharedLength", &Sequence::getSharedLength)
        .def("getOverlap", &Sequence::getOverlap)
        .def("countIdentical", &Sequence::countIdentical)
        .def("countTransversions", &Sequence::countTransversions)
        .def("getK2PDistance", &Sequence::getK2PDistance)
        .def("hasMinOverlap", &Sequence::hasMinOverlap);

    py::class_<SequenceList>(m, "SequenceList")
        .def(py::init<>())
        .def(py::init<const std::vector<SequencePtr> &>())
        .def_static("readFasta", &SequenceList::readFasta)
        .def_static("fromFastaText", &SequenceList::fromFastaText)
        .def_static("fromFastaTextOnDevice", &SequenceList::fromFastaTextOnDevice)
        .def("rawSequences", [](const SequenceList &l) {
            std::vector<py::bytes> v;
            for (auto &s : l.sequences()) v.emplace_back(s->getSequenceRaw());
            return v;
        })
        .def_readonly_static("SORT_BYNAME", &SequenceList::SORT_BYNAME)
        .def("add", &SequenceList::add)
        .def("count", &SequenceList::count)
        .def("get", &SequenceList::get)
        .def("resort", &SequenceList::resort)
        .def("names", [](const SequenceList &l) {
            std::vector<std::string> v;
            for (auto &s : l.sequences()) v.push_back(s->getFullName());
            return v;
        })
        .def("matrix", [](const SequenceList &l) {
            auto d = l.engine().matrix();
            int64_t n = l.count();
            py::array_t<double> a({n, n});
            std::memcpy(a.mutable_data(), d.data(), d.size() * 8);
            return a;
        })
        .def("speciesIds", [](const SequenceList &l) { return l.engine().speciesId(); });

    py::class_<SortedSequenceList>(m, "SortedSequenceList")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("sortAgainst", [](SortedSequenceList &s, const SequencePtr &q) { s.sortAgainst(q, nullptr); })
        .def("count", &SortedSequenceList::count)
        .def("get", &SortedSequenceList::get)
        .def("getQuery", &SortedSequenceList::getQuery)
        .def("getDistance", &SortedSequenceList::getDistance)
        .def("getOverlap", &SortedSequenceList::getOverlap)
        .def("position", &SortedSequenceList::position);

    py::class_<PairwiseDistribution, std::shared_ptr<PairwiseDistribution>>(m, "PairwiseDistribution")
        .def(py::init([](const SequenceList &l, int type) { return std::make_shared<PairwiseDistribution>(l, type, nullptr); }))
        .def_readonly_static("PD_INTRA", &PairwiseDistribution::PD_INTRA)
        .def_readonly_static("PD_INTER", &PairwiseDistribution::PD_INTER)
        .def("countSequences", &PairwiseDistribution::countSequences)
        .def("countValidComparisons", &PairwiseDistribution::countValidComparisons)
        .def("getZero", &PairwiseDistribution::getZero)
        .def("getOne", &PairwiseDistribution::getOne)
        .def("getDistributionAsString", &PairwiseDistribution::getDistributionAsString)
        .def("getBetween", &PairwiseDistribution::getBetween)
        .def("getBetweenIncl", &PairwiseDistribution::getBetweenIncl)
        .def("getMaximumDistance", &PairwiseDistribution::getMaximumDistance)
        .def("getMinimumDistance", &PairwiseDistribution::getMinimumDistance)
        .def("getDistancesBetween", &PairwiseDistribution::getDistancesBetween);

    py::class_<PairwiseDistances, std::shared_ptr<PairwiseDistances>>(m, "PairwiseDistances")
        .def(py::init([](const SequenceList &l, int type) { return std::make_shared<PairwiseDistances>(l, type, nullptr); }))
        .def_readonly_static("PD_INTRA", &PairwiseDistances::PD_INTRA)
        .def_readonly_static("PD_INTER", &PairwiseDistances::PD_INTER)
        .def("countSequences", &PairwiseDistances::countSequences)
        .def("countValidComparisons", &PairwiseDistances::countValidComparisons)
        .def("getZero", &PairwiseDistances::getZero)
        .def("getOne", &PairwiseDistances::getOne)
        .def("getBetween", &PairwiseDistances::getBetween)
        .def("getBetweenIncl", &PairwiseDistances::getBetweenIncl)
        .def("getMaximumDistance", &PairwiseDistances::getMaximumDistance)
        .def("getMinimumDistance", &PairwiseDistances::getMinimumDistance)
        .def("getAverageDistance", &PairwiseDistances::getAverageDistance)
        .def("getAveragedSequences", &PairwiseDistances::getAveragedSequences)
        .def("triples", [](const PairwiseDistances &p) {  // (full name A, full name B, distance) in sorted order
            std::vector<std::tuple<std::string, std::string, double>> out;
            for (auto &x : p.all()) out.emplace_back(x.seqA->getFullName(), x.seqB->getFullName(), x.distance);
            return out;
        });

    py::class_<BestMatch>(m, "BestMatch")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("setThreshold", &BestMatch::setThreshold)
        .def("run", [](BestMatch &b) { return b.run(nullptr); })
        .def_readwrite("force_exact", &BestMatch::force_exact)
        .def_readonly("rows_resolved_on_host", &BestMatch::rows_resolved_on_host)
        .def_readonly("best_match_correct", &BestMatch::best_match_correct)
        .def_readonly("best_match_ambiguous", &BestMatch::best_match_ambiguous)
        .def_readonly("best_match_incorrect", &BestMatch::best_match_incorrect)
        .def_readonly("best_close_match_correct", &BestMatch::best_close_match_correct)
        .def_readonly("best_close_match_nomatch", &BestMatch::best_close_match_nomatch)
        .def_readonly("count_no_matches", &BestMatch::count_no_matches);

    py::class_<BlockAnalysis>(m, "BlockAnalysis")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("setThreshold", &BlockAnalysis::setThreshold)
        .def("run", [](BlockAnalysis &b) { return b.run(nullptr); })
        .def_readwrite("force_exact", &BlockAnalysis::force_exact)
        .def_readonly("rows_resolved_on_host", &BlockAnalysis::rows_resolved_on_host)
        .def_readonly("block_correct", &BlockAnalysis::block_correct)
        .def_readonly("block_ambiguous", &BlockAnalysis::block_ambiguous);

    py::class_<PairwiseSummary>(m, "PairwiseSummary")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("run", [](PairwiseSummary &p) { return p.run(nullptr); })
        .def("getFivePercentCutoff", &PairwiseSummary::getFivePercentCutoff)
        .def("getPD_intra", &PairwiseSummary::getPD_intra)
        .def("getPD_inter", &PairwiseSummary::getPD_inter);

    py::class_<Cluster::Stat>(m, "ClusterStat")
        .def_readonly("size", &Cluster::Stat::size)
        .def_readonly("largest_pairwise", &Cluster::Stat::largest_pairwise)
        .def_readonly("valid_comparisons", &Cluster::Stat::valid_comparisons)
        .def_readonly("valid_comparisons_over", &Cluster::Stat::valid_comparisons_over)
        .def_readonly("percentage", &Cluster::Stat::percentage);
    py::class_<Cluster>(m, "Cluster")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("setThreshold", &Cluster::setThreshold)
        .def("run", [](Cluster &c) { c.run(nullptr); })
        .def("clusters", &Cluster::clusters)
        .def("consensusSequences", [](const Cluster &c) {
            std::vector<py::bytes> out;
            for (auto &s : c.consensusSequences()) out.emplace_back(s);
            return out;
        })
        .def("stats", &Cluster::stats);

    py::class_<ExtremePairwise>(m, "ExtremePairwise")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("run", [](ExtremePairwise &e) { return e.run(nullptr); })
        .def_readwrite("force_exact", &ExtremePairwise::force_exact)
        .def_readonly("rows_resolved_on_host", &ExtremePairwise::rows_resolved_on_host)
        .def("rows", [](const ExtremePairwise &e) {
            std::vector<std::pair<int, int>> out;
            for (auto &r : e.rows()) out.emplace_back(r.largest_intra, r.smallest_inter);
            return out;
        });

    py::class_<QuerySequence>(m, "QuerySequence")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("query", [](QuerySequence &q, const std::string &s) { return q.query(s, nullptr); })
        .def("positions", &QuerySequence::positions)
        .def("distances", &QuerySequence::distances);

    py::class_<PairwiseExplorer>(m, "PairwiseExplorer")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def("run", [](PairwiseExplorer &p, int type) {  // [(key, [(full name A, full name B, distance)])]
            std::vector<std::pair<std::string, std::vector<std::tuple<std::string, std::string, double>>>> out;
            for (auto &c : p.run(type, nullptr)) {
                std::vector<std::tuple<std::string, std::string, double>> v;
                for (auto &x : c.distances) v.emplace_back(x.seqA->getFullName(), x.seqB->getFullName(), x.distance);
                out.emplace_back(c.key, std::move(v));
            }
            return out;
        });

    py::class_<DistanceAnalysis>(m, "DistanceAnalysis")
        .def(py::init<const SequenceList &>(), py::keep_alive<1, 2>())
        .def_readwrite("force_exact", &DistanceAnalysis::force_exact)
        .def_readonly("rows_resolved_on_host", &DistanceAnalysis::rows_resolved_on_host)
        .def("run", [](DistanceAnalysis &d) { return d.run(nullptr); });
    py::class_<SequenceGrid>(m, "SequenceGrid")
        .def(py::init<const std::vector<std::string> &, const std::vector<std::string> &>())
        .def("setSequence", &SequenceGrid::setSequence)
        .def("setCancelled", &SequenceGrid::setCancelled)
        .def("getCharsets", &SequenceGrid::getCharsets)
        .def("getSequenceNames", &SequenceGrid::getSequenceNames)
        .def("getSequence", &SequenceGrid::getSequence);

    py::class_<FindDistances>(m, "FindDistances")
        .def(py::init<const SequenceGrid &>(), py::keep_alive<1, 2>())
        .def("displayResults", [](FindDistances &f, double from, double to, const std::string &col) { return f.displayResults(from, to, col, nullptr); },
             py::arg("from_"), py::arg("to"), py::arg("colName") = std::string())
        .def("results", [](const FindDistances &f) {
            std::vector<std::tuple<int, int, double>> out;
            for (auto &r : f.results()) out.emplace_back(r.outer, r.inner, r.distance);
            return out;
        });

    auto ddm = py::class_<DisplayDistancesMode>(m, "DisplayDistancesMode");
    ddm.attr("DIST_ILLEGAL") = (double)DisplayDistancesMode::DIST_ILLEGAL;
    ddm.attr("DIST_NO_COMPARE_SEQ") = (double)DisplayDistancesMode::DIST_NO_COMPARE_SEQ;
    ddm.attr("DIST_SEQ_NA") = (double)DisplayDistancesMode::DIST_SEQ_NA;
    ddm.attr("DIST_NO_OVERLAP") = (double)DisplayDistancesMode::DIST_NO_OVERLAP;
    ddm.attr("DIST_SEQ_ON_TOP") = (double)DisplayDistancesMode::DIST_SEQ_ON_TOP;
    ddm.attr("DIST_CANCELLED") = (double)DisplayDistancesMode::DIST_CANCELLED;
    ddm
        .def(py::init<const SequenceGrid &>(), py::keep_alive<1, 2>())
        .def("setDistanceMethod", &DisplayDistancesMode::setDistanceMethod)
        .def("activateDisplay", &DisplayDistancesMode::activateDisplay)
        .def("deactivateDisplay", &DisplayDistancesMode::deactivateDisplay)
        .def("getSortedSequences", &DisplayDistancesMode::getSortedSequences)
        .def("columns", &DisplayDistancesMode::columns)
        .def("distances", &DisplayDistancesMode::distances)
        .def("normDistances", &DisplayDistancesMode::normDistances)
        .def("ranks", &DisplayDistancesMode::ranks)
        .def("scores", &DisplayDistancesMode::scores);
}
