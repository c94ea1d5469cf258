"""taxondna_b200: B200-native all-vs-all pairwise-distance engine behind TaxonDNA's
Sequence / SequenceList / SpeciesIdentifier-module API.  The compute lives in
libtaxondna_cuda.so (C ABI, include/taxondna_cuda.h); there is no CPU path."""
from ._lib import (KERNEL_AUTO, KERNEL_POPC, KERNEL_TENSOR, OPT_COUNT_KERNEL, OPT_BEST_MATCH_PATH, OPT_TILE_COLUMNS, OPT_MULTI, OPT_HIST_SLOTS, OPT_SEGMENT_ITEMS, OPT_PIPELINE_CHUNKS, OPT_CLUSTER, OPT_UNIT_ITEMS, E_CANCELLED, E_TOO_LARGE, COMM_ID_BYTES, PATH_AUTO, PATH_DENSE, PATH_STREAM, PD_INTER, PD_INTRA,  # noqa: F401
                   PDM_K2P, PDM_TRANS_ONLY, PDM_UNCORRECTED, ROW_NONFINITE, ROW_RESULT_DTYPE, ROW_SELF_VALID, WANT_BEST_MATCH, WANT_EDGES, WANT_HIST_INTRA, WANT_HIST_INTER, HIST_ENTRY_DTYPE, EXTREME_RESULT_DTYPE, EXT_UNNAMED, EXT_TOO_MANY, WANT_INTER, WANT_INTRA, Alignment, Analysis,
                   Context, Fasta, HostList, host_list, best_match_host, FASTA_INVALID, FASTA_NEEDS_HOST, FASTA_URACIL, StreamPart, TdnaError, check, load)
