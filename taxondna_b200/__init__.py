"""taxondna_b200: B200-native all-vs-all pairwise-distance engine behind TaxonDNA's
Sequence / SequenceList / SpeciesIdentifier-module API.  The compute lives in
libtaxondna_cuda.so (C ABI, include/taxondna_cuda.h); there is no CPU path."""
from ._lib import (PD_INTER, PD_INTRA, PDM_K2P, PDM_TRANS_ONLY, PDM_UNCORRECTED, ROW_NONFINITE,  # noqa: F401
                   ROW_RESULT_DTYPE, ROW_SELF_VALID, Alignment, Context, TdnaError, check, load)
