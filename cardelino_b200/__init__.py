"""cardelino_b200 -- B200-native clone-assignment sampler (drop-in for cardelino's clone_id path).

Only the hot path of the reference is here: clone_id / clone_id_Gibbs / clone_id_EM / get_logLik
(R/clone_id.R).  The compute lives in libcardelino_b200.so (hand-written sm_100a CUDA behind the C ABI
of include/cardelino_b200.h); this package is the host-side mirror of the R interface.
"""
from ._lib import CdlError, Handle, load_library, LIB_PATH, EXPORTS  # noqa: F401
from .clone_id import (clone_id, clone_id_Gibbs, clone_id_EM, get_logLik, colMatch, devianceIC,  # noqa: F401
                       Geweke_Z, assign_cells_to_clones, SparseCounts)

from .assessment import (rowMax, rowArgmax, binaryPRC, binaryROC, multiPRC, assign_scores)  # noqa: F401

__all__ = ["rowMax", "rowArgmax", "binaryPRC", "binaryROC", "multiPRC", "assign_scores", "clone_id", "clone_id_Gibbs", "clone_id_EM", "get_logLik", "colMatch", "devianceIC", "Geweke_Z",
           "assign_cells_to_clones", "Handle", "CdlError", "load_library"]
