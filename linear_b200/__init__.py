"""linear_b200 -- B200-native approximate-map path of `linear filter` (xp3i4/linear) behind a C ABI.

Only what the hot path needs lives here: csrc/ (sm_100a kernels + the C ABI of include/lnr_b200.h),
api.py (ctypes mirror of the reference's operator interface for the path) and datagen.py (synthetic inputs).
"""
from . import datagen  # noqa: F401
from .api import (cords_to_records, Comm, create_index_sharded, Context, Features, Genome, Index, LnrError, Reads, apx_map_batch, apx_map_batch_packed, pack_dna5, cords_end, create_features,  # noqa: F401
                  create_index, hindex_shard_cuts, load_library, read_features, selftest_sort)
