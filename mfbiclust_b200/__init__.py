"""mfbiclust_b200 -- B200-native (sm_100a) factorization hot path of mfBiclust.

Host-side mirror of the reference's R API for the path (``als_nmf``, ``snmf``, ``svd_pca``,
``nipals_pca``, ``bcv``/``auto_bcv``, ``otsuHelper``/``threshold``, ``filter.biclust``, ``addStrat``) on top of the
C ABI in ``include/mfbiclust.h`` (``libmfbiclust_cuda.so``, hand-written CUDA).  There is no
CPU fallback: importing works anywhere the shared library is built, calling needs a B200.
"""
from . import _cabi  # noqa: F401
from ._cabi import Context, Matrix, MfbError, default_context  # noqa: F401
from .backends import (AlsSession, HostAllReduce, ShardedAlsSession, als_nmf_sharded, als_replicate_sharded,
                       gather_rows, row_range, BiclusterExperiment, BiclusterStrategy, HostRng, addStrat,  # noqa: F401
                       als_nmf, als_replicate, auto_bcv, bcv, bcv_batch, bcvGivenKs, cell_range, clean, draw_folds,
                       genericFactorization, genericFit, kLimiter, nipals, nipals_pca, nipals_pca_nocatch, nnls,
                       otsuHelper, pseudovalues, shard_cells, svd_pca, threshold, thresholdHelper, validateKM,
                       biclusterMatrix2List, filter_biclust, overlap, overlap_sizes, sizes, union_biclust,
                       SnmfSession, nmf_snmf, snmf)
