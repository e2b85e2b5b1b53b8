"""dual_threshold_optimization_b200 -- B200-native (sm_100a) implementation of the Dual Threshold Optimization hot
path behind the reference crate's own surface (lib.rs:9-10 re-exports RankedFeatureList, PermutedRankedFeatureList,
optimize).  All computation happens in the CUDA library (csrc/, include/dto_b200.h); there is no CPU fallback."""
from ._capi import DtoError, DtoPanic  # noqa: F401
from .collections import Feature, FeatureList, PermutedRankedFeatureList, RankedFeatureList  # noqa: F401
from .dto import OptimizationResultRecord, compute_population_size, optimize, process_threshold_pairs  # noqa: F401
from .engine import Engine, device_count  # noqa: F401
from .read import read_feature_list_from_file, read_ranked_feature_list_from_csv  # noqa: F401
from .run import Task, run_multi_gpu, run_pairs, run_single_node  # noqa: F401
from .stat_operations import empirical_pvalue, fdr, hypergeometric_pvalue, intersect_genes  # noqa: F401

__version__ = "0.2.0"
