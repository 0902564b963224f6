"""Drop-in replacement for the reference's `pointops` Python package.

Same public names, argument order, dtypes and return tuples as
`libs/pointops/functions/__init__.py:1-14` of HaoyiZhu/PointCloudMatters, backed by the sm_100a
kernels of libpcm_b200.so through its C ABI (include/pcm_b200.h).  Make it importable under the
reference's name with `pointcloudmatters_b200.install_as_pointops()`.
"""
from .ops import (
    aggregation,
    attention_fusion_step,
    attention_relation_step,
    ball_query,
    ball_query_and_group,
    batch2offset,
    farthest_point_sampling,
    grouping,
    grouping2,
    interpolation,
    interpolation2,
    knn_query,
    knn_query_and_group,
    offset2batch,
    query_and_group,
    random_ball_query,
    subtraction,
)

__all__ = [
    "aggregation", "attention_fusion_step", "attention_relation_step", "ball_query",
    "ball_query_and_group", "batch2offset", "farthest_point_sampling", "grouping", "grouping2",
    "interpolation", "interpolation2", "knn_query", "knn_query_and_group", "offset2batch",
    "query_and_group", "random_ball_query", "subtraction",
]
