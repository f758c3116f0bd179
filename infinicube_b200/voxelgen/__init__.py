"""Stage-1 adjacency of the hot path: the nearest-neighbour label transfer of the chunk merge."""
from .utils.color_util import color_from_points, knn_query_fast, semantic_from_points  # noqa: F401
