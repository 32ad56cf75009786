"""B200-native exact cosine vector search: drop-in for the chromadb collection of
parsakhaz/multimodal-image-similarity-search (import name: ``mmiss_b200``).

The directory is called ``multimodal-image-similarity-search_b200`` (not a valid Python
identifier); ``mmiss_b200.py`` at the repo root loads it under the importable name.
"""
from ._native import VecSearchError, LIB_PATH, launch_count, load as load_native  # noqa: F401
from .index import DeviceIndex  # noqa: F401
from .collection import Collection, PersistentClient, Client  # noqa: F401
from .sharded_index import ShardedIndex  # noqa: F401
from .group_index import GroupIndex  # noqa: F401
from .service import SearchService, MicroBatcher, apply_filters, similarity_from_distance  # noqa: F401
from .sharded import (ShardedSearcher, shard_bounds, triangle_bounds, replicate_index,  # noqa: F401
                      find_duplicates_sharded)

__all__ = ["VecSearchError", "DeviceIndex", "GroupIndex", "Collection", "ShardedIndex", "PersistentClient", "Client", "SearchService", "MicroBatcher",
           "apply_filters", "similarity_from_distance", "ShardedSearcher", "shard_bounds", "triangle_bounds",
           "replicate_index", "find_duplicates_sharded", "launch_count",
           "load_native", "LIB_PATH"]
