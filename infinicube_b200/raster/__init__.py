"""Host-side mirror of the reference's guidance-buffer API (camera, sparse grid, buffer generation)."""
from .camera import PinholeCamera  # noqa: F401
from .grid import VoxelGrid  # noqa: F401
from .fvdb_utils import generate_infinicube_buffer_from_fvdb_grid, points_to_fvdb  # noqa: F401
from .buffer_utils import generate_coordinate_buffer_from_memory_global_norm  # noqa: F401
from .semantic_utils import generate_rgb_semantic_buffer, semantic_to_color  # noqa: F401
from .sharding import gather_camera_shards, render_voxel_buffers_sharded, shard_cameras  # noqa: F401
