"""ABI element types of the hot path, as numpy structured dtypes.

Byte-for-byte the reference's two PODs (reference: surfel_fusion/src/elements.h:5-20 and
:22-31): ``Superpixel_seed`` is 60 bytes (14 floats + 2 bools + 2 pad bytes before the two
debug floats) and ``SurfelElement`` is 44 bytes (9 floats + 2 int32).  The C side of the same
contract is ``include/dsm.h`` (dsm_seed_t / dsm_surfel_t).
"""
import numpy as np

SEED_DTYPE = np.dtype(
    {
        "names": ["x", "y", "size", "norm_x", "norm_y", "norm_z", "posi_x", "posi_y", "posi_z",
                  "view_cos", "mean_depth", "mean_intensity", "fused", "stable",
                  "min_eigen_value", "max_eigen_value"],
        "formats": ["<f4"] * 12 + ["u1", "u1", "<f4", "<f4"],
        "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 49, 52, 56],
        "itemsize": 60,
    }
)

SURFEL_DTYPE = np.dtype(
    {
        "names": ["px", "py", "pz", "nx", "ny", "nz", "size", "color", "weight",
                  "update_times", "last_update"],
        "formats": ["<f4"] * 9 + ["<i4", "<i4"],
        "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40],
        "itemsize": 44,
    }
)

# pcl::PointXYZI as published/saved by SurfelMap (surfel_map.h:31-32), without PCL's padding:
# dsm_point_t of include/dsm.h
POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4")])

assert SEED_DTYPE.itemsize == 60 and SURFEL_DTYPE.itemsize == 44 and POINT_DTYPE.itemsize == 16

SURFEL_FLOAT_FIELDS = ["px", "py", "pz", "nx", "ny", "nz", "size", "color", "weight"]
SEED_FLOAT_FIELDS = ["x", "y", "size", "norm_x", "norm_y", "norm_z", "posi_x", "posi_y", "posi_z",
                     "view_cos", "mean_depth", "mean_intensity"]

SP_SIZE = 8  # reference: fusion_functions.h:10


def num_seeds(width: int, height: int) -> int:
    """S = floor(W/8) * floor(H/8) (reference: fusion_functions.cpp:14-15,24)."""
    return (width // SP_SIZE) * (height // SP_SIZE)
