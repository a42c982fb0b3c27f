"""Drop-in stand-ins for the reference's pybind11 extension modules.

Each sub-module has the reference module's NAME and the same positional function signatures
(``fn(N, out_ptr:int, in_ptr:int, ..., grid spec, D, min:list, max:list, boundary_check[, accum])``, reference
csrc/grid_feature/voxel_feature_cuda.cu:101-115 and friends), but forwards to libndjir_b200.so through its
C ABI.  ``N`` keeps the reference meaning (number of reference threads: B*D, B*D*3, L*B, B*R, B*R*M) and is
converted to points here.  ``install()`` registers them in ``sys.modules`` under the reference names so that
the reference's own ``python/grid_feature/*.py`` wrappers (``import voxel_feature_cuda``) bind to the B200
kernels unchanged.  Kernels go to the legacy default stream (0), like the reference.
"""
import importlib
import sys

MODULES = [
    "voxel_feature_cuda",
    "lanczos_voxel_feature_cuda",
    "voxel_hash_feature_cuda",
    "triplane_feature_cuda",
    "triline_feature_cuda",
    "cosine_voxel_feature_cuda",
    "cosine_triplane_feature_cuda",
    "cosine_triline_feature_cuda",
    "lanczos_triplane_feature_cuda",
    "lanczos_triline_feature_cuda",
    "lanczos_voxel_hash_feature_cuda",
    "total_variation_loss_cuda",
    "total_variation_loss_on_triplane_cuda",
    "total_variation_loss_on_triline_cuda",
    "total_variation_loss_on_voxel_hash_cuda",
    "ray_aabb_intersection_cuda",
    "ray_sphere_intersection_cuda",
    "inverse_transform_cuda",
    "squareplus_cuda",
]


def load(name):
    if name not in MODULES:
        raise ImportError(f"ndjir_b200.compat has no module {name}")
    return importlib.import_module(f"ndjir_b200.compat.{name}")


def install():
    """Make ``import voxel_feature_cuda`` (etc.) resolve to the B200 implementation."""
    for name in MODULES:
        sys.modules[name] = load(name)
