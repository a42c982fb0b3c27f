"""Stand-in for `ray_aabb_intersection_cuda` (csrc/intersection/ray_aabb_intersection_cuda.cu:145-169)."""
from .._lib import call


def ray_aabb_intersection(N, t_near_ptr, t_far_ptr, n_hits_ptr, camloc_ptr, raydir_ptr, B, R, min, max):
    call("ndjir_ray_aabb_intersection", N, t_near_ptr, t_far_ptr, n_hits_ptr, camloc_ptr, raydir_ptr, B, R, min,
         max, 0)
