"""Stand-in for `ray_sphere_intersection_cuda` (csrc/intersection/ray_sphere_intersection_cuda.cu:81-104)."""
from .._lib import call


def ray_sphere_intersection(N, t_near_ptr, t_far_ptr, n_hits_ptr, camloc_ptr, raydir_ptr, B, R, radius):
    call("ndjir_ray_sphere_intersection", N, t_near_ptr, t_far_ptr, n_hits_ptr, camloc_ptr, raydir_ptr, B, R,
         radius, 0)
