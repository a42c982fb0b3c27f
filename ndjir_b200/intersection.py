"""`ray_aabb_intersection` / `ray_sphere_intersection` with the reference's signatures
(python/intersection/ray_aabb_intersection.py:113-116, ray_sphere_intersection.py) on torch CUDA tensors.
camloc (B,3), raydir (B,R,3) -> t_near, t_far, n_hits, each (B,R,1); no gradient (the reference's backward_impl is
empty)."""
import torch

from ._lib import call


def _prep(camloc, raydir):
    if not (camloc.is_cuda and raydir.is_cuda):
        raise ValueError("ndjir_b200 intersections need CUDA tensors (there is no CPU path)")
    B, R, _ = raydir.shape
    if camloc.shape != (B, 3):
        raise ValueError("camloc must be (B,3) and raydir (B,R,3)")          # ray_aabb_intersection.py:62-66
    outs = [torch.empty((B, R, 1), dtype=torch.float32, device=raydir.device) for _ in range(3)]
    return B, R, camloc.contiguous().float(), raydir.contiguous().float(), outs


def ray_aabb_intersection(camloc, raydir, min, max, ctx=None):
    B, R, c, d, (tn, tf, nh) = _prep(camloc, raydir)
    call("ndjir_ray_aabb_intersection", B * R, tn, tf, nh, c, d, B, R, list(min), list(max),
         torch.cuda.current_stream().cuda_stream)
    return tn, tf, nh


def ray_sphere_intersection(camloc, raydir, radius, ctx=None):
    B, R, c, d, (tn, tf, nh) = _prep(camloc, raydir)
    call("ndjir_ray_sphere_intersection", B * R, tn, tf, nh, c, d, B, R, float(radius),
         torch.cuda.current_stream().cuda_stream)
    return tn, tf, nh
