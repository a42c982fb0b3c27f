"""Ray-sharded data parallelism of the hot path (new work: the reference is single-GPU).

One process per GPU; every rank renders its own rays with replicated parameters.  The only exchanges are
  * one scalar before the backward pass: sum(mask) over all ranks, so the loss denominators
    (sum(mask) * N + 1e-5, python/loss.py:64,74,95,118) equal the single-process values for the union of rays;
  * the gradient all-reduce (sum): the flat MLP gradient buffer (~5.8 MB) and the grid gradient.
The colour loss is normalised by the global ray count (engine passes 1/(B*R*world_size)).  NCCL over NVLink on the
GPUs; the same code runs on gloo/CPU tensors for the host-side tests.
"""
import torch
import torch.distributed as dist


def shard_rays(raydir, color_gt, rank, world_size, *extra):
    """Strong-scaling split of one batch along the ray axis R: rank k gets rays [k*R/w, (k+1)*R/w) of every view.
    raydir, color_gt (B,R,3); extra tensors are split along dim 1 the same way."""
    R = raydir.shape[1]
    if R % world_size != 0:
        raise ValueError(f"n_rays={R} is not divisible by world_size={world_size}")
    n = R // world_size
    sl = slice(rank * n, (rank + 1) * n)
    out = [raydir[:, sl].contiguous(), color_gt[:, sl].contiguous()]
    out += [e[:, sl].contiguous() for e in extra]
    return out


def allreduce_mask_sum(mask_sum, group=None):
    """mask_sum: 1-element tensor holding this rank's sum(mask); summed in place over the group."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mask_sum, op=dist.ReduceOp.SUM, group=group)
    return mask_sum


def allreduce_gradients(params, group=None, include_grid=True):
    """Sums params.grad (flat MLP gradients) and, unless the grid gradient was exchanged sparsely
    (allgather_rows + replicated scatter), every grid gradient over the group, in place."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    dist.all_reduce(params.grad, op=dist.ReduceOp.SUM, group=group)
    if include_grid:
        for v in params.grid_grad.values():
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)


def allgather_rows(out, local, group=None):
    """out (world*rows, c) <- concatenation over ranks of local (rows, c).  Used for the SPARSE exchange of the grid
    gradient: instead of all-reducing the dense table gradient (2 GiB for the 512^3 x 4 voxel grid), every rank
    gathers the per-sample scatter inputs of all ranks (query point + gradient rows, ~100 B per sample, ~26 MB per
    rank and step) and runs the scatter kernels over the union, so each replica ends with the full gradient."""
    dist.all_gather_into_tensor(out, local, group=group)
    return out
