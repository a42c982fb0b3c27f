"""Stand-in for `total_variation_loss_cuda` (csrc/grid_feature/total_variation_loss_cuda.cu:203-209)."""
from .._lib import call


def tv_loss_on_voxel(N, output_ptr, query_ptr, feature_ptr, grid_sizes, D, min, max, boundary_check):
    call("ndjir_tv_loss_on_voxel", N // D, output_ptr, query_ptr, feature_ptr, list(grid_sizes), D, min, max, 0)


def tv_loss_on_voxel_backward(N, grad_feature_ptr, grad_output_ptr, query_ptr, feature_ptr, grid_sizes, D, min,
                              max, sym_backward, boundary_check, accum):
    call("ndjir_tv_loss_on_voxel_backward", N // D, grad_feature_ptr, grad_output_ptr, query_ptr, feature_ptr,
         list(grid_sizes), D, min, max, int(sym_backward), 0)
