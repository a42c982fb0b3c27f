"""Stand-in for `total_variation_loss_on_voxel_hash_cuda`
(csrc/grid_feature/total_variation_loss_on_voxel_hash_cuda.cu:229-234).  N = L * B; values in the (D, L, B) layout."""
from .._lib import call


def tv_loss_on_voxel_hash(N, output_ptr, query_ptr, feature_ptr, G0, growth_factor, T0, L, D, min, max, boundary_check):
    call("ndjir_tv_loss_on_voxel_hash", N // L, output_ptr, query_ptr, feature_ptr, G0, growth_factor, T0, L, D, min,
         max, 0, 0)


def tv_loss_on_voxel_hash_backward(N, grad_feature_ptr, grad_output_ptr, query_ptr, feature_ptr, G0, growth_factor, T0,
                                   L, D, min, max, boundary_check):
    call("ndjir_tv_loss_on_voxel_hash_backward", N // L, grad_feature_ptr, grad_output_ptr, query_ptr, feature_ptr, G0,
         growth_factor, T0, L, D, min, max, 0, 0)
