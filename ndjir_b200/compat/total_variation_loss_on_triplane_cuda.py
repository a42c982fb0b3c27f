"""Stand-in for `total_variation_loss_on_triplane_cuda` (csrc/grid_feature/total_variation_loss_on_triplane_cuda.cu)."""
from .._lib import call


def tv_loss_on_triplane(N, output_ptr, query_ptr, feature_ptr, G, D, min, max, boundary_check):
    call("ndjir_tv_loss_on_triplane", N // (D * 3), output_ptr, query_ptr, feature_ptr, G, D, min, max, 0)


def tv_loss_on_triplane_backward(N, grad_feature_ptr, grad_output_ptr, query_ptr, feature_ptr, G, D, min, max,
                               sym_backward, boundary_check, accum):
    call("ndjir_tv_loss_on_triplane_backward", N // (D * 3), grad_feature_ptr, grad_output_ptr, query_ptr,
         feature_ptr, G, D, min, max, int(sym_backward), 0)
