"""Stand-in for `lanczos_voxel_hash_feature_cuda` (csrc/grid_feature/lanczos_voxel_hash_feature_cuda.cu:959-976).
Values use the reference kernel's (D, L, B) layout (layout=0), so the reference Python wrapper's in-place
transposes (voxel_hash_feature.py:152-155) keep working unchanged."""
from .._lib import call


def hash_index(N, output_ptr, query_ptr, T, boundary_check):
    call("ndjir_lanczos_voxel_hash_hash_index", N, output_ptr, query_ptr, T, 0)


def voxel_hash_feature(N, output_ptr, query_ptr, feature_ptr, G0, growth_factor, T0, L, D, min, max,
                       boundary_check):
    call("ndjir_lanczos_voxel_hash_voxel_hash_feature", N // L, output_ptr, query_ptr, feature_ptr, G0, growth_factor, T0,
         L, D, min, max, 0, 0, 0)


def grad_query(N, grad_query_ptr, grad_output_ptr, query_ptr, feature_ptr, G0, growth_factor, T0, L, D, min, max,
               boundary_check, accum):
    call("ndjir_lanczos_voxel_hash_grad_query", N // L, grad_query_ptr, grad_output_ptr, query_ptr, feature_ptr, G0,
         growth_factor, T0, L, D, min, max, 0, int(accum), 0)


def grad_feature(N, grad_feature_ptr, grad_output_ptr, query_ptr, G0, growth_factor, T0, L, D, min, max,
                 boundary_check, accum):
    call("ndjir_lanczos_voxel_hash_grad_feature", N // L, grad_feature_ptr, grad_output_ptr, query_ptr, G0, growth_factor,
         T0, L, D, min, max, 0, int(accum), 0)


def grad_query_grad_grad_output(N, grad_grad_output_ptr, grad_grad_query_ptr, query_ptr, feature_ptr, G0,
                                growth_factor, T0, L, D, min, max, boundary_check, accum):
    call("ndjir_lanczos_voxel_hash_grad_query_grad_grad_output", N // L, grad_grad_output_ptr, grad_grad_query_ptr,
         query_ptr, feature_ptr, G0, growth_factor, T0, L, D, min, max, 0, int(accum), 0)


def grad_query_grad_feature(N, grad_feature_ptr, grad_grad_query_ptr, grad_output_ptr, query_ptr, G0,
                            growth_factor, T0, L, D, min, max, boundary_check, accum):
    call("ndjir_lanczos_voxel_hash_grad_query_grad_feature", N // L, grad_feature_ptr, grad_grad_query_ptr,
         grad_output_ptr, query_ptr, G0, growth_factor, T0, L, D, min, max, 0, 0)
