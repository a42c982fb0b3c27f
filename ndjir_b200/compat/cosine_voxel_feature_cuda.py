"""Stand-in for the reference module `cosine_voxel_feature_cuda` (csrc/grid_feature/cosine_voxel_feature_cuda.cu:855-866;
the reference exports five functions, its remaining second-order ones are commented out)."""
from .._lib import call


def query_on_voxel(N, output_ptr, query_ptr, feature_ptr, grid_sizes, D, min, max, boundary_check):
    call("ndjir_cosine_voxel_query_on_voxel", N // D, output_ptr, query_ptr, feature_ptr, list(grid_sizes), D, min, max, 0, 0)


def grad_query(N, grad_query_ptr, grad_output_ptr, query_ptr, feature_ptr, grid_sizes, D, min, max,
               boundary_check, accum):
    call("ndjir_cosine_voxel_grad_query", N // D, grad_query_ptr, grad_output_ptr, query_ptr, feature_ptr,
         list(grid_sizes), D, min, max, int(accum), 0)


def grad_feature(N, grad_feature_ptr, grad_output_ptr, query_ptr, grid_sizes, D, min, max, boundary_check, accum):
    call("ndjir_cosine_voxel_grad_feature", N // D, grad_feature_ptr, grad_output_ptr, query_ptr, list(grid_sizes), D,
         min, max, int(accum), 0)


def grad_query_grad_grad_output(N, grad_grad_output_ptr, grad_grad_query_ptr, query_ptr, feature_ptr, grid_sizes,
                                D, min, max, boundary_check, accum):
    call("ndjir_cosine_voxel_grad_query_grad_grad_output", N // D, grad_grad_output_ptr, grad_grad_query_ptr, query_ptr,
         feature_ptr, list(grid_sizes), D, min, max, int(accum), 0)


def grad_query_grad_feature(N, grad_feature_ptr, grad_grad_query_ptr, grad_output_ptr, query_ptr, grid_sizes, D,
                            min, max, boundary_check, accum):
    call("ndjir_cosine_voxel_grad_query_grad_feature", N // D, grad_feature_ptr, grad_grad_query_ptr, grad_output_ptr,
         query_ptr, list(grid_sizes), D, min, max, 0)
