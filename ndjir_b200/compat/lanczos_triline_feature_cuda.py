"""Stand-in for `lanczos_triline_feature_cuda` (csrc/grid_feature/lanczos_triline_feature_cuda.cu, PYBIND11_MODULE at the end).
N = B * D * 3 in the reference (one thread per point, channel and plane/line)."""
from .._lib import call


def query_on_triline(N, output_ptr, query_ptr, feature_ptr, G, D, min, max, boundary_check):
    call("ndjir_lanczos_triline_query_on_triline", N // (D * 3), output_ptr, query_ptr, feature_ptr, G, D, min, max, 0, 0)


def grad_query(N, grad_query_ptr, grad_output_ptr, query_ptr, feature_ptr, G, D, min, max, boundary_check, accum):
    call("ndjir_lanczos_triline_grad_query", N // (D * 3), grad_query_ptr, grad_output_ptr, query_ptr, feature_ptr, G, D,
         min, max, int(accum), 0)


def grad_feature(N, grad_feature_ptr, grad_output_ptr, query_ptr, G, D, min, max, boundary_check, accum):
    call("ndjir_lanczos_triline_grad_feature", N // (D * 3), grad_feature_ptr, grad_output_ptr, query_ptr, G, D, min, max,
         int(accum), 0)


def grad_query_grad_grad_output(N, grad_grad_output_ptr, grad_grad_query_ptr, query_ptr, feature_ptr, G, D, min,
                                max, boundary_check, accum):
    call("ndjir_lanczos_triline_grad_query_grad_grad_output", N // (D * 3), grad_grad_output_ptr, grad_grad_query_ptr,
         query_ptr, feature_ptr, G, D, min, max, int(accum), 0)


def grad_query_grad_feature(N, grad_feature_ptr, grad_grad_query_ptr, grad_output_ptr, query_ptr, G, D, min, max,
                            boundary_check, accum):
    call("ndjir_lanczos_triline_grad_query_grad_feature", N // (D * 3), grad_feature_ptr, grad_grad_query_ptr,
         grad_output_ptr, query_ptr, G, D, min, max, 0)
