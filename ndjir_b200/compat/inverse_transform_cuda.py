"""Stand-in for `inverse_transform_cuda` (csrc/sampling/inverse_transform_cuda.cu:163-168)."""
from .._lib import call


def sample_uniform_directions(size, light_dirs_ptr, normal_ptr, cdf_the_ptr, cdf_phi_ptr, batch_size, n_lights,
                              n_thes, n_phis, eps):
    call("ndjir_sample_uniform_directions", size, light_dirs_ptr, normal_ptr, cdf_the_ptr, cdf_phi_ptr, batch_size,
         n_lights, n_thes, n_phis, eps, 0)


def sample_importance_directions(size, light_dirs_ptr, normal_ptr, cdf_the_ptr, cdf_phi_ptr, alpha_ptr, batch_size,
                                 n_lights, n_thes, n_phis, eps):
    call("ndjir_sample_importance_directions", size, light_dirs_ptr, normal_ptr, cdf_the_ptr, cdf_phi_ptr,
         alpha_ptr, batch_size, n_lights, n_thes, n_phis, eps, 0)
