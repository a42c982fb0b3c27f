"""`sample_points`, `sample_uniform_directions`, `sample_importance_directions` with the reference's signatures
(python/sampler.py:311-314, :401-408) on torch CUDA tensors.  None of them propagates gradients, like the
reference (SamplePoints.backward_impl / SampleDirections.backward_impl are empty, sampler.py:301, :391)."""
import torch

from ._lib import call
from .engine import get_engine


def sample_points(camloc, raydir, stratified_sample, background_sample, conf, ctx=None):
    """camloc (B,3), raydir (B,R,3), stratified_sample (B,R,N0,1) in [0,1), background_sample (B,R,Nb+1,1) in
    [1e-5,1) -> x_fg (B,R,N,3), t_fg (B,R,N+1,1), x_bg (B,R,Nb,4), t_bg (B,R,Nb+1,1), mask (B,R,1,1).
    The returned tensors are views of the engine's buffers (valid until the next call)."""
    return get_engine(conf).sample_points(camloc.contiguous(), raydir.contiguous(), stratified_sample.contiguous(),
                                          background_sample.contiguous())


def _dirs(name, normal, cdf_the, cdf_phi, alpha, eps):
    B, R, _ = normal.shape
    nt, nph = cdf_the.shape[-1], cdf_phi.shape[-1]
    M = nt * nph
    out = torch.empty((B, R, M, 3), dtype=torch.float32, device=normal.device)
    args = [B * R * M, out, normal.contiguous(), cdf_the.contiguous(), cdf_phi.contiguous()]
    if alpha is not None:
        args.append(alpha.contiguous())
    call(name, *args, B * R, M, nt, nph, float(eps), torch.cuda.current_stream().cuda_stream)
    return out


def sample_uniform_directions(normal, cdf_the, cdf_phi, eps=0.0, ctx=None):
    """normal (B,R,3), cdf_the (B,R,n_thetas), cdf_phi (B,R,n_phis) -> (B,R,n_thetas*n_phis,3)."""
    return _dirs("ndjir_sample_uniform_directions", normal, cdf_the, cdf_phi, None, eps)


def sample_importance_directions(normal, cdf_the, cdf_phi, alpha, eps=0.0, ctx=None):
    """GGX importance sampling around the normal; alpha (B,R,1) = roughness."""
    return _dirs("ndjir_sample_importance_directions", normal, cdf_the, cdf_phi, alpha, eps)
