"""Seeded inputs of the render-path golden vectors (tests/golden/render_*.npz): shared by the generator
(make_golden.py, which runs the reference's own Python on them) and by the tests (which run the oracle and the CUDA
path on the same inputs).  Parameters and inputs are rebuilt from seeds, so only OUTPUTS are stored in the fixtures.

  small_<kind>   reduced widths (64 / 32), the REAL per-ray shapes of default.yaml: 64 + 4 x 16 foreground samples,
                 32 background samples, n_thetas 8 (128 light directions), skip layer at 4; 2 views x 4 rays; every
                 gradient is stored
  full_default   default.yaml widths (256 / 128, skip 213 + 43), voxel 32^3 x 4, 1 view x 8 rays; large gradient
                 tensors are stored as norms + sampled entries (sampled_view)
  small_default_mask   small_default with train.mask_weight = 0.5 and a seeded object mask (obj_mask_of)
  small_default_l2     small_default with train.rgb_loss = l2
  small_spps008, small_pel4, small_terms_off   the shipped ablation configs that change shapes / switch terms off
  small_sphere_bounds  small_default with the ray bounds taken from the bounding sphere, cos_anneal_ratio 1
"""
import numpy as np

from ndjir_b200 import scene
from ndjir_b200.config import make_conf

CASES = {
    "small_default": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.3, G=16),
    "small_triplaneline": dict(kind="triplaneline", small=True, B=2, R=4, cos_anneal=0.0, G=32),
    "small_no_voxel": dict(kind="no_voxel", small=True, B=2, R=4, cos_anneal=1.0, G=None),
    "full_default": dict(kind="default", small=False, B=1, R=8, cos_anneal=0.5, G=32),
    # the mask term of loss.py:108-116 (weight 0 in every BASELINE config): mask_weight 0.5, a seeded 0 / 1 object mask
    "small_default_mask": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.6, G=16, mask_weight=0.5),
    # ray bounds from the sphere of radius bounding_sphere_radius (renderer.t_near_far_method: intersect_with_r_sphere,
    # sampler.py:84-91) instead of the box, at the annealing ratio images are rendered with
    # squared-error colour loss (train.rgb_loss: l2, loss.py:60-62; l1 in every shipped config)
    "small_default_l2": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.2, G=16, train={"rgb_loss": "l2"}),
    # shipped ablation configs: config/no_prior_varying_spps008.yaml (n_thetas 2 -> 8 light directions, no roughness /
    # specular priors), config/varying_pel4.yaml (4 encoding bands for the environment-light and soft-visibility
    # directions), config/varying_tv_weights0.0.yaml + varying_color_prior_weights0.00.yaml (terms switched off)
    "small_spps008": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16, renderer={"n_thetas": 2},
                          train={"roughness_prior_weight": 0.0, "specular_reflectance_prior_weight": 0.0}),
    "small_pel4": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                       extra={"environment_light_network": {"pe_bands": 4},
                              "soft_visibility_light_network": {"pe_bands": 4}}),
    "small_terms_off": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                            train={"tv_weight": 0.0, "base_color_prior_weight": 0.0}),
    # config/no_inv_distance_square.yaml: the photogrammetric light network without its 1 / d^2 input (network.py:410)
    "small_no_inv_dist": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                              extra={"photogrammetric_light_network": {"use_inverse_distance": False}}),
    # config/disentangle_diffuse.yaml: colour = VR(pl) (VR(bc) diffuse + specular)  (renderer.py:170-173)
    "small_disentangle": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                              extra={"diffuse_brdf": {"entangle": False}}),
    # config/uniform_sampling_on_sepcular.yaml: uniform specular directions, sBRDF = pi D V F (specular_brdf.py:104-108)
    "small_uniform_specular": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                                   extra={"specular_brdf": {"sampling": "uniform"}}),
    # config/no_implicit_illumination.yaml (network.py:308-309) and config/no_lightp.yaml (renderer.py:161, 174-176): the
    # network is never created, its parameters do not exist
    "small_no_ii": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                        extra={"implicit_illumination_network": {"use_me": False}}),
    "small_no_lightp": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16,
                            extra={"photogrammetric_light_network": {"use_me": False}}),
    # config/ste.yaml: geometric_network.voxel.use_ste, the normal does not differentiate the grid features
    "small_ste": dict(kind="default", small=True, B=2, R=4, cos_anneal=0.5, G=16, ste=True),
    "small_ste_triplaneline": dict(kind="triplaneline", small=True, B=2, R=4, cos_anneal=0.5, G=32, ste=True),
    "small_sphere_bounds": dict(kind="default", small=True, B=2, R=4, cos_anneal=1.0, G=16,
                                renderer={"t_near_far_method": "intersect_with_r_sphere"}),
}


def case_conf(name):
    c = CASES[name]
    over = dict(train={"batch_size": c["B"], "n_rays": c["R"], "mask_weight": c.get("mask_weight", 0.0)})
    if c["small"]:
        over.update(
            geometric_network={"feature_size": 64},
            base_color_network={"feature_size": 32}, environment_light_network={"feature_size": 32},
            soft_visibility_light_network={"feature_size": 32}, implicit_illumination_network={"feature_size": 32},
            photogrammetric_light_network={"feature_size": 32}, roughness_network={"feature_size": 32},
            specular_reflectance_network={"feature_size": 32},
            background_network={"feature_size0": 32, "feature_size1": 32})
    over["train"].update(c.get("train", {}))
    for sec, kv in c.get("extra", {}).items():
        over.setdefault(sec, {}).update(kv)
    if "renderer" in c:
        over["renderer"] = dict(c["renderer"])
    if c["G"] is not None:
        vox = {"grid_size": c["G"]}
        if c["kind"] == "triplaneline":
            vox["feature_size"] = 2
        if c.get("ste"):
            vox["use_ste"] = True
        over.setdefault("geometric_network", {})["voxel"] = vox
    return make_conf(c["kind"], **over)


def build_case(name):
    """-> conf, P (numpy parameters, reference layout), camloc, raydir, color_gt, rnd, cos_anneal"""
    c = CASES[name]
    conf = case_conf(name)
    P = scene.init_params(conf, seed=313, grid_std=0.05)
    rng = np.random.RandomState(7)       # make heads / SDF non-degenerate: perturb zero-initialised rows and biases
    for net, layers in P.items():
        if isinstance(layers, list):
            for (W, b) in layers:
                W += (rng.randn(*W.shape) * 0.02).astype(np.float32)
                b += (rng.randn(*b.shape) * 0.02).astype(np.float32)
    camloc, raydir, color_gt = scene.make_batch(conf, step=3, B=c["B"], R=c["R"])
    raydir[0, 0] = -raydir[0, 0]          # one ray that misses the box (mask 0)
    rnd = scene.make_randoms(conf, c["B"], c["R"], step=3)
    return conf, P, camloc, raydir, color_gt, rnd, c["cos_anneal"]


def obj_mask_of(name):
    """(B, R, 1) object mask of a case with a mask term (seeded), else None"""
    c = CASES[name]
    if not c.get("mask_weight"):
        return None
    return (np.random.RandomState(11).rand(c["B"], c["R"], 1) > 0.4).astype(np.float32)


BIG = 16384


def sampled_view(a, key):
    """What the fixtures keep of a large gradient tensor: its L2 norm, its sum, 1024 entries at seeded positions and the
    256 entries of largest magnitude (with their positions)."""
    flat = np.asarray(a, dtype=np.float64).reshape(-1)
    rng = np.random.RandomState(sum(map(ord, key)))
    pos = rng.randint(0, flat.size, 1024)
    top = np.argsort(-np.abs(flat))[:256]
    return dict(norm=np.float64(np.linalg.norm(flat)), sum=np.float64(flat.sum()), pos=pos, val=flat[pos], top=top,
                topval=flat[top])
