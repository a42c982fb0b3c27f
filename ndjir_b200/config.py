"""Configuration of the rendering hot path: the YAML keys the reference reads on the path
(config/default.yaml in the reference; SURVEY.md section 8b lists them), as a nested attribute object with the
same names so code reads like the reference (`conf.renderer.n_samples0`, `conf.geometric_network.voxel.type` ...).
Only keys consumed by the hot path are kept; dataset / validation / extraction keys are out of scope.
"""
import copy
from types import SimpleNamespace

DEFAULT = {
    "use_wn": False,
    "geometric_network": {
        "pe_bands": 6, "feature_size": 256, "layers": 8, "act": "softplus", "skip_layers": [4],
        "geometric_init": True, "initial_sphere_radius": 0.35, "use_inv_square": True,
        "voxel": {"type": "voxel", "grid_size": 512, "feature_size": 4, "use_ste": False},
    },
    "base_color_network": {"feature_size": 256, "layers": 4, "act": "softplus", "use_geometric_feature": True,
                           "use_normal": False},
    "environment_light_network": {"pe_bands": 6, "feature_size": 128, "layers": 4, "act": "softplus", "channels": 1,
                                  "act_last": "softplus", "inverse_black_degree": 1, "upper_bound": -1},
    "soft_visibility_light_network": {"pe_bands": 6, "feature_size": 128, "layers": 4, "act": "softplus",
                                      "channels": 1, "act_last": "sigmoid", "inverse_black_degree": 1,
                                      "use_geometric_feature": True, "use_normal": True},
    "implicit_illumination_network": {"use_me": True, "use_me_on_specular": False, "feature_size": 128, "layers": 4,
                                      "act": "softplus", "channels": 1, "use_geometric_feature": True,
                                      "use_normal": True, "act_last": "sigmoid", "inverse_black_degree": 1},
    "photogrammetric_light_network": {"use_me": True, "pe_bands": 4, "feature_size": 256, "layers": 4,
                                      "act": "softplus", "use_inverse_distance": True, "channels": 1},
    "roughness_network": {"feature_size": 128, "layers": 4, "act": "softplus", "lower_bound": 0.089,
                          "use_geometric_feature": True, "use_normal": True, "prior_value": 0.5},
    "specular_reflectance_network": {"fixme": False, "feature_size": 128, "layers": 4, "act": "softplus",
                                     "channels": 3, "use_geometric_feature": True, "use_normal": True,
                                     "upper_bound_scale": 0.16, "prior_value": 0.04},
    "diffuse_brdf": {"entangle": True},
    "specular_brdf": {"model": "filament", "remap": True, "sampling": "importance", "use_split_sum": False,
                      "weight": 1.0},
    "background_modeling": True,
    "background_color": 0.0,
    "background_network": {"pe_bands0": 6, "pe_bands1": 4, "feature_size0": 256, "feature_size1": 256,
                           "layers0": 4, "layers1": 2, "act": "softplus"},
    "renderer": {"n_samples0": 64, "n_upsamples": 4, "n_samples1": 16, "n_bg_samples": 32,
                 "sampling_sigmoid_gain": 64, "eps": 5.0e-05, "eps_dot": 1e-8, "eps_normal": 1e-16,
                 "bounding_sphere_radius": 1.0, "t_near_far_method": "intersect_with_aabb", "deterministic": True,
                 "n_thetas": 8,
                 "diffuse_cdf_the_seed": 412, "diffuse_cdf_phi_seed": 124, "specular_cdf_the_seed": 810,
                 "specular_cdf_phi_seed": 108, "stratified_sample_seed": 913, "background_sample_seed": 510},
    "train": {"batch_size": 4, "n_rays": 512, "epoch": 1500, "base_learning_rate_weight": 5e-4,
              "base_learning_rate_feat": 5e-4, "learning_rate_end_ratio": 0.01, "warmup_term_ratio": 0.015,
              "cos_anneal_term_ratio": 0.15, "weight_decay": 1e-3, "clip_grad_norm": 0, "sigmoid_gain": 0.3, "sigmoid_gain_lv_start": 1,
              "sigmoid_gain_lv_end": 1, "rgb_loss": "l1", "eikonal_weight": 0.1, "tv_weight": 0.1,
              "tv_sym_backward": True, "mask_weight": 0.0, "base_color_prior_weight": 0.1,
              "base_color_prior_sym_backward": True, "base_color_perturb_seed": 913,
              "roughness_prior_weight": 1e-5, "specular_reflectance_prior_weight": 1e-3},
}

# the named configs of BASELINE.json as overrides of default.yaml (reference config/no_voxel.yaml,
# config/triplaneline.yaml differ from default.yaml by exactly these keys on the hot path)
OVERRIDES = {
    "default": {},
    "no_voxel": {"geometric_network": {"voxel": {"type": "none"}}},
    "triplaneline": {"geometric_network": {"voxel": {"type": "triplaneline", "grid_size": 2048, "feature_size": 8}}},
}


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _ns(d):
    return SimpleNamespace(**{k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


def to_dict(conf):
    return {k: to_dict(v) if isinstance(v, SimpleNamespace) else v for k, v in vars(conf).items()}


def make_conf(name="default", **overrides):
    """make_conf("default", geometric_network={"voxel": {"grid_size": 32}}, train={"n_rays": 64})"""
    d = copy.deepcopy(DEFAULT)
    _merge(d, copy.deepcopy(OVERRIDES[name]))
    _merge(d, copy.deepcopy(overrides))
    return _ns(d)


def grid_channels(conf):
    """Width of the grid feature appended to the positional encoding (reference network.py:120-151)."""
    v = conf.geometric_network.voxel
    if v.type == "none":
        return 0
    if v.type == "voxel":
        return v.feature_size
    if v.type == "triplaneline":
        return 2 * 3 * v.feature_size
    raise ValueError(f"voxel.type {v.type!r} is not supported by the B200 path")


def _set_path(d, dotted, value):
    keys = dotted.split(".")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


_SCI = None


def _coerce(v):
    """PyYAML follows YAML 1.1, where `1e-8` (no dot) is a STRING; OmegaConf, which the reference loads its files
    with, reads it as a float (config/default.yaml: eps_dot, eps_normal, the prior weights)."""
    global _SCI
    if _SCI is None:
        import re
        _SCI = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)[eE][-+]?\d+$")
    if isinstance(v, dict):
        return {k: _coerce(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_coerce(x) for x in v]
    if isinstance(v, str) and _SCI.match(v.strip()):
        return float(v)
    return v


def load_conf(path, overrides=()):
    """One of the reference's `config/*.yaml` files (or a user's copy) -> the configuration object of this package,
    with command-line overrides in the `section.key=value` form train.py hands to hydra (python/train.py:168-179:
    `--config-name <name> key=value ...`).  Keys the hot path does not read (data_path, valid.*, extraction.* ...) are
    kept as they are; keys missing from the file fall back to default.yaml's values."""
    import yaml
    with open(path) as f:
        user = _coerce(yaml.safe_load(f) or {})
    d = copy.deepcopy(DEFAULT)
    _merge(d, user)
    for item in overrides:
        key, sep, val = item.partition("=")
        if not sep:
            raise ValueError(f"override {item!r} is not of the form section.key=value")
        _set_path(d, key.strip(), _coerce(yaml.safe_load(val)))
    return _ns(d)


def check_supported(conf):
    """Raises NotImplementedError naming the key when the configuration takes a branch the CUDA path does not have
    (it never falls back to default.yaml's behaviour silently).  Engine.__init__ calls this; it needs no GPU."""
    def need(ok, what):
        if not ok:
            raise NotImplementedError(f"ndjir_b200 does not implement this branch of the reference: {what}")
    g = conf.geometric_network
    need(g.voxel.type in ("none", "voxel", "triplaneline"),
         f"geometric_network.voxel.type {g.voxel.type!r} (none / voxel / triplaneline)")
    need(len(g.skip_layers) <= 1 and g.geometric_init and g.act == "softplus",
         "geometric_network (one skip layer, geometric_init, softplus)")
    sb = conf.specular_brdf
    need(sb.model == "filament" and sb.remap and sb.sampling in ("importance", "uniform") and not sb.use_split_sum,
         "specular_brdf (filament, remap, importance / uniform sampling, no split sum)")
    need(conf.background_modeling, "background_modeling")
    need(not conf.use_wn, "use_wn")
    el, sv, ii = conf.environment_light_network, conf.soft_visibility_light_network, conf.implicit_illumination_network
    need(el.act_last == "softplus" and el.upper_bound <= 0 and el.channels == 1, "environment_light_network head")
    need(sv.act_last == "sigmoid" and sv.channels == 1 and sv.use_geometric_feature and sv.use_normal,
         "soft_visibility_light_network head")
    need(ii.act_last == "sigmoid" and not (ii.use_me and ii.use_me_on_specular) and ii.channels == 1,
         "implicit_illumination_network head")
    need(not conf.specular_reflectance_network.fixme, "specular_reflectance_network.fixme")
    need(conf.train.rgb_loss in ("l1", "l2"), "train.rgb_loss l1 / l2")
    need(conf.renderer.t_near_far_method in ("intersect_with_aabb", "intersect_with_r_sphere"),
         f"renderer.t_near_far_method {conf.renderer.t_near_far_method!r}")
