"""CPU: configuration loading (ndjir_b200.config.load_conf: the reference's YAML files + `section.key=value` overrides
as python/train.py:168-179 passes them to hydra) and the support check that names the key of an unimplemented branch.

Where the reference is mounted (this container; not the GPU box) every file of its config/ directory is loaded: the
built-in defaults must equal config/default.yaml key by key, the named BASELINE configs must equal their files, and each
file is either supported or rejected with the key it was rejected for."""
import glob
import os

import pytest

from ndjir_b200 import config, scene

REF_CONFIG = "/root/reference/config"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF_CONFIG), reason="reference not mounted")


def test_load_conf_merges_file_defaults_and_overrides(tmp_path):
    p = tmp_path / "mine.yaml"
    p.write_text("data_path: DTU/scan24\n"
                 "geometric_network:\n  voxel:\n    type: triplaneline\n    grid_size: 128\n"
                 "train:\n  n_rays: 256\n  roughness_prior_weight: 1e-5\n"
                 "valid:\n  n_rays: 4000\n")
    conf = config.load_conf(str(p), ["train.batch_size=2", "renderer.eps_normal=1e-8", "renderer.t_near_far_method=intersect_with_r_sphere",
                                     "train.tv_weight=0.0", "diffuse_brdf.entangle=false"])
    assert conf.geometric_network.voxel.type == "triplaneline" and conf.geometric_network.voxel.grid_size == 128
    assert conf.geometric_network.voxel.feature_size == 4                  # default.yaml's value
    assert conf.train.n_rays == 256 and conf.train.batch_size == 2 and conf.train.tv_weight == 0.0
    assert conf.renderer.t_near_far_method == "intersect_with_r_sphere" and conf.diffuse_brdf.entangle is False
    assert conf.train.roughness_prior_weight == 1e-5 and conf.renderer.eps_normal == 1e-8      # floats, not strings
    assert conf.data_path == "DTU/scan24" and conf.valid.n_rays == 4000    # keys outside the hot path are kept
    config.check_supported(conf)
    with pytest.raises(ValueError):
        config.load_conf(str(p), ["train.batch_size"])


@pytest.mark.parametrize("over,key", [
    (dict(geometric_network={"geometric_init": False}), "geometric_network"),
    (dict(specular_brdf={"model": "ue4", "remap": False}), "specular_brdf"),
    (dict(geometric_network={"voxel": {"type": "lanczos_voxel"}}), "lanczos_voxel"),
    (dict(use_wn=True), "use_wn"),
    (dict(implicit_illumination_network={"use_me_on_specular": True}), "implicit_illumination_network"),
    (dict(train={"rgb_loss": "huber"}), "rgb_loss"),
])
def test_unimplemented_branches_are_named(over, key):
    with pytest.raises(NotImplementedError, match=key):
        config.check_supported(config.make_conf("default", **over))


def test_switched_off_networks_own_no_parameters():
    from ndjir_b200 import nnabla_names
    full = {n for n, _ in nnabla_names.parameter_names(config.make_conf("default"))}
    no_ii = {n for n, _ in nnabla_names.parameter_names(
        config.make_conf("default", implicit_illumination_network={"use_me": False}))}
    no_pl = {n for n, _ in nnabla_names.parameter_names(
        config.make_conf("default", photogrammetric_light_network={"use_me": False}))}
    assert {n.split("/")[0] for n in full - no_ii} == {"implicit-illumination-network"}
    assert {n.split("/")[0] for n in full - no_pl} == {"photogrammetric-light-network"}
    assert "photogrammetric-light-network/gain" in full - no_pl


def _subset_equal(ours, theirs, path=""):
    bad = []
    for k, v in ours.items():
        if k not in theirs:
            bad.append(f"{path}{k}: missing from the file")
        elif isinstance(v, dict):
            bad += _subset_equal(v, theirs[k], f"{path}{k}.")
        elif v != theirs[k] and not (isinstance(v, (int, float)) and isinstance(theirs[k], (int, float))
                                     and float(v) == float(theirs[k])):
            bad.append(f"{path}{k}: {v!r} != {theirs[k]!r}")
    return bad


@needs_reference
def test_builtin_defaults_equal_the_reference_files():
    import yaml
    for name in ("default", "no_voxel", "triplaneline"):
        with open(os.path.join(REF_CONFIG, f"{name}.yaml")) as f:
            theirs = config._coerce(yaml.safe_load(f))
        ours = config.to_dict(config.make_conf(name))
        assert not _subset_equal(ours, theirs), (name, _subset_equal(ours, theirs))


@needs_reference
def test_every_shipped_config_is_supported_or_rejected_by_name():
    rejected = {"custom.yaml": "lanczos_voxel", "ue4.yaml": "specular_brdf"}
    files = sorted(glob.glob(os.path.join(REF_CONFIG, "*.yaml")))
    assert len(files) >= 23
    for f in files:
        conf = config.load_conf(f)
        base = os.path.basename(f)
        if base in rejected:
            with pytest.raises(NotImplementedError, match=rejected[base]):
                config.check_supported(conf)
        else:
            config.check_supported(conf)
            dims = scene.network_dims(conf)
            assert set(scene.active_nets(conf)) <= set(dims)
