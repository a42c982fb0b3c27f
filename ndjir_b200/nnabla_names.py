"""Names under which the reference stores the parameters of the hot path (nnabla parameter scopes), so that a
parameter dictionary / file written by the reference can be loaded here and vice versa.

The names follow the scopes opened in the reference's python/network.py:
  geometric network      `geometric-network/affine-{l:02d}/affine/{W,b}`, last layer `affine-last` when
                         geometric_init (network.py:190-214), `affine-{L-1:02d}` otherwise (:181-186); the trainable
                         `geometric-network/gain` (:226-228); grid tables `geometric-network/{voxel,triplane,triline}_feature/F`
                         (python/grid_feature/voxel_feature.py:143-163)
  heads                  `<net>/affine-{l:02d}/affine/{W,b}` (network.py:235-424)
  roughness / specular   hidden layers are named with l - 1: `affine--1`, `affine-00`, `affine-01`, then `affine-03`
                         (the off-by-one of network.py:451,454,495,498 - SURVEY.md q16)
  photogrammetric light  plus the scheduled, non-trainable `photogrammetric-light-network/gain` (network.py:418-420)
  background             `background-network/geometric-network/...` and `background-network/lighting-network/...` (:537-559)
tests/golden/make_golden.py executes the reference's network.py against a parameter registry filled through this
table: a name the reference asks for that the table does not produce fails the golden generation.
"""
import numpy as np

from .scene import NET_ORDER, active_nets, network_dims

SCOPES = {
    "geo": "geometric-network", "bc": "base-color-network", "el": "environment-light-network",
    "sv": "soft-visibility-light-network", "ii": "implicit-illumination-network",
    "pl": "photogrammetric-light-network", "ro": "roughness-network", "sp": "specular-reflectance-network",
    "bg0": "background-network/geometric-network", "bg1": "background-network/lighting-network",
}
GRID_SCOPES = {"voxel": "voxel_feature", "triplane": "triplane_feature", "triline": "triline_feature"}


def layer_scope(conf, net, l, n_layers):
    """Scope name of layer l (0-based) of network `net` with n_layers affine layers."""
    last = l == n_layers - 1
    if net == "geo":
        if last and conf.geometric_network.geometric_init:
            return "affine-last"
        return f"affine-{l:02d}"
    if net in ("ro", "sp") and not last:
        return f"affine-{l - 1:02d}"          # sic: "affine--1", "affine-00", "affine-01"
    return f"affine-{l:02d}"


def parameter_names(conf):
    """[(nnabla name, key)] in registration order; key = (net, layer, "W" | "b"), ("geo_gain",), ("pl_gain",) or
    ("grid", part)."""
    dims = network_dims(conf)
    out = []
    for net in active_nets(conf):          # (networks the configuration switches off own no parameters)
        n = len(dims[net])
        for l in range(n):
            base = f"{SCOPES[net]}/{layer_scope(conf, net, l, n)}/affine"
            out.append((f"{base}/W", (net, l, "W")))
            out.append((f"{base}/b", (net, l, "b")))
        if net == "geo":
            v = conf.geometric_network.voxel.type
            parts = {"voxel": ["voxel"], "triplaneline": ["triplane", "triline"]}.get(v, [])
            for part in parts:
                out.append((f"{SCOPES['geo']}/{GRID_SCOPES[part]}/F", ("grid", part)))
            out.append((f"{SCOPES['geo']}/gain", ("geo_gain",)))
        if net == "pl":
            out.append((f"{SCOPES['pl']}/gain", ("pl_gain",)))
    return out


def to_nnabla(conf, P):
    """Parameter dictionary of scene.init_params / ParamStore.export_reference layout -> {nnabla name: array}."""
    out = {}
    for name, key in parameter_names(conf):
        if key[0] == "grid":
            v = P.get("grid", {}).get(key[1])
            if v is not None:
                out[name] = v
        elif key[0] in ("geo_gain", "pl_gain"):
            out[name] = P[key[0]]
        else:
            net, l, which = key
            out[name] = P[net][l][0 if which == "W" else 1]
    return out


def from_nnabla(conf, params):
    """{nnabla name: array} -> the dictionary layout ParamStore.load_reference takes.  Missing names raise KeyError."""
    dims = network_dims(conf)
    P = {net: [[None, None] for _ in dims[net]] for net in NET_ORDER}
    for net in set(NET_ORDER) - set(active_nets(conf)):      # switched off: no entries in the file, zeros in the store
        P[net] = [[np.zeros((di, do), np.float32), np.zeros(do, np.float32)] for di, do in dims[net]]
    if "pl" not in active_nets(conf):
        P["pl_gain"] = np.asarray([conf.train.sigmoid_gain_lv_start], np.float32)
    P["grid"] = {}
    for name, key in parameter_names(conf):
        if key[0] == "grid":
            if name in params:
                P["grid"][key[1]] = params[name]
            continue
        v = params[name]
        if key[0] in ("geo_gain", "pl_gain"):
            P[key[0]] = v
        else:
            net, l, which = key
            P[net][l][0 if which == "W" else 1] = v
    for net in NET_ORDER:
        P[net] = [tuple(x) for x in P[net]]
    return P
