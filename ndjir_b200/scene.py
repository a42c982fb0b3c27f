"""Host-side inputs of the hot path: parameter layout + initialisation, synthetic DTU-shaped rays and the explicit
random tensors.  All numpy, deterministic, shared by the CUDA path, the tests and the benchmark.

Parameter inventory and initialisers follow the reference: geometric initialisation
python/network.py:36-56,190-214; nnabla `PF.affine` default (Glorot-uniform W, zero b) for every other network;
grid features N(0, 1e-3) python/grid_feature/voxel_feature.py:155-163; ray generation python/helper.py:44-73; pixel
sampling python/dataset.py:170,180-189.  nnabla's RNG call order cannot be reproduced without nnabla, so the streams
are ours (numpy RandomState(313)); every consumer receives the SAME arrays, which is what parity needs.
"""
import numpy as np

from .config import grid_channels

f32 = np.float32


def pe_dim(in_dim, bands):
    return in_dim * (1 + 2 * bands) if bands > 0 else in_dim


def network_dims(conf):
    """{net: [(in, out), ...]} for every MLP on the path (reference python/network.py)."""
    g = conf.geometric_network
    D, L = g.feature_size, g.layers
    din = pe_dim(3, g.pe_bands) + grid_channels(conf)
    geo = []
    h = din
    for l in range(L):
        if l == L - 1:
            do = D + 1
        elif (l + 1) in g.skip_layers:
            do = D - din
        else:
            do = D
        geo.append((h, do))
        h = D if (l + 1) in g.skip_layers else do
    Df = D

    def head(cfg, din_, dout):
        dims, hh = [], din_
        for _ in range(cfg.layers - 1):
            dims.append((hh, cfg.feature_size))
            hh = cfg.feature_size
        dims.append((hh, dout))
        return dims
    bc = conf.base_color_network
    el = conf.environment_light_network
    sv = conf.soft_visibility_light_network
    ii = conf.implicit_illumination_network
    pl = conf.photogrammetric_light_network
    ro = conf.roughness_network
    sp = conf.specular_reflectance_network
    bg = conf.background_network
    nets = {
        "geo": geo,
        "bc": head(bc, 3 + Df, 3),
        "el": head(el, pe_dim(3, el.pe_bands), el.channels),
        "sv": head(sv, 3 + pe_dim(3, sv.pe_bands) + Df + 3, sv.channels),
        "ii": head(ii, 3 + Df + 3, ii.channels),
        "pl": head(pl, 3 + pe_dim(3, pl.pe_bands) + Df + 3 + (1 if pl.use_inverse_distance else 0), pl.channels),
        "ro": head(ro, 3 + Df + 3, 2),
        "sp": head(sp, 3 + Df + 3, 2 * sp.channels),
    }
    d0 = pe_dim(4, bg.pe_bands0)
    bg0, hh = [], d0
    for _ in range(bg.layers0 - 1):
        bg0.append((hh, bg.feature_size0)); hh = bg.feature_size0
    bg0.append((hh, bg.feature_size0 + 1))
    d1 = 4 + bg.feature_size0 + 3 + pe_dim(3, bg.pe_bands1)
    bg1, hh = [], d1
    for _ in range(bg.layers1 - 1):
        bg1.append((hh, bg.feature_size1)); hh = bg.feature_size1
    bg1.append((hh, 3))
    nets["bg0"], nets["bg1"] = bg0, bg1
    return nets


NET_ORDER = ["geo", "bc", "el", "sv", "ii", "pl", "ro", "sp", "bg0", "bg1"]


def active_nets(conf):
    """The networks the reference actually creates: implicit_illumination_network returns a constant 0 without touching a
    parameter when use_me is false (network.py:308-309, config/no_implicit_illumination.yaml) and the photogrammetric
    light network is never called (renderer.py:161, config/no_lightp.yaml)."""
    off = set()
    if not conf.implicit_illumination_network.use_me:
        off.add("ii")
    if not conf.photogrammetric_light_network.use_me:
        off.add("pl")
    return [n for n in NET_ORDER if n not in off]


def init_params(conf, seed=313, grid_std=1e-3, dtype=f32):
    """Returns {"geo": [(W,b),...], ..., "geo_gain": (1,), "pl_gain": (1,), "grid": {...}} as numpy arrays.
    W has shape (in, out) like nnabla's affine (x @ W + b)."""
    rng = np.random.RandomState(seed)
    dims = network_dims(conf)
    g = conf.geometric_network
    D, L = g.feature_size, g.layers
    din = dims["geo"][0][0]
    p = {}
    geo = []
    for l, (di, do) in enumerate(dims["geo"]):
        b = np.zeros(do)
        if not g.geometric_init:
            lim = np.sqrt(6.0 / (di + do))
            W = rng.uniform(-lim, lim, (di, do))
        elif l == 0:
            W = np.sqrt(2.0 / D) * rng.randn(di, do)
            W[3:, :] = 0.0
        elif l in g.skip_layers:
            W = np.sqrt(2.0 / (D - din)) * rng.randn(di, do)
            W[-din:, :] = 0.0
        elif l == L - 1:
            W = np.sqrt(2.0 / do) * rng.randn(di, do)
            W[:, 0] = np.sqrt(np.pi / di) * np.ones(di) + rng.randn(di) * 1e-4
            b[:] = -g.initial_sphere_radius
        else:
            W = np.sqrt(2.0 / do) * rng.randn(di, do)
        geo.append((W.astype(dtype), b.astype(dtype)))
    p["geo"] = geo
    for net in NET_ORDER[1:]:
        layers = []
        for (di, do) in dims[net]:
            lim = np.sqrt(6.0 / (di + do))
            layers.append((rng.uniform(-lim, lim, (di, do)).astype(dtype), np.zeros(do, dtype=dtype)))
        p[net] = layers
    p["geo_gain"] = np.asarray([conf.train.sigmoid_gain], dtype=dtype)
    p["pl_gain"] = np.asarray([conf.train.sigmoid_gain_lv_start], dtype=dtype)
    v = g.voxel
    grid = {}
    if v.type == "voxel":
        G = v.grid_size
        grid["voxel"] = (rng.randn(G, G, G, v.feature_size) * grid_std).astype(dtype) if G <= 128 else None
    elif v.type == "triplaneline":
        G = v.grid_size
        grid["triplane"] = (rng.randn(3, G, G, v.feature_size) * grid_std).astype(dtype) if G <= 512 else None
        grid["triline"] = (rng.randn(3, G, v.feature_size) * grid_std).astype(dtype)
    p["grid"] = grid   # None entries: too large for a host copy, create on the device instead
    return p


def grid_shapes(conf):
    v = conf.geometric_network.voxel
    G, D = v.grid_size, v.feature_size
    if v.type == "voxel":
        return {"voxel": (G, G, G, D)}
    if v.type == "triplaneline":
        return {"triplane": (3, G, G, D), "triline": (3, G, D)}
    return {}


# ----------------------------------------------------------------------------------------------------
# synthetic DTU-shaped rig: 49 pinhole cameras 1600x1200, f ~ 2892 px, on a sphere of
# radius 2.5-3 around the unit scene, looking at the origin
# ----------------------------------------------------------------------------------------------------
def make_cameras(n_views=49, W=1600, H=1200, focal=2892.0, seed=313):
    rng = np.random.RandomState(seed)
    poses = np.zeros((n_views, 4, 4))
    intr = np.zeros((n_views, 3, 3))
    for i in range(n_views):
        # cameras on the upper hemisphere, like a DTU arc
        th = np.arccos(rng.uniform(0.2, 0.95))
        ph = rng.uniform(0, 2 * np.pi)
        r = rng.uniform(2.5, 3.0)
        c = r * np.array([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        zc = -c / np.linalg.norm(c)                      # camera looks at the origin (+z camera axis)
        up = np.array([0.0, 0.0, 1.0])
        xc = np.cross(up, zc); xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        poses[i, :3, 0], poses[i, :3, 1], poses[i, :3, 2], poses[i, :3, 3] = xc, yc, zc, c
        poses[i, 3, 3] = 1.0
        intr[i] = [[focal, 0, W / 2], [0, focal, H / 2], [0, 0, 1]]
    return poses, intr, (W, H)


def generate_raydir_camloc(pose, intrinsic, xy):
    """Pixel -> world ray (reference python/helper.py:44-73). pose (B,4,4), intrinsic (B,3,3), xy (B,R,2)."""
    B, R, _ = xy.shape
    R_c2w = pose[:, np.newaxis, :3, :3]
    camloc = pose[:, :3, 3]
    K_inv = np.linalg.inv(intrinsic)[:, np.newaxis, :, :]
    pix = np.concatenate([xy, np.ones((B, R, 1))], axis=-1)[..., np.newaxis]
    world = np.matmul(R_c2w, np.matmul(K_inv, pix)).reshape(B, R, 3)
    raydir = world / np.sqrt(np.sum(world ** 2, axis=-1, keepdims=True))
    return raydir.astype(f32), camloc.astype(f32)


def make_batch(conf, step=0, n_views=49, B=None, R=None, seed=313):
    """One training batch: (camloc (B,3), raydir (B,R,3), color_gt (B,R,3)) for `step`."""
    B = B or conf.train.batch_size
    R = R or conf.train.n_rays
    poses, intr, (W, H) = make_cameras(n_views)
    rng = np.random.RandomState(seed + step)
    views = rng.permutation(n_views)[:B]
    pix = rng.randint(0, W * H, (B, R))
    xy = np.stack([pix % W, pix // W], axis=-1).astype(np.float64)
    raydir, camloc = generate_raydir_camloc(poses[views], intr[views], xy)
    color_gt = rng.rand(B, R, 3).astype(f32)
    return camloc, raydir, color_gt


def make_randoms(conf, B, R, step=0):
    """The stochastic tensors the reference draws with F.rand/F.randn (python/loss.py:40-41, renderer.py:97-98,
    131-132,191), as explicit inputs, seeded with the config's seed numbers."""
    r = conf.renderer
    N0, Nb, nt = r.n_samples0, r.n_bg_samples, r.n_thetas
    N = N0 + r.n_upsamples * r.n_samples1

    def gen(seed):
        return np.random.default_rng(seed + 1000003 * step)
    out = {
        "stratified": gen(r.stratified_sample_seed).random((B, R, N0, 1), dtype=f32),
        "background": (gen(r.background_sample_seed).random((B, R, Nb + 1, 1), dtype=f32) * f32(1 - 1e-5) + f32(1e-5)),
        "diffuse_cdf_the": gen(r.diffuse_cdf_the_seed).random((B, R, nt), dtype=f32),
        "diffuse_cdf_phi": gen(r.diffuse_cdf_phi_seed).random((B, R, 2 * nt), dtype=f32),
        "specular_cdf_the": gen(r.specular_cdf_the_seed).random((B, R, nt), dtype=f32),
        "specular_cdf_phi": gen(r.specular_cdf_phi_seed).random((B, R, 2 * nt), dtype=f32),
        "perturb": gen(conf.train.base_color_perturb_seed + 7).standard_normal((B, R, N, 3), dtype=f32),
    }
    return out
