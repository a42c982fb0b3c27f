"""Builds the REAL reference kernels into oracle/_ref/ (test infrastructure, never shipped or measured as ours).

The reference's native layer is 19 single-file pybind11 modules (reference Makefile:15-23:
``nvcc -shared --std=c++11 -fPIC -I./csrc `python -m pybind11 --includes` x.cu``).  They compile unmodified
from where they lie under /root/reference/csrc with this image's nvcc 12.9 + pybind11; we only add
``-gencode arch=compute_100a,code=sm_100a`` (the Makefile has no arch flag) and c++17 (pybind11 3.x needs it).
Default -fmad / no fast-math, exactly like the reference build.  Outputs go ONLY to oracle/_ref/
(git-ignored, not gpurun-ignored, so the modules travel to the GPU box).  No reference source is copied.

On the GPU box /root/reference does not exist: this script is a no-op there and the prebuilt modules are used.
Usage: python oracle/build_ref.py [--all]
"""
import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("NDJIR_REFERENCE", "/root/reference")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SCOPED = [
    "grid_feature/voxel_feature_cuda.cu",
    "grid_feature/lanczos_voxel_feature_cuda.cu",
    "grid_feature/voxel_hash_feature_cuda.cu",
    "grid_feature/triplane_feature_cuda.cu",
    "grid_feature/triline_feature_cuda.cu",
    "grid_feature/cosine_voxel_feature_cuda.cu",
    "grid_feature/cosine_triplane_feature_cuda.cu",
    "grid_feature/cosine_triline_feature_cuda.cu",
    "grid_feature/lanczos_triplane_feature_cuda.cu",
    "grid_feature/lanczos_triline_feature_cuda.cu",
    "grid_feature/lanczos_voxel_hash_feature_cuda.cu",
    "grid_feature/total_variation_loss_cuda.cu",
    "grid_feature/total_variation_loss_on_triplane_cuda.cu",
    "grid_feature/total_variation_loss_on_triline_cuda.cu",
    "grid_feature/total_variation_loss_on_voxel_hash_cuda.cu",
    "intersection/ray_aabb_intersection_cuda.cu",
    "intersection/ray_sphere_intersection_cuda.cu",
    "sampling/inverse_transform_cuda.cu",
    "activation/squareplus_cuda.cu",
]
EXTRA = []


def _includes():
    import pybind11
    return ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]


def _build_one(rel):
    src = os.path.join(REF, "csrc", rel)
    name = os.path.basename(rel)[:-3]
    out = os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return name, "cached"
    cmd = [NVCC, "-shared", "--std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
           "--compiler-options", "-fPIC", "-I" + os.path.join(REF, "csrc")] + _includes() + [src, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return name, "FAILED: " + r.stderr[-2000:]
    return name, "built"


def build(which=None, verbose=True):
    if not os.path.isdir(os.path.join(REF, "csrc")):
        if verbose:
            print(f"[oracle/build_ref] {REF} not present (GPU box): using prebuilt modules in {OUT}")
        return {}
    os.makedirs(OUT, exist_ok=True)
    rels = which if which is not None else SCOPED
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        res = dict(ex.map(_build_one, rels))
    if verbose:
        for k, v in res.items():
            print(f"[oracle/build_ref] {k}: {v}")
    bad = {k: v for k, v in res.items() if v.startswith("FAILED")}
    if bad:
        raise RuntimeError(f"reference modules failed to build: {bad}")
    return res


def available():
    """Names of the reference modules present in oracle/_ref."""
    if not os.path.isdir(OUT):
        return []
    suf = sysconfig.get_config_var("EXT_SUFFIX")
    return sorted(f[: -len(suf)] for f in os.listdir(OUT) if f.endswith(suf))


def load(name):
    """Imports a reference module from oracle/_ref (tests only)."""
    import importlib.util
    path = os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        raise ImportError(f"reference module {name} not built ({path})")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(SCOPED + EXTRA if "--all" in sys.argv else None)
