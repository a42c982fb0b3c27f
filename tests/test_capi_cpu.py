"""CPU: the C-ABI library builds/loads, exports every symbol include/ndjir_b200.h declares, and its host-only
entry points agree with the oracle.  No compute calls (there is no GPU here)."""
import ctypes
import os

import numpy as np
import pytest

from ndjir_b200 import _lib, compat
from oracle import cpu_ref as R


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 44
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(cdll, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_loader_fails_loudly_without_library(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NdjirError):
        _lib._Lib()


def test_compat_modules_mirror_reference_exports():
    """Same module names and function names as the reference's PYBIND11_MODULE blocks."""
    want = {
        "voxel_feature_cuda": ["query_on_voxel", "grad_query", "grad_feature", "grad_query_grad_grad_output",
                               "grad_query_grad_query", "grad_query_grad_feature", "grad_feature_grad_grad_output",
                               "grad_feature_grad_query"],
        "lanczos_voxel_feature_cuda": ["query_on_voxel", "grad_query", "grad_feature", "grad_query_grad_grad_output",
                                       "grad_query_grad_feature"],
        "voxel_hash_feature_cuda": ["hash_index", "voxel_hash_feature", "grad_query", "grad_feature",
                                    "grad_query_grad_grad_output", "grad_query_grad_feature"],
        "triplane_feature_cuda": ["query_on_triplane", "grad_query", "grad_feature", "grad_query_grad_grad_output",
                                  "grad_query_grad_feature"],
        "triline_feature_cuda": ["query_on_triline", "grad_query", "grad_feature", "grad_query_grad_grad_output",
                                 "grad_query_grad_feature"],
        "total_variation_loss_cuda": ["tv_loss_on_voxel", "tv_loss_on_voxel_backward"],
        "total_variation_loss_on_triplane_cuda": ["tv_loss_on_triplane", "tv_loss_on_triplane_backward"],
        "total_variation_loss_on_triline_cuda": ["tv_loss_on_triline", "tv_loss_on_triline_backward"],
        "ray_aabb_intersection_cuda": ["ray_aabb_intersection"],
        "ray_sphere_intersection_cuda": ["ray_sphere_intersection"],
        "inverse_transform_cuda": ["sample_uniform_directions", "sample_importance_directions"],
        "squareplus_cuda": ["forward", "backward"],
    }
    for mod, fns in want.items():
        m = compat.load(mod)
        for fn in fns:
            assert callable(getattr(m, fn)), (mod, fn)


@pytest.mark.parametrize("G0,gf,T0,L,D", [(16, 1.5, 2 ** 15, 16, 2), (2, 1.5, 2 ** 10, 4, 2), (4, 2.0, 2 ** 12, 6, 4),
                                          (3, 1.5, 2 ** 10, 5, 1)])
def test_hash_level_table_matches_oracle(G0, gf, T0, L, D):
    G = (ctypes.c_int * L)(); T = (ctypes.c_int * L)(); O = (ctypes.c_longlong * L)()
    _lib.call("ndjir_voxel_hash_level_table", G0, gf, T0, L, D, G, T, O)
    Gs, Ts, offs, total = R.hash_level_table(G0, gf, T0, L, D)
    assert list(G) == Gs and list(T) == Ts and list(O) == offs
    assert _lib.call("ndjir_voxel_hash_num_params", G0, gf, T0, L, D) == total
    # the Python-side formula of the reference wrapper (voxel_hash_feature.py:30-46) agrees for these growth factors
    py_total = 0
    for l in range(L):
        g = int(G0 * gf ** l)
        t = int(min(float(g) ** 3, T0))
        py_total += t * D + (t * D) % 8
    assert py_total == total


def test_bench_hash_table_size_matches_survey():
    assert _lib.call("ndjir_voxel_hash_num_params", 16, 1.5, 2 ** 15, 16, 2) == 953344


def test_empty_batches_are_noops_without_gpu():
    # n_points == 0 returns before any CUDA call, so this is safe on a CPU-only box
    _lib.call("ndjir_voxel_query_on_voxel", 0, None, None, None, [2, 2, 2], 4, [-1.] * 3, [1.] * 3, 0, 0)
    _lib.call("ndjir_triplane_query_on_triplane", 0, None, None, None, 8, 4, [-1.] * 3, [1.] * 3, 0, 0)
    _lib.call("ndjir_ray_aabb_intersection", 0, None, None, None, None, None, 0, 1, [-1.] * 3, [1.] * 3, 0)
    with pytest.raises(_lib.NdjirError):
        _lib.call("ndjir_set_option", "no_such_option", 1)


def test_binned_sweep_kernels_do_not_spill():
    """The brick-ordered sweeps run at 6-8 CTAs of 256 threads per SM (launch bounds cap them at 40 / 32 registers); a
    spill there costs 45 % (1.16 -> 1.68 ms measured when an experiment's extra arguments pushed the gather over)."""
    import os
    import re
    import subprocess
    from ndjir_b200 import build
    src = os.path.join(build.CSRC, "voxel_binned.cu")
    r = subprocess.run([build.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                        "-I" + os.path.join(build.ROOT, "include"), "-Xptxas", "-v", "-c", src, "-o", os.devnull],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    blocks = re.split(r"ptxas info\s+: Compiling entry function '", r.stderr)[1:]
    checked = 0
    for b in blocks:
        name = b.split("'")[0]
        if re.search(r"gather_kernelILi\dELb[01]ELb1E", name):
            continue      # the 256-bit z-pair load experiment (off by default, measured slower)
        if "gather_kernel" in name or "scatter_kernel" in name:
            m = re.search(r"(\d+) bytes spill stores", b)
            assert m and int(m.group(1)) == 0, f"{name}: {m.group(0) if m else 'no ptxas report'}"
            checked += 1
    assert checked >= 6


def test_set_option_keys():
    """Every documented option key is accepted (host-only call), unknown keys are an argument error."""
    from ndjir_b200 import _lib
    defaults = {"scatter_aggregate": 0, "mlp_tensor_cores": 1, "mlp_cta_pair": 0, "mlp_presplit": 1, "mlp_fused_colsum": 1, "mlp_dbg": 0,
                "mlp_mask_hi": 0, "mlp_h_dbg": 0, "mlp_h_pair": 0, "voxel_binned": -1, "voxel_bin_mb": 16, "voxel_pair256": 0, "voxel_tma": 0, "voxel_tma_bx": 16, "voxel_tma_l2": 1, "voxel_tma_dbg": 0, "hash_coarse_private": 1}
    for k, v in defaults.items():
        _lib.call("ndjir_set_option", k, v)
    import pytest
    with pytest.raises(_lib.NdjirError):
        _lib.call("ndjir_set_option", "no_such_option", 1)


def test_compat_modules_cover_every_reference_export():
    """Boundary coverage: every `m.def("<name>", ...)` of the reference's 19 pybind11 modules (csrc/**/<module>.cu) exists
    with the same name in the stand-in module ndjir_b200/compat/<module>.py.  Reads the reference where it is mounted
    (this container); skipped on machines without it."""
    import glob
    import importlib
    import os
    import re
    import pytest
    ref = os.environ.get("NDJIR_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "csrc")):
        pytest.skip("reference checkout not mounted")
    from ndjir_b200 import compat
    files = {os.path.basename(p)[:-3]: p for p in glob.glob(os.path.join(ref, "csrc", "*", "*_cuda.cu"))}
    assert len(files) == 19 and sorted(files) == sorted(compat.MODULES), sorted(set(files) ^ set(compat.MODULES))
    for name, path in files.items():
        text = re.sub(r"//[^\n]*", "", open(path).read())          # commented-out exports do not count
        exports = re.findall(r'm\.def\(\s*"(\w+)"', text)
        assert exports, name
        mod = importlib.import_module(f"ndjir_b200.compat.{name}")
        missing = [e for e in exports if not callable(getattr(mod, e, None))]
        assert not missing, f"{name}: stand-in lacks {missing}"
