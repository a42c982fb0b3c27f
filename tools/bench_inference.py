"""Inference measurement (BASELINE.json config 5): full-frame 1600 x 1200 render in chunks of valid.n_rays = 4000 rays
(renderer.render_image, the loop of the reference's python/renderer.py:212-272) and the 512^3 SDF lattice query of
extract_by_mc.py (renderer.sdf_volume), default.yaml networks with the 512^3 x 4 voxel grid, synthetic camera.
One JSON object to stdout / --out.  With torchrun the chunks / x-slabs are dealt to the ranks (no collective: every
rank keeps its part), time = max over ranks.

  python tools/bench_inference.py [--lattice 512] [--out gpurun_out/bench_inference.json]"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndjir_b200 import scene, renderer
from ndjir_b200.config import make_conf
from ndjir_b200.engine import get_engine

ap = argparse.ArgumentParser()
ap.add_argument("--lattice", type=int, default=512)
ap.add_argument("--width", type=int, default=1600)
ap.add_argument("--height", type=int, default=1200)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_inference.json"))
args = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
conf = make_conf("default")
eng = get_engine(conf)
eng.params.load_reference(scene.init_params(conf, seed=313))
eng.params.init_grid_on_device(scene.grid_shapes(conf), std=1e-3, seed=313)
poses, intr, _ = scene.make_cameras(49)
W, H = args.width, args.height


def timed(fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return r, dt


pg = dist.group.WORLD if world > 1 else None
# warm-up: buffers, kernel attributes, the captured per-chunk graph (same chunk size and frame as the timed call)
renderer.render_image(poses[1], intr[1], (W, H), conf, n_rays=4000, rank=rank, world_size=world, process_group=pg)
img, t_img = timed(lambda: renderer.render_image(poses[0], intr[0], (W, H), conf, n_rays=4000, rank=rank,
                                                 world_size=world, process_group=pg))
renderer.sdf_volume(conf, 16)
vol, t_vol = timed(lambda: renderer.sdf_volume(conf, args.lattice, batch_size=1 << 22, rank=rank, world_size=world,
                                               process_group=pg, gather=True))
if rank == 0:
    G = args.lattice
    out = {"config": "default.yaml, 512^3 x 4 voxel grid, synthetic DTU-shaped camera; chunks / x-planes dealt to the "
                     "ranks, results summed onto rank 0 inside the timed region", "n_gpus": world,
           "render_image": {"resolution": [W, H], "rays": W * H, "chunk_rays": 4000, "seconds": t_img,
                            "rays_per_s": W * H / t_img, "finite": bool(torch.isfinite(img).all())},
           "sdf_lattice": {"grid": G, "points": G ** 3, "seconds": t_vol, "points_per_s": G ** 3 / t_vol,
                           "finite": bool(torch.isfinite(vol).all())}}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)      # captured graphs: skip the communicator teardown
