"""Per-kernel time table of one eager train step (torch.profiler / CUPTI) and the product table of the engine's own
CUDA-event instrumentation.  python tools/exp/step_profile.py [--config default|triplaneline|no_voxel] [--mlp h16|fp32]"""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ndjir_b200 import scene  # noqa: E402
from ndjir_b200.config import make_conf  # noqa: E402
from ndjir_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="default")
ap.add_argument("--mlp", default="h16")
args = ap.parse_args()
conf = make_conf(args.config)
eng = Engine(conf, mlp=args.mlp)
eng.params.load_reference(scene.init_params(conf, seed=313))
eng.params.init_grid_on_device(scene.grid_shapes(conf), std=1e-3, seed=313)
tr = conf.train
B, R = tr.batch_size, tr.n_rays
camloc, raydir, color_gt = scene.make_batch(conf, step=0)
rnd = scene.make_randoms(conf, B, R, step=0)
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()  # noqa: E731
item = {k: d(v) for k, v in dict(camloc=camloc, raydir=raydir, color_gt=color_gt, **rnd).items()}


def step():
    return eng.train_step(item["camloc"], item["raydir"], item["color_gt"], item, cos_anneal_ratio=0.0)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lg = eng.train_step_graphed(item["camloc"], item["raydir"], item["color_gt"], {k: v for k, v in item.items() if k not in ("camloc", "raydir", "color_gt")})
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    eng.train_step_graphed(item["camloc"], item["raydir"], item["color_gt"], {k: v for k, v in item.items() if k not in ("camloc", "raydir", "color_gt")})
e1.record()
torch.cuda.synchronize()
print(f"graphed step: {e0.elapsed_time(e1) / 3:.3f} ms")

from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name
        k = k.replace("(anonymous namespace)::", "")
        for pre in ("void ndjir::", "ndjir::", "void "):
            if k.startswith(pre):
                k = k[len(pre):]
                break
        k = k.split("(")[0][:90]
        agg[k][0] += 1
        agg[k][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f"kernel time total {tot / 1e3:.3f} ms, {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t / 1e3:8.3f} ms {n:5d}  {k}")

eng.profile = True
eng.prof_events = []
step()
torch.cuda.synchronize()
shapes = collections.defaultdict(lambda: [0, 0.0, 0.0])
for a, b, fl, name in eng.prof_events:
    t = a.elapsed_time(b)
    shapes[name][0] += 1
    shapes[name][1] += t
    shapes[name][2] += fl
print("products:")
for k, (n, t, fl) in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:8.3f} ms {n:4d}  {k:<34s} {fl / t / 1e9:8.1f} TFLOP/s")
print(f"products total {sum(v[1] for v in shapes.values()):.3f} ms")
