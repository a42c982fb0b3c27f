"""Experiment: is the train step launch-gap bound?  Captures one Engine.train_step in a CUDA graph and compares the
replay time with the eager step (default.yaml, 1 GPU)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndjir_b200 import scene
from ndjir_b200.config import make_conf
from ndjir_b200.engine import Engine

conf = make_conf("default")
eng = Engine(conf)
eng.params.load_reference(scene.init_params(conf, seed=313))
eng.params.init_grid_on_device(scene.grid_shapes(conf), std=1e-3, seed=313)
tr = conf.train
camloc, raydir, color_gt = scene.make_batch(conf, step=0)
rnd = scene.make_randoms(conf, tr.batch_size, tr.n_rays, step=0)
item = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in {"camloc": camloc, "raydir": raydir, "color_gt": color_gt, **rnd}.items()}

def step():
    return eng.train_step(item["camloc"], item["raydir"], item["color_gt"], item, cos_anneal_ratio=0.0)

for _ in range(3):
    step()
torch.cuda.synchronize()
def timeit(fn, n=8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_host / n * 1e3
ms, host_ms = timeit(step)
print(f"eager: {ms:.2f} ms per step on the device, {host_ms:.2f} ms of host time to enqueue it, {eng.n_launches} launches so far")
l_eager = step().clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    l_graph = step()
torch.cuda.synchronize()
ms_g, host_g = timeit(g.replay)
print(f"graph replay: {ms_g:.2f} ms per step, {host_g:.3f} ms of host time")
print("loss eager", l_eager[:4].tolist(), "graph", l_graph[:4].tolist())
