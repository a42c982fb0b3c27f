"""Run under torchrun (N ranks, one GPU each): ray-sharded data parallelism gives every replica the gradient of the
single-process step over the union of all ranks' rays, with both grid-gradient exchanges (dense all-reduce, sparse
all-gather + replicated scatter).  Prints one line from rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ndjir_b200 import scene                      # noqa: E402
from ndjir_b200.engine import Engine             # noqa: E402
from test_engine_gpu import small_conf, dev      # noqa: E402


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    conf = small_conf("default")
    P = scene.init_params(conf, seed=313, grid_std=0.05)
    tr = conf.train
    B, R = tr.batch_size, tr.n_rays
    batches = [scene.make_batch(conf, step=r, B=B, R=R) for r in range(world)]
    rnds = [scene.make_randoms(conf, B, R, step=r) for r in range(world)]
    out = {}
    for mode in ("dense", "sparse"):
        eng = Engine(conf, world_size=world, process_group=dist.group.WORLD, grid_exchange=mode)
        eng.params.load_reference(P)
        camloc, raydir, gt = batches[rank]
        losses = eng.train_step(dev(camloc), dev(raydir), dev(gt), {k: dev(v) for k, v in rnds[rank].items()})
        torch.cuda.synchronize()
        out[mode] = (eng.params.export_reference("grad"), losses.cpu().numpy())
        # the same step replayed from ONE CUDA graph per rank, NCCL exchanges captured inside it
        lg = eng.train_step_graphed(dev(camloc), dev(raydir), dev(gt), {k: dev(v) for k, v in rnds[rank].items()})
        lg = eng.train_step_graphed(dev(camloc), dev(raydir), dev(gt), {k: dev(v) for k, v in rnds[rank].items()})
        torch.cuda.synchronize()
        gg = eng.params.export_reference("grad")
        out[mode + "_graph"] = max(relerr(gg[k], out[mode][0][k]) for k in gg if np.abs(out[mode][0][k]).max() > 0)
    if rank == 0:
        # single process over the union: views of all ranks stacked along B
        conf1 = small_conf("default")
        conf1.train.batch_size = B * world
        eng1 = Engine(conf1)
        eng1.params.load_reference(P)
        cat = lambda xs: np.concatenate(xs, axis=0)
        rnd1 = {k: dev(cat([r[k] for r in rnds])) for k in rnds[0]}
        l1 = eng1.train_step(dev(cat([b[0] for b in batches])), dev(cat([b[1] for b in batches])),
                             dev(cat([b[2] for b in batches])), rnd1)
        torch.cuda.synchronize()
        g1 = eng1.params.export_reference("grad")
        worst = {}
        for mode in ("dense", "sparse"):
            g, l = out[mode]
            worst[mode] = max(relerr(g[k], g1[k]) for k in g1 if np.abs(g1[k]).max() > 0)
            worst[mode + "_loss"] = abs(float(l[0]) - float(l1[0])) / abs(float(l1[0]))
            worst[mode + "_graph_vs_eager"] = out[mode + "_graph"]
        ok = all(v < 2e-4 for v in worst.values())
        print(("MULTI_GPU_OK " if ok else "MULTI_GPU_MISMATCH ") + str(worst), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)        # graphs hold NCCL work: skip the communicator teardown (it can block at exit)


if __name__ == "__main__":
    main()
