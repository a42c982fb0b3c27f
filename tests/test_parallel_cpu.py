"""Host-side logic of the ray-sharded data-parallel path on CPU: world_size-2 gloo processes exchange the mask sum
and the gradient buffers exactly like the NCCL path does on the GPUs (ndjir_b200/parallel.py), and the parameter
store round-trips the reference layout."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ndjir_b200 import scene
from ndjir_b200.config import make_conf
from ndjir_b200.engine import ParamStore
from ndjir_b200.parallel import allgather_rows, allreduce_gradients, allreduce_mask_sum, shard_rays


def small():
    return make_conf("default", geometric_network={"feature_size": 64, "voxel": {"grid_size": 8}},
                     base_color_network={"feature_size": 32}, train={"batch_size": 2, "n_rays": 8})


@pytest.mark.parametrize("kind", ["default", "triplaneline", "no_voxel"])
def test_param_store_round_trips_reference_layout(kind):
    conf = make_conf(kind, geometric_network={"feature_size": 128, "voxel": {"grid_size": 8}})
    P = scene.init_params(conf, seed=1)
    ps = ParamStore(conf, "cpu")
    ps.load_reference(P)
    ex = ps.export_reference("data")
    for net in scene.NET_ORDER:
        for i, (W, b) in enumerate(P[net]):
            assert np.array_equal(ex[f"{net}.W{i}"], W), (net, i)
            assert np.array_equal(ex[f"{net}.b{i}"], b), (net, i)
    assert ex["geo_gain"][0] == P["geo_gain"][0]
    for k, v in P["grid"].items():
        assert np.array_equal(ex[f"grid.{k}"], v)
    # head inputs are stored as [feature | x | normal | ...]: the feature rows of the first layer come first
    Df = conf.geometric_network.feature_size
    L0 = ps.nets["ii"][0]
    W0 = ps.data[L0.w_off:L0.w_off + L0.K * L0.ldw].reshape(L0.K, L0.ldw).numpy()
    assert np.array_equal(W0[:Df, :L0.N], P["ii"][0][0][3:3 + Df])
    assert np.array_equal(W0[Df:Df + 3, :L0.N], P["ii"][0][0][0:3])


def test_shard_rays_partitions_the_batch():
    conf = small()
    camloc, raydir, color_gt = scene.make_batch(conf, B=2, R=8)
    rd, gt = torch.from_numpy(raydir), torch.from_numpy(color_gt)
    parts = [shard_rays(rd, gt, r, 4) for r in range(4)]
    assert torch.equal(torch.cat([p[0] for p in parts], dim=1), rd)
    assert torch.equal(torch.cat([p[1] for p in parts], dim=1), gt)
    with pytest.raises(ValueError):
        shard_rays(rd, gt, 0, 3)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    conf = small()
    ps = ParamStore(conf, "cpu")
    ps.grid_grad["voxel"] = torch.zeros(8, 8, 8, 4)
    gen = torch.Generator().manual_seed(100 + rank)
    ps.grad.copy_(torch.randn(ps.grad.shape, generator=gen))
    ps.grid_grad["voxel"].copy_(torch.randn(8, 8, 8, 4, generator=gen))
    local = (ps.grad.clone(), ps.grid_grad["voxel"].clone())
    ms = torch.tensor([3.0 + rank])
    allreduce_mask_sum(ms)
    allreduce_gradients(ps)
    # sparse exchange: every rank ends with the rows of all ranks, in rank order
    rows = torch.full((5, 3), float(rank))
    allr = allgather_rows(torch.empty(5 * world, 3), rows)
    assert torch.equal(allr, torch.cat([torch.full((5, 3), float(r)) for r in range(world)]))
    flat_only = ps.grad.clone()
    grid_before = ps.grid_grad["voxel"].clone()
    allreduce_gradients(ps, include_grid=False)          # sparse mode: the grid gradient is not all-reduced
    assert torch.equal(ps.grid_grad["voxel"], grid_before) and torch.allclose(ps.grad, flat_only * world)
    ps.grad.copy_(flat_only)
    q.put((rank, local[0].numpy(), local[1].numpy(), ps.grad.numpy().copy(), ps.grid_grad["voxel"].numpy().copy(),
           float(ms)))
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world_size_2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_flat = res[0][1] + res[1][1]
    want_grid = res[0][2] + res[1][2]
    for r in res:
        assert np.allclose(r[3], want_flat, rtol=0, atol=1e-6)
        assert np.allclose(r[4], want_grid, rtol=0, atol=1e-6)
        assert r[5] == 7.0
