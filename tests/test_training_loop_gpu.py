"""GPU: the loop of python/train.py:124-148 end to end - Engine.train_step (fwd + bwd) followed by the fused optimizer
step of ndjir_b200.solver.Solvers, on ONE fixed batch.  Not a parity test (those pin every term and gradient): this one
checks the pieces as a user runs them - gradient signs, the per-tensor power-of-two scales of the split-fp16 engine
following weights that change every step, the refresh of the weight planes, graph replay - by asking for what training
must deliver: the loss goes down and everything stays finite."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,graphed", [("default", False), ("default", True), ("triplaneline", False)])
def test_loss_decreases_on_a_fixed_batch(kind, graphed):
    from test_engine_gpu import setup, dev
    from ndjir_b200.solver import Solvers
    conf, P, camloc, raydir, color_gt, rnd, eng, model = setup(kind, shape="small")
    conf.train.epoch = 200                      # warm-up term = 3 iterations, then the cosine schedule
    sol = Solvers(conf, eng)
    sol.set_parameters()
    eng.params.zero_grad()
    cam, ray, gt = dev(camloc), dev(raydir), dev(color_gt)
    drnd = {k: dev(v) for k, v in rnd.items()}
    hist = []
    for i in range(40):
        sol.update_learning_rate(i + 3)
        sol.cos_anneal_ratio = 0.0              # (a change of the annealing ratio would re-capture the graph every step)
        eng.params.pl_gain = 1.0
        if graphed:
            losses = eng.train_step_graphed(cam, ray, gt, drnd, cos_anneal_ratio=0.0)
        else:
            losses = eng.train_step(cam, ray, gt, drnd, cos_anneal_ratio=0.0, zero_grad=False)
        sol.step(loss=losses[0:1])
        hist.append(losses.detach().cpu().numpy().copy())
    torch.cuda.synchronize()
    hist = np.asarray(hist)
    assert np.isfinite(hist).all(), "a loss term went non-finite"
    assert np.isfinite(eng.params.data.cpu().numpy()).all()
    total, rgb = hist[:, 0], hist[:, 1]
    assert total[-5:].mean() < 0.92 * total[:3].mean(), (total[:3], total[-5:])      # (measured: 1.43 -> 1.21)
    assert rgb[-5:].mean() < rgb[:3].mean(), (rgb[:3], rgb[-5:])
    assert int(sol._flags.cpu().sum()) == 0, "an iteration was skipped (non-finite gradient or loss)"


def test_loop_over_a_device_resident_dataset():
    """The loop as train.py:124-148 runs it, with the data side on the device too: every step takes the next views and
    pixels from ndjir_b200.dataset.DeviceRaySource (one ndjir_train_batch launch: rays, colours, masks), runs
    Engine.train_step and the fused optimizer step.  The images are one flat colour, so the colour loss must fall even
    though every step sees other rays."""
    from test_engine_gpu import setup, dev
    from ndjir_b200 import scene
    from ndjir_b200.dataset import DeviceRaySource
    from ndjir_b200.solver import Solvers
    conf, P, camloc, raydir, color_gt, rnd, eng, model = setup("default", shape="small")
    tr = conf.train
    conf.train.epoch = 200
    conf.train.n_rays = 64                      # (the learning rate scales with batch_size * n_rays, solver.py:24-27)
    n, W, H = 7, 64, 48
    poses, intr, _ = scene.make_cameras(n, W=W, H=H, focal=90.0)
    images = np.broadcast_to(np.array([0.2, 0.5, 0.7], np.float32), (n, H, W, 3)).copy()
    src = DeviceRaySource(images, None, intr, poses, tr.n_rays, shuffle=True)
    sol = Solvers(conf, eng)
    sol.set_parameters()
    eng.params.zero_grad()
    before = eng.params.data.clone()
    hist = []
    for i in range(40):
        batch = src.next(tr.batch_size)
        drnd = {k: dev(v) for k, v in scene.make_randoms(conf, tr.batch_size, tr.n_rays, step=i).items()}
        sol.update_learning_rate(i + 3)
        sol.cos_anneal_ratio = 0.0
        eng.params.pl_gain = 1.0
        losses = eng.train_step(batch["camloc"], batch["raydir"], batch["color_gt"], drnd, cos_anneal_ratio=0.0,
                                zero_grad=False)
        sol.step(loss=losses[0:1])
        hist.append(losses.detach().cpu().numpy().copy())
    torch.cuda.synchronize()
    hist = np.asarray(hist)
    assert np.isfinite(hist).all() and np.isfinite(eng.params.data.cpu().numpy()).all()
    assert not torch.equal(before, eng.params.data)
    rgb = hist[:, 1]
    assert rgb[-8:].mean() < 0.95 * rgb[:4].mean(), (rgb[:4], rgb[-8:])
    assert int(sol._flags.cpu().sum()) == 0, "an iteration was skipped (non-finite gradient or loss)"
