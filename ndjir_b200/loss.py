"""`total_loss` with the reference's call shape (python/loss.py:27) plus the forward/backward execution that the
reference triggers with loss.forward() / loss.backward() (python/train.py:135-140)."""
from .engine import LOSS_NAMES, get_engine


def total_loss(camloc, raydir, color_gt, obj_mask, cos_anneal_ratio, conf, rnd, backward=True, zero_grad=True):
    """Returns {name: 0-dim device tensor} for the reference's 10 loss terms (loss.py:180-192).  `rnd` holds the
    random tensors the reference draws inside the graph (scene.make_randoms).  With backward=True the parameter
    gradients are accumulated in get_engine(conf).params.grad / grid_grad.  obj_mask (B, R, 1) is read only when
    train.mask_weight > 0 (the BCE term of loss.py:108-116; 0 in every BASELINE config)."""
    eng = get_engine(conf)
    losses = eng.train_step(camloc, raydir, color_gt, rnd, cos_anneal_ratio=cos_anneal_ratio, backward=backward,
                            zero_grad=zero_grad, obj_mask=obj_mask)
    return {k: losses[i] for i, k in enumerate(LOSS_NAMES)}
