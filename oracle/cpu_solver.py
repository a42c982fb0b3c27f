"""CPU restatement (numpy) of the reference's optimizer step and schedules: python/solver.py:20-119 driving nnabla's
S.Adam, in the call order of python/train.py:135-148.

TEST INFRASTRUCTURE ONLY (see oracle/cpu_ref.py header): imported by tests/, never by ndjir_b200/.

Schedules and call order are PINNED on the reference's own solver.py, executed through tests/golden/nnabla_standin.py
(tests/golden/solver.npz, tests/test_render_golden.py::test_solver_schedules_and_adam_equal_reference).  The Adam
arithmetic itself lives in nnabla (S.Adam), which is not installable here: it is nnabla's published update rule
(nnabla.solvers.Adam docs, v1.29), restated in the stand-in and here:
    m_t = beta1 m_{t-1} + (1 - beta1) g_t,   v_t = beta2 v_{t-1} + (1 - beta2) g_t^2,
    alpha_t = alpha sqrt(1 - beta2^t) / (1 - beta1^t),   w_t = w_{t-1} - alpha_t m_t / (sqrt(v_t) + eps),
defaults beta1 = 0.9, beta2 = 0.999, eps = 1e-8; Solver.weight_decay(r): g += r w; Solver.check_inf_or_nan_grad():
any(isinf(g) | isnan(g)).  All in float32 like the CUDA solver.
"""
import math

import numpy as np

F = np.float32


class Adam:
    """One nnabla S.Adam instance over a dict name -> (w, g) of float32 arrays (updated in place)."""

    def __init__(self, alpha=0.0, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = 0
        self.state = {}

    def set_learning_rate(self, lr):
        self.alpha = lr

    def weight_decay(self, params, rate):            # solver.py:48-50
        for w, g in params.values():
            g += F(rate) * w

    def check_inf_or_nan_grad(self, params):
        return any(not np.all(np.isfinite(g)) for _, g in params.values())

    def update(self, params):
        self.t += 1
        b1, b2 = self.beta1, self.beta2
        alpha_t = F(self.alpha * math.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t))
        for k, (w, g) in params.items():
            m, v = self.state.setdefault(k, (np.zeros_like(w), np.zeros_like(w)))
            m[...] = F(b1) * m + (F(1) - F(b1)) * g
            v[...] = F(b2) * v + (F(1) - F(b2)) * g * g
            w -= alpha_t * m / (np.sqrt(v) + F(self.eps))


def learning_rates(conf):
    """solver.py:23-27: both base rates scale with the number of rays per step relative to 1 x 512."""
    tr = conf.train
    scale = (tr.batch_size * tr.n_rays) / 512.0
    return tr.base_learning_rate_weight * scale, tr.base_learning_rate_feat * scale


def compute_learning_rate(conf, i, lr):
    """solver.py:82-98: linear warm-up over int(epoch * warmup_term_ratio) epochs, then a cosine from lr down to
    lr * learning_rate_end_ratio at the last epoch (continuous at the end of the warm-up)."""
    tr = conf.train
    E = tr.epoch
    wu = int(E * tr.warmup_term_ratio)
    if wu < 1:
        wu = 0
    if i < wu:
        return lr * i / wu
    end = tr.learning_rate_end_ratio
    a = (1.0 - end) * lr / (1.0 + math.cos(math.pi * wu / E))
    return math.cos(math.pi * (i - wu) / (E - wu)) * a + a + end * lr


def cos_anneal_ratio(conf, i):
    """solver.py:100-108 - as written there: 0.5 cos(pi x) + 0.5 while x = i / (epoch * cos_anneal_term_ratio) < 1
    (so it starts at 1 and falls towards 0), then jumps to 1."""
    x = i / (conf.train.epoch * conf.train.cos_anneal_term_ratio)
    return 0.5 * math.cos(math.pi * x) + 0.5 if x < 1.0 else 1.0


def light_visibility_gain(conf, i):
    """solver.py:110-119: cosine from 1 (epoch 0) to sigmoid_gain_lv_end (last epoch)."""
    M = conf.train.sigmoid_gain_lv_end
    b = (M + 1) * 0.5
    return (1 - b) * math.cos(math.pi * i / conf.train.epoch) + b


def train_iteration(conf, weight_solver, feat_solver, weights, feats, backward):
    """train.py:135-148 after loss.forward(): zero_grad, weight_decay (on the zeroed buffers), [clip off],
    backward (accumulates), skip when BOTH solvers report a non-finite gradient (`and`, solver.py:67-69), update.
    `backward()` must ADD dL/dw into the g arrays.  Returns False when the update was skipped."""
    for params in (weights, feats):
        for _, g in params.values():
            g[...] = 0
    weight_solver.weight_decay(weights, conf.train.weight_decay)
    feat_solver.weight_decay(feats, conf.train.weight_decay)
    assert conf.train.clip_grad_norm <= 0, "clip_grad_by_norm is off in every shipped config (solver.py:53-55)"
    backward()
    if weight_solver.check_inf_or_nan_grad(weights) and feat_solver.check_inf_or_nan_grad(feats):
        return False
    weight_solver.update(weights)
    feat_solver.update(feats)
    return True
