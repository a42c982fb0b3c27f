"""Host-side mirror of the reference's `Solvers` (python/solver.py:20-119) over the fused optimizer kernels
(csrc/optimizer.cu): two Adam groups - every MLP parameter in the engine's flat buffer ("weight") and the feature
grids (names ending in feature/F in the reference, "feat") - with the reference's learning-rate, cos-anneal and
light-visibility-gain schedules.

`step()` is the fused form of one training iteration's optimizer work (train.py:135-148: zero_grad, weight_decay,
backward, check_inf_or_nan_grad, update): ONE pass per buffer that adds the weight decay, applies Adam and leaves the
gradient zeroed for the next iteration, never synchronising with the host.  The reference's individual methods are
kept (same names, same order of use) for callers that drive the loop the way train.py does.
"""
import math

import torch

from . import _lib

BETA1, BETA2, EPS = 0.9, 0.999, 1e-8     # nnabla S.Adam defaults (solver.py:29-30 passes only alpha)


class Solvers:
    def __init__(self, conf, engine):
        tr = conf.train
        scale = (tr.batch_size * tr.n_rays) / (1 * 512)                      # solver.py:23-27
        self.learning_rate_weight = tr.base_learning_rate_weight * scale
        self.learning_rate_feat = tr.base_learning_rate_feat * scale
        self.lr_weight = 0.0                                                 # S.Adam(0)
        self.lr_feat = 0.0
        self.conf, self.eng = conf, engine
        self.cos_anneal_ratio = 0.0
        self._state = None

    # -- parameter groups ---------------------------------------------------------------------------
    def set_parameters(self):
        ps, dev = self.eng.params, self.eng.device
        self._groups = [("weight", ps.data, ps.grad)] + [("feat", ps.grid[k], ps.grid_grad[k]) for k in sorted(ps.grid)]
        self._state = [(torch.zeros_like(w), torch.zeros_like(w)) for _, w, _ in self._groups]
        self._t = torch.zeros(1, dtype=torch.int32, device=dev)              # Adam's update count, on the device
        self._flags = torch.zeros(2, dtype=torch.int32, device=dev)          # non-finite gradient: [weight, feat]
        self._decay = 0.0

    def _call(self, name, *args):
        _lib.call(name, *args, self.eng.stream())

    def _lr(self, kind):
        return self.lr_weight if kind == "weight" else self.lr_feat

    # -- the reference's methods (solver.py:48-69) ---------------------------------------------------
    def zero_grad(self):
        self.eng.params.zero_grad()

    def weight_decay(self):
        rate = self.conf.train.weight_decay
        for _, w, g in self._groups:
            self._call("ndjir_weight_decay", w.numel(), g, w, rate)

    def clip_grad_by_norm(self):
        if self.conf.train.clip_grad_norm <= 0:
            return
        raise NotImplementedError("clip_grad_by_norm > 0 is not used by any shipped config (solver.py:53-58)")

    def _scan(self):
        """flags[0] = weight group non-finite; the feat groups are scanned only if it is (the reference's `and`)."""
        self._flags.zero_()
        f0, f1 = self._flags.data_ptr(), self._flags.data_ptr() + 4
        for kind, _, g in self._groups:
            if kind == "weight":
                self._call("ndjir_nonfinite_flag", g.numel(), g, f0, None)
        for kind, _, g in self._groups:
            if kind == "feat":
                self._call("ndjir_nonfinite_flag", g.numel(), g, f1, f0)

    def check_inf_or_nan_grad(self):
        """True when BOTH solvers see a non-finite gradient (solver.py:67-69).  Synchronises (reads two ints)."""
        self._scan()
        fl = self._flags.cpu()
        # without a feature grid the feat solver is empty, reports False, and the `and` never trips
        return bool(fl[0]) and bool(fl[1])

    def update(self, fused_decay=0.0, zero_grad=False, skip_flags=None):
        self._call("ndjir_adam_tick", self._t, skip_flags)
        for (kind, w, g), (m, v) in zip(self._groups, self._state):
            self._call("ndjir_adam_step", w.numel(), w, g, m, v, self._lr(kind), BETA1, BETA2, EPS, fused_decay,
                       self._t, skip_flags, 1 if zero_grad else 0)
        # the W^T copies of the input-gradient products are refreshed at the start of Engine.train_step

    # -- fused iteration ----------------------------------------------------------------------------------
    def step(self, loss=None):
        """After train_step(zero_grad=False) accumulated dL/dw into zeroed buffers: decay + check + Adam + zero, fused.
        The iteration is skipped on the device - gradients zeroed, Adam's count not advanced - when both solvers see a
        non-finite gradient (solver.py:67-69) OR the loss is NaN (the loop's second guard, train.py:144-146; pass the
        device tensor train_step returned).  With `voxel.type: none` the feat group is empty and only the loss guard
        can trip, exactly as in the reference."""
        self._scan()
        if loss is not None:
            self._call("ndjir_nan_loss_flag", 1, loss, self._flags)
        self.update(fused_decay=self.conf.train.weight_decay, zero_grad=True, skip_flags=self._flags)

    # -- schedules (solver.py:71-119) -------------------------------------------------------------------
    def update_learning_rate(self, i):
        self.lr_weight = self.compute_learning_rate(i, self.learning_rate_weight)
        self.lr_feat = self.compute_learning_rate(i, self.learning_rate_feat)
        self.update_cos_anneal_ratio(i)
        self.update_light_visibility_gain(i)

    def compute_learning_rate(self, i, lr):
        tr = self.conf.train
        epoch = tr.epoch
        warmup_term = int(epoch * tr.warmup_term_ratio)
        warmup_term = 0 if warmup_term < 1 else warmup_term
        if i < warmup_term:
            return lr * i / warmup_term
        r_end = tr.learning_rate_end_ratio
        amp = (1 - r_end) * lr / (1 + math.cos(math.pi * warmup_term / epoch))
        return math.cos(math.pi * (i - warmup_term) / (epoch - warmup_term)) * amp + (amp + r_end * lr)

    def update_cos_anneal_ratio(self, i):
        tr = self.conf.train
        x = i / (tr.epoch * tr.cos_anneal_term_ratio)
        self.cos_anneal_ratio = 0.5 * math.cos(math.pi * x) + 0.5 if x < 1.0 else 1.0
        return self.cos_anneal_ratio

    def update_light_visibility_gain(self, i):
        tr = self.conf.train
        hi = tr.sigmoid_gain_lv_end
        mid = (hi + 1) * 0.5
        self.eng.params.pl_gain = float((1 - mid) * math.cos(math.pi * i / tr.epoch) + mid)
        return self.eng.params.pl_gain
