"""Device-resident ray source of the training loop: what `IDRDataSource` + `data_iterator_idr` +
`helper.generate_raydir_camloc` do on the host every step (reference python/dataset.py:28-189, python/train.py:126-130,
python/helper.py:44-73) as ONE kernel launch (`ndjir_train_batch`, csrc/inference.cu) over a dataset kept in HBM.

The draw order of the reference is kept: `reset()` takes `rng.permutation(size)` (when shuffling) and then
`rng.randint(0, H*W, (size, R))` from the same `RandomState(313)`; item `position` of an epoch is image
`img_indices[position]` with the pixels `pixel_idx[img_idx]` (dataset.py:33-40, 180-189); a batch is B successive
positions and an exhausted epoch starts the next one with a fresh `reset()`.  `device_rng=True` drops the host draw of
the pixels altogether (counter-based generator on the device: same distribution, not the same stream).

Only the default branch of `_get_data` is mirrored (`patch_ray_sampling: false`, `mask_ray_sample_ratio: 0`, the
values of every BASELINE config); the other two raise NotImplementedError.  Reading images / cameras from disk
(`_load_data`) is the caller's business: the constructor takes arrays."""
import numpy as np
import torch

from . import _lib


class DeviceRaySource:
    def __init__(self, images, masks, intrinsics, poses, n_rays, shuffle=True, rng=None, device="cuda", device_rng=False,
                 seed=313, patch_ray_sampling=False, mask_ray_sample_ratio=0.0):
        """images (n, H, W, 3) in [0, 1]; masks (n, H, W, 1) / (n, H, W) or None; intrinsics (n, 3, 3); poses (n, 4, 4)
        camera-to-world (what dataset.py's `_load_data` returns)."""
        if patch_ray_sampling or mask_ray_sample_ratio > 0:
            raise NotImplementedError("DeviceRaySource mirrors the default branch of IDRDataSource._get_data only "
                                      "(patch_ray_sampling: false, mask_ray_sample_ratio: 0)")
        images = np.asarray(images, np.float32)
        self.size, self.H, self.W = images.shape[0], images.shape[1], images.shape[2]
        self.pixels = self.H * self.W
        self.n_rays, self.shuffle, self.device_rng, self.seed = int(n_rays), shuffle, device_rng, seed
        self.rng = rng if rng is not None else np.random.RandomState(313)       # dataset.py:169-171
        dev = torch.device(device)
        self.device = dev
        self.images = torch.from_numpy(images.reshape(self.size, self.pixels, 3)).to(dev).contiguous()
        self.masks = None
        if masks is not None:
            m = np.asarray(masks, np.float32).reshape(self.size, self.pixels)
            self.masks = torch.from_numpy(m).to(dev).contiguous()
        intr = np.asarray(intrinsics, np.float64)[:, :3, :3]
        poses = np.asarray(poses, np.float64)
        self.kinv = torch.from_numpy(np.linalg.inv(intr).reshape(self.size, 9).copy()).to(dev)
        self.rot = torch.from_numpy(np.ascontiguousarray(poses[:, :3, :3]).reshape(self.size, 9).copy()).to(dev)
        self.camloc_all = torch.from_numpy(poses[:, :3, 3].astype(np.float32).copy()).to(dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)             # step counter of the device generator
        self.position = 0
        self.reset()

    def reset(self):
        """dataset.py:180-189"""
        self.img_indices = self.rng.permutation(self.size) if self.shuffle else np.arange(self.size)
        if not self.device_rng:
            self.pixel_idx = self.rng.randint(0, self.pixels, (self.size, self.n_rays))
            self.pixel_idx_dev = torch.from_numpy(self.pixel_idx.astype(np.int32)).to(self.device)
        self.position = 0

    def next(self, batch_size, out=None, stream=None):
        """-> dict(camloc (B,3), raydir (B,R,3), color_gt (B,R,3), obj_mask (B,R,1), views): the next B dataset items"""
        B, R = int(batch_size), self.n_rays
        views = []
        for _ in range(B):
            if self.position >= self.size:
                self.reset()
            views.append(int(self.img_indices[self.position]))
            self.position += 1
        dev = self.device
        if out is None:
            out = dict(camloc=torch.empty((B, 3), device=dev), raydir=torch.empty((B, R, 3), device=dev),
                       color_gt=torch.empty((B, R, 3), device=dev), obj_mask=torch.empty((B, R, 1), device=dev),
                       pixels=torch.empty((B, R), dtype=torch.int32, device=dev))
        view_ids = torch.tensor(views, dtype=torch.int32, device=dev)
        pix = None
        if not self.device_rng:
            pix = self.pixel_idx_dev[view_ids.long()].contiguous()
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _lib.call("ndjir_train_batch", B, R, self.W, self.pixels, view_ids, pix, self.seed, self.counter, self.images,
                  self.masks, self.kinv, self.rot, self.camloc_all, out["raydir"], out["camloc"], out["color_gt"],
                  out["obj_mask"], out.get("pixels"), st)
        if self.device_rng:
            _lib.call("ndjir_counter_add", self.counter, 1, st)
        out["views"] = views
        return out
