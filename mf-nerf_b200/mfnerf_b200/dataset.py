"""Data set resident in device memory + ray generation / batch sampling on the device (libmfnerf_b200.so: mfn_ray_batch).

Mirrors what the reference spreads over datasets/base.py:22-34 (random image / pixel indices, target gather),
train.py:83-96 (pose and direction gathers) and datasets/ray_utils.py:23-35, 60-68 (pixel-centre directions, camera-to-world
rotation): one kernel launch per batch, no DataLoader workers and no host-to-device copy of rays."""
import ctypes

import torch

from ._lib import call, ptr, stream_ptr


class Camera(ctypes.Structure):
    """mirror of `mfn_camera` (include/mfnerf_b200.h)"""
    _fields_ = [("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32)]

    @classmethod
    def from_K(cls, K, img_wh):
        """K: (3,3) intrinsics (datasets/nerf.py:26-31), img_wh = (W, H)"""
        K = torch.as_tensor(K, dtype=torch.float32).cpu()
        return cls(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), int(img_wh[0]), int(img_wh[1]))


STRATEGY = {"given": 0, "all_images": 1, "same_image": 2}


def ray_batch(camera, poses, n_rays, directions=None, pixels=None, img_idxs=None, pix_idxs=None, image=0, strategy="given", seed=0,
              call_counter=None, rays_o=None, rays_d=None, rgb=None, return_indices=False):
    """-> (rays_o, rays_d, rgb or None[, img_idxs, pix_idxs]) for `n_rays` rays; see mfn_ray_batch in the header for the index modes.
    poses (N,3,4) f32, pixels (N, H*W, C>=3) f32, directions (H*W,3) f32, indices int64 -- all CUDA, contiguous."""
    if not poses.is_cuda:
        raise RuntimeError("ray_batch needs CUDA tensors (there is no CPU fallback)")
    for t, dt in ((poses, torch.float32), (directions, torch.float32), (pixels, torch.float32), (img_idxs, torch.int64), (pix_idxs, torch.int64)):
        if t is not None and (t.dtype != dt or not t.is_contiguous() or not t.is_cuda):
            raise RuntimeError("ray_batch: tensors must be contiguous CUDA tensors (float32 data, int64 indices)")
    if poses.dim() != 3 or tuple(poses.shape[1:]) != (3, 4):
        raise RuntimeError("ray_batch: poses must be (N, 3, 4)")
    n_pix = camera.width * camera.height
    if pixels is not None and (pixels.dim() != 3 or pixels.shape[0] != poses.shape[0] or pixels.shape[1] != n_pix or pixels.shape[2] < 3):
        raise RuntimeError("ray_batch: pixels must be (N_images, H*W, C >= 3)")
    if directions is not None and tuple(directions.shape) != (n_pix, 3):
        raise RuntimeError("ray_batch: directions must be (H*W, 3)")
    for t in (img_idxs, pix_idxs):
        if t is not None and t.numel() != n_rays:
            raise RuntimeError("ray_batch: one index per ray")
    d, n = poses.device, int(n_rays)
    rays_o = torch.empty(n, 3, device=d) if rays_o is None else rays_o
    rays_d = torch.empty(n, 3, device=d) if rays_d is None else rays_d
    if rgb is None and pixels is not None:
        rgb = torch.empty(n, 3, device=d)
    io = po = None
    if return_indices:
        io, po = torch.empty(n, dtype=torch.int64, device=d), torch.empty(n, dtype=torch.int64, device=d)
    with torch.cuda.device(d):
        call("mfn_ray_batch", ctypes.byref(camera), ptr(directions), ptr(poses), poses.shape[0], ptr(pixels), pixels.shape[2] if pixels is not None else 0,
             ptr(img_idxs), ptr(pix_idxs), int(image), STRATEGY[strategy], int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(call_counter), n, ptr(rays_o), ptr(rays_d),
             ptr(rgb) if pixels is not None else None, ptr(io), ptr(po), stream_ptr(d))
    return (rays_o, rays_d, rgb, io, po) if return_indices else (rays_o, rays_d, rgb)


class ResidentDataset:
    """All training images and poses in HBM (100 images of 800 x 800 x 3 fp32 = 768 MB of the 180 GB): what the reference keeps in
    host memory as `dataset.rays` / `dataset.poses` and feeds through a 16-worker DataLoader (train.py:146-152)."""

    def __init__(self, K, img_wh, poses, pixels, ray_sampling_strategy="all_images", device="cuda"):
        if ray_sampling_strategy not in ("all_images", "same_image"):
            raise ValueError(f"ray_sampling_strategy {ray_sampling_strategy!r}")          # opt.py: choices of --ray_sampling_strategy
        self.camera = Camera.from_K(K, img_wh)
        self.K = torch.as_tensor(K, dtype=torch.float32).to(device).contiguous()
        self.img_wh = (int(img_wh[0]), int(img_wh[1]))
        self.poses = torch.as_tensor(poses, dtype=torch.float32).to(device).contiguous()
        self.pixels = torch.as_tensor(pixels, dtype=torch.float32).to(device).contiguous()
        self.strategy = ray_sampling_strategy
        n_pix = self.img_wh[0] * self.img_wh[1]
        if self.poses.dim() != 3 or tuple(self.poses.shape[1:]) != (3, 4):
            raise ValueError("poses must be (N, 3, 4)")
        if self.pixels.dim() != 3 or self.pixels.shape[0] != self.poses.shape[0] or self.pixels.shape[1] != n_pix or self.pixels.shape[2] < 3:
            raise ValueError("pixels must be (N_images, H*W, C >= 3)")

    def __len__(self):
        return self.poses.shape[0]

    def view_rays(self, image):
        """all rays of one view in pixel order (the test split of NeRFSystem.forward, train.py:88-96)"""
        n = self.img_wh[0] * self.img_wh[1]
        return ray_batch(self.camera, self.poses, n, pixels=self.pixels, image=int(image))
