"""Generates tests/golden/dataset_ref.npz by running the REFERENCE's own python functions on CPU in the build container
(where /root/reference exists):
  * datasets/ray_utils.py  get_ray_directions + get_rays                       (batched poses, as train.py:83-96 calls them)
  * models/networks.py     NGP.mark_invisible_cells                            (called unbound on a stand-in object)
`kornia` (requirements.txt) is not installed here; its create_meshgrid(H, W, normalized_coordinates=False) is a two-line
linspace/meshgrid, stubbed below.  tinycudann / vren resolve to this repo's drop-ins (import only; nothing of them is executed).

    python tests/golden/make_golden_dataset.py            # writes tests/golden/dataset_ref.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "mf-nerf_b200")); sys.path.insert(0, ROOT)


def _kornia_stub():
    def create_meshgrid(height, width, normalized_coordinates=True, device="cpu", dtype=torch.float32):
        assert not normalized_coordinates
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        return torch.stack(torch.meshgrid([xs, ys], indexing="ij"), dim=-1).permute(1, 0, 2).unsqueeze(0)     # 1 x H x W x 2, [..., 0] = x
    k = types.ModuleType("kornia"); k.create_meshgrid = create_meshgrid
    sys.modules["kornia"] = k


def main():
    from mfnerf_b200 import synthetic as syn
    from oracle import dataset_ref as dr
    _kornia_stub()
    sys.path.insert(0, REF)
    import importlib.util                              # the reference's file, unmodified, loaded by path: datasets/__init__.py would pull
    spec = importlib.util.spec_from_file_location("ref_ray_utils", os.path.join(REF, "datasets", "ray_utils.py"))   # in imageio / cv2 readers
    ray_utils = importlib.util.module_from_spec(spec); spec.loader.exec_module(ray_utils)
    from models import networks                        # idem (imports our tinycudann / vren drop-ins)
    sys.path.remove(REF)
    rng = np.random.default_rng(7)
    W, H = 64, 48
    K = np.array([[70.5, 0, 31.25], [0, 69.75, 24.5], [0, 0, 1]], np.float32)
    poses = syn.camera_poses(12, seed=3).astype(np.float32)                       # (12, 3, 4) upper-hemisphere cameras, radius 1.5
    n = 4096
    img = rng.integers(0, len(poses), n); pix = rng.integers(0, W * H, n)
    directions = ray_utils.get_ray_directions(H, W, torch.from_numpy(K))
    rays_o, rays_d = ray_utils.get_rays(directions[torch.from_numpy(pix)], torch.from_numpy(poses)[torch.from_numpy(img)])
    view_o, view_d = ray_utils.get_rays(directions, torch.from_numpy(poses[5]))   # test split: one pose, all pixels
    out = dict(K=K, img_wh=np.array([W, H]), poses=poses, img_idxs=img, pix_idxs=pix, directions=directions.numpy(),
               rays_o=rays_o.numpy().copy(), rays_d=rays_d.numpy(), view_o=view_o.numpy().copy(), view_d=view_d.numpy())
    # mark_invisible_cells on a stand-in for the NGP module: G = 32, 2 cascades (scale 1), cameras pulled in to radius 0.9 so that some
    # cells are too near and some are out of view
    G, cascades, scale = 32, 2, 1.0
    m = np.arange(G ** 3, dtype=np.uint32)
    coords = torch.from_numpy(np.stack([dr._compact(m), dr._compact(m >> 1), dr._compact(m >> 2)], 1).astype(np.int32))
    idx = torch.arange(G ** 3)
    near_poses = poses.copy(); near_poses[:, :, 3] *= 0.6
    me = types.SimpleNamespace(density_grid=torch.zeros(cascades, G ** 3), cascades=cascades, scale=scale, grid_size=G,
                               get_all_cells=lambda: [(idx, coords)] * cascades)
    networks.NGP.mark_invisible_cells(me, torch.from_numpy(K), torch.from_numpy(near_poses), (W, H))
    out.update(mi_poses=near_poses, mi_G=np.array(G), mi_cascades=np.array(cascades), mi_scale=np.array(scale, np.float32),
               mi_near=np.array(networks.NEAR_DISTANCE, np.float32), mi_density=me.density_grid.numpy(), mi_count=me.count_grid.numpy())
    path = os.path.join(ROOT, "tests", "golden", "dataset_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
    print("valid cells per cascade:", (me.density_grid == 0).sum(1).tolist(), "of", G ** 3)


if __name__ == "__main__":
    main()
