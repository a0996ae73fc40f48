"""SURVEY 8f-2 / 8f-3 on the device: ray generation + batch sampling (mfn_ray_batch) and mark_invisible_cells
(mfn_grid_mark_invisible) against the reference's own python functions (golden fixture) and the numpy restatement."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import scenes
from oracle import dataset_ref as dr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "dataset_ref.npz")
T = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).cuda() if dt is None else torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()


def test_ray_batch_matches_reference_golden():
    from mfnerf_b200 import dataset as mds
    g = np.load(GOLD)
    W, H = (int(v) for v in g["img_wh"])
    cam = mds.Camera.from_K(g["K"], (W, H))
    poses = T(g["poses"])
    rng = np.random.default_rng(1)
    pixels = rng.random((len(g["poses"]), W * H, 4), dtype=np.float32)            # 4 channels: HDR-NeRF style rays (base.py:33-34); rgb = [:, :3]
    img, pix = T(g["img_idxs"], torch.int64), T(g["pix_idxs"], torch.int64)
    o, d, rgb = mds.ray_batch(cam, poses, len(g["img_idxs"]), pixels=T(pixels), img_idxs=img, pix_idxs=pix)
    assert torch.equal(o.cpu(), torch.from_numpy(g["rays_o"]))
    # reference (torch bmm on CPU) vs our FMA chain: last-bit differences of a 3-term fp32 dot product (2 ulp of 1)
    torch.testing.assert_close(d.cpu(), torch.from_numpy(g["rays_d"]), rtol=0, atol=2.4e-7)
    ro, rd, rrgb = dr.rays_from_indices(g["K"], (W, H), g["poses"], g["img_idxs"], g["pix_idxs"], pixels)
    assert np.array_equal(d.cpu().numpy(), rd) and np.array_equal(o.cpu().numpy(), ro)          # the oracle's FMA emulation: bit-exact
    assert np.array_equal(rgb.cpu().numpy(), rrgb)
    # precomputed direction table (self.directions, train.py:86) instead of the intrinsics: same bits
    o2, d2, _ = mds.ray_batch(cam, poses, len(g["img_idxs"]), directions=T(g["directions"]), img_idxs=img, pix_idxs=pix)
    assert torch.equal(d2, d) and torch.equal(o2, o)
    # test split: one pose, every pixel in order
    vo, vd, vrgb = mds.ray_batch(cam, poses, W * H, pixels=T(pixels), image=5)
    torch.testing.assert_close(vd.cpu(), torch.from_numpy(g["view_d"]), rtol=0, atol=2.4e-7)
    assert torch.equal(vo.cpu(), torch.from_numpy(g["view_o"])) and np.array_equal(vrgb.cpu().numpy(), pixels[5, :, :3])
    # empty batch, bad arguments
    e = mds.ray_batch(cam, poses, 0)
    assert e[0].shape == (0, 3) and e[2] is None
    from mfnerf_b200._lib import MfnError
    with pytest.raises(MfnError):
        mds.ray_batch(cam, poses, 10, image=len(g["poses"]))
    with pytest.raises(RuntimeError):
        mds.ray_batch(cam, poses.cpu(), 10)
    bad = img.clone(); bad[3] = 99                                                 # out-of-range index: a dead ray, no wild read
    o3, d3, _ = mds.ray_batch(cam, poses, len(bad), img_idxs=bad, pix_idxs=pix)
    assert (d3[3] == 0).all() and torch.equal(d3[4:], d[4:])


def test_ray_batch_draws_are_uniform_and_counter_based():
    from mfnerf_b200 import dataset as mds
    W, H, N = 40, 30, 7
    K = np.array([[50.0, 0, 20.0], [0, 50.0, 15.0], [0, 0, 1]], np.float32)
    cam = mds.Camera.from_K(K, (W, H))
    poses = T(scenes.syn.camera_poses(N, seed=1))
    pixels = torch.rand(N, W * H, 3, device="cuda")
    ctr = torch.zeros(1, dtype=torch.int64, device="cuda")
    n = 1 << 18
    o, d, rgb, img, pix = mds.ray_batch(cam, poses, n, pixels=pixels, strategy="all_images", seed=11, call_counter=ctr, return_indices=True)
    assert int(img.min()) == 0 and int(img.max()) == N - 1 and int(pix.min()) == 0 and int(pix.max()) == W * H - 1
    # the indices reported are the indices used
    o2, d2, rgb2 = mds.ray_batch(cam, poses, n, pixels=pixels, img_idxs=img, pix_idxs=pix)
    assert torch.equal(o, o2) and torch.equal(d, d2) and torch.equal(rgb, rgb2)
    # uniform with replacement (np.random.choice, base.py:24-29): chi-square of the image and pixel histograms, 6-sigma bounds
    for idx, k in ((img, N), (pix, W * H)):
        h = torch.bincount(idx, minlength=k).double()
        chi2 = float(((h - n / k) ** 2 / (n / k)).sum())
        assert abs(chi2 - (k - 1)) < 6 * (2 * (k - 1)) ** 0.5, (k, chi2)
    joint = img * (W * H) + pix                                                   # image and pixel are independent
    assert abs(float(torch.corrcoef(torch.stack([img.double(), pix.double()]))[0, 1])) < 0.01 and joint.unique().numel() > 0.5 * N * W * H
    # a full-size image (640 000 pixels: the draw must not wrap in 64-bit arithmetic) -- coarse histogram over the whole index range
    big = mds.Camera.from_K(K, (800, 800))
    _, _, _, bi, bp = mds.ray_batch(big, poses, n, strategy="all_images", seed=5, call_counter=ctr, return_indices=True)
    assert int(bp.max()) > 0.999 * 640000 and int(bp.min()) < 0.001 * 640000
    h = torch.bincount(bp * 64 // 640000, minlength=64).double()
    chi2 = float(((h - n / 64) ** 2 / (n / 64)).sum())
    assert abs(chi2 - 63) < 6 * (2 * 63) ** 0.5, chi2
    # same (seed, counter) -> same batch; another counter value or seed -> another batch
    same = mds.ray_batch(cam, poses, n, strategy="all_images", seed=11, call_counter=ctr, return_indices=True)
    assert torch.equal(same[3], img) and torch.equal(same[4], pix)
    ctr += 1
    nxt = mds.ray_batch(cam, poses, n, strategy="all_images", seed=11, call_counter=ctr, return_indices=True)
    assert float((nxt[4] == pix).double().mean()) < 0.01
    other = mds.ray_batch(cam, poses, n, strategy="all_images", seed=12, call_counter=ctr, return_indices=True)
    assert float((other[4] == nxt[4]).double().mean()) < 0.01
    # 'same_image': one image per call, changing from call to call
    seen = set()
    for c in range(40):
        ctr.fill_(c)
        s = mds.ray_batch(cam, poses, 4096, strategy="same_image", seed=3, call_counter=ctr, return_indices=True)
        assert int(s[3].min()) == int(s[3].max())
        seen.add(int(s[3][0]))
    assert len(seen) == N


def _check_mark(dens, cnt, want_dens, want_cnt, fragile, cnt_atol=0.0):
    dens, cnt = dens.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(dens[~fragile], want_dens[~fragile])                     # every robustly decided cell: exact
    assert np.abs(cnt[~fragile] - want_cnt[~fragile]).max() <= cnt_atol
    assert (dens != want_dens).mean() < 2e-4                                       # cells on a frustum edge may go either way in fp32


def test_mark_invisible_cells_matches_reference():
    from mfnerf_b200._lib import call, ptr, stream_ptr
    g = np.load(GOLD)
    W, H = (int(v) for v in g["img_wh"])
    G, C, scale, near = int(g["mi_G"]), int(g["mi_cascades"]), float(g["mi_scale"]), float(g["mi_near"])
    _, _, fragile = dr.mark_invisible_cells(g["K"], g["mi_poses"], (W, H), C, scale, G, near)

    def run(K, poses, G, C, scale, near, wh):
        dens = torch.full((C, G ** 3), 7.0, device="cuda"); cnt = torch.full((C, G ** 3), 7.0, device="cuda")
        Kd, Pd = T(K), T(poses)          # (named: a temporary's block would be handed to the next allocation)
        call("mfn_grid_mark_invisible", ptr(Kd), ptr(Pd), len(poses), wh[0], wh[1], C, scale, G, near, ptr(dens), ptr(cnt), stream_ptr())
        return dens, cnt
    dens, cnt = run(g["K"], g["mi_poses"], G, C, scale, near, (W, H))
    # the reference's own output, produced on CPU (golden): the coverage fraction is k / N there and k * (1 / N) on a GPU -- one ulp
    _check_mark(dens, cnt, g["mi_density"], g["mi_count"], fragile, cnt_atol=6e-8)
    wd, wc, _ = dr.mark_invisible_cells(g["K"], g["mi_poses"], (W, H), C, scale, G, near, cuda_scalar_division=True)
    _check_mark(dens, cnt, wd, wc, fragile)
    # full-size grid, 100 cameras, a large near distance so that the too-near rule removes cells: against the restatement
    poses = scenes.syn.camera_poses(100, seed=2).astype(np.float32); poses[:, :, 3] *= 0.5
    K = np.array([[1111.111, 0, 400], [0, 1111.111, 400], [0, 0, 1]], np.float32)
    wd, wc, wf = dr.mark_invisible_cells(K, poses, (800, 800), 2, 1.0, 128, 0.3, cuda_scalar_division=True)
    assert 0.02 < (wd == -1).mean() < 0.98 and wf.mean() < 0.05
    dens, cnt = run(K, poses, 128, 2, 1.0, 0.3, (800, 800))
    _check_mark(dens, cnt, wd, wc, wf)


def test_engine_mark_invisible_cells_vs_reference_module():
    """NGPEngine.mark_invisible_cells against the reference's unmodified NGP.mark_invisible_cells run on the GPU (staged python layer)"""
    from test_engine_gpu import _load_reference_python, _hparams
    from mfnerf_b200.engine import NGPEngine
    _, networks, _ = _load_reference_python()
    eng = NGPEngine(scale=2.0, n_rays=256, sample_capacity=256 * 64, log2_T=15)
    model = networks.NGP(scale=2.0, hparams=_hparams(), rgb_act="Sigmoid").cuda()
    G = model.grid_size
    model.register_buffer("density_grid", torch.zeros(model.cascades, G ** 3, device="cuda"))      # train.py:78-81 registers both buffers;
    model.register_buffer("grid_coords", eng.cell_coords.clone())                  # create_meshgrid3d(G, G, G): any order of all cells will do
    poses = scenes.syn.camera_poses(24, seed=4).astype(np.float32)
    K = np.array([[1111.111, 0, 400], [0, 1111.111, 400], [0, 0, 1]], np.float32)
    model.mark_invisible_cells(T(K), T(poses), (800, 800))
    eng.mark_invisible_cells(K, poses, (800, 800))
    assert eng.cascades == model.cascades == 3
    _, _, fragile = dr.mark_invisible_cells(K, poses, (800, 800), eng.cascades, 2.0, G, 0.01)
    _check_mark(eng.density_grid, eng.count_grid, model.density_grid.cpu().numpy(), model.count_grid.cpu().numpy(), fragile)
    # cells marked -1 stay out of the occupancy update (networks.py:263-264)
    invisible = eng.density_grid < 0
    assert 0 < int(invisible.sum()) < invisible.numel()
    eng.update_density_grid(warmup=True)
    assert (eng.density_grid[invisible] == -1).all()


def test_train_step_resident_draws_from_the_dataset():
    """a data set resident in HBM: every step draws its batch on the device (no host-to-device copy), the batch is consistent with
    the data set, and training on it converges like training on explicitly supplied rays"""
    from mfnerf_b200 import dataset as mds
    from mfnerf_b200.engine import NGPEngine
    W = H = 100
    f = scenes.syn.FOCAL * W / 800
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], np.float32)
    poses = scenes.syn.camera_poses(8, seed=6).astype(np.float32)
    dirs = dr.ray_directions(H, W, K)
    pixels = np.stack([scenes.syn.analytic_render(torch.from_numpy(np.broadcast_to(p[:, 3], dirs.shape).copy()), torch.from_numpy(dirs @ p[:, :3].T)).numpy()
                       for p in poses]).astype(np.float32)
    ds = mds.ResidentDataset(K, (W, H), poses, pixels)
    R = 1024
    eng = NGPEngine(scale=0.5, n_rays=R, sample_capacity=R * 160, log2_T=15, seed=1)
    eng.density_grid.copy_(torch.from_numpy(scenes.syn.lego_density_grid(0.5, 1)).cuda()); eng.repack_bitfield(0.5)
    eng.attach_dataset(ds, seed=5)
    losses, batches = [], []
    for step in range(1, 151):
        if step == 4:
            eng.capture()
        ctr = eng._march_hdr[2:3].clone()
        eng.train_step_resident(global_step=step if step % 16 else step + 1)      # (step % 16 == 0 would rebuild the occupancy grid from the young network)
        torch.cuda.synchronize()
        losses.append(float(eng.loss_terms[0]) / R)
        if step in (2, 3, 10, 11):
            o, d, rgb, img, pix = mds.ray_batch(ds.camera, ds.poses, R, pixels=ds.pixels, strategy="all_images", seed=5, call_counter=ctr, return_indices=True)
            assert torch.equal(o, eng.rays_o) and torch.equal(d, eng.rays_d) and torch.equal(rgb, eng.target)
            assert torch.equal(rgb, ds.pixels[img, pix])
            batches.append(pix.clone())
    assert not torch.equal(batches[0], batches[1]) and not torch.equal(batches[2], batches[3])       # a new batch every step, graph or not
    assert np.isfinite(losses).all() and np.mean(losses[-10:]) < 0.7 * np.mean(losses[:5]), (losses[:5], losses[-10:])
    full = ds.view_rays(3)
    assert full[0].shape == (W * H, 3) and torch.equal(full[2], ds.pixels[3, :, :3])


def test_occupied_cell_draw_is_uniform_over_all_occupied_cells():
    """mfn_grid_draw_occupied (networks.py:186-193: indices2[randint(len(indices2), (M,))]) on a full-size cascade: every occupied cell is
    equally likely -- the k-th occupied cell for k uniform in [0, total), total far above 2^11 (a 64-bit product in the draw used to wrap)"""
    from mfnerf_b200._lib import call, ptr, stream_ptr
    G3 = 128 ** 3
    g = torch.Generator(device="cuda").manual_seed(0)
    occ = torch.rand(G3, device="cuda", generator=g) < 0.06
    occ[: G3 // 8] = False                                        # an empty stretch at the start, like a carved grid
    cs = torch.cumsum(occ.int(), 0, dtype=torch.int32)
    total = int(cs[-1])
    assert total > 100_000
    n = 1 << 19
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    call("mfn_grid_draw_occupied", ptr(cs), G3, n, 1234, ptr(idx), stream_ptr())
    idx = idx.long()
    assert bool(occ[idx].all())
    rank = cs[idx].long() - 1                                     # which occupied cell, 0 .. total-1
    assert int(rank.max()) > 0.999 * total and int(rank.min()) < 0.001 * total
    h = torch.bincount(rank * 64 // total, minlength=64).double()
    chi2 = float(((h - n / 64) ** 2 / (n / 64)).sum())
    assert abs(chi2 - 63) < 6 * (2 * 63) ** 0.5, chi2
    # no occupied cell: uniform over all cells
    zero = torch.zeros(G3, dtype=torch.int32, device="cuda")
    idx2 = torch.empty(n, dtype=torch.int32, device="cuda")
    call("mfn_grid_draw_occupied", ptr(zero), G3, n, 99, ptr(idx2), stream_ptr())
    idx2 = idx2.long()
    assert int(idx2.max()) > 0.999 * G3 and int(idx2.min()) < 0.001 * G3
    h = torch.bincount(idx2 * 64 // G3, minlength=64).double()
    chi2 = float(((h - n / 64) ** 2 / (n / 64)).sum())
    assert abs(chi2 - 63) < 6 * (2 * 63) ** 0.5, chi2
