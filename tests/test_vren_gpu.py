"""GPU parity tests for the vren op set, through the C ABI via the drop-in
`vren` module.  Bars: bit-exact for hit tables, rays_a, sample counts, ts/deltas/xyzs, N_eff, hits_t;
fp32 tolerance (written in tests/vren_cases.py: TOL_KEYS) for ws / opacity / depth / rgb / gradients."""
import glob
import os

import numpy as np
import pytest

import scenes
import vren_cases as vc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

SCENES = [("lego", 2048), ("full", 256), ("unbounded", 2048), ("axis", 1024)]


@pytest.mark.parametrize("name,n_rays", SCENES)
def test_cuda_vs_cpu_oracle(name, n_rays):
    import vren
    sc = scenes.scene(name, n_rays, seed=1)
    ours = vc.run_cuda(sc, vren)
    orc = vc.run_oracle(sc)
    problems = vc.compare(ours, orc, exact_expf=False)
    assert not problems, "\n".join(problems)
    assert ours["counter"][0] == ours["rays_a"][:, 2].sum()


@pytest.mark.parametrize("name,n_rays", SCENES)
def test_cuda_vs_reference_kernels(name, n_rays, ref_vren):
    """ours vs the reference's own CUDA kernels, live on this GPU (oracle/_ref/vren_ref*.so)"""
    if ref_vren is None:
        pytest.skip("oracle/_ref/vren_ref*.so not built")
    import vren
    sc = scenes.scene(name, n_rays, seed=2)
    ours = vc.run_cuda(sc, vren)
    ref = vc.run_cuda(sc, ref_vren, canonicalise=True)
    problems = vc.compare(ours, ref, exact_expf=True)
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "vren_ref_*.npz"))))
def test_cuda_vs_golden(path):
    """ours vs committed outputs of the reference's kernels (tests/golden/make_golden_vren.py)"""
    import vren
    g = dict(np.load(path))
    if "rays_o" not in g:
        pytest.skip("utility fixture")
    sc = {k: (g[k].item() if g[k].ndim == 0 else g[k]) for k in ("rays_o", "rays_d", "bitfield", "cascades", "scale", "esf", "grid_size",
                                                               "max_samples", "noise", "center", "half")}
    sc["cascades"], sc["grid_size"], sc["max_samples"] = int(sc["cascades"]), int(sc["grid_size"]), int(sc["max_samples"])
    sc["scale"], sc["esf"] = float(sc["scale"]), float(sc["esf"])
    ours = vc.run_cuda(sc, vren)
    problems = vc.compare(ours, g, exact_expf=True)
    assert not problems, "\n".join(problems)


def test_utils_exhaustive():
    """morton round trip over all 128^3 cells, packbits vs numpy on every dtype"""
    import torch
    import vren
    from mfnerf_b200 import synthetic as syn
    coords = torch.from_numpy(syn.morton_order_coords(128)).cuda()
    idx = vren.morton3D(coords)
    assert torch.equal(idx, torch.arange(128 ** 3, dtype=torch.int32, device="cuda"))
    assert torch.equal(vren.morton3D_invert(idx), coords)
    rng = np.random.RandomState(0)
    grid = rng.randn(3, 128 ** 3).astype(np.float32)
    grid[0, :16] = [0.5, np.nan, np.inf, -np.inf, 0.50000006, 0.49999997, -0.0, 0.0, 1e-45, -1, 7, 0.5, 0.5, 1, 2, 3]
    want = np.packbits(grid.reshape(-1) > 0.5, bitorder="little")
    for dt in (torch.float32, torch.float16, torch.float64):
        g = torch.from_numpy(grid).cuda().to(dt)
        bf = torch.zeros(3 * 128 ** 3 // 8, dtype=torch.uint8, device="cuda")
        assert vren.packbits(g, 0.5, bf) is None
        w = want if dt != torch.float16 else np.packbits(g.float().cpu().numpy().reshape(-1) > 0.5, bitorder="little")
        assert np.array_equal(bf.cpu().numpy(), w), dt


def test_empty_and_ragged_inputs():
    import torch
    import vren
    dev = "cuda"
    e3 = torch.zeros(0, 3, device=dev)
    cnt, ht, hi = vren.ray_aabb_intersect(e3, e3, torch.zeros(1, 3, device=dev), torch.ones(1, 3, device=dev), 1)
    assert cnt.shape == (0,) and ht.shape == (0, 1, 2) and hi.shape == (0, 1)
    bits = torch.full((128 ** 3 // 8,), 255, dtype=torch.uint8, device=dev)
    ra, x, d, dl, ts, c = vren.raymarching_train(e3, e3, torch.zeros(0, 2, device=dev), bits, 1, 0.5, 0.0, torch.zeros(0, device=dev), 128, 1024)
    assert ra.shape == (0, 3) and x.shape == (0, 3) and c.tolist() == [0, 0]
    # rays that all miss: every ray still gets a rays_a row with N=0
    o = torch.tensor([[5.0, 5, 5]] * 7, device=dev); dd = torch.tensor([[1.0, 0, 0]] * 7, device=dev)
    _, ht, _ = vren.ray_aabb_intersect(o, dd, torch.zeros(1, 3, device=dev), torch.full((1, 3), 0.5, device=dev), 1)
    assert (ht == -1).all()
    ra, x, d, dl, ts, c = vren.raymarching_train(o, dd, ht[:, 0].contiguous(), bits, 1, 0.5, 0.0, torch.rand(7, device=dev), 128, 1024)
    assert ra[:, 0].tolist() == list(range(7)) and (ra[:, 1:] == 0).all() and x.shape[0] == 0 and c.tolist() == [0, 7]
    tot, op, dp, rgb, ws = vren.composite_train_fw(torch.zeros(0, device=dev), e3, torch.zeros(0, device=dev), torch.zeros(0, device=dev), ra, 1e-4)
    assert (op == 0).all() and (tot == 0).all() and ws.shape == (0,)
    # max_samples cap (dt_min = sqrt(3)/max_samples also changes): must agree with the oracle, and a budget of 4
    # through a full grid stops at exactly 4 samples
    from oracle import vren_oracle as orc
    o = torch.tensor([[-1.0, 0.01, 0.02], [-1.0, 0.3, -0.2]], device=dev); dd = torch.tensor([[1.0, 0.0, 0.0], [1.0, 0.05, 0.1]], device=dev)
    _, ht, _ = vren.ray_aabb_intersect(o, dd, torch.zeros(1, 3, device=dev), torch.full((1, 3), 0.5, device=dev), 1)
    for ms in (4, 16, 100, 1024):
        ra, x, d, dl, ts, c = vren.raymarching_train(o, dd, ht[:, 0].contiguous(), bits, 1, 0.5, 0.0, torch.zeros(2, device=dev), 128, ms)
        ra_o, x_o, d_o, dl_o, ts_o, c_o = orc.raymarching_train(o.cpu().numpy(), dd.cpu().numpy(), ht[:, 0].cpu().numpy(), bits.cpu().numpy(), 1, 0.5,
                                                                0.0, np.zeros(2, np.float32), 128, ms)
        assert np.array_equal(ra.cpu().numpy(), ra_o) and np.array_equal(ts.cpu().numpy(), ts_o), ms
        assert ra[:, 2].max().item() <= ms
    # budget smaller than the lattice: exponential stepping off, tiny max_samples on the *test* marcher
    x, d, dl, ts, ne = vren.raymarching_test(o, dd, ht[:, 0].contiguous().clone(), torch.arange(2, device=dev), bits, 1, 0.5, 0.0, 128, 1024, 3)
    assert ne.tolist() == [3, 3]


def test_error_convention():
    """ref: include/utils.h:4-6 -- non-CUDA / non-contiguous inputs raise RuntimeError"""
    import torch
    import vren
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        vren.morton3D(torch.zeros(4, 3, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="must be contiguous"):
        vren.morton3D(torch.zeros(3, 4, dtype=torch.int32, device="cuda").t())
    with pytest.raises(RuntimeError):
        vren.morton3D(torch.zeros(4, 3, dtype=torch.int64, device="cuda"))
