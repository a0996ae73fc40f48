"""Generates tests/golden/vren_ref_<scene>.npz by running the REFERENCE's own kernels (oracle/_ref/vren_ref*.so,
built from /root/reference/models/csrc by oracle/build_ref_vren.sh) on a CUDA device.

Run on the GPU box:   python tests/golden/make_golden_vren.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/.  Inputs are stored with the outputs so the fixtures stay
valid if the synthetic-scene code changes.  The reference's rays_a comes out in atomic arrival order; it is
canonicalised here (rows sorted by ray index, samples re-packed in that order = prefix-sum layout).
"""
import glob
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mf-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402


def load_ref():
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "vren_ref*.so"))[0]
    spec = importlib.util.spec_from_file_location("vren_ref", so)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


def canonicalise(rays_a, arrays):
    """sort rows by ray idx and re-pack per-sample arrays into prefix-sum order"""
    ra = rays_a.cpu().numpy()
    order = np.argsort(ra[:, 0], kind="stable")
    ra = ra[order]
    idx = np.concatenate([np.arange(s, s + n) for _, s, n in ra]) if ra[:, 2].sum() > 0 else np.zeros(0, np.int64)
    new_start = np.concatenate([[0], np.cumsum(ra[:, 2])[:-1]])
    ra_c = np.stack([ra[:, 0], new_start, ra[:, 2]], 1).astype(np.int64)
    return ra_c, [a.cpu().numpy()[idx] for a in arrays]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = load_ref()
    dev = torch.device("cuda")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for name, n_rays in (("lego", 1024), ("full", 192), ("unbounded", 1024), ("axis", 512)):
        sc = scenes.scene(name, n_rays, seed=3)
        o, d, bits = T(sc["rays_o"]), T(sc["rays_d"]), T(sc["bitfield"])
        out = {k: v for k, v in sc.items()}
        # --- a1: ray/AABB -----------------------------------------------------------------------------
        cnt, hits_t, hidx = ref.ray_aabb_intersect(o, d, T(sc["center"]), T(sc["half"]), 1)
        out.update(aabb_cnt=cnt.cpu().numpy(), aabb_hits_t=hits_t.cpu().numpy(), aabb_idx=hidx.cpu().numpy())
        h = T(scenes.near_clamp(hits_t.cpu().numpy()))
        out["hits_t"] = h.cpu().numpy()
        # --- a4: train marcher ------------------------------------------------------------------------
        rays_a, xyzs, dirs, deltas, ts, counter = ref.raymarching_train(o, d, h, bits, sc["cascades"], sc["scale"], sc["esf"],
                                                                        T(sc["noise"]), sc["grid_size"], sc["max_samples"])
        tot = int(counter[0])
        ra_c, (xyzs_c, dirs_c, deltas_c, ts_c) = canonicalise(rays_a, [xyzs[:tot], dirs[:tot], deltas[:tot], ts[:tot]])
        out.update(rays_a=ra_c, xyzs=xyzs_c, deltas=deltas_c, ts=ts_c, counter=counter.cpu().numpy())
        # --- a8/a9: compositing on seeded field values ----------------------------------------------------
        sig, rgbs = scenes.field_values(tot, seed=5)
        ra_t = T(ra_c)
        total, opacity, depth, rgb, ws = ref.composite_train_fw(T(sig), T(rgbs), T(deltas_c), T(ts_c), ra_t, 1e-4)
        out.update(cf_total=total.cpu().numpy(), cf_opacity=opacity.cpu().numpy(), cf_depth=depth.cpu().numpy(), cf_rgb=rgb.cpu().numpy(),
                   cf_ws=ws.cpu().numpy())
        rng = np.random.RandomState(9)
        gO, gD = rng.randn(n_rays).astype(np.float32), rng.randn(n_rays).astype(np.float32)
        gRGB, gW = rng.randn(n_rays, 3).astype(np.float32), rng.randn(tot).astype(np.float32)
        dsig, drgbs = ref.composite_train_bw(T(gO), T(gD), T(gRGB), T(gW), T(sig), T(rgbs), ws, T(deltas_c), T(ts_c), ra_t, opacity, depth, rgb, 1e-4)
        out.update(cb_gO=gO, cb_gD=gD, cb_gRGB=gRGB, cb_gW=gW, cb_dsig=dsig.cpu().numpy(), cb_drgbs=drgbs.cpu().numpy())
        # --- a10: distortion loss ---------------------------------------------------------------------
        loss, wi, wti = ref.distortion_loss_fw(ws, T(deltas_c), T(ts_c), ra_t)
        gL = rng.randn(n_rays).astype(np.float32)
        dws = ref.distortion_loss_bw(T(gL), wi, wti, ws, T(deltas_c), T(ts_c), ra_t)
        out.update(dl_loss=loss.cpu().numpy(), dl_wi=wi.cpu().numpy(), dl_wti=wti.cpu().numpy(), dl_gL=gL, dl_dws=dws.cpu().numpy())
        # --- a11: test-time marcher + compositor, three rounds with growing N_samples -----------------------
        h2 = h.clone()
        alive = torch.arange(n_rays, device=dev)
        opacity = torch.zeros(n_rays, device=dev); depth = torch.zeros(n_rays, device=dev); rgb = torch.zeros(n_rays, 3, device=dev)
        for it, ns in enumerate((1, 4, 16)):
            x, dd, dl, tt, ne = ref.raymarching_test(o, d, h2, alive, bits, sc["cascades"], sc["scale"], sc["esf"], sc["grid_size"],
                                                     sc["max_samples"], ns)
            sg, cl = scenes.field_values(alive.numel() * ns, seed=20 + it)
            sg_t, cl_t = T(sg).view(-1, ns), T(cl).view(-1, ns, 3)
            out[f"mt{it}_alive_in"] = alive.cpu().numpy()
            ref.composite_test_fw(sg_t, cl_t, dl, tt, h2, alive, 1e-4, ne, opacity, depth, rgb)
            out.update({f"mt{it}_xyzs": x.cpu().numpy(), f"mt{it}_deltas": dl.cpu().numpy(), f"mt{it}_ts": tt.cpu().numpy(),
                        f"mt{it}_neff": ne.cpu().numpy(), f"mt{it}_hits_t": h2.cpu().numpy(), f"mt{it}_alive_out": alive.cpu().numpy(),
                        f"mt{it}_opacity": opacity.cpu().numpy(), f"mt{it}_depth": depth.cpu().numpy(), f"mt{it}_rgb": rgb.cpu().numpy()})
            alive = alive[alive >= 0]
        np.savez_compressed(os.path.join(out_dir, f"vren_ref_{name}.npz"), **out)
        print(name, "rays", n_rays, "samples", tot, "->", os.path.getsize(os.path.join(out_dir, f"vren_ref_{name}.npz")), "bytes")
    # --- a2/a3: integer utilities ---------------------------------------------------------------------
    rng = np.random.RandomState(1)
    coords = rng.randint(0, 128, (4096, 3)).astype(np.int32)
    idx = ref.morton3D(T(coords))
    inv = ref.morton3D_invert(idx)
    grid = rng.randn(8 * 4096).astype(np.float32)
    bf = torch.zeros(4096, dtype=torch.uint8, device=dev)
    ref.packbits(T(grid), 0.25, bf)
    np.savez_compressed(os.path.join(out_dir, "vren_ref_utils.npz"), coords=coords, morton=idx.cpu().numpy(), inv=inv.cpu().numpy(), grid=grid,
                        thr=np.float32(0.25), bitfield=bf.cpu().numpy())
    print("utils done")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
