"""Runs the whole vren op chain (a1..a11 of SURVEY.md section 8a) on one seeded scene through a given backend and
returns a dict of numpy outputs with the key names used by tests/golden/vren_ref_*.npz."""
import numpy as np

import scenes

T_THR = 1e-4
TEST_ROUNDS = (1, 4, 16)

# keys that must match bit for bit (integer outputs + everything the marcher produces + exact-order test compositor)
EXACT_KEYS = ["aabb_cnt", "aabb_hits_t", "aabb_idx", "rays_a", "xyzs", "deltas", "ts", "counter"] + \
             [f"mt{i}_{k}" for i in range(3) for k in ("xyzs", "deltas", "ts", "neff", "hits_t")]
# fp32 outputs compared with tolerance: (rtol, atol).  ws / total_samples / alive sets are exact against the GPU
# reference (same __expf), and within 1 sample / 1e-6 against the CPU oracle (libm expf).
TOL_KEYS = {
    "cf_opacity": (1e-5, 1e-6), "cf_depth": (1e-5, 1e-6), "cf_rgb": (1e-5, 1e-6), "cf_ws": (1e-5, 1e-7),
    "cb_dsig": (2e-4, 2e-5), "cb_drgbs": (1e-5, 1e-6),
    # wts_incl*ws_excl - ws_incl*wts_excl cancels catastrophically in fp32 (two O(1) products whose difference is
    # O(1e-2), summed over up to ~500 samples): the per-ray loss carries ~N*1e-7 absolute noise from the scan order
    # and from the last bits of ws -- the reference's own value is that far from the fp64 one as well
    "dl_loss": (5e-3, 2e-4), "dl_wi": (1e-5, 1e-6), "dl_wti": (1e-5, 1e-6), "dl_dws": (2e-4, 2e-5),
}
for _i in range(3):
    TOL_KEYS.update({f"mt{_i}_opacity": (1e-5, 1e-6), f"mt{_i}_depth": (1e-5, 1e-6), f"mt{_i}_rgb": (1e-5, 1e-6)})


def grads(n_rays, tot):
    rng = np.random.RandomState(9)
    gO, gD = rng.randn(n_rays).astype(np.float32), rng.randn(n_rays).astype(np.float32)
    gRGB, gW = rng.randn(n_rays, 3).astype(np.float32), rng.randn(tot).astype(np.float32)
    gL = rng.randn(n_rays).astype(np.float32)
    return gO, gD, gRGB, gW, gL


def run_oracle(sc):
    from oracle import vren_oracle as orc
    out = {}
    n_rays = sc["rays_o"].shape[0]
    cnt, ht, hi = orc.ray_aabb_intersect(sc["rays_o"], sc["rays_d"], sc["center"], sc["half"], 1)
    out.update(aabb_cnt=cnt, aabb_hits_t=ht, aabb_idx=hi)
    h = scenes.near_clamp(ht)
    out["hits_t"] = h
    ra, xyzs, dirs, deltas, ts, counter = orc.raymarching_train(sc["rays_o"], sc["rays_d"], h, sc["bitfield"], sc["cascades"], sc["scale"],
                                                                sc["esf"], sc["noise"], sc["grid_size"], sc["max_samples"])
    tot = int(counter[0])
    out.update(rays_a=ra, xyzs=xyzs, dirs=dirs, deltas=deltas, ts=ts, counter=counter)
    sig, rgbs = scenes.field_values(tot, seed=5)
    total, opacity, depth, rgb, ws = orc.composite_train_fw(sig, rgbs, deltas, ts, ra, T_THR)
    out.update(cf_total=total, cf_opacity=opacity, cf_depth=depth, cf_rgb=rgb, cf_ws=ws)
    gO, gD, gRGB, gW, gL = grads(n_rays, tot)
    dsig, drgbs = orc.composite_train_bw(gO, gD, gRGB, gW, sig, rgbs, ws, deltas, ts, ra, opacity, depth, rgb, T_THR)
    out.update(cb_dsig=dsig, cb_drgbs=drgbs)
    loss, wi, wti = orc.distortion_loss_fw(ws, deltas, ts, ra)
    out.update(dl_loss=loss, dl_wi=wi, dl_wti=wti, dl_dws=orc.distortion_loss_bw(gL, wi, wti, ws, deltas, ts, ra))
    h2 = h.copy(); alive = np.arange(n_rays, dtype=np.int64)
    opacity = np.zeros(n_rays, np.float32); depth = np.zeros(n_rays, np.float32); rgb = np.zeros((n_rays, 3), np.float32)
    for it, ns in enumerate(TEST_ROUNDS):
        x, dd, dl, tt, ne = orc.raymarching_test(sc["rays_o"], sc["rays_d"], h2, alive, sc["bitfield"], sc["cascades"], sc["scale"], sc["esf"],
                                                 sc["grid_size"], sc["max_samples"], ns)
        sg, cl = scenes.field_values(alive.shape[0] * ns, seed=20 + it)
        out[f"mt{it}_alive_in"] = alive.copy()
        orc.composite_test_fw(sg.reshape(-1, ns), cl.reshape(-1, ns, 3), dl, tt, h2, alive, T_THR, ne, opacity, depth, rgb)
        out.update({f"mt{it}_xyzs": x, f"mt{it}_deltas": dl, f"mt{it}_ts": tt, f"mt{it}_neff": ne, f"mt{it}_hits_t": h2.copy(),
                    f"mt{it}_alive_out": alive.copy(), f"mt{it}_opacity": opacity.copy(), f"mt{it}_depth": depth.copy(), f"mt{it}_rgb": rgb.copy()})
        alive = alive[alive >= 0]
    return out


def run_cuda(sc, vren, canonicalise=False):
    """`vren` is either our drop-in module or the compiled reference (vren_ref); canonicalise=True re-packs the
    reference's atomically ordered rays_a into ray order."""
    import torch
    dev = torch.device("cuda")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    N = lambda t: t.detach().cpu().numpy()
    out = {}
    n_rays = sc["rays_o"].shape[0]
    o, d, bits = T(sc["rays_o"]), T(sc["rays_d"]), T(sc["bitfield"])
    cnt, ht, hi = vren.ray_aabb_intersect(o, d, T(sc["center"]), T(sc["half"]), 1)
    out.update(aabb_cnt=N(cnt), aabb_hits_t=N(ht), aabb_idx=N(hi))
    h_np = scenes.near_clamp(N(ht)); h = T(h_np)
    out["hits_t"] = h_np
    ra, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(o, d, h, bits, sc["cascades"], sc["scale"], sc["esf"], T(sc["noise"]),
                                                                 sc["grid_size"], sc["max_samples"])
    tot = int(counter[0])
    xyzs, dirs, deltas, ts = xyzs[:tot], dirs[:tot], deltas[:tot], ts[:tot]
    if canonicalise:
        ra_np = N(ra); order = np.argsort(ra_np[:, 0], kind="stable"); ra_np = ra_np[order]
        idx = np.concatenate([np.arange(s, s + n) for _, s, n in ra_np]) if tot > 0 else np.zeros(0, np.int64)
        start = np.concatenate([[0], np.cumsum(ra_np[:, 2])[:-1]])
        ra = T(np.stack([ra_np[:, 0], start, ra_np[:, 2]], 1).astype(np.int64))
        idx_t = T(idx)
        xyzs, dirs, deltas, ts = (a[idx_t].contiguous() for a in (xyzs, dirs, deltas, ts))
    out.update(rays_a=N(ra), xyzs=N(xyzs), dirs=N(dirs), deltas=N(deltas), ts=N(ts), counter=N(counter))
    sig, rgbs = scenes.field_values(tot, seed=5)
    sig, rgbs = T(sig), T(rgbs)
    total, opacity, depth, rgb, ws = vren.composite_train_fw(sig, rgbs, deltas, ts, ra, T_THR)
    out.update(cf_total=N(total), cf_opacity=N(opacity), cf_depth=N(depth), cf_rgb=N(rgb), cf_ws=N(ws))
    gO, gD, gRGB, gW, gL = (T(g) for g in grads(n_rays, tot))
    dsig, drgbs = vren.composite_train_bw(gO, gD, gRGB, gW, sig, rgbs, ws, deltas, ts, ra, opacity, depth, rgb, T_THR)
    out.update(cb_dsig=N(dsig), cb_drgbs=N(drgbs))
    loss, wi, wti = vren.distortion_loss_fw(ws, deltas, ts, ra)
    out.update(dl_loss=N(loss), dl_wi=N(wi), dl_wti=N(wti), dl_dws=N(vren.distortion_loss_bw(gL, wi, wti, ws, deltas, ts, ra)))
    h2 = h.clone(); alive = torch.arange(n_rays, device=dev)
    opacity = torch.zeros(n_rays, device=dev); depth = torch.zeros(n_rays, device=dev); rgb = torch.zeros(n_rays, 3, device=dev)
    for it, ns in enumerate(TEST_ROUNDS):
        x, dd, dl, tt, ne = vren.raymarching_test(o, d, h2, alive, bits, sc["cascades"], sc["scale"], sc["esf"], sc["grid_size"],
                                                  sc["max_samples"], ns)
        sg, cl = scenes.field_values(alive.numel() * ns, seed=20 + it)
        out[f"mt{it}_alive_in"] = N(alive)
        vren.composite_test_fw(T(sg).view(-1, ns), T(cl).view(-1, ns, 3), dl, tt, h2, alive, T_THR, ne, opacity, depth, rgb)
        out.update({f"mt{it}_xyzs": N(x), f"mt{it}_deltas": N(dl), f"mt{it}_ts": N(tt), f"mt{it}_neff": N(ne), f"mt{it}_hits_t": N(h2),
                    f"mt{it}_alive_out": N(alive), f"mt{it}_opacity": N(opacity), f"mt{it}_depth": N(depth), f"mt{it}_rgb": N(rgb)})
        alive = alive[alive >= 0]
    return out


def compare(a, b, exact_expf, label=""):
    """a, b: output dicts.  exact_expf=True when both sides use the GPU's __expf (then ws, total_samples and the
    alive sets must be identical); False for GPU-vs-CPU-oracle where expf differs in the last bits."""
    problems = []
    for k in EXACT_KEYS:
        if k in a and k in b:
            if a[k].shape != b[k].shape or not np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)):
                nbad = -1 if a[k].shape != b[k].shape else int((a[k] != b[k]).sum())
                problems.append(f"{label}{k}: not bit-exact ({nbad} of {a[k].size} differ; shapes {a[k].shape} vs {b[k].shape})")
    for k, (rt, at) in TOL_KEYS.items():
        if k in a and k in b:
            if a[k].shape != b[k].shape:
                problems.append(f"{label}{k}: shape {a[k].shape} vs {b[k].shape}")
            elif not np.allclose(a[k], b[k], rtol=rt, atol=at):
                err = np.abs(a[k] - b[k]); i = int(err.argmax())
                # an early-termination decision that flips on the last bit of T moves one sample's weight
                frac_bad = float((err > at + rt * np.abs(b[k])).mean())
                if exact_expf or frac_bad > 2e-3:
                    problems.append(f"{label}{k}: max abs err {err.max():.3e} at {i} (a={a[k].flat[i]:.6g}, b={b[k].flat[i]:.6g}), {frac_bad:.2%} out of tol")
    for k in ["cf_total"] + [f"mt{i}_alive_out" for i in range(3)]:
        if k in a and k in b:
            diff = float((a[k] != b[k]).mean()) if a[k].shape == b[k].shape else 1.0
            if diff > (0.0 if exact_expf else 2e-3):
                problems.append(f"{label}{k}: {diff:.3%} entries differ")
    return problems
