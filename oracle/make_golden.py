"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (the reference cannot travel):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

Every fixture holds seeded synthetic inputs (oracle.synth_inputs) and the
REFERENCE's outputs for each stage of the path -- `Warper.forward` (models/nets/lvd.py:855-870,
run with the stable-sort tie rule of SURVEY.md §8c), `LVD.compute_occ` (:59-68),
`LVD.forward(mode="decode_output")` (:141-153), `WIF.forward` tail (models/nets/wif.py:50-54) --
plus reference autograd gradients of a fixed scalar loss.  The oracle restatement and the CUDA
path are both checked against these files.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import waldo_oracle as wo  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = {
    # name: (PathConfig kwargs, B, T, Tc, smooth)
    "city_x4": (dict(dim=16, load_dim=64, aspect_ratio=2.0, num_obj=3, obj_shape=(2, 2), latent_shape=(2, 4),
                     patch_size=8, num_lyt=4), 1, 3, 2, False),
    "kitti_x2": (dict(dim=16, load_dim=32, aspect_ratio=3.25, num_obj=4, obj_shape=(2, 2), latent_shape=(2, 6),
                      patch_size=8, num_lyt=5), 2, 4, 2, True),
    "train_lo": (dict(dim=16, load_dim=0, aspect_ratio=2.0, num_obj=3, obj_shape=(2, 2), latent_shape=(2, 4),
                      patch_size=8, num_lyt=4, restrict_to_ctx=False, include_self=True), 2, 3, 1, True),
    "cls_plain": (dict(dim=16, load_dim=32, aspect_ratio=2.0, num_obj=2, obj_shape=(2, 2), latent_shape=(2, 4),
                       patch_size=8, num_lyt=3, weight_cls=False, use_disocc=True), 1, 3, 2, False),
    # the real channel counts (Cityscapes C = 3+20, KITTI C = 3+19): the compile-time-specialised channel loops
    "c23_x4": (dict(dim=8, load_dim=32, aspect_ratio=2.0, num_obj=3, obj_shape=(2, 2), latent_shape=(2, 4),
                    patch_size=8, num_lyt=20), 1, 3, 2, True),
    "c22_x2": (dict(dim=8, load_dim=16, aspect_ratio=3.25, num_obj=3, obj_shape=(2, 2), latent_shape=(2, 6),
                    patch_size=8, num_lyt=19), 1, 3, 2, False),
    # 11 heavily overlapping layers: exercises the dense (> 8 live layers per warp) code path of the HD kernels
    "many_obj": (dict(dim=16, load_dim=32, aspect_ratio=2.0, num_obj=10, obj_shape=(2, 2), latent_shape=(2, 4),
                      patch_size=8, num_lyt=3), 1, 3, 2, True),
}


def ref_opt(cfg: wo.PathConfig):
    """A reference option namespace carrying this PathConfig (only the fields Warper/LVD.forward read)."""
    opt = ref_loader.parse_opts("scripts/cityscapes/test.sh")
    for k in ("dim", "load_dim", "aspect_ratio", "num_obj", "patch_size", "scale_factor", "num_lyt", "weight_cls",
              "min_cls", "include_self", "restrict_to_ctx", "use_disocc", "no_filter", "allow_ghost",
              "pad_obj_alpha", "pad_bg_alpha"):
        setattr(opt, k, getattr(cfg, k))
    opt.obj_shape = list(cfg.obj_shape)
    opt.latent_shape = list(cfg.latent_shape)
    opt.num_perm_grid = 1
    opt.time_dropout = False
    return opt


def ref_modules(cfg: wo.PathConfig):
    """(warper, lvd_like) where lvd_like exposes the reference's LVD.forward / compute_occ without
    constructing the (out-of-scope) encoder / transformer stacks."""
    ns = ref_loader.load()
    opt = ref_opt(cfg)
    warper = ns.lvd.Warper(opt)
    om, bg = wo.alpha_masks(cfg)
    fake = types.SimpleNamespace(
        warper=warper, restrict_to_ctx=cfg.restrict_to_ctx, use_disocc=cfg.use_disocc,
        include_self=cfg.include_self, diag=torch.eye(cfg.num_obj)[None, None],
        remove_obj=False, freeze_obj=False)
    fake.compute_occ = lambda s: ns.lvd.LVD.compute_occ(fake, s)
    fake.forward = lambda **kw: ns.lvd.LVD.forward(fake, **kw)
    return warper, fake


def ctx_pred(cfg, B, T, Tc):
    if cfg.restrict_to_ctx:
        Tp = T - Tc
        return torch.arange(Tc).view(1, Tc, 1).expand(B, Tc, Tp), torch.arange(Tc, T)
    # train_lvd: ctx_mode "prev" (synthesizer.py:833-839): previous frame is the context of every frame
    ts = torch.roll(torch.arange(T), 1).view(1, 1, T).expand(B, 1, T)
    return ts, torch.arange(T)


def run_reference(cfg: wo.PathConfig, B, T, Tc, smooth, seed=0):
    ns = ref_loader.load()
    warper, lvd = ref_modules(cfg)
    d = wo.synth_inputs(cfg, B, T, Tc, seed=seed, smooth=smooth, radius=0.2)
    d["ctx_ts"], d["pred_ts"] = ctx_pred(cfg, B, T, Tc)
    leaves = {k: d[k].clone().requires_grad_(True) for k in ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")}
    om, bg = wo.alpha_masks(cfg)
    obj_alpha = om * leaves["obj_alpha_raw"] + (1 - om) * (-1.0)          # lvd.py:132
    bg_alpha = bg.expand(B, -1, -1, -1)                                   # lvd.py:127
    with ref_loader.stable_sort():
        grid = warper(leaves["obj_pose"], leaves["bg_pose"])
    # unpatched run: hit masks / fields must not depend on the tie rule
    with torch.no_grad():
        grid_unpatched = warper(d["obj_pose"], d["bg_pose"])
    occ = lvd.compute_occ(leaves["occ_score"])
    out = lvd.forward(input=leaves["input"], grid=grid, occ=occ, obj_alpha=obj_alpha, bg_alpha=bg_alpha,
                      ctx_ts=d["ctx_ts"], pred_ts=d["pred_ts"], cls=leaves["cls"], mode="decode_output")
    output, flow, a_unflt, alpha, raw_alpha, raw_output, alpha_ctx = out
    # fixed scalar loss touching every differentiable output; weights are a seeded random projection
    gen = torch.Generator().manual_seed(1234)
    proj = {}
    loss = 0
    for name, t in (("output", output), ("flow", flow), ("alpha", alpha), ("raw_alpha", raw_alpha),
                    ("raw_output", raw_output)):
        w = torch.randn(t.shape, generator=gen)
        proj[name] = w
        loss = loss + (t * w).sum()
    loss.backward()
    # WIF fuse tail on a seeded fake UNet output (wif.py:50-54 via the real WIF.forward with a stub unet)
    Bq, Tcq, Tpq, Cq, Hq, Wq = raw_output.shape
    unet_out = torch.randn(Bq * Tpq * Tcq, 5, Hq, Wq, generator=gen)
    wif_self = types.SimpleNamespace(score=True, ab=True, unet=lambda x: unet_out)
    fused = ns.wif.WIF.forward(wif_self, raw_output.detach())
    res = dict(
        tgt_grid_obj=grid[0], src_grid_obj=grid[1], tgt_grid_bg=grid[2], src_grid_bg=grid[3],
        src_grid_obj_unpatched=grid_unpatched[1], src_grid_bg_unpatched=grid_unpatched[3],
        occ=occ, output=output, flow=flow, alpha=alpha, raw_alpha=raw_alpha, raw_output=raw_output,
        alpha_ctx=alpha_ctx, wif_unet_out=unet_out.view(Bq, Tpq, Tcq, 5, Hq, Wq), wif_fused=fused, loss=loss.detach())
    if a_unflt is not None:
        res["alpha_unflt"] = a_unflt
    for k, v in leaves.items():
        res["grad_" + k] = v.grad
    for k, v in proj.items():
        res["proj_" + k] = v
    for k in ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls", "ctx_ts", "pred_ts"):
        res["in_" + k] = d[k]
    return {k: v.detach().cpu().numpy() for k, v in res.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, (kw, B, T, Tc, smooth) in CASES.items():
        if only and name not in only:
            continue
        cfg = wo.PathConfig(**kw)
        res = run_reference(cfg, B, T, Tc, smooth)
        meta = dict(kw, B=B, T=T, Tc=Tc, smooth=smooth)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, meta=np.array(repr(meta)), **res)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  loss={float(res['loss']):.6f}")


if __name__ == "__main__":
    main()
