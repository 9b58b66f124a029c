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
    # The REAL structure of the benchmarked configurations at reduced resolution: 16 objects with 4x4
    # control points on 64x64 canvases, the 8x16 (Cityscapes, K = 131) / 8x26 (KITTI, K = 211) background systems, 4
    # contexts, 20 / 19 classes, scale_hd 4 / 2.  Stored lean (LEAN below): projections are re-drawn from the seed and
    # alpha_ctx is read back from raw_output by the tests.
    "city_real": (dict(dim=16, load_dim=64, aspect_ratio=2.0), 1, 5, 4, False),
    "kitti_real": (dict(dim=16, load_dim=32, aspect_ratio=3.25, latent_shape=(8, 26), num_lyt=19), 1, 6, 4, True),
}
LEAN = ("city_real", "kitti_real")
PROJ_SEED = 1234
PROJ_ORDER = ("output", "flow", "alpha", "raw_alpha", "raw_output")


def draw_projections(shapes):
    """The seeded random projections of the fixture loss, in PROJ_ORDER, then the fake UNet output (B*Tp*Tc, 5, Hd, Wd).
    `shapes`: name -> shape for PROJ_ORDER.  Shared with tests/parity.py (lean fixtures do not store them)."""
    gen = torch.Generator().manual_seed(PROJ_SEED)
    proj = {n: torch.randn(tuple(shapes[n]), generator=gen) for n in PROJ_ORDER}
    Bq, Tcq, Tpq, Cq, Hq, Wq = shapes["raw_output"]
    unet_out = torch.randn(Bq * Tpq * Tcq, 5, Hq, Wq, generator=gen)
    return proj, unet_out


from ref_runner import ref_opt, ref_modules, ctx_pred  # noqa: E402,F401


def run_reference(cfg: wo.PathConfig, B, T, Tc, smooth, seed=0, lean=False):
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
    outs = dict(output=output, flow=flow, alpha=alpha, raw_alpha=raw_alpha, raw_output=raw_output)
    proj, unet_out = draw_projections({n: t.shape for n, t in outs.items()})
    loss = 0
    for name in PROJ_ORDER:
        loss = loss + (outs[name] * proj[name]).sum()
    loss.backward()
    # WIF fuse tail on a seeded fake UNet output (wif.py:50-54 via the real WIF.forward with a stub unet)
    Bq, Tcq, Tpq, Cq, Hq, Wq = raw_output.shape
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
    if lean:
        del res["alpha_ctx"], res["wif_unet_out"]
    else:
        for k, v in proj.items():
            res["proj_" + k] = v
    for k in ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls", "ctx_ts", "pred_ts"):
        res["in_" + k] = d[k]
    return {k: v.detach().cpu().numpy() for k, v in res.items()}


def blur_fixture():
    """f-2: the reference's own `blur` (models/synthesizer.py:1114-1118, torchvision GaussianBlur) on seeded planes, with the
    autograd gradient of a seeded projection: default 23 taps / sigma 3 (:914 blur_sigma), the 3-tap sigma-2 call of :893,
    a ragged size and one barely larger than the reflect margin."""
    import importlib
    ref_loader.load()
    syn = importlib.import_module("models.synthesizer")
    gen = torch.Generator().manual_seed(77)
    res = {}
    for i, (shape, sigma, k) in enumerate((((2, 3, 2, 40, 72), 3.0, 23), ((1, 2, 5, 33, 47), 2.0, 3), ((1, 1, 1, 12, 13), 3.0, 23),
                                           ((2, 1, 3, 64, 128), 1.5, 9))):
        x = torch.randn(shape, generator=gen).requires_grad_(True)
        w = torch.randn(shape, generator=gen)
        y = syn.blur(x, sigma=sigma, kernel_size=k)
        (y * w).sum().backward()
        res.update({f"x{i}": x.detach(), f"w{i}": w, f"y{i}": y.detach(), f"g{i}": x.grad, f"p{i}": torch.tensor([sigma, float(k)])})
    path = os.path.join(OUT, "blur.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in res.items()})
    print(f"blur: {os.path.getsize(path) / 1e6:.2f} MB")


def pack_fixture():
    """f-3: what the reference's dataset class makes of an 8-bit frame and an 8-bit label map -- `BaseDataset.load_rgb_path` /
    `load_layout_path` with the transforms `__getitem__` builds (data/base_dataset.py:167-183, :213-220, :329-372; no
    augmentation, native size) -- on seeded PNG files; `input` = cat([rgb, layout]) as models/synthesizer.py:444 does.
    Stored: the raw bytes and the reference's fp32 result, for two frame sizes (one not a multiple of 4 pixels)."""
    import importlib
    import tempfile
    import PIL.Image
    ref_loader.load()
    bd = importlib.import_module("data.base_dataset")
    rng = np.random.default_rng(5)
    res, tmp = {}, tempfile.mkdtemp(prefix="waldo_pack_")
    for i, (H, W, Nl) in enumerate(((12, 20, 20), (7, 9, 19))):
        rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        lab = rng.integers(0, Nl, (H, W), dtype=np.uint8)
        PIL.Image.fromarray(rgb).save(os.path.join(tmp, "f.png"))
        PIL.Image.fromarray(lab, mode="L").save(os.path.join(tmp, "l.png"))
        fake = types.SimpleNamespace(opt=types.SimpleNamespace(remap_lyt=[], num_lyt=Nl))
        t_rgb = bd.get_transform(H, aspect_ratio=W / H, is_PIL=True)
        t_lyt = bd.get_transform(H, aspect_ratio=W / H, is_PIL=False, normalize=False)
        img = bd.BaseDataset.load_rgb_path(fake, os.path.join(tmp, "f.png"), t_rgb)
        lyt = bd.BaseDataset.load_layout_path(fake, os.path.join(tmp, "l.png"), t_lyt)
        res.update({f"rgb{i}": rgb.transpose(2, 0, 1).copy(), f"lab{i}": lab, f"input{i}": torch.cat([img, lyt], dim=0).numpy()})
    path = os.path.join(OUT, "pack.npz")
    np.savez_compressed(path, **res)
    print(f"pack: {os.path.getsize(path) / 1e3:.1f} kB")


def pose_dis_fixture():
    """f-2: `cell_dis` / `center_dis` (models/synthesizer.py:965-979).  The statements sit inside `Synthesizer.extract_object`
    and cannot be called on their own, so the generator EXECUTES those source lines of the reference file unchanged (read from
    /root/reference at generation time, nothing copied into the repo) in a namespace holding seeded inputs, and lets autograd
    supply the gradients of a seeded combination of the two scalars."""
    import textwrap
    import types as _t
    src = open(os.path.join(ref_loader._find_root(), "models", "synthesizer.py")).read().splitlines()
    first = next(i for i, l in enumerate(src) if "grid = self.net_pe.module.warper.src_grid" in l)
    last = next(i for i, l in enumerate(src) if l.strip().startswith('log_dic["scalar"]["center_dis"] ='))
    assert (first, last) == (964, 978), (first, last)   # lines 965-979, 1-based
    code = compile(textwrap.dedent("\n".join(src[first:last + 1])), "synthesizer.py:965-979", "exec")
    gen = torch.Generator().manual_seed(31)
    res = {}
    for i, (B, T, No, obj_shape, H, W, eps, zeros) in enumerate(((2, 3, 16, (4, 4), 16, 32, 0.0, False), (1, 2, 5, (3, 2), 9, 13, 0.05, False),
                                                                 (1, 2, 16, (4, 4), 16, 32, 0.0, True))):
        ys, xs = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
        grid = torch.stack([xs, ys], dim=-1)[None]                       # (1, H, W, 2) like warper.src_grid
        obj_pose = (torch.rand(B, T, No, obj_shape[0] * obj_shape[1], 2, generator=gen) * 2 - 1).requires_grad_(True)
        mov = torch.rand(B, T, 1, H, W, generator=gen)
        if zeros:                                                        # a blurred 0/1 mask is exactly 0 far from every moving object
            mov = mov * (torch.rand(B, T, 1, H, W, generator=gen) > 0.6).float()
        fg = (torch.rand(B, T, 1, H, W, generator=gen) * 1.1).requires_grad_(True)   # a blurred sum of opacities can pass 1
        opt = _t.SimpleNamespace(num_obj=No, obj_shape=obj_shape, cell_dis_eps=eps, blur_alpha=False, blur_sigma=3.0)
        ns = {"self": _t.SimpleNamespace(opt=opt, net_pe=_t.SimpleNamespace(module=_t.SimpleNamespace(warper=_t.SimpleNamespace(src_grid=grid)))),
              "obj_pose": obj_pose, "mov_obj_mask": mov, "fg_mask": fg, "log_dic": {"scalar": {}}, "blur": None}
        exec(code, ns)
        cell, center = ns["log_dic"]["scalar"]["cell_dis"], ns["log_dic"]["scalar"]["center_dis"]
        (cell * 1.5 + center * 0.75).backward()
        res.update({f"grid{i}": grid, f"pose{i}": obj_pose.detach(), f"mov{i}": mov, f"fg{i}": fg.detach(), f"cell{i}": cell.detach(), f"center{i}": center.detach(),
                    f"d_pose{i}": obj_pose.grad, f"d_fg{i}": fg.grad, f"p{i}": torch.tensor([float(obj_shape[0]), float(obj_shape[1]), eps])})
    path = os.path.join(OUT, "pose_dis.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in res.items()})
    print(f"pose_dis: {os.path.getsize(path) / 1e6:.2f} MB")


def obj_flow_fixture():
    """f-2: `obj_flow` (models/synthesizer.py:865-868), produced like pose_dis_fixture: the reference's own source lines executed
    unchanged on seeded inputs, autograd for the gradient with respect to the layer opacities."""
    import textwrap
    src = open(os.path.join(ref_loader._find_root(), "models", "synthesizer.py")).read().splitlines()
    first = next(i for i, l in enumerate(src) if l.strip().startswith("a = (rec_output_alpha[:, :, 1:] + 1) / 2 + 1e-6"))
    last = next(i for i, l in enumerate(src) if l.strip().startswith('log_dic["scalar"]["obj_flow"] ='))
    assert (first, last) == (864, 867), (first, last)   # lines 865-868, 1-based
    code = compile(textwrap.dedent("\n".join(src[first:last + 1])), "synthesizer.py:865-868", "exec")
    gen = torch.Generator().manual_seed(41)
    res = {}
    for i, (B, T, L, H, W, sat) in enumerate(((2, 3, 17, 16, 32, False), (1, 2, 3, 9, 13, False), (1, 1, 6, 12, 20, True))):
        alpha = torch.tanh(2 * torch.randn(B, T, L, H, W, generator=gen))
        if sat:   # saturated opacities (+-1 exactly): layers that are empty everywhere keep sum_a = 1e-6 HW
            alpha = torch.where(alpha > 0.3, torch.ones(()), -torch.ones(()))
        alpha = alpha.requires_grad_(True)
        flow = torch.randn(B, T, 2, H, W, generator=gen) * 0.1
        ns = {"rec_output_alpha": alpha, "real_flow": flow, "log_dic": {"scalar": {}}}
        exec(code, ns)
        val = ns["log_dic"]["scalar"]["obj_flow"]
        (val * 3.0).backward()
        res.update({f"alpha{i}": alpha.detach(), f"flow{i}": flow, f"val{i}": val.detach(), f"d_alpha{i}": alpha.grad})
    path = os.path.join(OUT, "obj_flow.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in res.items()})
    print(f"obj_flow: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]
    if not only or "blur" in only:
        blur_fixture()
    if not only or "pack" in only:
        pack_fixture()
    if not only or "pose_dis" in only:
        pose_dis_fixture()
    if not only or "obj_flow" in only:
        obj_flow_fixture()
    for name, (kw, B, T, Tc, smooth) in CASES.items():
        if only and name not in only:
            continue
        cfg = wo.PathConfig(**kw)
        res = run_reference(cfg, B, T, Tc, smooth, lean=name in LEAN)
        meta = dict(kw, B=B, T=T, Tc=Tc, smooth=smooth)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, meta=np.array(repr(meta)), **res)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  loss={float(res['loss']):.6f}")


if __name__ == "__main__":
    main()
