"""Run the UNMODIFIED reference's hot path (imported by oracle/ref_loader.py from /root/reference, or from the staged
copy oracle/_ref on the GPU box) on seeded synthetic inputs, on CPU or -- stock PyTorch, the honest GPU comparator --
on CUDA.

TEST / BENCH INFRASTRUCTURE.  Users: oracle/make_golden.py (fixtures), tests/ (the reference's own code as the checker
at the benchmarked shapes), bench.py --impl reference / reference-gpu.  Never imported by waldo_b200/.

What runs is the reference's own `Warper.forward` (models/nets/lvd.py:855-870), `LVD.compute_occ` (:59-68),
`LVD.forward(mode="decode_output")` (:141-153) and `WIF.forward` (models/nets/wif.py:37-57, with the UNet replaced by a
fixed tensor where stated).  The LVD object is a bare namespace carrying the fields those methods read: constructing
the real LVD would build the (out-of-scope) encoder / transformer stacks.
"""
from __future__ import annotations

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import waldo_oracle as wo  # noqa: E402

available = ref_loader.available

OPT_FIELDS = ("dim", "load_dim", "aspect_ratio", "num_obj", "patch_size", "scale_factor", "num_lyt", "weight_cls",
              "min_cls", "include_self", "restrict_to_ctx", "use_disocc", "no_filter", "allow_ghost",
              "pad_obj_alpha", "pad_bg_alpha")


def ref_opt(cfg: wo.PathConfig):
    """The reference's own option namespace (its parser on its own cityscapes/test.sh) carrying this PathConfig."""
    opt = ref_loader.parse_opts("scripts/cityscapes/test.sh")
    for k in OPT_FIELDS:
        setattr(opt, k, getattr(cfg, k))
    opt.obj_shape = list(cfg.obj_shape)
    opt.latent_shape = list(cfg.latent_shape)
    opt.num_perm_grid = 1
    opt.time_dropout = False
    return opt


def fake_lvd(ns, cfg: wo.PathConfig, warper, device="cpu"):
    """An object on which the reference's unbound LVD.forward / LVD.compute_occ run unchanged."""
    fake = types.SimpleNamespace(
        warper=warper, restrict_to_ctx=cfg.restrict_to_ctx, use_disocc=cfg.use_disocc, include_self=cfg.include_self,
        diag=torch.eye(cfg.num_obj, device=device)[None, None], remove_obj=False, freeze_obj=False)
    fake.compute_occ = lambda s: ns.lvd.LVD.compute_occ(fake, s)
    fake.forward = lambda **kw: ns.lvd.LVD.forward(fake, **kw)
    return fake


def ref_modules(cfg: wo.PathConfig, device="cpu"):
    ns = ref_loader.load()
    warper = ns.lvd.Warper(ref_opt(cfg)).to(device)
    return warper, fake_lvd(ns, cfg, warper, device)


def ctx_pred(cfg, B, T, Tc):
    if cfg.restrict_to_ctx:
        Tp = T - Tc
        return torch.arange(Tc).view(1, Tc, 1).expand(B, Tc, Tp), torch.arange(Tc, T)
    # train_lvd: ctx_mode "prev" (synthesizer.py:833-839): the previous frame is the context of every frame
    ts = torch.roll(torch.arange(T), 1).view(1, 1, T).expand(B, 1, T)
    return ts, torch.arange(T)


LEAVES = ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")


class RefPath:
    """The reference's path, set up once for a config and a device."""

    def __init__(self, cfg: wo.PathConfig, device="cpu"):
        self.cfg, self.device = cfg, torch.device(device)
        self.ns = ref_loader.load()
        self.warper, self.lvd = ref_modules(cfg, self.device)
        om, bg = wo.alpha_masks(cfg)
        self.om = om.to(self.device) if torch.is_tensor(om) else om
        self.bg = bg.to(self.device)

    def stage_a(self, obj_alpha_raw, obj_pose, bg_pose, occ_score, stable=True):
        """estimate_alpha_grid_occ after the decoder (lvd.py:127-135)."""
        B = obj_alpha_raw.size(0)
        obj_alpha = self.om * obj_alpha_raw + (1 - self.om) * (-1.0)     # lvd.py:132
        bg_alpha = self.bg.expand(B, -1, -1, -1)                          # lvd.py:127
        if stable:
            with ref_loader.stable_sort():
                grid = self.warper(obj_pose, bg_pose)
        else:
            grid = self.warper(obj_pose, bg_pose)
        occ = self.lvd.compute_occ(occ_score)
        return occ, obj_alpha, bg_alpha, grid

    def decode(self, input, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts):
        return self.lvd.forward(input=input, grid=grid, occ=occ, obj_alpha=obj_alpha, bg_alpha=bg_alpha,
                                ctx_ts=ctx_ts, pred_ts=pred_ts, cls=cls, mode="decode_output")

    def chain(self, d, backward=False, stable=True, loss_fn=None):
        """control points -> grids -> occ -> decode_output (-> loss.backward): one step of the hot path."""
        dev = self.device
        lv = {k: d[k].to(dev).clone().requires_grad_(backward) for k in LEAVES}
        with torch.set_grad_enabled(backward):
            occ, oa, ba, grid = self.stage_a(lv["obj_alpha_raw"], lv["obj_pose"], lv["bg_pose"], lv["occ_score"], stable)
            out = self.decode(lv["input"], grid, occ, oa, ba, lv["cls"], d["ctx_ts"].to(dev), d["pred_ts"].to(dev))
            loss = None
            if backward:
                loss = loss_fn(out) if loss_fn else out[0].abs().mean() + out[1].abs().mean()
                loss.backward()
        return out, grid, occ, {k: v.grad for k, v in lv.items()}, loss
