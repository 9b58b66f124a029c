"""Workload definitions of the benchmarked configurations (BASELINE.json `configs`) and their seeded synthetic inputs
 -- the product-side helper bench.py and the examples use.  It restates, for the product, what the
oracle (test infrastructure, never imported from here) also defines; tests/test_workloads.py pins the two against each other.
"""
from __future__ import annotations

import math
import types
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn.functional as F


@dataclass
class PathConfig:
    """The subset of the reference's options the path reads (models/nets/lvd.py:470-499,
    :15-44; defaults = scripts/cityscapes/test.sh resolved by tools/options.py)."""
    dim: int = 128
    load_dim: int = 512
    aspect_ratio: float = 2.0
    num_obj: int = 16
    obj_shape: Tuple[int, int] = (4, 4)
    latent_shape: Tuple[int, int] = (8, 16)
    patch_size: int = 16
    scale_factor: float = 1.0
    num_lyt: int = 20
    weight_cls: bool = True
    min_cls: float = 0.1
    include_self: bool = False
    restrict_to_ctx: bool = True
    use_disocc: bool = False
    no_filter: bool = False
    allow_ghost: bool = False
    pad_obj_alpha: int = 3
    pad_bg_alpha: int = 3

    @property
    def lo_shape(self):  # H, W           lvd.py:479
        return (self.dim, int(self.dim * self.aspect_ratio))

    @property
    def hd_shape(self):  # Hd, Wd         lvd.py:480
        if self.load_dim > 0:
            return (self.load_dim, int(self.load_dim * self.aspect_ratio))
        return self.lo_shape

    @property
    def obj_hw(self):  # Ho, Wo           lvd.py:478
        return (int(self.obj_shape[0] * self.patch_size * self.scale_factor),
                int(self.obj_shape[1] * self.patch_size * self.scale_factor))

    @property
    def scale_hd(self):  # lvd.py:495
        return self.load_dim / self.dim if self.load_dim > 0 else 1

    @property
    def fast(self):  # lvd.py:494
        return self.load_dim == 0



def _pixel_grid(h: int, w: int, dtype=torch.float32) -> torch.Tensor:
    """tools/utils.py:293-297 -- pixel-centre normalised lattice, (1,h,w,2), last dim (x,y).
    The reference builds it with fp32 `linspace`; we do the same and then cast, so that the
    fp32 oracle is bit-identical and the fp64 twin starts from the same lattice."""
    xs = torch.linspace(-1.0 + 1.0 / w, 1.0 - 1.0 / w, w)
    ys = torch.linspace(-1.0 + 1.0 / h, 1.0 - 1.0 / h, h)
    g = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)], dim=-1)
    return g.unsqueeze(0).to(dtype)



def synth_inputs(cfg: PathConfig, B: int, T: int, Tc: int, seed: int = 0, dtype=torch.float32, smooth: bool = False,
                 radius: float = 0.5):
    """Seeded synthetic inputs of SURVEY.md §8(d).  Returns a dict of CPU tensors.
    `radius` = radius of the circle the object centres sit on (0.5 in the survey; the small golden
    cases use 0.2 so that the few objects overlap and the occlusion matrix matters)."""
    g = torch.Generator().manual_seed(seed)
    Hd, Wd = cfg.hd_shape
    H, W = cfg.lo_shape
    Ho, Wo = cfg.obj_hw
    No, Nl = cfg.num_obj, cfg.num_lyt
    Lo = cfg.obj_shape[0] * cfg.obj_shape[1]
    L = cfg.latent_shape[0] * cfg.latent_shape[1]
    if smooth:
        def lowpass(c, amp):
            z = torch.randn(B * T, c, max(Hd // 16, 2), max(Wd // 16, 2), generator=g) * amp
            return F.interpolate(z, size=(Hd, Wd), mode="bicubic", align_corners=False).view(B, T, c, Hd, Wd)
        vid = lowpass(3, 0.6).clamp(-1, 1)
        lyt = lowpass(Nl, 3.0)
    else:
        vid = torch.rand(B, T, 3, Hd, Wd, generator=g) * 2 - 1
        lab = torch.randint(0, Nl, (B, T, Hd, Wd), generator=g)
        lyt = 5 * (2 * F.one_hot(lab, Nl).permute(0, 1, 4, 2, 3).float() - 1)
    inp = torch.cat([vid, lyt], dim=2)
    obj_alpha = torch.tanh(2 * torch.randn(B, No, 1, Ho, Wo, generator=g))
    theta = 2 * math.pi * torch.arange(No) / No
    centre = radius * torch.stack([theta.cos(), theta.sin()], dim=-1)                       # No 2
    base = 0.25 * _pixel_grid(*cfg.obj_shape).view(1, 1, 1, Lo, 2) * torch.tensor([1.0, cfg.aspect_ratio])
    drift = torch.linspace(0, 0.1, T).view(1, T, 1, 1, 1)
    obj_pose = base + centre.view(1, 1, No, 1, 2) + drift + 0.01 * torch.randn(B, T, No, Lo, 2, generator=g)
    bg_pose = 1.2 * _pixel_grid(*cfg.latent_shape).view(1, 1, 1, L, 2) + 0.005 * torch.randn(B, T, 1, L, 2, generator=g)
    occ_score = torch.randn(B, T, No, generator=g)
    cls = torch.randn(B, No, Nl, generator=g).softmax(dim=-1)
    Tp = T - Tc
    ctx_ts = torch.arange(Tc).view(1, Tc, 1).expand(B, Tc, Tp)
    pred_ts = torch.arange(Tc, T)
    d = dict(input=inp, obj_alpha_raw=obj_alpha, obj_pose=obj_pose, bg_pose=bg_pose, occ_score=occ_score, cls=cls)
    d = {k: v.to(dtype) for k, v in d.items()}
    d.update(ctx_ts=ctx_ts, pred_ts=pred_ts)
    return d


def make_opt(cfg: PathConfig):
    """The option namespace the reference's Warper reads (models/nets/lvd.py:470-499), filled from a PathConfig."""
    return types.SimpleNamespace(
        latent_shape=list(cfg.latent_shape), obj_shape=list(cfg.obj_shape), time_dropout=False, num_obj=cfg.num_obj,
        patch_size=cfg.patch_size, scale_factor=cfg.scale_factor, dim=cfg.dim, aspect_ratio=cfg.aspect_ratio,
        load_dim=cfg.load_dim, num_perm_grid=1, normalize_alpha=False, use_lyt_filtering=True, use_lyt_opacity=True,
        weight_cls=cfg.weight_cls, min_cls=cfg.min_cls, include_self=cfg.include_self, no_filter=cfg.no_filter,
        allow_ghost=cfg.allow_ghost, use_disocc=cfg.use_disocc, pad_obj_alpha=cfg.pad_obj_alpha,
        pad_bg_alpha=cfg.pad_bg_alpha)


WORKLOADS = {
    # name: (PathConfig kwargs, B per GPU, T, Tc, backward, label)          BASELINE.json configs[1], [3], [2], [4]
    "city_train": (dict(), 8, 5, 4, True, "cityscapes 512x1024 fwd+bwd B=8/GPU Tc=4->Tp=1"),
    "city_rollout": (dict(), 1, 14, 4, False, "cityscapes 512x1024 rollout fwd B=1/GPU Tc=4->Tp=10"),
    "kitti_rollout": (dict(dim=128, load_dim=256, aspect_ratio=3.25, latent_shape=(8, 26), num_lyt=19), 1, 9, 4, False,
                      "kitti 256x832 rollout fwd B=1/GPU Tc=4->Tp=5"),
    # the reference ships no UCF-Sports / H3.6M script: SURVEY 8d C5 shape
    "nonrigid_train": (dict(dim=128, load_dim=256, aspect_ratio=1.0, latent_shape=(8, 8)), 8, 5, 4, True,
                       "non-rigid 256x256 fwd+bwd B=8/GPU Tc=4->Tp=1"),
}


def workload(name: str):
    if name not in WORKLOADS:
        raise SystemExit(f"unknown workload {name}")
    kw, B, T, Tc, backward, label = WORKLOADS[name]
    return PathConfig(**kw), dict(B=B, T=T, Tc=Tc, backward=backward, label=label)


def alg_bytes(cfg: PathConfig, B, Tc, Tp, backward, elem=4):
    """Algorithmic HBM bytes of one step (SURVEY.md 8d / BASELINE.md 3): (forward, backward)."""
    Hd, Wd = cfg.hd_shape
    px, s = Hd * Wd, elem
    C, L, Nl = 3 + cfg.num_lyt, cfg.num_obj + 1, cfg.num_lyt
    fwd = px * s * (B * Tc * Tp * (C + L) + B * Tc * Tp * ((C + L) + 2 + 1) + B * Tp * (C + 1) + B * Tc * (Nl + L))
    bwd = px * s * (B * Tc * Tp * ((C + L) + 2 + C + L) + B * Tp * (C + 1) + B * Tc * (C + L))
    return fwd, (bwd if backward else 0)
