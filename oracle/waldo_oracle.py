"""CPU oracle for the WALDO warp+composite hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import this file.
`waldo_b200/` never does (tests/test_no_oracle_in_product.py enforces it).

What it is: a from-scratch restatement, stage by stage,
of the reference's algorithm for the path
    control points -> TPS grids -> inverse warp -> occlusion matrix
    -> context alpha prep -> per-layer flow -> warp + composite -> context fusion
    -> WIF fuse tail,
written against torch CPU tensors so that (a) it is dtype-generic (fp32 to compare
with, fp64 to arbitrate, SURVEY.md §8d tiers) and (b) `torch.autograd` supplies the
gradient oracle.  The third-party arithmetic the reference relies on (ATen
`grid_sampler_2d`, `upsample_bilinear2d`; reference pin torch 1.11, here torch 2.11,
semantics unchanged -- SURVEY.md §7 item 7) is used through the same public calls
(`F.grid_sample`, `F.interpolate`); `bil0_explicit` / `resize_explicit` restate their
published formulas from first principles and tests pin one against the other.

Pinning: the reference ships no tests or golden vectors.  The oracle is
pinned (1) against the reference itself, imported unmodified from /root/reference in
the build container, by `oracle/make_golden.py`, which writes the committed fixtures in
tests/golden/ (reference outputs, not oracle outputs), and (2) against the KATs of
SURVEY.md §4.  tests/test_oracle_golden.py replays both without /root/reference.

Tie rule: where several lattice samples round to the same target cell in
the inverse warp, the LOWEST source index wins (== the reference with a stable sort).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- config
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


# --------------------------------------------------------------------------- a-0 helpers
def pixel_grid(h: int, w: int, dtype=torch.float32) -> torch.Tensor:
    """tools/utils.py:293-297 -- pixel-centre normalised lattice, (1,h,w,2), last dim (x,y).
    The reference builds it with fp32 `linspace`; we do the same and then cast, so that the
    fp32 oracle is bit-identical and the fp64 twin starts from the same lattice."""
    xs = torch.linspace(-1.0 + 1.0 / w, 1.0 - 1.0 / w, w)
    ys = torch.linspace(-1.0 + 1.0 / h, 1.0 - 1.0 / h, h)
    g = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)], dim=-1)
    return g.unsqueeze(0).to(dtype)


def gaussian3(k: int = 3, sigma_div: float = 6.0) -> torch.Tensor:
    """tools/utils.py:273-291 -- kxk Gaussian, sigma = k / sigma_div, normalised to sum 1."""
    c = torch.arange(k, dtype=torch.float32) - (k - 1) / 2.0
    var = (k / sigma_div) ** 2.0
    g = (1.0 / (2.0 * math.pi * var)) * torch.exp(-(c.view(1, k) ** 2 + c.view(k, 1) ** 2) / (2 * var))
    return g / g.sum()


def bil0(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """`F.grid_sample(x, g)` with the defaults the reference uses everywhere
    (bilinear, zeros padding, align_corners=False): lvd.py:548,559,678,801,837."""
    return F.grid_sample(x, g, mode="bilinear", padding_mode="zeros", align_corners=False)


def bil0_explicit(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """First-principles restatement of ATen grid_sampler_2d (bilinear/zeros/align_corners=False),
    SURVEY.md Appendix C.  x (N,C,H,W), g (N,h,w,2) -> (N,C,h,w)."""
    N, C, H, W = x.shape
    gx = ((g[..., 0] + 1) * W - 1) / 2
    gy = ((g[..., 1] + 1) * H - 1) / 2
    x0 = torch.floor(gx)
    y0 = torch.floor(gy)
    out = torch.zeros(N, C, *g.shape[1:3], dtype=x.dtype)
    flat = x.reshape(N, C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            wx = (gx - x0) if dx else (x0 + 1 - gx)
            wy = (gy - y0) if dy else (y0 + 1 - gy)
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long().view(N, 1, -1).expand(-1, C, -1)
            v = flat.gather(2, idx).view(N, C, *g.shape[1:3])
            out = out + v * (wx * wy * ok).unsqueeze(1)
    return out


def resize(x: torch.Tensor, factor: float) -> torch.Tensor:
    """lvd.py:175-179 `scale`: bilinear `F.interpolate(scale_factor=...)` on the last two dims of
    a tensor flattened to 4-D; identity when factor == 1."""
    if factor == 1:
        return x
    lead = x.shape[:-3]
    y = F.interpolate(x.reshape(-1, *x.shape[-3:]), scale_factor=factor, mode="bilinear", align_corners=False)
    return y.reshape(*lead, *y.shape[-3:])


def resize_explicit(x: torch.Tensor, out_hw: Sequence[int], ratio_hw: Sequence[float]) -> torch.Tensor:
    """First-principles restatement of ATen upsample_bilinear2d (align_corners=False), Appendix C:
    src = max(r*(dst+0.5)-0.5, 0), i0 = floor(src), i1 = min(i0+1, in-1), lam = src-i0."""
    def axis(n_in, n_out, r):
        d = torch.arange(n_out, dtype=x.dtype)
        s = (r * (d + 0.5) - 0.5).clamp(min=0)
        i0 = s.floor().long().clamp(max=n_in - 1)
        i1 = (i0 + 1).clamp(max=n_in - 1)
        return i0, i1, s - i0
    H, W = x.shape[-2:]
    y0, y1, ly = axis(H, out_hw[0], ratio_hw[0])
    x0, x1, lx = axis(W, out_hw[1], ratio_hw[1])
    rows = x[..., y0, :] * (1 - ly).view(-1, 1) + x[..., y1, :] * ly.view(-1, 1)
    return rows[..., x0] * (1 - lx) + rows[..., x1] * lx


# --------------------------------------------------------------------------- a-1 TPS
def tps_phi(p: torch.Tensor, q: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """warp.py:15-18 -- phi = 0.5 * d * log(d + eps), d = |p|^2 + |q|^2 - 2 p.q (the expanded form,
    which is NOT exactly the squared distance in floating point)."""
    d = (p * p).sum(-1).view(-1, 1) + (q * q).sum(-1).view(1, -1) - 2 * p @ q.t()
    return 0.5 * d * (d + eps).log()


@dataclass
class TPSBasis:
    """warp.py:22-47: the two constant matrices of a TPSWarp."""
    h: int
    w: int
    inverse_kernel: torch.Tensor  # (N+3, N+3)
    tgt_grid_repr: torch.Tensor   # (h*w, N+3)


def tps_basis(h: int, w: int, ctrl: torch.Tensor, dtype=torch.float32) -> TPSBasis:
    """warp.py:22-47.  System matrix [[phi(P,P) 1 P],[1^T 0 0],[P^T 0 0]] inverted once;
    lattice representation [phi(lattice,P) | 1 | lattice].  Built in fp32 like the reference
    (hard-coded .float(), warp.py:28-42) then cast, so the fp64 twin uses the same constants."""
    ctrl = ctrl.float()
    n = ctrl.shape[0]
    a = torch.zeros(n + 3, n + 3)
    a[:n, :n] = tps_phi(ctrl, ctrl)
    a[:n, n] = 1
    a[n, :n] = 1
    a[:n, n + 1:] = ctrl
    a[n + 1:, :n] = ctrl.t()
    inv = torch.inverse(a)
    lattice = pixel_grid(h, w).view(-1, 2)
    rep = torch.cat([tps_phi(lattice, ctrl), torch.ones(h * w, 1), lattice], dim=1)
    return TPSBasis(h, w, inv.to(dtype), rep.to(dtype))


def tps_eval(basis: TPSBasis, pts: torch.Tensor) -> torch.Tensor:
    """warp.py:49-55.  pts (n,N,2) -> (n,h,w,2); keeps the reference's association
    (inverse_kernel @ [pts;0]) first, then tgt_grid_repr @ mapping."""
    n = pts.shape[0]
    padded = torch.cat([pts, pts.new_zeros(n, 3, 2)], dim=1)
    mapping = basis.inverse_kernel @ padded
    return (basis.tgt_grid_repr @ mapping).view(n, basis.h, basis.w, 2)


# --------------------------------------------------------------------------- a-2 inverse warp
@dataclass
class InverseWarpTrace:
    grid: torch.Tensor            # (n,Ht,Wt,2) the inverse map (normalised coords)
    field: torch.Tensor           # (n,Hs*Ws... as Ht*Wt) int64 target cell per lattice sample, -1 = outside
    hit: torch.Tensor             # (n,Ht,Wt) bool: cells that received a sample
    known: torch.Tensor           # (n,Ht,Wt) bool: cells known after fill (+erode), cropped
    winner: torch.Tensor          # (n,Ht*Wt) int64: winning sample index per cell, Ht*Wt = none


def inverse_warp(fwd_grid: torch.Tensor, tgt_hw: Sequence[int], niter: int = 5, erode: bool = True,
                 trace: bool = False):
    """warp.py:71-174 (num_perm == 1 branch, pad=True).

    fwd_grid (n,Hs,Ws,2): forward map lattice -> image coords.  Output (n,Ht,Wt,2): for every
    image cell the lattice coordinate that maps onto it; unknown cells point far outside.
    Steps follow SURVEY.md Appendix B A2.1-A2.6.
    """
    n, Hs, Ws, _ = fwd_grid.shape
    Ht, Wt = tgt_hw
    dt = fwd_grid.dtype
    # 1. displacement, resampled on the target lattice, in pixels            warp.py:76-79
    disp = (fwd_grid - pixel_grid(Hs, Ws, dt)).permute(0, 3, 1, 2)
    disp = F.interpolate(disp, size=(Ht, Wt), mode="bilinear", align_corners=False)
    dx = disp[:, 0].reshape(n, -1) * Wt / 2
    dy = disp[:, 1].reshape(n, -1) * Ht / 2
    # 2. rounded landing cell (half-to-even), -1 when outside                warp.py:80-88
    col = torch.arange(Wt, dtype=dt).repeat(Ht)
    row = torch.arange(Ht, dtype=dt).repeat_interleave(Wt)
    tx = (col + dx).round().long()
    ty = (row + dy).round().long()
    inside = (tx >= 0) & (ty >= 0) & (tx <= Wt - 1) & (ty <= Ht - 1)
    cell = torch.where(inside, ty * Wt + tx, torch.full_like(tx, -1))
    # 3. one survivor per cell: lowest sample index                          warp.py:113-123 + tie rule
    P = Ht * Wt
    src_idx = torch.arange(P).expand(n, P)
    slot = torch.where(inside, cell, torch.full_like(cell, P))               # dump slot P for outsiders
    winner = torch.full((n, P + 1), P, dtype=torch.long).scatter_reduce(1, slot, src_idx, "amin")[:, :P]
    hit = winner < P
    take = winner.clamp(max=P - 1)
    zero = dx.new_zeros(())
    inv_dx = torch.where(hit, -dx.gather(1, take), zero).view(n, Ht, Wt)
    inv_dy = torch.where(hit, -dy.gather(1, take), zero).view(n, Ht, Wt)
    mask = hit.view(n, Ht, Wt)
    # 4. pad by niter+1 and grow: frontier cells take the normalised-Gaussian mean   warp.py:125-151
    m = niter + 1
    inv_dx = F.pad(inv_dx, (m, m, m, m))
    inv_dy = F.pad(inv_dy, (m, m, m, m))
    mask = F.pad(mask, (m, m, m, m))
    g = gaussian3(3).to(dt).view(1, 1, 3, 3)

    def nb4(mk):  # any 4-neighbour set (no wrap)
        out = torch.zeros_like(mk)
        out[:, 1:] |= mk[:, :-1]
        out[:, :-1] |= mk[:, 1:]
        out[:, :, 1:] |= mk[:, :, :-1]
        out[:, :, :-1] |= mk[:, :, 1:]
        return out

    for _ in range(niter):
        frontier = ~mask & nb4(mask)
        sx = F.conv2d(inv_dx.unsqueeze(1), g, padding=1).squeeze(1)
        sy = F.conv2d(inv_dy.unsqueeze(1), g, padding=1).squeeze(1)
        sw = F.conv2d(mask.to(dt).unsqueeze(1), g, padding=1).squeeze(1)
        inv_dx = torch.where(frontier, sx / sw, inv_dx)
        inv_dy = torch.where(frontier, sy / sw, inv_dy)
        mask = mask | frontier
    # 5. erosion (objects only)                                              warp.py:153-162
    if erode:
        for _ in range(niter):
            mask = mask & ~(mask & nb4(~mask))
    # 6. sentinel + crop + back to normalised coords                         warp.py:164-174
    inv_dx = torch.where(mask, inv_dx, inv_dx.new_full((), 2.0 * Wt))[:, m:-m, m:-m]
    inv_dy = torch.where(mask, inv_dy, inv_dy.new_full((), 2.0 * Ht))[:, m:-m, m:-m]
    out = pixel_grid(Ht, Wt, dt) + torch.stack([inv_dx * 2 / Wt, inv_dy * 2 / Ht], dim=3)
    if trace:
        return InverseWarpTrace(out, cell, hit.view(n, Ht, Wt), mask[:, m:-m, m:-m], winner)
    return out


# --------------------------------------------------------------------------- a-3 Warper.forward
@dataclass
class WarperState:
    """Constant buffers of the reference Warper (lvd.py:470-499)."""
    cfg: PathConfig
    tps_obj: TPSBasis
    tps_bg: TPSBasis
    dtype: torch.dtype = torch.float32


def make_state(cfg: PathConfig, dtype=torch.float32) -> WarperState:
    Ho, Wo = cfg.obj_hw
    H, W = cfg.lo_shape
    obj_ctrl = pixel_grid(*cfg.obj_shape).view(-1, 2)       # lvd.py:473 tgt_pts
    bg_ctrl = pixel_grid(*cfg.latent_shape).view(-1, 2)     # lvd.py:472 src_pts
    return WarperState(cfg, tps_basis(Ho, Wo, obj_ctrl, dtype), tps_basis(H, W, bg_ctrl, dtype), dtype)


def warper_forward(st: WarperState, obj_pose: torch.Tensor, bg_pose: torch.Tensor):
    """lvd.py:855-870.  obj_pose (B,T,No,Lo,2), bg_pose (B,T,1,L,2) -> 4-tuple `grid`."""
    B, T, No = obj_pose.shape[:3]
    H, W = st.cfg.lo_shape
    Ho, Wo = st.cfg.obj_hw
    tgo = tps_eval(st.tps_obj, obj_pose.reshape(B * T * No, -1, 2))
    sgo = inverse_warp(tgo, (H, W), erode=True)
    tgb = tps_eval(st.tps_bg, bg_pose.reshape(B * T, -1, 2))
    sgb = inverse_warp(tgb, (H, W), erode=False)
    return (tgo.view(B, T, No, Ho, Wo, 2), sgo.view(B, T, No, H, W, 2),
            tgb.view(B, T, H, W, 2), sgb.view(B, T, H, W, 2))


# --------------------------------------------------------------------------- a-4 occlusion matrix
def compute_occ(occ_score: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """lvd.py:59-68.  (B,T,No) -> (B,T,No+1,No+1); entry [j,i] = how much layer j hides layer i."""
    B, T, No = occ_score.shape
    e = torch.exp(-occ_score ** 2) + eps
    pair = e.unsqueeze(-1) / (e.unsqueeze(-1) + e.unsqueeze(-2))
    pair = pair - 0.5 * torch.eye(No, dtype=occ_score.dtype)
    occ = occ_score.new_zeros(B, T, No + 1, No + 1)
    occ[:, :, 1:, 0] = 1            # every object hides the background
    occ[:, :, 1:, 1:] = pair        # row 0 (background hides nobody) stays 0
    return occ


def alpha_masks(cfg: PathConfig, dtype=torch.float32):
    """lvd.py:25-44: object alpha pad mask (1,1,1,Ho,Wo) and the constant background alpha (1,1,H,W)."""
    Ho, Wo = cfg.obj_hw
    H, W = cfg.lo_shape
    om = torch.ones(Ho, Wo, dtype=dtype)
    if cfg.pad_obj_alpha > 0:
        p = int(cfg.pad_obj_alpha * cfg.scale_factor)
        om[:p] = 0; om[-p:] = 0; om[:, :p] = 0; om[:, -p:] = 0
    bg = torch.ones(1, 1, H, W, dtype=dtype)
    if cfg.pad_bg_alpha > 0:
        p = int(cfg.pad_bg_alpha * cfg.scale_factor)
        bg[:, :, :p] = -1; bg[:, :, -p:] = -1; bg[:, :, :, :p] = -1; bg[:, :, :, -p:] = -1
    return om.view(1, 1, 1, Ho, Wo), bg


# --------------------------------------------------------------------------- a-5 layer projection
def project_layers(obj_field: torch.Tensor, bg_field: torch.Tensor, sgo: torch.Tensor, sgb: torch.Tensor) -> torch.Tensor:
    """lvd.py:533-559 with delta=0 (every live call site): sample per-layer canonical fields into the
    image.  obj_field (B,T,No,c,Ho,Wo), bg_field (B,T,c,H,W), sgo (B,T,No,H,W,2), sgb (B,T,H,W,2)
    -> (B,T,No+1,c,H,W) with the background first."""
    B, T, No, c = obj_field.shape[:4]
    H, W = sgo.shape[3:5]
    o = bil0(obj_field.reshape(B * T * No, c, *obj_field.shape[-2:]), sgo.reshape(B * T * No, H, W, 2))
    b = bil0(bg_field.reshape(B * T, c, *bg_field.shape[-2:]), sgb.reshape(B * T, H, W, 2))
    return torch.cat([b.view(B, T, 1, c, H, W), o.view(B, T, No, c, H, W)], dim=2)


def take_time(x: torch.Tensor, ts: torch.Tensor) -> torch.Tensor:
    """lvd.py:462-467 gather_time: x (B,T,...) , ts (B,Tc,Tp) -> (B,Tc,Tp,...)."""
    B, Tc, Tp = ts.shape
    b = torch.arange(B).view(B, 1, 1).expand(B, Tc, Tp)
    return x[b, ts]


def occlude(a: torch.Tensor, occ: torch.Tensor) -> torch.Tensor:
    """lvd.py:763-765 / :806-815.  a (...,L,h,w) opacities, occ (...,L,L):
    out[i] = a[i] * prod_j (1 - a[j] * occ[j,i])."""
    L = a.shape[-3]
    vis = torch.stack([(1 - a * occ[..., :, i, None, None]).prod(dim=-3) for i in range(L)], dim=-3)
    return vis * a


# --------------------------------------------------------------------------- a-6 context alpha prep (B1-B4)
def class_profile(cfg: PathConfig, lyt_lo: torch.Tensor, a_obj_lo: torch.Tensor, cls: Optional[torch.Tensor]):
    """lvd.py:735-742 (== :629-636): alpha-weighted mean layout logits per (b, object).
    lyt_lo (B,Tw,Nl,H,W), a_obj_lo (B,Tw,No,H,W) in [0,1]  ->  M (B,No,Nl)."""
    w = a_obj_lo + 1e-6
    if cfg.weight_cls:
        sm = lyt_lo.softmax(dim=2)                                              # B Tw Nl H W
        w = w * torch.einsum("bkc,btchw->btkhw", cls + cfg.min_cls, sm)
    den = w.sum(dim=(1, 3, 4))                                                  # B No
    num = torch.einsum("btchw,btkhw->bkc", lyt_lo, w)
    return num / den.unsqueeze(-1)


def layout_agreement(profile_probs: torch.Tensor, hd_lyt: torch.Tensor) -> torch.Tensor:
    """lvd.py:744-745 / :753-754: 1 - 0.5 * || p_obj - softmax(hd_lyt) ||_1 per HD pixel.
    profile_probs (B,No,Nl), hd_lyt (B,Tw,Nl,Hd,Wd) -> (B,Tw,No,Hd,Wd)."""
    sm = hd_lyt.softmax(dim=2)
    out = []
    for k in range(profile_probs.shape[1]):
        out.append(1 - (profile_probs[:, None, k, :, None, None] - sm).abs().sum(dim=2) / 2)
    return torch.stack(out, dim=2)


def context_alpha(st: WarperState, inp: torch.Tensor, grid, occ: torch.Tensor, obj_alpha: torch.Tensor,
                  bg_alpha: torch.Tensor, cls: Optional[torch.Tensor], n_win: int, use_filter: bool = True):
    """B1-B4 (lvd.py:723-766 with n_win=Tc; :617-653 with n_win=T).  Returns A in [0,1], (B,n_win,L,Hd,Wd)."""
    cfg = st.cfg
    _, sgo, _, sgb = grid
    B, T, No = sgo.shape[:3]
    s = cfg.scale_hd
    oa = ((obj_alpha + 1) / 2).unsqueeze(1).expand(-1, T, -1, -1, -1, -1)       # B T No 1 Ho Wo
    ba = ((bg_alpha + 1) / 2).unsqueeze(1).expand(-1, T, -1, -1, -1)            # B T 1 H W
    a_lo = project_layers(oa, ba, sgo, sgb)[:, :n_win, :, 0]                    # B Tw L H W
    a_hd = resize(a_lo, s)                                                      # B Tw L Hd Wd
    if use_filter:
        hd_lyt = inp[:, :n_win, 3:]
        lyt_lo = resize(inp[:, :n_win], 1 / s)[:, :, 3:]
        if cls is None or cfg.weight_cls:
            probs = class_profile(cfg, lyt_lo, a_lo[:, :, 1:], cls).softmax(dim=-1)
        else:
            probs = cls
        agree = layout_agreement(probs, hd_lyt)
        a_hd = torch.cat([a_hd[:, :, :1], a_hd[:, :, 1:] * agree], dim=2)
    return occlude(a_hd, occ[:, :n_win])


# --------------------------------------------------------------------------- a-6 flows + context warp (B5-B9)
def layer_flows(st: WarperState, grid, ctx_ts: torch.Tensor, pred_ts: torch.Tensor):
    """B5 (lvd.py:771-794 / :656-672): per-layer backward flow of every (ctx, pred) pair at HD.
    Returns Fl (B,Tc,Tp,L,2,Hd,Wd) and the low-res object support s_lo (B,Tp,No,H,W) whose
    up-sampled version thresholds into is_obj (lvd.py:788-791)."""
    cfg = st.cfg
    tgo, sgo, tgb, sgb = grid
    B, _, No = sgo.shape[:3]
    Tc, Tp = ctx_ts.shape[1], pred_ts.shape[0]
    H, W = cfg.lo_shape
    sgo_p = sgo[:, pred_ts].unsqueeze(1).expand(-1, Tc, -1, -1, -1, -1, -1).reshape(B * Tc, Tp, No, H, W, 2)
    sgb_p = sgb[:, pred_ts].unsqueeze(1).expand(-1, Tc, -1, -1, -1, -1).reshape(B * Tc, Tp, H, W, 2)
    of = take_time(tgo, ctx_ts) - tgo[:, pred_ts].unsqueeze(1)                  # B Tc Tp No Ho Wo 2
    bf = take_time(tgb, ctx_ts) - tgb[:, pred_ts].unsqueeze(1)                  # B Tc Tp H W 2
    of = of.permute(0, 1, 2, 3, 6, 4, 5).reshape(B * Tc, Tp, No, 2, *of.shape[4:6])
    bf = bf.permute(0, 1, 2, 5, 3, 4).reshape(B * Tc, Tp, 2, H, W)
    fl = project_layers(of, bf, sgo_p, sgb_p).view(B, Tc, Tp, No + 1, 2, H, W)
    ones = of.new_ones(B, Tp, No, 1, *of.shape[-2:])
    s_lo = bil0(ones.reshape(B * Tp * No, 1, *of.shape[-2:]),
                sgo[:, pred_ts].reshape(B * Tp * No, H, W, 2)).view(B, Tp, No, H, W)
    return resize(fl, cfg.scale_hd), s_lo


def grid_to_flow(st: WarperState, inp, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts, trace: Optional[dict] = None):
    """lvd.py:707-828 when cfg.restrict_to_ctx else lvd.py:602-705.
    Returns (flow, alpha_unflt|None, alpha, alpha_ctx, disocc) with the reference's shapes.
    `trace` (a dict) receives the index-class intermediates: is_obj (B,Tp,L,Hd,Wd) bool, lvd.py:788-791."""
    cfg = st.cfg
    B, T = inp.shape[:2]
    Tc, Tp = ctx_ts.shape[1], pred_ts.shape[0]
    Hd, Wd = cfg.hd_shape
    L = cfg.num_obj + 1
    if cfg.restrict_to_ctx:
        A = context_alpha(st, inp, grid, occ, obj_alpha, bg_alpha, cls, Tc, True)
    else:
        A = context_alpha(st, inp, grid, occ, obj_alpha, bg_alpha, cls, T, not cfg.no_filter)
    Fl, s_lo = layer_flows(st, grid, ctx_ts, pred_ts)
    # B6: sample the context opacities through each layer's own flow       lvd.py:796-802 / :673-679
    id_hd = pixel_grid(Hd, Wd, inp.dtype)
    g = id_hd + Fl.permute(0, 1, 2, 3, 5, 6, 4).reshape(-1, Hd, Wd, 2)
    R = bil0(take_time(A, ctx_ts).reshape(-1, 1, Hd, Wd), g).view(B, Tc, Tp, L, Hd, Wd)
    if cfg.restrict_to_ctx and not cfg.allow_ghost:
        is_obj = (resize(s_lo.unsqueeze(3), cfg.scale_hd).squeeze(3) > 0.9).to(inp.dtype)   # B Tp No Hd Wd
        is_obj = torch.cat([torch.ones_like(is_obj[:, :, :1]), is_obj], dim=2)
        if trace is not None:
            trace["is_obj"] = is_obj.detach() > 0
        R = R * is_obj.unsqueeze(1)
    disocc = R.max(dim=3, keepdim=True)[0]                                      # B7  lvd.py:803 / :680
    Actx = occlude(R, occ[:, pred_ts].unsqueeze(1))                             # B8  lvd.py:806-815
    flow = (Actx.unsqueeze(4) * Fl).sum(dim=3)                                  # B9  lvd.py:818
    alpha = A * 2 - 1
    return flow, (alpha if cfg.fast else None), alpha, Actx * 2 - 1, disocc


# --------------------------------------------------------------------------- a-7 warp frames + fuse contexts
def input_to_output(st: WarperState, inp, alpha_ctx, flow, ctx_ts, eps: float = 1e-6):
    """lvd.py:830-853.  -> output (B,Tp,C+1,Hd,Wd), raw_output (B,Tc[+1],Tp,C+L,Hd,Wd)."""
    cfg = st.cfg
    B, Tc, Tp = flow.shape[:3]
    Hd, Wd = cfg.hd_shape
    C = inp.shape[2]
    g = pixel_grid(Hd, Wd, inp.dtype) + flow.permute(0, 1, 2, 4, 5, 3).reshape(-1, Hd, Wd, 2)
    O = bil0(take_time(inp, ctx_ts).reshape(-1, C, Hd, Wd), g).view(B, Tc, Tp, C, Hd, Wd)
    s = ((alpha_ctx + 1) / 2).sum(dim=3, keepdim=True)
    if cfg.include_self and Tp == inp.shape[1]:
        s = torch.cat([s, torch.ones_like(s[:, :1])], dim=1)
        alpha_ctx = torch.cat([alpha_ctx, torch.ones_like(alpha_ctx[:, :1])], dim=1)
        O = torch.cat([O, inp.unsqueeze(1)], dim=1)
    raw = torch.cat([O, alpha_ctx], dim=3)
    wgt = F.normalize(s + eps, p=1, dim=1)
    out = (torch.cat([O, s * 2 - 1], dim=3) * wgt).sum(dim=1)
    return out, raw


def decode_output(st: WarperState, inp, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts, trace: Optional[dict] = None):
    """lvd.py:141-153.  Returns the reference's 7-tuple
    (output, flow, alpha_unflt, alpha, raw_alpha, raw_output, alpha_ctx)."""
    cfg = st.cfg
    flow, a_unflt, alpha, alpha_ctx, disocc = grid_to_flow(st, inp, grid, occ, obj_alpha, bg_alpha, cls, ctx_ts, pred_ts, trace)
    out, raw = input_to_output(st, inp, alpha_ctx, flow, ctx_ts)
    raw_alpha = out[:, :, -1:]
    if cfg.use_disocc:
        if cfg.include_self:
            disocc = torch.cat([disocc, torch.ones_like(disocc[:, :1])], dim=1)
        raw = torch.cat([raw, disocc], dim=3)
    return out[:, :, :-1], flow, a_unflt, alpha, raw_alpha, raw, alpha_ctx


def estimate_alpha_grid_occ(st: WarperState, obj_alpha_raw, obj_pose, bg_pose, occ_score):
    """lvd.py:126-135 minus the decoder: pad-mask the object alphas, expand the constant
    background alpha, build grids and the occlusion matrix."""
    om, bg = alpha_masks(st.cfg, obj_alpha_raw.dtype)
    obj_alpha = om * obj_alpha_raw + (1 - om) * (-1.0)
    bg_alpha = bg.expand(obj_alpha_raw.shape[0], -1, -1, -1)
    return compute_occ(occ_score), obj_alpha, bg_alpha, warper_forward(st, obj_pose, bg_pose)


# --------------------------------------------------------------------------- a-9 WIF fuse tail
def wif_fuse(raw_output: torch.Tensor, unet_out: torch.Tensor, ab: bool = True) -> torch.Tensor:
    """wif.py:50-54.  raw_output (B,Tc,Tp,Cin,H,W) (as produced by decode_output),
    unet_out (B,Tp,Tc,5|4,H,W) -> (B,Tp,3,H,W).  Note channel 4 of the INPUT is the gate."""
    v = raw_output.permute(0, 2, 1, 3, 4, 5)
    beta, score = unet_out[:, :, :, :3], unet_out[:, :, :, 3:4].softmax(dim=2)
    gate = (v[:, :, :, 4:5] + 5).sigmoid() if ab else 0
    return ((gate * v[:, :, :, :3] + beta) * score).sum(dim=2)


# --------------------------------------------------------------------------- f-2: loss epilogues
def gaussian_taps(kernel_size: int, sigma: float, dtype=torch.float32) -> torch.Tensor:
    """The 1-D taps torchvision's gaussian_blur builds (the reference calls it through models/synthesizer.py:1116
    `GaussianBlur(kernel_size=kernel_size, sigma=sigma)`): linspace(-half, half, k), exp(-0.5 (x / sigma)^2), normalised."""
    half = (kernel_size - 1) * 0.5
    x = torch.linspace(-half, half, steps=kernel_size, dtype=dtype)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def blur(vid: torch.Tensor, sigma: float = 3.0, kernel_size: int = 23) -> torch.Tensor:
    """models/synthesizer.py:1114-1118: every (H, W) plane of vid (..., C, H, W) convolved with the outer product of the
    Gaussian taps after reflect padding (what torchvision's GaussianBlur does; restated with plain torch ops so that the
    oracle does not need torchvision and has an fp64 twin)."""
    k1 = gaussian_taps(kernel_size, sigma, vid.dtype)
    k2 = torch.mm(k1[:, None], k1[None, :])
    H, W = vid.shape[-2:]
    img = vid.reshape(-1, 1, H, W)
    r = kernel_size // 2
    img = F.pad(img, [r, r, r, r], mode="reflect")
    return F.conv2d(img, k2[None, None]).view_as(vid)


def layer_entropy(alpha: torch.Tensor):
    """models/synthesizer.py:886-889 and :933 -> (entropy (B,T,1,H,W), fg_mask (B,T,1,H,W)) of alpha (B,T,L,H,W)."""
    entropy = (alpha + 1) / 2
    entropy = F.normalize(entropy + 1e-6, p=1, dim=2)
    entropy = -torch.sum(torch.mul(entropy, torch.log(entropy + 1e-6)), dim=2, keepdim=True) / 0.37
    fg_mask = ((alpha[:, :, 1:] + 1) / 2).sum(dim=2, keepdim=True)
    return entropy, fg_mask


def pose_distances(mov_obj_mask: torch.Tensor, fg_mask: torch.Tensor, obj_pose: torch.Tensor, grid: torch.Tensor, obj_shape, eps: float):
    """models/synthesizer.py:965-979 up to (not including) the two `.mean()`s: (cell_min, center_min), each (B, T, H, W).
    mov_obj_mask, fg_mask (B, T, 1, H, W); obj_pose (B, T, No, ho*wo, 2); grid (1, H, W, 2)."""
    num_obj = obj_pose.shape[2]
    obj_grid = obj_pose.view(*obj_pose.shape[:3], *obj_shape, 2)                                     # :966
    obj_cell = (obj_grid[:, :, :, 1:, 1:] + obj_grid[:, :, :, 1:, :-1] + obj_grid[:, :, :, :-1, 1:] + obj_grid[:, :, :, :-1, :-1]) / 4   # :967
    obj_center = obj_grid.view(*obj_pose.shape[:3], -1, 2).mean(dim=3)                               # :968
    obj_cell_dis = (grid ** 2).sum(dim=-1).view(1, -1) + (obj_cell ** 2).sum(dim=-1).view(-1, 1) - 2 * obj_cell.reshape(-1, 2) @ grid.view(-1, 2).t()   # :969
    obj_cell_dis = obj_cell_dis.view(*obj_grid.shape[:2], num_obj, -1, *grid.shape[1:3]).sum(dim=3)  # :970-971
    obj_center_dis = (grid ** 2).sum(dim=-1).view(1, -1) + (obj_center ** 2).sum(dim=-1).view(-1, 1) - 2 * obj_center.view(-1, 2) @ grid.view(-1, 2).t()   # :972
    obj_center_dis = obj_center_dis.view(*obj_grid.shape[:2], num_obj, *grid.shape[1:3])             # :973
    cell = ((mov_obj_mask + eps) * (1 - fg_mask) * obj_cell_dis).min(dim=2)[0]                       # :976
    center = (mov_obj_mask * obj_center_dis).min(dim=2)[0]                                           # :978
    return cell, center


def obj_flow(rec_output_alpha: torch.Tensor, real_flow: torch.Tensor) -> torch.Tensor:
    """models/synthesizer.py:864-868 -> the scalar `obj_flow`.  rec_output_alpha (B, T, No+1, H, W), real_flow (B, T, 2, H, W)."""
    a = (rec_output_alpha[:, :, 1:] + 1) / 2 + 1e-6                                                   # :865
    sum_a = a.sum(dim=3, keepdim=True).sum(dim=4, keepdim=True)                                       # :866
    mean_flow = (real_flow.unsqueeze(2) * a.unsqueeze(3)).sum(dim=4, keepdim=True).sum(dim=5, keepdim=True) / sum_a.unsqueeze(3)   # :867
    return (a * (real_flow.unsqueeze(2) - mean_flow).abs().sum(dim=3)).mean()                         # :868


# --------------------------------------------------------------------------- synthetic inputs
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
    base = 0.25 * pixel_grid(*cfg.obj_shape).view(1, 1, 1, Lo, 2) * torch.tensor([1.0, cfg.aspect_ratio])
    drift = torch.linspace(0, 0.1, T).view(1, T, 1, 1, 1)
    obj_pose = base + centre.view(1, 1, No, 1, 2) + drift + 0.01 * torch.randn(B, T, No, Lo, 2, generator=g)
    bg_pose = 1.2 * pixel_grid(*cfg.latent_shape).view(1, 1, 1, L, 2) + 0.005 * torch.randn(B, T, 1, L, 2, generator=g)
    occ_score = torch.randn(B, T, No, generator=g)
    cls = torch.randn(B, No, Nl, generator=g).softmax(dim=-1)
    Tp = T - Tc
    ctx_ts = torch.arange(Tc).view(1, Tc, 1).expand(B, Tc, Tp)
    pred_ts = torch.arange(Tc, T)
    d = dict(input=inp, obj_alpha_raw=obj_alpha, obj_pose=obj_pose, bg_pose=bg_pose, occ_score=occ_score, cls=cls)
    d = {k: v.to(dtype) for k, v in d.items()}
    d.update(ctx_ts=ctx_ts, pred_ts=pred_ts)
    return d
