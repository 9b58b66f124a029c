"""A/B of two builds of the SAME kernel sources on the GPU, output by output and gradient by gradient (the
lanes forms of the layer / context-alpha kernels against the generic per-pixel forms the host emulation compiles).

    python -m waldo_b200.build --force -DWB_NO_LANES -o$PWD/scratch/libwaldo_b200_nolanes.so
    python profiles/tools/ab_forms.py scratch/libwaldo_b200_nolanes.so [shape ...]

Each side runs in its own process (the library is chosen at import through WALDO_B200_LIB), decodes the full-shape cases of
tests/parity.py forward and backward in deterministic mode (64-bit fixed-point accumulation: the result does not depend on the
order of the reductions, so every difference is a difference of the addends) and dumps its tensors; the parent compares."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child(shape, out_path):
    import torch
    import waldo_b200 as wb
    from tests import parity
    from tests.parity import wo
    kw, B, T, Tc = parity.FULL_SHAPES[shape]
    cfg = wo.PathConfig(**kw)
    dev = torch.device("cuda:0")
    d = wo.synth_inputs(cfg, B, T, Tc, seed=0)
    opt = parity.make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    om, bg = wb.alpha_masks(opt)
    om, bg = (om.to(dev) if torch.is_tensor(om) else om), bg.to(dev)
    wb.set_deterministic(True)
    lv = {k: d[k].clone().to(dev).requires_grad_(True) for k in ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls")}
    occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, lv["obj_alpha_raw"], om, bg, lv["obj_pose"], lv["bg_pose"], lv["occ_score"])
    out = wb.decode_output(warper, lv["input"], grid, occ, oa, ba, lv["cls"], d["ctx_ts"].to(dev), d["pred_ts"].to(dev), cfg.restrict_to_ctx)
    gen = torch.Generator().manual_seed(5)
    loss = 0
    for o in out:
        if o is not None and o.requires_grad:
            loss = loss + (o * torch.randn(o.shape, generator=gen).to(dev)).sum()
    loss.backward()
    blob = {f"out_{n}": o.detach().cpu() for n, o in zip(parity.OUT_NAMES, out) if o is not None}
    blob.update({f"d_{k}": v.grad.cpu() for k, v in lv.items()})
    blob["lib"] = wb._lib.LIB_PATH
    torch.save(blob, out_path)


def main():
    import torch
    alt = os.path.abspath(sys.argv[1])
    shapes = sys.argv[2:] or ["city_512x1024", "kitti_256x832"]
    report = {}
    with tempfile.TemporaryDirectory() as tmp:
        for shape in shapes:
            blobs = []
            for tag, lib in (("default", None), ("alt", alt)):
                env = dict(os.environ)
                env.pop("WALDO_B200_LIB", None)
                if lib:
                    env["WALDO_B200_LIB"] = lib
                p = os.path.join(tmp, f"{shape}_{tag}.pt")
                subprocess.run([sys.executable, os.path.abspath(__file__), "--child", shape, p], check=True, env=env, cwd=ROOT)
                blobs.append(torch.load(p))
            a, b = blobs
            assert a["lib"] != b["lib"], (a["lib"], b["lib"])
            rep = {"libs": [os.path.relpath(a["lib"], ROOT), os.path.relpath(b["lib"], ROOT)]}
            for k in a:
                if k == "lib":
                    continue
                x, y = a[k].double(), b[k].double()
                sc = max(float(x.abs().max()), 1e-30)
                rep[k] = {"bit_identical": bool(torch.equal(a[k], b[k])), "max_abs": float((x - y).abs().max()),
                          "max_abs_over_scale": float((x - y).abs().max()) / sc, "differing": int((a[k] != b[k]).sum()), "numel": a[k].numel()}
                if rep[k]["max_abs_over_scale"] > 2e-5 and a[k].dim() >= 3:   # where: per-channel counts (dim -3) and the five worst elements
                    dlt = (x - y).abs()
                    per_ch = (dlt > 1e-5 * sc).sum(dim=tuple(i for i in range(dlt.dim()) if i != dlt.dim() - 3))
                    rep[k]["per_channel_over_1e-5"] = per_ch.tolist()
                    top = torch.topk(dlt.flatten(), 5)
                    rep[k]["worst"] = [{"index": [int(v) for v in torch.unravel_index(i, dlt.shape)], "default": float(a[k].flatten()[i]),
                                        "alt": float(b[k].flatten()[i])} for i in top.indices]
            report[shape] = rep
    print(json.dumps(report, indent=1))
    # Bound: well inside the reference arithmetic's own fp32-vs-fp64 floor at these shapes (profiles/r2/parity_*.json: raw_output
    # 7.4e-3 absolute = 1.5e-3 of its range, gradients 2e-2 .. 1.6e-1 of theirs); a one-ulp difference of a sampling coordinate near
    # 1.0 moves a bilinear tap by 6e-5 pixel = up to 6e-4 on inputs that differ by 10 between neighbouring pixels.
    bad = [(s, k) for s, r in report.items() for k, v in r.items() if k != "libs" and v["max_abs_over_scale"] > 5e-4]
    if bad:
        raise SystemExit(f"forms disagree beyond rounding: {bad}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3])
    else:
        main()
