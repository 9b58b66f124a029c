"""The product-side workload helper (waldo_b200/workloads.py: what bench.py uses) against the oracle's own definitions."""
import torch

from tests.parity import wo
from waldo_b200 import workloads as wl


def test_synthetic_inputs_and_options_match_the_oracle():
    for kw in (dict(dim=16, load_dim=32), dict(dim=16, load_dim=64, aspect_ratio=3.25, latent_shape=(8, 26), num_lyt=19)):
        a, b = wo.PathConfig(**kw), wl.PathConfig(**kw)
        assert a.lo_shape == b.lo_shape and a.hd_shape == b.hd_shape and a.obj_hw == b.obj_hw and a.scale_hd == b.scale_hd
        x, y = wo.synth_inputs(a, 2, 6, 4, seed=7), wl.synth_inputs(b, 2, 6, 4, seed=7)
        assert x.keys() == y.keys() and all(torch.equal(x[k], y[k]) for k in x)
        x, y = wo.synth_inputs(a, 1, 5, 4, seed=1, smooth=True), wl.synth_inputs(b, 1, 5, 4, seed=1, smooth=True)
        assert all(torch.equal(x[k], y[k]) for k in x)


def test_benchmark_workloads_are_the_baseline_configs():
    cfg, spec = wl.workload("city_train")
    assert cfg.hd_shape == (512, 1024) and cfg.num_obj == 16 and cfg.latent_shape == (8, 16) and (spec["B"], spec["T"], spec["Tc"]) == (8, 5, 4)
    cfg, spec = wl.workload("kitti_rollout")
    assert cfg.hd_shape == (256, 832) and cfg.num_lyt == 19 and cfg.latent_shape == (8, 26) and (spec["T"], spec["Tc"]) == (9, 4)
    cfg, spec = wl.workload("nonrigid_train")
    assert cfg.hd_shape == (256, 256) and spec["backward"]
    f, b = wl.alg_bytes(wl.PathConfig(), 8, 4, 1, True)
    assert abs(f - 8.46e9) < 0.02e9 and abs(b - 8.59e9) < 0.02e9      # SURVEY.md 8d worked example
