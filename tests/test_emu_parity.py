"""CPU suite: the kernels' logic, compiled for the host (tests/emu, one thread per CTA), against the oracle and the
reference fixtures.  This checks index arithmetic and formulas without a GPU; tests/test_gpu_parity.py runs the very same
checks through the real sm_100a library."""
import pytest
import torch

from tests import parity
from tests.emu.harness import emulated

DEV = torch.device("cpu")


@pytest.fixture(autouse=True)
def _emu():
    with emulated():
        yield


@pytest.mark.parametrize("case", parity.CASES)
def test_tps(case):
    parity.check_tps(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_inverse_warp(case):
    parity.check_inverse_warp(DEV, case)


@pytest.mark.parametrize("knob", [("WALDO_INV_CAP", "64"), ("WALDO_INV_CAP", "1500"), ("WALDO_INV_MARGIN", "-6"), ("WALDO_INV_UNFUSED", "1")])
@pytest.mark.parametrize("case", ["city_x4", "city_real", "kitti_real"])
def test_inverse_warp_paths(case, knob, monkeypatch):
    """The fused one-CTA-per-item object kernel with its work area in global memory (box over the cell budget: all / some
    items), after a sample landed outside the predicted box (margin shrunk artificially), and the phase-per-kernel path."""
    monkeypatch.setenv(*knob)
    parity.check_inverse_warp(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_occ(case):
    parity.check_occ(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode(case):
    parity.check_decode(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode_deterministic(case):
    parity.check_decode_deterministic(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode_bf16_storage(case):
    parity.check_decode_bf16(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_end_to_end(case):
    parity.check_end_to_end(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_wif(case):
    parity.check_wif(DEV, case)


def test_kats():
    parity.check_kats(DEV)


def test_conv3x3():
    parity.check_conv3x3(DEV)


def test_blur():
    parity.check_blur(DEV)


def test_layer_entropy():
    parity.check_layer_entropy(DEV)


def test_pose_distances():
    parity.check_pose_distances(DEV)


def test_obj_flow():
    parity.check_obj_flow(DEV)


def test_loss_epilogue_shapes():
    parity.check_loss_epilogue_shapes(DEV)


def test_pack_input():
    parity.check_pack_input(DEV)


def test_frames_to_u8():
    parity.check_frames_to_u8(DEV)


@pytest.mark.parametrize("case", ["city_x4", "kitti_x2", "train_lo"])
def test_field_warps(case):
    parity.check_field_warps(DEV, case)


@pytest.mark.parametrize("case", list(parity.SYNTH_CASES))
def test_decode_oracle_direct(case):
    parity.check_decode_synth(DEV, case)


def test_small_chain_deterministic():
    cfg, (B, T, Tc), _ = parity.load_case("train_lo")
    parity.check_chain_deterministic(DEV, cfg, B, T, Tc, seed=5)


def test_deterministic_mode_rejects_targets_outside_the_arena(monkeypatch):
    """C ABI contract of the det_* fields: every scatter target must lie inside [det_base, det_base + det_n)."""
    import waldo_b200 as wb
    from waldo_b200 import functional as F
    orig = F._scatter_targets

    def broken(refs, det, fill=True):
        out, arena, shadow = orig(refs, det, fill)
        if det:   # d_input allocated on its own instead of as a view of the arena
            out[0] = torch.zeros_like(refs[0]) if refs[0] is not None else None
        return out, arena, shadow
    monkeypatch.setattr(F, "_scatter_targets", broken)
    cfg, _, z = parity.load_case("train_lo")
    wb.set_deterministic(True)
    try:
        with pytest.raises(RuntimeError, match="outside the det arena"):
            parity.kernel_decode(DEV, cfg, z)
    finally:
        wb.set_deterministic(False)


def test_deterministic_flag_follows_torch():
    import waldo_b200 as wb
    assert not wb.is_deterministic()
    torch.use_deterministic_algorithms(True)
    try:
        assert wb.is_deterministic()
    finally:
        torch.use_deterministic_algorithms(False)
    assert not wb.is_deterministic()


def test_scatter_targets_are_aligned_views_of_one_arena():
    from waldo_b200 import functional as F
    refs = [torch.empty(3, 5), None, torch.empty(7), torch.empty(2, 2, 2)]
    out, arena, shadow = F._scatter_targets(refs, True)
    assert out[1] is None and arena.dtype == torch.float32 and shadow.dtype == torch.int64 and arena.numel() == shadow.numel()
    base = arena.data_ptr()
    for r, t in zip(refs, out):
        if r is None:
            continue
        assert t.shape == r.shape and float(t.abs().sum()) == 0.0
        off = t.data_ptr() - base
        assert 0 <= off and off + 4 * t.numel() <= 4 * arena.numel() and off % 16 == 0
    plain, a2, s2 = F._scatter_targets(refs, False)
    assert a2 is None and s2 is None and plain[1] is None and plain[0].shape == refs[0].shape


@pytest.mark.parametrize("case", list(parity.SYNTH_CASES_EMU_ONLY))
def test_decode_oracle_direct_emu_only(case):
    parity.check_decode_synth(DEV, case)


def test_nonrigid_workload_shape_runs_the_whole_chain():
    """bench.py --workload nonrigid_train (256x256, scale 2, 8x8 background control points) at B=1: the chain runs, every
    gradient is finite, deterministic mode agrees with the default mode."""
    import bench
    cfg, spec = bench.workload_cfg("nonrigid_train")
    assert cfg.hd_shape == (256, 256) and cfg.lo_shape == (128, 128) and spec["backward"]
    parity.check_chain_deterministic(DEV, cfg, 1, spec["T"], spec["Tc"], seed=2)


def test_no_kernel_reads_uninitialised_scratch():
    """torch.use_deterministic_algorithms(True) fills every torch.empty buffer with NaN: the whole chain (control points ->
    loss -> every leaf) must still give finite results, bit-identical to the ones obtained with ordinary allocations --
    i.e. no kernel depends on the previous contents of an output / scratch buffer."""
    import waldo_b200 as wb
    cfg, (B, T, Tc), _ = parity.load_case("train_lo")
    opt = parity.make_opt(cfg)
    warper = wb.Warper(opt).to(DEV)
    d = parity.wo.synth_inputs(cfg, B, T, Tc, seed=9)
    om, bg = wb.alpha_masks(opt)
    wb.set_deterministic(True)
    try:
        g1 = parity._chain_grads(DEV, cfg, d, warper, om, bg)
    finally:
        wb.set_deterministic(False)
    torch.use_deterministic_algorithms(True)
    try:
        assert torch.utils.deterministic.fill_uninitialized_memory
        g2 = parity._chain_grads(DEV, cfg, d, warper, om, bg)
    finally:
        torch.use_deterministic_algorithms(False)
    for k in g1:
        assert bool(torch.isfinite(g2[k]).all()), k
        assert torch.equal(g1[k], g2[k]), k
