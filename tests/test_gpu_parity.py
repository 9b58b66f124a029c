"""GPU suite (`-m gpu`): the real sm_100a library, called through the C ABI, against the oracle and the committed
reference outputs -- the parity tests proper.  Same checks as tests/test_emu_parity.py, plus full-size properties."""
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from waldo_b200 import _lib
    lib = _lib.load()
    assert lib.waldo_has_device_code() == 1
    return torch.device("cuda:0")


@pytest.mark.parametrize("case", parity.CASES)
def test_tps(dev, case):
    parity.check_tps(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_inverse_warp(dev, case):
    parity.check_inverse_warp(dev, case)


@pytest.mark.parametrize("knob", [("WALDO_INV_CAP", "64"), ("WALDO_INV_CAP", "1500"), ("WALDO_INV_MARGIN", "-6"), ("WALDO_INV_UNFUSED", "1")])
@pytest.mark.parametrize("case", ["city_x4", "city_real", "kitti_real"])
def test_inverse_warp_paths(dev, case, knob, monkeypatch):
    """k_inv_fused with its work area in global memory (box over the cell budget), after a sample landed outside the
    predicted box (margin shrunk artificially), and the phase-per-kernel path: all bit-identical index maps."""
    monkeypatch.setenv(*knob)
    parity.check_inverse_warp(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_occ(dev, case):
    parity.check_occ(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode(dev, case):
    parity.check_decode(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode_bf16_storage(dev, case):
    parity.check_decode_bf16(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode_deterministic(dev, case):
    parity.check_decode_deterministic(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_end_to_end(dev, case):
    parity.check_end_to_end(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_wif(dev, case):
    parity.check_wif(dev, case)


def test_kats(dev):
    parity.check_kats(dev)


def test_cpu_tensor_is_refused(dev):
    """No CPU path: host tensors must fail loudly."""
    import waldo_b200 as wb
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        wb.compute_occ(torch.zeros(1, 2, 3))


def test_full_size_properties(dev):
    """BASELINE config shape (Cityscapes 512x1024, 16 objects, 20 classes), B=1, 4 contexts -> 1 future frame:
    size-independent properties of the fused path."""
    parity.check_full_size(dev)


def test_full_size_deterministic_gradients(dev):
    """BASELINE shape (Cityscapes 512x1024, 17 layers, 23 channels), B=1: the whole backward in deterministic mode is
    bit-identical from run to run and agrees with the default (float-reduction) mode up to rounding."""
    parity.check_chain_deterministic(dev)


def test_small_chain_deterministic_gradients(dev):
    cfg, (B, T, Tc), _ = parity.load_case("train_lo")
    parity.check_chain_deterministic(dev, cfg, B, T, Tc, seed=5)


def test_deterministic_mode_survives_large_intermediate_gradients(dev):
    parity.check_deterministic_large_gradients(dev)


def test_conv3x3(dev):
    parity.check_conv3x3(dev)


def test_blur(dev):
    parity.check_blur(dev)


def test_layer_entropy(dev):
    parity.check_layer_entropy(dev)


def test_pose_distances(dev):
    parity.check_pose_distances(dev)


def test_pack_input(dev):
    parity.check_pack_input(dev)
    parity.check_pack_input(dev, B=1, T=2, Hd=256, Wd=832, num_lyt=19)


def test_frames_to_u8(dev):
    parity.check_frames_to_u8(dev)


def test_device_prefetcher(dev):
    """Batches arrive on the device intact and in order while the previous one is still being consumed; 8-bit RGB +
    labels are expanded to the fp32 `input` exactly as the host formulas do."""
    import waldo_b200 as wb
    g = torch.Generator().manual_seed(3)
    host = []
    for i in range(5):
        rgb = torch.randint(0, 256, (2, 3, 3, 32, 64), generator=g, dtype=torch.uint8).pin_memory()
        lab = torch.randint(0, 20, (2, 3, 32, 64), generator=g, dtype=torch.uint8).pin_memory()
        pose = torch.randn(2, 3, 4, generator=g).pin_memory()
        host.append({"rgb": rgb, "label": lab, "pose": pose})
    seen = 0
    sink = torch.zeros(4096, 4096, device=dev)
    for i, batch in enumerate(wb.DevicePrefetcher(host, dev, num_lyt=20)):
        sink = sink @ sink * 0 + 1          # keep the consumer stream busy while the next copy runs
        want = parity.reference_pack(host[i]["rgb"], host[i]["label"], 20)
        assert torch.equal(batch["input"].cpu(), want), i
        assert torch.equal(batch["pose"].cpu(), host[i]["pose"]), i
        seen += 1
    assert seen == 5


def test_graphed_decode_matches_eager(dev):
    """The CUDA-graph replay of the inference chain returns bit-identical tensors to the eager calls, also after the
    input buffers have been refilled in place."""
    import waldo_b200 as wb
    cfg, (B, T, Tc), z = parity.load_case("kitti_x2")
    opt = parity.make_opt(cfg)
    warper = wb.Warper(opt).to(dev)
    om, bg = wb.alpha_masks(opt)
    om, bg = (om.to(dev) if torch.is_tensor(om) else om), bg.to(dev)
    d = parity.wo.synth_inputs(cfg, B, T, Tc, seed=3)
    keys = ("input", "obj_alpha_raw", "obj_pose", "bg_pose", "occ_score", "cls", "ctx_ts", "pred_ts")
    bufs = {k: d[k].contiguous().to(dev) for k in keys}
    gd = wb.GraphedDecode(warper, om, bg, cfg.restrict_to_ctx)
    for seed in (3, 4):
        d = parity.wo.synth_inputs(cfg, B, T, Tc, seed=seed)
        for k in keys:
            bufs[k].copy_(d[k])
        got = gd(*[bufs[k] for k in keys])
        with torch.no_grad():
            occ, oa, ba, grid = wb.estimate_alpha_grid_occ(warper, bufs["obj_alpha_raw"], om, bg, bufs["obj_pose"], bufs["bg_pose"], bufs["occ_score"])
            want = wb.decode_output(warper, bufs["input"], grid, occ, oa, ba, bufs["cls"], bufs["ctx_ts"], bufs["pred_ts"], cfg.restrict_to_ctx)
        for a, b in zip(got, want):
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.equal(a, b)
    assert len(gd.cache) == 1


@pytest.mark.parametrize("case", ["city_x4", "kitti_x2", "train_lo"])
def test_field_warps(dev, case):
    parity.check_field_warps(dev, case)


@pytest.mark.parametrize("case", list(parity.SYNTH_CASES))
def test_decode_oracle_direct(dev, case):
    parity.check_decode_synth(dev, case)
