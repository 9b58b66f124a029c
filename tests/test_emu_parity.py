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
def test_end_to_end(case):
    parity.check_end_to_end(DEV, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_wif(case):
    parity.check_wif(DEV, case)


def test_kats():
    parity.check_kats(DEV)


def test_pack_input():
    parity.check_pack_input(DEV)


@pytest.mark.parametrize("case", ["city_x4", "kitti_x2", "train_lo"])
def test_field_warps(case):
    parity.check_field_warps(DEV, case)


@pytest.mark.parametrize("case", list(parity.SYNTH_CASES))
def test_decode_oracle_direct(case):
    parity.check_decode_synth(DEV, case)


def test_small_chain_deterministic():
    cfg, (B, T, Tc), _ = parity.load_case("train_lo")
    parity.check_chain_deterministic(DEV, cfg, B, T, Tc, seed=5)
