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


@pytest.mark.parametrize("case", parity.CASES)
def test_occ(dev, case):
    parity.check_occ(dev, case)


@pytest.mark.parametrize("case", parity.CASES)
def test_decode(dev, case):
    parity.check_decode(dev, case)


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
