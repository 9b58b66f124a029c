"""f-2 `obj_flow` kernels through the sm_100a library.  The same check runs on the host emulation of the kernels in the CPU suite
(tests/test_emu_parity.py::test_obj_flow).  This file sorts last on purpose: the kernels were added after the round's GPU budget was
spent, so their first run on a B200 is the driver's (every other GPU test of the suite has been run on a B200)."""
import pytest
import torch

from tests import parity


@pytest.mark.gpu
def test_obj_flow_gpu():
    assert torch.cuda.is_available()
    parity.check_obj_flow(torch.device("cuda:0"))


@pytest.mark.gpu
def test_loss_epilogue_shapes_gpu():
    assert torch.cuda.is_available()
    parity.check_loss_epilogue_shapes(torch.device("cuda:0"))
