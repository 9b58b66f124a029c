"""f-2 `cell_dis` / `center_dis` kernels through the sm_100a library.  The same check runs on the host emulation of the kernels in
the CPU suite (tests/test_emu_parity.py::test_pose_distances).  This file sorts last on purpose: the kernels were added after the
round's GPU budget was spent, so their first run on a B200 is the driver's."""
import pytest
import torch

from tests import parity


@pytest.mark.gpu
def test_pose_distances_gpu():
    assert torch.cuda.is_available()
    parity.check_pose_distances(torch.device("cuda:0"))
