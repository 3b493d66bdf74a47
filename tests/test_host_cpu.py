"""CPU-only tests of the C++ host engine (the reference-API mirror): builder / factory / compiled-layers
logic via its device-free self-test, and the weight-file layouts of the sample networks against the
reference's hard-coded tables (through the oracle's independent restatement)."""
import pytest

import fyn_oracle as fo
from fyusenet_b200 import hostapi


def test_host_selftest():
    failures, report = hostapi.selftest()
    assert failures == 0, report
    assert report.count("ok ") >= 25


@pytest.mark.parametrize("ksize", [3, 9])
def test_stylenet_weight_offsets(ksize):
    """stylenet9x9.cpp:41-56 / stylenet3x3.cpp:41-50; layer numbers follow the enum of stylenet9x9.h:39-57."""
    net = hostapi.StyleNet(ksize, 64, 48, device=-1)
    offs = fo.stylenet_offsets(ksize)
    nres = 5 if ksize == 9 else 2
    names = ["conv1", "conv2", "conv3"] + [f"res{r}_{i}" for r in range(1, nres + 1) for i in (1, 2)] + ["deconv1", "deconv2", "deconv3"]
    assert net.weight_floats == offs["_total"] == (169059 if ksize == 9 else 77235)
    for i, name in enumerate(names):
        assert net.weight_offset(1 + i) == offs[name], name
    assert net.weight_offset(0) == -1 and net.weight_offset(1 + len(names)) == -1
    net.destroy()


def test_resnet50_weight_offsets():
    """resnet50.cpp:539-677 (every layer), via the oracle table that is itself checked against the reference bytes."""
    net = hostapi.ResNet50(device=-1)
    offs = fo.resnet50_offsets()
    assert net.weight_floats * 4 == 102304184
    for no in range(0, 74):
        want = offs.get(no, -1)
        assert net.weight_offset(no) == want, no
    net.destroy()


def test_setup_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = hostapi.StyleNet(3, 64, 48, device=-1)
    net.load_weights(fo.stylenet_synthetic_weights(3))
    with pytest.raises(hostapi.HostError, match="no CUDA device|no CPU fallback"):
        net.setup()
    with pytest.raises(hostapi.HostError, match="expected"):
        net.load_weights(fo.stylenet_synthetic_weights(9))
    with pytest.raises(hostapi.HostError, match="multiples of 4"):
        hostapi.StyleNet(3, 62, 48, device=-1)
