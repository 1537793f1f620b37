import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(REPO, "tests", "golden")


@pytest.fixture(autouse=True)
def _emulated_gpu(request, monkeypatch):
    """FSNET_EMULATED_GPU=1: run `-m gpu` tests that stay inside this process on the CPU -- the kernels under the SIMT emulator
    (tests/host_emulation), `.cuda()` a no-op.  A dry run for GPU tests that have not seen a B200 yet: it proves their logic and the
    non-tensor-core kernels they launch, nothing about the tcgen05 convolutions or about speed.  Tests that spawn the training
    scripts or capture CUDA graphs can not run this way."""
    if os.environ.get("FSNET_EMULATED_GPU") != "1" or request.node.get_closest_marker("gpu") is None:
        yield
        return
    import torch
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from host_emulation import fixture
    fixture.install(monkeypatch)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "cpu", lambda self, *a, **k: self)
    plain_to = torch.Tensor.to

    def to(self, *args, **kwargs):
        def host(d):
            return "cpu" if (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda") else d
        if "device" in kwargs:
            kwargs["device"] = host(kwargs["device"])
        return plain_to(self, *[host(a) for a in args], **kwargs)

    monkeypatch.setattr(torch.Tensor, "to", to)
    yield
