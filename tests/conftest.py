"""Shared fixtures.  `-m "not gpu"` runs here (CPU box); `-m gpu` runs on a B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "graph-hypernetwork-forge_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "grad: runs with autograd enabled (every other test runs under torch.no_grad())")


@pytest.fixture(autouse=True)
def _no_grad_unless_marked(request):
    """Forward parity tests compare plain tensors: they run under torch.no_grad() (with gradients enabled the
    drop-in records an autograd graph, as the reference does).  Tests of the gradients are marked `grad`."""
    import torch
    if request.node.get_closest_marker("grad"):
        yield
    else:
        with torch.no_grad():
            yield


@pytest.fixture(scope="session")
def toy_kg():
    from graph_hypernetwork_forge import ToyKnowledgeGraph
    return ToyKnowledgeGraph(feat_dim=16)
