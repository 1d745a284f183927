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


@pytest.fixture(scope="session")
def toy_kg():
    from graph_hypernetwork_forge import ToyKnowledgeGraph
    return ToyKnowledgeGraph(feat_dim=16)
