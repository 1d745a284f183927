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


def pytest_terminal_summary(terminalreporter):
    """Measured error of every parity comparison next to its bound (see _util.RECORDS)."""
    import json
    from _util import RECORDS
    if not RECORDS:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_errors.json"), "w") as f:
            json.dump(RECORDS, f, indent=0)
    worst = {}
    for r in RECORDS:                                   # per (test function, quantity): the largest share of the bound
        key = (r["test"].split("::")[-1].split("[")[0], r["what"].split(" ")[0], r["kind"])
        if key not in worst or r["used"] > worst[key]["used"]:
            worst[key] = r
    tr = terminalreporter
    tr.write_line("")
    tr.write_line(f"parity: {len(RECORDS)} comparisons; largest measured error per (test, quantity, tolerance):")
    for (test, what, kind), r in sorted(worst.items()):
        tr.write_line(f"  {test:58s} {what:14s} err {r['max_err']:.3e}  ref {r['max_ref']:.3e}  "
                      f"bound {r['bound']:.3e} ({100 * r['used']:.0f}% used)  [{kind}]")
