import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """The ledger of achieved CUDA-vs-reference / CUDA-vs-oracle errors (tests/parity.py::record)."""
    import json
    import parity as P
    if not P.ACHIEVED:
        return
    tr = terminalreporter
    tr.write_sep("=", "parity ledger: achieved error | bar (tests/parity.py)")
    worst = {}
    for r in P.ACHIEVED:
        k = (r["case"], r["what"])
        if k not in worst or r["achieved"] > worst[k]["achieved"]:
            worst[k] = r
    for (case, what), r in sorted(worst.items()):
        bar = "-" if r["bar"] is None else f"{r['bar']:.1e}"
        tr.write_line(f"{case:34s} {what:34s} {r['achieved']:.3e} | {bar} {r['note']}")
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_achieved.json"), "w") as f:
            json.dump(sorted(worst.values(), key=lambda r: (r["case"], r["what"])), f, indent=1)
    except OSError:
        pass
