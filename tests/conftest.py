import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that stalls must fail, not hang the box: every `gpu` test gets a 300 s limit (pytest-timeout, if
    installed; the whole CUDA suite normally takes ~20 s)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None and item.get_closest_marker("timeout") is None:
            # method 'thread': a test stuck inside a CUDA call never returns to Python, so a signal would not fire
            item.add_marker(pytest.mark.timeout(300, method="thread"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    import json
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


GOLDEN_CASES = [
    "s8_nb1_240_refinit", "s8_nb3_480_trained", "s8_nb3_240_b2_trained", "s8_nb2_224_trained",
    "s8_nb1_64_trained", "s8_nb3_480_refinit", "b8_nb4_240_trained", "b8_nb4_240_refinit",
    "s8_nb1_240_linear5_refinit",
    # BASELINE.json configs[3] / configs[4] at their full per-frame shape (two frames each)
    "s8_nb3_960_b2_refinit", "b8_nb4_480_b2_refinit",
]
