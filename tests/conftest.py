import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def load_autophase_cases():
    raw = load_golden("autophase_cases")
    n = int(raw["count"])
    cases = []
    for i in range(n):
        pre = f"{i:02d}__"
        c = {k[len(pre):]: v for k, v in raw.items() if k.startswith(pre)}
        c["name"] = str(c["name"])
        c["variant"] = eval(str(c["variant"]))  # list of (key, value) pairs written by make_golden.py
        c["kwargs"] = dict(c["variant"])
        for k in ("p0", "p1", "pivot", "lb"):
            c[k] = float(c[k])
        c["zf"] = int(c["zf"])
        cases.append(c)
    return cases


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="session")
def golden():
    return load_golden
