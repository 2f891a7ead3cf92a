import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))


_CACHE = {}


def load_case(name):
    """Mesh + model of a committed fixture, and the oracle's assembly of it (cached per session)."""
    if name in _CACHE:
        return _CACHE[name]
    from oracle import fem
    g = json.load(open(os.path.join(GOLDEN_DIR, "golden.json")))[name]
    z = np.load(os.path.join(GOLDEN_DIR, "inputs_%s.npz" % name))
    mesh = dict(ntet=int(z["ele"].shape[0]), nvert=int(z["node"].shape[0]), ele=z["ele"].astype(np.int64),
                neigh=z["neigh"].astype(np.int64), node=z["node"])
    model = dict(vp=z["vp"], vs=z["vs"], rho=z["rho"], g0=z["g0"] if "g0" in z.files else None)
    mats, topo, num, geo = fem.assemble(mesh, model, g["porder"], g["job"])
    out = dict(g=g, mesh=mesh, model=model, mats=mats, topo=topo, num=num, geo=geo)
    _CACHE[name] = out
    return out


@pytest.fixture(scope="session")
def case_loader():
    return load_case


def to_coomat(mats):
    """oracle CSR dicts -> normalmodes_b200.matvec.COOmat (single rank)."""
    from normalmodes_b200.matvec import COOmat
    out = {}
    for k, m in mats.items():
        nr, nc = m["shape"]
        out[k] = COOmat([0, nr], m["ia"], m["ja"], m["a"], coldist=[0, nc])
    return out
