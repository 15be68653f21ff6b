import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden.npz")))


@pytest.fixture(scope="session")
def cora():
    z = np.load(os.path.join(GOLDEN_DIR, "cora.npz"))
    shape = tuple(int(v) for v in z["feat_shape"])
    feats = np.zeros(shape, np.float32)
    feats[z["feat_rows"].astype(np.int64), z["feat_cols"].astype(np.int64)] = z["feat_vals"]
    return dict(rowptr64=z["rowptr64"], rowptr=z["rowptr64"].astype(np.uint32), colidx=z["colidx"], labels=z["labels"],
                split=z["split"], feats=feats, ncls=int(z["ncls"]), nv=shape[0], feat_len=shape[1])


@pytest.fixture(scope="session")
def ref_inputs(tmp_path_factory):
    """DATASET_PATH holding byte-identical copies of the reference's own inputs/cora/* (tests/golden/cora_ref.tar.xz, digests in cora_ref.json)."""
    import tarfile
    d = tmp_path_factory.mktemp("ref_inputs")
    with tarfile.open(os.path.join(GOLDEN_DIR, "cora_ref.tar.xz")) as t:
        t.extractall(d, filter="data")
    return str(d) + "/"


@pytest.fixture(scope="session")
def small_graph(golden):
    """Power-law graph with a hub row (>1024 neighbours) and isolated vertices; inputs regenerated from seeds."""
    rp64, ci = golden["sg_rowptr64"], golden["sg_colidx"]
    n = len(rp64) - 1
    rng = np.random.default_rng(5)
    xs = {}
    for F in (7, 16, 47, 100, 256):
        xs[F] = rng.standard_normal((n, F), dtype=np.float32)
        assert sha(xs[F]) == str(golden[f"sg_x_{F}_sha"]), "numpy RNG stream changed; regenerate goldens"
    return dict(rowptr64=rp64, rowptr=rp64.astype(np.uint32), colidx=ci, n=n, x=xs)


@pytest.fixture(scope="session")
def liborc():
    import oracle
    return oracle.liborc()


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a CUDA device is required for -m gpu tests (no CPU fallback exists)")
