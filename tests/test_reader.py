"""the loader on the REFERENCE's bytes. tests/golden/cora_ref.tar.xz holds byte-identical copies of the dataset the
reference ships (inputs/cora/*, sha256s in cora_ref.json); cora_ref.json also holds digests of what the reference's own Reader
(src/gnn/reader.cpp:248-457, compiled into oracle/_ref) returns for them. This repository's Reader (host/gai_graph.cpp) must return the
same bytes. Host-only code: runs on CPU here and on the GPU box alike."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tarfile

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT

sys.path.insert(0, GOLDEN_DIR)
import make_cora_ref as mk  # noqa: E402  (shares the two-call loader protocol with the golden generator)

GOLD = json.load(open(os.path.join(GOLDEN_DIR, "cora_ref.json")))


@pytest.fixture(scope="module")
def cora_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("ref_inputs")
    with tarfile.open(os.path.join(GOLDEN_DIR, "cora_ref.tar.xz")) as t:
        t.extractall(d, filter="data")
    return str(d)


def test_fixture_is_byte_identical_to_the_reference_files(cora_dir):
    for name, digest in GOLD["files"].items():
        data = open(os.path.join(cora_dir, "cora", name), "rb").read()
        assert hashlib.sha256(data).hexdigest() == digest, name
        ref = os.path.join("/root/reference/inputs/cora", name)
        if os.path.exists(ref):  # build container: against the reference tree itself
            assert open(ref, "rb").read() == data, name


def _ours(cora_dir):
    """Our Reader in a child process (DATASET_PATH is part of the reference's CLI contract and is read from the environment)."""
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import make_cora_ref as mk; "
            "from graphaibench_b200 import model; model.hostlib(); "
            "print(json.dumps(mk.load_with(model.HOSTLIB_PATH, 'gai_reader_load', 'cora')))") % (ROOT, GOLDEN_DIR)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DATASET_PATH=cora_dir + "/"), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reader_matches_the_reference_reader_on_reference_bytes(cora_dir):
    got = _ours(cora_dir)
    want = GOLD["reference_reader"]
    assert got["meta"] == want["meta"]
    for k in ("rowptr_u32", "colidx", "feats", "labels_single", "labels_multi"):
        assert got[k] == want[k], k


def test_reader_against_live_reference_reader():
    """Build container only: both loaders on the reference tree's own files, in the same run."""
    import oracle
    if not (os.path.isdir("/root/reference/inputs/cora") and oracle.have_ref()):
        pytest.skip("needs /root/reference and oracle/_ref (build container)")
    env = dict(os.environ, DATASET_PATH="/root/reference/inputs/")
    ref = subprocess.run([sys.executable, os.path.join(GOLDEN_DIR, "make_cora_ref.py"), "--child"], env=env, capture_output=True, text=True)
    assert ref.returncode == 0, ref.stderr[-2000:]
    assert json.loads(ref.stdout.strip().splitlines()[-1]) == _ours("/root/reference/inputs")
    # a second shipped dataset, live only. Its meta file stops after the class counts: the reference leaves the split fields uninitialised
    # (garbage), this Reader leaves them 0 — everything else must agree
    for ds in ("citeseer",):
        code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import make_cora_ref as mk; import os; "
                "from graphaibench_b200 import model; model.hostlib(); "
                "a = mk.load_with(model.HOSTLIB_PATH, 'gai_reader_load', %r); "
                "b = mk.load_with(os.path.join(%r, 'oracle', '_ref', 'libref_gnn.so'), 'ref_reader_load', %r); "
                "a['meta'] = a['meta'][:4]; b['meta'] = b['meta'][:4]; print(json.dumps(a == b))") % (ROOT, GOLDEN_DIR, ds, ROOT, ds)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-2000:]
        assert json.loads(out.stdout.strip().splitlines()[-1]) is True


def test_golden_npz_agrees_with_reference_bytes(cora_dir, cora):
    """The sparse cora fixture used by the training tests is the same data as the reference files."""
    rp = np.fromfile(os.path.join(cora_dir, "cora", "graph.vertex.bin"), np.int64)
    ci = np.fromfile(os.path.join(cora_dir, "cora", "graph.edge.bin"), np.uint32)
    fe = np.fromfile(os.path.join(cora_dir, "cora", "graph.feats.bin"), np.float32).reshape(cora["nv"], cora["feat_len"])
    assert np.array_equal(rp, cora["rowptr64"]) and np.array_equal(ci, cora["colidx"]) and np.array_equal(fe, cora["feats"])
    assert np.array_equal(np.fromfile(os.path.join(cora_dir, "cora", "graph.vlabel.bin"), np.uint8), cora["labels"])


def test_legacy_csgr_reader_matches_the_reference_reader(tmp_path):
    """reader.cpp:16-246 (.csgr graph, -labels.txt, -feats.bin + -dims.txt, -*_mask.txt) on byte-identical copies of inputs/gnn-tester/*
    (tests/golden/gnn-tester), against digests of what the reference's Reader returns for them."""
    import shutil
    src = os.path.join(GOLDEN_DIR, "gnn-tester")
    for name, digest in GOLD["gnn_tester_files"].items():
        assert hashlib.sha256(open(os.path.join(src, name), "rb").read()).hexdigest() == digest, name
    shutil.copytree(src, tmp_path / "tester")
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import make_cora_ref as mk; "
            "from graphaibench_b200 import model; model.hostlib(); "
            "print(json.dumps(mk.load_csgr_with(model.HOSTLIB_PATH, 'gai_reader_load_csgr', 'tester')))") % (ROOT, GOLDEN_DIR)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DATASET_PATH=str(tmp_path) + "/"), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert json.loads(out.stdout.strip().splitlines()[-1]) == GOLD["reference_reader_csgr"]
