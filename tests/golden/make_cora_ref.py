"""Build-container only: pins the loader to the REFERENCE's bytes and the REFERENCE's own Reader.

    python tests/golden/make_cora_ref.py

Writes
  tests/golden/cora_ref.tar.xz   byte-identical copies of /root/reference/inputs/cora/* (the dataset the reference ships)
  tests/golden/cora_ref.json     sha256 of every file, and sha256 of what the reference's Reader (src/gnn/reader.cpp:248-457, compiled
                                 into oracle/_ref/libref_gnn.so) returns for them: row pointers (as u32), column indices, features,
                                 single-class labels, multi-hot labels, and the meta tuple.
tests/test_reader.py then feeds the same bytes to this repository's Reader, here and on the GPU box.
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = "/root/reference/inputs/cora"


def sha_bytes(b):
    return hashlib.sha256(b).hexdigest()


def load_with(lib_path, fn_name, dataset):
    """Two-call protocol shared by ref_reader_load (oracle/ref_harness.cpp) and gai_reader_load (host/gai_host_capi.cpp)."""
    L = C.CDLL(lib_path)
    fn = getattr(L, fn_name)
    fn.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    out = {}
    for single in (1, 0):
        meta = np.zeros(13, np.int64)
        fn(dataset.encode(), single, meta.ctypes.data_as(C.c_void_p), None, None, None, None)
        nv, ne, flen, ncls = (int(x) for x in meta[:4])
        rp, ci = np.zeros(nv + 1, np.uint32), np.zeros(ne, np.uint32)
        feats = np.zeros((nv, flen), np.float32)
        labels = np.zeros(nv if single else nv * ncls, np.uint8)
        fn(dataset.encode(), single, meta.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
           feats.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p))
        out["meta"] = [int(x) for x in meta]
        out["rowptr_u32"], out["colidx"], out["feats"] = sha_bytes(rp.tobytes()), sha_bytes(ci.tobytes()), sha_bytes(feats.tobytes())
        out["labels_single" if single else "labels_multi"] = sha_bytes(labels.tobytes())
    return out


def load_csgr_with(lib_path, fn_name, dataset):
    """Legacy .csgr / text layout (reader.cpp:16-246) through ref_reader_load_csgr / gai_reader_load_csgr (same two-call protocol)."""
    L = C.CDLL(lib_path)
    fn = getattr(L, fn_name)
    fn.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 6
    out = {}
    for single in (1, 0):
        meta = np.zeros(13, np.int64)
        fn(dataset.encode(), single, meta.ctypes.data_as(C.c_void_p), None, None, None, None, None)
        nv, ne, flen, ncls = (int(x) for x in meta[:4])
        rp, ci = np.zeros(nv + 1, np.uint32), np.zeros(ne, np.uint32)
        feats = np.zeros((nv, flen), np.float32)
        labels = np.zeros(nv if single else nv * ncls, np.uint8)
        masks = np.zeros(3 * nv, np.uint8)
        fn(dataset.encode(), single, meta.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
           feats.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p), masks.ctypes.data_as(C.c_void_p))
        out["meta"] = [int(x) for x in meta]
        out["rowptr_u32"], out["colidx"], out["feats"] = sha_bytes(rp.tobytes()), sha_bytes(ci.tobytes()), sha_bytes(feats.tobytes())
        out["masks"] = sha_bytes(masks.tobytes())
        out["labels_single" if single else "labels_multi"] = sha_bytes(labels.tobytes())
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child-csgr":   # DATASET_PATH = a directory holding tester/ (copy of inputs/gnn-tester)
        print(json.dumps(load_csgr_with(os.path.join(ROOT, "oracle", "_ref", "libref_gnn.so"), "ref_reader_load_csgr", "tester")))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--child":   # DATASET_PATH must be set before the reference library loads (configs.h:5)
        print(json.dumps(load_with(os.path.join(ROOT, "oracle", "_ref", "libref_gnn.so"), "ref_reader_load", "cora")))
        sys.exit(0)
    names = sorted(os.listdir(SRC))
    files = {n: sha_bytes(open(os.path.join(SRC, n), "rb").read()) for n in names}
    tar_path = os.path.join(HERE, "cora_ref.tar.xz")
    with tarfile.open(tar_path, "w:xz", preset=9) as t:
        for n in names:
            ti = t.gettarinfo(os.path.join(SRC, n), arcname=f"cora/{n}")
            ti.mtime, ti.uid, ti.gid, ti.uname, ti.gname = 0, 0, 0, "", ""
            with open(os.path.join(SRC, n), "rb") as f:
                t.addfile(ti, f)
    env = dict(os.environ, DATASET_PATH="/root/reference/inputs/")
    child = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, check=True)
    ref = json.loads(child.stdout.strip().splitlines()[-1])
    # legacy layout: tests/golden/gnn-tester/* are byte-identical copies of inputs/gnn-tester/* (the reader wants them in a directory named
    # after the dataset, "tester")
    import shutil, tempfile
    with tempfile.TemporaryDirectory() as d:
        shutil.copytree("/root/reference/inputs/gnn-tester", os.path.join(d, "tester"))
        child = subprocess.run([sys.executable, os.path.abspath(__file__), "--child-csgr"], env=dict(os.environ, DATASET_PATH=d + "/"),
                               capture_output=True, text=True, check=True)
    csgr = json.loads(child.stdout.strip().splitlines()[-1])
    csgr_files = {n: sha_bytes(open(os.path.join("/root/reference/inputs/gnn-tester", n), "rb").read()) for n in sorted(os.listdir("/root/reference/inputs/gnn-tester"))}
    json.dump({"source": "chenxuhao/GraphAIBench inputs/cora", "files": files, "reference_reader": ref,
               "gnn_tester_files": csgr_files, "reference_reader_csgr": csgr},
              open(os.path.join(HERE, "cora_ref.json"), "w"), indent=1)
    print("wrote", tar_path, os.path.getsize(tar_path), "bytes")
