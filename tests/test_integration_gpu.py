"""INTEGRATION.md §B, run: the reference's UNCHANGED driver (src/gnn/train.cpp, net.cpp, reader.cpp, loss_layer.cpp, sampler.cpp,
random.cpp, layers/l2norm_layer.cpp, dense_layer.cpp, compiled from the reference tree by integration/build.sh in the build container)
linked against integration/b200_objset.cpp — the symbols of the reference's `.cu` twins, each body one call into include/gai_b200.h —
and libgai_b200.so. The binaries travel to the GPU box prebuilt. They must reproduce the reference's own CPU results on the reference's
own dataset files (tests/golden/cora_ref.tar.xz): test accuracy 0.795 (GCN, 200 epochs), 0.784 (SAGE, 100), 0.771 (GAT, 100)."""
import os
import subprocess

import pytest

from conftest import ROOT, require_cuda

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "integration", "_build")


@pytest.mark.parametrize("arch,epochs,want_acc,loss0", [("gcn", 200, "cora_gcn_test_acc", 1.946), ("sage", 100, "cora_sage_test_acc", None),
                                                         ("gat", 100, "cora_gat_test_acc", None)])
def test_reference_driver_over_the_b200_object_set(golden, ref_inputs, arch, epochs, want_acc, loss0):
    require_cuda()
    exe = os.path.join(BUILD, f"gpu_train_{arch}_b200")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run integration/build.sh in the build container (it needs /root/reference)")
    from graphaibench_b200 import build
    build.build_all()
    out = subprocess.run([exe, "cora", str(epochs), "1", "softmax"], env=dict(os.environ, DATASET_PATH=ref_inputs), capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-1500:])
    lines = out.stdout.splitlines()
    test = float([l for l in lines if l.startswith("Test accuracy:")][0].split()[2])
    assert abs(test - float(golden[want_acc])) < 1e-3, (test, float(golden[want_acc]))
    if loss0 is not None:
        ep0 = [l for l in lines if l.startswith("Epoch   0")][0]
        assert f"train_loss {loss0:.3f}" in ep0, ep0
    # the reference's per-epoch losses (goldens from its CPU build) are reproduced by its own driver over this object set
    ref_losses = golden[f"cora_{arch}_losses"]
    got = [float(l.split("train_loss")[1].split()[0]) for l in lines if l.startswith("Epoch") and "train_loss" in l]
    assert len(got) == epochs
    for ep in (0, 1, 5, 20, epochs - 1):
        assert abs(got[ep] - float(ref_losses[ep])) <= 0.002 + 0.02 * float(ref_losses[ep]), (ep, got[ep], float(ref_losses[ep]))
