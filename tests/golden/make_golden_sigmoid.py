"""Golden vectors for the sigmoid (multi-label) loss, generated from the LIVE reference (oracle/_ref/libref_gnn.so:
sigmoid_loss_layer + masked_accuracy_multi of /root/reference). Run in the build container: python tests/golden/make_golden_sigmoid.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

R = oracle.libref()
rng = np.random.default_rng(21)
nv, nc = 500, 13
x = (rng.standard_normal((nv, nc)) * 4).astype(np.float32)
x[3, :4] = [0.0, -0.0, 60.0, -60.0]
y = (rng.random((nv, nc)) < 0.25).astype(np.uint8)
m = (rng.random(nv) < 0.7).astype(np.uint8)
b, e = 17, 431
cnt = int(m[b:e].sum())
probs = np.zeros((nv, nc), np.float32); losses = np.zeros(nv, np.float32); grad = np.zeros((nv, nc), np.float32); f1 = C.c_float()
loss = R.ref_sigmoid_loss(nv, nc, x.reshape(-1), y.reshape(-1), m.ctypes.data_as(C.c_void_p), b, e, cnt, probs.ctypes.data_as(C.c_void_p),
                          losses.ctypes.data_as(C.c_void_p), grad.ctypes.data_as(C.c_void_p), C.byref(f1))
sel = m.astype(bool); sel[:b] = False; sel[e:] = False
losses[~sel] = 0
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sigmoid.npz"), x=x, y=y, m=m, b=b, e=e, probs=probs, losses=losses, grad=grad,
                    loss=np.float32(loss), f1=np.float32(f1.value))
print("mean loss", loss, "micro-F1", f1.value, "rows", cnt)
