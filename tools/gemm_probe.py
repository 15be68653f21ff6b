#!/usr/bin/env python
"""Times the dense-transform kernels on the C2 shapes (N = 2 449 029 rows), one JSON line per (shape, knob) to stdout.

    python tools/gemm_probe.py            # all shapes, all GAI_TC_DEBUG knob settings (one subprocess per setting)
    python tools/gemm_probe.py --knob 0   # one setting, in-process

GAI_TC_DEBUG knobs are timing experiments (gemm_tc.cu): 1 = weights loaded once per stage slot, 2 = no global stores,
4 = no hi/lo split, 8 = no MMA issue. Results under a non-zero knob are numerically wrong by design."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N = 2_449_029


def run(knob, n, only=""):
    import torch
    from graphaibench_b200 import ops
    dev = "cuda:0"
    g = torch.Generator(device=dev); g.manual_seed(1)
    R = lambda r, c, ld=None: (torch.randn(r, ld or c, generator=g, device=dev)[:, :c])
    x100, h100 = R(n, 100), R(n, 100)
    h256 = R(n, 256)
    g48a, g48b = R(n, 47, 48), R(n, 47, 48)
    out256 = torch.empty(n, 256, device=dev)
    o48a, o48b = torch.empty(n, 48, device=dev)[:, :47], torch.empty(n, 48, device=dev)[:, :47]
    W = lambda r, c: torch.randn(r, c, generator=g, device=dev) * 0.1
    w1, w2, v1, v2 = W(100, 256), W(100, 256), W(256, 47), W(256, 47)
    cases = {
        "fwd 256x100": (lambda: ops.matmul(x100, w1, out=out256), 4.0 * n * (100 + 256)),
        "fwd kcat 100+100->256 relu": (lambda: ops.matmul_kcat(h100, w1, x100, w2, out=out256, flags=ops.EPI_RELU), 4.0 * n * (200 + 256)),
        "fwd ncat 256->47,47": (lambda: ops.matmul_ncat(h256, v1, v2, out1=o48a, out2=o48b), 4.0 * n * (256 + 94)),
        "bwd 47->256 TB": (lambda: ops.matmul(g48a, v1, out=out256, transB=True), 4.0 * n * (47 + 256)),
        "bwd kcat 47+47->256 TB mask": (lambda: ops.matmul_kcat(g48a, v1, g48b, v2, out=out256, transB=True, flags=ops.EPI_MASK, mask=h256),
                                        4.0 * n * (94 + 512)),
        "wgrad two_a 100,100 x 256": (lambda: ops.wgrad_two_a(h100, x100, h256), 4.0 * n * (200 + 256)),
        "wgrad two_b 256 x 47,47": (lambda: ops.wgrad_two_b(h256, g48a, g48b), 4.0 * n * (256 + 94)),
    }
    for name, (fn, nbytes) in cases.items():
        if name.startswith("wgrad") and (knob & 3):
            continue
        if only and not name.startswith(only):
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"case": name, "knob": knob, "rows": n, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--knob", type=int, default=None)
    ap.add_argument("--rows", type=int, default=N)
    ap.add_argument("--only", default="", help="case-name prefix filter")
    a = ap.parse_args()
    if a.knob is not None:
        run(a.knob, a.rows, a.only)
    else:
        for k in [int(v) for v in os.environ.get("GAI_PROBE_KNOBS", "0,1,2,4,8,3,7,15").split(",")]:
            env = dict(os.environ, GAI_TC_DEBUG=str(k))
            subprocess.run([sys.executable, os.path.abspath(__file__), "--knob", str(k), "--rows", str(a.rows), "--only", a.only], env=env, check=False)
