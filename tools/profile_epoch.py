"""One warm-up epoch + N epochs of the bench workload, for ncu (launch list / --set full captures). Never a bench value.
usage: python tools/profile_epoch.py [--arch sage] [--scale 1] [--epochs 2]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--gemm-mode", type=int, default=0)
    a = ap.parse_args()
    import torch
    from graphaibench_b200 import _abi, model as gmodel
    _abi.lib().gai_set_gemm_mode(a.gemm_mode)
    w = bench.make_workload(a.scale, "cuda")
    s = torch.cuda.Stream()
    with bench.quiet_stdout():
        m = gmodel.GnnModel("sage", w["rowptr"], w["colidx"], w["feats"], w["labels"], w["split"], bench.C2["hid"], bench.C2["ncls"],
                            num_layers=bench.C2["layers"], lr=bench.C2["lr"], stream=s.cuda_stream)
    for _ in range(1 + a.epochs):
        m.train_epoch()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
