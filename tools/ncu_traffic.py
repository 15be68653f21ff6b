#!/usr/bin/env python
"""profiles/r1_hot_kernels.csv (tools/ncu_table.py output of one epoch's hot kernels, in launch order) -> profiles/traffic.json:
measured DRAM bytes per launch for the ops bench.py names. The C2 SAGE epoch launches, in order: AGGR mean F=100, K-cat transform,
N-cat transform, AGGR mean F=47, [loss], AGGR meanT F=47, two-B dW, K-cat+mask transform, two-A dW."""
import csv, json, sys
rows = list(csv.DictReader(open(sys.argv[1])))
names = ["AGGR mean F=100", "LINEAR 2449029x256x100+100 kcat", "LINEAR 2449029x47x256 ncat2", "AGGR mean F=47", "AGGR meanT F=47",
         "LINEAR 256x47x2449029 TA two_b", "LINEAR 2449029x256x47+47 TB kcat bitmask", "LINEAR 100x256x2449029 TA two_a"]
kinds = ["spmm_rows", "gemm_tc_kernel", "gemm_tc_kernel", "spmm_rows", "spmm_rows", "gemm_tc_wgrad", "gemm_tc_kernel", "gemm_tc_wgrad"]
# find the first launch of the epoch: an spmm_rows kernel followed by two gemm_tc_kernel launches
start = next(i for i in range(len(rows) - 2) if "spmm_rows" in rows[i]["kernel"] and "gemm_tc_kernel" in rows[i + 1]["kernel"] and "gemm_tc_kernel" in rows[i + 2]["kernel"])
out, i = {}, start
for name, kind in zip(names, kinds):
    while i < len(rows) and kind not in rows[i]["kernel"]:
        i += 1
    if i == len(rows):
        break
    r = rows[i]
    out[name] = int((float(r["dram_read_GB"]) + float(r["dram_write_GB"])) * 1e9)
    i += 1
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
