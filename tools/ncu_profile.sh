#!/bin/bash
# Profiling recipe (run under gpurun, one GPU): launch list of the bench command + one full-section capture of the hot kernels.
# usage: tools/ncu_profile.sh <tag>
TAG=${1:-r1}
OURS='regex:spmm_|gemm_|sgemm_|wgrad_|pad_|prep_b|splitk|adam_k|relu_k|softmax_ce|loss_acc|norms_k|fill_k|transpose_perm|gather_rows|scores_k|sddmm|softmax_bwd|colsum|alpha_grad|el_er|l2norm|hub_'
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
# second epoch's hot kernels (skip the first epoch: 14 matching launches per SAGE epoch), all sections
ncu --set full --clock-control none --import-source on -k 'regex:spmm_rows|spmm_hub|gemm_tc' -s ${HOT_SKIP:-14} -c ${HOT_COUNT:-14} -f -o gpurun_out/${TAG}_hot \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_hot.log 2>&1
ls -la gpurun_out/
