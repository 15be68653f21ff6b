#!/bin/bash
# Profiling recipe whose outputs fit gpurun's 64 MiB return limit: launch list of the bench command + one full-section capture of the
# hot kernels, post-processed ON THE BOX (raw page -> per-launch table, source page -> stall hot spots of the aggregation and the dense
# transform), after which the .ncu-rep itself is deleted.   usage: tools/ncu_profile_small.sh <tag>
TAG=${1:-r2}
OURS='regex:spmm_|gemm_|sgemm_|wgrad_|pad_|prep_b|splitk|adam_k|relu_k|softmax_ce|loss_acc|norms_k|fill_k|transpose_perm|gather_rows|scores_k|sddmm|softmax_bwd|colsum|alpha_grad|el_er|l2norm|hub_'
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
REP=/tmp/${TAG}_hot
ncu --set full --clock-control none --import-source on -k 'regex:spmm_rows|spmm_hub|gemm_tc' -s ${HOT_SKIP:-14} -c ${HOT_COUNT:-11} -f -o $REP \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_hot.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv > /tmp/${TAG}_raw.csv 2>/dev/null
python tools/ncu_table.py /tmp/${TAG}_raw.csv > gpurun_out/${TAG}_hot_kernels.csv
for K in spmm_rows gemm_tc_kernel gemm_tc_wgrad; do
  ncu -i $REP.ncu-rep --page source --csv --kernel-name regex:$K --launch-count 1 > /tmp/${TAG}_src_$K.csv 2>/dev/null
  python tools/ncu_hotspots.py /tmp/${TAG}_src_$K.csv 25 > gpurun_out/${TAG}_hotspots_$K.txt 2>&1
done
rm -f $REP.ncu-rep
du -sh gpurun_out
