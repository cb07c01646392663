#!/bin/bash
# ncu summaries of the kernels added late in round 1 (ingest, batch witness, verify); text only.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
P="python tools/new_kernels_probe.py"
$P > gpurun_out/new_kernels.log 2>&1 || { tail -5 gpurun_out/new_kernels.log; exit 1; }
cat gpurun_out/new_kernels.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_late.csv $P > /dev/null 2>&1
for k in k_pass_ingest k_witness k_plonk_verify; do
  skip=1; [ $k = k_plonk_verify ] && skip=3
  ncu --set full --clock-control none -k regex:"$k" --launch-skip $skip -c 1 -o /tmp/ncu/$k $P > /dev/null 2>&1
  python tools/ncu_keys.py /tmp/ncu/$k.ncu-rep 0 > gpurun_out/ncu_late_$k.txt 2>/dev/null
done
grep -h "gpu__time_duration\|dram__bytes_read.sum \|dram__bytes_write.sum \|launch__grid" gpurun_out/ncu_late_*.txt
