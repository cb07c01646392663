#!/bin/bash
# Round profile capture on the GPU box: launch list of one single-lane proof, ncu --set full of the dominant kernel
# (dense 3-commitment batch) and of the other hot kernels; only text summaries are kept (gpurun_out is size-capped).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
export NZCB_LANES=1
B="python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 3 > gpurun_out/launch_summary.txt
ncu --set full --clock-control none --import-source on -k regex:k_msm_accum --launch-skip 14 -c 1 -o /tmp/ncu/accum $B > /dev/null 2>&1
python tools/ncu_keys.py /tmp/ncu/accum.ncu-rep > gpurun_out/ncu_msm_accum.txt
ncu -i /tmp/ncu/accum.ncu-rep --page details > gpurun_out/ncu_msm_accum_details.txt 2>/dev/null
for k in k_round3 k_ntt_pass k_msm_digits k_bred k_witness; do
  skip=6; [ $k = k_ntt_pass ] && skip=150; [ $k = k_witness ] && skip=3; [ $k = k_round3 ] && skip=2; [ $k = k_msm_digits ] && skip=28; [ $k = k_bred ] && skip=60
  ncu --set full --clock-control none -k regex:"^$k" --launch-skip $skip -c 2 -o /tmp/ncu/$k $B > /dev/null 2>&1
  python tools/ncu_keys.py /tmp/ncu/$k.ncu-rep 0 > gpurun_out/ncu_$k.txt 2>/dev/null
  python tools/ncu_keys.py /tmp/ncu/$k.ncu-rep 1 >> gpurun_out/ncu_$k.txt 2>/dev/null
done
ls -la gpurun_out
