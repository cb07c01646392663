#!/bin/bash
# ncu summaries of the kernels added late in round 1 (ingest, batch witness, verify); text only.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
P="python tools/new_kernels_probe.py"
$P > gpurun_out/new_kernels.log 2>&1 || { tail -5 gpurun_out/new_kernels.log; exit 1; }
cat gpurun_out/new_kernels.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_late.csv $P > /dev/null 2>&1
# k_witness<128, 3>: the 325-pass launch of the 444-pass batch; k_plonk_verify: one proof (launch 0) and 32,768 proofs, 16 per warp (launch 8)
ncu --set full --clock-control none -k regex:"k_witness<128" --launch-skip 0 -c 1 -o /tmp/ncu/wit128 $P > /dev/null 2>&1
python tools/ncu_keys.py /tmp/ncu/wit128.ncu-rep 0 > gpurun_out/ncu_late_k_witness_128x3.txt 2>/dev/null
ncu --set full --clock-control none -k regex:"k_plonk_verify" --launch-skip 1 -c 1 -o /tmp/ncu/ver1 $P > /dev/null 2>&1
python tools/ncu_keys.py /tmp/ncu/ver1.ncu-rep 0 > gpurun_out/ncu_late_k_plonk_verify_1.txt 2>/dev/null
ncu --set full --clock-control none -k regex:"k_plonk_verify" --launch-skip 9 -c 1 -o /tmp/ncu/ver16 $P > /dev/null 2>&1
python tools/ncu_keys.py /tmp/ncu/ver16.ncu-rep 0 > gpurun_out/ncu_late_k_plonk_verify_32768x16.txt 2>/dev/null
grep -h "gpu__time_duration\|launch__grid\|inst_executed.avg.per_cycle_active\|smsp__thread_inst_executed_per_inst" gpurun_out/ncu_late_k_witness_128x3.txt gpurun_out/ncu_late_k_plonk_verify_*.txt
