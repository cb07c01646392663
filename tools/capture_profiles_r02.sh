#!/bin/bash
# Round-2 profile capture on the GPU box.  Every command is first run plain (it must exit 0) and only then under ncu
# (B200_PROFILING.md).  Text / JSON summaries only come back (gpurun_out is size-capped); numbers printed by a run
# under ncu are never bench values.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
OUT=gpurun_out
export NZCB_LANES=1
B="python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline"
M="python tools/msm_affine_probe.py 21 3 1 3"
export WR_B=444
W="python tools/witness_rate.py"

# 1. launch list of one single-lane proof
$B > $OUT/r02_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/r02_plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file /tmp/ncu/launches.csv $B > /dev/null 2>&1
python tools/launch_summary.py /tmp/ncu/launches.csv 3 > $OUT/r02_launches_single_lane.txt
gzip -c /tmp/ncu/launches.csv > $OUT/r02_launches_single_lane.csv.gz

# 2. the accumulation of a dense 3-commitment MSM batch: every kernel of it, --set full
$M > $OUT/r02_plain_msm.log 2>&1 || { echo "plain msm probe failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k_aff|k_msm_accum" -c 10 -o /tmp/ncu/accum $M > /dev/null 2>&1
python tools/ncu_accum_summary.py /tmp/ncu/accum.ncu-rep $OUT/r02_msm_accum_ncu.json > $OUT/r02_ncu_msm_accum.txt

# 3. the witness program, 325 passes in flight (the first launch of a 444-pass batch)
$W > $OUT/r02_plain_witness.log 2>&1 || { echo "plain witness probe failed"; exit 1; }
ncu --set full --clock-control none -k regex:"k_witness" --launch-skip 2 -c 2 -o /tmp/ncu/wit $W > /dev/null 2>&1
python tools/ncu_keys.py /tmp/ncu/wit.ncu-rep 0 > $OUT/r02_ncu_k_witness_325_passes.txt 2>/dev/null
python tools/ncu_keys.py /tmp/ncu/wit.ncu-rep 1 > $OUT/r02_ncu_k_witness_119_passes.txt 2>/dev/null

# 4. the other hot kernels of a proof
for k in k_ntt_pass k_round3 k_msm_digits k_bred; do
  skip=6; [ $k = k_ntt_pass ] && skip=150; [ $k = k_round3 ] && skip=2; [ $k = k_msm_digits ] && skip=28; [ $k = k_bred ] && skip=12
  ncu --set full --clock-control none -k regex:"^$k|^void $k" --launch-skip $skip -c 1 -o /tmp/ncu/$k $B > /dev/null 2>&1
  python tools/ncu_keys.py /tmp/ncu/$k.ncu-rep 0 > $OUT/r02_ncu_$k.txt 2>/dev/null
done
cat $OUT/r02_launches_single_lane.txt | head -30
ls -la $OUT | grep r02_
