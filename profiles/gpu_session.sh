#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu --set full of the HD kernels.
# usage (under gpurun): bash profiles/gpu_session.sh <tag> [tests] [bench] [launches] [full]
TAG=${1:-x}; shift
OUT=gpurun_out
mkdir -p $OUT
for what in "$@"; do
case $what in
tests)
  timeout ${TESTS_TIMEOUT:-1500} python -m pytest tests -m gpu -x -q --durations=8 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log;;
bench)
  timeout 600 python bench.py > $OUT/${TAG}_bench.jsonl 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.jsonl; tail -3 $OUT/${TAG}_bench.err;;
benchq)
  timeout 600 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.jsonl 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.jsonl; tail -3 $OUT/${TAG}_bench.err;;
benchdet)
  timeout 600 python bench.py --no-cpu-baseline --deterministic > $OUT/${TAG}_bench_det.jsonl 2> $OUT/${TAG}_bench_det.err; echo "benchdet rc=$?"; cat $OUT/${TAG}_bench_det.jsonl; tail -3 $OUT/${TAG}_bench_det.err;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/${TAG}_launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches.log 2>&1; echo "launches rc=$?";;
launchesw)
  # launch list of another workload, eager (no graph) so that every kernel shows:  LW=kitti_rollout
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/${TAG}_launches_${LW}.csv \
     python bench.py --workload $LW --no-graph --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-others > $OUT/${TAG}_launches_${LW}.log 2>&1; echo "launchesw rc=$?";;
benchw)
  # one line of another workload:  LW=nonrigid_train
  timeout 600 python bench.py --workload $LW --no-cpu-baseline --no-others > $OUT/${TAG}_bench_$LW.jsonl 2> $OUT/${TAG}_bench_$LW.err; echo "bench $LW rc=$?"; cat $OUT/${TAG}_bench_$LW.jsonl; tail -2 $OUT/${TAG}_bench_$LW.err;;
workloads)
  for w in city_rollout kitti_rollout; do
    timeout 600 python bench.py --workload $w --no-cpu-baseline > $OUT/${TAG}_bench_$w.jsonl 2> $OUT/${TAG}_bench_$w.err; echo "bench $w rc=$?"; cat $OUT/${TAG}_bench_$w.jsonl; tail -2 $OUT/${TAG}_bench_$w.err
  done;;
refgpu)
  timeout 600 python bench.py --impl reference-gpu --steps 3 > $OUT/${TAG}_bench_refgpu.jsonl 2> $OUT/${TAG}_bench_refgpu.err; echo "refgpu rc=$?"; cat $OUT/${TAG}_bench_refgpu.jsonl; tail -2 $OUT/${TAG}_bench_refgpu.err;;
ref)
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.jsonl 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"; cat $OUT/${TAG}_bench_ref.jsonl; tail -2 $OUT/${TAG}_bench_ref.err;;
exp)
  timeout ${EXP_TIMEOUT:-300} python scratch/exp.py $EXP_NAMES > $OUT/${TAG}_exp.log 2>&1; echo "exp rc=$?"; cat $OUT/${TAG}_exp.log;;
fullk)
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$FULL_K" -c ${FULL_C:-1} \
     -o $OUT/${TAG}_fullk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_fullk.log 2>&1; echo "fullk rc=$?"; tail -2 $OUT/${TAG}_fullk.log;;
full)
  timeout 1500 ncu --set full --clock-control none --import-source on \
     -k 'regex:k_gather_bwd|k_layers_bwd|k_alpha_prep_bwd|k_gather_fwd|k_layers_fwd|k_alpha_prep|k_class_profile|k_inv_fused' -c 9 \
     -o $OUT/${TAG}_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/${TAG}_full.log 2>&1; echo "full rc=$?"; tail -2 $OUT/${TAG}_full.log;;
esac
done
