#!/bin/bash
# r02 final visits (1 GPU).  gpurun brings back at most 64 MiB per call, so the work is split by mode:
#   bench   both bench arms as the driver runs them + the ncu launch list of the bench command
#   lexmarch / lexprep   ncu --set full: the march launches of one lexingtonHII20 shoot / its first emission kernels
#   lean    ncu --set full: one launch of the coherent walk per 256^3 grid
MODE=${1:-bench}
TAG=${2:-r02final}
OUT=gpurun_out
mkdir -p $OUT
if [ $MODE = bench ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 5 > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err
  echo "reference arm rc=$?"; cut -c1-300 $OUT/bench_reference_$TAG.json
  timeout 900 python bench.py > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err
  echo "bench rc=$?"; tail -2 $OUT/bench_1gpu_$TAG.err | cut -c1-300
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_$TAG.json 2> $OUT/bench_under_ncu_$TAG.err
  echo "ncu launch list rc=$?"; wc -l $OUT/launches_bench_$TAG.csv
elif [ $MODE = lexmarch ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k 'regex:march_kernel' --launch-count 16 \
    -f -o $OUT/lex_march_$TAG python tools/profile_shoot.py --problem lexington --packets 16777216 --spinup-packets 2000000 > $OUT/ncu_lex_march_$TAG.log 2>&1
  echo "ncu lexington march rc=$?"; grep "shoot of" $OUT/ncu_lex_march_$TAG.log | cut -c1-300
elif [ $MODE = lexprep ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:prepare_kernel|tail_kernel|reemit_decide' --launch-count 4 \
    -f -o $OUT/lex_prepare_$TAG python tools/profile_shoot.py --problem lexington --packets 16777216 --spinup-packets 2000000 > $OUT/ncu_lex_prepare_$TAG.log 2>&1
  echo "ncu lexington prepare rc=$?"
else
  for prob in stromgren256 clumpy256; do
    timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:march_lean_kernel' --launch-count 1 \
      -f -o $OUT/lean_${prob}_$TAG python tools/profile_shoot.py --problem $prob --packets 16000000 --spinup-packets 16000000 > $OUT/ncu_lean_${prob}_$TAG.log 2>&1
    echo "ncu $prob rc=$?"; grep "shoot of" $OUT/ncu_lean_${prob}_$TAG.log | cut -c1-300
  done
fi
ls -la $OUT/*$TAG* ; du -sh $OUT
