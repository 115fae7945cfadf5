#!/bin/bash
# One GPU call that produces everything profiles/ holds for a round (run under gpurun from the repo root):
#   bash tools/profile_round.sh r01e
# 1. bench.py (default flags) -> gpurun_out/bench_<tag>.json        (never under a profiler)
# 2. tools/bench_ops.py at batch 8 and 64 -> gpurun_out/ops_<tag>_b{8,64}.json
# 3. ncu launch list of the same bench command -> gpurun_out/launches_<tag>.csv
# 4. ncu --set full of the bench chain's kernels and of XCorrVol -> gpurun_out/prof_<tag>.ncu-rep, xc_<tag>.ncu-rep
# Summaries are made afterwards, here: tools/ncu_summary.py, tools/ncu_traffic.py.
tag=${1:-rXX}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
python tools/bench_ops.py --batch 8 > gpurun_out/ops_${tag}_b8.json 2> gpurun_out/ops_${tag}_b8.err
python tools/bench_ops.py --batch 64 --only calib,photometric,warp,geometric,disparity,lcn,xcorrvol,reduce > gpurun_out/ops_${tag}_b64.json 2> gpurun_out/ops_${tag}_b64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -c 16 -f -o gpurun_out/prof_$tag python tools/profile_target.py > gpurun_out/ncu_prof_$tag.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:xcorr -s 6 -c 6 -f -o gpurun_out/xc_$tag python tools/experiments/xcorr_profile_target.py 8 > gpurun_out/ncu_xc_$tag.log 2>&1
ls -la gpurun_out/*$tag*
