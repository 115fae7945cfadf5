# Round-2 measurement run (one B200): tests, smoke, the contract benchmark (both arms), the per-op tables at batch 8 and 64,
# the ncu launch list of the benchmark and one `ncu --set full` capture of every kernel family.  Only small summaries are
# left in gpurun_out/ (the .ncu-rep files stay in /tmp on the box: gpurun_out is limited to 64 MiB).
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -n 3 > gpurun_out/r02n_tests.log
python __graft_entry__.py --smoke > gpurun_out/r02n_smoke.log 2>&1
python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err
python tools/bench_ops.py --iters 20 > gpurun_out/r02n_ops_b8.json 2> gpurun_out/r02n_ops_b8.err
python tools/bench_ops.py --iters 10 --batch 64 --only calib,photometric,warp,geometric,disparity,lcn,reduce,xcorrvol > gpurun_out/r02n_ops_b64.json 2> gpurun_out/r02n_ops_b64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --no-strong > gpurun_out/r02n_ncu_bench.log 2>&1
ncu --set full --clock-control none -c 40 -o /tmp/r02n_allops python tools/experiments/all_ops_profile_target.py > gpurun_out/r02n_ncu_allops.log 2>&1
python tools/ncu_summary.py /tmp/r02n_allops.ncu-rep > gpurun_out/r02n_ncu_allops_summary.md 2> gpurun_out/r02n_ncu_summary.err
python tools/ncu_traffic.py /tmp/r02n_allops.ncu-rep > gpurun_out/r02n_traffic.json 2>> gpurun_out/r02n_ncu_summary.err
tail -n 2 gpurun_out/r02n_tests.log gpurun_out/r02n_smoke.log gpurun_out/r02n_ncu_allops.log
du -sh gpurun_out
