#!/bin/bash
# Per-kernel durations of one batch-8 XCorrVol call (ncu launch list) -> gpurun_out/xc_launches.csv
set -e
python - <<'PY' > gpurun_out/xc_target.log 2>&1
print("warm")
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/xc_launches.csv \
  python tools/experiments/xcorr_profile_target.py 8 > gpurun_out/xc_ncu.log 2>&1
