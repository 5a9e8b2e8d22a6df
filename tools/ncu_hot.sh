#!/bin/bash
# ncu --set full captures of the hot kernels of one steady-state iteration (torus, LMC, maxdepth 8, 2^20 chains).
# usage (on the GPU box): tools/ncu_hot.sh TAG   -> gpurun_out/TAG_<kernel>.ncu-rep
TAG=${1:-hot}
cap() {  # name regex skip
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 \
      -f -o /tmp/${TAG}_$1 python tools/prof_run.py 20 8 1 > gpurun_out/${TAG}_$1.log 2>&1
  # the reports are ~30 MB each and gpurun_out/ carries 64 MiB: keep the raw metrics and the per-line source view as csv
  ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_$1.source.csv.gz
}
# invocation index of wave w of iteration 5: 3 x 6 (pre-calibration) + 15 (calibration) + 7 + w
cap shade_pcam_w0 'k_shade<.int.8, .int.2>' 40
cap shade_pcam_w2 'k_shade<.int.8, .int.2>' 42
cap shade_gcam_w0 'k_shade<.int.8, .int.4>' 40

cap start_small 'k_prop_start<.int.8, .int.0>' 5
cap finish 'k_wave_finish<.int.8, .int.1>' 5
cap grad_prop 'k_wave_grad<.int.8, .int.1>' 11
ls -la gpurun_out/${TAG}_*
