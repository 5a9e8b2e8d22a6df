#!/bin/bash
# ncu --set full captures of door-scene kernels in steady state.  usage (GPU box): tools/ncu_door.sh TAG
TAG=${1:-door}
export LMC_SCENE=veachdoor/lmc.xml LMC_MAXDEPTH=12
cap() {
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o /tmp/$1 python tools/prof_run.py 20 16 4 > gpurun_out/${TAG}_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_$1.source.csv.gz
}
cap glgt 'k_shade<.int.12, .int.3>' 600
cap plgt 'k_shade<.int.12, .int.1>' 600
cap connect 'k_connect' 40
cap pcam 'k_shade<.int.12, .int.2>' 800
ls -la gpurun_out/${TAG}_*
