cap() {
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o /tmp/$1 python tools/prof_run.py 20 8 1 > gpurun_out/r3q_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r3q_$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r3q_$1.source.csv.gz
}
cap finish 'k_wave_finish<.int.8, .int.1>' 5
cap start_small 'k_prop_start<.int.8, .int.0>' 5
