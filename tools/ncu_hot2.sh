cap() {
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o /tmp/$1 python tools/prof_run.py 20 8 1 > gpurun_out/r3d_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r3d_$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r3d_$1.source.csv.gz
}
cap trace_w0 'k_trace' 40
cap gcam_w1 'k_shade<.int.8, .int.4>' 41
cap shadow 'k_shadow' 5
cap tail 'k_shade_tail' 2
ls -la gpurun_out/r3d_*
