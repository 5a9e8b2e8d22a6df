#!/bin/bash
# compute-sanitizer passes over small jobs of every configuration (wavefront path forced).  usage (GPU box): tools/sanitize.sh OUT.txt
OUT=${1:-gpurun_out/sanitizer.txt}
export LMC_WAVEFRONT=1
run() {  # tool label env... -- args
  local tool=$1 label=$2; shift 2
  local res
  res=$(env "$@" compute-sanitizer --tool $tool --print-limit 5 python tools/prof_run.py $ARGS 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | tail -3 | tr '\n' ' ')
  printf "%-10s %-58s %s\n" "$tool" "$label" "$res" >> $OUT
}
echo "compute-sanitizer runs on one B200 (round 2 build, $(date -u +%F)):" > $OUT
ARGS="13 6 1";  run memcheck "torus LMC L8, 8192 chains x 6 (reverse-sweep gradient)" LMC_SCENE=torus/lmc.xml
ARGS="12 4 1";  run memcheck "door LMC L12, 4096 chains x 4" LMC_SCENE=veachdoor/lmc.xml LMC_MAXDEPTH=12
ARGS="11 4 1";  run memcheck "torus H2MC L8, 2048 chains x 4 (FoR Hessian + k_h2mc_gaussian)" LMC_SCENE=torus/h2mc.xml
ARGS="12 6 1";  run memcheck "torus textured.xml L6, 4096 chains x 6 (textured parameters)" LMC_SCENE=torus/textured.xml LMC_MAXDEPTH=6
ARGS="12 130 1"; run memcheck "torus LMC L6 global cache on, 4096 chains x 130 (k_cache_*, k_cache_prequery)" LMC_SCENE=torus/lmc.xml LMC_MAXDEPTH=6 LMC_OPTS=globalcache=1
ARGS="12 130 1"; run racecheck "torus LMC L6 global cache on (shared lists of k_cache_prequery)" LMC_SCENE=torus/lmc.xml LMC_MAXDEPTH=6 LMC_OPTS=globalcache=1
ARGS="12 130 1"; run synccheck "torus LMC L6 global cache on" LMC_SCENE=torus/lmc.xml LMC_MAXDEPTH=6 LMC_OPTS=globalcache=1
ARGS="13 6 1";  run synccheck "torus LMC L8" LMC_SCENE=torus/lmc.xml
ARGS="11 4 1";  run synccheck "torus H2MC L8 (half-warp __syncwarp masks of k_h2mc_gaussian)" LMC_SCENE=torus/h2mc.xml
ARGS="11 3 1";  run racecheck "torus H2MC L8 (shared-memory Jacobi)" LMC_SCENE=torus/h2mc.xml
ARGS="12 3 1";  run racecheck "torus LMC L8" LMC_SCENE=torus/lmc.xml
unset LMC_WAVEFRONT
ARGS="12 6 1";  LMC_WAVEFRONT=0 run memcheck "torus LMC L8 monolithic form, 4096 chains x 6 (large-step proposals on the auxiliary stream)" LMC_SCENE=torus/lmc.xml LMC_WAVEFRONT=0
ARGS="12 6 1";  run synccheck "door LMC L12 (warp-aggregated queue counters)" LMC_SCENE=veachdoor/lmc.xml LMC_MAXDEPTH=12 LMC_WAVEFRONT=1
cat $OUT
