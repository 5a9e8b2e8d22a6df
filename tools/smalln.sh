# throughput of the two device forms of the proposal phase over the chain count (development aid)
for lg in ${LGS:-17 18 19}; do for wf in 1 0; do echo "n=2^$lg wavefront=$wf: $(LMC_WAVEFRONT=$wf timeout 120 python tools/prof_run.py $lg ${STEPS:-32} 2 2>&1 | grep 'launch 1')"; done; done
