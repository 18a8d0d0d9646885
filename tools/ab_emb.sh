#!/bin/bash
# A/B of the embeddedness formulations / launch shapes on the bench workload (per-kernel CUDA-event times + phase cycle counters)
R=${R:-1024}; T=${T:-96}
for cfg in ${CFGS:-"POC_EMB_V=1" "POC_EMB_V=2" "POC_EMB_V=2,POC_EMB_BLOCKS=6" "POC_EMB_V=2,POC_EMB_G=128" "POC_EMB_V=2,POC_EMB_G=512"}; do
  echo "=== $cfg"
  env ${cfg//,/ } python tools/profile_run.py --replicas $R --ticks $T --timing 2>&1 | grep -v "^$"
done
