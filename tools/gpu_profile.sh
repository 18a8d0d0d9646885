#!/bin/bash
# ncu evidence for the two traversal kernels on the bench workload (run on the GPU box under gpurun):
#   tools/gpu_profile.sh TAG [replicas] [ticks]
# writes gpurun_out/TAG_{emb,search}_{raw,source}.csv (+ the .ncu-rep files when small) after a plain run exited 0.
TAG=${1:-r02}; R=${2:-1024}; T=${3:-16}
OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/profile_run.py --replicas $R --ticks $T ${EXTRA:-}"
$CMD --timing > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
for K in ${KERNELS:-emb:k_emb search:k_search}; do
  N=${K%%:*}; KN=${K##*:}
  ncu --set full --clock-control none --import-source on -k regex:$KN -s ${SKIP:-2} -c 2 -f -o $OUT/${TAG}_$N $CMD > $OUT/${TAG}_${N}_ncu.log 2>&1
  ncu -i $OUT/${TAG}_$N.ncu-rep --page raw --csv > $OUT/${TAG}_${N}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$N.ncu-rep --page source --csv --print-source cuda,sass > $OUT/${TAG}_${N}_source.csv 2>/dev/null
  ls -la $OUT/${TAG}_$N.ncu-rep
  [ $(stat -c %s $OUT/${TAG}_$N.ncu-rep) -gt 20000000 ] && rm -f $OUT/${TAG}_$N.ncu-rep
done
tail -12 $OUT/${TAG}_plain.log
