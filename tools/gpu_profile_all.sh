#!/bin/bash
# every launch of the two traversal kernels in a short run under ncu --set full (for tools/ncu_traffic.py), plus the
# source-level view of the dominant launch:  tools/gpu_profile_all.sh TAG [replicas] [ticks] [warm-ticks]
TAG=${1:-r02}; R=${2:-1024}; T=${3:-2}; W=${4:-0}
OUT=gpurun_out; mkdir -p $OUT
CMD="python tools/profile_run.py --replicas $R --ticks $T --warm-ticks $W"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
# launches of the warm-up ticks to skip: 4 replica groups x (2 rounds x 2 kernels | 3 search launches) per tick
for K in emb:k_emb3:16 search:k_search:12; do
  N=${K%%:*}; KN=${K#*:}; PER=${KN##*:}; KN=${KN%%:*}
  ncu --set full --clock-control none --import-source on -k regex:$KN -s $((W * PER)) -f -o $OUT/${TAG}_$N $CMD > $OUT/${TAG}_${N}_ncu.log 2>&1
  ncu -i $OUT/${TAG}_$N.ncu-rep --page raw --csv > $OUT/${TAG}_${N}_raw.csv 2>/dev/null
  ls -la $OUT/${TAG}_$N.ncu-rep
  rm -f $OUT/${TAG}_$N.ncu-rep
done
tail -4 $OUT/${TAG}_plain.log
