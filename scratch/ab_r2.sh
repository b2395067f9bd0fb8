#!/bin/bash
# A/B of the round-2 forward variants on one box: each line = bench.py --skip-extras (20 steps) under one setting
out=gpurun_out/r2_ab_$1.txt
: > $out
for rep in 1 2; do
for v in "B200SURV_P1=ring B200SURV_PDL=0" "B200SURV_P1=ring B200SURV_PDL=1" "B200SURV_PDL=0" "B200SURV_PDL=1"; do
  echo "== $v" >> $out
  env $v python bench.py --steps 30 --warmup 5 --skip-extras >> $out 2>&1
done
done
