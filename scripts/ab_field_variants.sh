for rep in 1 2; do
for v in "1 1" "1 0" "0 1" "0 0"; do
  set -- $v
  SAHS_FIELD_DUO=$1 SAHS_FIELD_WIDE=$2 timeout 300 python bench.py --steps 8 --warmup 3 --no-train --no-cpu-baseline --no-torch-gpu > gpurun_out/ab_duo$1_wide$2_$rep.json 2>/dev/null
done
done
