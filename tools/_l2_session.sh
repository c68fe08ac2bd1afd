#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
exec > gpurun_out/l2.log 2>&1
date
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sector_op_read_hit_rate.pct,gpu__time_duration.sum
for v in default r1state; do
  if [ $v = default ]; then unset PTG_B200_SO; else export PTG_B200_SO=$PWD/variants/$v.so; fi
  for pol in sticky uniform; do
    timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_step --launch-skip 700 -c 3 --csv --log-file gpurun_out/l2_${v}_$pol.csv python tools/microbench.py --steps 100 --policy $pol --no-rollout > /dev/null 2>&1
    echo "== $v $pol"; grep -E "dram__bytes|hit_rate|time_duration" gpurun_out/l2_${v}_$pol.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
  done
done
date
