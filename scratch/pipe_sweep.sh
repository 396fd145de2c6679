#!/bin/bash
python -m pytest tests/test_gpu_tsdf.py -k "sharded_fused or fused_batch_equals" -x -q 2>&1 | tail -3
for LB in 6 5; do
  python - <<PY
import re
p='pyslam_b200/csrc/b2v_tsdf.cu'
s=open(p).read()
s=re.sub(r'__launch_bounds__\(kIntThreads, kPipe \? \d : 8\)', '__launch_bounds__(kIntThreads, kPipe ? $LB : 8)', s)
open(p,'w').write(s)
PY
  python -m pyslam_b200.build > /dev/null 2>&1
  grep -n "integrate_group_kernelILb1" -A 3 pyslam_b200/build/ptxas.log | grep -E "registers" | sed "s/^/LB=$LB /"
  for lat in 0 1; do for s in 8 4 2 1; do echo -n "LB=$LB latency_variant=$lat "; B2V_LATENCY_VARIANT=$lat python scratch/tl8.py $s 2>&1 | grep -E "^shards"; done; done
done
