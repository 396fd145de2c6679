#!/bin/bash
# sweep the allocate_group_kernel residency hint (rebuilds on the GPU box)
for LB in 8 6 4; do
  python - <<PY
import re
p='pyslam_b200/csrc/b2v_tsdf.cu'
s=open(p).read()
s=re.sub(r'__launch_bounds__\(kAllocThreads, \d\)\nallocate_group_kernel\(', '__launch_bounds__(kAllocThreads, $LB)\nallocate_group_kernel(', s)
open(p,'w').write(s)
PY
  python -m pyslam_b200.build > /dev/null 2>&1
  grep -n "allocate_group_kernelILb1" -A 3 pyslam_b200/build/ptxas.log | grep -E "registers" | sed "s/^/LB=$LB /"
  for s in 8 1; do python scratch/tl8.py $s 2>&1 | grep -E "shards"; done
done
