"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the volumetric-integration hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package; the product (`pyslam_b200/`) never does.

* `RefGrid`     - the UNMODIFIED reference `volumetric::VoxelBlockGrid` compiled from
                  /root/reference/cpp/volumetric (oracle/_ref/libref_grid.so): truth for block keys,
                  `BlockKeyHash`, local voxel index, counts and point/colour running sums.
* `RefSemanticGrid` - the UNMODIFIED reference `VoxelBlockSemanticGrid` / `VoxelBlockSemanticProbabilisticGrid`
                  (oracle/_ref/libref_semantic.so): truth for label voting / Bayesian label fusion.
* `TsdfOracle`  - C restatement of Open3D's legacy ScalableTSDFVolume under decision D1
                  (oracle/tsdf_oracle.c): truth for tsdf / weight / rgb and the mesh.
                  PARITY UNPINNED against Open3D itself (not installed, not vendored).
* `numpy_tsdf`  - a second, independent numpy restatement of A.3 used to pin the C oracle.
"""

from .oracle import (RefGrid, RefSemanticGrid, TsdfOracle, build, canonical_mesh, have_ref, have_ref_semantic, numpy_integrate_block,
                     numpy_touched_blocks, ref_block_key_hash, ref_floor_div, ref_keys)

__all__ = ["RefGrid", "RefSemanticGrid", "TsdfOracle", "build", "canonical_mesh", "have_ref", "have_ref_semantic", "numpy_integrate_block",
           "numpy_touched_blocks", "ref_block_key_hash", "ref_floor_div", "ref_keys"]
