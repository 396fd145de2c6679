import numpy as np, oracle
from pyslam_b200 import VoxelBlockSemanticProbabilisticGrid
from tests._util import sort_dump
rng = np.random.default_rng(11)
ref = oracle.RefSemanticGrid(0.05, "probabilistic"); ref.set_depth_threshold(1.5); ref.set_depth_decay_rate(0.8)
grid = VoxelBlockSemanticProbabilisticGrid(0.05, 8, capacity_blocks=4096); grid.set_depth_threshold(1.5); grid.set_depth_decay_rate(0.8)
variants = [dict(f64=True, u8=False, inst=True, depth=True), dict(f64=False, u8=True, inst=True, depth=False),
            dict(f64=True, u8=False, inst=False, depth=True), dict(f64=False, u8=False, inst=False, depth=False)]
for vi, var in enumerate(variants):
    n=30000
    dirs = rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pts = dirs * (0.5 + 0.01 * rng.normal(size=(n, 1))) + [0.05, -0.1, 0.02]
    pts = pts.astype(np.float64 if var["f64"] else np.float32)
    cols_u8 = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    cols_f = (cols_u8.astype(np.float32) * (np.float32(1.0) / np.float32(255.0))) if var["u8"] else rng.random((n, 3)).astype(np.float32)
    side = (pts[:, 0] > 0).astype(np.int32)
    flip, noise = rng.random(n) < 0.2, rng.integers(-1, 1, n)
    cls = np.where(flip, noise, 1 + side).astype(np.int32)
    ins = np.where(flip, noise, 10 + side).astype(np.int32)
    dep = rng.uniform(0.5, 4.0, n).astype(np.float32)
    ref.integrate(pts, cols_f, cls, ins if var["inst"] else None, dep if var["depth"] else None)
    grid.integrate(pts, cols_u8 if var["u8"] else cols_f, cls, ins if var["inst"] else None, dep if var["depth"] else None)
    a, b = sort_dump(grid.dump_blocks(8)), sort_dump(ref.dump_blocks(8))
    ca, cb = a["confidence"].reshape(-1), b["confidence"].reshape(-1)
    rel = np.abs(ca - cb) / np.maximum(cb, 1e-12)
    bad = np.where(rel > 2e-6)[0]
    la, lb = a["lab_logp"].reshape(-1), b["lab_logp"].reshape(-1)
    fin = np.isfinite(la) & np.isfinite(lb)
    print("   logp: finite-mask equal", np.array_equal(np.isfinite(la), np.isfinite(lb)), "n diff", int((la[fin] != lb[fin]).sum()),
          "max abs diff", float(np.abs(la[fin] - lb[fin]).max()), "max conf rel", float(rel.max()),
          "pos equal", np.array_equal(a["pos_sum"], b["pos_sum"]), "col equal", np.array_equal(a["col_sum"], b["col_sum"]))
    print("variant", vi, "mismatches", len(bad), "of", int((b["count"] > 0).sum()), "labels equal",
          np.array_equal(a["lab_logp"], b["lab_logp"]), np.array_equal(a["object_id"], b["object_id"]))
    for v in bad[:6]:
        print("  v", v, "gpu", ca[v], "ref", cb[v], "count", b["count"].reshape(-1)[v], "nl", b["aux"].reshape(-1)[v],
              "obj/cls", b["object_id"].reshape(-1)[v], b["class_id"].reshape(-1)[v],
              "lab", list(zip(b["lab_obj"].reshape(-1, 8)[v], b["lab_cls"].reshape(-1, 8)[v], b["lab_logp"].reshape(-1, 8)[v]))[:6])
