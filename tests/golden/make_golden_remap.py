"""Golden vectors of the rectification row, produced by OpenCV itself in the build container
(cv2 4.13 wheels): python tests/golden/make_golden_remap.py

remap_T0.npz: distorted synthetic frame + the maps of cv2.initUndistortRectifyMap (TUM1 distortion coefficients,
settings/TUM1.yaml:32-36, alpha-scaled new camera matrix as volumetric_integrator_base.py:766-778) + the outputs
of cv2.remap (colour INTER_LINEAR, depth / labels INTER_NEAREST) and of cv2.cvtColor(BGR2RGB)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyslam_b200 import synthetic as S  # noqa: E402
from tests._util import GOLDEN  # noqa: E402


def main():
    cfg = S.CONFIGS["T0"]
    d, c, T = S.render_frame(cfg, 2)
    h, w = d.shape
    K = np.array([[cfg.fx, 0, cfg.cx], [0, cfg.fy, cfg.cy], [0, 0, 1]], float)
    D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314])  # TUM1 k1 k2 p1 p2 k3
    new_K, _ = cv2.getOptimalNewCameraMatrix(K, D, (w, h), 0.5, (w, h))
    map1, map2 = cv2.initUndistortRectifyMap(K, D, None, new_K, (w, h), cv2.CV_32FC1)
    bgr = np.ascontiguousarray(c[..., ::-1])
    color_u = cv2.remap(bgr, map1, map2, interpolation=cv2.INTER_LINEAR)
    depth_u = cv2.remap(d, map1, map2, interpolation=cv2.INTER_NEAREST)
    labels = (np.arange(h * w, dtype=np.int32).reshape(h, w) // 37) % 11
    labels_u = np.ascontiguousarray(cv2.remap(labels.astype(np.float32), map1, map2, interpolation=cv2.INTER_NEAREST),
                                    dtype=np.int32)
    rgb_u = cv2.cvtColor(color_u, cv2.COLOR_BGR2RGB)
    np.savez_compressed(os.path.join(GOLDEN, "remap_T0.npz"), depth=d, bgr=bgr, Tcw=T, K=cfg.K, new_K=new_K,
                        map1=map1, map2=map2, color_u=color_u, depth_u=depth_u, rgb_u=rgb_u, labels=labels,
                        labels_u=labels_u)
    print("remap_T0:", map1.shape, "new_K", np.round(new_K, 3).tolist(), "zero border pixels",
          int((depth_u == 0).sum()))


if __name__ == "__main__":
    main()
