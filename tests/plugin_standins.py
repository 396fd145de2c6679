"""Stand-ins for the pieces of pySLAM's plugin contract the adapter touches, with the same names
and fields, so `VolumetricIntegratorB200` can run (and be tested) without pySLAM installed.

Mirrors, field for field:
  VolumetricIntegrationTaskType      pyslam/dense/volumetric_integrator_base.py:90-96
  VolumetricIntegrationKeyframeData  base.py:100-137   (id, pose = Tcw, img BGR, depth, ...)
  VolumetricIntegrationTask          base.py:140-156
  VolumetricIntegrationPointCloud    base.py:159-210
  VolumetricIntegrationMesh          base.py:213-226
  VolumetricIntegrationOutput        base.py:308-322
  StandaloneIntegratorBase           the subset of VolumetricIntegratorBase (base.py:328-1394) that
                                     `init` / `volume_integration` rely on, run in-process with
                                     queue.Queue instead of a spawned process.
"""

from __future__ import annotations

import queue
import threading
import time
from enum import Enum
from types import SimpleNamespace

import numpy as np


class VolumetricIntegrationTaskType(Enum):
    NONE = 0
    INTEGRATE = 1
    SAVE = 2
    LOAD = 3
    RESET = 4
    UPDATE_OUTPUT = 5


class DatasetEnvironmentType(Enum):
    INDOOR = 1
    OUTDOOR = 2


class VolumetricIntegrationKeyframeData:
    def __init__(self, id=-1, pose=None, img=None, depth=None, img_right=None, semantic_img=None,
                 semantic_instances_img=None, camera=None, timestamp=-1):
        self.id = id
        self.kid = id
        self.img_id = id
        self.timestamp = timestamp
        self.pose = pose  # Tcw
        self.camera = camera
        self.img = img
        self.img_right = img_right
        self.depth = depth
        self.semantic_img = semantic_img
        self.semantic_instances_img = semantic_instances_img


class VolumetricIntegrationTask:
    def __init__(self, keyframe_data=None, task_type=VolumetricIntegrationTaskType.NONE,
                 load_save_path=None):
        self.task_type = task_type
        self.keyframe_data = keyframe_data
        self.load_save_path = load_save_path


class VolumetricIntegrationPointCloud:
    def __init__(self, point_cloud=None, points=None, colors=None):
        if point_cloud is not None:
            points, colors = point_cloud.points, point_cloud.colors
        self.points = np.asarray(points) if points is not None else None
        self.colors = np.asarray(colors) if colors is not None else None
        self.semantics = None
        self.object_ids = None
        self.semantic_colors = None


class VolumetricIntegrationMesh:
    def __init__(self, mesh):
        self.vertices = np.asarray(mesh.vertices)
        self.triangles = np.asarray(mesh.triangles)
        self.vertex_colors = np.asarray(mesh.vertex_colors)
        self.vertex_normals = np.asarray(mesh.vertex_normals)


class VolumetricIntegrationOutput:
    def __init__(self, task_type, id=-1, point_cloud=None, mesh=None, objects=None):
        self.task_type = task_type
        self.id = id
        self.point_cloud = point_cloud
        self.mesh = mesh
        self.objects = objects
        self.timestamp = time.perf_counter()


class VolumetricIntegrationObject:
    """Fields of pyslam/dense/volumetric_integrator_base.py:262-282."""

    def __init__(self, object_data):
        self.points = np.ascontiguousarray(object_data.points, np.float32)
        self.colors = np.ascontiguousarray(object_data.colors, np.float32)
        self.class_id, self.object_id = object_data.class_id, object_data.object_id
        self.confidence_min, self.confidence_max = object_data.confidence_min, object_data.confidence_max
        self.box_matrix = np.ascontiguousarray(np.asarray(object_data.oriented_bounding_box.get_matrix()).T)
        self.box_size = np.ascontiguousarray(object_data.oriented_bounding_box.size, np.float64)


class VolumetricIntegrationObjectList:
    """Fields of pyslam/dense/volumetric_integrator_base.py:285-305."""

    def __init__(self, object_data_group, semantic_colors, object_colors, num_objects):
        self.object_list = [VolumetricIntegrationObject(o) for o in object_data_group.object_vector]
        self.semantic_colors = np.ascontiguousarray(semantic_colors, np.float32)
        self.object_colors = np.ascontiguousarray(object_colors, np.float32)
        self.num_objects = num_objects


class _Value:
    def __init__(self, v):
        self.value = v


class StandaloneIntegratorBase:
    """In-process subset of `VolumetricIntegratorBase`: same constructor signature, `init`,
    `add_keyframe`-style task submission, `save`, `request_reset`-style RESET, `pop_output`, `quit`;
    `step()` runs one `volume_integration` call (the reference's `run()` loop body, base.py:905-915)."""

    print = staticmethod(lambda *a, **k: None)

    def __init__(self, camera, environment_type, sensor_type, volumetric_integrator_type,
                 viewer_queue=None, **kwargs):
        self.camera = camera
        self.environment_type = environment_type
        self.sensor_type = sensor_type
        self.volumetric_integrator_type = volumetric_integrator_type
        self.viewer_queue = viewer_queue
        self.constructor_kwargs = kwargs
        self.volume = None
        self.q_in, self.q_out, self.q_management = queue.Queue(), queue.Queue(), queue.Queue()
        self.q_out_condition = threading.Condition()
        self.save_request_condition = threading.Condition()
        self.load_request_condition = threading.Condition()
        self.is_running = _Value(1)
        self.save_request_completed = _Value(-1)
        self.load_request_completed = _Value(-1)
        self.time_volumetric_integration = _Value(0.0)
        self.last_input_task = None
        self.last_output = None
        self.last_integrated_id = -1
        self.parameters_dict = dict(kwargs.get("parameters", {}))
        self.init(camera, environment_type, sensor_type, self.parameters_dict, kwargs)

    def init(self, camera, environment_type, sensor_type, parameters_dict, constructor_kwargs):
        self.depth_factor = 1.0  # base.py:713
        # base.py:766-778 computes these with cv2.initUndistortRectifyMap from camera.D; the stand-in takes
        # ready-made maps (`calib_maps=(map1, map2)`) so that it does not need OpenCV
        maps = (constructor_kwargs or {}).get("calib_maps")
        self.calib_map1, self.calib_map2 = (maps if maps is not None else (None, None))
        self.depth_estimator = None

    def get_camera_intrinsics_for_depth(self):
        c = self.camera
        return c.fx, c.fy, c.cx, c.cy

    def estimate_depth_if_needed_and_rectify(self, kd):
        """depth -> float32 metres; undistortion with cv2.remap when maps exist (base.py:1017-1047; the CPU
        path the plugin's GPU rectification replaces); colour BGR -> RGB (base.py:1054)."""
        if kd.depth is None or kd.depth.size == 0:
            return None, None, None, None, None
        depth = kd.depth if kd.depth.dtype == np.float32 else kd.depth.astype(np.float32)
        color = kd.img
        if self.calib_map1 is not None and self.calib_map2 is not None:
            import cv2  # only this CPU leg of the stand-in needs OpenCV
            color = cv2.remap(color, self.calib_map1, self.calib_map2, interpolation=cv2.INTER_LINEAR)
            depth = cv2.remap(depth, self.calib_map1, self.calib_map2, interpolation=cv2.INTER_NEAREST)
        color = np.ascontiguousarray(color[..., ::-1])
        return color, depth, None, kd.semantic_img, kd.semantic_instances_img

    # -- front-end API --
    def add_task(self, task):
        self.q_in.put(task)

    def add_keyframe_data(self, kd):
        self.add_task(VolumetricIntegrationTask(kd, VolumetricIntegrationTaskType.INTEGRATE))

    def add_update_output_task(self):
        self.add_task(VolumetricIntegrationTask(None, VolumetricIntegrationTaskType.UPDATE_OUTPUT))

    def save(self, path):
        self.save_request_completed.value = 0
        self.add_task(VolumetricIntegrationTask(None, VolumetricIntegrationTaskType.SAVE,
                                                load_save_path=path + "/dense_map.ply"))

    def reset(self):
        self.q_management.put(VolumetricIntegrationTask(None, VolumetricIntegrationTaskType.RESET))

    def step(self):
        self.volume_integration(self.q_in, self.q_out, self.q_out_condition, self.q_management,
                                self.viewer_queue, self.is_running, self.load_request_completed,
                                self.load_request_condition, self.save_request_completed,
                                self.save_request_condition, self.time_volumetric_integration)

    def run_pending(self):
        while not self.q_in.empty() and self.is_running.value == 1:
            self.step()

    def pop_output(self):
        try:
            return self.q_out.get_nowait()
        except queue.Empty:
            return None

    def quit(self):
        self.is_running.value = 0
        stop = getattr(self, "_stop_volume_integrator_implementation", None)
        if stop:
            stop()


API = SimpleNamespace(
    VolumetricIntegrationTaskType=VolumetricIntegrationTaskType,
    VolumetricIntegrationOutput=VolumetricIntegrationOutput,
    VolumetricIntegrationMesh=VolumetricIntegrationMesh,
    VolumetricIntegrationPointCloud=VolumetricIntegrationPointCloud,
    VolumetricIntegrationObjectList=VolumetricIntegrationObjectList,
    DatasetEnvironmentType=DatasetEnvironmentType, Parameters=None)


def standalone_integrator_class():
    from pyslam_b200.integrator import make_integrator_class
    return make_integrator_class(StandaloneIntegratorBase, API)


def standalone_voxel_grid_integrator_class():
    from pyslam_b200.integrator_semantic import make_voxel_grid_integrator_class
    return make_voxel_grid_integrator_class(StandaloneIntegratorBase, API)


def standalone_semantic_integrator_class():
    from pyslam_b200.integrator_semantic import make_semantic_integrator_class
    return make_semantic_integrator_class(StandaloneIntegratorBase, API)
