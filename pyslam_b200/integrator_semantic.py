"""`VolumetricIntegratorB200SemanticGrid` — the plugin class for pySLAM's semantic dense backend
(`VolumetricIntegratorVoxelSemanticGrid`, pyslam/dense/volumetric_integrator_voxel_semantic_grid.py) on the GPU.

Same contract as `integrator.py`: a subclass of `VolumetricIntegratorBase` built against whichever base / task /
output types it is given (pySLAM's real ones, or the in-process stand-ins of `tests/plugin_standins.py`).  The per-frame
loop body (reference :326-461) maps one-to-one onto the C ABI:

    filter_shadow_points + depth2pointcloud + world transform + integrate  ->  b2v_sgrid_integrate_rgbd
    assign_object_ids_to_instance_ids (+ carving)                           ->  b2v_sgrid_assign_object_ids_to_instance_ids
    remap_instance_ids                                                      ->  pyslam_b200.remap_instance_ids
    carve (no instance ids)                                                 ->  b2v_sgrid_carve
    get_voxels(min_count, min_confidence)                                   ->  b2v_sgrid_get_voxels / copy_voxels

    get_object_segments(min_count, min_confidence)                          ->  b2v_sgrid_get_voxels + grouping / PCA boxes

Outputs: when 2-D instance ids are integrated, the reference's OBJECTS representation (:523-587:
`get_object_segments` -> `VolumetricIntegrationObjectList`, one entry per object with its points, colours, class id,
confidence range and oriented box); otherwise its single-point-cloud representation (:590-700): points, colours,
class ids, object ids.
"""

from __future__ import annotations

import time
import traceback

import numpy as np

from .integrator import write_ply_points
from .volume import (CameraFrustrum, VoxelBlockGrid, VoxelBlockSemanticGrid, VoxelBlockSemanticProbabilisticGrid,
                     filter_shadow_points, remap_instance_ids)

# defaults of pyslam/config_parameters.py:311-380 (overridable through parameters_dict / constructor kwargs)
DEFAULT_PARAMETERS = {
    "kVolumetricIntegrationVoxelLength": 0.015,
    "kVolumetricIntegrationBlockSize": 8,
    "kVolumetricIntegrationTsdfDepthTruncIndoor": 4.0,
    "kVolumetricIntegrationTsdfDepthTruncOutdoor": 10.0,
    "kVolumetricIntegrationOutputTimeInterval": 1.0,
    "kVolumetricIntegrationVoxelGridMinCount": 3,
    "kVolumetricIntegrationVoxelGridMinConfidence": 0.6,
    "kVolumetricIntegrationVoxelGridUseCarving": False,
    "kVolumetricIntegrationVoxelGridCarvingDepthMin": 1e-2,
    "kVolumetricIntegrationVoxelGridCarvingDepthMaxIndoor": 8.0,
    "kVolumetricIntegrationVoxelGridCarvingDepthMaxOutdoor": 15.0,
    "kVolumetricIntegrationVoxelGridCarvingDepthThreshold": 3e-2,
    "kVolumetricIntegrationVoxelGridShadowPointsFilter": True,
    "kVolumetricSemanticProbabilisticIntegrationUseDepth": True,
    "kVolumetricSemanticProbabilisticIntegrationDepthThresholdIndoor": 5.0,
    "kVolumetricSemanticProbabilisticIntegrationDepthThresholdOutdoor": 10.0,
    "kVolumetricSemanticProbabilisticIntegrationDepthDecayRateIndoor": 0.1,
    "kVolumetricSemanticProbabilisticIntegrationDepthDecayRateOutdoor": 0.05,
    "kVolumetricSemanticIntegrationUseInstanceIds": True,
    "kVolumetricSemanticIntegrationMinVoteRatio": 0.5,
    "kVolumetricSemanticIntegrationMinVotes": 3,
    "kVolumetricIntegrationB200CapacityBlocks": 1 << 15,
    "kVolumetricIntegrationB200Device": 0,
    "kVolumetricIntegrationB200GenerateObjects": True,   # kGenerateObjectsDefault (reference :84)
}


def make_semantic_integrator_class(Base, api):
    """Build the semantic plugin class against a base class and an `api` namespace (see `integrator.py`)."""
    TaskType = api.VolumetricIntegrationTaskType

    class VolumetricIntegratorB200SemanticGrid(Base):
        """GPU semantic voxel-grid integrator; `use_semantic_probabilistic` selects Bayesian fusion (:131-141)."""

        def __init__(self, camera, environment_type, sensor_type, volumetric_integrator_type,
                     viewer_queue=None, **kwargs):
            super().__init__(camera, environment_type, sensor_type, volumetric_integrator_type,
                             viewer_queue, **kwargs)

        # -- runs inside the integrator process: the CUDA context is created here, never in the parent
        def init(self, camera, environment_type, sensor_type, parameters_dict, constructor_kwargs):
            Base.init(self, camera, environment_type, sensor_type, parameters_dict, constructor_kwargs)
            p = dict(DEFAULT_PARAMETERS)
            if parameters_dict:
                p.update({k: parameters_dict[k] for k in DEFAULT_PARAMETERS if k in parameters_dict})
            if constructor_kwargs:
                p.update({k: v for k, v in constructor_kwargs.items() if k in DEFAULT_PARAMETERS})
            self.b200_parameters = p
            indoor = True
            env_t = getattr(api, "DatasetEnvironmentType", None)
            if env_t is not None and hasattr(env_t, "INDOOR"):
                indoor = environment_type == env_t.INDOOR
            side = "Indoor" if indoor else "Outdoor"
            self.volumetric_integration_depth_trunc = p[f"kVolumetricIntegrationTsdfDepthTrunc{side}"]
            probabilistic = bool((constructor_kwargs or {}).get("use_semantic_probabilistic", False))
            grid_t = VoxelBlockSemanticProbabilisticGrid if probabilistic else VoxelBlockSemanticGrid
            self.volume = grid_t(voxel_size=p["kVolumetricIntegrationVoxelLength"],
                                 block_size=p["kVolumetricIntegrationBlockSize"],
                                 capacity_blocks=int(p["kVolumetricIntegrationB200CapacityBlocks"]),
                                 device=int(p["kVolumetricIntegrationB200Device"]))
            self.volume.set_depth_threshold(p[f"kVolumetricSemanticProbabilisticIntegrationDepthThreshold{side}"])
            self.volume.set_depth_decay_rate(p[f"kVolumetricSemanticProbabilisticIntegrationDepthDecayRate{side}"])
            fx, fy, cx, cy = self._intrinsics()
            self.camera_frustrum = CameraFrustrum(
                fx, fy, cx, cy, self.camera.width, self.camera.height, np.eye(4),
                depth_max=p[f"kVolumetricIntegrationVoxelGridCarvingDepthMax{side}"],
                depth_min=p["kVolumetricIntegrationVoxelGridCarvingDepthMin"])
            self.last_output = None
            self.last_integrated_id = -1
            self.last_instance_map = {}

        def _intrinsics(self):
            if hasattr(self, "get_camera_intrinsics_for_depth"):
                return self.get_camera_intrinsics_for_depth()
            c = self.camera
            return c.fx, c.fy, c.cx, c.cy

        def _integrate_keyframe(self, kd):
            """The reference loop body for one keyframe (:300-461)."""
            p = self.b200_parameters
            rect = self.estimate_depth_if_needed_and_rectify(kd)
            color, depth = rect[0], rect[1]
            classes = rect[3] if len(rect) > 3 else None
            instances = rect[4] if len(rect) > 4 else None
            if color is None or depth is None:
                return False
            depth = np.ascontiguousarray(depth, np.float32)
            flt = bool(p["kVolumetricIntegrationVoxelGridShadowPointsFilter"])
            self.integrated_instance_ids = False
            use_instances = (bool(p["kVolumetricSemanticIntegrationUseInstanceIds"]) and instances is not None
                             and np.asarray(instances).size > 0 and classes is not None)
            self.camera_frustrum.set_T_cw(kd.pose)
            carve_thr = float(p["kVolumetricIntegrationVoxelGridCarvingDepthThreshold"])
            object_image = None
            if use_instances or p["kVolumetricIntegrationVoxelGridUseCarving"]:
                # association and carving look at the FILTERED depth image (:349-362, :371-388)
                depth_used = filter_shadow_points(depth) if flt else depth
                if use_instances:
                    self.last_instance_map = self.volume.assign_object_ids_to_instance_ids(
                        self.camera_frustrum, classes, instances, depth_used, depth_threshold=carve_thr,
                        do_carving=bool(p["kVolumetricIntegrationVoxelGridUseCarving"]),
                        min_vote_ratio=float(p["kVolumetricSemanticIntegrationMinVoteRatio"]),
                        min_votes=int(p["kVolumetricSemanticIntegrationMinVotes"]))
                    object_image = remap_instance_ids(instances, self.last_instance_map)
                    self.integrated_instance_ids = True
                else:
                    self.volume.carve(self.camera_frustrum, depth_used, carve_thr)
            fx, fy, cx, cy = self._intrinsics()
            Twc = np.linalg.inv(np.asarray(kd.pose, np.float64).reshape(4, 4))
            self.volume.integrate_rgbd(
                depth, color, (fx, fy, cx, cy), Twc, class_image=classes, object_image=object_image,
                max_depth=self.volumetric_integration_depth_trunc,
                use_depths=bool(p["kVolumetricSemanticProbabilisticIntegrationUseDepth"]), filter_shadow_points=flt)
            self.last_integrated_id = kd.id
            return True

        def _make_output(self, task_type):
            p = self.b200_parameters
            ObjList = getattr(api, "VolumetricIntegrationObjectList", None)
            if (p["kVolumetricIntegrationB200GenerateObjects"] and getattr(self, "integrated_instance_ids", False)
                    and ObjList is not None):
                # reference :523-587: objects only when instance ids are available
                grp = self.volume.get_object_segments(
                    min_count=int(p["kVolumetricIntegrationVoxelGridMinCount"]),
                    min_confidence=float(p["kVolumetricIntegrationVoxelGridMinConfidence"]))
                sem_rgb = getattr(api, "sem_img_to_rgb", None)      # SemanticMappingShared.sem_img_to_rgb
                ids_rgb = getattr(api, "ids_to_rgb_float", None)    # IdsColorTable.ids_to_rgb_float
                n = len(grp.object_vector)
                sem_cols = (np.ascontiguousarray(sem_rgb(np.asarray(grp.class_ids), bgr=True), np.float32) / 255.0
                            if sem_rgb is not None and n else np.zeros((n, 3), np.float32))
                obj_cols = (np.ascontiguousarray(ids_rgb(np.asarray(grp.object_ids), bgr=True), np.float32)
                            if ids_rgb is not None and n else np.zeros((n, 3), np.float32))
                objects = ObjList(grp, sem_cols, obj_cols, n)
                return api.VolumetricIntegrationOutput(task_type, self.last_integrated_id, None, None, objects)
            v = self.volume.get_voxels(min_count=int(p["kVolumetricIntegrationVoxelGridMinCount"]),
                                       min_confidence=float(p["kVolumetricIntegrationVoxelGridMinConfidence"]))
            pc = api.VolumetricIntegrationPointCloud(points=np.ascontiguousarray(v.points, np.float32),
                                                     colors=np.ascontiguousarray(v.colors, np.float32))
            pc.semantics = v.class_ids if len(v.class_ids) else None
            pc.object_ids = v.object_ids if len(v.object_ids) else None
            return api.VolumetricIntegrationOutput(task_type, self.last_integrated_id, pc, None)

        def volume_integration(self, q_in, q_out, q_out_condition, q_management, viewer_queue,
                               is_running, load_request_completed, load_request_condition,
                               save_request_completed, save_request_condition,
                               time_volumetric_integration):
            t_start = time.perf_counter()
            last_output = None
            do_output = False
            try:
                if is_running.value == 1:
                    task = None
                    try:
                        task = q_management.get_nowait()
                    except Exception:
                        pass
                    if task is not None and task.task_type == TaskType.RESET:
                        self.volume.reset()
                    self.last_input_task = q_in.get()  # blocking
                    if self.last_input_task is None:
                        is_running.value = 0
                    else:
                        ttype = self.last_input_task.task_type
                        if ttype == TaskType.INTEGRATE:
                            if self._integrate_keyframe(self.last_input_task.keyframe_data):
                                do_output = True
                                if self.last_output is not None:
                                    dt = time.perf_counter() - self.last_output.timestamp
                                    if dt < self.b200_parameters["kVolumetricIntegrationOutputTimeInterval"]:
                                        do_output = False
                        elif ttype == TaskType.SAVE:
                            p = self.b200_parameters
                            v = self.volume.get_voxels(
                                min_count=int(p["kVolumetricIntegrationVoxelGridMinCount"]),
                                min_confidence=float(p["kVolumetricIntegrationVoxelGridMinConfidence"]))
                            if len(v.points):
                                write_ply_points(self.last_input_task.load_save_path, v.points, v.colors)
                            last_output = api.VolumetricIntegrationOutput(ttype)
                        elif ttype == TaskType.UPDATE_OUTPUT:
                            do_output = True
                        if do_output:
                            last_output = self._make_output(ttype)
                            self.last_output = last_output
                        if is_running.value == 1 and last_output is not None:
                            if last_output.task_type in (TaskType.INTEGRATE, TaskType.UPDATE_OUTPUT):
                                with q_out_condition:
                                    last_output.timestamp = time.perf_counter()
                                    q_out.put(last_output)
                                    q_out_condition.notify_all()
                            elif last_output.task_type == TaskType.SAVE:
                                with save_request_condition:
                                    save_request_completed.value = 1
                                    save_request_condition.notify_all()
            except Exception as e:  # the reference logs and keeps the loop alive (:720-730)
                printer = getattr(Base, "print", print)
                printer(f"VolumetricIntegratorB200SemanticGrid: EXCEPTION: {e} !!!")
                printer(traceback.format_exc())
            time_volumetric_integration.value = time.perf_counter() - t_start

        def _stop_volume_integrator_implementation(self):
            if getattr(self, "volume", None) is not None:
                self.volume.close()

    return VolumetricIntegratorB200SemanticGrid


def make_voxel_grid_integrator_class(Base, api):
    """The point-average backend (`VolumetricIntegratorVoxelGrid`, pyslam/dense/volumetric_integrator_voxel_grid.py)
    on the GPU compat grid.  Same task loop as the semantic class; the loop body is reference :232-300: optional
    shadow filter -> depth2pointcloud -> world transform -> optional carve (with the UNFILTERED depth, :283-296) ->
    integrate, all of which `VoxelBlockGrid.carve` / `integrate_rgbd` do on the device."""
    SemanticCls = make_semantic_integrator_class(Base, api)

    class VolumetricIntegratorB200VoxelGrid(SemanticCls):
        def init(self, camera, environment_type, sensor_type, parameters_dict, constructor_kwargs):
            Base.init(self, camera, environment_type, sensor_type, parameters_dict, constructor_kwargs)
            p = dict(DEFAULT_PARAMETERS)
            p["kVolumetricIntegrationB200CapacityBlocks"] = 1 << 17
            if parameters_dict:
                p.update({k: parameters_dict[k] for k in DEFAULT_PARAMETERS if k in parameters_dict})
            if constructor_kwargs:
                p.update({k: v for k, v in constructor_kwargs.items() if k in DEFAULT_PARAMETERS})
            self.b200_parameters = p
            indoor = True
            env_t = getattr(api, "DatasetEnvironmentType", None)
            if env_t is not None and hasattr(env_t, "INDOOR"):
                indoor = environment_type == env_t.INDOOR
            side = "Indoor" if indoor else "Outdoor"
            self.volumetric_integration_depth_trunc = p[f"kVolumetricIntegrationTsdfDepthTrunc{side}"]
            self.volume = VoxelBlockGrid(p["kVolumetricIntegrationVoxelLength"], p["kVolumetricIntegrationBlockSize"],
                                         capacity_blocks=int(p["kVolumetricIntegrationB200CapacityBlocks"]),
                                         device=int(p["kVolumetricIntegrationB200Device"]))
            fx, fy, cx, cy = self._intrinsics()
            self.camera_frustrum = CameraFrustrum(
                fx, fy, cx, cy, self.camera.width, self.camera.height, np.eye(4),
                depth_max=p[f"kVolumetricIntegrationVoxelGridCarvingDepthMax{side}"],
                depth_min=p["kVolumetricIntegrationVoxelGridCarvingDepthMin"])
            self.last_output = None
            self.last_integrated_id = -1

        def _integrate_keyframe(self, kd):
            p = self.b200_parameters
            rect = self.estimate_depth_if_needed_and_rectify(kd)
            color, depth = rect[0], rect[1]
            if color is None or depth is None:
                return False
            depth = np.ascontiguousarray(depth, np.float32)
            if p["kVolumetricIntegrationVoxelGridUseCarving"]:
                self.camera_frustrum.set_T_cw(kd.pose)
                self.volume.carve(self.camera_frustrum, depth,
                                  float(p["kVolumetricIntegrationVoxelGridCarvingDepthThreshold"]))
            fx, fy, cx, cy = self._intrinsics()
            Twc = np.linalg.inv(np.asarray(kd.pose, np.float64).reshape(4, 4))
            self.volume.integrate_rgbd(depth, color, (fx, fy, cx, cy), Twc,
                                       max_depth=self.volumetric_integration_depth_trunc,
                                       filter_shadow_points=bool(p["kVolumetricIntegrationVoxelGridShadowPointsFilter"]))
            self.last_integrated_id = kd.id
            return True

        def _make_output(self, task_type):
            v = self.volume.get_voxels(min_count=int(self.b200_parameters["kVolumetricIntegrationVoxelGridMinCount"]))
            pc = api.VolumetricIntegrationPointCloud(points=np.ascontiguousarray(v.points, np.float32),
                                                     colors=np.ascontiguousarray(v.colors, np.float32))
            return api.VolumetricIntegrationOutput(task_type, self.last_integrated_id, pc, None)

    return VolumetricIntegratorB200VoxelGrid


def load_pyslam_semantic_plugin():
    """The semantic plugin built against the real pySLAM types (requires pySLAM on sys.path)."""
    from types import SimpleNamespace

    from pyslam.config_parameters import Parameters
    from pyslam.dense import volumetric_integrator_base as B
    from pyslam.io.dataset_types import DatasetEnvironmentType

    api = SimpleNamespace(
        VolumetricIntegrationTaskType=B.VolumetricIntegrationTaskType,
        VolumetricIntegrationOutput=B.VolumetricIntegrationOutput,
        VolumetricIntegrationMesh=B.VolumetricIntegrationMesh,
        VolumetricIntegrationPointCloud=B.VolumetricIntegrationPointCloud,
        VolumetricIntegrationObjectList=B.VolumetricIntegrationObjectList,
        DatasetEnvironmentType=DatasetEnvironmentType, Parameters=Parameters)
    try:   # colours of the viewer: semantic palette and per-object id colours (reference :535-575)
        from pyslam.semantics.semantic_mapping_shared import SemanticMappingShared
        from pyslam.utilities.colors import IdsColorTable
        api.sem_img_to_rgb = SemanticMappingShared.sem_img_to_rgb
        api.ids_to_rgb_float = IdsColorTable().ids_to_rgb_float
    except Exception:
        pass
    return make_semantic_integrator_class(B.VolumetricIntegratorBase, api)
