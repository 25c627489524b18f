"""Point painting — mirrors lav/utils/point_painting.py and the GPU twin in
team_code_v2/model_inference.py:75-93,255-297 (same names: CoordConverter, point_painting).

The projection + validity test + gather of all cameras is ONE CUDA kernel (csrc/paint.cu).
Camera matrices follow CARLA's Transform.get_matrix (UE4 yaw-pitch-roll, degrees); when the
``carla`` package is importable it is used, otherwise the same matrix is restated here.
"""
import math

import numpy as np
import torch

from . import ops

CAMERA_YAWS = [-60, 0, 60]   # team_code_v2/model_inference.py:12


def _transform_matrix(x, y, z, yaw=0.0, pitch=0.0, roll=0.0):
    try:   # the real thing when running inside the CARLA agent
        import carla
        return np.array(carla.Transform(carla.Location(x, y, z), carla.Rotation(pitch=pitch, yaw=yaw, roll=roll)).get_matrix(),
                        dtype=np.float64)
    except ImportError:
        pass
    cy, sy = math.cos(math.radians(yaw)), math.sin(math.radians(yaw))
    cr, sr = math.cos(math.radians(roll)), math.sin(math.radians(roll))
    cp, sp = math.cos(math.radians(pitch)), math.sin(math.radians(pitch))
    return np.array([[cp * cy, cy * sp * sr - sy * cr, -cy * sp * cr - sy * sr, x],
                     [cp * sy, sy * sp * sr + cy * cr, -sy * sp * cr + cy * sr, y],
                     [sp, -cp * sr, cp * cr, z],
                     [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)


class CoordConverter:
    """Same constructor as point_painting.py:5-25 / model_inference.py:255-278; holds K, lidar_to_world,
    world_to_cam (fp64 numpy, like the reference's numpy class) and the packed fp32 row the kernel reads."""

    def __init__(self, cam_yaw, lidar_xyz=[0, 0, 2.5], cam_xyz=[1.4, 0, 2.5], rgb_h=320, rgb_w=320, fov=60):
        focal = rgb_w / (2.0 * np.tan(fov * np.pi / 360.0))
        K = np.identity(3)
        K[0, 0] = K[1, 1] = focal
        K[0, 2] = rgb_w / 2.0
        K[1, 2] = rgb_h / 2.0
        self.K = K
        self.lidar_to_world = _transform_matrix(*lidar_xyz)
        self.world_to_cam = np.linalg.inv(_transform_matrix(*cam_xyz, yaw=cam_yaw))
        self.rgb_h, self.rgb_w = rgb_h, rgb_w

    def packed(self):
        return np.concatenate([self.K.reshape(-1), self.lidar_to_world.reshape(-1), self.world_to_cam.reshape(-1)]).astype(np.float32)


def make_converters(camera_x=1.5, camera_z=2.4, rgb_h=288, rgb_w=256, fov=64, yaws=CAMERA_YAWS):
    """the converters the agents build (lav_agent_fast.py:131-134)."""
    return [CoordConverter(yaw, lidar_xyz=[0, 0, camera_z], cam_xyz=[camera_x, 0, camera_z], rgb_h=rgb_h, rgb_w=rgb_w, fov=fov)
            for yaw in yaws]


def _cams(coord_converters):
    return np.stack([c.packed() for c in coord_converters])


def _as_sem(sems):
    if isinstance(sems, (list, tuple)):
        sems = torch.stack(list(sems))
    return sems


def point_painting(lidar, sems, coord_converters):
    """point_painting(lidar (N,>=3), sems (ncam,C,H,W), converters) -> painted (N,C) fp32 (CUDA tensors)."""
    sems = _as_sem(sems)
    assert len(sems) == len(coord_converters)
    return ops.paint(lidar.float(), sems.float(), _cams(coord_converters), mode=0)


def forward_paint(cur_lidar, pred_sem, coord_converters, logits=False):
    """InferModel.forward_paint (model_inference.py:44-50): (N,4) + softmaxed (ncam,5,H,W) -> fused (N,8).
    With logits=True the softmax is fused into the gather as well (only the hit pixels are exponentiated)."""
    pred_sem = _as_sem(pred_sem)
    return ops.paint(cur_lidar.float(), pred_sem.float(), _cams(coord_converters), mode=2 if logits else 1,
                     copy_cols=cur_lidar.shape[1])
