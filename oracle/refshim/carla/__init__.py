"""Stand-in for the CARLA 0.9.10.1 PythonAPI types the painting code touches
(Transform/Location/Rotation.get_matrix) so the REFERENCE painter can be imported by
oracle/pin_against_reference.py.  Test infrastructure only; restates UE4 yaw-pitch-roll."""
import math
import numpy as np


class Location:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class Rotation:
    def __init__(self, pitch=0.0, yaw=0.0, roll=0.0):
        self.pitch, self.yaw, self.roll = float(pitch), float(yaw), float(roll)


class Transform:
    def __init__(self, location=None, rotation=None):
        self.location = location or Location()
        self.rotation = rotation or Rotation()

    def get_matrix(self):
        r, l = self.rotation, self.location
        cy, sy = math.cos(math.radians(r.yaw)), math.sin(math.radians(r.yaw))
        cr, sr = math.cos(math.radians(r.roll)), math.sin(math.radians(r.roll))
        cp, sp = math.cos(math.radians(r.pitch)), math.sin(math.radians(r.pitch))
        return [[cp * cy, cy * sp * sr - sy * cr, -cy * sp * cr - sy * sr, l.x],
                [cp * sy, sy * sp * sr + cy * cr, -sy * sp * cr + cy * sr, l.y],
                [sp, -cp * sr, cp * cr, l.z],
                [0.0, 0.0, 0.0, 1.0]]

    def get_inverse_matrix(self):
        return np.linalg.inv(np.array(self.get_matrix())).tolist()


class VehicleControl:
    def __init__(self, steer=0.0, throttle=0.0, brake=0.0):
        self.steer, self.throttle, self.brake = steer, throttle, brake
