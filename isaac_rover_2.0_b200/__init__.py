"""B200-native drop-in for the isaac_rover_2.0 rover task's per-step hot path.

Mirrors the reference's call signatures; every numeric op runs in hand-written sm_100a
kernels behind the C ABI of include/rover_b200.h (librover_b200.so, bound with ctypes).  No CPU path.
"""
from . import _lib, dist, hooks, model, synth
from .hooks import ObsHooks, TeacherRecorder
from ._lib import SEM_TORCH_CPU, SEM_TORCH_CUDA
from .camera import Camera, cast_rays
from .heightmap_distribution import Heightmap
from .host import HostPipeline
from .kinematics import Ackermann
from .ray_casting import ray_distance
from .rock_detect import Rock_Detection
from .rover import Memory, RoverTask, STAT_NAMES
from .tensor_quat_to_euler import tensor_quat_to_eul
from .terrain import TerrainLayer, build_knn_index
from .terrain_utils import read_stone_info, stone_info_from_array

__all__ = ["Camera", "Heightmap", "Ackermann", "ray_distance", "Rock_Detection", "Memory", "RoverTask",
           "tensor_quat_to_eul", "TerrainLayer", "build_knn_index", "read_stone_info", "stone_info_from_array",
           "cast_rays", "HostPipeline", "synth", "model", "hooks", "ObsHooks", "TeacherRecorder", "SEM_TORCH_CPU", "SEM_TORCH_CUDA", "STAT_NAMES"]
