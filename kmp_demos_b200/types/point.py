"""``Point`` record of the demonstration database (kmp/types/point.py:7-32).  Instances are pickled inside
``quaternion_trajectories/pose_data.npy`` under the name ``kmp.types.point.Point``; the field defaults use
``default_factory`` so that the class also imports on Python >= 3.11."""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Point:
    time: float = 0.0
    pose: np.ndarray = field(default_factory=lambda: np.zeros(6))
    twist: np.ndarray = field(default_factory=lambda: np.zeros(6))
    quat: np.ndarray = field(default_factory=lambda: np.zeros(4))
    quat_eucl: np.ndarray = field(default_factory=lambda: np.zeros(3))
    wrench: np.ndarray = field(default_factory=lambda: np.zeros(6))
