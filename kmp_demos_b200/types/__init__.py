from .point import Point
from .quaternion import quaternion
