from .point import Point
