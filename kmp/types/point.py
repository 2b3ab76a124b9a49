from kmp_demos_b200.types.point import Point
