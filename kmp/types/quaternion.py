from kmp_demos_b200.types.quaternion import quaternion
