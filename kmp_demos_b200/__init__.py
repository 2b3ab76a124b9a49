"""kmp_demos_b200 - B200 (sm_100a) implementation of the KMP learn-and-adapt hot path of lbusellato/KMP_demos
behind the reference's own Python API.  Numerical work happens in libkmpx.so (csrc/, C ABI in include/kmpx.h);
importing the package needs neither CUDA nor the library, using it does (no CPU fallback)."""

__all__ = ['KMP', 'KMPBatch', 'GaussianMixtureModel', 'KMeans', 'Uniform', 'quaternion', 'Point']


def __getattr__(name):
    if name == 'KMP':
        from .model import KMP
        return KMP
    if name == 'KMPBatch':
        from .batch import KMPBatch
        return KMPBatch
    if name == 'GaussianMixtureModel':
        from .mixture import GaussianMixtureModel
        return GaussianMixtureModel
    if name in ('KMeans', 'Uniform'):
        from . import cluster
        return getattr(cluster, name)
    if name == 'quaternion':
        from .types.quaternion import quaternion
        return quaternion
    if name == 'Point':
        from .types.point import Point
        return Point
    raise AttributeError(name)
