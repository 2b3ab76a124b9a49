"""``KMeans`` and ``Uniform`` - drop-ins for the reference's ``kmp.cluster`` (kmp/cluster/Cluster.py).

``KMeans.fit_predict`` reproduces the reference bit for bit (labels are integers, so parity is exact):
distances and centroid means are device kernels that follow numpy's summation orders without FMA
contraction (kmpx_refkmeans_assign / kmpx_refkmeans_update); the initial indices come from the global
numpy RNG exactly as in Cluster.py:88, and the ``np.allclose`` stop test (Cluster.py:97) runs on the K x d
centroid matrix on the host."""
from __future__ import annotations

import logging

import numpy as np

from . import _lib

logging.basicConfig(
    level=logging.INFO,
    format='%(asctime)s.%(msecs)03d %(levelname)s %(message)s',
    datefmt='%Y-%m-%d,%H:%M:%S'
)


class Uniform:
    """Clusterization by even distribution of data points (Cluster.py:14-47): pure index arithmetic."""

    def __init__(self, n_components: int, n_demos: int) -> None:
        if n_components <= 0:
            raise ValueError('n_components must be strictly positive.')
        if n_demos <= 0:
            raise ValueError('n_demos must be strictly positive.')
        self.n_components = n_components
        self.n_demos = n_demos

    def fit_predict(self, X):
        m = int(X.shape[1] / self.n_demos)
        one_demo = (np.arange(m) * self.n_components / m).astype(int)
        return np.tile(one_demo, self.n_demos)


class KMeans:
    """Lloyd's algorithm with the reference's exact arithmetic (Cluster.py:49-102).  X is (n_features, n_samples)."""

    def __init__(self, n_components: int, n_iter: int = 100, tol: float = 1e-4) -> None:
        if n_components <= 0:
            raise ValueError('n_components must be strictly positive.')
        if n_iter <= 0:
            raise ValueError('n_iter must be strictly positive.')
        if tol <= 0:
            raise ValueError('tol must be strictly positive.')
        self.n_components = n_components
        self.n_iter = n_iter
        self.tol = tol
        self.__logger = logging.getLogger(__name__)

    def fit_predict(self, X):
        torch = _lib.require_cuda()
        X = np.ascontiguousarray(X, dtype=np.float64)
        d, n = X.shape
        K = self.n_components
        init_idx = np.random.choice(n, K, replace=False)           # global numpy RNG, Cluster.py:88
        centroids = np.ascontiguousarray(X[:, init_idx].T)
        Xd = torch.as_tensor(X, device='cuda')
        cen = torch.as_tensor(centroids, device='cuda')
        new = torch.empty_like(cen)
        labels = torch.empty(n, dtype=torch.int64, device='cuda')
        st = _lib.stream()
        for iteration in range(self.n_iter):
            _lib.call('kmpx_refkmeans_assign', _lib.ptr(Xd), n, d, K, _lib.ptr(cen), _lib.ptr(labels), st)
            _lib.call('kmpx_refkmeans_update', _lib.ptr(Xd), n, d, K, _lib.ptr(labels), _lib.ptr(new), st)
            new_h = new.cpu().numpy()
            if np.allclose(centroids, new_h, rtol=self.tol):
                self.__logger.info(f"KMeans converged after {iteration+1} iterations")
                break
            centroids = new_h
            cen.copy_(new)
        # labels of the last computed distances, i.e. w.r.t. the pre-update centroids (Cluster.py:102)
        return labels.cpu().numpy()
