"""Headless ``demo1`` / ``demo2`` / ``demo3``.

The reference's demos (kmp/demos.py) are matplotlib front-ends: each constructor runs a learn-and-adapt pipeline
(GMM fit -> GMR -> KMP fit -> predict -> via points -> predict) and then draws it.  The classes here run the same
pipelines with the same parameters on the device path of this package and keep every intermediate array as an
attribute (``results()`` returns them as a dict, ``save()`` writes an .npz) instead of plotting, so that
``from kmp import demo1, demo2, demo3`` (main.py:12) works on a machine without a display.

Data files are looked up like the reference does - relative to the working directory (``2Dletters/<L>.mat``,
``quaternion_trajectories/pose_data.npy``) - unless ``root`` is given.
"""
from __future__ import annotations

import logging
import os

import numpy as np

from . import utils
from .mixture import GaussianMixtureModel
from .model import KMP
from .types import quaternion as _quat
from .types.quaternion import quaternion

LETTERS = tuple('ABCDEFGHIJKLMNOPQRSTUVWXYZ')


class _Headless:
    """Common part: a result dictionary instead of figures."""

    _fields: tuple = ()

    def results(self):
        return {name: np.asarray(getattr(self, name)) for name in self._fields if getattr(self, name, None) is not None}

    def save(self, path):
        np.savez(path, **self.results())
        return path


class demo1(_Headless):
    """Position-only learning and adaptation on one letter of the 2Dletters set (demos.py:13-170).

    ``update`` is the reference's "Update" button (letter and number of components are its two text boxes),
    ``add_waypoint(x, y)`` its ctrl-click on the KMP axes."""

    _fields = ('mu', 'sigma', 'mu_kmp', 'sigma_kmp', 'waypoints')

    def __init__(self, letter='G', C=8, root=None):
        self._log = logging.getLogger(__name__)
        self.root = os.getcwd() if root is None else root
        self.letter, self.C = letter, C
        self.waypoints = None
        self.update()

    def update(self, letter=None, C=None):
        self.letter = str(self.letter if letter is None else letter).upper()
        self.C = int(self.C if C is None else C)
        if self.letter not in LETTERS:
            raise ValueError(f'unknown letter {self.letter!r}')
        self.waypoints = None
        self.H = 5                                      # demonstrations used
        self.demos = utils.read_struct(os.path.join(self.root, '2Dletters', self.letter + '.mat'), max_cell=self.H)
        pos = np.concatenate([d['pos'] for d in self.demos], axis=1)[:, ::4]       # every fourth sample
        self.N = pos.shape[1]
        T = self.N // self.H
        self.kmp_dt = 0.01
        stamps = 0.01 * np.tile(np.linspace(1, T, T), self.H).reshape(1, -1)
        self.gmm = GaussianMixtureModel(n_components=self.C, diag_reg_factor=1e-6)
        self.gmm.fit(np.vstack((stamps, pos)).T)
        self.mu, self.sigma = self.gmm.predict(0.01 * np.arange(T).reshape(1, -1))
        self.time = self.kmp_dt * np.arange(1, T + 1).reshape(1, -1)
        self.kmp = KMP(l=0.5, alpha=40, sigma_f=200, time_driven_kernel=False)
        self.kmp.fit(self.time, self.mu, self.sigma)
        self.mu_kmp, self.sigma_kmp = self.kmp.predict(self.time)
        return self

    def add_waypoint(self, x, y):
        """Insert the via point (x, y) at the time of the closest point of the GMR mean (demos.py:91-112)."""
        p = np.array([float(x), float(y)]).reshape(1, -1)
        t = self.kmp_dt * (utils.find_closest_index(self.mu[:2, :], p) + 1)
        self._log.info('Inserting new point x=%.2f, y=%.2f, t=%.2f' % (x, y, t))
        self.kmp.set_waypoint([t], [p], [np.eye(2) * 1e-4])
        self.waypoints = p if self.waypoints is None else np.vstack((self.waypoints, p))
        self.mu_kmp, self.sigma_kmp = self.kmp.predict(self.time)
        return t


class demo2(_Headless):
    """Position + orientation (tangent space of the first quaternion) learning on the UR5e recordings, one via point
    for each (demos.py:172-330)."""

    _fields = ('time_single', 'mu_pos', 'sigma_pos', 'mu_pos_kmp', 'sigma_pos_kmp', 'mu_pos_kmp_ad', 'sigma_pos_kmp_ad',
               'mu_quat', 'sigma_quat', 'mu_quat_kmp', 'sigma_quat_kmp', 'mu_quat_kmp_ad', 'sigma_quat_kmp_ad',
               'kmp_quats', 'kmp_quats_ad', 'waypoint_pos', 'waypoint_quat', 'des_quat', 'qa')

    def __init__(self, root=None):
        self._log = logging.getLogger(__name__)
        root = os.getcwd() if root is None else root
        folder = os.path.join(root, 'quaternion_trajectories')
        dataset = os.path.join(folder, 'pose_data.npy')
        self.demos = np.load(dataset, allow_pickle=True) if os.path.isfile(dataset) else utils.create_dataset(folder + os.sep, subsample=50)
        demo_num = 6
        demo_len = len(self.demos) // demo_num
        stamps = np.array([s.time for s in self.demos]).reshape(1, -1)
        self.time_single = stamps[:, :demo_len]
        new_kmp = lambda: KMP(l=0.5, alpha=60, sigma_f=4, time_driven_kernel=False)

        def learn(Y):
            gmm = GaussianMixtureModel(n_demos=demo_num)
            gmm.fit(np.vstack((stamps, Y)).T)
            mu, sigma = gmm.predict(self.time_single)
            kmp = new_kmp()
            kmp.fit(self.time_single, mu, sigma)
            return gmm, mu, sigma, kmp, kmp.predict(self.time_single)

        def adapt(mu, sigma, waypoint):
            kmp = new_kmp()
            kmp.fit(self.time_single, mu, sigma)
            kmp.set_waypoint([1], np.asarray(waypoint, float).reshape(1, -1), np.eye(3) * 1e-6)     # un-listed, as upstream
            return kmp, kmp.predict(self.time_single)

        # position
        self.gmm, self.mu_pos, self.sigma_pos, self.kmp_pos, (self.mu_pos_kmp, self.sigma_pos_kmp) = \
            learn(np.vstack([s.pose[:3] for s in self.demos]).T)
        self.waypoint_pos = np.array([0.0, 0.85, 0.3])
        self.kmp_pos_ad, (self.mu_pos_kmp_ad, self.sigma_pos_kmp_ad) = adapt(self.mu_pos, self.sigma_pos, self.waypoint_pos)
        # orientation in the tangent space at qa, mapped back with exp(.) * qa
        self.gmm_quat, self.mu_quat, self.sigma_quat, self.kmp_quat, (self.mu_quat_kmp, self.sigma_quat_kmp) = \
            learn(np.vstack([s.quat_eucl for s in self.demos]).T)
        qa = self.demos[0].quat
        qa = qa if isinstance(qa, quaternion) else quaternion.from_array(np.asarray(qa, float))
        self.qa = qa.as_array()
        self.kmp_quats = self._to_quaternions(self.mu_quat_kmp)
        des = quaternion(-0.2, np.array([0.8, -0.7, -1.6]))
        self.des_quat = des.as_array()
        self.waypoint_quat = (des * (~qa)).log()
        self.kmp_quat_ad, (self.mu_quat_kmp_ad, self.sigma_quat_kmp_ad) = adapt(self.mu_quat, self.sigma_quat, self.waypoint_quat)
        self.kmp_quats_ad = self._to_quaternions(self.mu_quat_kmp_ad)

    def _to_quaternions(self, mu):
        """(3, M) tangent vectors -> (4, M) unit quaternions exp(w) * qa, scalar first (demos.py:262-266), batched on the device."""
        return _quat.mul_batch(_quat.exp_batch(np.ascontiguousarray(mu[:3].T)), self.qa.reshape(1, 4)).T


class demo3(_Headless):
    """Time-driven kernel on position + velocity of the letter G with four via points (demos.py:332-390)."""

    _fields = ('mu', 'sigma', 'time', 'via_t', 'via_p', 'mu_kmp', 'sigma_kmp')

    def __init__(self, root=None, letter='G'):
        self._log = logging.getLogger(__name__)
        root = os.getcwd() if root is None else root
        H = 5
        self.demos = utils.read_struct(os.path.join(root, '2Dletters', letter + '.mat'), max_cell=H)
        pos = np.concatenate([d['pos'] for d in self.demos], axis=1)
        vel = np.concatenate([d['vel'] for d in self.demos], axis=1)
        T = pos.shape[1] // H
        stamps = 0.01 * np.tile(np.linspace(1, T, T), H).reshape(1, -1)
        self.gmm = GaussianMixtureModel(n_demos=H)
        self.gmm.fit(np.vstack((stamps, pos, vel)).T)
        self.mu, self.sigma = self.gmm.predict(0.01 * np.arange(T).reshape(1, -1))
        self.kmp_dt = 0.01
        self.time = self.kmp_dt * np.arange(1, T + 1).reshape(1, -1)
        self.kmp = KMP(l=1, sigma_f=6)
        self.via_t = [0.01, 0.25, 1.2, 2]
        self.via_p = [np.array(v, float).reshape(1, -1) for v in ([8, 10, -50, 0], [-1, 6, -25, -40], [8, -4, 30, 10], [-3, 1, -10, 3])]
        self.kmp.fit(self.time, self.mu, self.sigma)
        self.kmp.set_waypoint(self.via_t, self.via_p, [np.eye(4) * 1e-6] * 4)
        self.mu_kmp, self.sigma_kmp = self.kmp.predict(self.time)
        self.kmp_pos = self.mu_kmp                       # the attribute the reference ends up with (demos.py:367)


def main(argv=None):
    """``python -m kmp_demos_b200.demos {1,2,3} [--root DIR] [--out FILE.npz]``: main.py without the GUI."""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument('demo', choices=['1', '2', '3'])
    ap.add_argument('--root', default=None)
    ap.add_argument('--out', default=None)
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s.%(msecs)03d %(levelname)s %(message)s', datefmt='%Y-%m-%d,%H:%M:%S')
    d = {'1': demo1, '2': demo2, '3': demo3}[a.demo](root=a.root)
    for name, v in d.results().items():
        print(f'{name}: shape {v.shape}')
    if a.out:
        print('saved', d.save(a.out))
    return d


if __name__ == '__main__':
    main()
