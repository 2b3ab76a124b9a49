"""Host I/O helpers with the reference's names (kmp/utils.py): MATLAB demo reader, nearest-index search and
the UR5e CSV -> Point dataset builder.  No device work happens here except the quaternion maps used by
create_dataset, which go through kmpx_quat_* like everything else."""
import csv
import math
import os

import numpy as np

from .types.point import Point
from .types.quaternion import quaternion


def read_struct(filepath, cell_array_name='demos', fields=['pos', 'vel', 'acc'], max_cell=math.inf):
    """1 x n MATLAB cell array of structs -> list of dicts of arrays (kmp/utils.py:13-28)."""
    import scipy.io as sio
    cells = sio.loadmat(filepath)[cell_array_name][0]
    out = []
    for idx, cell in enumerate(cells):
        if idx >= max_cell:
            break
        out.append({name: cell[name][0][0] for name in fields})
    return out


def find_closest_index(list, val):
    """Index of the column of ``list`` closest to ``val`` (first minimum), kmp/utils.py:31-38."""
    cols = np.asarray(list).T
    dists = np.array([np.linalg.norm(val - c) for c in cols])
    dists = np.where(np.isnan(dists), np.inf, dists)      # `dist < smallest_dist` never accepts a NaN (utils.py:34)
    return int(np.argmin(dists))


def create_dataset(path, subsample=100):
    """UR5e recordings (space-separated CSV) -> list of Point, saved as pose_data.npy (kmp/utils.py:41-81),
    including the sign-continuity rule for the stored quaternion and the tangent-space projection."""
    cols, out = [], []
    prev_quat, qa = None, None
    dt = 0.01 / subsample
    sign = 1
    files = [f for f in os.listdir(path) if f != 'pose_data.npy' and os.path.isfile(os.path.join(path, f))]
    wanted = ['actual_TCP_pose_', 'actual_TCP_speed_', 'actual_TCP_force_']
    for name in files:
        with open(os.path.join(path, name)) as fh:
            for j, row in enumerate(csv.reader(fh, delimiter=' ')):
                if j % subsample != 0 or j > 7500:
                    continue
                if len(cols) == 0:
                    cols = [i for i, val in enumerate(row) if any(s in val for s in wanted)]
                elif j != 0:
                    row = np.array(row, dtype=float)
                    pose, twist = row[cols[:6]], row[cols[6:12]]
                    quat = quaternion.from_rotation_vector(row[cols[3:6]])
                    if qa is None:
                        qa = quat
                    if prev_quat is not None:
                        max_prev = np.argmax(quat.abs())
                        max_curr = np.argmax(prev_quat.abs())
                        if max_prev == max_curr and np.sign(quat[max_curr]) != np.sign(prev_quat[max_prev]):
                            sign = -sign
                    quat_eucl = (quat * ~qa).log()
                    out.append(Point(j * dt, pose, twist, -quat * sign, quat_eucl, row[cols[-6:]]))
                    prev_quat = quat
    np.save(os.path.join(path + "pose_data.npy"), out)
    return out
