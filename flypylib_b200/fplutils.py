"""Drop-in for the two helpers of flypylib/fplutils.py that sit on the detection path
(to3d :9-12, set_filter :14-22).  Host-side only; the device kernels use the same footprint
``dz^2+dy^2+dx^2 <= r^2`` in integer arithmetic."""
from collections import namedtuple

import numpy as np

szyx = namedtuple('szyx', 'size z y x')     # flypylib/fplutils.py:6-7


def to3d(vv):
    """Scalar -> 3-tuple, anything else unchanged (flypylib/fplutils.py:9-12)."""
    if np.size(vv) == 1:
        vv = (vv, vv, vv)
    return vv


def set_filter(radius, return_dist=False):
    """Boolean ball of shape (2r+1,)*3, True where the Euclidean distance to the centre is
    <= radius (flypylib/fplutils.py:14-22)."""
    ax = np.arange(-radius, radius + 1)
    dd = np.sqrt(ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2)
    if return_dist:
        return dd <= radius, dd
    return dd <= radius


def roi_from_txt(filename):
    """Substack list of an ROI text file, one ``size,z,y,x`` line per substack, wrapped in a one-element
    list like DVID's ``get_roi_partition`` result (flypylib/fplutils.py:24-30, fplobjdetect.py:1210-1216)."""
    with open(filename, 'r') as f_in:
        lines = f_in.read().splitlines()
    return [[szyx(*[int(nn) for nn in ss.split(',')]) for ss in lines if ss.strip()], ]
