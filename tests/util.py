import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    if nb == 0.0:
        return float(np.linalg.norm(a))
    return float(np.linalg.norm(a - b) / nb)


def local_pts(mesh, arr, comps):
    """slice of a per-point array that belongs to local cells"""
    off = mesh.vs_off()
    return arr[: off[mesh.n_local] * comps]
