"""Gadget-2 (format 1) snapshots through lib/libp2p_host.so (include/p2p_host.h: p2p_snapshot_*), the reader / writer
the device-resident stepping needs to start from and return to the reference's files
(1_Indexing/src/snapshot.c:211-293,397-503; a rank reads the slab [N r / P, N (r + 1) / P) as 1_Indexing/src/initial.c:648-692 does)."""
import ctypes as C

import numpy as np

from .host import load_host_library

_dp = C.POINTER(C.c_double)


class SnapshotInfo(C.Structure):
    _fields_ = [("npart", C.c_int64 * 6), ("nfile", C.c_int64), ("npart_total", C.c_uint32 * 6), ("mass", C.c_double * 6),
                ("time", C.c_double), ("redshift", C.c_double), ("box", C.c_double), ("omega0", C.c_double),
                ("omega_lambda", C.c_double), ("hubble", C.c_double), ("num_files", C.c_int)]


def _lib():
    L = load_host_library()
    L.p2p_snapshot_header.argtypes = [C.c_char_p, C.POINTER(SnapshotInfo)]
    L.p2p_snapshot_read.argtypes = [C.c_char_p, C.c_int64, C.c_int64, _dp, C.c_int64, _dp, C.c_int64]
    L.p2p_snapshot_write.argtypes = [C.c_char_p, C.POINTER(SnapshotInfo), C.c_int64, _dp, C.c_int64, _dp, C.c_int64]
    return L


def header(path):
    info = SnapshotInfo()
    rc = _lib().p2p_snapshot_header(str(path).encode(), C.byref(info))
    if rc:
        raise IOError(f"cannot read the Gadget-2 header of {path} ({rc})")
    return info


def read(path, n_start=0, n_count=None, velocities=True):
    """(positions [n, 3] float64, velocities [n, 3] float64 or None, header) of the particles [n_start, n_start + n_count)"""
    info = header(path)
    n = info.nfile - n_start if n_count is None else n_count
    pos = np.empty((n, 3), np.float64)
    vel = np.empty((n, 3), np.float64) if velocities else None
    rc = _lib().p2p_snapshot_read(str(path).encode(), int(n_start), int(n), pos.ctypes.data_as(_dp), 3,
                                  vel.ctypes.data_as(_dp) if velocities else None, 3)
    if rc:
        raise IOError(f"cannot read particles [{n_start}, {n_start + n}) of {path} ({rc})")
    return pos, vel, info


def write(path, pos, vel, box, mass, redshift, npart_total=None, omega0=0.0, omega_lambda=0.0, hubble=0.0):
    pos = np.ascontiguousarray(pos, np.float64)
    vel = np.ascontiguousarray(vel, np.float64) if vel is not None else None
    info = SnapshotInfo()
    info.mass[1] = mass
    info.npart_total[1] = int(npart_total if npart_total is not None else len(pos))
    info.box, info.redshift, info.omega0, info.omega_lambda, info.hubble = box, redshift, omega0, omega_lambda, hubble
    rc = _lib().p2p_snapshot_write(str(path).encode(), C.byref(info), len(pos), pos.ctypes.data_as(_dp), 3,
                                   vel.ctypes.data_as(_dp) if vel is not None else None, 3)
    if rc:
        raise IOError(f"cannot write {path} ({rc})")
