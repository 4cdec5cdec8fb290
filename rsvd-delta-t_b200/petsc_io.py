"""PETSc binary readers/writers (big-endian), PETSc-free.

Formats follow what the reference reads and writes through PetscViewerBinaryOpen + MatLoad / MatView /
VecLoad / VecView (LNSStructure.c:22-27, CreateRandomMat.c:23-31, SaveResults.c:32-84,
TransientRunRK4.c:82-86,201-209):

  Mat  : int MAT_FILE_CLASSID (1211216), M, N, nnz; M row lengths; nnz column indices; nnz scalars.
         Dense matrices written by MatView in the default format use the same record with every entry
         present (row by row).  nnz == -1 is PETSc's "native dense" record (column-major values) and is
         accepted on input.
  Vec  : int VEC_FILE_CLASSID (1211214), n; n scalars.

Integers are 32- or 64-bit depending on the PETSc build (the shipped A_GL uses 64-bit); the reader
auto-detects from the class id.  Scalars are complex (re, im doubles) in the reference's build; real
scalar files are detected from the record length.
"""
from __future__ import annotations

import os

import numpy as np

MAT_FILE_CLASSID = 1211216
VEC_FILE_CLASSID = 1211214


def _detect_int(buf: bytes, classid: int):
    if len(buf) >= 4 and int(np.frombuffer(buf[:4], ">i4")[0]) == classid:
        return ">i4", 4
    if len(buf) >= 8 and int(np.frombuffer(buf[:8], ">i8")[0]) == classid:
        return ">i8", 8
    raise ValueError("not a PETSc binary record of the expected class")


def read_mat(path: str):
    """-> (M, N, rowptr[int64], col[int64], val[complex128]).  Trailing bytes (e.g. the MATLAB
    PetscBinaryWrite option records after A_GL) are ignored, as MatLoad does."""
    with open(path, "rb") as f:
        buf = f.read()
    it, isz = _detect_int(buf, MAT_FILE_CLASSID)
    hdr = np.frombuffer(buf[: 4 * isz], it).astype(np.int64)
    M, N, nnz = int(hdr[1]), int(hdr[2]), int(hdr[3])
    off = 4 * isz
    if nnz == -1:  # native dense: column-major scalars
        rest = len(buf) - off
        cplx = rest >= M * N * 16
        if cplx:
            vals = np.frombuffer(buf[off: off + M * N * 16], ">f8").reshape(-1, 2)
            dense = (vals[:, 0] + 1j * vals[:, 1]).reshape(N, M).T
        else:
            dense = np.frombuffer(buf[off: off + M * N * 8], ">f8").astype(np.complex128).reshape(N, M).T
        rowptr = np.arange(M + 1, dtype=np.int64) * N
        col = np.tile(np.arange(N, dtype=np.int64), M)
        return M, N, rowptr, col, np.ascontiguousarray(dense).reshape(-1)
    rl = np.frombuffer(buf[off: off + M * isz], it).astype(np.int64)
    off += M * isz
    col = np.frombuffer(buf[off: off + nnz * isz], it).astype(np.int64)
    off += nnz * isz
    if len(buf) - off >= nnz * 16:
        v = np.frombuffer(buf[off: off + nnz * 16], ">f8").reshape(-1, 2)
        val = v[:, 0] + 1j * v[:, 1]
    else:
        val = np.frombuffer(buf[off: off + nnz * 8], ">f8").astype(np.complex128)
    rowptr = np.zeros(M + 1, np.int64)
    np.cumsum(rl, out=rowptr[1:])
    if rowptr[-1] != nnz:
        raise ValueError("row lengths do not sum to nnz")
    return M, N, rowptr, col, np.ascontiguousarray(val, dtype=np.complex128)


def read_mat_dense(path: str) -> np.ndarray:
    """Any Mat record -> dense (M, N) complex128 in Fortran order."""
    M, N, rowptr, col, val = read_mat(path)
    D = np.zeros((M, N), np.complex128, order="F")
    rows = np.repeat(np.arange(M), np.diff(rowptr))
    D[rows, col] = val
    return D


def _ints(a, int64: bool):
    return np.asarray(a).astype(">i8" if int64 else ">i4").tobytes()


def _scalars(v):
    v = np.asarray(v, dtype=np.complex128).reshape(-1)
    out = np.empty((v.size, 2), ">f8")
    out[:, 0] = v.real
    out[:, 1] = v.imag
    return out.tobytes()


def write_mat_aij(path: str, M: int, N: int, rowptr, col, val, int64: bool = True) -> None:
    rowptr = np.asarray(rowptr, np.int64)
    nnz = int(rowptr[-1])
    with open(path, "wb") as f:
        f.write(_ints([MAT_FILE_CLASSID, M, N, nnz], int64))
        f.write(_ints(np.diff(rowptr), int64))
        f.write(_ints(np.asarray(col)[:nnz], int64))
        f.write(_scalars(np.asarray(val)[:nnz]))


def write_mat_dense(path: str, D, int64: bool = True) -> None:
    """What MatView(dense, binary viewer) produces in the default format: an AIJ record with all M*N entries."""
    D = np.asarray(D, dtype=np.complex128)
    M, N = D.shape
    with open(path, "wb") as f:
        f.write(_ints([MAT_FILE_CLASSID, M, N, M * N], int64))
        f.write(_ints(np.full(M, N), int64))
        f.write(_ints(np.tile(np.arange(N), M), int64))
        f.write(_scalars(np.ascontiguousarray(D).reshape(-1)))  # row by row


def read_vec(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        buf = f.read()
    it, isz = _detect_int(buf, VEC_FILE_CLASSID)
    n = int(np.frombuffer(buf[isz: 2 * isz], it)[0])
    off = 2 * isz
    if len(buf) - off >= n * 16:
        v = np.frombuffer(buf[off: off + n * 16], ">f8").reshape(-1, 2)
        return v[:, 0] + 1j * v[:, 1]
    return np.frombuffer(buf[off: off + n * 8], ">f8").astype(np.complex128)


def write_vec(path: str, v, int64: bool = True) -> None:
    v = np.asarray(v, dtype=np.complex128).reshape(-1)
    with open(path, "wb") as f:
        f.write(_ints([VEC_FILE_CLASSID, v.size], int64))
        f.write(_scalars(v))


def file_exists(path: str) -> bool:
    return os.path.exists(path)
