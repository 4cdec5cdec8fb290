"""PETSc binary readers/writers (big-endian), PETSc-free.

Formats follow what the reference reads and writes through PetscViewerBinaryOpen + MatLoad / MatView /
VecLoad / VecView (LNSStructure.c:22-27, CreateRandomMat.c:23-31, SaveResults.c:32-84,
TransientRunRK4.c:82-86,201-209):

  Mat  : int MAT_FILE_CLASSID (1211216), M, N, nnz; M row lengths; nnz column indices; nnz scalars.
         Dense matrices written by MatView in the default format use the same record with every entry
         present (row by row).  nnz == -1 is PETSc's "native dense" record (values row by row, as MatView_Dense_Binary of PETSc >= 3.14 writes them) and is
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
    if nnz == -1:  # native dense: M*N scalars, row by row (MatLoad_Dense_Binary, PETSc 3.19)
        rest = len(buf) - off
        cplx = rest >= M * N * 16
        if cplx:
            vals = np.frombuffer(buf[off: off + M * N * 16], ">f8").reshape(-1, 2)
            dense = (vals[:, 0] + 1j * vals[:, 1]).reshape(M, N)
        else:
            dense = np.frombuffer(buf[off: off + M * N * 8], ">f8").astype(np.complex128).reshape(M, N)
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


# ----------------------------------------------------------------------------------------------------------------------
# Partitioned I/O: every rank reads / writes only its rows of a record ("next" row 4 of SURVEY.md section 8f: MatLoad of a
# MATMPIAIJ gives each rank its PETSC_DECIDE row range, LNSStructure.c:22-27; MatView / VecView collect them in row order).
# ----------------------------------------------------------------------------------------------------------------------
class MatRowReader:
    """Row-range reader of a binary Mat record.  Reads the header and the M row lengths once (8*M bytes), then any
    row range [r0, r1) costs two seeks and reads only its own columns and values."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(8)
            self.it, self.isz = _detect_int(head, MAT_FILE_CLASSID)
            f.seek(0)
            hdr = np.frombuffer(f.read(4 * self.isz), self.it).astype(np.int64)
            self.M, self.N, self.nnz = int(hdr[1]), int(hdr[2]), int(hdr[3])
            self.native_dense = self.nnz == -1
            size = os.fstat(f.fileno()).st_size
            if self.native_dense:
                self.rowptr = np.arange(self.M + 1, dtype=np.int64) * self.N
                self.off_val = 4 * self.isz
                self.cplx = size - self.off_val >= self.M * self.N * 16
            else:
                rl = np.frombuffer(f.read(self.M * self.isz), self.it).astype(np.int64)
                self.rowptr = np.zeros(self.M + 1, np.int64)
                np.cumsum(rl, out=self.rowptr[1:])
                if self.rowptr[-1] != self.nnz:
                    raise ValueError("row lengths do not sum to nnz")
                self.off_col = (4 + self.M) * self.isz
                self.off_val = self.off_col + self.nnz * self.isz
                self.cplx = size - self.off_val >= self.nnz * 16

    def _scalars(self, f, first, count):
        w = 16 if self.cplx else 8
        f.seek(self.off_val + first * w)
        raw = np.frombuffer(f.read(count * w), ">f8")
        if self.cplx:
            raw = raw.reshape(-1, 2)
            return raw[:, 0] + 1j * raw[:, 1]
        return raw.astype(np.complex128)

    def rows(self, r0: int, r1: int):
        """-> (rowptr [r1-r0+1, starting at 0], global columns, values) of rows [r0, r1)"""
        r0, r1 = max(0, int(r0)), min(self.M, int(r1))
        p0, p1 = int(self.rowptr[r0]), int(self.rowptr[r1])
        rp = (self.rowptr[r0:r1 + 1] - p0).astype(np.int64)
        with open(self.path, "rb") as f:
            if self.native_dense:   # values row by row: the rows asked for are one contiguous range
                out = self._scalars(f, r0 * self.N, (r1 - r0) * self.N)
                return rp, np.tile(np.arange(self.N, dtype=np.int64), r1 - r0), np.ascontiguousarray(out, dtype=np.complex128)
            f.seek(self.off_col + p0 * self.isz)
            col = np.frombuffer(f.read((p1 - p0) * self.isz), self.it).astype(np.int64)
            val = self._scalars(f, p0, p1 - p0)
        return rp, col, np.ascontiguousarray(val, dtype=np.complex128)

    def dense_rows(self, r0: int, r1: int) -> np.ndarray:
        """rows [r0, r1) as a dense (r1-r0, N) complex128 Fortran array (for F_hat files, CreateRandomMat.c:23-31)"""
        rp, col, val = self.rows(r0, r1)
        D = np.zeros((r1 - r0, self.N), np.complex128, order="F")
        D[np.repeat(np.arange(r1 - r0), np.diff(rp)), col] = val
        return D


def read_vec_rows(path: str, r0: int, r1: int):
    """-> (n_global, entries [r0, r1)) of a binary Vec record"""
    with open(path, "rb") as f:
        it, isz = _detect_int(f.read(8), VEC_FILE_CLASSID)
        f.seek(isz)
        n = int(np.frombuffer(f.read(isz), it)[0])
        size = os.fstat(f.fileno()).st_size
        cplx = size - 2 * isz >= n * 16
        w = 16 if cplx else 8
        f.seek(2 * isz + r0 * w)
        raw = np.frombuffer(f.read((r1 - r0) * w), ">f8")
    if cplx:
        raw = raw.reshape(-1, 2)
        return n, raw[:, 0] + 1j * raw[:, 1]
    return n, raw.astype(np.complex128)


class DenseMatWriter:
    """Collective writer of the record MatView(dense) produces (an AIJ record holding every entry, row by row).
    Rank 0 calls create() (header + file size), then -- after a barrier -- every rank calls write_rows() with its own
    row block; the three sections of a row block (row lengths, column indices, values) are contiguous slabs."""

    def __init__(self, path: str, M: int, N: int, int64: bool = True):
        self.path, self.M, self.N = path, int(M), int(N)
        self.it, self.isz = (">i8", 8) if int64 else (">i4", 4)
        self.off_rl = 4 * self.isz
        self.off_col = self.off_rl + self.M * self.isz
        self.off_val = self.off_col + self.M * self.N * self.isz
        self.size = self.off_val + self.M * self.N * 16

    def create(self):
        with open(self.path, "wb") as f:
            f.write(np.asarray([MAT_FILE_CLASSID, self.M, self.N, self.M * self.N]).astype(self.it).tobytes())
            f.truncate(self.size)

    def write_rows(self, r0: int, block) -> None:
        block = np.asarray(block, dtype=np.complex128)
        n = block.shape[0]
        assert block.ndim == 2 and block.shape[1] == self.N and 0 <= r0 and r0 + n <= self.M
        if n == 0:
            return
        with open(self.path, "r+b") as f:
            f.seek(self.off_rl + r0 * self.isz)
            f.write(np.full(n, self.N).astype(self.it).tobytes())
            f.seek(self.off_col + r0 * self.N * self.isz)
            f.write(np.tile(np.arange(self.N), n).astype(self.it).tobytes())
            f.seek(self.off_val + r0 * self.N * 16)
            f.write(_scalars(np.ascontiguousarray(block).reshape(-1)))


class VecWriter:
    """Collective VecView: rank 0 create(), barrier, every rank write_rows(r0, values)."""

    def __init__(self, path: str, n: int, int64: bool = True):
        self.path, self.n = path, int(n)
        self.it, self.isz = (">i8", 8) if int64 else (">i4", 4)

    def create(self):
        with open(self.path, "wb") as f:
            f.write(np.asarray([VEC_FILE_CLASSID, self.n]).astype(self.it).tobytes())
            f.truncate(2 * self.isz + self.n * 16)

    def write_rows(self, r0: int, v) -> None:
        v = np.asarray(v, dtype=np.complex128).reshape(-1)
        if v.size == 0:
            return
        with open(self.path, "r+b") as f:
            f.seek(2 * self.isz + r0 * 16)
            f.write(_scalars(v))


class AijStreamWriter:
    """Writes a binary AIJ Mat record from row blocks that arrive in order, without holding the matrix: the three sections
    (row lengths, column indices, values) are streamed to side files and concatenated behind the header on close()."""

    def __init__(self, path: str, M: int, N: int, int64: bool = True):
        self.path, self.M, self.N = path, int(M), int(N)
        self.it = ">i8" if int64 else ">i4"
        self.parts = [open(path + ext, "wb") for ext in (".rl", ".ci", ".va")]
        self.rows = 0
        self.nnz = 0

    def append_rows(self, rowptr, col, val) -> None:
        rowptr = np.asarray(rowptr, np.int64)
        n = int(rowptr[-1] - rowptr[0])
        self.parts[0].write(np.diff(rowptr).astype(self.it).tobytes())
        self.parts[1].write(np.asarray(col)[:n].astype(self.it).tobytes())
        self.parts[2].write(_scalars(np.asarray(val)[:n]))
        self.rows += rowptr.size - 1
        self.nnz += n

    def close(self) -> None:
        for f in self.parts:
            f.close()
        if self.rows != self.M:
            raise ValueError(f"{self.rows} rows written, {self.M} announced")
        with open(self.path, "wb") as out:
            out.write(np.asarray([MAT_FILE_CLASSID, self.M, self.N, self.nnz]).astype(self.it).tobytes())
            for ext in (".rl", ".ci", ".va"):
                with open(self.path + ext, "rb") as f:
                    while True:
                        buf = f.read(1 << 26)
                        if not buf:
                            break
                        out.write(buf)
                os.remove(self.path + ext)
