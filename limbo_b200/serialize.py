"""serialize::TextArchive / BinaryArchive of the reference (src/limbo/serialize/text_archive.hpp:63-151,
binary_archive.hpp:63-162): one file per object in a directory.  Text: rows separated by new lines, values by single
spaces, full precision.  Binary: Eigen::Index rows, cols (int64 on the LP64 targets limbo builds on) followed by the
column-major doubles; a vector list is an int32 count followed by the vectors.  GPs written by the reference load here and
vice versa (same six objects as GP::save, model/gp.hpp:448-460)."""
from __future__ import annotations

import os

import numpy as np


class TextArchive:
    def __init__(self, dir_name: str):
        self._dir_name = dir_name

    def directory(self) -> str:
        return self._dir_name

    def fname(self, object_name: str) -> str:  # text_archive.hpp:114-117
        return os.path.join(self._dir_name, object_name + ".dat")

    def save(self, v, object_name: str) -> None:
        """A matrix (2-D), a vector (written as a column, like Eigen) or a list of vectors (one per line)."""
        os.makedirs(self._dir_name, exist_ok=True)
        if isinstance(v, (list, tuple)):
            rows = [np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in v]
        else:
            m = np.asarray(v, dtype=np.float64)
            rows = list(m.reshape(-1, 1)) if m.ndim == 1 else list(m)
        with open(self.fname(object_name), "w") as f:
            for r in rows:
                f.write(" ".join(repr(float(x)) for x in r) + "\n")

    def _load(self, object_name: str):
        path = self.fname(object_name)
        assert os.path.exists(path), "file not found"
        rows = []
        with open(path) as f:
            for line in f:
                line = line.rstrip("\n")
                if line == "":
                    continue
                rows.append([float(c) for c in line.split(" ") if c != ""])
        assert rows, "empty file"
        return rows

    def load_matrix(self, object_name: str) -> np.ndarray:
        return np.array(self._load(object_name), dtype=np.float64)

    def load_vector(self, object_name: str) -> np.ndarray:
        return self.load_matrix(object_name).reshape(-1)

    def load_vector_list(self, object_name: str):
        return [np.array(r, dtype=np.float64) for r in self._load(object_name)]


class BinaryArchive:
    """binary_archive.hpp:63-162; same interface as TextArchive."""

    def __init__(self, dir_name: str):
        self._dir_name = dir_name

    def directory(self) -> str:
        return self._dir_name

    def fname(self, object_name: str) -> str:  # binary_archive.hpp:125-128
        return os.path.join(self._dir_name, object_name + ".bin")

    @staticmethod
    def _matrix_bytes(m: np.ndarray) -> bytes:  # binary_archive.hpp:144-150 (_write_binary)
        m = np.asarray(m, dtype=np.float64)
        if m.ndim == 1:
            m = m.reshape(-1, 1)  # an Eigen::VectorXd is rows x 1
        return np.array(m.shape, dtype=np.int64).tobytes() + np.asfortranarray(m).tobytes(order="F")

    def save(self, v, object_name: str) -> None:
        os.makedirs(self._dir_name, exist_ok=True)
        with open(self.fname(object_name), "wb") as f:
            if isinstance(v, (list, tuple)):  # binary_archive.hpp:78-94
                f.write(np.array([len(v)], dtype=np.int32).tobytes())
                for x in v:
                    f.write(self._matrix_bytes(np.atleast_1d(np.asarray(x, dtype=np.float64))))
            else:
                f.write(self._matrix_bytes(v))

    @staticmethod
    def _read_matrix(buf: bytes, off: int):  # binary_archive.hpp:152-160 (_read_binary)
        rows, cols = np.frombuffer(buf, dtype=np.int64, count=2, offset=off)
        off += 16
        n = int(rows) * int(cols)
        m = np.frombuffer(buf, dtype=np.float64, count=n, offset=off).reshape((int(rows), int(cols)), order="F").copy()
        return m, off + 8 * n

    def _bytes(self, object_name: str) -> bytes:
        path = self.fname(object_name)
        assert os.path.exists(path), "file not found"
        with open(path, "rb") as f:
            return f.read()

    def load_matrix(self, object_name: str) -> np.ndarray:
        return self._read_matrix(self._bytes(object_name), 0)[0]

    def load_vector(self, object_name: str) -> np.ndarray:
        return self.load_matrix(object_name).reshape(-1, order="F")

    def load_vector_list(self, object_name: str):
        buf = self._bytes(object_name)
        n = int(np.frombuffer(buf, dtype=np.int32, count=1)[0])
        off, out = 4, []
        for _ in range(n):
            m, off = self._read_matrix(buf, off)
            out.append(m.reshape(-1, order="F"))
        assert out, "empty list"
        return out
