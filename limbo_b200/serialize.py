"""serialize::TextArchive of the reference (src/limbo/serialize/text_archive.hpp:63-151): one `<name>.dat` file per
object in a directory, rows separated by new lines, values by single spaces, full precision.  GPs written by the
reference load here and vice versa (same six objects as GP::save, model/gp.hpp:448-460)."""
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
