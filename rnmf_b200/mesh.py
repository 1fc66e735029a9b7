"""Mirror of src/mesh/cartesian2d/mesh.rs (host-only metadata; only n, dx, lo cross the C ABI)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np


@dataclass
class CartesianMesh:
    lo: Tuple[float, ...]
    hi: Tuple[float, ...]
    n: Tuple[int, ...]
    dx: Tuple[float, ...]
    dim: int

    @classmethod
    def new(cls, lo: Sequence[float], hi: Sequence[float], n: Sequence[int]) -> "CartesianMesh":
        """CartesianMesh2D::new (mesh.rs:36-45): dx = (hi - lo) / n."""
        dim = len(n)
        if dim not in (2, 3) or len(lo) != dim or len(hi) != dim:
            raise ValueError("mesh must be 2D or 3D with matching lo/hi/n")
        dx = tuple((float(hi[d]) - float(lo[d])) / float(n[d]) for d in range(dim))
        return cls(tuple(map(float, lo)), tuple(map(float, hi)), tuple(map(int, n)), dx, dim)

    def centres(self, d: int) -> np.ndarray:
        """cell-centre coordinates lo + (i + 0.5) * dx (mesh.rs:69-70)."""
        return self.lo[d] + (np.arange(self.n[d], dtype=np.float64) + 0.5) * self.dx[d]

    @property
    def node_pos(self) -> np.ndarray:
        """SoA table of mesh.rs:65-74 (x block then y block [then z])."""
        grids = np.meshgrid(*[self.centres(d) for d in reversed(range(self.dim))], indexing="ij")
        return np.concatenate([g.reshape(-1) for g in reversed(grids)])


CartesianMesh2D = CartesianMesh
CartesianMesh3D = CartesianMesh
