"""Mirror of src/boundary_conditions.rs: BcType, ComponentBCs, BCs (same names and meaning)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

from . import _capi


@dataclass(frozen=True)
class BcType:
    """boundary_conditions.rs:11-24.  Use the constructors Dirichlet/Neumann/Reflect/Prescribed."""
    kind: int
    value: float = 0.0
    values: Optional[tuple] = None

    @staticmethod
    def Dirichlet(value: float) -> "BcType":
        return BcType(_capi.BC_DIRICHLET, float(value))

    @staticmethod
    def Neumann(gradient: float) -> "BcType":
        return BcType(_capi.BC_NEUMANN, float(gradient))

    @staticmethod
    def Reflect() -> "BcType":
        return BcType(_capi.BC_REFLECT)

    @staticmethod
    def Prescribed(values: Sequence[float]) -> "BcType":
        return BcType(_capi.BC_PRESCRIBED, 0.0, tuple(float(v) for v in values))


@dataclass
class ComponentBCs:
    """boundary_conditions.rs:41-54: lo[d], hi[d] per direction."""
    lo: List[BcType]
    hi: List[BcType]


@dataclass
class BCs:
    """boundary_conditions.rs:57-70."""
    bcs: List[ComponentBCs] = field(default_factory=list)

    @staticmethod
    def empty() -> "BCs":
        return BCs([])

    def clone(self) -> "BCs":
        return BCs([ComponentBCs(list(c.lo), list(c.hi)) for c in self.bcs])
