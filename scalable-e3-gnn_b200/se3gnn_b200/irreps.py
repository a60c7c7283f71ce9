"""Minimal O(3) irreps metadata (``Irrep``, ``MulIr``, ``Irreps``, ``Instruction``).

The reference tensor product uses e3nn purely as *metadata*
(``/root/reference/models/segnn/l1_tensor_prod.py:5,13-21,29-36,122-151``):
``Irreps.spherical_harmonics(1)``, ``.lmax``, ``.dim``, ``len()``, iteration
yielding ``.mul/.dim/.ir.l/.ir.p/.ir.dim`` and the ``Instruction`` named tuple.
e3nn is not installable in the build image, so this module supplies exactly
that surface.  Objects coming from a real e3nn install are accepted wherever an
``Irreps`` is expected (they are re-parsed through ``str()``).
"""
from __future__ import annotations

import re
from typing import Iterable, NamedTuple, Tuple, Union

__all__ = ["Irrep", "MulIr", "Irreps", "Instruction", "as_irreps"]


class Irrep(NamedTuple):
    l: int
    p: int  # +1 even, -1 odd

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def __repr__(self) -> str:  # "1o", "0e"
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    @staticmethod
    def parse(s: Union[str, "Irrep", Tuple[int, int]]) -> "Irrep":
        if isinstance(s, Irrep):
            return s
        if isinstance(s, tuple):
            l, p = s
            return Irrep(int(l), int(p))
        m = re.fullmatch(r"\s*(\d+)([eoy])\s*", str(s))
        if m is None:
            raise ValueError(f"cannot parse irrep {s!r}")
        l = int(m.group(1))
        c = m.group(2)
        p = {"e": 1, "o": -1, "y": (-1) ** l}[c]
        return Irrep(l, p)


class MulIr(NamedTuple):
    mul: int
    ir: Irrep

    @property
    def dim(self) -> int:
        return self.mul * self.ir.dim

    def __repr__(self) -> str:
        return f"{self.mul}x{self.ir!r}"


class Irreps(tuple):
    """Direct sum of irreps, e.g. ``Irreps("34x0e+10x1o")``."""

    def __new__(cls, spec: Union[str, "Irreps", Iterable, None] = None):
        if isinstance(spec, Irreps):
            return super().__new__(cls, tuple(spec))
        items = []
        if spec is None:
            pass
        elif isinstance(spec, str):
            s = spec.strip()
            if s:
                for tok in s.split("+"):
                    tok = tok.strip()
                    if "x" in tok:
                        mul, ir = tok.split("x")
                        items.append(MulIr(int(mul), Irrep.parse(ir)))
                    else:
                        items.append(MulIr(1, Irrep.parse(tok)))
        elif isinstance(spec, Irrep):
            items.append(MulIr(1, spec))
        elif hasattr(spec, "__iter__"):
            for it in spec:
                if isinstance(it, MulIr):
                    items.append(it)
                elif hasattr(it, "mul") and hasattr(it, "ir"):  # foreign (e3nn) _MulIr
                    items.append(MulIr(int(it.mul), Irrep(int(it.ir.l), int(it.ir.p))))
                elif isinstance(it, (tuple, list)) and len(it) == 2:
                    items.append(MulIr(int(it[0]), Irrep.parse(it[1])))
                else:
                    items.append(MulIr(1, Irrep.parse(it)))
        else:
            raise TypeError(f"cannot build Irreps from {type(spec)}")
        for it in items:
            if it.mul < 0:
                raise ValueError("negative multiplicity")
        return super().__new__(cls, items)

    # -- e3nn-compatible surface ------------------------------------------------
    @staticmethod
    def spherical_harmonics(lmax: int, p: int = -1) -> "Irreps":
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self) -> int:
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self) -> int:
        return sum(mi.mul for mi in self)

    @property
    def lmax(self) -> int:
        if len(self) == 0:
            raise ValueError("Cannot get lmax of empty Irreps")
        return max(mi.ir.l for mi in self)

    @property
    def ls(self):
        return [mi.ir.l for mi in self for _ in range(mi.mul)]

    def simplify(self) -> "Irreps":
        out = []
        for mi in self:
            if mi.mul == 0:
                continue
            if out and out[-1].ir == mi.ir:
                out[-1] = MulIr(out[-1].mul + mi.mul, mi.ir)
            else:
                out.append(mi)
        return Irreps(out)

    def count(self, ir) -> int:  # type: ignore[override]
        ir = Irrep.parse(ir)
        return sum(mi.mul for mi in self if mi.ir == ir)

    def slices(self):
        out, i = [], 0
        for mi in self:
            out.append(slice(i, i + mi.dim))
            i += mi.dim
        return out

    def __add__(self, other) -> "Irreps":  # type: ignore[override]
        return Irreps(tuple(self) + tuple(Irreps(other)))

    def __getitem__(self, i):
        r = super().__getitem__(i)
        return Irreps(r) if isinstance(i, slice) else r

    def __repr__(self) -> str:
        return "+".join(repr(mi) for mi in self)

    __str__ = __repr__


class Instruction(NamedTuple):
    """Same field order as ``e3nn.o3.Instruction`` (used at ``L1TP:151,193``)."""

    i_in1: int
    i_in2: int
    i_out: int
    connection_mode: str
    has_weight: bool
    path_weight: float
    path_shape: tuple


def as_irreps(x) -> Irreps:
    """Accept our Irreps, a string, or a foreign (e3nn) Irreps object."""
    if isinstance(x, Irreps):
        return x
    if isinstance(x, str):
        return Irreps(x)
    if hasattr(x, "__iter__"):
        return Irreps(list(x))
    return Irreps(str(x))
