"""TEST INFRASTRUCTURE ONLY.  The subset of `e3nn.o3` that the reference tensor
product touches: `Irreps` (parse, iterate, .dim/.lmax/len, spherical_harmonics)
and the `Instruction` named tuple (reference L1TP:5,13-21,29-36,98,122-151,193)."""
from collections import namedtuple

Instruction = namedtuple(
    "Instruction",
    "i_in1 i_in2 i_out connection_mode has_weight path_weight path_shape",
)


class _Ir(namedtuple("_Ir", "l p")):
    @property
    def dim(self):
        return 2 * self.l + 1

    def __repr__(self):
        return "%d%s" % (self.l, "e" if self.p == 1 else "o")


class _MulIr(namedtuple("_MulIr", "mul ir")):
    @property
    def dim(self):
        return self.mul * self.ir.dim

    def __repr__(self):
        return "%dx%r" % (self.mul, self.ir)


class Irreps(tuple):
    def __new__(cls, spec=""):
        if isinstance(spec, Irreps):
            return tuple.__new__(cls, spec)
        items = []
        if isinstance(spec, str):
            for tok in [t.strip() for t in spec.split("+") if t.strip()]:
                mul, _, ir = tok.rpartition("x")
                items.append(_MulIr(int(mul) if mul else 1,
                                    _Ir(int(ir[:-1]), 1 if ir[-1] == "e" else -1)))
        else:
            for mul, (l, p) in spec:
                items.append(_MulIr(int(mul), _Ir(int(l), int(p))))
        return tuple.__new__(cls, items)

    @staticmethod
    def spherical_harmonics(lmax, p=-1):
        return Irreps([(1, (l, p ** l)) for l in range(lmax + 1)])

    @property
    def dim(self):
        return sum(m.dim for m in self)

    @property
    def lmax(self):
        return max(m.ir.l for m in self)

    def __repr__(self):
        return "+".join(repr(m) for m in self)
