"""Python mirror of the reference's operator classes for the hot path (eul/ variant signatures).

Same names, constructor arguments and ``assemble(...)`` argument meaning as eul/Assembly.h:
``assemble`` only records (level, scale, flags, coefficient) -- there is no sparse matrix -- and
``mult(x)`` is the reference's ``MatMult(op->M, x, y)``, executed by the CUDA kernels.  Fields are
column-layout device tensors (see Engine); ``lev`` may address a single level (a one-column field,
the MatShell use) or, with ``nlev`` columns, levels lev .. lev+nlev-1 in one launch.

The C++ mirror with the PETSc MatShell adaptor lives in mimsem_b200/host/.
"""
from .engine import Engine, FIXED_LEVEL  # noqa: F401

SCALE = 1.0e8  # eul/Assembly.cpp:20


class _Op:
    def __init__(self, engine):
        self.engine = engine
        self.lev, self.scale, self.tpow, self.flags, self.coeff = 0, 1.0, 0, 0, None

    def mult(self, x, out=None):
        return self.engine.apply(self.OP, x, coeff=self.coeff, out=out, lev0=self.lev, scale=self.scale, tpow=self.tpow,
                                 flags=self.flags)


class Umat(_Op):
    """1-form mass matrix.  eul/Assembly.cpp:26-153: Umat(topo, geom, l, e); assemble(lev, scale, vert_scale)."""
    OP = "M1"

    def assemble(self, lev, scale, vert_scale):
        self.lev, self.scale, self.tpow = lev, scale, 1 if vert_scale else 0


class Wmat(_Op):
    """2-form mass matrix.  eul/Assembly.cpp:288-373: assemble(lev, scale, vert_scale)."""
    OP = "M2"

    def assemble(self, lev, scale, vert_scale):
        self.lev, self.scale, self.tpow = lev, scale, 1 if vert_scale else 0


class Pmat(_Op):
    """0-form mass matrix.  eul/Assembly.cpp:2004-2098: assemble(lev, scale); assemble_h(lev, scale, h2)."""
    OP = "M0"

    def assemble(self, lev, scale):
        self.OP, self.lev, self.scale, self.tpow, self.coeff = "M0", lev, scale, 1, None

    def assemble_h(self, lev, scale, h2):
        self.OP, self.lev, self.scale, self.tpow, self.coeff = "M0h", lev, scale, 2, h2


class Uhmat(_Op):
    """1-form mass matrix weighted by a 2-form.  eul/Assembly.cpp:376-474: assemble(h2, lev, const_vert, scale)."""
    OP = "M1h"

    def assemble(self, h2, lev, const_vert, scale):
        self.coeff, self.lev, self.scale, self.tpow = h2, lev, scale, 2 if const_vert else 1


class Whmat(_Op):
    """2-form mass matrix weighted by a 2-form.  eul/Assembly.cpp:1243-1299: assemble(rho, lev, scale, vert_scale_rho)."""
    OP = "M2h"

    def assemble(self, rho, lev, scale, vert_scale_rho):
        self.coeff, self.lev, self.scale, self.tpow = rho, lev, scale, 2 if vert_scale_rho else 1


class WtQUmat(_Op):
    """Kinetic-energy operator K(u1): 1-form -> 2-form.  eul/Assembly.cpp:908-986: assemble(u1, lev, scale)."""
    OP = "K"

    def assemble(self, u1, lev, scale):
        self.coeff, self.lev, self.scale, self.tpow = u1, lev, scale, 2


class E10mat:
    """Edge-node incidence.  eul/Assembly.cpp:1102-1162: members E10 and E01 = -E10^T."""

    def __init__(self, engine):
        self.engine = engine

    def mult_E10(self, x, out=None):
        return self.engine.apply("E10", x, out=out)

    def mult_E01(self, x, out=None):
        return self.engine.apply("E01", x, out=out)


class E21mat:
    """Face-edge incidence.  eul/Assembly.cpp:1170-1220: members E21 and E12 = -E21^T."""

    def __init__(self, engine):
        self.engine = engine

    def mult_E21(self, x, out=None):
        return self.engine.apply("E21", x, out=out)

    def mult_E12(self, x, out=None):
        return self.engine.apply("E12", x, out=out)
