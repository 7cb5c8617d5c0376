"""Applied-field factories: what a caller hands to ``solve(applied_field=...)`` (reference sources/constant.py,
sources/dipole.py, sources/vortex.py).  Host-side numpy: an applied field is evaluated once per solve at the mesh
vertices and uploaded."""
from __future__ import annotations

import numpy as np


class Parameter:
    """A callable ``f(x, y[, z])`` with bound keyword arguments (minimal stand-in for
    superscreen/parameter.py; arithmetic composition is out of scope)."""

    def __init__(self, func, **kwargs):
        self.func = func
        self.kwargs = kwargs

    def __call__(self, x, y, z=None):
        x, y = np.atleast_1d(x, y)
        if z is None:
            return np.asarray(self.func(x, y, **self.kwargs))
        z = np.atleast_1d(z)
        if z.ndim > 1:
            z = np.squeeze(z, axis=tuple(range(1, z.ndim)))
        return np.asarray(self.func(x, y, z, **self.kwargs))

    def __repr__(self):
        return f"Parameter<{getattr(self.func, '__name__', 'f')}({self.kwargs})>"

    # arithmetic between Parameters and real numbers gives a CompositeParameter (reference parameter.py:146-184)
    def __add__(self, other):
        return CompositeParameter(self, other, "+")

    def __radd__(self, other):
        return CompositeParameter(other, self, "+")

    def __sub__(self, other):
        return CompositeParameter(self, other, "-")

    def __rsub__(self, other):
        return CompositeParameter(other, self, "-")

    def __mul__(self, other):
        return CompositeParameter(self, other, "*")

    def __rmul__(self, other):
        return CompositeParameter(other, self, "*")

    def __truediv__(self, other):
        return CompositeParameter(self, other, "/")

    def __rtruediv__(self, other):
        return CompositeParameter(other, self, "/")

    def __pow__(self, other):
        return CompositeParameter(self, other, "**")

    def __rpow__(self, other):
        return CompositeParameter(other, self, "**")


class CompositeParameter(Parameter):
    """The result of ``+ - * / **`` between Parameters and / or real numbers: evaluates both sides at the same
    ``(x, y[, z])`` and combines them (reference parameter.py:200-273), e.g.
    ``ConstantField(0.1) + DipoleField(...)`` as the applied field of a solve."""

    _OPS = {"+": np.add, "-": np.subtract, "*": np.multiply, "/": np.divide, "**": np.power}

    def __init__(self, left, right, op):
        import numbers

        for side, what in ((left, "Left"), (right, "Right")):
            if not isinstance(side, (numbers.Real, Parameter)):
                raise TypeError(f"{what} must be a number, Parameter, or CompositeParameter, not {type(side)!r}.")
        if isinstance(left, numbers.Real) and isinstance(right, numbers.Real):
            raise TypeError("Either left or right must be a Parameter or CompositeParameter.")
        op = op.strip() if isinstance(op, str) else op
        if op not in self._OPS:
            raise ValueError(f"Unknown operator, {op!r}. Valid operators are {list(self._OPS)!r}.")
        self.left, self.right, self.operator = left, right, op
        self.func, self.kwargs = None, {}

    def __call__(self, x, y, z=None):
        lv = self.left(x, y, z) if isinstance(self.left, Parameter) else self.left
        rv = self.right(x, y, z) if isinstance(self.right, Parameter) else self.right
        return self._OPS[self.operator](lv, rv)

    def __repr__(self):
        return f"({self.left!r} {self.operator} {self.right!r})"


class Constant(Parameter):
    def __init__(self, value, dimensions: int = 2):
        if dimensions == 2:
            super().__init__(lambda x, y, value=0: value * np.ones_like(x, dtype=float), value=value)
        else:
            super().__init__(lambda x, y, z, value=0: value * np.ones_like(x, dtype=float), value=value)


def constant(x, y, z, value=0):
    return value * np.ones_like(x, dtype=float)


def ConstantField(value: float = 0) -> Parameter:
    """A Parameter that returns ``value`` at all ``x, y, z``."""
    return Parameter(constant, value=float(value))


# ---------------------------------------------------------------------------------------------------------
# point dipoles (reference sources/dipole.py)
# ---------------------------------------------------------------------------------------------------------
def dipole_field(eval_coords, r0=(0, 0, 0), moment=(0, 0, 0)) -> np.ndarray:
    """Field ``B = mu_0 / 4 pi * (3 r (r . m) / |r|^5 - m / |r|^3)`` in tesla of one dipole with moment ``moment``
    (A m^2) at ``r0``, evaluated at ``eval_coords`` (``(3,)`` or ``(n, 3)``, metres); ``r`` is measured from the
    dipole (reference sources/dipole.py:11-55).  Returns ``(3,)`` or ``(n, 3)``."""
    from scipy.constants import mu_0

    m = np.atleast_1d(np.asarray(moment, dtype=float)).reshape(-1)
    r = np.atleast_2d(eval_coords).reshape((-1, 3)) - np.atleast_1d(np.asarray(r0, dtype=float)).reshape(-1)
    dist = np.sqrt(np.einsum("ij,ij->i", r, r))[:, None]
    along = (r @ m)[:, None]
    B = 3.0 * r * along / dist**5 - m[None, :] / dist**3
    return mu_0 / (4.0 * np.pi) * B.squeeze()


def dipole_distribution(x, y, z, *, dipole_positions, dipole_moments, component=None, length_units: str = "um",
                        moment_units: str = "mu_B") -> np.ndarray:
    """Field (tesla), or one Cartesian component of it, of a set of dipoles at ``dipole_positions`` (``(m, 3)``, in
    ``length_units``) with ``dipole_moments`` (``(3,)`` for all or ``(m, 3)``, in ``moment_units``) at the points
    ``(x, y, z)`` given in ``length_units`` (reference sources/dipole.py:58-129)."""
    from . import units as _u

    index = Ellipsis if component is None else "xyz".index(component)
    to_m = _u.conversion_factor(length_units, "m")
    moments = np.atleast_2d(np.asarray(dipole_moments, dtype=float) * _u.conversion_factor(moment_units, "A * m ** 2"))
    positions = np.atleast_2d(np.asarray(dipole_positions, dtype=float) * to_m)
    x, y, z = (np.atleast_1d(np.asarray(v, dtype=float) * to_m) for v in (x, y, z))
    if len(z) == 1:
        z = z * np.ones_like(x)
    points = np.stack([x, y, z], axis=1)
    if moments.shape[0] == 1:
        moments = np.repeat(moments, positions.shape[0], axis=0)
    elif moments.shape[0] != positions.shape[0]:
        raise ValueError(
            f"The number of dipole moments ({moments.shape[0]}) must be either"
            f"1 or equal to the the number of dipole positions ({positions.shape[0]}).")
    B = sum(dipole_field(points, moment=mom, r0=pos) for mom, pos in zip(moments, positions))
    return np.atleast_2d(B)[:, index]


def DipoleField(*, dipole_positions, dipole_moments, component=None, length_units: str = "um",
                moment_units: str = "mu_B") -> Parameter:
    """A Parameter for the field ``mu_0 H`` (tesla; one component or the vector) of a distribution of point dipoles
    (reference sources/dipole.py:132-182)."""
    if component not in (None, "x", "y", "z"):
        raise ValueError(f"Component must be 'x', 'y', 'z', or None (got {component!r}).")
    return Parameter(dipole_distribution, dipole_positions=dipole_positions, dipole_moments=dipole_moments,
                     component=component, length_units=length_units, moment_units=moment_units)


# ---------------------------------------------------------------------------------------------------------
# vortices: a flux monopole and the Pearl vortex of a thin film (reference sources/vortex.py)
# ---------------------------------------------------------------------------------------------------------
def monopole(x, y, z, *, r0=(0, 0, 0), nPhi0=1, vector: bool = False):
    """``mu_0 H = n Phi_0 / (2 pi) * (r - r0) / |r - r0|^3`` in units of ``Phi_0 / length_units**2``: the field of an
    isolated vortex seen from far away (reference sources/vortex.py:8-44).  z component, or the vector."""
    d = [np.asarray(c, dtype=float) - c0 for c, c0 in zip((x, y, z), r0)]
    scale = nPhi0 / (2.0 * np.pi * (d[0] ** 2 + d[1] ** 2 + d[2] ** 2) ** 1.5)
    if vector:
        return np.stack([d[0] * scale, d[1] * scale, d[2] * scale], axis=1)
    return d[2] * scale


def MonopoleField(r0=(0, 0, 0), nPhi0=1, vector: bool = False) -> Parameter:
    """reference sources/vortex.py:47-78"""
    return Parameter(monopole, r0=r0, nPhi0=nPhi0, vector=vector)


VortexField = MonopoleField


def pearl_vortex(x, y, z, *, xs, ys, Lambda: float = 0, r0=(0, 0, 0), nPhi0=1):
    """``mu_0 H_z`` (``Phi_0 / length_units**2``) of a Pearl vortex in a uniform film with effective penetration depth
    ``Lambda``, at a plane ``z = const`` (reference sources/vortex.py:84-170).  In Fourier space
    ``F{mu_0 H_z}(k, z) = n Phi_0 exp(-k z) / (1 + 2 Lambda k)``; the inverse transform is taken on the grid
    ``xs x ys`` and interpolated linearly to ``(x, y)``, which must lie inside that grid."""
    from scipy.interpolate import LinearNDInterpolator

    x, y, z = np.atleast_1d(x, y, z)
    if not np.allclose(z, z[0]):
        raise ValueError("All elements of the vector z must be equal.")
    x, y, height = x - r0[0], y - r0[1], abs(z[0] - r0[2])
    xs, ys = np.sort(xs), np.sort(ys)
    if x.min() < xs.min() or x.max() > xs.max() or y.min() < ys.min() or y.max() > ys.max():
        raise ValueError("The evaluation coordinates (x, y) must lie within the domain defined by (xs, ys).")
    dx, dy = xs[1] - xs[0], ys[1] - ys[0]
    kx = np.linspace(-np.pi / dx, np.pi / dx, xs.shape[0], endpoint=False)
    ky = np.linspace(-np.pi / dy, np.pi / dy, ys.shape[0], endpoint=False)
    KX, KY = np.meshgrid(kx, ky)
    k = np.sqrt(KX**2 + KY**2)
    spectrum = np.fft.fftshift(nPhi0 * np.exp(-k * height) / (1.0 + 2.0 * Lambda * k))
    hz = np.abs(np.fft.fftshift(np.fft.ifft2(spectrum))) / (dx * dy)
    X, Y = np.meshgrid(xs, ys)
    interp = LinearNDInterpolator(np.array([X.ravel(), Y.ravel()]).T, hz.ravel())
    return interp(np.array([x, y]).T).squeeze()


def PearlVortexField(*, r0=(0, 0, 0), Lambda: float = 0, nPhi0=1, xs, ys) -> Parameter:
    """reference sources/vortex.py:173-227"""
    return Parameter(pearl_vortex, xs=xs, ys=ys, Lambda=Lambda, r0=r0, nPhi0=nPhi0)
