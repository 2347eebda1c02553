"""Scalar mirror-prescription helpers mirrored from the reference's ``conicsolve.py``
(host-side numpy; nothing here touches rays).  Only the functions the hot-path wrappers and
the known-answer tests need: ``primrad`` (conicsolve.py:7-15), ``secrad`` (:29-37),
``woltparam`` (:51-59), ``primfocus`` (:62-64), ``wsRMS`` (:250-254), ``wsFoc`` (:257-261),
``ellipsoidFunction`` (:263-281).
"""
from numpy import arcsin, arctan, cos, sin, sqrt, tan


def _vs(r0, z0, psi):
    alpha = .25 * arctan(r0 / z0)
    thetah = 2 * (1 + 2 * psi) / (1 + psi) * alpha
    thetap = 2 * psi / (1 + psi) * alpha
    p = z0 * tan(4 * alpha) * tan(thetap)
    d = z0 * tan(4 * alpha) * tan(4 * alpha - thetah)
    e = cos(4 * alpha) * (1 + tan(4 * alpha) * tan(thetah))
    return alpha, p, d, e


def primrad(z, r0, z0, psi=1.):
    """Radius of the Wolter-I primary at axial position z."""
    alpha, p, d, e = _vs(r0, z0, psi)
    return sqrt(p ** 2 + 2 * p * z + (4 * e ** 2 * p * d) / (e ** 2 - 1))


def secrad(z, r0, z0, psi=1.):
    """Radius of the Wolter-I secondary at axial position z."""
    alpha, p, d, e = _vs(r0, z0, psi)
    return sqrt(e ** 2 * (d + z) ** 2 - z ** 2)


def woltparam(r0, z0):
    """(alpha, p, d, e) with thetah = 3 alpha, thetap = alpha (psi = 1)."""
    alpha = .25 * arctan(r0 / z0)
    thetah = 3 * alpha
    thetap = alpha
    p = z0 * tan(4 * alpha) * tan(thetap)
    d = z0 * tan(4 * alpha) * tan(4 * alpha - thetah)
    e = cos(4 * alpha) * (1 + tan(4 * alpha) * tan(thetah))
    return (alpha, p, d, e)


def primfocus(r0, z0):
    """Distance to the primary (paraboloid) focus."""
    alpha, p, d, e = woltparam(r0, z0)
    return z0 + 2 * e ** 2 * d / (e ** 2 - 1)


def wsRMS(psi, theta, alpha, L1, z0):
    """RMS blur at the optimum focal surface (Chase & Van Speybroeck Eq. 13)."""
    return .135 * (psi + 1) * (tan(theta) ** 2 / tan(alpha)) * L1 / z0


def wsFoc(r, psi, L1, z0, alpha):
    """Optimum focal-surface height at radius r (Chase & Van Speybroeck)."""
    return .0625 * (psi + 1) * (r ** 2 * L1 / z0 ** 2) / tan(alpha) ** 2


def ellipsoidFunction(S, psi, R, F):
    """(P, a, b, e, f) of the ellipsoid primary of an ellipsoid-hyperboloid telescope."""
    P = R / sin((psi * arcsin(R / F) - arcsin(R / S)) / (1 + psi))
    f = (S + P) / 2.
    a = 1.
    b = -(R ** 2 + (f - P) ** 2 + f ** 2)
    c = f ** 2 * (f - P) ** 2
    a = sqrt((-b + sqrt(b ** 2 - 4 * a * c)) / (2 * a))
    b = sqrt(a ** 2 - f ** 2)
    e = f / a
    return P, a, b, e, f
