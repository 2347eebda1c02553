"""Scalar mirror-prescription helpers mirrored from the reference's ``conicsolve.py``
(host-side numpy; nothing here touches rays): ``primrad`` (conicsolve.py:7-15), ``secrad`` (:29-37),
``woltparam`` (:51-59), ``primfocus`` (:62-64), ``wsRMS`` (:250-254), ``wsFoc`` (:257-261),
``ellipsoidFunction`` (:263-281) and, at the end of the file, the sag / radius / intersection helpers.  Not mirrored:
the reference's debugging scratch (``mathraytrace``, ``wsPrimFunction*``, ``wsSecFunction*`` stop in ``pdb.set_trace()``;
``primaryintercept`` only prints).
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


# ---- the remaining closed-form prescription helpers (design-time scalars; conicsolve.py:17-27, 39-49, 85-89, 283-325) ----
def _sag(z, r, z0, z1):
    """Sag of a mirror segment: the quadratic term of a parabola fitted to its radius profile, over half its length."""
    import numpy as np
    return np.abs(np.polyfit(z, r, 2)[0] * ((z1 - z0) / 2.) ** 2)


def primsag(z1, r0, z0):
    """Sag of a primary that runs from the node z0 to z1 (conicsolve.py:17-27)."""
    import numpy as np
    z = np.linspace(z0, z1, 100)
    return _sag(z, primrad(z, r0, z0), z0, z1)


def secsag(z1, z0, r0, F, psi=1.):
    """Sag of a secondary that runs from z0 to z1 for the prescription (r0, F) (conicsolve.py:39-49)."""
    import numpy as np
    z = np.linspace(z0, z1, 100)
    return _sag(z, secrad(z, r0, F, psi=psi), z0, z1)


def rGoal_to_rMax(rgoal, z0, zmax):
    """The node radius r0 whose primary reaches radius rgoal at zmax: a 10000-point scan below rgoal (conicsolve.py:85-89)."""
    import numpy as np
    rguess = np.linspace(rgoal - 2., rgoal, 10000)
    return rguess[np.argmin(np.abs(rgoal - primrad(zmax, rguess, z0)))]


def ellipsoidRad(S, psi, R, F, z):
    """Radius of an ellipsoid primary at height z above the two-mirror focus (conicsolve.py:283-290)."""
    P, a, b, e, f = ellipsoidFunction(S, psi, R, F)
    return sqrt(1 - (z - (f - P + F)) ** 2 / a ** 2) * b


def ehSecRad(S, psi, R, F, z):
    """Radius of the hyperboloid secondary of an ellipsoid-hyperboloid pair at height z (conicsolve.py:292-300)."""
    P, a, b, e, f = ellipsoidFunction(S, psi, R, F)
    psi_eff = arctan(R / P) / (arctan(R / F) - arctan(R / P))
    return secrad(z, R, F, psi=psi_eff)


def ellipsoidSag(S, psi, R0, F, z1, z0):
    """Sag of an ellipsoid primary between z0 and z1 (conicsolve.py:302-309)."""
    import numpy as np
    z = np.linspace(z0, z1, 100)
    return _sag(z, ellipsoidRad(S, psi, R0, F, z), z0, z1)


def solveS(P, a, b, e, f, x, y, z, l, m, n):
    """Analytic ray / conic intersection: the two path lengths and their quadratic's coefficients (conicsolve.py:311-325)."""
    K = -e ** 2
    R = b ** 2 / a
    denom = l ** 2 + m ** 2 + (K + 1) * n ** 2
    b2 = (l * x + m * y - R * n + (K + 1) * n * z) / denom
    c2 = (x ** 2 + y ** 2 - 2 * R * z + (K + 1) * z ** 2) / denom
    return b2, c2, -b2 + sqrt(b2 ** 2 - c2), -b2 - sqrt(b2 ** 2 - c2)
