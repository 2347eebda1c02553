"""GPU parity for the OPD-map pipeline's scattered-data interpolation: analyses.interpolateVec /
wavefront (analyses.py:189-230, 305-334) against the oracle restatement, which calls scipy's own griddata -- the
reference's third-party dependency -- and the C oracle's reconstruct.  The device path builds no triangulation (the
Delaunay triangle of each query from its natural neighbours), so agreement with Qhull's triangulation on every grid
point, NaN mask included, is the test."""
import numpy as np
import pytest

from oracle import refapi
from util import chains, copy, pyref

pytestmark = pytest.mark.gpu
scipy_interpolate = pytest.importorskip("scipy.interpolate")


@pytest.fixture(scope="module")
def pxf():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pyxfocus_b200
    return pyxfocus_b200


def to_dev(rays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(r)).cuda() for r in rays]


def bundle(x, y, seed):
    """A bundle on the given footprint whose slopes are a smooth field plus noise (so that wrong triangles show)."""
    rng = np.random.default_rng(seed)
    n = x.size
    s = max(np.ptp(x), np.ptp(y))
    l = 1e-3 * np.sin(3 * x / s) * np.cos(2 * y / s) + 1e-5 * rng.normal(size=n)
    m = 2e-3 * (x / s) ** 2 - 1e-3 * (y / s) + 1e-5 * rng.normal(size=n)
    nn = np.sqrt(1. - l ** 2 - m ** 2)
    z = np.zeros(n)
    return [rng.normal(size=n), x.copy(), y.copy(), z, l, m, nn, z.copy(), z.copy(), z + 1.]


def footprints():
    rng = np.random.default_rng(5)
    out = {}
    n = 20_000
    out["square"] = (rng.uniform(-10, 10, n), rng.uniform(-7, 7, n))
    r, t = np.sqrt(rng.uniform(200. ** 2, 230. ** 2, n)), rng.uniform(-.15, .15, n)
    out["annulus_sector"] = (r * np.cos(t), r * np.sin(t))
    r, t = 12.5 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
    out["disc"] = (r * np.cos(t), r * np.sin(t))
    # strongly non-uniform density: a tight cluster inside a sparse field (stresses the cell grid)
    out["clustered"] = (np.concatenate([rng.normal(0, .05, 15_000), rng.uniform(-10, 10, 400)]),
                        np.concatenate([rng.normal(0, .05, 15_000), rng.uniform(-10, 10, 400)]))
    out["elongated"] = (rng.uniform(0, 1000, 6000), rng.uniform(0, 1, 6000))
    out["tiny"] = (np.array([0., 1., .2, .9, .5]), np.array([0., .1, 1., .8, .45]))
    return out


def compare(got, want, scale, what, tol=1e-10):
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    assert got.shape == want.shape, what
    ng, nw = np.isnan(got), np.isnan(want)
    assert np.array_equal(ng, nw), "%s: NaN mask differs at %d of %d grid points" % (what, (ng != nw).sum(), ng.size)
    err = np.abs(np.where(nw, 0., got - want)).max()
    assert err <= tol * scale, "%s: max |delta| %.3e (scale %.3e)" % (what, err, scale)
    return int((~nw).sum())


@pytest.mark.parametrize("name", ["square", "annulus_sector", "disc", "clustered", "elongated", "tiny"])
def test_interpolatevec_linear_and_nearest(pxf, name):
    x, y = footprints()[name]
    rays = bundle(x, y, 7)
    dev = to_dev(rays)
    Nx, Ny = (64, 48) if name != "tiny" else (9, 7)
    filled = 0
    for I in (4, 5, 0):
        for method in ("linear", "nearest"):
            want, dxw, dyw = refapi.interpolateVec(copy(rays), I, Nx, Ny, method=method)
            got, dx, dy = pxf.analyses.interpolateVec(dev, I, Nx, Ny, method=method)
            assert dx == dxw and dy == dyw
            filled += compare(got, want, np.abs(rays[I]).max(), "%s I=%d %s" % (name, I, method))
    assert filled > 0
    # a caller-supplied range (partly outside the data) and vector
    vec = np.cos(x) + y
    xr, yr = [x.min() - .1 * np.ptp(x), x.mean()], [y.mean(), y.max() + .2 * np.ptp(y)]
    want, _, _ = refapi.interpolateVec(copy(rays), 1, 33, 21, xr=xr, yr=yr, interpVec=vec)
    got, _, _ = pxf.analyses.interpolateVec(dev, 1, 33, 21, xr=xr, yr=yr, interpVec=vec)
    compare(got, want, np.abs(vec).max(), name + " xr/yr/interpVec")
    with pytest.raises(ValueError):
        pxf.analyses.interpolateVec(dev, 4, 8, 8, method="quintic")


@pytest.mark.parametrize("name", ["square", "annulus_sector", "disc", "clustered", "tiny"])
def test_interpolatevec_cubic(pxf, name):
    """method='cubic': scipy's Clough-Tocher interpolant.  Its vertex gradients are Gauss-Seidel sweeps in input order to
    a relative change of 1e-6; the device runs the same sweeps level by level, so the maps agree to rounding (1e-9 of
    the data's scale leaves room for the different order of the neighbour sums), not merely to the 1e-6 tolerance."""
    x, y = footprints()[name]
    rays = bundle(x, y, 9)
    dev = to_dev(rays)
    Nx, Ny = (64, 48) if name != "tiny" else (9, 7)
    for I in (4, 5):
        want, _, _ = refapi.interpolateVec(copy(rays), I, Nx, Ny, method="cubic")
        got, _, _ = pxf.analyses.interpolateVec(dev, I, Nx, Ny, method="cubic")
        assert compare(got, want, np.abs(rays[I]).max(), "%s I=%d cubic" % (name, I), tol=1e-9) > 0


@pytest.mark.parametrize("name", ["annulus_sector", "disc"])
def test_interpolatevec_polar(pxf, name):
    x, y = footprints()[name]
    rays = bundle(x, y, 8)
    dev = to_dev(rays)
    want, _, _ = refapi.interpolateVec(copy(rays), 5, 40, 30, method="linear", polar=True)
    got, _, _ = pxf.analyses.interpolateVec(dev, 5, 40, 30, method="linear", polar=True)
    # the charts multiply an angle by a radius of ~200: coordinates agree to ~1e-13 relative, and a query that sits
    # on a triangle edge to that precision may take the neighbouring triangle -- same plane to first order
    compare(got, want, np.abs(rays[5]).max(), name + " polar", tol=1e-6)


def test_griddata_queries_on_data_points_and_outside(pxf):
    import torch
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-1, 1, 5000), rng.uniform(-1, 1, 5000)
    v = np.sin(4 * x) * y
    qx = np.concatenate([x[:100], [5., -5., 0.], [x.min() - 1e-9]])
    qy = np.concatenate([y[:100], [0., 0., 7.], [0.]])
    got = pxf.analyses.griddata(*(torch.from_numpy(a).cuda() for a in (x, y, v, qx, qy))).cpu().numpy()
    want = scipy_interpolate.griddata((x, y), v, (qx, qy), method="linear")
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[:100], v[:100])                   # a query on a data point returns its value
    assert np.isnan(got[100:]).all()


def test_interpolatevec_on_a_traced_bundle_and_wavefront(pxf):
    """The metrology use: slopes of a traced bundle a little out of focus -> maps -> Southwell phase."""
    rays = chains.wolter1_source(30_000, 3, dphi=.3)
    chains.run_steps_cpu(rays, chains.wolter1_steps())
    pyref.transform(rays, 0, 0, 40., 0, 0, 0)
    from oracle import f2py as of
    of.surfacesf.flat(*rays[1:])
    dev = to_dev(rays)
    for I in (4, 5):
        want, _, _ = refapi.interpolateVec(copy(rays), I, 50, 40)
        got, _, _ = pxf.analyses.interpolateVec(dev, I, 50, 40)
        assert compare(got, want, np.abs(rays[I]).max(), "traced I=%d" % I) > 500
    pw, xw, yw = refapi.wavefront(copy(rays), 30, 24, method="linear", maxiter=2000)
    pg, xg, yg = pxf.analyses.wavefront(dev, 30, 24, method="linear", maxiter=2000)
    compare(xg, xw, np.nanmax(np.abs(xw)), "wavefront x slopes")
    compare(yg, yw, np.nanmax(np.abs(yw)), "wavefront y slopes")
    # the reconstruction amplifies slope differences of 1e-16 through a few hundred SOR sweeps
    compare(pg, pw, np.nanmax(np.abs(pw)), "wavefront phase", tol=1e-8)
    assert np.isfinite(pg).sum() > 100
    # the default method ('cubic') on a filled footprint
    x, y = footprints()["disc"]
    rays = bundle(x, y, 11)
    dev = to_dev(rays)
    pw, xw, yw = refapi.wavefront(copy(rays), 40, 32, maxiter=3000)
    pg, xg, yg = pxf.analyses.wavefront(dev, 40, 32, maxiter=3000)
    compare(xg, xw, np.nanmax(np.abs(xw)), "wavefront (cubic) x slopes", tol=1e-9)
    compare(yg, yw, np.nanmax(np.abs(yw)), "wavefront (cubic) y slopes", tol=1e-9)
    compare(pg, pw, np.nanmax(np.abs(pw)), "wavefront (cubic) phase", tol=1e-7)


@pytest.mark.parametrize("name", ["square", "annulus_sector", "disc", "clustered", "tiny"])
def test_delaunay_neighbor_rings(pxf, name):
    """The gift-wrapped neighbour rings against Qhull's triangulation: same neighbour SETS for every vertex, hull flags
    equal, and each ring counter-clockwise."""
    import torch
    from scipy.spatial import Delaunay
    x, y = footprints()[name]
    tri = Delaunay(np.column_stack([x, y]))
    indptr, indices = tri.vertex_neighbor_vertices
    ring, deg, hull = (t.cpu().numpy() for t in pxf.analyses.delaunay_neighbors(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()))
    on_hull = np.zeros(x.size, dtype=bool)
    on_hull[np.unique(tri.convex_hull)] = True
    bad = 0
    for i in range(x.size):
        want = set(indices[indptr[i]:indptr[i + 1]].tolist())
        got = ring[i, :deg[i]].tolist()
        if set(got) != want or len(got) != len(want):
            bad += 1
            continue
        ang = np.unwrap(np.arctan2(y[got] - y[i], x[got] - x[i]))
        assert np.all(np.diff(ang) > 0), "ring of vertex %d is not counter-clockwise" % i
    assert bad == 0, "%d of %d vertices have a different neighbour set" % (bad, x.size)
    assert np.array_equal(hull.astype(bool), on_hull)


def test_interpolatevec_cubic_polar_and_degenerate_inputs(pxf):
    import torch
    x, y = footprints()["disc"]
    rays = bundle(x, y, 12)
    dev = to_dev(rays)
    want, _, _ = refapi.interpolateVec(copy(rays), 4, 30, 24, method="cubic", polar=True)
    got, _, _ = pxf.analyses.interpolateVec(dev, 4, 30, 24, method="cubic", polar=True)
    compare(got, want, np.abs(rays[4]).max(), "cubic polar", tol=1e-6)          # (chart coordinates: see the linear test)
    # three points: one triangle
    t = [torch.tensor(v, dtype=torch.float64).cuda() for v in ([0., 1., 0.], [0., 0., 1.], [1., 2., 3.])]
    q = [torch.tensor(v, dtype=torch.float64).cuda() for v in ([.25, .9, .2], [.25, .9, .2])]
    for method in ("linear", "cubic", "nearest"):
        got = pxf.analyses.griddata(*t, *q, method=method).cpu().numpy()
        want = scipy_interpolate.griddata((t[0].cpu().numpy(), t[1].cpu().numpy()), t[2].cpu().numpy(),
                                          (q[0].cpu().numpy(), q[1].cpu().numpy()), method=method)
        assert np.allclose(got, want, rtol=0, atol=1e-12, equal_nan=True), method
    # collinear points have no triangulation (Qhull raises): every query is outside the (flat) hull here, never a hang
    line = [torch.linspace(0, 1, 50, dtype=torch.float64).cuda() for _ in range(2)]
    got = pxf.analyses.griddata(line[0], line[1], line[0], q[0], q[1] + .05, method="linear").cpu().numpy()
    assert np.isnan(got).all()
    # duplicate points: a loud error for the methods that need the triangulation
    dup = [torch.tensor(v, dtype=torch.float64).cuda() for v in ([0., 1., 0., 1., .5, .5], [0., 0., 1., 1., .5, .5], [1., 2., 3., 4., 5., 5.])]
    with pytest.raises(pxf.PxfError):
        pxf.analyses.griddata(*dup, q[0], q[1], method="cubic")


def test_interpolatevec_at_scale(pxf):
    """2e5 points -> 128 x 128 grid, all three methods against scipy (the largest size scipy finishes in seconds)."""
    rng = np.random.default_rng(21)
    n = 200_000
    r, t = 12.5 * np.sqrt(rng.uniform(0, 1, n)), rng.uniform(0, 2 * np.pi, n)
    rays = bundle(r * np.cos(t), r * np.sin(t), 22)
    dev = to_dev(rays)
    for method, tol in (("linear", 1e-10), ("cubic", 1e-9), ("nearest", 0.)):
        want, _, _ = refapi.interpolateVec(copy(rays), 5, 128, 128, method=method)
        got, _, _ = pxf.analyses.interpolateVec(dev, 5, 128, 128, method=method)
        assert compare(got, want, np.abs(rays[5]).max(), "2e5 points " + method, tol=tol) > 10_000
