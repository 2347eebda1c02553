"""One routine alone on a resident bundle, a few launches -- the command the per-routine ncu captures wrap.
    python profiles/routine_probe.py <routine> [rays] [launches]
routine: refract woltersine wsprimary wssecondary spocone tracezern wolterprimll woltersecll radgrat conic
Prints the best CUDA-event time of the launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyxfocus_b200 as pxf  # noqa: E402
from pyxfocus_b200 import surfacesf as SF, transformationsf as TF, woltsurf as WS, zernsurf as ZS  # noqa: E402
from pyxfocus_b200._call import bundle_alloc  # noqa: E402


def main():
    name = sys.argv[1]
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    dev = torch.device("cuda", 0)
    src = pxf.sources.subannulus(220., 220.6, 2 * np.pi, n, zhat=-1., rng="philox", seed=0, device=dev)
    state = bundle_alloc(n, dev)
    W = bundle_alloc(n, dev)

    def cp(dst, s_):
        for a, b in zip(dst, s_):
            a.copy_(b)
    cp(state, src)
    TF.transform(*state[1:], 0., 0., 8400., 0., 0., 0.)            # above the primary
    alpha = pxf.conicsolve.woltparam(220., 8400.)[0]
    if name in ("refract", "woltersecll", "woltersecondary"):
        WS.wolterprimary(*state[1:], 220., 8400., 1.)
        TF.reflect(*state[4:])
    if name == "wssecondary":
        WS.wsprimary(*state[1:], alpha, 8400., 1.)
        TF.reflect(*state[4:])
    if name in ("tracezern", "conic", "radgrat"):
        WS.wolterprimary(*state[1:], 220., 8400., 1.)
        TF.reflect(*state[4:])
        WS.woltersecondary(*state[1:], 220., 8400., 1.)
        TF.reflect(*state[4:])
    ro = [r for r in range(8) for _ in range(r + 1)]
    ao = [m for r in range(8) for m in range(-r, r + 1, 2)]
    zc = np.random.default_rng(0).normal(0., 1e-4, len(ro))
    zc[:3] = 0.
    llc = np.random.default_rng(1).normal(0., 1e-5, 36)
    lla = np.repeat(np.arange(6), 6).astype(np.int32)
    llz = np.tile(np.arange(6), 6).astype(np.int32)
    fns = {
        "refract": lambda: TF.refract(*W[4:], 1., 1.5),
        "woltersine": lambda: WS.woltersine(*W[1:], 220., 8400., 1e-4, .05),
        "wsprimary": lambda: WS.wsprimary(*W[1:], alpha, 8400., 1.),
        "wssecondary": lambda: WS.wssecondary(*W[1:], alpha, 8400., 1.),
        "spocone": lambda: WS.spocone(*W[1:], 220., .0065),
        "conic": lambda: SF.conic(*W[1:], 2e4, -1.),
        "radgrat": lambda: TF.radgrat(W[1], W[2], W[4], W[5], W[6], 2.4e-6, 160. / 11832.911, -1.),
        "tracezern": lambda: ZS.tracezern(*W[1:], zc, np.array(ro), np.array(ao), 230.),
        "wolterprimll": lambda: WS.wolterprimll(*W[1:], 220., 8400., 8500., 8400., 2 * np.pi, llc, lla, llz),
        "woltersecondary": lambda: WS.woltersecondary(*W[1:], 220., 8400., 1.),
        "woltersecll": lambda: WS.woltersecll(*W[1:], 220., 8400., 1., 8400., 8300., 2 * np.pi, llc, lla, llz),
    }
    fn = fns[name]
    best = 1e30
    for _ in range(reps):
        cp(W, state)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s: %d rays, best of %d: %.3f ms = %.2f Grays/s" % (name, n, reps, best, n / best / 1e6))


if __name__ == "__main__":
    main()
