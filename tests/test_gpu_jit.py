"""Run-time specialised kernels (pxf_jit.cu) against the interpreter: same programs, same inputs, two processes,
identical bits -- rows, alive flags, centroid sums, surviving indices, compacted bundles, side arrays."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra):
    env = dict(os.environ)
    for k in ("PXF_JIT", "PXF_JIT_MIN_RAYS", "PXF_NO_SPECIALIZE"):
        env.pop(k, None)
    env.update(env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_jit_compare.py")], env=env, capture_output=True,
                       text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    out = {}
    for line in p.stdout.splitlines():
        f = line.split()
        if len(f) == 3:
            out[f[0]] = (f[1], f[2])
        elif len(f) == 2 and f[0] == "STATUS":
            out["STATUS"] = f[1]
    return out


def test_specialised_kernels_equal_the_interpreter_bit_for_bit():
    interp = run({"PXF_JIT": "0"})
    jit = run({"PXF_JIT_MIN_RAYS": "0", "PXF_JIT_VERBOSE": "1"})
    assert set(interp) == set(jit) and len(interp) >= 12
    for name in interp:
        if name == "STATUS":
            continue
        assert interp[name][0] == jit[name][0], "%s: specialised kernel and interpreter differ" % name
        assert "interpreter" in interp[name][1] or "k_chain" in interp[name][1], (name, interp[name][1])
    # the specialised path really ran (and not the interpreter) wherever no built-in chain exists
    ran_jit = [n for n in jit if n != "STATUS" and "pxf_jit_chain" in jit[n][1]]
    assert len(ran_jit) >= 10, (ran_jit, jit.get("STATUS"))
    assert not any("interpreter" in jit[n][1] for n in jit if n != "STATUS"), jit
