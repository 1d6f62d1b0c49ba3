"""Optional kernel variants behind environment switches (read once per process by the library) must give the results
of the default configuration: each runs in a child process on the same model / batch."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# switch -> (what it selects, outputs bit-identical to the default?)
VARIANTS = {
    "CGR_AP_BN=208": ("persistent atom projection with 208-wide slices / two operand stages", True),
    "CGR_AP_MC=2": ("persistent atom projection, weight chunks multicast over clusters of two CTAs", True),
    "CGR_AP_OLD=1": ("one-unit-per-CTA atom projection kernel", True),
    "CGR_AP_CG2=1": ("persistent atom projection on CTA pairs (tcgen05 cta_group::2, M = 256)", True),
    "CGR_NO_FUSED_TRAIN_FWD=1": ("training forward through the per-layer kernels", False),
    "CGR_BWD_FORK=1": ("weight-gradient GEMMs on a side stream", False),
}


@pytest.fixture(scope="module")
def default_results():
    from tests.variant_child import run
    return run()


@pytest.mark.parametrize("switch", sorted(VARIANTS))
def test_variant_matches_default(switch, default_results, tmp_path):
    name, value = switch.split("=")
    env = dict(os.environ)
    env[name] = value
    path = str(tmp_path / "variant.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_child.py"), path], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(path)
    ref = default_results
    scale = np.abs(ref["out"]).mean()
    if VARIANTS[switch][1]:
        assert np.array_equal(z["out"], ref["out"])            # same arithmetic per output element
    assert np.abs(z["out"] - ref["out"]).max() <= 1e-5 * scale
    assert np.abs(z["out_t"] - ref["out_t"]).max() <= 1e-5 * scale
    gmax = np.abs(ref["grads"]).max()
    assert np.abs(z["grads"] - ref["grads"]).max() <= 2e-5 * gmax
