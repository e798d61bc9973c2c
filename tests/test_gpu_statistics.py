"""Statistical agreement of the default (counter-based Philox) Langevin noise with the reference.

north_star: "thermostatted runs must agree statistically on the energy estimators within error bars". The reference's
own thermostatted regression cases (tests/cases/bosonic_quadratic_harmonic, dist_harmonic: 100 000 Langevin steps,
estimators every 100) run unchanged through pimdb_gpu with its DEFAULT noise stream, which shares nothing with the
reference's RANMAR sequence; the means of the estimator columns are compared with the means of the reference's golden
simulation.out (tests/golden/refcases.npz, made by tests/golden/make_fixtures.py) within k standard errors, each
series' standard error of the mean taken from a Flyvbjerg-Petersen blocking analysis (the series are correlated).
Reference: tests/main.py:77-107 (what the reference compares), src/thermostats/langevin.cpp:10-27.
"""
import subprocess

import numpy as np
import pytest

from pimd_b_b200 import io as pio
from tests.helpers import GOLDEN_DIR, ROOT

REFCASES = np.load(GOLDEN_DIR / "refcases.npz")
BIN = ROOT / "pimd_b_b200" / "pimdb_gpu"
COLUMNS = ("kinetic", "potential", "virial", "temperature", "cl_kinetic", "cl_spring")
SKIP_ROWS = 200          # equilibration from the random initial state (20 000 steps)
K_SIGMA = 4.5


def blocked_sem(a: np.ndarray) -> float:
    """Standard error of the mean of a correlated series: the largest of the naive estimates over block sizes 1, 2, 4, ...
    (blocks of at least 16; Flyvbjerg & Petersen, J. Chem. Phys. 91, 461 (1989))."""
    a = np.asarray(a, dtype=np.float64)
    best = 0.0
    while a.size >= 16:
        best = max(best, float(a.std(ddof=1) / np.sqrt(a.size)))
        a = 0.5 * (a[0:a.size - a.size % 2:2] + a[1:a.size - a.size % 2 + 1:2])
    return best


def test_blocking_analysis_on_a_known_series():
    rng = np.random.default_rng(0)
    white = rng.normal(size=4096)
    assert abs(blocked_sem(white) * np.sqrt(4096) - 1.0) < 0.35
    ar = np.empty(8192)                      # AR(1), rho = 0.9: the true error of the mean is sqrt((1+rho)/(1-rho)) times the naive one
    ar[0] = 0.0
    for i in range(1, ar.size):
        ar[i] = 0.9 * ar[i - 1] + rng.normal()
    naive = ar.std(ddof=1) / np.sqrt(ar.size)
    assert 2.5 < blocked_sem(ar) / naive < 6.5


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["bosonic_quadratic_harmonic", "dist_harmonic"])
def test_default_noise_stream_agrees_with_reference_within_error_bars(gpu_required, case, tmp_path, capsys):
    if not BIN.exists():
        from pimd_b_b200 import build
        build.build_host()
    (tmp_path / "config.ini").write_text(str(REFCASES[f"{case}/ini"]))
    r = subprocess.run([str(BIN), "-in", "config.ini", "--dim", "3"], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert "finished running successfully" in r.stdout, r.stdout + r.stderr
    got = pio.read_simulation_out(str(tmp_path / "output" / "simulation.out"))
    cols = [str(c) for c in REFCASES[f"{case}/simout_columns"]]
    ref = REFCASES[f"{case}/simout"]
    lines = []
    for c in COLUMNS:
        if c not in got:
            continue
        g, e = got[c][SKIP_ROWS:], ref[SKIP_ROWS:, cols.index(c)]
        assert g.shape == e.shape and g.size > 500
        mg, me, sg, se = g.mean(), e.mean(), blocked_sem(g), blocked_sem(e)
        sigma = np.hypot(sg, se)
        lines.append(f"{case:28s} {c:12s} gpu {mg:13.6e} +- {sg:9.2e}   reference {me:13.6e} +- {se:9.2e}   "
                     f"difference {abs(mg - me) / sigma:4.2f} sigma")
        assert abs(mg - me) <= K_SIGMA * sigma, lines[-1]
        # the fluctuations agree too (variance of the estimator, 25 %)
        assert 0.75 < g.std() / e.std() < 1.33, (c, g.std(), e.std())
    with capsys.disabled():
        print()
        for ln in lines:
            print("   ", ln)
    # the noise really is a different realisation: the trajectories must NOT coincide
    assert not np.allclose(got["kinetic"][1:50], ref[1:50, cols.index("kinetic")], rtol=1e-3)
