"""Parity at BASELINE.json's FULL sizes (round-1 review: every config was only compared shrunk): C2 / C3 at 50 k LiDAR
features, C4 with its 40-knot prior and the Schur complement of 20 knots, and C5 — the benchmarked window, 1 M LiDAR +
200 k IMU + 20 k visual — against the oracle on all host threads.  Bars as everywhere: indices bit-exact, H / b / cost
1e-9 relative, solved poses 1e-7 m / 1e-7 rad after the same iteration count."""
import os

import numpy as np
import pytest

from clic_b200 import estimator as E
from clic_b200 import synthetic as S
from test_gpu_parity import build, compare_eval, rel, rel_scaled
from test_gpu_solve import solve_both

pytestmark = pytest.mark.gpu


@pytest.fixture()
def oracle_mt(oracle):
    oracle.clic_oracle_set_threads(int(os.cpu_count() or 1))
    yield oracle
    oracle.clic_oracle_set_threads(1)


def test_c2_full(cuda, oracle_mt):
    sw = S.config(2)
    compare_eval(cuda, oracle_mt, sw, unlock_bias=True, unlock_toff=True)
    solve_both(cuda, oracle_mt, sw, 8, unlock_bias=True, unlock_toff=True)


def test_c3_full(cuda, oracle_mt):
    sw = S.config(3)
    compare_eval(cuda, oracle_mt, sw, unlock_bias=True)
    solve_both(cuda, oracle_mt, sw, 8, unlock_bias=True)


def test_c4_full_solve(cuda, oracle_mt):
    sw = S.config(4)
    compare_eval(cuda, oracle_mt, sw, unlock_bias=True)
    solve_both(cuda, oracle_mt, sw, 6, unlock_bias=True)


def test_c5_full_one_million_features(cuda, oracle_mt):
    """The benchmarked configuration itself."""
    sw = S.config(5)
    assert len(sw.lidar["t_point"]) == 1_000_000 and len(sw.imu["timestamp"]) == 200_000
    Hg, bg = compare_eval(cuda, oracle_mt, sw, unlock_bias=True)
    assert Hg.shape == (72, 72)
    solve_both(cuda, oracle_mt, sw, 8, unlock_bias=True)
