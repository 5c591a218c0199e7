"""Writes tests/golden/factors_extra.npz: residual and Jacobian of every single-factor case of tests/factor_cases.py
(the reference's test_analytic_factor.cpp inputs, seeds 1-3) from the oracle — which tests/test_oracle_factors_extra.py
pins to the torch-f64 auto-diff lineage.  Run from the repo root:  python tests/golden/make_golden_factors.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import factor_cases as FC  # noqa: E402
import oracle_lib  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from test_oracle_autodiff import reference_test_trajectory  # noqa: E402


def main():
    lib = oracle_lib.load()
    out = {}
    for seed in (1, 2, 3):
        w = reference_test_trajectory(seed)
        est = E.TrajectoryEstimator(w, E.default_options(lib), lib)
        for c in FC.cases(w, seed):
            r, J = FC.evaluate(est, c)
            out[f"s{seed}/{c['name']}/r"] = r
            out[f"s{seed}/{c['name']}/J"] = J
        est.close()
    path = os.path.join(HERE, "factors_extra.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out) // 2, "cases")


if __name__ == "__main__":
    main()
