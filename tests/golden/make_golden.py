"""Generates the committed golden vectors from the oracle (run in the build container: python tests/golden/make_golden.py).
Each .npz holds the inputs' identity (generator arguments) and the oracle outputs for one small window, so the -m gpu
tests can check the CUDA path on a box where only the prebuilt oracle .so travels, and so oracle regressions are caught.
The reference has no numeric golden vectors of its own: these are oracle outputs, the oracle itself is
pinned by tests/test_oracle_autodiff.py."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib  # noqa: E402
from clic_b200 import estimator as E  # noqa: E402
from clic_b200 import synthetic as S  # noqa: E402

CASES = {
    # name: (make_window kwargs, estimator setup)
    "c1_small": (dict(n_knots=10, n_lidar=400, n_imu=60, seed=101), dict(fixed=-1, unlock_bias=False, bias_factor=False)),
    "c2_small": (dict(n_knots=10, n_lidar=600, n_imu=60, line_fraction=0.3, seed=102, t_offset_ns=(1.0e6, 0.0, 0.0),
                      time_margin_ns=12_000_000), dict(fixed=-1, unlock_bias=True, unlock_toff=True, bias_factor=True)),
    "c3_small": (dict(n_knots=10, n_lidar=500, n_imu=60, n_landmarks=25, obs_per_landmark=8, line_fraction=0.3, seed=103),
                 dict(fixed=2, unlock_bias=True, bias_factor=True)),
    "c4_small": (dict(n_knots=40, n_lidar=1500, n_imu=200, seed=104), dict(fixed=-1, unlock_bias=True, bias_factor=True)),
}


def build(lib, sw, fixed=-1, unlock_bias=False, unlock_toff=False, bias_factor=True):
    opt = E.default_options(lib)
    if unlock_bias:
        opt.lock_ab = opt.lock_wb = 0
    if unlock_toff:
        opt.lock_t_offset[0] = 0
        opt.t_offset_padding_ns = 10_000_000
    est = E.TrajectoryEstimator(sw.window, opt, lib)
    est.SetFixedIndex(fixed)
    S.add_all_factors(est, sw, bias_factor=bias_factor)
    return est


def main():
    lib = oracle_lib.load()
    for name, (wkw, ekw) in CASES.items():
        sw = S.make_window(**wkw)
        est = build(lib, sw, **ekw)
        H, b, cost, fcost = est.Evaluate()
        idx = {f"s{k}": est.indices(k)[0] for k in range(4)}
        idx.update({f"u{k}": est.indices(k)[1] for k in range(4)})
        summ = est.Solve(8)
        x = est.read_state()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), H=H, b=b, cost=cost, fixed_cost=fcost,
                            knot_q=x["knot_q"], knot_p=x["knot_p"], bias=x["bias"], t_offset_ns=x["t_offset_ns"],
                            summary=json.dumps(summ), window_kwargs=json.dumps(wkw), estimator_kwargs=json.dumps(ekw),
                            **idx)
        print(name, "D", H.shape[0], "cost", cost, "iters", summ["iterations"])


if __name__ == "__main__":
    main()
