"""Runs a few applySmooth passes on a synthetic shard for profiling under ncu.
usage: [AGD_PAIR=1 | AGD_TWO=1] [AGD_OPTS=key=value,...] python tools/k1_prof.py <logistic|least_squares|hinge> [rows] [d] [store] [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import spark_agd_b200 as S  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "logistic"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
store = sys.argv[4] if len(sys.argv) > 4 else "f32"
passes = int(sys.argv[5]) if len(sys.argv) > 5 else 4
grad = {"logistic": S.LogisticGradient(), "least_squares": S.LeastSquaresGradient(), "hinge": S.HingeGradient()}[kind]
ds = S.Context(devices=[0]).synthetic(rows, d, grad, seed=42, store=store)
for k, v in (kv.split("=") for kv in os.environ.get("AGD_OPTS", "").split(",") if kv):
    ds.set_option(k, v)
w = np.random.default_rng(0).standard_normal(d) * 0.02
w2 = w + np.random.default_rng(1).standard_normal(d) * 0.01
for _ in range(passes):
    if os.environ.get("AGD_TWO"):       # the speculative sweep of the memoised pass structure: two complete evaluations
        loss, g, cnt, loss2, g2 = ds.smooth_two(grad, w, w2)
    elif os.environ.get("AGD_PAIR"):    # the fused sweep of agd_run: applySmooth at w + the loss at w2
        loss, g, cnt, loss2 = ds.smooth_pair(grad, w, w2)
    else:
        loss, g, cnt = ds.smooth(grad, w)
print(kind, rows, d, store, "loss", loss, "count", cnt)
