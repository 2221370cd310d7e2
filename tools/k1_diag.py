"""Bisects the K1 ring: stream-only / +phase1 / full kernels, a few ring depths (run on the GPU box)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S
rows, d = 10_000_000, 1024
ds = S.Context(devices=[0]).synthetic(rows, d, S.LogisticGradient(), seed=42, store="f32")
bytes_pass = rows * (d * 4 + 8)
w0 = np.zeros(d)
def t(label, grad=S.LogisticGradient(), **opts):
    for k, v in opts.items():
        ds.set_option(k, v)
    S.run_with_stats(ds, grad, S.SimpleUpdater(), 0.0, 1, 0.0, w0)
    _, _, st = S.run_with_stats(ds, grad, S.SimpleUpdater(), 0.0, 4, 0.0, w0)
    ms = st.k1_ms_total / st.k1_launches
    print(json.dumps(dict(label=label, **opts, k1_ms=round(ms, 3), gbs=round(bytes_pass / ms / 1e6, 1),
                          frac=round(bytes_pass / ms / 1e6 / 6566.1, 4))), flush=True)
for rep in range(3):
    t("ring logistic", k1_variant="ring")
    t("ring LS", grad=S.LeastSquaresGradient())
    t("ring hinge", grad=S.HingeGradient())
