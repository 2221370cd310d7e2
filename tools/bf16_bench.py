"""Times the bf16 gradient kernels (tcgen05 vs CUDA-core ring) on a config-4-shaped shard."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ds = S.Context(devices=[0]).synthetic(rows, d, S.LeastSquaresGradient(), seed=42, store="bf16")
w0 = np.zeros(d)
bytes_pass = rows * (d * 2 + 8)
def t(label, grad, **opts):
    for k, v in opts.items():
        ds.set_option(k, v)
    S.run_with_stats(ds, grad, S.SimpleUpdater(), 0.0, 1, 0.0, w0, 8.0, 8.0, 1.0, 0.9, False)
    _, h, st = S.run_with_stats(ds, grad, S.SimpleUpdater(), 0.0, 5, 0.0, w0, 8.0, 8.0, 1.0, 0.9, False)
    ms = st.k1_ms_total / st.k1_launches
    print(json.dumps(dict(label=label, rows=rows, d=d, **opts, k1_ms=round(ms, 3), gbs=round(bytes_pass / ms / 1e6, 1),
                          frac=round(bytes_pass / ms / 1e6 / 6566.1, 4), loss=h[-1])), flush=True)
t("tc LS", S.LeastSquaresGradient(), k1_variant="tc")
t("tc logistic", S.LogisticGradient())
for st_ in (9, 10, 12):
    t("tc LS", S.LeastSquaresGradient(), ring_stages=st_)
t("ring LS", S.LeastSquaresGradient(), k1_variant="ring", ring_stages=0)
