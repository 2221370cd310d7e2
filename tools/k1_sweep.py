"""Times the K1 gradient kernel variants on the BASELINE config-2 shard (run on the GPU box).
usage: python tools/k1_sweep.py [rows] [d]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import spark_agd_b200 as S  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
store = sys.argv[3] if len(sys.argv) > 3 else "f32"
peak = 6566.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ctx = S.Context(devices=[0])
ds = ctx.synthetic(rows, d, S.LogisticGradient(), seed=42, store=store)
eb = 4 if store == "f32" else 8
bytes_pass = rows * d * eb + rows * 8
w0 = np.zeros(d)
out = []
for (r, c, s) in [(8, 2, 0), (4, 3, 0)]:
    ds.set_option("ring_rows", r); ds.set_option("ring_ctas", c); ds.set_option("ring_stages", s)
    for grad in (S.LogisticGradient(), S.LeastSquaresGradient()):
        S.run_with_stats(ds, grad, S.SimpleUpdater(), 0.0, 2, 0.0, w0)  # warm-up
        w, h, st = S.run_with_stats(ds, grad, S.SimpleUpdater(), 0.0, 6, 0.0, w0)
        ms = st.k1_ms_total / st.k1_launches
        gbs = bytes_pass / ms / 1e6
        rec = dict(rows_per_tile=r, ctas=c, stages=s, grad=type(grad).__name__, k1_ms=round(ms, 4), gbs=round(gbs, 1),
                   frac=round(gbs / peak, 4), passes=st.passes, total_s=round(st.seconds_total, 4))
        print(json.dumps(rec), flush=True)
        out.append(rec)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "k1_sweep.json"), "w"), indent=1)
