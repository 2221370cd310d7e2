"""Times the CSR gradient path on a config-3-shaped shard (hinge, d = 1M, k stored entries per row): the row-major kernel
and the column-blocked, row-tiled kernels (option csr_format).  K1 time = everything between the events around a sweep
(for the tiled form: memsets + margin kernel + row kernel + gradient kernel)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ds = S.Context(devices=[0]).synthetic_csr(rows, d, k, S.HingeGradient(), seed=42, store="f32")
w0 = np.zeros(d)
bytes_pass = rows * k * 8 + rows * 16 + 8
res = {}
for fmt in ("rows", "tiles", "rows", "tiles"):
    ds.set_option("csr_format", fmt)
    for grad in (S.HingeGradient(), S.LogisticGradient()):
        t0 = time.time()
        S.run_with_stats(ds, grad, S.SquaredL2Updater(), 0.0, 1, 0.1, w0)      # first tiled call builds the twin
        t_first = time.time() - t0
        w, h, st = S.run_with_stats(ds, grad, S.SquaredL2Updater(), 0.0, 5, 0.1, w0, fuse=False)
        ms = st.k1_ms_total / st.k1_launches
        wf, hf, sf = S.run_with_stats(ds, grad, S.SquaredL2Updater(), 0.0, 5, 0.1, w0)
        rec = dict(format=fmt, grad=type(grad).__name__, rows=rows, d=d, k=k, k1_ms=round(ms, 3), gbs=round(bytes_pass / ms / 1e6, 1),
                   frac=round(bytes_pass / ms / 1e6 / 6566.1, 4), ms_per_pass=round(st.device_ms_total / st.passes, 3),
                   examples_per_s=round(rows * st.passes / st.device_ms_total * 1e3),
                   fused_examples_per_s=round(rows * sf.passes / sf.device_ms_total * 1e3), first_call_s=round(t_first, 2), loss=h[-1],
                   w_l2=float(np.linalg.norm(w)))
        print(json.dumps(rec), flush=True)
        res.setdefault(type(grad).__name__, {})[fmt] = (h[-1], np.linalg.norm(w))
for g, r in res.items():
    print(g, "loss rel diff rows vs tiles", abs(r["rows"][0] - r["tiles"][0]) / abs(r["rows"][0]), "w", abs(r["rows"][1] - r["tiles"][1]) / r["rows"][1])
