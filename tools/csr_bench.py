"""Times the CSR gradient kernel on a config-3-shaped shard (hinge, d = 1M, k stored entries per row)."""
import json, os, sys
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
for grad, opt in ((S.HingeGradient(), 0), (S.HingeGradient(), 1), (S.HingeGradient(), 0), (S.HingeGradient(), 1), (S.LogisticGradient(), 0)):
    ds.set_option("ring_rows", opt)      # 0 = pipelined row loop (default), 1 = simple loop
    S.run_with_stats(ds, grad, S.SquaredL2Updater(), 0.0, 1, 0.1, w0)
    w, h, st = S.run_with_stats(ds, grad, S.SquaredL2Updater(), 0.0, 5, 0.1, w0, fuse=False)
    ms = st.k1_ms_total / st.k1_launches
    wf, hf, sf = S.run_with_stats(ds, grad, S.SquaredL2Updater(), 0.0, 5, 0.1, w0)
    print(json.dumps(dict(grad=type(grad).__name__, simple_loop=opt, rows=rows, d=d, k=k, k1_ms=round(ms, 3), gbs=round(bytes_pass / ms / 1e6, 1),
                          frac=round(bytes_pass / ms / 1e6 / 6566.1, 4), ms_per_pass=round(st.device_ms_total / st.passes, 3),
                          examples_per_s=round(rows * st.passes / st.device_ms_total * 1e3),
                          fused_examples_per_s=round(rows * sf.passes / sf.device_ms_total * 1e3), loss=h[-1])), flush=True)
