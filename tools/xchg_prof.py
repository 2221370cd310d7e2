"""Short 2-GPU run in ONE process (Context(devices=[0, 1])) for ncu: the fused reduce + publish kernel
(k1_reduce_kernel<true>: fixed-order slab sums stored straight into the peer's HBM over NVLink) and xchg_gather_kernel.
One process so that ncu's kernel serialisation cannot deadlock the flag wait: every publish is launched before any gather.
usage: ncu --set full --section Nvlink -k regex:"k1_reduce_kernel|xchg_gather" -c 8 -o gpurun_out/xchg python tools/xchg_prof.py [rows] [d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ctx = S.Context(devices=[0, 1])
ds = ctx.synthetic(rows, d, S.LogisticGradient(), seed=42, store="f32")
w, hist, st = S.run_with_stats(ds, S.LogisticGradient(), S.SimpleUpdater(), 0.0, 3, 0.0, np.zeros(d))
print("collective_kind", st.collective_kind, "collective_calls", st.collective_calls, "loss", hist[-1])
ds.close()
