"""A few applySmooth passes over a synthetic CSR shard, for profiling k1_csr_kernel under ncu.
usage: python tools/csr_prof.py [rows] [d] [nnz_per_row]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spark_agd_b200 as S
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ds = S.Context(devices=[0]).synthetic_csr(rows, d, k, S.HingeGradient(), seed=42, store="f32")
w = np.random.default_rng(0).standard_normal(d) * 0.01
for _ in range(4):
    loss, g, cnt = ds.smooth(S.HingeGradient(), w)
print("csr", rows, d, k, "loss", loss, "count", cnt)
