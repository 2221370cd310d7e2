"""Thread scaling of the CPU oracle on this host (diagnostic for bench.py's reference arm).
usage: python tools/cpu_scaling.py [rows_per_thread] [d]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O

rpt = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ncpu = O.host_threads()
out = {"loadavg": open("/proc/loadavg").read().split()[:3], "cpus": ncpu, "rows_per_thread": rpt, "d": d, "runs": []}
for T in [1, 8, 16, 32, 64, 96, 128]:
    if T > ncpu:
        continue
    for bind in (True, False):
        rows = rpt * T
        if bind:
            O.bind_threads(T)
        else:
            O.unbind_threads(); O.lib().oracle_set_threads(T)
        X = O.synth_dense_f32_placed(42, 0, rows, d, T, T)
        y = O.synth_labels(42, "logistic", 0, X, O.synth_wtrue(42, d))
        D = O.Data(y, X=X)
        O.agd_run(D, "logistic", "simple", np.zeros(d), convergence_tol=0.0, num_iterations=1, partitions=T, threads=T)
        t0 = time.perf_counter()
        r = O.agd_run(D, "logistic", "simple", np.zeros(d), convergence_tol=0.0, num_iterations=3, partitions=T, threads=T)
        dt = time.perf_counter() - t0
        O.unbind_threads()
        out["runs"].append({"threads": T, "pinned": bind, "row_passes_per_s": rows * r.passes / dt,
                            "per_thread": rows * r.passes / dt / T})
        del X, D
out["loadavg_after"] = open("/proc/loadavg").read().split()[:3]
print(json.dumps(out))
