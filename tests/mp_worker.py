"""One rank of a multi-process world (spawned by tests/test_multirank_gpu.py; not a test module itself).

  python tests/mp_worker.py RANK WORLD PORT DEVICE TRANSPORT OUT.json

Every rank loads its contiguous slice of a seeded host dataset (and, second, generates its slice of the synthetic
workload in place), runs applySmooth and the AGD loop through the C-ABI mirror, and rank 0 writes the results."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def make_data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    w_true = rng.standard_normal(d) / np.sqrt(d)
    y = (X.astype(np.float64) @ w_true + rng.logistic(size=n) > 0).astype(np.float64)
    return X, y


def main():
    rank, world, port, dev, transport, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(dev)
    backend = "nccl" if transport == "nccl" else "gloo"
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", dev)} if backend == "nccl" else {}))
    import spark_agd_b200 as S
    ctx = S.Context.from_torch_distributed(dev, transport=transport)
    res = {}
    n, d = 6001, 1024
    X, y = make_data(n, d, 7)
    lo, hi = rank * n // world, (rank + 1) * n // world
    data = ctx.parallelize(y[lo:hi], X[lo:hi], store="f32")
    rng = np.random.default_rng(11)
    w = rng.standard_normal(d) * 0.1
    loss, g, cnt = data.smooth(S.LogisticGradient(), w)
    res["smooth"] = {"loss": loss, "grad": g.tolist(), "count": cnt}
    wf, hist, st = S.run_with_stats(data, S.LogisticGradient(), S.SquaredL2Updater(), 0.0, 8, 0.01, np.zeros(d))
    res["run"] = {"w": wf.tolist(), "hist": hist.tolist(), "passes": st.passes, "backtracks": st.backtracks,
                  "restarts": st.restarts, "collective_kind": st.collective_kind, "collective_calls": st.collective_calls}
    # the other pass structures exchange different payloads (d + 4 or 2 (d + 4) doubles per sweep) through the same buffers:
    # they must give the same bits on every rank
    wm, hm, sm = S.run_with_stats(data, S.LogisticGradient(), S.SquaredL2Updater(), 0.0, 8, 0.01, np.zeros(d), memoize=True)
    wu, hu, su = S.run_with_stats(data, S.LogisticGradient(), S.SquaredL2Updater(), 0.0, 8, 0.01, np.zeros(d), fuse=False)
    same = bool(np.array_equal(wm, wf) and np.array_equal(hm, hist) and np.array_equal(wu, wf) and np.array_equal(hu, hist))
    flags = [None] * world
    dist.all_gather_object(flags, same)
    res["modes_bit_identical_on_every_rank"] = bool(all(flags))
    res["memo_fused_passes"] = sm.fused_passes
    # a second dimension on the same handle: the exchange is rebuilt (export / import again under transport="ipc")
    data.unpersist()
    d2 = 260
    X2, y2 = make_data(3000, d2, 9)
    lo, hi = rank * 3000 // world, (rank + 1) * 3000 // world
    data.load_dense(y2[lo:hi], X2[lo:hi], store="f64")
    loss2, g2, cnt2 = data.smooth(S.LeastSquaresGradient(), np.full(d2, 0.01))
    res["smooth_d2"] = {"loss": loss2, "grad": g2.tolist(), "count": cnt2}
    data.close()
    # the synthetic workload generated in place: every rank owns rows [r*n/W, (r+1)*n/W) of the same global matrix
    syn = ctx.synthetic(20000, 512, S.LogisticGradient(), seed=42, store="f32")
    ws, hs, ss = S.run_with_stats(syn, S.LogisticGradient(), S.SimpleUpdater(), 0.0, 5, 0.0, np.zeros(512))
    res["synthetic"] = {"w": ws.tolist(), "hist": hs.tolist(), "passes": ss.passes, "rows_local": syn.local_rows(0)}
    syn.close()
    if rank == 0:
        with open(out, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
