"""AGD_TRACE=1 diagnostic: per-kernel time (gap before the launch + run time) of default and memoised runs.
usage (N ranks): AGD_TRACE=1 python -m torch.distributed.run --nproc-per-node N tools/trace_run.py [rows] [d] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import spark_agd_b200 as S
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = S.Context.from_torch_distributed(local)
else:
    ctx = S.Context(devices=[local])
ds = ctx.synthetic(rows, d, S.LogisticGradient(), seed=42, store="f32")
g, u, w0 = S.LogisticGradient(), S.SimpleUpdater(), np.zeros(d)
for memo in (False, True, False, True):
    if world > 1:
        torch.cuda.synchronize(); dist.barrier()
    w, h, st = S.run_with_stats(ds, g, u, 0.0, iters, 0.0, w0, memoize=memo)
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"memo={memo} iters/s {st.iterations / st.device_ms_total * 1e3:.2f} sweeps {st.k1_launches} k1 {st.k1_ms_total / st.k1_launches:.3f} ms device {st.device_ms_total:.2f} ms", flush=True)
ds.close()
if world > 1:
    dist.destroy_process_group()
