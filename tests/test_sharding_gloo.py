"""N>1 host logic on CPU with gloo, world_size 2: the row partition every rank derives, the
unique-id exchange used by Context.from_torch_distributed, and the algebra of the one exchange step
(sum of per-shard [grad | loss | count] then normalise == applySmooth over all rows,
AGD.scala:196-207) with the oracle standing in for the GPU kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import spark_agd_b200 as S
        from oracle import oracle as O
        # (1) the id exchange closure of Context.from_torch_distributed
        buf = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(buf, src=0)
        assert buf[0] == bytes(range(128))
        # (2) row partition used by agd_generate / bench.py: rank r owns [r*n/W, (r+1)*n/W)
        n, d = 1001, 24
        lo, hi = (rank * n) // world, ((rank + 1) * n) // world
        X = O.synth_dense_f32(42, 0, n, d)
        w_true = O.synth_wtrue(42, d)
        y = O.synth_labels(42, "logistic", 0, X, w_true)
        w = np.linspace(-0.5, 0.5, d)
        # local shard statistics, un-normalised: [grad_sum | loss_sum | count]
        l, g, c = O.smooth(O.Data(y[lo:hi], X=X[lo:hi]), "logistic", w, partitions=1)
        packed = torch.from_numpy(np.concatenate([g * c, [l * c, float(c)]]))
        dist.all_reduce(packed)                        # the ONE exchange step (ncclAllReduce on the GPU path)
        cnt = packed[-1].item()
        got_loss, got_grad = packed[-2].item() / cnt, packed[:-2].numpy() / cnt
        ref_loss, ref_grad, ref_cnt = O.smooth(O.Data(y, X=X), "logistic", w, partitions=world)
        assert cnt == ref_cnt == n
        np.testing.assert_allclose(got_loss, ref_loss, rtol=1e-13)
        np.testing.assert_allclose(got_grad, ref_grad, rtol=1e-11, atol=1e-15)
        out[rank] = (lo, hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_exchange():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    spans = [out[r] for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == 1001 and spans[0][1] == spans[1][0]
