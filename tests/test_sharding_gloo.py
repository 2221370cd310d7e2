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
        # (3) the handle exchange closure of transport="ipc" (agd_xchg_export -> all-gather in rank order -> agd_xchg_import):
        # every rank must end up with every rank's blob, concatenated by rank, and the slot arithmetic of the exchange
        # (fixed stride 2 (d + 4) per rank slot, two parity buffers) must keep sweeps of both payload sizes apart
        blob = bytes([rank]) * S._native.XCHG_HANDLE_BYTES
        gathered = [None] * world
        dist.all_gather_object(gathered, blob)
        everyone = b"".join(gathered)
        assert len(everyone) == world * S._native.XCHG_HANDLE_BYTES
        assert all(everyone[r * S._native.XCHG_HANDLE_BYTES] == r for r in range(world))
        stride = 2 * (d + 4)
        spans = {}
        for buf_i in (0, 1):
            for r in range(world):
                for payload in (d + 4, stride):
                    a = (buf_i * world + r) * stride
                    spans[(buf_i, r, payload)] = (a, a + payload)
        for k1, (a1, b1) in spans.items():
            for k2, (a2, b2) in spans.items():
                if k1[:2] != k2[:2]:
                    assert b1 <= a2 or b2 <= a1, (k1, k2)       # different (buffer, rank) slots never overlap
            assert b1 <= 2 * world * stride                         # inside the allocation xchg_alloc makes
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
