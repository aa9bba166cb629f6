"""Multi-rank host logic on CPU (gloo, world_size 2): class-batch sharding + the single all-reduce of [A || B].

The per-class-batch factor increments come from the oracle here (the kernels need a GPU); what is under test is the
product's sharding schedule and reduction (bayesvlm_b200.hessians.class_batch_schedule / reduce_factors), i.e. that
rank-sharded estimation followed by ONE all-reduce reproduces the single-process result of the reference loop."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import laplace_oracle as O

LS = math.log(100.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_cb, ncls, bs, emb_s, act_s, emb_t, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bayesvlm_b200.hessians import class_batch_schedule, reduce_factors

        d_in, d = act_s.shape[1], emb_s.shape[1]
        A = torch.zeros(d_in, d_in, dtype=torch.float64)
        B = torch.zeros(d, d, dtype=torch.float64)
        mine = class_batch_schedule(n_cb, rank, world)
        for i in mine:
            lo, hi = i * ncls, (i + 1) * ncls
            used = (ncls // bs) * bs
            B += torch.from_numpy(O.infonce_ggn_collapsed(emb_s[lo:lo + used], emb_t[lo:hi], LS))
            A += torch.from_numpy(act_s[lo:hi].astype(np.float64).T @ act_s[lo:hi].astype(np.float64))
        A, B = reduce_factors(A, B)
        n = n_cb * ncls
        if rank == 0:
            out["A"], out["B"], out["mine"] = (A / math.sqrt(n)).numpy(), (B / math.sqrt(n)).numpy(), mine
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_kfac_equals_single_process():
    rng = np.random.default_rng(7)
    n_cb, ncls, bs, d, d_in = 5, 32, 5, 12, 10
    n = n_cb * ncls + 7  # remainder class batch is dropped (hessian_estimation.py:55)
    z = rng.standard_normal((n, d))
    emb_s = (z + 1.5 * rng.standard_normal((n, d))).astype(np.float32)
    emb_t = (z + 1.5 * rng.standard_normal((n, d))).astype(np.float32)
    act_s = rng.standard_normal((n, d_in)).astype(np.float32)
    A_ref, B_ref = O.kfac_ggn(emb_s, act_s, emb_t, ncls, bs, LS)
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_cb, ncls, bs, emb_s, act_s, emb_t, out), nprocs=world, join=True)
        assert out["mine"] == [0, 1, 2]  # contiguous blocks: rank 0 owns 3 of the 5 class batches
        np.testing.assert_allclose(out["A"], A_ref, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(out["B"], B_ref, rtol=1e-9, atol=1e-12)


def test_schedule_partitions_class_batches():
    from bayesvlm_b200.hessians import class_batch_schedule

    for n_cb in (1, 7, 32):
        for world in (1, 2, 4, 8):
            parts = [class_batch_schedule(n_cb, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n_cb))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_reduce_factors_is_identity_without_process_group():
    from bayesvlm_b200.hessians import reduce_factors

    A, B = torch.eye(3), torch.ones(2, 2)
    A2, B2 = reduce_factors(A, B)
    assert A2 is A and B2 is B
