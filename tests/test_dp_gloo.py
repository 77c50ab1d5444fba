"""Host-side logic of data-parallel mode A on CPU with the gloo backend (world_size 2):
slicing, the packed all-reduce buffer (vae_b200.dist.FlatLayout / allreduce_flat) and the
decomposition it relies on -- sum over ranks of the local data-term gradients plus the KL
gradient built from the all-reduced batch counts and normalisers equals the single-process
gradient on the global batch.  The arithmetic here is the fp64 oracle; the CUDA
kernels are checked against the same decomposition in tests/test_gpu_dp.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as gu
from oracle import vfm_math
from vae_b200.dist import DP_TAIL, T_NLL, T_RESID, T_SQERR, FlatLayout, allreduce_flat, local_slice


def _global_inputs():
    meta, g = gu.load("sampled_reg_d64")
    x, y = gu.batch_of(meta, g, 0)
    noise = [g[f"step0.noise{i}"] for i in range(3)]
    return meta, g, x, y, noise


def _local_noise(noise, uniq_global, uniq_local):
    idx = np.searchsorted(uniq_global, uniq_local)
    return [noise[0], noise[1][:, idx], noise[2][:, idx]]


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        meta, g, x, y, noise = _global_inputs()
        N, M, d = meta["N"], meta["M"], meta["d"]
        params = gu.sampled_math_params(gu.state(g, "init"))
        sl = local_slice(len(x), rank, world)
        xl, yl = x[sl], y[sl]
        uniq_g, uniq_l = np.unique(x), np.unique(xl)
        loc = vfm_math.sampled_step(params, xl, yl, _local_noise(noise, uniq_g, uniq_l), g["train_counts"],
                                    meta["n_train"] / world, [N, M], output=meta["output"], kl_scale=0.0)
        lay = FlatLayout(N + M, d)
        flat = torch.zeros(lay.numel, dtype=torch.float64)
        ge, gb, counts, tail = lay.views(flat)
        ge += torch.from_numpy(loc["grads"]["entity"])
        gb += torch.from_numpy(loc["grads"]["bias"])
        counts[torch.from_numpy(uniq_l)] = torch.from_numpy(loc["plan"]["counts"].astype(np.float64))
        tail[:2] = torch.from_numpy(loc["z"])
        tail[T_NLL], tail[T_SQERR] = float(loc["nll_sum"]), float(loc["sq_err"])
        tail[T_RESID] = float(loc["resid"].sum())
        allreduce_flat(flat)
        if rank == 0:
            out_q.put(flat.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_two_rank_decomposition_equals_global_step():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    meta, g, x, y, noise = _global_inputs()
    N, M, d = meta["N"], meta["M"], meta["d"]
    lay = FlatLayout(N + M, d)
    ge, gb, counts, tail = lay.views(torch.from_numpy(flat))
    glob = vfm_math.sampled_step(gu.sampled_math_params(gu.state(g, "init")), x, y, noise, g["train_counts"],
                                 meta["n_train"], [N, M], output=meta["output"])
    # reduced integers / scalars are the global ones
    uniq = glob["plan"]["uniq"]
    assert np.array_equal(np.nonzero(counts.numpy())[0], uniq)
    assert np.array_equal(counts.numpy()[uniq], glob["plan"]["counts"])
    np.testing.assert_allclose(tail[:2].numpy(), glob["z"], rtol=1e-12)
    np.testing.assert_allclose(tail[T_NLL].item(), glob["nll_sum"], rtol=1e-12)
    # data-term sum + KL gradient from the reduced counts == global gradient
    c = glob["kl_weight"]
    P = gu.sampled_math_params(gu.state(g, "init"))
    mu, rho = P["entity"][uniq, :d].astype(np.float64), P["entity"][uniq, d:].astype(np.float64)
    a, b = P["bias"][uniq, 0].astype(np.float64), P["bias"][uniq, 1].astype(np.float64)
    full_e = ge.numpy().copy()
    full_e[uniq, :d] += c[:, None] * mu
    full_e[uniq, d:] += np.sign(rho) * c[:, None] * (np.abs(rho) - 1 / np.abs(rho))
    full_b = gb.numpy().copy()
    full_b[uniq, 0] += c * a
    full_b[uniq, 1] += np.sign(b) * c * (np.abs(b) - 1 / np.abs(b))
    np.testing.assert_allclose(full_e, glob["grads"]["entity"], rtol=1e-9, atol=1e-9 * np.abs(full_e).max())
    np.testing.assert_allclose(full_b, glob["grads"]["bias"], rtol=1e-9, atol=1e-9 * np.abs(full_b).max())
    np.testing.assert_allclose(tail[T_RESID].item() + float(P["global_bias_mean"][0]), glob["grads"]["global_bias_mean"][0],
                               rtol=1e-9)


def test_layout_and_slices():
    lay = FlatLayout(10, 4)
    assert lay.numel == 10 * 8 + 10 * 2 + 10 + DP_TAIL
    flat = torch.arange(lay.numel, dtype=torch.float32)
    ge, gb, counts, tail = lay.views(flat)
    assert ge.shape == (10, 8) and gb.shape == (10, 2) and counts.shape == (10,) and tail.shape == (DP_TAIL,)
    assert ge.data_ptr() == flat.data_ptr() and tail[-1] == lay.numel - 1
    assert [local_slice(8, r, 2) for r in range(2)] == [slice(0, 4), slice(4, 8)]
    with pytest.raises(AssertionError):
        local_slice(7, 0, 2)


# ---------------------------------------------------------------------------------------------
# mode B host logic: owner bucketing + fixed-shape all-to-all (gloo, world_size 2)
def _a2a_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vae_b200.dist import TorchExchange, bucket_by_owner
        rng = np.random.default_rng(rank)
        U_cap, CAP = 40, 24
        ids = np.sort(rng.choice(100, size=30, replace=False)).astype(np.int32)
        cnt = rng.integers(1, 9, size=30).astype(np.int32)
        pad = np.full(U_cap - 30, -7, dtype=np.int32)                     # garbage beyond n_valid
        send, dest, over = bucket_by_owner(torch.from_numpy(np.concatenate([ids, pad])),
                                           torch.from_numpy(np.concatenate([cnt, pad])), torch.tensor(30), world, CAP)
        assert not bool(over)
        recv = TorchExchange().all_to_all(send[: world * CAP].view(world, CAP, 2).contiguous())
        # every received real id is owned by this rank and carries its sender's count
        real = recv[..., 0] >= 0
        assert bool(((recv[..., 0][real] % world) == rank).all())
        # echo the ids back through the same slots: the requester recovers its list via `dest`
        back = TorchExchange().all_to_all(recv)
        got = back.reshape(world * CAP, 2)[dest[:30]]
        assert np.array_equal(got[:, 0].numpy(), ids) and np.array_equal(got[:, 1].numpy(), cnt)
        t = TorchExchange().all_reduce(torch.tensor([float(rank + 1)]))
        assert t.item() == 3.0
        if rank == 0:
            out_q.put("ok")
    finally:
        dist.destroy_process_group()


def test_mode_b_bucketing_and_all_to_all_round_trip():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_a2a_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    assert q.get(timeout=120) == "ok"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0


def test_bucket_overflow_is_flagged():
    from vae_b200.dist import bucket_by_owner
    ids = torch.arange(0, 20, 2, dtype=torch.int32)                       # all owned by rank 0 of 2
    send, dest, over = bucket_by_owner(ids, torch.ones_like(ids), torch.tensor(10), 2, 4)
    assert bool(over) and int((dest == 8).sum()) == 6                     # 6 ids went to the dump slot
    assert send[:4, 0].tolist() == [0, 2, 4, 6] and send[4:8, 0].tolist() == [-1] * 4
