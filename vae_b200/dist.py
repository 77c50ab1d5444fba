"""Multi-GPU execution of the VFM step (one process per GPU, ``torch.distributed`` / NCCL).

The reference is single-process, so these modes have no reference counterpart;
their contract is: the step on the GLOBAL batch (the concatenation of the ranks' local batches)
equals the single-process step on that batch -- same per-entity noise (Philox is keyed by row id
and step, never by rank or batch position), same global batch counts in the KL weights.

Mode A -- ``DataParallelSampled``: replicated tables, batch data parallel, ONE dense all-reduce
per step over ``[grad_entity | grad_bias | batch counts | tail]``.  Exact reference semantics
(including dense Adam over every row) at a cost of ``R*(2d+3)*4`` bytes per step, so it is the
mode for small tables (BASELINE configs 1-2).  Mode B -- ``ShardedSampled``: row-sharded tables (``owner = row mod P``), three fixed-shape
all-to-alls per step (unique ids, sampled rows, row gradients) and a 16-float all-reduce; what
scales (BASELINE configs 3-5).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib as L
from .engine import current_stream, make_config

DP_TAIL = 16
T_NLL, T_RESID, T_SQERR = 8, 9, 10


@dataclass(frozen=True)
class FlatLayout:
    """Layout of the all-reduced fp32 buffer: dense entity gradient, dense bias gradient, dense
    per-row batch counts, then a 16-float tail {Z_f (8), sum nll, sum resid, sum sq. err}."""
    R: int
    d: int

    @property
    def n_entity(self) -> int:
        return self.R * 2 * self.d

    @property
    def off_bias(self) -> int:
        return self.n_entity

    @property
    def off_counts(self) -> int:
        return self.off_bias + 2 * self.R

    @property
    def off_tail(self) -> int:
        return self.off_counts + self.R

    @property
    def numel(self) -> int:
        return self.off_tail + DP_TAIL

    def views(self, flat: torch.Tensor):
        return (flat[: self.n_entity].view(self.R, 2 * self.d),
                flat[self.off_bias: self.off_counts].view(self.R, 2),
                flat[self.off_counts: self.off_tail],
                flat[self.off_tail:])


def allreduce_flat(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the packed buffer over the ranks (NCCL on GPUs; any backend in tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def local_slice(n_global: int, rank: int, world: int):
    """Contiguous slice of the global batch owned by ``rank`` (batches are never shuffled)."""
    per = n_global // world
    assert per * world == n_global, "global batch must divide evenly over the ranks"
    return slice(rank * per, (rank + 1) * per)


class DataParallelSampled:
    """Mode A driver around a replicated ``vae_b200.vfm_torch.CF``."""

    def __init__(self, model, world: int, group=None, dense_adam: bool = True):
        self.model, self.world, self.group, self.dense_adam = model, int(world), group, bool(dense_adam)
        self.layout = FlatLayout(model.R, model.d)
        self.flat = torch.zeros(self.layout.numel, dtype=torch.float32, device=model.device)
        self._cfg_local, self._cfg_global = {}, {}

    def _configs(self, B_local: int):
        m = self.model
        if B_local not in self._cfg_local:
            mk = lambda B, n_train: make_config(B, m.F, m.d, m.R, m.S, m.output, m.link_name, m._class_bounds,
                                                m._class_sizes, n_train, m.seed)
            # local kernels scale residuals by n_train_eff / B_local = n_train / B_global
            self._cfg_local[B_local] = mk(B_local, m.n_train / self.world)
            self._cfg_global[B_local] = mk(B_local * self.world, m.n_train)
        return self._cfg_local[B_local], self._cfg_global[B_local]

    @torch.no_grad()
    def local_backward(self, x: torch.Tensor, y: torch.Tensor,
                       noise: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """Forward + data-term backward on the local slice; returns the packed buffer to reduce.
        ``noise`` (tests) is indexed by the LOCAL unique rank, as in ``CF.fused_step``."""
        m, lay = self.model, self.layout
        m._sync_scalars()
        x = x.to(m.device).contiguous()
        y = y.to(m.device, torch.float32).contiguous()
        B = int(x.shape[0])
        cfg_l, _ = self._configs(B)
        m._ensure(B)
        m._cfg = cfg_l
        m._plan = m._pipe.acquire(cfg_l, x, m.train_counts)
        noise = m._prep_noise(noise)
        self.flat.zero_()
        g_entity, g_bias, counts, tail = lay.views(self.flat)
        io = m._buf.io(y=y, noise=noise, grad_bias=g_bias, grad_entity=g_entity)
        io.grad_scalars = None
        tab, s, lib = m._tables(), current_stream(m.device), L.lib()
        L.check(lib.vfmb_sampled_forward(C.byref(cfg_l), C.byref(tab), C.byref(m._plan.struct), C.byref(io), s),
                "vfmb_sampled_forward")
        L.check(lib.vfmb_sampled_backward(C.byref(cfg_l), C.byref(tab), C.byref(m._plan.struct), C.byref(io),
                                          C.byref(m.adam), L.GRAD_ONLY, 0.0, s), "vfmb_sampled_backward")
        L.check(lib.vfmb_dp_scatter_counts(C.byref(cfg_l), C.byref(m._plan.struct), C.byref(io),
                                           counts.data_ptr(), tail.data_ptr(), s), "vfmb_dp_scatter_counts")
        m._pipe.release(m._plan)
        self._B_local = B
        self._eps0 = noise[0] if noise is not None else None
        return self.flat

    @torch.no_grad()
    def apply(self, flat: torch.Tensor) -> dict:
        """KL gradient with global counts + Adam + scalar update from the reduced buffer."""
        m, lay = self.model, self.layout
        _, cfg_g = self._configs(self._B_local)
        g_entity, g_bias, counts, tail = lay.views(flat)
        L.check(L.lib().vfmb_dp_apply_sampled(C.byref(cfg_g), C.byref(m._tables()), g_entity.data_ptr(),
                                              g_bias.data_ptr(), counts.data_ptr(), tail.data_ptr(),
                                              L.ptr(self._eps0), C.byref(m.adam), int(self.dense_adam),
                                              m._buf.partials.data_ptr(), m._buf.counters.data_ptr(),
                                              m._buf.stats.data_ptr(), current_stream(m.device)),
                "vfmb_dp_apply_sampled")
        st = m._buf.stats
        return {"loss": st[L.ST_LOSS], "kl": st[L.ST_KL], "nll_mean": st[L.ST_NLL_MEAN],
                "pred": m._buf.mean[: self._B_local], "stats": st}

    def step(self, x_local: torch.Tensor, y_local: torch.Tensor, noise=None) -> dict:
        flat = self.local_backward(x_local, y_local, noise)
        allreduce_flat(flat, self.group)
        return self.apply(flat)


# =============================================================================================
# Mode B -- row-sharded tables, all-to-all of sampled rows and of row gradients
# =============================================================================================
T_KLROWS = 11


def bucket_by_owner(ids: torch.Tensor, cnt: torch.Tensor, n_valid: torch.Tensor, P: int, CAP: int):
    """Place the first ``n_valid`` sorted unique ids into fixed-capacity per-owner slots.

    owner(id) = id mod P; inside an owner's bucket the ids keep their (ascending) order.
    Returns ``send [P*CAP+1, 2] int32`` rows {id, count} (-1 = empty; the last row is a dump slot),
    ``dest [len(ids)]`` = slot of every unique rank (dump slot for invalid / overflowing ones) and
    the overflow flag.  Pure tensor code, no host synchronisation (``n_valid`` stays on the device)."""
    M = P * CAP
    dev = ids.device
    valid = torch.arange(ids.numel(), device=dev) < n_valid
    owner = torch.where(valid, ids % P, torch.zeros_like(ids)).long()
    # [P, n] layout: the scan runs along the contiguous dimension (a [n, P] cumsum over dim 0 took 8 ms)
    onehot = (owner[None, :] == torch.arange(P, device=dev)[:, None]) & valid[None, :]
    ordinal = (onehot.cumsum(1, dtype=torch.int32) - 1).gather(0, owner[None, :]).squeeze(0).long()
    over = valid & (ordinal >= CAP)
    ok = valid & ~over
    dest = torch.where(ok, owner * CAP + ordinal, torch.full_like(owner, M))
    send = torch.full((M + 1, 2), -1, dtype=torch.int32, device=dev)
    send[dest] = torch.stack((ids.int(), cnt.int()), dim=1)
    send[M] = -1
    return send, dest, over.any()


class TorchExchange:
    """The collectives of mode B over torch.distributed (NCCL on GPUs)."""

    def __init__(self, group=None):
        self.group = group

    def all_to_all(self, send: torch.Tensor) -> torch.Tensor:
        """send[q] goes to rank q; returns recv with recv[q] = what rank q sent to this rank."""
        import torch.distributed as dist
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


SMALL_PITCH = 32                                            # floats per rank slot of the small-vector regions
N_SLOTS = 3                                                 # batches in flight in the pipelined loop (ShardedPipeline)


def _region_layout(M: int, d: int, P: int):
    """Regions of one rank's exchange buffer (bytes, 256-aligned): requests and Z_f once per pipeline slot (the
    requests of batches i+1 and i+2 travel while step i runs), sampled rows (+ one spare slot of zeros that overflowed rows are
    routed to), row gradients, the ranks' additive scalars.  Slot pitch of rows / gradients: d + 4 floats."""
    sp = d + 4
    sizes = [(f"ids{k}", M * 2 * 4) for k in range(N_SLOTS)] + [("rows", (M + 1) * sp * 4), ("grads", M * sp * 4)] + \
            [(f"z{k}", P * SMALL_PITCH * 4) for k in range(N_SLOTS)] + [("tail", P * SMALL_PITCH * 4)]
    off, total = {}, 0
    for name, nbytes in sizes:
        off[name] = total
        total += (nbytes + 255) // 256 * 256
    return off, total


class _ExchangeViews:
    """Typed views of a rank's own exchange buffer + the pointer tables over all ranks' buffers."""

    def _bind(self, buf: torch.Tensor, ptrs, M: int, d: int, P: int, rank: int):
        self.P, self.rank, self.M, self.d, self.SP = P, rank, M, d, d + 4
        self.buf = buf
        self.tables = {name: (C.c_void_p * P)(*[p + o for p in ptrs]) for name, o in self.off.items()}
        view = lambda name, n, dt: buf[self.off[name]: self.off[name] + n * 4].view(dt)
        self.ids = [view(f"ids{k}", M * 2, torch.int32).view(M, 2) for k in range(N_SLOTS)]
        self.rows = view("rows", (M + 1) * self.SP, torch.float32).view(M + 1, self.SP)
        self.grads = view("grads", M * self.SP, torch.float32).view(M, self.SP)
        self.z = [view(f"z{k}", P * SMALL_PITCH, torch.float32) for k in range(N_SLOTS)]
        self.tail = view("tail", P * SMALL_PITCH, torch.float32)


class PeerExchange(_ExchangeViews):
    """Exchange buffers of mode B as NVLink peer memory (torch symmetric memory): one symmetric
    allocation per rank, carved into the request / sampled-row / gradient / small-vector regions
    and mapped into every process.  The step kernels of ``csrc/`` store straight into the peers'
    regions and read their own in place, so there is no collective call on the data path -- only
    ``barrier()`` (a signal-pad barrier kernel on the current stream) between producer and consumer."""

    def __init__(self, M: int, d: int, P: int, rank: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.off, total = _region_layout(M, d, P)
        buf = symm.empty(total, dtype=torch.uint8, device=device)
        buf.zero_()
        self.hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == P and ptrs[rank] == buf.data_ptr()
        self._bind(buf, ptrs, M, d, P, rank)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def barrier(self, channel: int = 0) -> None:
        self.hdl.barrier(channel=channel)


class LocalPeerGroup:
    """P emulated ranks on ONE device (tests, single-GPU debugging): the same exchange buffers as
    ``PeerExchange``, ordinary allocations addressed through the same pointer tables.  Pass the group as
    ``exchange=`` to every emulated ``ShardedSampled``.  The ranks must be driven phase by phase in lock
    step (every rank finishes a phase before any rank starts the next); ``barrier()`` is then a no-op."""

    def __init__(self, P: int):
        self.P, self.bufs, self.off = int(P), None, None

    def member(self, rank: int, M: int, d: int, device) -> "_ExchangeViews":
        if self.bufs is None:
            self.off, total = _region_layout(M, d, self.P)
            self.bufs = [torch.zeros(total, dtype=torch.uint8, device=device) for _ in range(self.P)]
        ex = _ExchangeViews()
        ex.off = self.off
        ex._bind(self.bufs[rank], [b.data_ptr() for b in self.bufs], M, d, self.P, rank)
        ex.barrier = lambda channel=0: None
        return ex


class ShardedSampled:
    """Sampled-ELBO VFM with ROW-SHARDED tables (mode B).

    Row ``r`` (parameters, Adam moments, train count) lives on rank ``r mod P`` at local index
    ``r // P``.  One step on the global batch (each rank holds a slice):

      1. local plan of the rank's slice; its unique ids are bucketed by owner into fixed-capacity
         slots and exchanged (all-to-all #1: ids + local batch counts);
      2. the owner plans the received ids (same plan kernels, F = 1), sums the batch counts,
         draws the noise (Philox keyed by the GLOBAL id -> identical whoever asks) and returns
         the sampled rows (all-to-all #2);
      3. each rank scores its samples and runs the ordered segmented reduction locally, then
         sends the row gradients back (all-to-all #3);
      4. the owner adds the gradients of a row in source-rank order (deterministic), applies the
         KL gradient with the global counts and Adam; a small all-reduce carries the scalar sums.

    Slots are padded to a fixed capacity so that every exchange has a static shape (no host
    synchronisation); padding maps to a sentinel row ``R_loc`` kept at the KL minimum.
    """

    def __init__(self, embedding_size: int, field_sizes: Sequence[int], train_counts: torch.Tensor,
                 n_train: float, batch_local: int, world: int, rank: int, output: str = "reg",
                 link: str = "abs", kl_weighting: str = "torch", seed: int = 7, lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, device="cuda", exchange=None,
                 init: Optional[dict] = None, noise_tables=None, slack: Optional[float] = None):
        from .engine import BatchPlan, StepBuffers, require_cuda
        self.device = require_cuda(device)
        L.lib()
        self.d, self.P, self.p = int(embedding_size), int(world), int(rank)
        self.field_sizes = list(map(int, field_sizes))
        self.F, self.R = len(self.field_sizes), int(sum(self.field_sizes))
        self.B, self.n_train = int(batch_local), float(n_train)
        self.output, self.link = output, link
        self.adam = L.Adam(lr, betas[0], betas[1], eps)
        self._want_peer = isinstance(exchange, str) and exchange == "peer"
        self._local_group = exchange if isinstance(exchange, LocalPeerGroup) else None
        self.exchange = TorchExchange() if (exchange is None or self._want_peer or self._local_group) else exchange
        self.peer = None
        self.noise_tables = noise_tables
        dev, d, P, p = self.device, self.d, self.P, self.p
        if kl_weighting == "torch":
            bounds, sizes = [self.field_sizes[0] + 1], self.field_sizes[:2]
        else:
            import numpy as np
            bounds, sizes = list(np.cumsum(self.field_sizes)[:-1]), self.field_sizes
        self.R_loc = (self.R + P - 1) // P
        Rl = self.R_loc + 1                                   # + sentinel row for padding slots
        # ---- local shards
        tc = torch.as_tensor(train_counts).reshape(-1).to(torch.float32)
        self.train_counts = tc.to(dev)                        # replicated [R]: 4 B/row, feeds Z_f
        tcl = torch.full((Rl,), float("inf"), dtype=torch.float32)
        mine = tc[p::P]
        tcl[: len(mine)] = mine
        self.train_counts_loc = tcl.to(dev)
        if init is not None:
            bias = torch.zeros(Rl, 2); ent = torch.zeros(Rl, 2 * d)
            bias[: len(mine)] = init["bias"][p::P]; ent[: len(mine)] = init["entity"][p::P]
            scal = torch.zeros(L.S_COUNT)
            scal[L.S_ALPHA], scal[L.S_GB_MEAN], scal[L.S_GB_SCALE] = (float(init["alpha"]), float(init["global_bias_mean"]),
                                                                       float(init["global_bias_scale"]))
        else:
            # random init straight on the device: a shard of the 100 M-row table is tens of GB
            g = torch.Generator(device=dev).manual_seed(seed * 1000003 + p)
            bias = torch.randn(Rl, 2, generator=g, device=dev)
            ent = torch.randn(Rl, 2 * d, generator=g, device=dev)
            # CUDA randn returns exactly 0 with probability ~2^-24 per draw (Box-Muller on u = 1): a raw
            # scale of 0 is sigma = 0 under the abs link, i.e. an infinite KL -- in the reference too
            # (vfm-torch.py:126, 290-295).  Millions of draws per shard hit it; keep |raw scale| >= 1e-4.
            for t in (bias[:, 1:], ent[:, d:]):
                t.copy_(torch.where(t.abs() < 1e-4, torch.full_like(t, 1e-4), t))
            scal = torch.tensor([0.5, 0.0, 1.0, 0.0])
        bias[len(mine):] = torch.tensor([0.0, 1.0], device=bias.device)   # unused + sentinel rows: KL minimum
        ent[len(mine):, :d] = 0.0
        ent[len(mine):, d:] = 1.0
        z = torch.zeros_like
        self.bias, self.entity = bias.to(dev), ent.to(dev)
        self.bias_m, self.bias_v, self.entity_m, self.entity_v = z(self.bias), z(self.bias), z(self.entity), z(self.entity)
        self.scalars = scal.to(dev)
        self.scalars_m, self.scalars_v = z(self.scalars), z(self.scalars)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.noise_step = torch.zeros(2, dtype=torch.int32, device=dev)    # see vfmb_tables.noise_step
        # ---- requester side (local slice, global ids)
        mk = lambda B, F, R, n_tr, stride, off: self._cfg(B, F, R, n_tr, bounds, sizes, seed, stride, off)
        self.cfg_l = mk(self.B, self.F, self.R, self.n_train / P, 0, 0)
        self.cfg_g = mk(self.B * P, self.F, self.R, self.n_train, 0, 0)
        self.plan_l = BatchPlan(self.B, self.F, self.R, dev)
        self.buf_l = StepBuffers(self.cfg_l, self.plan_l, dev, L.S_COUNT, need_msg=self.F > 2)
        # ---- owner side (received ids, local row indices)
        u_cap = self.plan_l.u_cap
        # slot capacity per owner.  Default: the expected share of a batch whose rows are all distinct plus
        # six standard deviations of the multinomial split (ids are spread by id mod P), never more than an
        # owner has rows.  `slack` (fraction of u_cap / P) overrides, e.g. 0.75 for Zipf batches whose unique
        # count is known to stay below that; an overflow is flagged on the device and turns the loss into NaN
        per = u_cap / P
        cap = int(per + 6.0 * per ** 0.5 + 16) if slack is None else int(-(-int(u_cap * slack) // P))
        self.CAP = max(1, min(cap, self.R_loc + 1))
        self.M = self.CAP * P
        self.cfg_o = mk(self.M, 1, Rl, self.n_train, P, p)
        self.plan_o = BatchPlan(self.M, 1, Rl, dev)
        self.buf_o = StepBuffers(self.cfg_o, self.plan_o, dev, L.S_COUNT, need_msg=False)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.tail = torch.zeros(DP_TAIL, dtype=torch.float32, device=dev)
        # exchange buffers (slot layout [P, CAP, w]) and glue scratch, allocated once
        M = self.M
        i32 = lambda *shape: torch.zeros(shape, dtype=torch.int32, device=dev)
        f32 = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        self._send, self._dest = i32(M, 2), i32(u_cap)
        self._bws = torch.zeros(int(L.lib().vfmb_shard_bucket_workspace(u_cap)), dtype=torch.uint8, device=dev)
        self._loc = torch.zeros((M, 1), dtype=torch.int64, device=dev)
        self._reply, self._gsend, self._table = f32(M, d + 1), f32(M, d + 1), f32(M, d)
        self._tail_idx = (C.c_int32 * 4)(T_NLL, T_RESID, T_SQERR, T_KLROWS)
        self._ids_copy = i32(M, 2)
        self._zg, self._tailg = f32(L.MAX_FIELDS), f32(DP_TAIL)
        self._y = self._io_o = self._recv_ids = self._noise_o = None
        # Per-batch state exists N_SLOTS times ("slots"): everything that depends only on the ids of a
        # batch -- both plans, the slot map, the routing tables, the owner's copy of the requests, Z_f --
        # so that it can be prepared for batches i+1 and i+2 while step i runs (ShardedPipeline).
        # _use_slot(k) makes slot k the one the phase methods work on.
        # fused peer mode: where every row of the batch is read from / its gradient written to (vfmb_shard_route)
        N = self.B * self.F
        self.SP = d + 4
        self.n_real = int(len(mine))
        self._inv_slot, self._partner_slot = i32(N), i32(N)
        self._gptr = torch.zeros(u_cap, dtype=torch.int64, device=dev)
        self._dump = f32(self.SP)
        # outputs of a step (predictions, loss terms) also exist per slot: step i+1 writes other buffers, so
        # a caller may copy step i's outputs to the host on its own stream while step i+1 runs
        self._pred, self._mean = f32(self.B), f32(self.B)
        self._stats = torch.zeros(L.STATS, dtype=torch.float32, device=dev)
        self._slot_keys = ("plan_l", "plan_o", "_dest", "_ids_copy", "_loc", "_zg", "_y", "_io_o", "_recv_ids", "_noise_o",
                           "_inv_slot", "_partner_slot", "_gptr", "_pred", "_mean", "_stats")
        mk_slot = lambda: {"plan_l": BatchPlan(self.B, self.F, self.R, dev), "plan_o": BatchPlan(self.M, 1, Rl, dev),
                           "_dest": i32(u_cap), "_ids_copy": i32(M, 2),
                           "_loc": torch.zeros((M, 1), dtype=torch.int64, device=dev), "_zg": f32(L.MAX_FIELDS),
                           "_y": None, "_io_o": None, "_recv_ids": None, "_noise_o": None,
                           "_inv_slot": i32(N), "_partner_slot": i32(N),
                           "_gptr": torch.zeros(u_cap, dtype=torch.int64, device=dev),
                           "_pred": f32(self.B), "_mean": f32(self.B),
                           "_stats": torch.zeros(L.STATS, dtype=torch.float32, device=dev)}
        self._slots = [{k: getattr(self, k) for k in self._slot_keys}] + [mk_slot() for _ in range(N_SLOTS - 1)]
        self._k = 0
        if self._want_peer:                                  # collective: every rank constructs it
            self.peer = PeerExchange(M, d, P, p, dev)
        elif self._local_group is not None:
            self.peer = self._local_group.member(p, M, d, dev)

    def _cfg(self, B, F, R, n_train, bounds, sizes, seed, stride, off):
        cfg = make_config(B, F, self.d, R, 1, self.output, self.link, bounds, sizes, n_train, seed,
                          "pairwise" if self.F > 2 else "prod")      # the owner side (F = 1) learns the model's interaction
        cfg.row_stride, cfg.row_offset = int(stride), int(off)
        return cfg

    def _tables(self) -> L.Tables:
        return L.Tables(L.ptr(self.bias), L.ptr(self.bias_m), L.ptr(self.bias_v), L.ptr(self.entity),
                        L.ptr(self.entity_m), L.ptr(self.entity_v), L.ptr(self.train_counts_loc),
                        L.ptr(self.scalars), L.ptr(self.scalars_m), L.ptr(self.scalars_v), L.ptr(self.adam_step),
                        L.ptr(self.noise_step))

    def _use_slot(self, k: int) -> None:
        for key in self._slot_keys:                           # save what the phases assigned, switch
            self._slots[self._k][key] = getattr(self, key)
        self._k = k
        for key in self._slot_keys:
            setattr(self, key, self._slots[k][key])

    # ------------------------------------------------------------------ phases
    @torch.no_grad()
    def phase_request(self, x: torch.Tensor, y: torch.Tensor):
        """Local plan; unique ids bucketed by owner.  Returns (send [P,CAP,2] int32, z_local [8])."""
        P, CAP, M = self.P, self.CAP, self.M
        self._y = y.to(self.device, torch.float32).contiguous()
        x = x.to(self.device).contiguous()
        self.plan_l.build(self.cfg_l, x, self.train_counts)
        pl = self.plan_l
        pe = self.peer
        L.check(L.lib().vfmb_shard_bucket(C.byref(pl.struct), pl.u_cap, P, CAP, self._send.data_ptr(),
                                          self._dest.data_ptr(), self.overflow.data_ptr(), self._bws.data_ptr(),
                                          pe.tables[f"ids{self._k}"] if pe else None, self.p,
                                          current_stream(self.device)), "vfmb_shard_bucket")
        if pe:
            s = current_stream(self.device)
            L.check(L.lib().vfmb_shard_put_small(pl.z.data_ptr(), L.MAX_FIELDS, SMALL_PITCH, pe.tables[f"z{self._k}"], P, self.p,
                                                 s), "vfmb_shard_put_small")
            # id-only routing tables of the fused step: slots to read rows from, peer addresses of the gradients
            L.check(L.lib().vfmb_shard_route(C.byref(pl.struct), self._dest.data_ptr(), self.B, self.F, pl.u_cap, M, CAP,
                                             self.SP, pe.tables["grads"], P, self.p, self._dump.data_ptr(),
                                             self._inv_slot.data_ptr(), self._partner_slot.data_ptr(), self._gptr.data_ptr(), s),
                    "vfmb_shard_route")
            return None, None
        return self._send.view(P, CAP, 2), pl.z.clone()

    def phase_owner_stage(self, recv: torch.Tensor, z_global: torch.Tensor) -> torch.Tensor:
        """Owner: plan the received ids, global batch counts, sample the rows.  Returns the reply
        [P,CAP,d+1] (sampled factor row | sampled bias) in the slot layout of the request."""
        self._owner_prepare(recv, z_global)
        return self._owner_sample()

    @torch.no_grad()
    def _owner_prepare(self, recv, z_global) -> None:
        """The part of the owner's work that depends only on the requested ids."""
        M, d, P, p = self.M, self.d, self.P, self.p
        pe = self.peer
        po, bo, lib, s = self.plan_o, self.buf_o, L.lib(), current_stream(self.device)
        if pe:      # the requests are in this rank's peer region; keep a private copy (the region is
            # rewritten by the next step), Z_f = sum of the ranks' slots in rank order
            L.check(lib.vfmb_shard_owner_ids(pe.ids[self._k].data_ptr(), M, P, self.R_loc, self._loc.data_ptr(),
                                             self._ids_copy.data_ptr(), s), "vfmb_shard_owner_ids")
            recv = self._ids_copy
            L.check(lib.vfmb_shard_sum_small(pe.z[self._k].data_ptr(), P, L.MAX_FIELDS, SMALL_PITCH, self._zg.data_ptr(), s),
                    "vfmb_shard_sum_small")
            z_global = self._zg
        else:
            recv = recv.contiguous()
            L.check(lib.vfmb_shard_owner_ids(recv.data_ptr(), M, P, self.R_loc, self._loc.data_ptr(), None, s),
                    "vfmb_shard_owner_ids")
        self._recv_ids = recv
        po.build(self.cfg_o, self._loc, self.train_counts_loc)
        # batch count of every owned row summed over the requesting ranks -> urec[:, 3]
        L.check(lib.vfmb_shard_owner_pack(C.byref(po.struct), recv.data_ptr(), M, self.CAP, d, None, None, None, 1,
                                          None, self.p, s), "vfmb_shard_owner_pack")
        po.z.copy_(z_global)

    @torch.no_grad()
    def _owner_sample(self):
        """Owner (collective path): draw the noise of the requested rows, reply with the sampled rows."""
        M, d, P, p = self.M, self.d, self.P, self.p
        recv = self._recv_ids
        po, bo, lib, s = self.plan_o, self.buf_o, L.lib(), current_stream(self.device)
        noise = None
        if self.noise_tables is not None:                     # tests: per-entity noise tables
            e0, eb_t, ee_t = self.noise_tables
            gid = (po.uniq.long() * P + p).clamp_(min=0, max=self.R - 1)   # ranks >= U hold garbage ids
            noise = (e0.reshape(1).contiguous(), eb_t[gid].contiguous(), ee_t[gid].contiguous())
        self._noise_o = noise
        self._io_o = bo.io(noise=noise)
        L.check(L.lib().vfmb_sampled_stage(C.byref(self.cfg_o), C.byref(self._tables()), C.byref(po.struct),
                                           C.byref(self._io_o), current_stream(self.device)), "vfmb_sampled_stage")
        L.check(lib.vfmb_shard_owner_pack(C.byref(po.struct), recv.data_ptr(), M, self.CAP, d, bo.vs.data_ptr(),
                                          bo.ws.data_ptr(), self._reply.data_ptr(), 0, None, self.p, s), "vfmb_shard_owner_pack")
        return self._reply.view(P, self.CAP, d + 1)

    @torch.no_grad()
    def phase_local(self, recv_rows: torch.Tensor):
        """Requester (collective path): place the sampled rows, score the samples, ordered segmented
        reduction.  Returns (row gradients [P,CAP,d+1] in slot layout, additive scalars tail [16])."""
        M, d, P = self.M, self.d, self.P
        pl, bl = self.plan_l, self.buf_l
        recv_rows = recv_rows.contiguous()
        lib, s = L.lib(), current_stream(self.device)
        L.check(lib.vfmb_shard_unpack_rows(C.byref(pl.struct), recv_rows.data_ptr(), self._dest.data_ptr(), pl.u_cap, M, d,
                                           bl.vs.data_ptr(), bl.ws.data_ptr(), s), "vfmb_shard_unpack_rows")
        e0 = self.noise_tables[0].reshape(1).contiguous() if self.noise_tables is not None else None
        io = bl.io(y=self._y)
        io.eps_global = L.ptr(e0)
        self._io_l = io
        tab = self._tables()
        L.check(lib.vfmb_sampled_score(C.byref(self.cfg_l), C.byref(tab), C.byref(pl.struct), C.byref(io), s),
                "vfmb_sampled_score")
        L.check(lib.vfmb_sampled_gather(C.byref(self.cfg_l), C.byref(pl.struct), C.byref(io), None, 0, s),
                "vfmb_sampled_gather")
        L.check(lib.vfmb_shard_pack_grads(C.byref(pl.struct), bl.grow.data_ptr(), bl.gws.data_ptr(), self._dest.data_ptr(),
                                          pl.u_cap, M, self.CAP, d, self._gsend.data_ptr(), bl.stats.data_ptr(),
                                          self.buf_o.stats.data_ptr(), float(self.B), self.tail.data_ptr(),
                                          self._tail_idx, DP_TAIL, None, self.p, s), "vfmb_shard_pack_grads")
        self.tail[L.DP_T_OVERFLOW] = self.overflow[0].to(torch.float32)      # any rank's overflow -> NaN loss everywhere
        return self._gsend.view(P, self.CAP, d + 1), self.tail

    @torch.no_grad()
    def phase_owner_update(self, recv_g: torch.Tensor, tail_global: torch.Tensor) -> dict:
        """Owner (collective path): add a row's gradients in source-rank order, KL gradient + Adam on the
        owned rows, replicated scalar update."""
        M, d = self.M, self.d
        po, bo = self.plan_o, self.buf_o
        io, tab, s, lib = self._io_o, self._tables(), current_stream(self.device), L.lib()
        recv_g = recv_g.contiguous()
        L.check(lib.vfmb_shard_unpack_grads(C.byref(po.struct), recv_g.data_ptr(), None, M, d, self._table.data_ptr(),
                                            bo.rsorted.data_ptr(), s), "vfmb_shard_unpack_grads")
        L.check(lib.vfmb_sampled_gather(C.byref(self.cfg_o), C.byref(po.struct), C.byref(io), self._table.data_ptr(), 1, s),
                "vfmb_sampled_gather")
        L.check(lib.vfmb_sampled_adam_rows(C.byref(self.cfg_o), C.byref(tab), C.byref(po.struct), C.byref(io),
                                           C.byref(self.adam), L.ADAM_TOUCHED, 1.0, s), "vfmb_sampled_adam_rows")
        st = self.buf_l.stats
        st[L.ST_KL_ROWS] = tail_global[T_KLROWS]
        e0 = self.noise_tables[0].reshape(1).contiguous() if self.noise_tables is not None else None
        L.check(lib.vfmb_dp_final(C.byref(self.cfg_g), C.byref(tab), tail_global.data_ptr(), L.ptr(e0),
                                  C.byref(self.adam), st.data_ptr(), s), "vfmb_dp_final")
        return {"loss": st[L.ST_LOSS], "kl": st[L.ST_KL], "nll_mean": st[L.ST_NLL_MEAN],
                "pred": self.buf_l.mean[: self.B], "stats": st}

    def step(self, x_local: torch.Tensor, y_local: torch.Tensor) -> dict:
        ex = self.exchange
        mark = self._mark
        if self.peer is not None:       # peer memory: the pack kernels are the all-to-alls
            self._use_slot(0)
            mark("start")
            self._phase_a(x_local, y_local, mark)
            return self._phase_b(mark)
        mark("start")
        send, z = self.phase_request(x_local, y_local)
        mark("request")
        recv = ex.all_to_all(send)
        z = ex.all_reduce(z)
        mark("a2a_ids")
        reply = self.phase_owner_stage(recv, z)
        mark("owner_stage")
        rows = ex.all_to_all(reply)
        mark("a2a_rows")
        grads, tail = self.phase_local(rows)
        mark("local")
        recv_g = ex.all_to_all(grads)
        tail = ex.all_reduce(tail)
        mark("a2a_grads")
        out = self.phase_owner_update(recv_g, tail)
        mark("owner_update")
        return out

    def _phase_a(self, x_local, y_local, mark=lambda name: None) -> None:
        """Peer mode, everything that depends only on the ids of the batch (current slot): local
        plan, request exchange, the owner's plan and global counts."""
        self._a1(x_local, y_local)
        mark("request")
        self._a2(mark)

    def _a1(self, x_local, y_local) -> None:
        """Part A1 (requester, no barrier): local plan, requests and Z_f stored into the owners' regions
        of the current slot, routing tables of the fused step."""
        self.phase_request(x_local, y_local)

    def _a2(self, mark=lambda name: None) -> None:
        """Part A2 (owner): wait for every rank's requests of the current slot, plan them, global counts."""
        self.peer.barrier(0)
        mark("a2a_ids")
        self._owner_prepare(None, None)

    def _phase_b(self, mark=lambda name: None) -> dict:
        """Peer mode, the parameter-dependent part of the step (current slot), fused: five kernels and two
        barriers -- the sampled rows are stored by the stage kernel straight into the requesters' slots,
        scored and gathered in place, the gradient rows stored by the gather kernel into the owners' slots
        and summed / applied there together with the scalar parameters."""
        self._b_stage()
        mark("owner_stage")
        self.peer.barrier(1)
        mark("a2a_rows")
        self._b_local()
        mark("local")
        self.peer.barrier(2)
        mark("a2a_grads")
        out = self._b_update()
        mark("owner_update")
        return out

    @torch.no_grad()
    def _b_stage(self) -> None:
        """Owner: sample the requested rows into the requesters' slots (k_stage with peer stores)."""
        pe, po, bo, P, p = self.peer, self.plan_o, self.buf_o, self.P, self.p
        noise = None
        if self.noise_tables is not None:                     # tests: per-entity noise tables
            e0_t, eb_t, ee_t = self.noise_tables
            gid = (po.uniq.long() * P + p).clamp_(min=0, max=self.R - 1)   # ranks >= U hold garbage ids
            noise = (e0_t.reshape(1).contiguous(), eb_t[gid].contiguous(), ee_t[gid].contiguous())
        self._noise_o = noise
        self._io_o = bo.io(noise=noise)
        L.check(L.lib().vfmb_shard_stage_put(C.byref(self.cfg_o), C.byref(self._tables()), C.byref(po.struct),
                                             C.byref(self._io_o), self.CAP, self.SP, self.n_real, pe.tables["rows"], P, p,
                                             current_stream(self.device)), "vfmb_shard_stage_put")

    @torch.no_grad()
    def _b_local(self) -> None:
        """Requester: score the samples and reduce the gradients on the received slots, in place; the
        finished gradient rows and the rank's additive scalars go to their owners / to every rank."""
        pe, pl, bl, bo, P, p, SP = self.peer, self.plan_l, self.buf_l, self.buf_o, self.P, self.p, self.SP
        lib, s, tab = L.lib(), current_stream(self.device), self._tables()
        e0 = self._noise_o[0] if self._noise_o is not None else None
        io = bl.io(y=self._y)
        io.eps_global = L.ptr(e0)
        io.pred, io.mean, io.stats = L.ptr(self._pred), L.ptr(self._mean), L.ptr(self._stats)
        self._io_l = io
        L.check(lib.vfmb_shard_score(C.byref(self.cfg_l), C.byref(tab), C.byref(pl.struct), C.byref(io), pe.rows.data_ptr(), SP,
                                     self._inv_slot.data_ptr(), pe.tables["tail"], P, p, bo.stats.data_ptr(),
                                     self.overflow.data_ptr(), float(self.B), SMALL_PITCH, s), "vfmb_shard_score")
        L.check(lib.vfmb_shard_gather_put(C.byref(self.cfg_l), C.byref(pl.struct), C.byref(io), pe.rows.data_ptr(), SP,
                                          self._partner_slot.data_ptr(), self._gptr.data_ptr(), self._dest.data_ptr(), s),
                "vfmb_shard_gather_put")

    @torch.no_grad()
    def _b_update(self) -> dict:
        """Owner: ordered sum of the received gradients, Adam on the owned rows, scalar parameters, loss."""
        pe, po = self.peer, self.plan_o
        e0 = self._noise_o[0] if self._noise_o is not None else None
        st = self._stats
        L.check(L.lib().vfmb_shard_owner_update(C.byref(self.cfg_o), C.byref(self._tables()), C.byref(po.struct),
                                                C.byref(self._io_o), C.byref(self.adam), pe.grads.data_ptr(), self.SP,
                                                self.n_real, pe.tail.data_ptr(), self.P, SMALL_PITCH, self.B * self.P, float(self.n_train),
                                                st.data_ptr(), L.ptr(e0), current_stream(self.device)),
                "vfmb_shard_owner_update")
        return {"loss": st[L.ST_LOSS], "kl": st[L.ST_KL], "nll_mean": st[L.ST_NLL_MEAN],
                "pred": self._mean[: self.B], "stats": st}

    def graphed_step(self):
        """Capture one whole step -- both plans, every kernel and the NCCL collectives -- in a CUDA
        graph; returns ``run(x_local, y_local) -> dict`` that copies the batch into the graph's
        static input buffers and replays it.  The eager step is host-bound (~45 launches and 5
        collectives enqueued from Python per step take longer than the GPU needs to run them).
        Call after at least one eager ``step`` (communicators and lazy allocations exist)."""
        assert getattr(self, "_timing", None) is None, "disable phase timing before capturing"
        xs = torch.zeros((self.B, self.F), dtype=torch.int64, device=self.device)
        ys = torch.zeros(self.B, dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.step(xs, ys)

        def run(x_local: torch.Tensor, y_local: torch.Tensor) -> dict:
            xs.copy_(x_local, non_blocking=True)
            ys.copy_(y_local, non_blocking=True)
            g.replay()
            return out
        run.graph = g
        return run

    # ---- optional per-phase timing (CUDA events on the step's stream; read with phase_times())
    def enable_timing(self, on: bool = True) -> None:
        self._timing = [] if on else None

    def _mark(self, name: str) -> None:
        if getattr(self, "_timing", None) is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._timing.append((name, ev))

    def phase_times(self) -> dict:
        """Mean milliseconds per phase over the steps recorded since enable_timing() (synchronises)."""
        torch.cuda.synchronize(self.device)
        acc, cnt = {}, {}
        tl = self._timing or []
        for (n0, e0), (n1, e1) in zip(tl[:-1], tl[1:]):
            if n1 == "start":
                continue
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        return {k: acc[k] / cnt[k] for k in acc}

    def check_overflow(self) -> None:
        """Raises if a bucket ever exceeded the slot capacity (synchronises; call occasionally)."""
        if int(self.overflow.item()) != 0:
            raise RuntimeError("ShardedSampled: owner bucket overflow -- increase `slack`")

    # ------------------------------------------------------------------ helpers (tests / checkpoints)
    def gather_tables(self):
        """This rank's rows as (global ids, bias rows, entity rows) -- for assembling full tables."""
        n = len(range(self.p, self.R, self.P))
        gid = torch.arange(self.p, self.R, self.P, device=self.device)
        return gid, self.bias[:n], self.entity[:n]


class ShardedPipeline:
    """Software-pipelined, CUDA-graphed mode-B loop over NVLink peer memory (fixed batch size).

    A step has two id-only parts -- A1 (requester: local plan, requests into the owners' regions, routing
    tables) and A2 (owner: plan of the received requests, global batch counts) -- and the parameter-
    dependent part B (sample rows into the requesters' slots -> score / segmented sums in place, gradient
    rows into the owners' slots -> owner update).  Three batches are in flight: graph s holds B of the
    batch in slot s on the main branch, A2 of the next batch (slot s+1) and A1 of the one after (slot s+2)
    on two side branches, so both plans and the id exchange are off the critical path and each of the
    three chains is about as long as the others.  A2 uses barrier channel 0, B channels 1 and 2; the
    request / Z_f regions exist once per slot.

        pipe = ShardedPipeline(model)
        pipe.start(x0, y0, x1, y1)
        for x_next, y_next in batches[2:]:
            out = pipe.step(x_next, y_next)      # runs the oldest staged batch, stages (x_next, y_next)
        out = pipe.step(); out = pipe.step()      # the last two staged batches
    """

    def __init__(self, model: ShardedSampled, reserve: int = 1, side_priority: int = -1):
        assert model.peer is not None, "ShardedPipeline needs exchange='peer'"
        assert getattr(model, "_timing", None) is None, "disable phase timing before capturing"
        m, dev = model, model.device
        self.m, D = m, N_SLOTS
        self.xs = [torch.zeros((m.B, m.F), dtype=torch.int64, device=dev) for _ in range(D)]
        self.ys = [torch.zeros(m.B, dtype=torch.float32, device=dev) for _ in range(D)]
        self.side1 = torch.cuda.Stream(device=dev, priority=side_priority)
        self.side2 = torch.cuda.Stream(device=dev, priority=side_priority)
        self.head = 0
        self.graphs, self.outs = [], []
        torch.cuda.synchronize(dev)
        L.check(L.lib().vfmb_set_grid_reserve(reserve), "vfmb_set_grid_reserve")
        n0 = int(L.lib().vfmb_launch_count())
        try:
            for s in range(D):
                s1, s2 = (s + 1) % D, (s + 2) % D
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    cur = torch.cuda.current_stream(dev)
                    self.side1.wait_stream(cur)                  # fork
                    self.side2.wait_stream(cur)
                    with torch.cuda.stream(self.side2):          # A1 of batch i+2
                        m._use_slot(s2)
                        m._a1(self.xs[s2], self.ys[s2])
                    with torch.cuda.stream(self.side1):          # A2 of batch i+1
                        m._use_slot(s1)
                        m._a2()
                    m._use_slot(s)                               # B of batch i
                    m._y = self.ys[s]                            # targets of the batch staged in slot s
                    out = m._phase_b()
                    cur.wait_stream(self.side1)                  # join
                    cur.wait_stream(self.side2)
                self.graphs.append(g)
                self.outs.append(out)
        finally:
            L.lib().vfmb_set_grid_reserve(0)
        self.launches_per_graph = (int(L.lib().vfmb_launch_count()) - n0) // D

    def start(self, x0: torch.Tensor, y0: torch.Tensor, x1: torch.Tensor, y1: torch.Tensor) -> None:
        """Stage the first two batches and run their id-only parts (outside the graphs): A1 and A2 of
        batch 0, A1 of batch 1."""
        m = self.m
        self.head = 0
        for k, (x, y) in enumerate(((x0, y0), (x1, y1))):
            self.xs[k].copy_(x, non_blocking=True)
            self.ys[k].copy_(y, non_blocking=True)
            m._use_slot(k)
            m._a1(self.xs[k], self.ys[k])
        m._use_slot(0)
        m._a2()

    def step(self, x_next: Optional[torch.Tensor] = None, y_next: Optional[torch.Tensor] = None) -> dict:
        """Run the step on the oldest staged batch; ``(x_next, y_next)`` is staged two batches ahead
        (without it the slot's old contents are planned again, harmlessly, at the end of a run)."""
        s = self.head
        if x_next is not None:
            t = (s + 2) % N_SLOTS
            self.xs[t].copy_(x_next, non_blocking=True)
            self.ys[t].copy_(y_next, non_blocking=True)
        self.graphs[s].replay()
        self.head = (s + 1) % N_SLOTS
        return self.outs[s]
