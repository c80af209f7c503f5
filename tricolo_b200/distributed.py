"""Multi-GPU forms of the hot path: one process per GPU, torch.distributed (NCCL over NVLink).

Training — global-negative trimodal InfoNCE (BASELINE.json configs[3]).  The reference never
gathers negatives; the semantics here are "the reference loss evaluated on the
concatenated global batch".  Each rank holds B/W rows of every modality and owns that row block
of each logit matrix:
    gather       16-bit normalised embeddings, all modalities at once (interleaved rows, strided TMA operands):
                 K1 stores its rows straight into every peer's buffer over NVLink (symmetric memory), or one
                 NCCL all-gather                                 (W-1)/W * 3*B*D*2 bytes in
    reduce       sum-exp statistics [3, P, B] fp32 (column partials + every rank's row sums and positives; fixed
                 shift -> plain sums): one-shot peer read + add in rank order, or one NCCL all-reduce
Backward (peer-memory transport, 256 < dim <= 512, rows per rank a multiple of 128): the shared-G form sharded as
north_star describes it - the rank forms its row block of every pair's softmax-gradient matrix G once, the row-side
gradients G Zcol are complete locally, and the column-side products G^T Zrow_local are reduce-scattered by the GEMM
kernel itself: its accumulator drain TMA-stores each piece into the owner rank's receive buffer over NVLink
(tcl_ntxent_bwd_sharded_gemm); one barrier, then the owner adds the W partials in a fixed order inside the normalise
backward (tcl_ntxent_bwd_sharded_finish).  6 b B D executed per pair.  Otherwise (NCCL transport, other shapes,
TRICOLO_B200_SHARDED_BWD=pc): the directional kernel of one GPU run for the local rows of each modality against ALL
rows of the partner - complete local gradients without any exchange, at 8 b B D.

Retrieval — gallery sharded over ranks (configs[4]): every rank scores all queries against its
shard; ground-truth similarities are all-reduced (owner contributes, others add 0), local top-k
lists are all-gathered and merged, rank counts are all-reduced.
"""
from __future__ import annotations

from itertools import combinations
from typing import List, Sequence

import torch
import torch.distributed as dist

import os
import weakref

from . import ops
from .loss.nt_xent import DEFAULT_OP_FORMAT

# ---------------------------------------------------------------------------------------------------------------
# NVLink peer-memory transport of the sharded loss (default when torch symmetric memory is available;
# TRICOLO_B200_SYMM=0 selects NCCL).  The gathered operand buffer and the statistics buffer of every rank are
# symmetric-memory allocations mapped into every peer:
#   * K1 writes each normalised row straight into all ranks' gathered buffers (tcl_l2norm_fwd_bcast: the all-gather
#     is the kernel's own store traffic over NVLink), bracketed by two device-side barriers;
#   * the sum-exp statistics are exchanged through per-rank slots: every rank writes its slot, one barrier, then the
#     finalise kernel pulls all slots over NVLink and adds them in rank order (tcl_ntxent_finalize_sharded) -
#     bit-identical sums everywhere, no NCCL launch.
# Buffers are cached per (group, shapes); a forward that needs gradients holds its buffer until its backward ran.
# ---------------------------------------------------------------------------------------------------------------
_SYMM_CACHE = {}


class _SymmWorkspace:
    def __init__(self, group, world, rank, b_loc, n, dim, dt, p, dev):
        import torch.distributed._symmetric_memory as symm_mem

        b_glob = b_loc * world
        self.z = symm_mem.empty((b_glob, n * dim), dtype=dt, device=dev)
        self.hz = symm_mem.rendezvous(self.z, group if group is not None else dist.group.WORLD)
        self.z_peers = [self.hz.get_buffer(r, self.z.shape, dt) for r in range(world)]
        # statistics slots [world][col P x B | row P x b_loc | diag P x b_loc]: a rank writes slot `rank` of its own
        # buffer and every rank pulls slot s from rank s (barrier form), or pushes it into every buffer (flag form)
        g = group if group is not None else dist.group.WORLD
        nst = ops.shard_stats_bytes(p, b_loc, world) // 4
        self.stats_slots = symm_mem.empty((nst,), dtype=torch.float32, device=dev)
        self.hstats = symm_mem.rendezvous(self.stats_slots, g)
        self.stats_addrs = [int(self.hstats.get_buffer(r, self.stats_slots.shape, torch.float32).data_ptr())
                            for r in range(world)]
        lo = rank * b_loc
        esz = self.z.element_size()
        self.z_row_stride = n * dim
        # destinations of K1's stores: address of this rank's first row of modality m inside each target: one store
        # per peer-mapped buffer.  TRICOLO_B200_MULTICAST=1 stores ONCE to the NVLink multicast (NVLS) mapping instead
        # and lets the switch replicate; correct, but measured slower at N=2 (47 vs 29 us for K1), so opt-in.
        mc = int(self.hz.multicast_ptr) if os.environ.get("TRICOLO_B200_MULTICAST", "0") == "1" else 0
        bases = [mc] if mc else [int(zp.data_ptr()) for zp in self.z_peers]
        self.multicast = bool(mc)
        self.dsts = [[base + (lo * n * dim + m * dim) * esz for m in range(n)] for base in bases]
        if not mc:  # own buffer first, then the peers in a rank-staggered order (no all-to-one bursts)
            self.dsts = self.dsts[rank:] + self.dsts[:rank]
        self.side = torch.cuda.Stream(device=dev)  # copy-engine transfers that overlap the forward tile kernel
        self.busy = False  # a forward with autograd holds the gathered operands until its backward
        self.group, self.world, self.rank, self.b_loc, self.n, self.dim = group, world, rank, b_loc, n, dim
        self._bwd = {}
        # flag-based protocol (no barrier kernels): sync pad + statistics slots, peer-mapped
        self.flags_ok = b_loc % 128 == 0 and b_loc <= 8192 and dim <= 512 and dim % 64 == 0
        if self.flags_ok:
            self.sync = symm_mem.empty((ops.shard_sync_bytes() // 4,), dtype=torch.int32, device=dev)
            self.hsync = symm_mem.rendezvous(self.sync, g)
            self.sync.zero_()
            self.hsync.barrier(channel=0)  # every pad is zero before any rank signals into it
            self.sync_addrs = [int(self.hsync.get_buffer(r, self.sync.shape, torch.int32).data_ptr()) for r in range(world)]
            self.dsts_all = [[int(zp.data_ptr()) + (lo * n * dim + m * dim) * esz for m in range(n)] for zp in self.z_peers]
            self.z_base_addrs = [int(zp.data_ptr()) for zp in self.z_peers]

    def sharded_bwd(self, pairs, need_grad):
        """(plan, local workspace, receive-buffer addresses per rank, symmetric handle) of the sharded shared-G backward.
        First use for a (pairs, need_grad) combination allocates and rendezvouses the receive buffer: collective -
        every rank reaches it in the same backward because the flags must agree across ranks."""
        import torch.distributed._symmetric_memory as symm_mem

        key = (tuple(pairs), tuple(bool(g) for g in need_grad))
        if key not in self._bwd:
            plan = ops.ShardedBwdPlan(self.n, pairs, need_grad, self.b_loc, self.world, self.dim)
            dev = self.z.device
            recv = symm_mem.empty((plan.recv_bytes,), dtype=torch.uint8, device=dev)
            hr = symm_mem.rendezvous(recv, self.group if self.group is not None else dist.group.WORLD)
            addrs = [int(hr.get_buffer(r, (plan.recv_bytes,), torch.uint8).data_ptr()) for r in range(self.world)]
            work = torch.empty((plan.workspace_bytes,), dtype=torch.uint8, device=dev)
            self._bwd[key] = (plan, work, addrs, hr, recv)
        return self._bwd[key][:4]


def _symm_enabled() -> bool:
    return os.environ.get("TRICOLO_B200_SYMM", "1") != "0"


def _symm_workspace(group, world, rank, b_loc, n, dim, dt, p, dev):
    """Cached symmetric workspace, or None (transport unavailable / all cached buffers still held by a pending
    backward: the caller uses NCCL for this step).  Collective: every rank takes the same branch because the key and
    the busy flags evolve identically on all ranks."""
    if not _symm_enabled() or world > 8 or world < 2:
        return None
    key = (id(group) if group is not None else 0, world, b_loc, n, dim, dt, p, dev.index)
    if key not in _SYMM_CACHE:
        try:
            _SYMM_CACHE[key] = _SymmWorkspace(group, world, rank, b_loc, n, dim, dt, p, dev)
        except Exception as e:  # symmetric memory not supported on this system: say so once, use NCCL
            _SYMM_CACHE[key] = None
            if rank == 0:
                print(f"tricolo_b200.distributed: symmetric memory unavailable ({e!r:.120}); using NCCL collectives")
    ws = _SYMM_CACHE[key]
    if ws is None or ws.busy:
        return None
    return ws


def _hold(ws, ctx) -> None:
    """A forward that needs gradients holds the workspace (its gathered operands) until its backward has run - or until
    the autograd node is dropped without one (loss discarded, exception): the finalizer releases it then, so that a
    lost backward does not leave every later step on the NCCL path."""
    ws.busy = True
    ctx.symm_ws = ws
    token = object()
    ws.holder = token
    try:
        weakref.finalize(ctx, _release, weakref.ref(ws), token)
    except TypeError:  # the autograd node cannot be weakly referenced on this torch build: released by backward only
        pass


def _release(ws_ref, token) -> None:
    ws = ws_ref()
    if ws is not None and getattr(ws, "holder", None) is token:
        ws.busy = False


def _flags_enabled() -> bool:
    """TRICOLO_B200_SHARD_SYNC=flags selects the barrier-free protocol (device-side flags, arrival-gated tile sweep,
    optional fused all-gather).  Same results; measured SLOWER than the barrier form on 8 B200 (0.29 vs 0.22 ms per
    step at B=8192: every flag costs a system-scope fence behind a burst of NVLink stores), so it is opt-in."""
    return os.environ.get("TRICOLO_B200_SHARD_SYNC", "barrier") == "flags"


def _fused_push_enabled() -> bool:
    """TRICOLO_B200_PUSH=k1: the normalise kernel stores the rows to the peers itself (gather before the tile kernel);
    default: the tile kernel's push warps do it while its MMAs run."""
    return os.environ.get("TRICOLO_B200_PUSH", "fwd") != "k1"


def _sharded_g_enabled(b_loc: int = 1 << 30) -> bool:
    """Backward form of a sharded step: TRICOLO_B200_SHARDED_BWD=sharedg|pc forces one.  Default: the sharded shared-G
    form (6 b B D, in-kernel reduce-scatter, two modalities gathered) from 2048 rows per rank on, the directional
    kernel (8 b B D, no exchange, three modalities gathered) below - measured at B=8192 (profiles/r2m_*, r2n_*):
    N=2 0.407 vs 0.470 ms per step, N=4 0.279 vs 0.287, N=8 0.214 vs 0.207 (with 1024 rows per rank kernel A, kernel B
    and the summing normalise backward each carry ~10 us of fixed cost that the saved recompute no longer covers)."""
    e = os.environ.get("TRICOLO_B200_SHARDED_BWD", "")
    if e in ("sharedg", "pc"):
        return e == "sharedg"
    return b_loc >= 2048


def _gather_plan(dsts, pairs, n, use_sg: bool, needs_grad: bool, multicast: bool):
    """(destination addresses for K1's stores, modalities sent later by the copy engines).

    `dsts[r][m]`: address of this rank's rows of modality m in destination r, the own buffer first.  Only the COLUMN side of
    a pair is read from other ranks by the forward, the G recompute and the row-side gradient GEMM; the directional
    backward alone also reads the row side (text) of every rank.  So under the sharded shared-G backward, or without
    gradients, the row-only modalities get address 0 (= skipped by tcl_l2norm_fwd_bcast) at every remote destination.
    For the directional backward K1 sends everything, unless TRICOLO_B200_DEFER_GATHER=1 hands those rows to the copy
    engines.  TRICOLO_B200_GATHER_ALL=1 or the multicast mapping: K1 sends everything."""
    if multicast or os.environ.get("TRICOLO_B200_GATHER_ALL", "0") == "1":
        return dsts, []
    cols = {b for _, b in pairs}
    local_only = [[a if m in cols else 0 for m, a in enumerate(d)] for d in dsts[1:]]
    if use_sg or not needs_grad:
        return [dsts[0]] + local_only, []
    if os.environ.get("TRICOLO_B200_DEFER_GATHER", "0") == "1":
        return [dsts[0]] + local_only, [m for m in range(n) if m not in cols]
    return dsts, []


def _world(group=None):
    return dist.get_world_size(group), dist.get_rank(group)


class _GlobalNTXent(torch.autograd.Function):
    """Two collectives in the forward (one all-gather of all modalities, one all-reduce of the sum-exp
    statistics); none in the backward."""

    @staticmethod
    def forward(ctx, temperature, alpha, op_format, pairs, group, grad_world_scale, *feats):
        world, rank = _world(group)
        inv_tau = 1.0 / float(temperature)
        feats = [f.detach() for f in feats]
        n, p = len(feats), len(pairs)
        b_loc, dim = feats[0].shape
        b_glob = b_loc * world
        row_offset = rank * b_loc
        dev = feats[0].device
        dt = ops.L.op_torch_dtype(op_format)
        # all modalities interleaved per row: [b_loc, n*dim]; modality m = columns [m*dim, (m+1)*dim) with row
        # stride n*dim, so ONE gathered buffer yields every [B, dim] operand without a copy
        needs_grad = any(ctx.needs_input_grad[6:])
        ws = _symm_workspace(group, world, rank, b_loc, n, dim, dt, p, dev)
        ctx.flags = False
        if ws is not None and ws.flags_ok and _flags_enabled():
            # no barrier kernels: K1 stores every row into all ranks' buffers as soon as the destination is ready and
            # flags each 128-row chunk; the tile kernel consumes column tiles in arrival order; statistics are pushed
            # into per-source slots and the finalise kernel waits for the W flags
            use_sg = _sharded_g_enabled(b_loc) and ops.ShardedBwdPlan.supported(b_loc, dim, world)
            fused = _fused_push_enabled()
            inv_all, xs = ops.l2norm_fwd_push(feats, ws.dsts_all, ws.z_row_stride, rank, world, ws.sync_addrs, op_format,
                                              remote=not fused)
            invs = [inv_all[m] for m in range(n)]
            z_glob3 = ws.z.view(b_glob, n, dim)
            z_all = [z_glob3[:, m] for m in range(n)]
            z_own = [z[row_offset:row_offset + b_loc] for z in z_all]
            # modalities whose remote rows anyone reads: the column side of a pair (forward, G recompute, row-side
            # gradient GEMM); the directional backward also reads the row side of every pair from all ranks
            cols = sorted({b for _, b in pairs})
            push_mods = cols if (use_sg or not needs_grad) else list(range(n))
            ops.ntxent_fwd_sharded([z_own[a] for a, _ in pairs], [z_all[b] for _, b in pairs], rank, world, inv_tau,
                                   ws.stats_addrs, ws.sync_addrs, op_format, z_base_addrs=ws.z_base_addrs,
                                   push_offsets=[m * dim for m in push_mods] if fused else ())
            ctx.use_sg = use_sg
            lse2_row_all, lse2_col, loss = ops.ntxent_finalize_sharded(p, b_loc, rank, world, inv_tau, alpha, ws.stats_addrs,
                                                                       ws.sync_addrs[rank], dev)
            if needs_grad:
                _hold(ws, ctx)
                ctx.flags = True
            ctx.cfg = (inv_tau, float(alpha), op_format, tuple(pairs), row_offset, b_glob, float(grad_world_scale), n)
            ctx.save_for_backward(lse2_row_all, lse2_col, ws.z, *xs, *invs)
            return loss
        if ws is not None:
            # peer-memory transport: K1 stores its rows into the ranks' buffers; barriers before (nobody still reads
            # the previous step's operands) and after (every rank's rows have landed everywhere).  Which modalities
            # cross NVLink follows the backward form (_gather_plan): a third less traffic under the sharded shared-G
            # backward.  The copy-engine variant (TRICOLO_B200_DEFER_GATHER=1) is correct (multi-rank parity) and cut
            # K1 + gather 42 -> 32 us on 8 B200, but the step got SLOWER (0.207 -> 0.212 ms; pipelined e2e 0.37 -> 0.46 ms):
            # 1 KB rows at a 3 KB pitch are a poor copy-engine pattern and the statistics barrier ends up waiting for them.
            use_sg = _sharded_g_enabled(b_loc) and ops.ShardedBwdPlan.supported(b_loc, dim, world)
            ctx.use_sg = use_sg
            dsts, deferred = _gather_plan(ws.dsts, pairs, n, use_sg, needs_grad, ws.multicast)
            ws.hz.barrier(channel=0)
            invs, xs = ops.l2norm_fwd_bcast(feats, dsts, ws.z_row_stride, op_format)
            ev_sent = None
            if deferred:
                cur = torch.cuda.current_stream(dev)
                ev_k1 = torch.cuda.Event()
                ev_k1.record(cur)
                esz = ws.z.element_size()
                pitch = ws.z_row_stride * esz
                with torch.cuda.stream(ws.side):
                    ws.side.wait_event(ev_k1)
                    for d in ws.dsts[1:]:  # peers in the rank-staggered order; ws.dsts[0] is the own buffer
                        for m in deferred:
                            ops.copy_rows(d[m], pitch, ws.dsts[0][m], pitch, dim * esz, b_loc, ws.side.cuda_stream)
                    ev_sent = torch.cuda.Event()
                    ev_sent.record(ws.side)
            ws.hz.barrier(channel=0)
            z_glob3 = ws.z.view(b_glob, n, dim)
            z_all = [z_glob3[:, m] for m in range(n)]
            z_own = [z[row_offset:row_offset + b_loc] for z in z_all]
            # statistics exchange: the reduce kernel writes this rank's slot (column sum-exp partials, row sum-exp and
            # positives of its rows), ONE barrier, then the finalise kernel pulls every rank's slot over NVLink, adds
            # the column partials in rank order (bit-identical everywhere) and finalises ALL rows - the row LSEs of all
            # ranks are needed by the backward
            ops.ntxent_fwd_sharded([z_own[a] for a, _ in pairs], [z_all[b] for _, b in pairs], rank, world, inv_tau,
                                   ws.stats_addrs, None, op_format)
            if ev_sent is not None:  # this rank's deferred rows have left before it enters the barrier: after the barrier
                torch.cuda.current_stream(dev).wait_event(ev_sent)  # every rank's have landed everywhere
            ws.hstats.barrier(channel=0)
            lse2_row_all, lse2_col, loss = ops.ntxent_finalize_sharded(p, b_loc, rank, world, inv_tau, alpha, ws.stats_addrs,
                                                                       0, dev)
            if needs_grad:
                _hold(ws, ctx)
            ctx.cfg = (inv_tau, float(alpha), op_format, tuple(pairs), row_offset, b_glob, float(grad_world_scale), n)
            ctx.save_for_backward(lse2_row_all, lse2_col, ws.z, *xs, *invs)
            return loss
        else:
            z_loc = torch.empty((b_loc, n * dim), dtype=dt, device=dev)
            z_loc3 = z_loc.view(b_loc, n, dim)
            _, invs, xs = ops.l2norm_fwd(feats, op_format, out=[z_loc3[:, m] for m in range(n)])
            z_glob = torch.empty((b_glob, n * dim), dtype=dt, device=dev)
            dist.all_gather_into_tensor(z_glob, z_loc, group=group)
        z_glob3 = z_glob.view(b_glob, n, dim)
        z_all = [z_glob3[:, m] for m in range(n)]
        z_own = [z[row_offset:row_offset + b_loc] for z in z_all]  # local rows inside the gathered buffer
        row_sum, col_sum, diag2 = ops.ntxent_fwd([z_own[a] for a, _ in pairs], [z_all[b] for _, b in pairs],
                                                 row_offset, inv_tau, op_format)
        # NCCL: ONE all-reduce carries the column sum-exp partials and, each rank writing only its own row range of a
        # zeroed buffer, every rank's row sums and positives: afterwards every rank finalises ALL rows itself
        stats = torch.empty((3, p, b_glob), dtype=torch.float32, device=dev)
        stats.zero_()
        stats[0].copy_(col_sum)
        stats[1, :, row_offset:row_offset + b_loc].copy_(row_sum)
        stats[2, :, row_offset:row_offset + b_loc].copy_(diag2)
        dist.all_reduce(stats, group=group)
        lse2_row_all, lse2_col, _, loss = ops.ntxent_finalize(stats[1], stats[0], stats[2], 0, inv_tau, alpha,
                                                              want_loss=True)
        ctx.cfg = (inv_tau, float(alpha), op_format, tuple(pairs), row_offset, b_glob, float(grad_world_scale), n)
        ctx.save_for_backward(lse2_row_all, lse2_col, z_glob, *xs, *invs)
        return loss

    @staticmethod
    def backward(ctx, grad_losses):
        inv_tau, alpha, op_format, pairs, row_offset, b_glob, gscale, n = ctx.cfg
        saved = ctx.saved_tensors
        lse2_row_all, lse2_col, z_glob = saved[0], saved[1], saved[2]
        xs, invs = saved[3:3 + n], saved[3 + n:3 + 2 * n]
        b_loc, dim = xs[0].shape
        z_glob3 = z_glob.view(b_glob, n, dim)
        z_all = [z_glob3[:, m] for m in range(n)]
        z_own = [z[row_offset:row_offset + b_loc] for z in z_all]
        grad_losses = grad_losses.to(torch.float32)
        if gscale != 1.0:
            grad_losses = grad_losses * gscale
        grad_losses = grad_losses.contiguous()
        ws = getattr(ctx, "symm_ws", None)
        need = [bool(ctx.needs_input_grad[6 + m]) for m in range(n)]
        use_sg = getattr(ctx, "use_sg", _sharded_g_enabled(b_loc) and ops.ShardedBwdPlan.supported(b_loc, dim, ws.world if ws else 1))
        if ws is not None and any(need) and use_sg:
            # row block of G once per pair; column-side partials land in their owners' receive buffers over NVLink
            plan, work, addrs, hr = ws.sharded_bwd(pairs, need)
            rank = row_offset // b_loc
            flags = getattr(ctx, "flags", False)
            ops.ntxent_bwd_sharded_gemm(plan, z_all, rank, inv_tau, alpha, lse2_row_all, lse2_col, grad_losses, work,
                                        addrs, op_format, sync_addrs=ws.sync_addrs if flags else None)
            if not flags:
                hr.barrier(channel=0)  # every rank's partials have landed
            inv_all = invs[0]._base if invs[0]._base is not None and invs[0]._base.shape == (n, b_loc) else torch.stack(list(invs))
            grads = ops.ntxent_bwd_sharded_finish(plan, list(xs), inv_all, rank, work, addrs[rank],
                                                  sync_own_addr=ws.sync_addrs[rank] if flags else 0)
            ws.busy = False
            return (None, None, None, None, None, None, *grads)
        zts, ld_t = ops.transpose_for_bwd(z_all)  # none for the default dim-512 kernel
        jobs, owners = [], []
        sl = slice(row_offset, row_offset + b_loc)
        for m in range(n):
            if not ctx.needs_input_grad[6 + m]:
                continue
            segs = []
            for p, (a, b) in enumerate(pairs):
                if m == a:
                    segs.append(ops.BwdSegmentSpec(z_all[b], zts[b], lse2_row_all[p, sl], lse2_col[p],
                                                   grad_losses[p:p + 1], alpha, 1.0 - alpha))
                elif m == b:
                    segs.append(ops.BwdSegmentSpec(z_all[a], zts[a], lse2_col[p, sl], lse2_row_all[p],
                                                   grad_losses[p:p + 1], 1.0 - alpha, alpha))
            if segs:
                jobs.append(ops.BwdJobSpec(z_own[m], xs[m], invs[m], segs))
                owners.append(m)
        grads: List = [None] * n
        if jobs:
            for m, dx in zip(owners, ops.ntxent_bwd(jobs, b_glob, row_offset, ld_t, inv_tau, op_format)):
                grads[m] = dx
        if ws is not None:
            ws.busy = False  # the gathered operands may be overwritten by the next forward (after its first barrier)
        return (None, None, None, None, None, None, *grads)


def global_trimodal_ntxent(feats: Sequence[torch.Tensor], temperature: float, alpha: float, group=None,
                           op_format: int = DEFAULT_OP_FORMAT, grad_world_scale: float = 1.0) -> torch.Tensor:
    """Per-pair losses of the GLOBAL batch (identical on every rank); gradients are
    d(global loss)/d(local rows). Under DDP (which averages parameter gradients over ranks) pass
    grad_world_scale=world_size to reproduce the single-process global-batch gradient."""
    feats = list(feats)
    pairs = list(combinations(range(len(feats)), 2))
    return _GlobalNTXent.apply(float(temperature), float(alpha), op_format, pairs, group, grad_world_scale, *feats)


def global_calculate_losses(output_dict, loss_prefix, temperature, alpha, group=None, **kw):
    """Sharded counterpart of TriCoLoNet._calculate_losses (tricolo_net.py:56-65), same keys."""
    keys = list(output_dict.keys())
    losses = global_trimodal_ntxent([output_dict[k] for k in keys], temperature, alpha, group, **kw)
    out = {}
    for p, (a, b) in enumerate(combinations(keys, 2)):
        out[f"{loss_prefix}/{a[:-9]}_{b[:-9]}_loss"] = losses[p]
    out[f"{loss_prefix}/total_loss"] = sum(out.values())
    return out


def sharded_retrieve(text: torch.Tensor, gallery_shard: torch.Tensor, labels: torch.Tensor, idx_base: int,
                     k: int = 5, group=None, block_queries: int = None, operand_format: int = ops.BF16,
                     fused: bool = None):
    """Gallery-sharded retrieval. text [Q,D] (all queries, replicated), gallery_shard [G_loc,D] with global
    index base idx_base, labels [Q] global gallery indices. Returns (topk_val, topk_idx, rank) — identical on
    every rank.  fused: see tricolo_b200.evaluation.retrieve."""
    world, _ = _world(group)
    dt = ops.L.op_torch_dtype(operand_format)
    g16 = gallery_shard if gallery_shard.dtype == dt else ops.cast_16bit(gallery_shard, operand_format)
    n_q, n_g, dim = text.shape[0], gallery_shard.shape[0], text.shape[1]
    dev = text.device
    if fused is None:
        fused = dim % 64 == 0 and 64 <= dim <= 512 and k <= 16
    if block_queries is None:
        from .evaluation.eval_retrieval import two_kernel_block_queries

        block_queries = 148 * 128 * 8 if fused else two_kernel_block_queries(n_g)
    labels = labels.to(torch.int64)
    val = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    nb = torch.empty((n_q,), dtype=torch.int32, device=dev)
    if not fused:
        ld = (n_g + 31) // 32 * 32
        buf = torch.empty((min(block_queries, n_q), ld), dtype=torch.float32, device=dev)
    for s in range(0, n_q, block_queries):
        e = min(s + block_queries, n_q)
        tq = text[s:e]
        q16 = tq if tq.dtype == dt else ops.cast_16bit(tq, operand_format)
        if fused:
            gt = ops.gt_sim_mma(q16, g16, labels[s:e], idx_base)
            dist.all_reduce(gt, group=group)  # owner shard contributes, the others add 0
            v, i, b = ops.sim_topk_fused(q16, g16, k, labels[s:e], gt, idx_base)
        else:
            sim, _ = ops.sim_gemm(q16, g16, out=buf[: e - s])
            gt = ops.gather_gt_sim(sim, n_g, labels[s:e], idx_base)
            dist.all_reduce(gt, group=group)
            v, i, _, b = ops.topk_rank(sim, n_g, k, labels[s:e], idx_base, gt)
        val[s:e], idx[s:e], nb[s:e] = v, i, b
    cand_v = torch.empty((world * n_q, k), dtype=torch.float32, device=dev)
    cand_i = torch.empty((world * n_q, k), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(cand_v, val, group=group)
    dist.all_gather_into_tensor(cand_i, idx, group=group)
    dist.all_reduce(nb, group=group)
    mv, mi = ops.topk_merge(cand_v.view(world, n_q, k), cand_i.view(world, n_q, k))
    return mv, mi, nb + 1
