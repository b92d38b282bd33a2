"""The head stage of ``MSFWSI.forward`` (src/models/backbone.py:161-186, 205-212): 12 projectors + 12 predictors applied
to both views, executed DEPTH BY DEPTH over all heads at once instead of head by head.

The reference runs 24 ``nn.Sequential`` calls per view (60 Linear + 48 BatchNorm1d layers x 2 views = ~1100 kernels forward
and backward).  All 12 heads are independent and have the same layer structure

    projector:  Linear -> BN -> ReLU -> Linear -> BN -> ReLU -> Linear -> BN(affine=False)       = z
    predictor:  Linear(d -> d/4) -> BN -> ReLU -> Linear(d/4 -> d) + bias                         = p

so depth k of every head and both views goes into ONE grouped tcgen05 launch (``ops.gemm_grouped``), whose epilogue leaves
the batch-norm statistics of that depth and whose A-operand prologue applies the previous depth's batch norm + ReLU; one
small kernel per depth (``msf_head_bn_finalize``) turns the statistics into scale / shift, with the SyncBatchNorm exchange
(tools/ssl_train.py:160) inside it.  Forward: 5 GEMM + 4 finalize + 1 apply launches; backward: 1 + 5 GEMM + 4 x 3.
The backward rebuilds the ReLU activations from the saved raw Linear outputs (nothing but those is stored), runs dX and dW
of a depth in one launch (dW with K = the rows of both views, deterministic split-K), and the batch-norm backward as
reduce -> finalize (+ exchange) -> element-wise launches over all heads.

fp32 inputs / parameters outside autocast take the same schedule on the plain-FMA SIMT GEMM with explicit statistics and
apply kernels (exact fp32: the <= 1e-5 parity path).  There is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from .ops import GemmSpec

BN_EPS, BN_MOMENTUM = 1e-5, 0.1  # nn.BatchNorm1d defaults used at backbone.py:15,18,21,28


# ------------------------------------------------------------------------------------------------------------------
# cross-rank exchange workspace of the in-kernel SyncBatchNorm reduction
# ------------------------------------------------------------------------------------------------------------------
class HeadSync:
    """Symmetric workspace (``torch.distributed._symmetric_memory``) for ``msf_head_bn_finalize`` /
    ``msf_head_bn_bwd_finalize``: every rank's copy is mapped into every process, and the kernels exchange the batch-norm
    sums of a whole depth with NVLink loads / stores -- no NCCL launch.  Creating one is a collective."""

    CAPACITY = 1 << 16  # doubles per parity and source rank: 4 per (head, column) of a depth = 42240 for the reference widths
    TIMEOUT_MS = int(float(os.environ.get("MSFWSI_PEER_TIMEOUT_S", "120")) * 1000)
    _cache = {}

    def __init__(self, group, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        nbytes = L.lib().msf_head_sync_workspace_bytes(self.CAPACITY)
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        if self.world > L.MSF_PEER_MAX_WORLD:
            raise RuntimeError(f"HeadSync supports up to {L.MSF_PEER_MAX_WORLD} ranks")
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every workspace is zeroed before anybody publishes a flag
        self.seq = 0

    def next(self) -> Tuple[int, int, int, int, int, int]:
        """(peers ptr, world, rank, seq, capacity, timeout_ms) of the next exchange; all ranks call in the same order."""
        self.seq += 1
        return L.ptr(self.peers), self.world, self.rank, self.seq, self.CAPACITY, self.TIMEOUT_MS

    @classmethod
    def get(cls, group, device: torch.device) -> "HeadSync":
        key = (getattr(group, "group_name", None) or id(group), device.index)
        if key not in cls._cache:
            cls._cache[key] = cls(group, device)  # no silent NCCL fallback here: a box without peer memory raises
        return cls._cache[key]


_NO_SYNC = (0, 1, 0, 0, 0, 1)
_side_streams = {}


def _world(group) -> int:
    if group is None or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def _side_stream(device) -> torch.cuda.Stream:
    s = _side_streams.get(device.index)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _side_streams[device.index] = s
    return s


def _sync_args(group, device):
    if group is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return _NO_SYNC
    return HeadSync.get(group, device).next()


# ------------------------------------------------------------------------------------------------------------------
# parameter access: the modules of module.py are containers; the head stage reads their tensors
# ------------------------------------------------------------------------------------------------------------------
def _inner(m: torch.nn.Module) -> torch.nn.Module:
    """Look through a checkpoint_wrapper (backbone.py:106-119 wraps every Linear when use_checkpoint=True); the head stage
    recomputes its activations itself, so the wrapper has nothing left to do."""
    return getattr(m, "_checkpoint_wrapped_module", m)


class HeadRefs:
    """The modules of one head (projector nn.Sequential indices 0,1,3,4,6,7; predictor 0,1,3 -- backbone.py:12-31)."""

    def __init__(self, projector: torch.nn.Sequential, predictor: torch.nn.Sequential):
        self.lin = [_inner(projector[0]), _inner(projector[3]), _inner(projector[6]), _inner(predictor[0]), _inner(predictor[3])]
        self.bn = [projector[1], projector[4], projector[7], predictor[1]]
        self.d, self.dq = self.lin[0].in_features, self.lin[3].out_features

    def tensors(self) -> List[torch.Tensor]:
        """[W1, g1, b1, W2, g2, b2, W3, W4, g4, b4, W5, bias5] (BN3 has no affine)."""
        l, b = self.lin, self.bn
        return [l[0].weight, b[0].weight, b[0].bias, l[1].weight, b[1].weight, b[1].bias, l[2].weight, l[3].weight, b[3].weight, b[3].bias,
                l[4].weight, l[4].bias]


N_PARAM = 12
PRE_APPLY_MIN_WIDTH = 512  # heads wider than this materialise relu(bn(y)) instead of normalising inside the GEMM (see _HeadStage.forward)
_W_IDX = (0, 3, 6, 7, 10)          # positions of the five Linear weights in HeadRefs.tensors()
_G_IDX = ((1, 2), (4, 5), None, (8, 9))  # (gamma, beta) positions of BN1, BN2, BN3 (none), BN4


def _arr(cls, items):
    a = (cls * len(items))()
    for i, it in enumerate(items):
        a[i] = it
    return a


def _p2(a, b):
    return (C.c_void_p * 2)(a, b)


class _Carver:
    """Sub-allocates many small fp32 vectors / matrices from one tensor (one allocator call instead of hundreds)."""

    def __init__(self, numel: int, dtype: torch.dtype, device, zero: bool = False):
        self.buf = (torch.zeros if zero else torch.empty)(max(numel, 1), dtype=dtype, device=device)
        self.off = 0

    def take(self, *shape: int) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        pad = (n + 63) // 64 * 64  # keeps every piece 256-byte aligned for fp32, 128-byte for 16-bit
        t = self.buf[self.off:self.off + n].view(*shape)
        self.off += pad
        assert self.off <= self.buf.numel() + 64
        return t

    @staticmethod
    def size(shapes: Sequence[Tuple[int, ...]]) -> int:
        tot = 0
        for sh in shapes:
            n = 1
            for s in sh:
                n *= s
            tot += (n + 63) // 64 * 64
        return tot


class _HeadStage(torch.autograd.Function):
    """inputs: 12 stacked activations x0[h] (2, R_h, d_h) followed by 12 x N_PARAM parameter tensors.
    outputs: p[h] (2, R_h, d_h) for h in 0..11, then z[h] (non-differentiable: the reference detaches every z that leaves
    the module, backbone.py:188-191, 214-215; the gradient that reaches z through the predictor is handled inside)."""

    @staticmethod
    def forward(ctx, cfg, *tensors):
        heads: List[HeadRefs] = cfg["heads"]
        nh = len(heads)
        x0_in = tensors[:nh]
        params = tensors[nh:]
        L.require_cuda(*x0_in, *[t for t in params if t is not None])
        dev = x0_in[0].device
        dt = cfg["dtype"]
        fused = dt in (torch.bfloat16, torch.float16)  # tcgen05 GEMMs with statistics epilogue / batch-norm prologue
        training, group = cfg["training"], cfg["group"]
        want_hat, want_rowsq = cfg["want_keys"], cfg["want_rowsq"]
        code = L.dtype_code(dt)
        lib, st = L.lib(), L.stream_ptr()
        x0 = [ops._contig(t if t.dtype == dt else t.to(dt)) for t in x0_in]
        R = [int(t.shape[1]) for t in x0]
        for h, t in enumerate(x0):
            if t.dim() != 3 or t.shape[0] != 2 or t.shape[2] != heads[h].d:
                raise ValueError(f"head {h}: expected a (2, rows, {heads[h].d}) activation, got {tuple(t.shape)}")
        # operand copies of the Linear weights in the compute dtype (maintained by FusedAdam, see module.bind_optimizer)
        W = [[heads[h].lin[k].lowp_weight(dt) if fused else params[h * N_PARAM + _W_IDX[k]].detach() for k in range(5)] for h in range(nh)]
        widths = [[heads[h].d, heads[h].d, heads[h].d, heads[h].dq] for h in range(nh)]  # BN1..BN4 widths

        # Wide heads (the fuser heads: d = 576 .. 4608 over few rows) take their normalised activations from ONE element-wise
        # pass per depth instead of the GEMM's A prologue: a 256-row problem with N = 4608 has 36 column tiles (x split-K), and
        # each of those units would re-normalise the same A panel -- the four transform warps then bound the launch (measured
        # at B = 256: 201 us per projector depth with the prologue everywhere, 67 us for the same GEMMs without).  The narrow
        # heads (context / target, d <= 512: one to four column tiles over many rows) keep the prologue, where it saves a
        # round trip of the activation through HBM.  The materialised activations are kept for the dW GEMMs of the backward.
        pre = [fused and heads[h].d > PRE_APPLY_MIN_WIDTH for h in range(nh)]

        # ---- buffers ----
        grows = 128 if fused else 32  # rows per column-statistics group: one per GEMM M tile, or msf_head_bn_stats' 32-row groups
        act_shapes, f32_shapes = [], []
        for h in range(nh):
            d, dq, r = heads[h].d, heads[h].dq, R[h]
            act_shapes += [(2, r, d)] * 3 + [(2, r, dq)]                      # y1, y2, y3, y4 (raw Linear outputs)
            if not fused or pre[h]:
                act_shapes += [(2, r, d)] * 2 + [(2, r, dq)]                  # a1, a2, a4 kept (exact path; wide heads)
            for k in range(4):
                c = widths[h][k]
                f32_shapes += [(2, (r + grows - 1) // grows, 2, c)] + [(2, c)] * 4  # col_stats, scale, shift, mean, invstd
            if want_rowsq:
                f32_shapes += [(2, (d + 63) // 64, r)]
        acts = _Carver(_Carver.size(act_shapes), dt, dev)
        f32 = _Carver(_Carver.size(f32_shapes), torch.float32, dev)
        y = [[acts.take(2, R[h], heads[h].d) for _ in range(3)] + [acts.take(2, R[h], heads[h].dq)] for h in range(nh)]
        a_keep = [[acts.take(2, R[h], heads[h].d), acts.take(2, R[h], heads[h].d), None, acts.take(2, R[h], heads[h].dq)] if (not fused or pre[h]) else None
                  for h in range(nh)]
        cs = [[None] * 4 for _ in range(nh)]
        sc, sh, mu, istd = ([[None] * 4 for _ in range(nh)] for _ in range(4))
        rowsq = [None] * nh
        for h in range(nh):
            for k in range(4):
                c = widths[h][k]
                cs[h][k] = f32.take(2, (R[h] + grows - 1) // grows, 2, c)
                sc[h][k], sh[h][k], mu[h][k], istd[h][k] = (f32.take(2, c) for _ in range(4))
            if want_rowsq:
                rowsq[h] = f32.take(2, (heads[h].d + 63) // 64, R[h])
        z = [torch.empty((2, R[h], heads[h].d), dtype=dt, device=dev) for h in range(nh)]
        p = [torch.empty((2, R[h], heads[h].d), dtype=dt, device=dev) for h in range(nh)]
        # the L2-normalised keys of all heads live in ONE buffer: at world > 1 a single all-gather moves them
        khat = kflat = kgath = kready = None
        if want_hat:
            kflat = _Carver(_Carver.size([(2, R[h], heads[h].d) for h in range(nh)]), dt, dev)
            khat = [kflat.take(2, R[h], heads[h].d) for h in range(nh)]
        kinv = None

        def finalize(k):
            items = []
            for h in range(nh):
                bn = heads[h].bn[k]
                gi = _G_IDX[k]
                gam = params[h * N_PARAM + gi[0]] if gi else None
                bet = params[h * N_PARAM + gi[1]] if gi else None
                items.append(L.HeadBnItem(_p2(L.ptr(cs[h][k][0]), L.ptr(cs[h][k][1])), _p2(L.ptr(sc[h][k][0]), L.ptr(sc[h][k][1])),
                                          _p2(L.ptr(sh[h][k][0]), L.ptr(sh[h][k][1])), _p2(L.ptr(mu[h][k][0]), L.ptr(mu[h][k][1])),
                                          _p2(L.ptr(istd[h][k][0]), L.ptr(istd[h][k][1])), L.ptr(gam), L.ptr(bet), L.ptr(bn.running_mean),
                                          L.ptr(bn.running_var), R[h], widths[h][k], 2, 0 if fused else 1, grows))
            peers, world, rank, seq, cap, tmo = _sync_args(group, dev) if training else _NO_SYNC
            L.check(lib.msf_head_bn_finalize(_arr(L.HeadBnItem, items), nh, BN_EPS, BN_MOMENTUM, int(training), peers, world, rank, seq, cap, tmo, st),
                    "msf_head_bn_finalize")
            L.launch_count += 1

        def stats_of(k):  # exact path: column statistics straight from the fp32 Linear outputs
            mats = [L.HeadMat(L.ptr(y[h][k][v]), L.ptr(cs[h][k][v]), R[h], widths[h][k]) for h in range(nh) for v in range(2)]
            L.check(lib.msf_head_bn_stats(_arr(L.HeadMat, mats), len(mats), code, st), "msf_head_bn_stats")
            L.launch_count += 1

        def apply(items):
            for lo in range(0, len(items), L.MSF_HEAD_MAX_MATS):
                chunk = items[lo:lo + L.MSF_HEAD_MAX_MATS]
                L.check(lib.msf_head_bn_apply(_arr(L.HeadApplyItem, chunk), len(chunk), code, ops.COS_EPS, st), "msf_head_bn_apply")
                L.launch_count += 1

        def gemm(specs):
            ops.gemm_grouped(specs)

        # ---- depth 1..3: projector ----
        def pre_apply(k):  # a_k = relu(bn_k(y_k)) of the wide heads, one launch
            items = [L.HeadApplyItem(L.ptr(y[h][k][v]), L.ptr(a_keep[h][k][v]), 0, 0, L.ptr(sc[h][k][v]), L.ptr(sh[h][k][v]), 0, R[h], widths[h][k], 1, 0)
                     for h in range(nh) if pre[h] for v in range(2)]
            if items:
                apply(items)

        src = x0
        for k in range(3):
            if fused and k > 0:
                pre_apply(k - 1)
            specs = []
            for h in range(nh):
                d = heads[h].d
                for v in range(2):
                    a_in = a_keep[h][k - 1][v] if (fused and k > 0 and pre[h]) else src[h][v]
                    g = GemmSpec(a_in, W[h][k], R[h], d, d, C=y[h][k][v])
                    if fused:
                        g.col_stats = cs[h][k][v] if training else None
                        if k > 0 and not pre[h]:
                            g.a_scale, g.a_shift, g.a_relu = sc[h][k - 1][v], sh[h][k - 1][v], True
                    specs.append(g)
            gemm(specs)
            if not fused and training:
                stats_of(k)
            finalize(k)
            if fused:
                src = [y[h][k] for h in range(nh)]  # the next GEMM normalises on the fly
            elif k < 2:
                apply([L.HeadApplyItem(L.ptr(y[h][k][v]), L.ptr(a_keep[h][k][v]), 0, 0, L.ptr(sc[h][k][v]), L.ptr(sh[h][k][v]), L.ptr(mu[h][k][v]), R[h], heads[h].d, 1, 0)
                       for h in range(nh) for v in range(2)])
                src = [a_keep[h][k] for h in range(nh)]
        # ---- z = BN3(y3) (no ReLU); with it the L2-normalised keys of the InfoNCE objective ----
        if want_hat:
            kinv = [torch.empty((2, R[h]), dtype=torch.float32, device=dev) for h in range(nh)]
        apply([L.HeadApplyItem(L.ptr(y[h][2][v]), L.ptr(z[h][v]), L.ptr(khat[h][v]) if want_hat else 0, L.ptr(kinv[h][v]) if want_hat else 0,
                               L.ptr(sc[h][2][v]), L.ptr(sh[h][2][v]), 0 if fused else L.ptr(mu[h][2][v]), R[h], heads[h].d, 0, 0) for h in range(nh) for v in range(2)])
        if want_hat and training and _world(group) > 1:
            # ONE exchange for the keys of all 24 pairs, on a side stream: it overlaps the two predictor depths below, and
            # the loss kernels (which need the predictor outputs anyway) wait for it
            side = _side_stream(dev)
            done = torch.cuda.Event()
            done.record()
            kgath = torch.empty(_world(group) * kflat.buf.numel(), dtype=dt, device=dev)
            with torch.cuda.stream(side):
                side.wait_event(done)
                dist.all_gather_into_tensor(kgath, kflat.buf, group=group)
                kready = torch.cuda.Event()
                kready.record(side)
            kgath.record_stream(side)
            kflat.buf.record_stream(side)
        # ---- depth 4, 5: predictor ----
        specs = []
        for h in range(nh):
            for v in range(2):
                g = GemmSpec(z[h][v], W[h][3], R[h], heads[h].dq, heads[h].d, C=y[h][3][v])
                if fused and training:
                    g.col_stats = cs[h][3][v]
                specs.append(g)
        gemm(specs)
        if not fused and training:
            stats_of(3)
        finalize(3)
        if not fused:
            apply([L.HeadApplyItem(L.ptr(y[h][3][v]), L.ptr(a_keep[h][3][v]), 0, 0, L.ptr(sc[h][3][v]), L.ptr(sh[h][3][v]), L.ptr(mu[h][3][v]), R[h], heads[h].dq, 1, 0)
                   for h in range(nh) for v in range(2)])
        if fused:
            pre_apply(3)
        specs = []
        for h in range(nh):
            bias = params[h * N_PARAM + 11]
            for v in range(2):
                if fused:
                    if pre[h]:
                        g = GemmSpec(a_keep[h][3][v], W[h][4], R[h], heads[h].d, heads[h].dq, C=p[h][v], bias=bias)
                    else:
                        g = GemmSpec(y[h][3][v], W[h][4], R[h], heads[h].d, heads[h].dq, C=p[h][v], bias=bias, a_scale=sc[h][3][v], a_shift=sh[h][3][v],
                                     a_relu=True)
                    if want_rowsq:
                        g.row_sumsq = rowsq[h][v]
                else:
                    g = GemmSpec(a_keep[h][3][v], W[h][4], R[h], heads[h].d, heads[h].dq, C=p[h][v], bias=bias)
                specs.append(g)
        gemm(specs)

        ctx.cfg = cfg
        ctx.meta = (nh, R, dt, fused, training, [t.dtype for t in x0_in], [None if t is None else t.dtype for t in params])
        # raw Linear outputs + statistics + operand weights: everything the backward rebuilds the rest from
        ctx.keep = (x0, y, z, a_keep, sc, sh, mu, istd, W, acts.buf, f32.buf)
        ctx.save_for_backward(*[t for t in params if t is not None])
        ctx.mark_non_differentiable(*z)
        cfg["extras"] = {"khat": khat, "kinv": kinv, "rowsq": rowsq, "kflat": None if kflat is None else kflat.buf, "kgathered": kgath,
                         "kready": kready, "world": _world(group) if (want_hat and training) else 1,
                         "rank": dist.get_rank(group) if (want_hat and training and _world(group) > 1) else 0}
        return (*p, *z)

    @staticmethod
    def backward(ctx, *grads):
        nh, R, dt, fused, training, x_dtypes, p_dtypes = ctx.meta
        cfg = ctx.cfg
        heads: List[HeadRefs] = cfg["heads"]
        group = cfg["group"]
        x0, y, z, a_keep, sc, sh, mu, istd, W, _, _ = ctx.keep
        dev = x0[0].device
        code = L.dtype_code(dt)
        lib, st = L.lib(), L.stream_ptr()
        gp = []
        for h in range(nh):
            g = grads[h]
            gp.append(torch.zeros_like(z[h]) if g is None else ops._contig(g if g.dtype == dt else g.to(dt)))
        widths = [[heads[h].d, heads[h].d, heads[h].d, heads[h].dq] for h in range(nh)]

        act_shapes, f32_shapes = [], []
        for h in range(nh):
            d, dq, r = heads[h].d, heads[h].dq, R[h]
            act_shapes += [(2, r, dq), (2, r, dq), (2, r, d), (2, r, d)]       # da4/dy4 ... reused per depth: g buffer + dy buffer at both widths
            if fused and a_keep[h] is None:
                act_shapes += [(2, r, d)] * 2 + [(2, r, dq)]                    # a1, a2, a4 rebuilt (narrow heads: the forward normalised on the fly)
            rbs = (r + 255) // 256
            f32_shapes += [(2, rbs, 2, d)] * 4 + [(2, rbs, 2, dq)]              # reduce partials: BN1..3 + bias column sums (width d), BN4
            f32_shapes += [(2, d)] * 6 + [(2, dq)] * 2                          # c1, c2 per BN
        acts = _Carver(_Carver.size(act_shapes), dt, dev)
        f32 = _Carver(_Carver.size(f32_shapes), torch.float32, dev)
        gq, dyq, gd, dyd = ([None] * nh for _ in range(4))
        a = [[None] * 4 for _ in range(nh)]
        part = [[None] * 5 for _ in range(nh)]
        c1, c2 = ([[None] * 4 for _ in range(nh)] for _ in range(2))
        for h in range(nh):
            d, dq, r = heads[h].d, heads[h].dq, R[h]
            gq[h], dyq[h], gd[h], dyd[h] = acts.take(2, r, dq), acts.take(2, r, dq), acts.take(2, r, d), acts.take(2, r, d)
            if a_keep[h] is None:
                a[h][0], a[h][1], a[h][3] = acts.take(2, r, d), acts.take(2, r, d), acts.take(2, r, dq)
            else:
                a[h] = a_keep[h]  # exact path and wide heads: kept by the forward
            rbs = (r + 255) // 256
            for k in range(5):
                part[h][k] = f32.take(2, rbs, 2, dq if k == 3 else d)
            for k in range(4):
                c1[h][k], c2[h][k] = f32.take(2, widths[h][k]), f32.take(2, widths[h][k])

        # parameter gradients (fp32, the parameters' dtype)
        dW = [[torch.empty_like(W[h][k], dtype=torch.float32) for k in range(5)] for h in range(nh)]
        dgam = [[torch.empty(widths[h][k], dtype=torch.float32, device=dev) if _G_IDX[k] else None for k in range(4)] for h in range(nh)]
        dbet = [[torch.empty(widths[h][k], dtype=torch.float32, device=dev) if _G_IDX[k] else None for k in range(4)] for h in range(nh)]
        dbias = [torch.empty(heads[h].d, dtype=torch.float32, device=dev) for h in range(nh)]
        need_dx = [ctx.needs_input_grad[1 + h] for h in range(nh)]
        dx0 = [torch.empty_like(x0[h]) if need_dx[h] else None for h in range(nh)]

        def chunks(items, cap):
            for lo in range(0, len(items), cap):
                yield items[lo:lo + cap]

        if fused:  # rebuild the ReLU activations the dW GEMMs read (one launch for all narrow heads, layers and views)
            items = [L.HeadApplyItem(L.ptr(y[h][k][v]), L.ptr(a[h][k][v]), 0, 0, L.ptr(sc[h][k][v]), L.ptr(sh[h][k][v]), 0, R[h], widths[h][k], 1, 0)
                     for h in range(nh) if a_keep[h] is None for k in (0, 1, 3) for v in range(2)]
            for ch in chunks(items, L.MSF_HEAD_MAX_MATS):
                L.check(lib.msf_head_bn_apply(_arr(L.HeadApplyItem, ch), len(ch), code, ops.COS_EPS, st), "msf_head_bn_apply")
                L.launch_count += 1

        def bn_backward(k, g_in, dy_out, with_bias_sums=False):
            """g_in[h] (2,R,C): gradient w.r.t. relu?(bn_k(y_k)); writes dy_out[h] = gradient w.r.t. y_k."""
            relu = 0 if k == 2 else 1
            cen = 0 if fused else 1  # the exact path evaluated (y - mean) * scale + beta in the forward: rebuild the mask the same way
            items = [L.HeadBwdItem(L.ptr(g_in[h][v]), L.ptr(y[h][k][v]), 0, L.ptr(part[h][k][v]), L.ptr(sc[h][k][v]), L.ptr(sh[h][k][v]),
                                   L.ptr(mu[h][k][v]), L.ptr(istd[h][k][v]), 0, 0, R[h], widths[h][k], relu, cen) for h in range(nh) for v in range(2)]
            if with_bias_sums:  # column sums of dp = the bias gradient of the predictor's last Linear, in the same launch
                items += [L.HeadBwdItem(L.ptr(gp[h][v]), 0, 0, L.ptr(part[h][4][v]), 0, 0, 0, 0, 0, 0, R[h], heads[h].d, 0, 0) for h in range(nh) for v in range(2)]
            for ch in chunks(items, L.MSF_HEAD_MAX_MATS):
                L.check(lib.msf_head_bn_bwd_reduce(_arr(L.HeadBwdItem, ch), len(ch), code, st), "msf_head_bn_bwd_reduce")
                L.launch_count += 1
            fin = [L.HeadBwdFinItem(_p2(L.ptr(part[h][k][0]), L.ptr(part[h][k][1])), _p2(L.ptr(c1[h][k][0]), L.ptr(c1[h][k][1])),
                                    _p2(L.ptr(c2[h][k][0]), L.ptr(c2[h][k][1])), L.ptr(dgam[h][k]), L.ptr(dbet[h][k]), R[h], widths[h][k], 2, 0) for h in range(nh)]
            if with_bias_sums:
                fin += [L.HeadBwdFinItem(_p2(L.ptr(part[h][4][0]), L.ptr(part[h][4][1])), _p2(0, 0), _p2(0, 0), 0, L.ptr(dbias[h]), R[h], heads[h].d, 2, 1)
                        for h in range(nh)]
            peers, world, rank, seq, cap, tmo = _sync_args(group, dev) if training else _NO_SYNC
            L.check(lib.msf_head_bn_bwd_finalize(_arr(L.HeadBwdFinItem, fin), len(fin), int(training), peers, world, rank, seq, cap, tmo, st),
                    "msf_head_bn_bwd_finalize")
            L.launch_count += 1
            items = [L.HeadBwdItem(L.ptr(g_in[h][v]), L.ptr(y[h][k][v]), L.ptr(dy_out[h][v]), 0, L.ptr(sc[h][k][v]), L.ptr(sh[h][k][v]),
                                   L.ptr(mu[h][k][v]), L.ptr(istd[h][k][v]), L.ptr(c1[h][k][v]), L.ptr(c2[h][k][v]), R[h], widths[h][k], relu, cen)
                     for h in range(nh) for v in range(2)]
            for ch in chunks(items, L.MSF_HEAD_MAX_MATS):
                L.check(lib.msf_head_bn_bwd_elemt(_arr(L.HeadBwdItem, ch), len(ch), code, st), "msf_head_bn_bwd_elemt")
                L.launch_count += 1

        def gemm_bwd(k, dy, a_in, dx_out):
            """Layer k (0-based Linear index): dX = dY W_k per view (if dx_out) and dW_k = dY^T A over both views, one launch."""
            specs = []
            for h in range(nh):
                fout, fin = W[h][k].shape
                if dx_out is not None and dx_out[h] is not None:
                    for v in range(2):
                        specs.append(GemmSpec(dy[h][v], W[h][k], R[h], fin, fout, b_is_kn=True, C=dx_out[h][v]))
                specs.append(GemmSpec(dy[h].view(2 * R[h], fout), a_in[h].view(2 * R[h], fin), fout, fin, 2 * R[h], a_is_km=True, b_is_kn=True,
                                      out_dtype=torch.float32, C=dW[h][k]))
            ops.gemm_grouped(specs)

        # ---- predictor tail: p = a4 W5^T + b5 ----
        gemm_bwd(4, gp, [a[h][3] for h in range(nh)], gq)                     # da4 -> gq, dW5
        bn_backward(3, gq, dyq, with_bias_sums=True)                           # dy4 -> dyq, d gamma4 / beta4, d bias5
        gemm_bwd(3, dyq, z, gd)                                                # dz -> gd, dW4   (z carries no other gradient: it leaves detached)
        bn_backward(2, gd, dyd)                                                # dy3 -> dyd
        gemm_bwd(2, dyd, [a[h][1] for h in range(nh)], gd)                     # da2 -> gd, dW3
        bn_backward(1, gd, dyd)                                                # dy2
        gemm_bwd(1, dyd, [a[h][0] for h in range(nh)], gd)                     # da1, dW2
        bn_backward(0, gd, dyd)                                                # dy1
        gemm_bwd(0, dyd, x0, dx0)                                              # dx0 (only where needed), dW1

        out: List[Optional[torch.Tensor]] = [None]  # cfg
        for h in range(nh):
            out.append(None if dx0[h] is None else (dx0[h] if x_dtypes[h] == dt else dx0[h].to(x_dtypes[h])))
        for h in range(nh):
            per = [dW[h][0], dgam[h][0], dbet[h][0], dW[h][1], dgam[h][1], dbet[h][1], dW[h][2], dW[h][3], dgam[h][3], dbet[h][3], dW[h][4], dbias[h]]
            for j, g in enumerate(per):
                pd = p_dtypes[h * N_PARAM + j]
                out.append(None if pd is None else (g if g.dtype == pd else g.to(pd)))
        ctx.keep = None
        return tuple(out)


def head_stage(x0: Sequence[torch.Tensor], heads: Sequence[HeadRefs], training: bool, group=None, dtype: Optional[torch.dtype] = None,
               want_keys: bool = False, want_rowsq: bool = False):
    """Run all heads over stacked two-view activations ``x0[h]`` (2, rows_h, d_h).  Returns (p list, z list, extras):
    ``p[h]``, ``z[h]`` (2, rows_h, d_h); extras = {"khat": L2-normalised z (InfoNCE keys) or None, "kinv", "rowsq": per-64-
    column-block sum of squares of p rows or None}."""
    if dtype is None:
        dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x0[0].dtype
    if dtype not in (torch.bfloat16, torch.float16, torch.float32):
        raise TypeError(f"head_stage: unsupported compute dtype {dtype}")
    cfg = {"heads": list(heads), "training": bool(training), "group": group, "dtype": dtype, "want_keys": bool(want_keys), "want_rowsq": bool(want_rowsq)}
    params = [t for h in heads for t in h.tensors()]
    nh = len(heads)
    out = _HeadStage.apply(cfg, *x0, *params)
    return list(out[:nh]), list(out[nh:]), cfg.pop("extras")
