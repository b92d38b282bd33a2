"""Host-side operators over the C ABI: torch.autograd.Functions that hand raw device pointers
and the current CUDA stream to libmsfwsi_b200.so.  PyTorch is plumbing here (device memory,
streams, autograd graph); every operator fails loudly without the CUDA library."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib as L

COS_EPS = 1e-8  # nn.CosineSimilarity default (tools/ssl_train.py:422)


def _contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------
# A1: inverse-jigsaw gather + fuser concat      (src/models/backbone.py:147-158, 195-202)
# ------------------------------------------------------------------------------------------
class _GatherConcat(torch.autograd.Function):
    """``stacked=False``: one (tgt_sorted, ms_f) pair per item.  ``stacked=True``: the items are [view 0 levels ..., view 1
    levels ...] and the two views of a level are written into ONE (2, ...) tensor each -- the layout the head stage consumes
    (its weight-gradient GEMMs contract over the rows of both views at once)."""

    @staticmethod
    def forward(ctx, K: int, n_keep: int, n_items: int, validate: bool, stacked: bool, *tensors):
        ctx_f = [_contig(t) for t in tensors[:n_items]]
        tgt_f = [_contig(t) for t in tensors[n_items:2 * n_items]]
        rev = [_contig(t) for t in tensors[2 * n_items:3 * n_items]]
        L.require_cuda(*ctx_f, *tgt_f, *rev)
        dt = tgt_f[0].dtype
        B = ctx_f[0].shape[0]
        dev = tgt_f[0].device
        nl = n_items // 2 if stacked else n_items
        if stacked and n_items % 2:
            raise ValueError("gather_concat: stacked output needs the items of two views")
        items = (L.GatherItem * n_items)()
        outs_sorted, outs_ms = [], []
        if stacked:
            for i in range(nl):
                d = ctx_f[i].shape[1]
                outs_sorted.append(torch.empty((2, B * K, d), dtype=dt, device=dev))
                outs_ms.append(torch.empty((2, B, (n_keep + 1) * d), dtype=dt, device=dev))
        for i in range(n_items):
            if tgt_f[i].dtype != dt or ctx_f[i].dtype != dt:
                raise TypeError("gather_concat: all feature tensors must share one dtype")
            d = ctx_f[i].shape[1]
            if tgt_f[i].shape != (B * K, d) or ctx_f[i].shape != (B, d):
                raise ValueError(f"gather_concat: item {i} has shapes {tuple(ctx_f[i].shape)} / {tuple(tgt_f[i].shape)}; "
                                 f"expected ({B},{d}) / ({B * K},{d})")
            if rev[i].shape != (B, K) or rev[i].dtype != torch.int64:
                # the reference asserts batch_idx.shape == jigsaw_idx[v].shape (backbone.py:152)
                raise AssertionError(f"jigsaw_idx must be int64 of shape ({B},{K}); got {rev[i].dtype} {tuple(rev[i].shape)}")
            if stacked:
                if ctx_f[i].shape[1] != ctx_f[i % nl].shape[1]:
                    raise ValueError("gather_concat: the two views of a level must have the same width")
                s, m = outs_sorted[i % nl][i // nl], outs_ms[i % nl][i // nl]
            else:
                s = torch.empty_like(tgt_f[i])
                m = torch.empty((B, (n_keep + 1) * d), dtype=dt, device=dev)
                outs_sorted.append(s)
                outs_ms.append(m)
            items[i] = L.GatherItem(L.ptr(tgt_f[i]), L.ptr(ctx_f[i]), L.ptr(rev[i]), L.ptr(s), L.ptr(m), d, 0)
        flag = torch.zeros(1, dtype=torch.int32, device=dev) if validate else None
        L.check(L.lib().msf_gather_concat_fwd(items, n_items, B, K, n_keep, L.dtype_code(dt), L.ptr(flag), L.stream_ptr()),
                "msf_gather_concat_fwd")
        L.launch_count += 1
        if validate:  # host syncs: only on the validating path
            if int(flag.item()) != 0:
                raise IndexError(f"jigsaw_idx holds values outside [-{K}, {K})")
            want = torch.arange(K, device=rev[0].device)
            for r in {t.data_ptr(): t for t in rev}.values():
                if not bool((torch.sort(r.remainder(K), dim=1).values == want).all().item()):
                    raise ValueError("jigsaw_idx rows must be permutations of range(K) (argsort(randperm(K)), bcss.py:171-177): "
                                     "the backward scatters by them")
        ctx.save_for_backward(*rev)
        ctx.meta = (K, n_keep, n_items, B, dt, [t.shape[1] for t in ctx_f], stacked)
        return (*outs_sorted, *outs_ms)

    @staticmethod
    def backward(ctx, *grads):
        K, n_keep, n_items, B, dt, dims, stacked = ctx.meta
        rev = ctx.saved_tensors
        nl = n_items // 2 if stacked else n_items
        g_sorted, g_ms = grads[:nl], grads[nl:]
        items = (L.GatherGradItem * n_items)()
        keep, g_ctx, g_tgt = [], [], []
        gs_all = [None if g is None else _contig(g).to(dt) for g in g_sorted]
        gm_all = [None if g is None else _contig(g).to(dt) for g in g_ms]
        for i in range(n_items):
            if stacked:
                gs = None if gs_all[i % nl] is None else gs_all[i % nl][i // nl]
                gm = None if gm_all[i % nl] is None else gm_all[i % nl][i // nl]
            else:
                gs, gm = gs_all[i], gm_all[i]
            keep += [gs, gm]
            d = dims[i]
            dev = rev[i].device
            gt = torch.empty((B * K, d), dtype=dt, device=dev)
            gc = torch.empty((B, d), dtype=dt, device=dev)
            g_tgt.append(gt)
            g_ctx.append(gc)
            items[i] = L.GatherGradItem(L.ptr(gs), L.ptr(gm), L.ptr(rev[i]), L.ptr(gt), L.ptr(gc), d, 0)
        L.check(L.lib().msf_gather_concat_bwd(items, n_items, B, K, n_keep, L.dtype_code(dt), L.stream_ptr()),
                "msf_gather_concat_bwd")
        L.launch_count += 1
        return (None, None, None, None, None, *g_ctx, *g_tgt, *([None] * n_items))


def gather_concat(ctx_f: Sequence[torch.Tensor], tgt_f: Sequence[torch.Tensor], rev: Sequence[torch.Tensor],
                  K: int = 16, n_keep: int = 8, validate: bool = False, stacked: bool = False):
    """All items in one launch.  ``ctx_f[i]`` (B,d_i), ``tgt_f[i]`` (B*K,d_i) shuffled, ``rev[i]`` (B,K) int64
    -> ``(tgt_sorted, ms_f)`` lists.  The forward is defined for any in-range ``rev`` (``ms_f`` is built from
    ``tgt_f`` alone, as in the reference); the backward needs ``rev`` rows to be permutations (argsort(randperm),
    bcss.py:171-177) -- ``validate=True`` checks both, at the cost of host syncs.
    ``stacked=True``: items = [view-0 levels ..., view-1 levels ...]; returns one (2, B*K, d) / (2, B, (n_keep+1)*d) tensor
    per level instead of one per (level, view)."""
    n = len(ctx_f)
    if not (n == len(tgt_f) == len(rev)) or n == 0 or n > L.MSF_GATHER_MAX_ITEMS:
        raise ValueError("gather_concat: need 1..16 (ctx, tgt, rev) triples")
    out = _GatherConcat.apply(K, n_keep, n, validate, stacked, *ctx_f, *tgt_f, *rev)
    no = n // 2 if stacked else n
    return list(out[:no]), list(out[no:])


# ------------------------------------------------------------------------------------------
# L1 cosine mode: the loss block of tools/ssl_train.py:448-466 in one launch
# ------------------------------------------------------------------------------------------
class _CosineLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coefs: Tuple[float, ...], eps: float, *tensors):
        n = len(coefs)
        ps = [_contig(t) for t in tensors[:n]]
        zs = [_contig(t.detach()) for t in tensors[n:]]
        L.require_cuda(*ps, *zs)
        dts = {t.dtype for t in ps} | {t.dtype for t in zs}
        dt = ps[0].dtype if len(dts) == 1 else torch.float32  # mixed inputs: exact up-cast, fp32 kernel
        ps_k = [t.to(dt) for t in ps]
        zs_k = [t.to(dt) for t in zs]
        dev = ps[0].device
        pairs = (L.CosPair * n)()
        total_rows = sum(p.shape[0] for p in ps_k)
        stats = torch.empty((max(total_rows, 1), 4), dtype=torch.float32, device=dev)
        off = 0
        for i in range(n):
            if ps_k[i].shape != zs_k[i].shape or ps_k[i].dim() != 2:
                raise ValueError(f"cosine_loss: pair {i} shapes {tuple(ps_k[i].shape)} vs {tuple(zs_k[i].shape)}")
            rows, dim = ps_k[i].shape
            pairs[i] = L.CosPair(L.ptr(ps_k[i]), L.ptr(zs_k[i]), stats[off:].data_ptr(), 0, rows, dim, float(coefs[i]))
            off += rows
        ws_bytes = L.lib().msf_cosine_loss_workspace_bytes(pairs, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L.check(L.lib().msf_cosine_loss_fwd(pairs, n, L.dtype_code(dt), eps, L.ptr(loss), L.ptr(ws), ws_bytes, L.stream_ptr()),
                "msf_cosine_loss_fwd")
        L.launch_count += 2
        ctx.save_for_backward(stats, *ps_k, *zs_k)
        ctx.meta = (coefs, n, dt, [t.dtype for t in ps])
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        coefs, n, dt, orig_dt = ctx.meta
        stats, *rest = ctx.saved_tensors
        ps, zs = rest[:n], rest[n:]
        g = _contig(grad_out.to(torch.float32))
        pairs = (L.CosPair * n)()
        grads = []
        off = 0
        for i in range(n):
            rows, dim = ps[i].shape
            gp = torch.empty_like(ps[i])
            grads.append(gp)
            pairs[i] = L.CosPair(L.ptr(ps[i]), L.ptr(zs[i]), stats[off:].data_ptr(), L.ptr(gp), rows, dim, float(coefs[i]))
            off += rows
        L.check(L.lib().msf_cosine_loss_bwd(pairs, n, L.dtype_code(dt), L.ptr(g), L.stream_ptr()), "msf_cosine_loss_bwd")
        L.launch_count += 1
        grads = [gp.to(orig_dt[i]) for i, gp in enumerate(grads)]
        return (None, None, *grads, *([None] * n))


class _CosineLossStacked(torch.autograd.Function):
    """The loss block over two-view stacks: for every head h, pairs (p[h][0], z[h][1]) and (p[h][1], z[h][0]) with
    coefficient coefs[h] each (tools/ssl_train.py:449-464: cos(p1, z2) and cos(p2, z1)).  One forward launch (+ the
    1-CTA final sum) and one backward launch that writes the gradient of both views of a head into one (2, rows, dim) tensor."""

    @staticmethod
    def forward(ctx, coefs: Tuple[float, ...], eps: float, *tensors):
        nh = len(coefs)
        ps = [_contig(t) for t in tensors[:nh]]
        zs = [_contig(t.detach()) for t in tensors[nh:]]
        L.require_cuda(*ps, *zs)
        dts = {t.dtype for t in ps} | {t.dtype for t in zs}
        dt = ps[0].dtype if len(dts) == 1 else torch.float32
        ps_k, zs_k = [t.to(dt) for t in ps], [t.to(dt) for t in zs]
        dev = ps[0].device
        n = 2 * nh
        pairs = (L.CosPair * n)()
        total_rows = sum(2 * p.shape[1] for p in ps_k)
        stats = torch.empty((max(total_rows, 1), 4), dtype=torch.float32, device=dev)
        off = 0
        for h in range(nh):
            if ps_k[h].shape != zs_k[h].shape or ps_k[h].dim() != 3 or ps_k[h].shape[0] != 2:
                raise ValueError(f"cosine_loss_stacked: head {h} shapes {tuple(ps_k[h].shape)} vs {tuple(zs_k[h].shape)} (expected (2, rows, dim))")
            _, rows, dim = ps_k[h].shape
            for v in range(2):
                pairs[2 * h + v] = L.CosPair(L.ptr(ps_k[h][v]), L.ptr(zs_k[h][1 - v]), stats[off:].data_ptr(), 0, rows, dim, float(coefs[h]))
                off += rows
        ws_bytes = L.lib().msf_cosine_loss_workspace_bytes(pairs, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L.check(L.lib().msf_cosine_loss_fwd(pairs, n, L.dtype_code(dt), eps, L.ptr(loss), L.ptr(ws), ws_bytes, L.stream_ptr()), "msf_cosine_loss_fwd")
        L.launch_count += 2
        ctx.save_for_backward(stats, *ps_k, *zs_k)
        ctx.meta = (coefs, nh, dt, [t.dtype for t in ps])
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        coefs, nh, dt, orig_dt = ctx.meta
        stats, *rest = ctx.saved_tensors
        ps, zs = rest[:nh], rest[nh:]
        g = _contig(grad_out.to(torch.float32))
        n = 2 * nh
        pairs = (L.CosPair * n)()
        grads = []
        off = 0
        for h in range(nh):
            _, rows, dim = ps[h].shape
            gp = torch.empty_like(ps[h])
            grads.append(gp)
            for v in range(2):
                pairs[2 * h + v] = L.CosPair(L.ptr(ps[h][v]), L.ptr(zs[h][1 - v]), stats[off:].data_ptr(), L.ptr(gp[v]), rows, dim, float(coefs[h]))
                off += rows
        L.check(L.lib().msf_cosine_loss_bwd(pairs, n, L.dtype_code(dt), L.ptr(g), L.stream_ptr()), "msf_cosine_loss_bwd")
        L.launch_count += 1
        grads = [gp if gp.dtype == orig_dt[h] else gp.to(orig_dt[h]) for h, gp in enumerate(grads)]
        return (None, None, *grads, *([None] * nh))


def cosine_loss_stacked(p_stacks: Sequence[torch.Tensor], z_stacks: Sequence[torch.Tensor], coefs: Sequence[float], eps: float = COS_EPS):
    """``sum_h coefs[h] * (mean_rows cos(p[h][0], z[h][1]) + mean_rows cos(p[h][1], z[h][0]))`` over (2, rows, dim) stacks."""
    nh = len(p_stacks)
    if not (nh == len(z_stacks) == len(coefs)) or nh == 0 or 2 * nh > L.MSF_COS_MAX_PAIRS:
        raise ValueError("cosine_loss_stacked: need 1..16 (p, z, coef) triples")
    return _CosineLossStacked.apply(tuple(float(c) for c in coefs), float(eps), *p_stacks, *z_stacks)


def cosine_loss(ps: Sequence[torch.Tensor], zs: Sequence[torch.Tensor], coefs: Sequence[float], eps: float = COS_EPS):
    """``sum_i coefs[i] * mean_rows cos(ps[i], zs[i])`` (fp32 scalar); zs are treated as detached."""
    n = len(ps)
    if not (n == len(zs) == len(coefs)) or n == 0 or n > L.MSF_COS_MAX_PAIRS:
        raise ValueError("cosine_loss: need 1..32 (p, z, coef) triples")
    return _CosineLoss.apply(tuple(float(c) for c in coefs), float(eps), *ps, *zs)


# ------------------------------------------------------------------------------------------
# L1 infonce mode (extension): fused flash-style InfoNCE with global negatives
# ------------------------------------------------------------------------------------------
def rownorm(x: torch.Tensor, out_dtype: torch.dtype, eps: float = COS_EPS):
    """x (rows, dim) -> (x / max(||x||, eps) in out_dtype, 1/max(||x||, eps) fp32)."""
    x = _contig(x)
    L.require_cuda(x)
    rows, dim = x.shape
    xh = torch.empty((rows, dim), dtype=out_dtype, device=x.device)
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device)
    L.check(L.lib().msf_rownorm(L.ptr(x), rows, dim, L.dtype_code(x.dtype), eps, L.ptr(xh), L.dtype_code(out_dtype), L.ptr(inv),
                                L.stream_ptr()), "msf_rownorm")
    L.launch_count += 1
    return xh, inv


def infonce_precision_for(dim: int, dtype: torch.dtype) -> torch.dtype:
    """bf16 tcgen05 paths for 16-bit inputs: flash kernel for dim in {64,128,256}, two-pass GEMMs for larger multiples
    of 64 (512 and the fuser widths); fp32 SIMT otherwise (fp32 inputs keep <=1e-5 parity)."""
    tc = dim in (64, 128, 256) or (dim > 256 and dim % 64 == 0)
    return torch.bfloat16 if (dtype in (torch.bfloat16, torch.float16) and tc) else torch.float32


def all_gather_keys(k_hat: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """Rank-major all-gather of the normalised keys.  Returns (keys_all, pos_offset).
    ``group=None`` is the default (world) group; ``group=False`` means "this rank holds every key" (no collective)."""
    if group is False or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return k_hat, 0
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = torch.empty((world * k_hat.shape[0], k_hat.shape[1]), dtype=k_hat.dtype, device=k_hat.device)
    dist.all_gather_into_tensor(out, k_hat, group=group)
    return out, rank * k_hat.shape[0]


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, z, tau: float, precision: torch.dtype, eps: float, group, detach_keys: bool):
        L.require_cuda(p, z)
        if p.dim() != 2 or p.shape != z.shape:
            raise ValueError(f"infonce: p {tuple(p.shape)} and z {tuple(z.shape)} must be equal 2-D shapes")
        nq, dim = p.shape
        q_hat, q_inv = rownorm(p, precision, eps)
        k_hat, k_inv = rownorm(z.detach(), precision, eps)
        k_all, pos_offset = all_gather_keys(k_hat, group)
        n_keys = k_all.shape[0]
        prec = L.dtype_code(precision)
        ws_bytes = L.lib().msf_infonce_workspace_bytes(nq, n_keys, dim, prec)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p.device)
        loss_sum = torch.empty((), dtype=torch.float32, device=p.device)
        L.check(L.lib().msf_infonce_fwd(L.ptr(q_hat), L.ptr(k_all), nq, n_keys, dim, pos_offset, tau, prec, L.ptr(loss_sum), 0,
                                        L.ptr(ws), ws_bytes, L.stream_ptr()), "msf_infonce_fwd")
        L.launch_count += 3
        ctx.save_for_backward(q_hat, k_all, q_inv, ws, k_hat, k_inv)
        ctx.meta = (nq, n_keys, dim, pos_offset, tau, prec, ws_bytes, p.dtype, z.dtype, group, detach_keys)
        return loss_sum / nq  # per-rank mean, like the reference's per-rank .mean() under DDP

    @staticmethod
    def backward(ctx, grad_out):
        nq, n_keys, dim, pos_offset, tau, prec, ws_bytes, p_dtype, z_dtype, group, detach_keys = ctx.meta
        q_hat, k_all, q_inv, ws, k_hat, k_inv = ctx.saved_tensors
        g = _contig(grad_out.to(torch.float32))
        grad_q = torch.empty((nq, dim), dtype=p_dtype, device=q_hat.device)
        L.check(L.lib().msf_infonce_bwd(L.ptr(q_hat), L.ptr(k_all), L.ptr(q_inv), nq, n_keys, dim, pos_offset, tau, prec, L.ptr(g),
                                        1.0 / nq, L.ptr(ws), ws_bytes, L.ptr(grad_q), L.dtype_code(p_dtype), L.stream_ptr()),
                "msf_infonce_bwd")
        L.launch_count += 1
        grad_z = None
        if not detach_keys and ctx.needs_input_grad[1]:
            grad_z = _infonce_key_grad(q_hat, k_all, k_hat, k_inv, ws, ws_bytes, g, nq, n_keys, dim, pos_offset, tau, prec, z_dtype, group)
        return grad_q, grad_z, None, None, None, None, None


def _infonce_key_grad(q_hat, k_all, k_hat, k_inv, ws, ws_bytes, g, nq, n_keys, dim, pos_offset, tau, prec, z_dtype, group):
    """north_star (4), keys NOT detached: this rank's queries contribute to the gradient of EVERY global key
    (``msf_infonce_dk``: the flash pass with the roles swapped); the partials are reduce-scattered to the keys' owners over
    NCCL (the backward of the rank-major all-gather), where ``msf_infonce_dk_finish`` subtracts the positives and applies the
    normalise Jacobian of z."""
    dev = q_hat.device
    dk_ws_bytes = L.lib().msf_infonce_dk_workspace_bytes(nq, n_keys, dim, prec)
    dk_ws = torch.empty(dk_ws_bytes, dtype=torch.uint8, device=dev)
    dk_all = torch.empty((n_keys, dim), dtype=torch.float32, device=dev)
    L.check(L.lib().msf_infonce_dk(L.ptr(q_hat), L.ptr(k_all), nq, n_keys, dim, pos_offset, tau, prec, L.ptr(g), 1.0 / nq, L.ptr(ws), ws_bytes,
                                   L.ptr(dk_all), L.ptr(dk_ws), dk_ws_bytes, L.stream_ptr()), "msf_infonce_dk")
    L.launch_count += 3
    rows = k_hat.shape[0]
    if n_keys != rows:  # keys were gathered: sum the partials of all ranks, every rank keeps the rows of its own keys
        dk_local = torch.empty((rows, dim), dtype=torch.float32, device=dev)
        dist.reduce_scatter_tensor(dk_local, dk_all, op=dist.ReduceOp.SUM, group=group if group is not False else None)
    else:
        dk_local = dk_all
    grad_z = torch.empty((rows, dim), dtype=z_dtype, device=dev)
    L.check(L.lib().msf_infonce_dk_finish(L.ptr(dk_local), L.ptr(q_hat), L.ptr(k_hat), L.ptr(k_inv), rows, nq, dim, tau, prec, L.ptr(g), 1.0 / nq,
                                          L.ptr(grad_z), L.dtype_code(z_dtype), L.stream_ptr()), "msf_infonce_dk_finish")
    L.launch_count += 1
    return grad_z


def infonce_loss(p: torch.Tensor, z: torch.Tensor, tau: float = 0.07, precision: Optional[torch.dtype] = None,
                 eps: float = COS_EPS, group=None, detach_keys: bool = True) -> torch.Tensor:
    """mean_i [logsumexp_j(p_hat_i . z_hat_j / tau) - p_hat_i . z_hat_pos(i) / tau] with keys all-gathered over
    ``group`` (rank-major).  Positive of local row i on rank r is global row r*rows+i.
    ``detach_keys=True`` (default) is the reference's semantics (every z is detached, backbone.py:188-191): z receives no
    gradient.  ``detach_keys=False`` is the non-detached variant of north_star (4): z receives the gradient of the losses of
    ALL ranks (each rank's loss being its local mean, as DDP sums them before averaging), reduce-scattered over ``group``."""
    if precision is None:
        precision = infonce_precision_for(p.shape[1], p.dtype)
    return _InfoNCE.apply(p, z, float(tau), precision, float(eps), group, bool(detach_keys))


class _InfoNCEGrouped(torch.autograd.Function):
    """All InfoNCE pairs of the loss block in one fused call (``msf_nce_grouped_fwd`` / ``_bwd``): for every head h the pairs
    (p[h][0], keys[h][1]) and (p[h][1], keys[h][0]) with weight coefs[h] (tools/ssl_train.py:449-464 pairing).  Queries are
    the raw bf16 predictor outputs (their norms come from ``rowsq``, the predictor-tail GEMM's epilogue); keys are the
    normalised projector outputs in rank-major blocks (this rank's, or the single all-gather of all heads' keys)."""

    @staticmethod
    def forward(ctx, cfg, *p_stacks):
        nh = len(p_stacks)
        ps = [_contig(t) for t in p_stacks]
        L.require_cuda(*ps)
        dev = ps[0].device
        khat, rowsq, coefs, tau = cfg["khat"], cfg["rowsq"], cfg["coefs"], cfg["tau"]
        world, rank, kgath, kflat = cfg["world"], cfg["rank"], cfg["kgathered"], cfg["kflat"]
        if cfg["kready"] is not None:
            torch.cuda.current_stream(dev).wait_event(cfg["kready"])  # the side-stream all-gather of the keys
        n = 2 * nh
        pairs = (L.NcePair * n)()
        stride = 0 if world == 1 else kflat.numel()
        for h in range(nh):
            if ps[h].dtype != torch.bfloat16 or ps[h].dim() != 3 or ps[h].shape[0] != 2 or ps[h].shape != khat[h].shape:
                raise ValueError(f"infonce_grouped: head {h}: expected (2, rows, dim) bfloat16 stacks, got {tuple(ps[h].shape)} {ps[h].dtype}")
            _, rows, dim = ps[h].shape
            for v in range(2):
                keys = khat[h][1 - v]
                if world > 1:  # the same block inside the gathered buffer: rank r's copy sits r * stride elements further
                    off = (keys.data_ptr() - kflat.data_ptr()) // 2
                    kptr = kgath.data_ptr() + 2 * off
                else:
                    kptr = keys.data_ptr()
                pairs[2 * h + v] = L.NcePair(L.ptr(ps[h][v]), 0 if rowsq is None else L.ptr(rowsq[h][v]), kptr, 0, stride, rows, rows, world, dim,
                                             rank if world > 1 else 0, float(coefs[h]))
        ws_bytes = L.lib().msf_nce_grouped_workspace_bytes(pairs, n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L.check(L.lib().msf_nce_grouped_fwd(pairs, n, L.MSF_BF16, tau, COS_EPS, L.ptr(loss), L.ptr(ws), ws_bytes, L.stream_ptr()), "msf_nce_grouped_fwd")
        L.launch_count += 7
        ctx.keep = (ps, khat, rowsq, kgath, kflat, ws, ws_bytes, pairs, n, tau)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        ps, khat, rowsq, kgath, kflat, ws, ws_bytes, pairs, n, tau = ctx.keep
        g = _contig(grad_out.to(torch.float32))
        grads = [torch.empty_like(t) for t in ps]
        for h in range(len(ps)):
            for v in range(2):
                pairs[2 * h + v].grad_q = L.ptr(grads[h][v])
        L.check(L.lib().msf_nce_grouped_bwd(pairs, n, L.MSF_BF16, tau, COS_EPS, L.ptr(g), L.ptr(ws), ws_bytes, L.stream_ptr()), "msf_nce_grouped_bwd")
        L.launch_count += 1
        ctx.keep = None
        return (None, *grads)


def infonce_grouped(p_stacks: Sequence[torch.Tensor], extras: dict, coefs: Sequence[float], tau: float = 0.07) -> torch.Tensor:
    """``sum_h coefs[h] * (InfoNCE(p[h][0], keys[h][1]) + InfoNCE(p[h][1], keys[h][0]))`` with global negatives: ``extras`` is
    what ``heads.head_stage(..., want_keys=True, want_rowsq=True)`` returned (normalised keys of all heads in one buffer,
    gathered across ranks once; row sums of squares of p).  Each term is the mean over the LOCAL queries (DDP averages)."""
    nh = len(p_stacks)
    if nh == 0 or 2 * nh > L.MSF_NCE_MAX_PAIRS or len(coefs) != nh:
        raise ValueError("infonce_grouped: need 1..16 (p, coef) pairs")
    cfg = dict(extras)
    cfg["coefs"], cfg["tau"] = [float(c) for c in coefs], float(tau)
    return _InfoNCEGrouped.apply(cfg, *p_stacks)


# ------------------------------------------------------------------------------------------
# G1: Linear layers of the heads on the tcgen05 GEMM      (src/models/backbone.py:14,17,20,27,30)
# ------------------------------------------------------------------------------------------
def gemm_bf16(A, B, M, N, K, a_is_km=False, b_is_kn=False, out_dtype=torch.bfloat16, alpha=1.0, bias=None):
    """C[M,N] = alpha * op(A) op(B) (+ bias) on the persistent tcgen05 kernel; A/B bf16 row-major 2-D tensors."""
    C_ = torch.empty((M, N), dtype=out_dtype, device=A.device)
    L.check(L.lib().msf_gemm_bf16(L.ptr(A), A.stride(0), L.ptr(B), B.stride(0), L.ptr(C_), C_.stride(0), M, N, K, int(a_is_km), int(b_is_kn),
                                  L.dtype_code(out_dtype), float(alpha), L.ptr(bias), L.stream_ptr()), "msf_gemm_bf16")
    L.launch_count += 1
    return C_


class GemmSpec:
    """One problem of a grouped launch (``msf_gemm_problem``): ``C[M,N] = epi(alpha * pro(A) op(B))``.
    ``A`` (M,K) row-major (or (K,M) with ``a_is_km``), ``B`` (N,K) (or (K,N) with ``b_is_kn``); 2-D CUDA tensors whose last
    dimension is contiguous (row stride = leading dimension).  Optional epilogue outputs are allocated by
    :func:`gemm_grouped` when requested (``want_col_stats`` / ``want_row_sumsq``)."""
    __slots__ = ("A", "B", "C", "M", "N", "K", "a_is_km", "b_is_kn", "out_dtype", "alpha", "bias", "col_stats", "row_sumsq", "a_scale", "a_shift",
                 "a_relu", "tile_n", "split_k", "no_tma_store")

    def __init__(self, A, B, M, N, K, a_is_km=False, b_is_kn=False, out_dtype=None, alpha=1.0, bias=None, C=None, col_stats=None, row_sumsq=None,
                 a_scale=None, a_shift=None, a_relu=False, tile_n=0, split_k=0, no_tma_store=False):
        self.A, self.B, self.C, self.M, self.N, self.K = A, B, C, int(M), int(N), int(K)
        self.a_is_km, self.b_is_kn, self.out_dtype, self.alpha, self.bias = bool(a_is_km), bool(b_is_kn), out_dtype, float(alpha), bias
        self.col_stats, self.row_sumsq, self.a_scale, self.a_shift, self.a_relu = col_stats, row_sumsq, a_scale, a_shift, bool(a_relu)
        self.tile_n, self.split_k, self.no_tma_store = int(tile_n), int(split_k), bool(no_tma_store)


_gemm_counters = {}


def _counters_for(device: torch.device) -> torch.Tensor:
    """The zeroed split-K tile counters of (device, current stream); every launch leaves them zero."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _gemm_counters.get(key)
    if t is None:
        t = torch.zeros(L.MSF_GEMM_MAX_COUNTERS, dtype=torch.int32, device=device)
        _gemm_counters[key] = t
    return t


def gemm_grouped(specs: Sequence[GemmSpec], want_col_stats: bool = False, want_row_sumsq: bool = False) -> List[torch.Tensor]:
    """All problems in ONE persistent tcgen05 launch (``msf_gemm_grouped``); more than MSF_GEMM_MAX_PROBLEMS are chunked.
    Operands must share one 16-bit dtype (bf16 or fp16).  Returns the C tensors (allocated here unless given)."""
    if not specs:
        return []
    op_dt = specs[0].A.dtype
    if op_dt == torch.float32:
        return _gemm_grouped_f32(specs)
    if op_dt not in (torch.bfloat16, torch.float16):
        raise TypeError(f"gemm_grouped: operands must be bfloat16, float16 or float32, got {op_dt}")
    dev = specs[0].A.device
    outs = []
    for lo in range(0, len(specs), L.MSF_GEMM_MAX_PROBLEMS):
        chunk = specs[lo:lo + L.MSF_GEMM_MAX_PROBLEMS]
        arr = (L.GemmProblem * len(chunk))()
        keep = []
        for i, g in enumerate(chunk):
            L.require_cuda(g.A, g.B)
            if g.A.dtype != op_dt or g.B.dtype != op_dt or g.A.dim() != 2 or g.B.dim() != 2 or g.A.stride(1) != 1 or g.B.stride(1) != 1:
                raise ValueError(f"gemm_grouped: problem {lo + i}: A and B must be 2-D {op_dt} tensors with a contiguous last dimension")
            odt = g.out_dtype or op_dt
            if g.C is None:
                g.C = torch.empty((g.M, g.N), dtype=odt, device=dev)
            elif g.C.dtype != odt or g.C.stride(-1) != 1:
                raise ValueError(f"gemm_grouped: problem {lo + i}: C must be {odt} with a contiguous last dimension")
            if want_col_stats and g.col_stats is None:
                g.col_stats = torch.empty(((g.M + 127) // 128, 2, g.N), dtype=torch.float32, device=dev)  # one entry per 128-row M tile
            if want_row_sumsq and g.row_sumsq is None:
                g.row_sumsq = torch.empty(((g.N + 63) // 64, g.M), dtype=torch.float32, device=dev)
            bias = None if g.bias is None else _contig(g.bias.detach().to(torch.float32))
            keep.append(bias)
            arr[i] = L.GemmProblem(L.ptr(g.A), g.A.stride(0), L.ptr(g.B), g.B.stride(0), L.ptr(g.C), g.C.stride(0), g.M, g.N, g.K,
                                   int(g.a_is_km), int(g.b_is_kn), L.dtype_code(odt), g.alpha, L.ptr(bias), L.ptr(g.col_stats), L.ptr(g.row_sumsq),
                                   L.ptr(g.a_scale), L.ptr(g.a_shift), int(g.a_relu), g.tile_n, g.split_k, int(g.no_tma_store), 0.0, 0, 0)
        ws_bytes = L.lib().msf_gemm_grouped_workspace_bytes(arr, len(chunk))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes > 256 else None
        L.check(L.lib().msf_gemm_grouped(arr, len(chunk), L.dtype_code(op_dt), L.ptr(ws), ws_bytes if ws is not None else 0,
                                         L.ptr(_counters_for(dev)) if ws is not None else 0, L.stream_ptr()), "msf_gemm_grouped")
        L.launch_count += 1
        if ws is not None:
            ws.record_stream(torch.cuda.current_stream(dev))
        outs += [g.C for g in chunk]
    return outs


def _gemm_grouped_f32(specs: Sequence[GemmSpec]) -> List[torch.Tensor]:
    """fp32 operands: the plain-FMA SIMT kernel (``msf_gemm_grouped_f32``) -- exact fp32 accumulation, no tensor cores."""
    dev = specs[0].A.device
    outs = []
    for lo in range(0, len(specs), L.MSF_GEMM_MAX_PROBLEMS):
        chunk = specs[lo:lo + L.MSF_GEMM_MAX_PROBLEMS]
        arr = (L.GemmProblem * len(chunk))()
        keep = []
        for i, g in enumerate(chunk):
            L.require_cuda(g.A, g.B)
            if g.A.dtype != torch.float32 or g.B.dtype != torch.float32 or g.A.dim() != 2 or g.B.dim() != 2 or g.A.stride(1) != 1 or g.B.stride(1) != 1:
                raise ValueError(f"gemm_grouped: problem {lo + i}: A and B must be 2-D float32 tensors with a contiguous last dimension")
            if g.a_scale is not None or g.col_stats is not None or g.row_sumsq is not None:
                raise ValueError("gemm_grouped: the fp32 path has no fused prologue / statistics")
            if g.C is None:
                g.C = torch.empty((g.M, g.N), dtype=torch.float32, device=dev)
            bias = None if g.bias is None else _contig(g.bias.detach().to(torch.float32))
            keep.append(bias)
            arr[i] = L.GemmProblem(L.ptr(g.A), g.A.stride(0), L.ptr(g.B), g.B.stride(0), L.ptr(g.C), g.C.stride(0), g.M, g.N, g.K,
                                   int(g.a_is_km), int(g.b_is_kn), L.MSF_F32, g.alpha, L.ptr(bias), 0, 0, 0, 0, 0, 0, 0, 0, 0.0, 0, 0)
        L.check(L.lib().msf_gemm_grouped_f32(arr, len(chunk), L.stream_ptr()), "msf_gemm_grouped_f32")
        L.launch_count += 1
        outs += [g.C for g in chunk]
    return outs


def column_sums(g: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of a (rows, C) matrix on the head kernels (fixed-order, deterministic): the bias gradient of a Linear."""
    g = _contig(g)
    rows, Cc = g.shape
    dev = g.device
    part = torch.empty(((rows + 255) // 256, 2, Cc), dtype=torch.float32, device=dev)
    out = torch.empty(Cc, dtype=torch.float32, device=dev)
    it = (L.HeadBwdItem * 1)(L.HeadBwdItem(L.ptr(g), 0, 0, L.ptr(part), 0, 0, 0, 0, 0, 0, rows, Cc, 0, 0))
    L.check(L.lib().msf_head_bn_bwd_reduce(it, 1, L.dtype_code(g.dtype), L.stream_ptr()), "msf_head_bn_bwd_reduce")
    fin = (L.HeadBwdFinItem * 1)(L.HeadBwdFinItem((C.c_void_p * 2)(L.ptr(part), 0), (C.c_void_p * 2)(0, 0), (C.c_void_p * 2)(0, 0), 0, L.ptr(out), rows, Cc, 1, 1))
    L.check(L.lib().msf_head_bn_bwd_finalize(fin, 1, 0, 0, 1, 0, 0, 0, 1, L.stream_ptr()), "msf_head_bn_bwd_finalize")
    L.launch_count += 2
    return out


class _LinearTC(torch.autograd.Function):
    """One Linear on the grouped GEMM kernels: tcgen05 for 16-bit operands, the plain-FMA SIMT kernel for fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_op):
        L.require_cuda(x, weight)
        dt = w_op.dtype
        xo = _contig(x if x.dtype == dt else x.to(dt))
        rows, fin = xo.shape
        fout = w_op.shape[0]
        (y,) = gemm_grouped([GemmSpec(xo, w_op, rows, fout, fin, bias=bias)])
        ctx.save_for_backward(xo, w_op)
        ctx.meta = (weight.dtype, None if bias is None else bias.dtype, x.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        xo, w_op = ctx.saved_tensors
        wdt, bdt, xdt = ctx.meta
        dt = w_op.dtype
        gyo = _contig(gy if gy.dtype == dt else gy.to(dt))
        rows, fin = xo.shape
        fout = w_op.shape[0]
        specs = []
        if ctx.needs_input_grad[0]:
            specs.append(GemmSpec(gyo, w_op, rows, fin, fout, b_is_kn=True))                                             # dX = dY W
        if ctx.needs_input_grad[1]:
            specs.append(GemmSpec(gyo, xo, fout, fin, rows, a_is_km=True, b_is_kn=True, out_dtype=torch.float32))        # dW = dY^T X
        outs = gemm_grouped(specs)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = outs.pop(0)
            gx = gx if gx.dtype == xdt else gx.to(xdt)
        if ctx.needs_input_grad[1]:
            gw = outs.pop(0).to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            gb = column_sums(gyo).to(bdt)
        return gx, gw, gb, None


def linear_tc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
              weight_op: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x W^T (+ b) with fp32 accumulation on this repo's GEMM kernels (forward, dX and dW).  ``weight_op`` = the
    operand copy of ``weight`` in the compute dtype (bf16 / fp16: tcgen05; fp32: SIMT); defaults to ``weight`` itself.
    x (rows, in), W (out, in); in and out must be multiples of 8 for the 16-bit path."""
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise ValueError(f"linear_tc: x {tuple(x.shape)} vs weight {tuple(weight.shape)}")
    w_op = weight.detach() if weight_op is None else weight_op
    if w_op.dtype != torch.float32 and (x.shape[1] % 8 or weight.shape[0] % 8):
        raise ValueError("linear_tc: in/out features must be multiples of 8")
    return _LinearTC.apply(x, weight, bias, w_op)


# ------------------------------------------------------------------------------------------
# A2: crop + bilinear resample            (integer case: src/models/hooknet.py:29-32)
# ------------------------------------------------------------------------------------------
class _CropResample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, boxes, oh: int, ow: int):
        feat = _contig(feat)
        boxes = _contig(boxes.to(torch.float32))
        L.require_cuda(feat, boxes)
        B, Cc, H, W = feat.shape
        if boxes.dim() != 3 or boxes.shape[0] != B or boxes.shape[2] != 4:
            raise ValueError(f"crop_resample: boxes must be (B,K,4); got {tuple(boxes.shape)}")
        K = boxes.shape[1]
        out = torch.empty((B, K, Cc, oh, ow), dtype=feat.dtype, device=feat.device)
        L.check(L.lib().msf_crop_resample_fwd(L.ptr(feat), B, Cc, H, W, L.ptr(boxes), K, oh, ow, L.dtype_code(feat.dtype), L.ptr(out),
                                              L.stream_ptr()), "msf_crop_resample_fwd")
        L.launch_count += 1
        ctx.save_for_backward(boxes)
        ctx.meta = (B, Cc, H, W, K, oh, ow, feat.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B, Cc, H, W, K, oh, ow, dt = ctx.meta
        (boxes,) = ctx.saved_tensors
        go = _contig(grad_out.to(dt))
        gfeat = torch.empty((B, Cc, H, W), dtype=torch.float32, device=go.device)  # every element is written (gather form)
        L.check(L.lib().msf_crop_resample_bwd(L.ptr(go), B, Cc, H, W, L.ptr(boxes), K, oh, ow, L.dtype_code(dt), L.ptr(gfeat),
                                              L.stream_ptr()), "msf_crop_resample_bwd")
        L.launch_count += 1
        return gfeat.to(dt), None, None, None


def crop_resample(feat: torch.Tensor, boxes: torch.Tensor, out_hw: Tuple[int, int]) -> torch.Tensor:
    """feat (B,C,H,W), boxes (B,K,4) [y0,x0,y1,x1) -> (B,K,C,oh,ow)."""
    return _CropResample.apply(feat, boxes, int(out_hw[0]), int(out_hw[1]))


def footprint_boxes(B: int, scale: int, H: int, W: int, device) -> torch.Tensor:
    """Footprints of the scale x scale high-magnification tiles on a low-magnification (H,W) map, in the raster
    order of ``blockshaped`` (src/utils/data/bcss.py:203-216): tile t -> rows [H/scale*(t//scale), +H/scale)."""
    th, tw = H / scale, W / scale
    t = torch.arange(scale * scale, device=device)
    y0 = (t // scale).to(torch.float32) * th
    x0 = (t % scale).to(torch.float32) * tw
    box = torch.stack((y0, x0, y0 + th, x0 + tw), dim=1)
    return box.unsqueeze(0).expand(B, -1, -1).contiguous()


# ------------------------------------------------------------------------------------------
# N1: channels-last BatchNorm2d (+ReLU, +residual, +stem max-pool) of the encoders   (src/models/resnet.py:59-82, 244-247)
# ------------------------------------------------------------------------------------------
def _nhwc(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def _sync_world(group) -> int:
    if group is None or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class PeerReducer:
    """Sum-all-reduce of small fp64 vectors over NVLink peer memory (``msf_peer_allreduce_f64``): one single-CTA kernel
    on the current stream per call instead of an NCCL launch.  The symmetric workspace (every rank's copy mapped into
    every process) comes from ``torch.distributed._symmetric_memory``; creating a reducer is a collective."""

    CAPACITY = 16384  # doubles per slot: 2 * 4608 + 1 (the widest head BatchNorm) fits with room to spare
    # how long a rank waits for its peers inside the kernel before trapping (first steps can be seconds apart while
    # cuDNN autotunes); MSFWSI_PEER_TIMEOUT_S overrides
    TIMEOUT_MS = int(float(__import__("os").environ.get("MSFWSI_PEER_TIMEOUT_S", "120")) * 1000)
    _cache = {}

    def __init__(self, group, device: torch.device):
        import torch.distributed._symmetric_memory as symm
        nbytes = L.lib().msf_peer_workspace_bytes(self.CAPACITY)
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.world, self.rank = int(self.hdl.world_size), int(self.hdl.rank)
        if self.world > L.MSF_PEER_MAX_WORLD:
            raise RuntimeError(f"PeerReducer supports up to {L.MSF_PEER_MAX_WORLD} ranks")
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every workspace is zeroed before anybody publishes a flag
        self.seq = 0

    def all_reduce_(self, vec: torch.Tensor) -> torch.Tensor:
        if vec.dtype != torch.float64 or not vec.is_contiguous() or vec.numel() > self.CAPACITY:
            raise ValueError("PeerReducer.all_reduce_: contiguous float64 vector of at most CAPACITY elements expected")
        self.seq += 1
        L.check(L.lib().msf_peer_allreduce_f64(L.ptr(vec), vec.numel(), L.ptr(self.peers), self.world, self.rank, self.seq, self.CAPACITY,
                                               self.TIMEOUT_MS, L.stream_ptr()), "msf_peer_allreduce_f64")
        L.launch_count += 1
        return vec

    @classmethod
    def get(cls, group, device: torch.device):
        """The reducer of (group, device), created on first use (collectively).  If symmetric memory cannot be set up
        on ANY rank, all ranks agree to use NCCL instead (returns None)."""
        key = (getattr(group, "group_name", None) or id(group), device.index)
        if key not in cls._cache:
            red, ok = None, 1
            try:
                red = cls(group, device)
            except Exception as e:  # noqa: BLE001 -- any failure means "no peer memory here"
                import warnings
                warnings.warn(f"msfwsi_b200: NVLink peer-memory all-reduce unavailable ({e!r:.200}); batch-norm statistics use NCCL", stacklevel=2)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            cls._cache[key] = red if int(flag.item()) == 1 else None
        return cls._cache[key]


USE_PEER_ALLREDUCE = True  # batch-norm statistics over NVLink peer memory when available (else NCCL)


def _all_reduce_stats(sums: torch.Tensor, group) -> None:
    red = PeerReducer.get(group, sums.device) if USE_PEER_ALLREDUCE else None
    if red is not None:
        red.all_reduce_(sums)
    else:
        dist.all_reduce(sums, group=group)


class _BNAct2d(torch.autograd.Function):
    """act(batch_norm(x) (+ residual)) with batch statistics (train mode), optionally followed by the stem's
    3x3/2 max-pool.  ``sync_group`` = process group the statistics are reduced over (SyncBatchNorm semantics), or None."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, eps: float, momentum: float, relu: bool, pool: bool,
                sync_group, want_mean: bool = False):
        L.require_cuda(x, weight, bias, residual)
        flat = x.dim() == 2  # (rows, C): BatchNorm1d of the heads -- the kernels see a [rows][C] matrix either way
        if flat:
            if pool:
                raise ValueError("bn_act2d: the pooled variant needs a 4-D input")
            x = _contig(x)[:, :, None, None]
            residual = None if residual is None else _contig(residual)[:, :, None, None]
        if x.dim() != 4:
            raise ValueError(f"bn_act2d: expected (N,C,H,W) or (rows,C), got {tuple(x.shape)}")
        if pool and (residual is not None or not relu):
            raise ValueError("bn_act2d: the pooled variant is bn -> relu -> maxpool without a residual")
        if want_mean and (pool or residual is None or not relu):
            raise ValueError("bn_act2d: the spatial mean output is implemented for relu(bn(x) + residual) (a BasicBlock output)")
        if not flat:
            x = _nhwc(x)
        N, Cc, H, W = x.shape
        dt, dev = x.dtype, x.device
        code = L.dtype_code(dt)
        rows = N * H * W
        gamma = None if weight is None else _contig(weight.detach().float())
        beta = None if bias is None else _contig(bias.detach().float())
        lib, st = L.lib(), L.stream_ptr()
        ws_bytes = lib.msf_bn2d_workspace_bytes(rows, Cc)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sums = torch.empty(2 * Cc + 1, dtype=torch.float64, device=dev)
        mean = torch.empty(Cc, dtype=torch.float32, device=dev)
        invstd = torch.empty(Cc, dtype=torch.float32, device=dev)
        world = _sync_world(sync_group)
        if world > 1:
            L.check(lib.msf_bn2d_stats(L.ptr(x), rows, Cc, code, L.ptr(sums), L.ptr(ws), ws_bytes, st), "msf_bn2d_stats")
            _all_reduce_stats(sums, sync_group)  # one sum-reducible fp64 vector {sum, sum of squares, count}
            L.check(lib.msf_bn2d_finalize(L.ptr(sums), Cc, eps, momentum, L.ptr(mean), L.ptr(invstd), L.ptr(running_mean),
                                          L.ptr(running_var), st), "msf_bn2d_finalize")
        else:
            L.check(lib.msf_bn2d_stats_finalize(L.ptr(x), rows, Cc, code, eps, momentum, L.ptr(sums), L.ptr(mean), L.ptr(invstd),
                                                L.ptr(running_mean), L.ptr(running_var), L.ptr(ws), ws_bytes, st), "msf_bn2d_stats_finalize")
        tap = x_arg = bits = None
        if pool:
            PH, PW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            y = torch.empty((N, Cc, PH, PW), dtype=dt, device=dev, memory_format=torch.channels_last)
            tap = torch.empty((N, PH, PW, Cc), dtype=torch.uint8, device=dev)
            x_arg = torch.empty((N, PH, PW, Cc), dtype=dt, device=dev)
            L.check(lib.msf_bn2d_apply_pool(L.ptr(x), L.ptr(y), L.ptr(tap), L.ptr(x_arg), N, H, W, Cc, code, L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                                            L.ptr(beta), st), "msf_bn2d_apply_pool")
        else:
            res = None
            if residual is not None:
                if residual.shape != x.shape or residual.dtype != dt:
                    raise ValueError(f"bn_act2d: residual {tuple(residual.shape)} {residual.dtype} vs x {tuple(x.shape)} {dt}")
                res = residual if flat else _nhwc(residual)
            y = torch.empty_like(x) if flat else torch.empty_like(x, memory_format=torch.channels_last)
            # with a residual the ReLU mask cannot be recomputed from x alone: one bit per element is written here
            # (16x less backward traffic than re-reading the bf16 output in both backward passes)
            if residual is not None and relu:
                bits = torch.empty(rows * Cc // (16 // x.element_size()), dtype=torch.uint8, device=dev)
            L.check(lib.msf_bn2d_apply(L.ptr(x), L.ptr(res), L.ptr(y), L.ptr(bits), rows, Cc, code, L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                                       L.ptr(beta), int(relu), st), "msf_bn2d_apply")
        L.launch_count += 4
        y_mask = y if pool else bits  # stem: the pooled output is the mask of the pooled-grid reduction
        ctx.save_for_backward(x, gamma, beta, mean, invstd, sums, y_mask, tap, x_arg)
        ctx.meta = (N, Cc, H, W, code, relu, pool, residual is not None, sync_group, world,
                    None if weight is None else weight.dtype, None if bias is None else bias.dtype, want_mean, flat)
        if flat:
            return y[:, :, 0, 0]
        if want_mean:
            # global average pool of the block output (src/models/resnet.py:250-254); its backward is folded into the
            # batch-norm backward kernels below instead of being expanded and added to the main gradient by ATen
            return y, y.mean(dim=(2, 3))
        return y

    @staticmethod
    def backward(ctx, gy, gmean=None):
        x, gamma, beta, mean, invstd, sums_fwd, y_mask, tap, x_arg = ctx.saved_tensors
        N, Cc, H, W, code, relu, pool, has_res, group, world, wdt, bdt, want_mean, flat = ctx.meta
        dev = x.device
        if gy is None:  # only the pooled branch carries gradient
            gy = torch.zeros_like(x, memory_format=torch.channels_last)
        gy = gy if gy.dtype == x.dtype else gy.to(x.dtype)
        gy = _contig(gy)[:, :, None, None] if flat else _nhwc(gy)
        like = (lambda t: torch.empty_like(t)) if flat else (lambda t: torch.empty_like(t, memory_format=torch.channels_last))
        gp = None
        if want_mean and gmean is not None:
            gp = _contig(gmean if gmean.dtype == x.dtype else gmean.to(x.dtype))
        hw = H * W
        lib, st = L.lib(), L.stream_ptr()
        rows = N * H * W
        sums = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
        count_ptr = sums_fwd.data_ptr() + 16 * Cc  # element 2C of the forward sums: the (global) element count
        dx = like(x)
        dres = None
        if pool:
            PH, PW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            ws_bytes = lib.msf_bn2d_workspace_bytes(N * PH * PW, Cc)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            # dy' lives on the pooled grid: reduce over (x at the arg-max, pooled gradient, pooled output as ReLU mask)
            L.check(lib.msf_bn2d_bwd_reduce(L.ptr(x_arg), L.ptr(gy), L.ptr(y_mask), 0, N * PH * PW, Cc, code, L.ptr(mean), L.ptr(invstd),
                                            L.ptr(gamma), L.ptr(beta), 1, None, 0, L.ptr(sums), L.ptr(ws), ws_bytes, st), "msf_bn2d_bwd_reduce")
        else:
            ws_bytes = lib.msf_bn2d_workspace_bytes(rows, Cc)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            L.check(lib.msf_bn2d_bwd_reduce(L.ptr(x), L.ptr(gy), L.ptr(y_mask), int(y_mask is not None), rows, Cc, code, L.ptr(mean), L.ptr(invstd), L.ptr(gamma),
                                            L.ptr(beta), int(relu), L.ptr(gp), hw, L.ptr(sums), L.ptr(ws), ws_bytes, st), "msf_bn2d_bwd_reduce")
        # parameter gradients are the LOCAL sums (DDP averages them over ranks, as with SyncBatchNorm)
        gw = None if wdt is None else sums[Cc:].to(wdt)
        gb = None if bdt is None else sums[:Cc].to(bdt)
        if world > 1:
            _all_reduce_stats(sums, group)
        if pool:
            L.check(lib.msf_bn2d_pool_bwd_elemt(L.ptr(x), L.ptr(gy), L.ptr(tap), L.ptr(dx), N, H, W, Cc, code, L.ptr(mean), L.ptr(invstd),
                                                L.ptr(gamma), L.ptr(sums), count_ptr, st), "msf_bn2d_pool_bwd_elemt")
        else:
            if has_res and ctx.needs_input_grad[3]:
                dres = like(x)
            L.check(lib.msf_bn2d_bwd_elemt(L.ptr(x), L.ptr(gy), L.ptr(y_mask), int(y_mask is not None), L.ptr(dx), L.ptr(dres), rows, Cc, code, L.ptr(mean),
                                           L.ptr(invstd), L.ptr(gamma), L.ptr(beta), int(relu), L.ptr(gp), hw, L.ptr(sums), count_ptr, st),
                    "msf_bn2d_bwd_elemt")
        L.launch_count += 3
        if flat:
            dx = dx[:, :, 0, 0]
            dres = None if dres is None else dres[:, :, 0, 0]
        return dx, gw, gb, dres, None, None, None, None, None, None, None, None


def bn_act2d(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], running_mean: Optional[torch.Tensor],
             running_var: Optional[torch.Tensor], eps: float = 1e-5, momentum: float = 0.1, relu: bool = False,
             residual: Optional[torch.Tensor] = None, pool: bool = False, sync_group=None, want_mean: bool = False):
    """Train-mode ``[maxpool3x3/2](relu?(batch_norm(x) (+ residual)))`` on channels-last CUDA tensors.  Statistics are
    biased batch statistics (all-reduced over ``sync_group`` when given); ``running_*`` receive the momentum update with
    the unbiased variance.  x (N,C,H,W) in NHWC memory order (converted if not).  ``want_mean=True`` additionally
    returns the spatial mean (N,C) of the output (the encoder's pooled pyramid feature), whose gradient is folded into
    the batch-norm backward kernels."""
    return _BNAct2d.apply(x, weight, bias, residual, running_mean, running_var, float(eps), float(momentum), bool(relu), bool(pool),
                          sync_group, bool(want_mean))


def bn_eval_apply(x: torch.Tensor, weight, bias, running_mean: torch.Tensor, running_var: torch.Tensor, eps: float = 1e-5, relu: bool = False):
    """Eval-mode batch norm of a (rows, C) matrix with the running statistics (no gradient): ``msf_head_bn_finalize`` in
    eval mode turns them into scale / shift, ``msf_head_bn_apply`` applies them."""
    x = _contig(x)
    L.require_cuda(x, running_mean, running_var)
    rows, Cc = x.shape
    dev = x.device
    sc, sh = torch.empty(Cc, dtype=torch.float32, device=dev), torch.empty(Cc, dtype=torch.float32, device=dev)
    P2 = C.c_void_p * 2
    gam = None if weight is None else _contig(weight.detach().float())
    bet = None if bias is None else _contig(bias.detach().float())
    it = (L.HeadBnItem * 1)(L.HeadBnItem(P2(0, 0), P2(L.ptr(sc), 0), P2(L.ptr(sh), 0), P2(0, 0), P2(0, 0), L.ptr(gam), L.ptr(bet),
                                         L.ptr(running_mean), L.ptr(running_var), rows, Cc, 1, 0))
    L.check(L.lib().msf_head_bn_finalize(it, 1, float(eps), 0.0, 0, 0, 1, 0, 0, 0, 1, L.stream_ptr()), "msf_head_bn_finalize")
    y = torch.empty_like(x)
    ap = (L.HeadApplyItem * 1)(L.HeadApplyItem(L.ptr(x), L.ptr(y), 0, 0, L.ptr(sc), L.ptr(sh), 0, rows, Cc, int(relu), 0))
    L.check(L.lib().msf_head_bn_apply(ap, 1, L.dtype_code(x.dtype), COS_EPS, L.stream_ptr()), "msf_head_bn_apply")
    L.launch_count += 2
    return y


# ------------------------------------------------------------------------------------------
# S1: space-to-depth input layout for the stem convolution          (src/models/resnet.py:155, 244)
# ------------------------------------------------------------------------------------------
def stem_s2d(x: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """x (N, C<=4, H, W) (any strides, H and W even) -> the zero-padded (3 px), 2x2 pixel-unshuffled image as an
    (N, 16, (H+6)/2, (W+6)/2) channels-last tensor in ``out_dtype``: channel c*4 + dy*2 + dx of pixel (oy, ox) is
    x[n, c, 2*oy+dy-3, 2*ox+dx-3].  No gradient flows to x (the encoder input never needs one)."""
    L.require_cuda(x)
    if x.dim() != 4:
        raise ValueError(f"stem_s2d: expected (N,C,H,W), got {tuple(x.shape)}")
    N, Cin, H, W = x.shape
    out_dtype = out_dtype or x.dtype
    x = x.detach()
    out = torch.empty((N, 16, (H + 6) // 2, (W + 6) // 2), dtype=out_dtype, device=x.device, memory_format=torch.channels_last)
    sn, sc, sy, sx = x.stride()
    L.check(L.lib().msf_stem_s2d(L.ptr(x), N, Cin, H, W, sn, sc, sy, sx, L.dtype_code(x.dtype), L.ptr(out), L.dtype_code(out_dtype),
                                 L.stream_ptr()), "msf_stem_s2d")
    L.launch_count += 1
    return out


def stem_s2d_weight(w: torch.Tensor) -> torch.Tensor:
    """The (O, C, 7, 7) stem kernel in the layout that pairs with :func:`stem_s2d`: zero-padded to 8x8, split into
    (ky, dy) x (kx, dx) and rearranged to (O, 16, 4, 4) with input channel c*4 + dy*2 + dx.  Plain differentiable torch
    ops on a 9408-element tensor, so autograd turns the 4x4 kernel's gradient back into the 7x7 one."""
    O, Cin, kh, kw = w.shape
    if (kh, kw) != (7, 7) or Cin > 4:
        raise ValueError(f"stem_s2d_weight: expected (O, C<=4, 7, 7), got {tuple(w.shape)}")
    w8 = torch.nn.functional.pad(w, (0, 1, 0, 1))                       # (O, C, 8, 8), tap 7 = 0
    w8 = w8.reshape(O, Cin, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4)       # (O, C, dy, dx, ky, kx)
    w16 = w8.reshape(O, Cin * 4, 4, 4)
    return torch.nn.functional.pad(w16, (0, 0, 0, 0, 0, 16 - Cin * 4))  # channels -> 16


# ------------------------------------------------------------------------------------------
# D1: on-device data path of the views      (src/utils/data/bcss.py:171-177, 203-216 + transforms[2])
# ------------------------------------------------------------------------------------------
def jigsaw_tiles(src: torch.Tensor, perm: Optional[torch.Tensor], grid: int = 4, out_hw: Tuple[int, int] = (224, 224),
                 mean: Sequence[float] = (0.485, 0.456, 0.406), std: Sequence[float] = (0.229, 0.224, 0.225),
                 out_dtype: torch.dtype = torch.bfloat16, validate: bool = False) -> torch.Tensor:
    """src (B, H, W, 3) uint8 HWC on the device; perm (B, grid*grid) int64 = jigsaw_idx (forward shuffle) or None.
    Returns the (B*grid*grid, 3, oh, ow) channels-last views: tile j of sample b is source tile perm[b, j] of
    ``blockshaped``, resized and normalised.  ``grid=1`` gives the context view of the whole image."""
    L.require_cuda(src, perm)
    if src.dim() != 4 or src.shape[3] != 3 or src.dtype != torch.uint8:
        raise ValueError(f"jigsaw_tiles: src must be (B,H,W,3) uint8, got {tuple(src.shape)} {src.dtype}")
    src = _contig(src)
    B, H, W, _ = src.shape
    K = grid * grid
    if perm is not None:
        if tuple(perm.shape) != (B, K) or perm.dtype != torch.int64:
            raise AssertionError(f"jigsaw_idx must be int64 of shape ({B},{K}); got {perm.dtype} {tuple(perm.shape)}")
        perm = _contig(perm)
    if H % grid or W % grid:
        raise AssertionError(f"{H} x {W} is not evenly divisible into a {grid} x {grid} grid")  # bcss.py:212-213
    oh, ow = int(out_hw[0]), int(out_hw[1])
    out = torch.empty((B * K, 3, oh, ow), dtype=out_dtype, device=src.device, memory_format=torch.channels_last)
    flag = torch.zeros(1, dtype=torch.int32, device=src.device) if validate else None
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    L.check(L.lib().msf_jigsaw_tiles(L.ptr(src), B, H, W, grid, L.ptr(perm), oh, ow, m3, s3, L.ptr(out), L.dtype_code(out_dtype), L.ptr(flag),
                                     L.stream_ptr()), "msf_jigsaw_tiles")
    L.launch_count += 1
    if validate and int(flag.item()) != 0:
        raise IndexError(f"jigsaw_idx holds values outside [-{K}, {K})")
    return out


def view_crops_s2d(src: torch.Tensor, crops: torch.Tensor, out_hw: Tuple[int, int] = (224, 224), mean: Sequence[float] = (0.485, 0.456, 0.406),
                   std: Sequence[float] = (0.229, 0.224, 0.225), out_dtype: torch.dtype = torch.bfloat16, validate: bool = False) -> torch.Tensor:
    """Every view of a step from the uint8 source tiles in one launch, in the stem convolution's input layout.
    src (B, H, W, 3) uint8 on the device; crops (n, 6) int32 rows ``[sample, y0, x0, y1, x1, flip]`` (integer source-pixel
    boxes, what albumentations' RandomResizedCrop draws; see :func:`jigsaw_view_crops` for the target views).  Returns the
    (n, 16, (oh+6)/2, (ow+6)/2) channels-last tensor ``stem_s2d`` would produce from the normalised (n, 3, oh, ow) views;
    the encoders take it as is (``ResNet._stem_conv``)."""
    L.require_cuda(src, crops)
    if src.dim() != 4 or src.shape[3] != 3 or src.dtype != torch.uint8:
        raise ValueError(f"view_crops_s2d: src must be (B,H,W,3) uint8, got {tuple(src.shape)} {src.dtype}")
    if crops.dim() != 2 or crops.shape[1] != 6 or crops.dtype != torch.int32:
        raise ValueError(f"view_crops_s2d: crops must be (n, 6) int32, got {tuple(crops.shape)} {crops.dtype}")
    src, crops = _contig(src), _contig(crops)
    B, H, W, _ = src.shape
    n = crops.shape[0]
    oh, ow = int(out_hw[0]), int(out_hw[1])
    out = torch.empty((n, 16, (oh + 6) // 2, (ow + 6) // 2), dtype=out_dtype, device=src.device, memory_format=torch.channels_last)
    flag = torch.zeros(1, dtype=torch.int32, device=src.device) if validate else None
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    L.check(L.lib().msf_view_crops_s2d(L.ptr(src), B, H, W, L.ptr(crops), n, oh, ow, m3, s3, L.ptr(out), L.dtype_code(out_dtype), L.ptr(flag),
                                       L.stream_ptr()), "msf_view_crops_s2d")
    L.launch_count += 1
    if validate and int(flag.item()) != 0:
        raise IndexError("view_crops_s2d: a crop box or sample index lies outside the source images")
    return out


def jigsaw_view_crops(perm: torch.Tensor, boxes: torch.Tensor, flips: torch.Tensor, H: int, W: int, grid: int = 4) -> torch.Tensor:
    """Crop rows for the target views: view (b, j) is tile ``perm[b, j]`` of ``blockshaped(img_b, H/grid, W/grid)``
    (src/utils/data/bcss.py:171-177; raster tile t covers rows [th*(t//grid), +th), cols [tw*(t%grid), +tw)) cropped to
    ``boxes[b, j] = [y0, x0, y1, x1)`` INSIDE the tile.  perm (B, K) int64, boxes (B, K, 4) integer, flips (B, K) -> (B*K, 6)
    int32 for :func:`view_crops_s2d` (integer arithmetic only: bit-exact tiling)."""
    B, K = perm.shape
    th, tw = H // grid, W // grid
    t = perm.remainder(K)
    oy, ox = (t // grid) * th, (t % grid) * tw
    b = boxes.to(torch.int64)
    sample = torch.arange(B, device=perm.device).view(B, 1).expand(B, K)
    rows = torch.stack((sample, oy + b[..., 0], ox + b[..., 1], oy + b[..., 2], ox + b[..., 3], flips.to(torch.int64)), dim=2)
    return rows.reshape(B * K, 6).to(torch.int32)


# ------------------------------------------------------------------------------------------
# E1: multi-tensor EMA (extension)
# ------------------------------------------------------------------------------------------
class EmaUpdater:
    """``teacher <- m * teacher + (1 - m) * student`` over a fixed parameter list, one launch per step.
    The device-side tensor table is built once here."""

    def __init__(self, teacher: Sequence[torch.Tensor], student: Sequence[torch.Tensor]):
        teacher, student = list(teacher), list(student)
        if len(teacher) != len(student) or not teacher:
            raise ValueError("EmaUpdater: need equally long, non-empty tensor lists")
        L.require_cuda(*teacher, *student)
        self.tdt, self.sdt = teacher[0].dtype, student[0].dtype
        n = len(teacher)
        numels = (C.c_int64 * n)()
        for i, (t, s) in enumerate(zip(teacher, student)):
            if t.shape != s.shape or t.dtype != self.tdt or s.dtype != self.sdt:
                raise ValueError(f"EmaUpdater: tensor {i} shape/dtype mismatch")
            dense = t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))
            if not dense or t.stride() != s.stride():
                raise ValueError(f"EmaUpdater: tensor {i} must be dense (contiguous or channels-last) with equal strides on both sides")
            numels[i] = t.numel()
        prefix = (C.c_int32 * (n + 1))()
        L.check(L.lib().msf_ema_plan(numels, n, prefix), "msf_ema_plan")
        self.n, self.total_chunks = n, int(prefix[n])
        table = torch.empty((n, 3), dtype=torch.int64)
        for i, (t, s) in enumerate(zip(teacher, student)):
            table[i, 0], table[i, 1], table[i, 2] = t.data_ptr(), s.data_ptr(), t.numel()
        dev = teacher[0].device
        self._table = table.to(dev)
        self._prefix = torch.tensor(list(prefix), dtype=torch.int32, device=dev)
        self._keep = (teacher, student)  # the table holds raw pointers: keep the tensors alive
        self.numel = int(sum(numels))

    def step(self, momentum: float) -> None:
        L.check(L.lib().msf_ema_multi(L.ptr(self._table), L.ptr(self._prefix), self.n, self.total_chunks, L.dtype_code(self.tdt),
                                      L.dtype_code(self.sdt), float(momentum), float(1.0 - float(momentum)), L.stream_ptr()), "msf_ema_multi")
        L.launch_count += 1
        L.bump_param_epoch()  # teachers changed through raw pointers: derived 16-bit copies must refresh
