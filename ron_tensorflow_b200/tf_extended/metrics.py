"""Drop-in for the hot-path consumers in the reference's ``tf_extended/metrics.py``:
``streaming_tp_fp_arrays`` :133-206, ``precision_recall`` :100-130, ``average_precision_voc07``
:237-258, ``average_precision_voc12`` :212-234.

The reference accumulates TP/FP records in TF local variables inside one process.  Here they are
accumulated per rank -- on the device (``TpFpDeviceState``: nothing is read back per batch) or on the host
(``streaming_tp_fp_arrays``, the reference's function) -- and ``gather_tp_fp`` / ``gather_detections`` are the
collectives of the whole path: every rank contributes its (score, tp, fp) records, ground-truth counts and,
when asked, its detections; NCCL (torch.distributed) all-gathers them once at the end of the evaluation and AP
is computed exactly as the reference does.
"""
import numpy as np
import torch

__all__ = ['streaming_tp_fp_arrays', 'precision_recall', 'average_precision_voc07', 'average_precision_voc12',
           'TpFpAccumulator', 'TpFpDeviceState', 'gather_tp_fp', 'gather_detections', 'average_precision_records',
           'gather_tp_fp_records', 'PAD_RECORD']

# a record whose class index is 0xffffff is padding: ronk_average_precision_records sorts it behind every class and ignores it
PAD_RECORD = 0xffffff << 8


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class TpFpAccumulator(object):
    """The five local variables of reference :177-181 for one class."""

    def __init__(self):
        self.n_gt = 0
        self.scores = np.zeros((0,), np.float32)
        self.tp = np.zeros((0,), bool)
        self.fp = np.zeros((0,), bool)

    def update(self, num_gbboxes, tp, fp, scores, remove_zero_scores=True):
        s = _np(scores).astype(np.float32).reshape(-1)
        t = _np(tp).astype(bool).reshape(-1)
        f = _np(fp).astype(bool).reshape(-1)
        mask = t | f                                              # reference :167
        if remove_zero_scores:
            mask = mask & (s > np.float32(1e-4))                  # reference :169-171
            s, t, f = s[mask], t[mask], f[mask]                   # reference :172-174
        self.n_gt += int(_np(num_gbboxes).astype(np.int64).sum())
        self.scores = np.concatenate([self.scores, s])
        self.tp = np.concatenate([self.tp, t])
        self.fp = np.concatenate([self.fp, f])
        return self.value()

    def value(self):
        return self.n_gt, self.scores.shape[0], self.tp, self.fp, self.scores


class TpFpDeviceState(object):
    """The accumulators of ``streaming_tp_fp_arrays`` for all classes, resident on the device: every batch appends its
    filtered (score, tp, fp) records behind the earlier ones with two small kernels (ronk_tpfp_records_append) and
    nothing is read back until ``to_host()`` / ``gather_tp_fp`` at the end of the evaluation -- one copy instead of
    4 x (C - 1) device->host reads per batch.  ``capacity``: records the buffer can hold (an overflow raises at the end)."""

    def __init__(self, num_classes, capacity=1 << 22, device=None, max_updates=4096):
        from .. import core
        core._require_cuda()
        self.num_classes = int(num_classes)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.capacity = int(capacity)
        self.records = torch.empty((self.capacity,), dtype=torch.int64, device=self.device)
        self.totals = torch.zeros((4,), dtype=torch.int32, device=self.device)
        self.n_gt = torch.zeros((self.num_classes - 1,), dtype=torch.int64, device=self.device)
        self.seg = torch.zeros((int(max_updates), self.num_classes - 1), dtype=torch.int32, device=self.device)
        self.calls = 0
        self._ws = None

    def update(self, num_gbboxes, tp, fp, scores, min_score=1e-4):
        """num_gbboxes int64 [B,C-1], tp / fp bool or uint8 [B,C-1,M], scores float32 [B,C-1,M] (the outputs of
        core.tpfp_match and of the NMS), all on this device."""
        import ctypes
        from .. import _ffi, core
        s = core.as_cuda(scores, torch.float32, self.device)
        t = tp.view(torch.uint8) if tp.dtype == torch.bool else core.as_cuda(tp, torch.uint8, self.device)
        f = fp.view(torch.uint8) if fp.dtype == torch.bool else core.as_cuda(fp, torch.uint8, self.device)
        g = core.as_cuda(num_gbboxes, torch.int64, self.device)
        B, CM, M = (int(v) for v in s.shape)
        if CM != self.num_classes - 1 or t.shape != s.shape or f.shape != s.shape or tuple(g.shape) != (B, CM):
            raise ValueError('expected scores / tp / fp [B,%d,M] and num_gbboxes [B,%d]' % (self.num_classes - 1, self.num_classes - 1))
        L = _ffi.lib()
        need = int(L.ronk_tpfp_records_workspace_bytes(B, CM + 1, M))
        if self._ws is None or self._ws.numel() * 4 < need:
            self._ws = torch.empty(((need + 3) // 4,), dtype=torch.int32, device=self.device)
        if self.calls >= self.seg.shape[0]:
            raise RuntimeError('TpFpDeviceState: more than %d updates, raise `max_updates`' % self.seg.shape[0])
        p = lambda x: ctypes.c_void_p(x.data_ptr())
        with torch.cuda.device(self.device):
            _ffi.check(L.ronk_tpfp_records_append(p(s.contiguous()), p(t.contiguous()), p(f.contiguous()), p(g.contiguous()), B, CM + 1, M,
                                                  float(min_score), p(self.records), self.capacity, p(self.totals), self.calls & 1,
                                                  p(self.n_gt), p(self.seg[self.calls]), p(self._ws),
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.calls += 1
        return self

    def count_tensor(self):
        """int32 device scalar tensor: records so far."""
        return self.totals[self.calls & 1]

    @staticmethod
    def _assemble(parts, n_gt, num_classes):
        """parts: per rank (records int64 [n] in arrival order, seg int32 [updates, C-1]) on the host.  Every update
        appended its records class after class, so class c of update u is one contiguous run: the per-class arrays
        are concatenations of runs (rank-major, then update order) -- no sorting.  -> dict class -> TpFpAccumulator."""
        CM = num_classes - 1
        runs = [[] for _ in range(CM)]
        for rec, seg in parts:
            w = np.ascontiguousarray(rec).view(np.uint32).reshape(-1, 2)          # little endian: (meta, score bits)
            ends = np.cumsum(seg.reshape(-1).astype(np.int64))
            starts = ends - seg.reshape(-1)
            for u in range(seg.shape[0]):
                for c in range(CM):
                    k = u * CM + c
                    if ends[k] > starts[k]:
                        runs[c].append(w[starts[k]:ends[k]])
        out = {}
        for c in range(1, num_classes):
            blk = np.concatenate(runs[c - 1]) if runs[c - 1] else np.zeros((0, 2), np.uint32)
            acc = TpFpAccumulator()
            acc.n_gt = int(n_gt[c - 1])
            acc.scores = blk[:, 1].copy().view(np.float32)
            acc.tp = (blk[:, 0] & 1).astype(bool)
            acc.fp = ((blk[:, 0] >> 1) & 1).astype(bool)
            out[c] = acc
        return out

    def average_precision(self, group=None, curves=False):
        """AP of every class without moving the records to the host: the records of all ranks are all-gathered on the
        device (``gather_tp_fp_records``; a no-op without a process group) and ``average_precision_records`` sorts and
        integrates them there.  Returns ``(ap07, ap12)``: dicts class -> float, read back in one copy of 2 x (C - 1)
        doubles (with ``curves=True`` a third value: dict class -> (precision, recall) float64 device tensors)."""
        rec, n_gt = gather_tp_fp_records(self, group)
        r = average_precision_records(rec, n_gt, self.num_classes, curves=curves)
        both = torch.stack([r['ap07'], r['ap12']]).cpu().numpy()
        ap07 = {c: float(both[0, c - 1]) for c in range(1, self.num_classes)}
        ap12 = {c: float(both[1, c - 1]) for c in range(1, self.num_classes)}
        if not curves:
            return ap07, ap12
        off = r['offsets'].cpu().numpy()
        cur = {c: (r['precision'][off[c - 1]:off[c]], r['recall'][off[c - 1]:off[c]]) for c in range(1, self.num_classes)}
        return ap07, ap12, cur

    def to_host(self):
        """dict class -> TpFpAccumulator (the state ``streaming_tp_fp_arrays`` would have built), one device->host copy."""
        tot = self.totals.cpu().numpy()
        if tot[2]:
            raise RuntimeError('TpFpDeviceState: more than %d records, raise `capacity`' % self.capacity)
        n = int(tot[self.calls & 1])
        return self._assemble([(self.records[:n].cpu().numpy(), self.seg[:self.calls].cpu().numpy())], self.n_gt.cpu().numpy(),
                              self.num_classes)


def streaming_tp_fp_arrays(num_gbboxes, tp, fp, scores, remove_zero_scores=True, metrics_collections=None,
                           updates_collections=None, name=None, state=None):
    """reference :133-206.  Returns (value, state): ``value`` is the reference's tuple
    (n_objects, n_detections, tp, fp, scores) after this update (dict of tuples for dict inputs);
    pass ``state`` back in on the next batch to keep accumulating."""
    if isinstance(scores, dict) or isinstance(fp, dict):
        state = state if state is not None else {}
        vals = {}
        for c in num_gbboxes.keys():
            acc = state.setdefault(c, TpFpAccumulator())
            vals[c] = acc.update(num_gbboxes[c], tp[c], fp[c], scores[c], remove_zero_scores)
        return vals, state
    acc = state if state is not None else TpFpAccumulator()
    return acc.update(num_gbboxes, tp, fp, scores, remove_zero_scores), acc


def precision_recall(num_gbboxes, num_detections, tp, fp, scores, dtype=np.float64, scope=None):
    """reference :100-130: sort by decreasing score (lower index first on ties), cumulative
    sums in float64, recall = tp / n_gt, precision = tp / (tp + fp), both 0 for a zero denominator."""
    if isinstance(scores, dict):
        d_p, d_r = {}, {}
        for c in num_gbboxes.keys():
            d_p[c], d_r[c] = precision_recall(num_gbboxes[c], num_detections[c], tp[c], fp[c], scores[c], dtype)
        return d_p, d_r
    s = _np(scores).astype(np.float32).reshape(-1)
    k = int(num_detections)
    order = np.argsort(-s.astype(np.float64), kind='stable')[:k]
    t = np.cumsum(_np(tp).astype(bool).reshape(-1)[order].astype(dtype))
    f = np.cumsum(_np(fp).astype(bool).reshape(-1)[order].astype(dtype))
    ng = dtype(int(_np(num_gbboxes)))
    recall = t / ng if ng > 0 else np.zeros_like(t)
    den = t + f
    precision = np.where(den > 0, t / np.where(den > 0, den, 1), 0.)
    return precision, recall


def average_precision_voc12(precision, recall, name=None):
    """reference :212-234."""
    p = np.concatenate([[0.], _np(precision).astype(np.float64), [0.]])
    r = np.concatenate([[0.], _np(recall).astype(np.float64), [1.]])
    p = np.maximum.accumulate(p[::-1])[::-1]
    return float(np.sum(p[1:] * (r[1:] - r[:-1])))


def average_precision_voc07(precision, recall, name=None):
    """reference :237-258: 11-point interpolation."""
    p = np.concatenate([_np(precision).astype(np.float64), [0.]])
    r = np.concatenate([_np(recall).astype(np.float64), [np.inf]])
    ap = 0.
    for t in np.arange(0., 1.1, 0.1):
        ap = ap + np.max(p[r >= t]) / 11.
    return float(ap)


def average_precision_records(records, n_gt, num_classes, thresholds=None, curves=False):
    """precision_recall + average_precision_voc07 / _voc12 (reference :100-130, :237-258, :212-234) of every class in one
    call, on the device: ``records`` int64 [n] = the packed (score bits << 32 | class index << 8 | fp << 1 | tp) records
    of ``TpFpDeviceState`` in concatenation order (``PAD_RECORD`` entries are ignored), ``n_gt`` int64 [C-1].  A stable
    radix sort by (class, descending score) gives tf.nn.top_k's order (lower index first among equal scores); sums and
    quotients are the reference's float64 ones.  -> dict(ap07, ap12: float64 [C-1] device tensors, offsets int32 [C];
    with ``curves``: sorted int64 [n], precision / recall float64 [n] in sorted order, class c at offsets[c-1]:offsets[c])."""
    import ctypes
    from .. import _ffi, core
    core._require_cuda()
    rec = records if (isinstance(records, torch.Tensor) and records.is_cuda) else core.as_cuda(records, torch.int64)
    rec = rec.contiguous().view(-1)
    if rec.dtype != torch.int64:
        raise ValueError('records must be int64 (packed uint64 bit patterns)')
    dev = rec.device
    g = core.as_cuda(n_gt, torch.int64, dev).contiguous()
    C = int(num_classes)
    if g.numel() != C - 1:
        raise ValueError('n_gt must have num_classes - 1 entries')
    n = int(rec.numel())
    thr = np.arange(0., 1.1, 0.1) if thresholds is None else np.asarray(thresholds, np.float64).reshape(-1)
    thr_c = (ctypes.c_double * len(thr))(*[float(t) for t in thr])
    L = _ffi.lib()
    need = int(L.ronk_average_precision_workspace_bytes(n, C))
    ws = torch.empty(((need + 7) // 8,), dtype=torch.int64, device=dev)
    out = dict(ap07=torch.empty((C - 1,), dtype=torch.float64, device=dev), ap12=torch.empty((C - 1,), dtype=torch.float64, device=dev),
               offsets=torch.empty((C,), dtype=torch.int32, device=dev))
    if curves:
        out['sorted'] = torch.empty((n,), dtype=torch.int64, device=dev)
        out['precision'] = torch.zeros((n,), dtype=torch.float64, device=dev)     # (padding entries stay 0)
        out['recall'] = torch.zeros((n,), dtype=torch.float64, device=dev)
    p = lambda k: ctypes.c_void_p(out[k].data_ptr()) if k in out else None
    with torch.cuda.device(dev):
        _ffi.check(L.ronk_average_precision_records(ctypes.c_void_p(rec.data_ptr()) if n else None, n, ctypes.c_void_p(g.data_ptr()), C,
                                                    thr_c, len(thr), p('ap07'), p('ap12'), p('offsets'), p('sorted'),
                                                    p('precision'), p('recall'), ctypes.c_void_p(ws.data_ptr()), need,
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def gather_tp_fp_records(state, group=None):
    """The device-resident half of ``gather_tp_fp``: all-gather the packed records of every rank's ``TpFpDeviceState``
    (NCCL: one small all_gather of the counts, one all_gather_into_tensor of the records, one all_reduce of the
    ground-truth counts) and leave them on the device.  -> (records int64 [world * width] rank-major, the unused tail of
    every rank's row set to ``PAD_RECORD``; n_gt int64 [C-1]).  Without a process group: this rank's own records."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    par = state.calls & 1
    if not multi:
        tot = state.totals.cpu().numpy()
        if tot[2]:
            raise RuntimeError('TpFpDeviceState: more than %d records, raise `capacity`' % state.capacity)
        return state.records[:int(tot[par])], state.n_gt
    world = dist.get_world_size(group)
    meta = torch.stack([state.totals[par], state.totals[2]]).to(torch.int64)
    all_meta = torch.empty((world, 2), dtype=torch.int64, device=state.device)
    dist.all_gather_into_tensor(all_meta, meta, group=group)
    n_gt = state.n_gt.clone()
    dist.all_reduce(n_gt, op=dist.ReduceOp.SUM, group=group)
    host_meta = all_meta.cpu().numpy()                       # the one sync: the width of the gather
    if host_meta[:, 1].any():
        raise RuntimeError('TpFpDeviceState: a rank ran out of record capacity (%d)' % state.capacity)
    width = max(int(host_meta[:, 0].max()), 1)
    if width > state.capacity:
        raise RuntimeError('TpFpDeviceState: capacity %d below the largest rank (%d records)' % (state.capacity, width))
    out = torch.empty((world, width), dtype=torch.int64, device=state.device)
    dist.all_gather_into_tensor(out, state.records[:width].contiguous(), group=group)
    valid = torch.arange(width, device=state.device)[None, :] < all_meta[:, :1]
    out = torch.where(valid, out, torch.full((), PAD_RECORD, dtype=torch.int64, device=state.device))
    return out.view(-1), n_gt


def gather_tp_fp(state, num_classes, group=None):
    """All-gather the TP/FP accumulators of every rank (NCCL when the default process group is NCCL, Gloo on CPU test
    runs) and return the merged per-class ``state`` (dict class -> TpFpAccumulator) on every rank; the concatenation
    order is rank-major, which is the order a single process would have produced for a contiguous image split.
    ``state`` is a ``TpFpDeviceState`` (records never left the device: one all_gather of the counts, one
    all_gather_into_tensor of the packed records, one all_reduce of the ground-truth counts, one copy to the host) or
    the dict of host accumulators ``streaming_tp_fp_arrays`` returns (same collectives after one upload)."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if isinstance(state, TpFpDeviceState):
        if not multi:
            return state.to_host()
        world = dist.get_world_size(group)
        CM = state.num_classes - 1
        # one small all_gather: record count, overflow flag, number of updates of every rank
        meta = torch.cat([state.totals[(state.calls & 1):(state.calls & 1) + 1], state.totals[2:3]]).to(torch.int64)
        meta = torch.cat([meta, torch.tensor([state.calls], dtype=torch.int64, device=state.device)])
        all_meta = torch.empty((world, 3), dtype=torch.int64, device=state.device)
        dist.all_gather_into_tensor(all_meta, meta, group=group)
        n_gt = state.n_gt.clone()
        dist.all_reduce(n_gt, op=dist.ReduceOp.SUM, group=group)
        all_meta = all_meta.cpu().numpy()                       # the one sync before the big gather: its width
        if all_meta[:, 1].any():
            raise RuntimeError('TpFpDeviceState: a rank ran out of record capacity (%d)' % state.capacity)
        width = max(int(all_meta[:, 0].max()), 1)
        calls = max(int(all_meta[:, 2].max()), 1)
        if width > state.capacity or calls > state.seg.shape[0]:
            raise RuntimeError('TpFpDeviceState: capacity %d / max_updates %d below the largest rank (%d records, %d updates)'
                               % (state.capacity, state.seg.shape[0], width, calls))
        out = torch.empty((world, width), dtype=torch.int64, device=state.device)
        dist.all_gather_into_tensor(out, state.records[:width].contiguous(), group=group)
        segs = torch.empty((world, calls, CM), dtype=torch.int32, device=state.device)
        dist.all_gather_into_tensor(segs, state.seg[:calls].contiguous(), group=group)
        host, segs = out.cpu().numpy(), segs.cpu().numpy()       # (a pinned staging buffer only pays when it is reused: measured)
        parts = [(host[r, :int(all_meta[r, 0])], segs[r, :int(all_meta[r, 2])]) for r in range(world)]
        return TpFpDeviceState._assemble(parts, n_gt.cpu().numpy(), state.num_classes)
    if not multi:
        return state
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    classes = list(range(1, num_classes))
    empty = TpFpAccumulator()
    accs = [state.get(c, empty) for c in classes]
    counts = torch.tensor([a.scores.shape[0] for a in accs], dtype=torch.int64)
    # one packed upload: per class block of (score bits << 32 | fp << 1 | tp)
    packed = np.concatenate([(a.scores.view(np.uint32).astype(np.int64) << 32) | (a.fp.astype(np.int64) << 1) | a.tp.astype(np.int64)
                             for a in accs]) if int(counts.sum()) else np.zeros((0,), np.int64)
    head = torch.cat([counts, torch.tensor([a.n_gt for a in accs], dtype=torch.int64)]).to(dev)
    all_head = torch.empty((world, head.numel()), dtype=torch.int64, device=dev)
    if backend == 'nccl':
        dist.all_gather_into_tensor(all_head, head, group=group)
    else:
        parts = [torch.empty_like(head) for _ in range(world)]
        dist.all_gather(parts, head, group=group)
        all_head = torch.stack(parts)
    all_head = all_head.cpu().numpy()
    all_counts = all_head[:, :len(classes)]
    width = max(int(all_counts.sum(1).max()), 1)
    rec = torch.zeros((width,), dtype=torch.int64)
    rec[:packed.shape[0]] = torch.from_numpy(packed)
    rec = rec.to(dev)
    if backend == 'nccl':
        out = torch.empty((world, width), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(out, rec, group=group)
    else:
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec, group=group)
        out = torch.stack(parts)
    host = out.cpu().numpy()
    merged = {}
    for i, c in enumerate(classes):
        acc = TpFpAccumulator()
        acc.n_gt = int(all_head[:, len(classes) + i].sum())
        blocks = []
        for r in range(world):
            off = int(all_counts[r, :i].sum())
            blocks.append(host[r, off:off + int(all_counts[r, i])])
        blk = np.concatenate(blocks) if blocks else np.zeros((0,), np.int64)
        acc.scores = (blk >> 32).astype(np.uint32).view(np.float32).copy()
        acc.tp = (blk & 1).astype(bool)
        acc.fp = ((blk >> 1) & 1).astype(bool)
        merged[c] = acc
    return merged


def gather_detections(scores, bboxes, group=None):
    """All-gather the final detections of every rank (the reference keeps them for the debug dump and the TP/FP
    matching, eval_ron_network.py:230-252): scores [B,C-1,M] / bboxes [B,C-1,M,4] (or dicts class -> [B,M] / [B,M,4])
    -> the same structures for the world_size * B images, rank-major (a contiguous image split).  One
    all_gather_into_tensor per tensor over NCCL (Gloo on CPU test runs); no-op without a process group."""
    import torch.distributed as dist
    if isinstance(scores, dict):
        keys = sorted(scores.keys())
        s, b = gather_detections(torch.stack([scores[k] for k in keys], 1), torch.stack([bboxes[k] for k in keys], 1), group)
        return {k: s[:, i] for i, k in enumerate(keys)}, {k: b[:, i] for i, k in enumerate(keys)}
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return scores, bboxes
    world = dist.get_world_size(group)
    outs = []
    for t in (scores, bboxes):
        t = t.contiguous()
        if dist.get_backend(group) == 'nccl':
            o = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(o, t, group=group)
        else:
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t, group=group)
            o = torch.cat(parts, 0)
        outs.append(o)
    return outs[0], outs[1]
