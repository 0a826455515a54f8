"""Drop-in for the hot-path consumers in the reference's ``tf_extended/metrics.py``:
``streaming_tp_fp_arrays`` :133-206, ``precision_recall`` :100-130, ``average_precision_voc07``
:237-258, ``average_precision_voc12`` :212-234.

The reference accumulates TP/FP records in TF local variables inside one process.  Here the
records live on the host in float64/bool NumPy arrays (they are tiny), and ``gather_tp_fp``
is the one collective of the whole path: every rank contributes its per-class
(score, tp, fp) records and ground-truth counts, NCCL (torch.distributed) all-gathers them
once at the end of the evaluation, and rank 0 computes AP exactly as the reference does.
"""
import numpy as np
import torch

__all__ = ['streaming_tp_fp_arrays', 'precision_recall', 'average_precision_voc07', 'average_precision_voc12',
           'TpFpAccumulator', 'gather_tp_fp']


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


def gather_tp_fp(state, num_classes, group=None):
    """All-gather the per-class accumulators of every rank (NCCL when the default process group
    is NCCL, Gloo on CPU test runs): one all_gather of padded (score, tp, fp) records plus one
    all_reduce(SUM) of the ground-truth counts.  Returns a merged ``state`` on every rank; the
    concatenation order is rank-major, which is the order a single process would have produced
    for a contiguous image split."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return state
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    classes = list(range(1, num_classes))
    counts = torch.tensor([state[c].scores.shape[0] if c in state else 0 for c in classes], dtype=torch.int64, device=dev)
    n_gt = torch.tensor([state[c].n_gt if c in state else 0 for c in classes], dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    dist.all_reduce(n_gt, op=dist.ReduceOp.SUM, group=group)
    width = int(torch.stack(all_counts).sum(1).max().item())
    rec = torch.zeros((max(width, 1), 2), dtype=torch.float32, device=dev)
    o = 0
    for c in classes:
        if c in state and state[c].scores.shape[0]:
            n = state[c].scores.shape[0]
            rec[o:o + n, 0] = torch.from_numpy(state[c].scores).to(dev)
            rec[o:o + n, 1] = torch.from_numpy(state[c].tp.astype(np.float32) + 2 * state[c].fp.astype(np.float32)).to(dev)
            o += n
    all_rec = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(all_rec, rec, group=group)
    merged = {}
    for i, c in enumerate(classes):
        acc = TpFpAccumulator()
        acc.n_gt = int(n_gt[i].item())
        parts_s, parts_t, parts_f = [], [], []
        for r in range(world):
            cnt = all_counts[r].cpu().numpy()
            off = int(cnt[:i].sum())
            n = int(cnt[i])
            blk = all_rec[r][off:off + n].cpu().numpy()
            parts_s.append(blk[:, 0].astype(np.float32))
            code = blk[:, 1].astype(np.int64)
            parts_t.append((code & 1).astype(bool))
            parts_f.append((code & 2).astype(bool))
        acc.scores = np.concatenate(parts_s) if parts_s else acc.scores
        acc.tp = np.concatenate(parts_t) if parts_t else acc.tp
        acc.fp = np.concatenate(parts_f) if parts_f else acc.fp
        merged[c] = acc
    return merged
