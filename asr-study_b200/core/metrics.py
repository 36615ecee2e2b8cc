"""Label error rate — core/metrics.py:4-8: mean over the batch of
tf.edit_distance(hyp, truth, normalize=True).  Operates on the -1-padded dense label
matrices the decode kernels emit; the dynamic program itself is host arithmetic on a
few dozen integers per utterance (negligible), not a device kernel."""
import numpy as np


def _lev(a, b):
    prev = np.arange(len(b) + 1)
    for i, ca in enumerate(a, 1):
        cur = np.empty(len(b) + 1, dtype=np.int64)
        cur[0] = i
        for j, cb in enumerate(b, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb))
        prev = cur
    return int(prev[len(b)])


def _rows(x):
    if hasattr(x, "tocsr"):                          # scipy sparse labels (the batch contract)
        x = x.tocsr()
        return [x.data[x.indptr[i]:x.indptr[i + 1]].tolist() for i in range(x.shape[0])]
    if hasattr(x, "cpu"):
        x = x.cpu().numpy()
    if isinstance(x, np.ndarray) and x.ndim == 2:
        return [[int(v) for v in r if v >= 0] for r in x]
    return [list(map(int, r)) for r in x]


def ler(y_true, y_pred, **kwargs):
    t, h = _rows(y_true), _rows(y_pred)
    vals = []
    for a, b in zip(t, h):
        d = _lev(b, a)
        vals.append(d / len(a) if len(a) else (float("inf") if d else 0.0))
    return float(np.mean(vals)) if vals else 0.0
