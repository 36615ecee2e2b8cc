"""helpers shared by the -m gpu parity tests (CUDA path through the C ABI vs oracle/)."""
import numpy as np
import torch


def norm_err(a, b):
    """max |a-b| / max |b| — the norm-wise relative error the 1e-3 activation bar is stated in."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()
