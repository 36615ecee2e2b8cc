"""CTC ops with the reference's call shapes (core/ctc_utils.py:8-82), on the CUDA kernels.

``ctc_lambda_func([y_pred, labels, inputs_length])`` -> per-utterance loss [N]
``decode([y_pred, seq_len], is_greedy=True, beam_width=100, top_paths=1, merge_repeated=True)``
    -> int32 matrix [N, T] padded with -1 (the to_dense form, core/layers_utils.py:54-57)
y_pred is the batch-major logits tensor [N, T, C] (CUDA, fp32) as in the reference; it is
re-laid out time-major like the reference's own tf.transpose before the kernels run.
"""
import numpy as np
import torch

from .._lib import cur_stream, lib, ptr


def _labels_to_flat(labels, device):
    if hasattr(labels, "tocsr"):
        m = labels.tocsr()
        rows = [m.data[m.indptr[i]:m.indptr[i + 1]] for i in range(m.shape[0])]
    else:
        rows = [np.asarray(r) for r in labels]
    lens = [len(r) for r in rows]
    off = np.zeros(len(rows) + 1, np.int32)
    off[1:] = np.cumsum(lens)
    flat = np.concatenate(rows).astype(np.int32) if sum(lens) else np.zeros(1, np.int32)
    return torch.as_tensor(flat, device=device), torch.as_tensor(off, device=device), int(max(lens) if lens else 0)


def _seq_len(inputs_length, device):
    x = torch.as_tensor(np.asarray(inputs_length)) if not torch.is_tensor(inputs_length) else inputs_length
    if x.dim() == 2:
        x = x[:, 0]                                   # the reference passes [N,1] and takes [:, 0]
    return x.to(device=device, dtype=torch.int32).contiguous()


def ctc_lambda_func(args):
    y_pred, labels, inputs_length = args
    N, T, C = y_pred.shape
    lg = y_pred.transpose(0, 1).contiguous()          # time-major, as tf.transpose(perm=[1,0,2])
    flat, off, mx = _labels_to_flat(labels, y_pred.device)
    ws = torch.empty(lib.asr_ctc_workspace_bytes(T, N, mx) // 4 + 1, dtype=torch.float32, device=y_pred.device)
    loss = torch.empty(N, dtype=torch.float32, device=y_pred.device)
    grad = torch.empty(T, N, C, dtype=torch.float32, device=y_pred.device)
    lib.asr_ctc_loss_grad(ptr(lg), T, N, C, ptr(_seq_len(inputs_length, y_pred.device)), ptr(flat), ptr(off), mx,
                          C - 1, 1.0, ptr(loss), ptr(grad), ptr(ws), cur_stream())
    return loss


def decode(inputs, **kwargs):
    is_greedy = kwargs.get("is_greedy", True)
    y_pred, seq_len = inputs
    N, T, C = y_pred.shape
    lg = y_pred.transpose(0, 1).contiguous()
    sl = _seq_len(seq_len, y_pred.device)
    out = torch.empty(N, T, dtype=torch.int32, device=y_pred.device)
    out_len = torch.empty(N, dtype=torch.int32, device=y_pred.device)
    if is_greedy:
        lib.asr_ctc_greedy(ptr(lg), T, N, C, ptr(sl), C - 1, 1, ptr(out), ptr(out_len), cur_stream())
    else:
        bw = int(kwargs.get("beam_width", 100))
        if int(kwargs.get("top_paths", 1)) != 1:
            raise ValueError("only top_paths=1 is built (the reference uses [0][0] only)")
        ws = torch.empty(lib.asr_ctc_beam_workspace_bytes(T, N, C, bw) // 4 + 1, dtype=torch.int32, device=y_pred.device)
        lib.asr_ctc_beam(ptr(lg), T, N, C, ptr(sl), C - 1, bw, int(kwargs.get("merge_repeated", True)), ptr(out),
                         ptr(out_len), ptr(ws), cur_stream())
    return out


def decode_output_shape(inputs_shape):
    y_pred_shape, _ = inputs_shape
    return (y_pred_shape[:1], None)


def ctc_dummy_loss(y_true, y_pred):
    return y_pred


def decoder_dummy_loss(y_true, y_pred):
    return 0.0
