"""Monotonic alignment search — mirror of reference fs2/attn/alignment.py:21-85.

Same call signatures (numpy float32 `mel × text` log-probabilities in, same-shape 0/1 out), but
the DP + backtrack run in the batched CUDA kernel (`fs2k_mas_fwd`).  `mas_cuda` is the tensor
entry point the model uses (no host round trip)."""
import numpy as np
import torch

from ... import ops


def mas_cuda(attn: torch.Tensor, in_lens: torch.Tensor, out_lens: torch.Tensor, take_log: bool = False, dense: bool = True):
    """attn [B,1,F,T] CUDA → (path [B,F] int32, durations [B,T] int32, hard [B,1,F,T] or None)."""
    return ops.mas(attn, in_lens, out_lens, take_log=take_log, dense=dense)


def _device():
    from ..._lib import require_device

    require_device()
    return torch.device("cuda", torch.cuda.current_device())


def b_mas(b_log_attn_map, in_lens, out_lens, width=1):
    assert width == 1
    dev = _device()
    x = torch.as_tensor(np.ascontiguousarray(b_log_attn_map, dtype=np.float32), device=dev)
    il = torch.as_tensor(np.asarray(in_lens, dtype=np.int32), device=dev)
    ol = torch.as_tensor(np.asarray(out_lens, dtype=np.int32), device=dev)
    _, _, hard = ops.mas(x, il, ol)
    return hard.cpu().numpy()


def mas_width1(log_attn_map):
    """mas with hardcoded width=1 (alignment.py:48-74)."""
    a = np.ascontiguousarray(log_attn_map, dtype=np.float32)
    n_mel, n_text = a.shape
    return b_mas(a[None, None], [n_text], [n_mel])[0, 0]


def mas(log_attn_map, width=1):
    """General-width entry point of the reference (alignment.py:21-45); only width 1 is ever used."""
    if width != 1:
        raise NotImplementedError("only width=1 is used by FastSpeech2_lightning")
    return mas_width1(log_attn_map)
