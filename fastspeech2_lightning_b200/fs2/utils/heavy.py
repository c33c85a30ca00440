"""mask_from_lens — mirror of reference fs2/utils/heavy.py:11-15 (plotting helpers are out of scope)."""
from typing import Optional


from ... import ops


def mask_from_lens(lens, max_len: Optional[int] = None):
    """True on valid positions; lens int tensor [B] on the GPU."""
    if max_len is None:
        max_len = int(lens.max())
    return ops.lens_mask(lens, int(max_len))
