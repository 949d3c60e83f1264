"""Stand-in for lucidrains' `rotary_embedding_torch` (>=0.2.3, un-pinned by the reference README).

The real package is NOT in this container, so this restates its published 0.2.x algorithm:
freqs = theta^(-arange(0,dim,2)/dim) held as a frozen nn.Parameter called `freqs`;
rotate_queries_or_keys rotates interleaved pairs (x[2i], x[2i+1]) by angle pos*freqs[i],
positions 0..n-1 along dim -2.  "parity unpinned" at this boundary: see DESIGN.md.
Test infrastructure only (oracle/make_golden.py).
"""
import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000):
        super().__init__()
        exponents = torch.arange(0, dim, 2)[: dim // 2].float() / dim
        self.freqs = nn.Parameter(1.0 / (theta ** exponents), requires_grad=False)

    def rotate_queries_or_keys(self, t, seq_dim=-2, offset=0):
        n = t.shape[seq_dim]
        pos = torch.arange(n, device=t.device).type(self.freqs.dtype) + offset
        ang = pos[:, None] * self.freqs[None, :]          # (n, dim/2)
        ang = ang.repeat_interleave(2, dim=-1)            # (n, dim): a0 a0 a1 a1 ...
        even, odd = t[..., 0::2], t[..., 1::2]
        rot = torch.stack((-odd, even), dim=-1).flatten(-2)
        return t * ang.cos() + rot * ang.sin()
