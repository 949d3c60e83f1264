"""Stand-in for the un-vendored `einops_exts` package (test infrastructure only).

Only the two helpers the reference imports (VDDP:17) are provided; neither does arithmetic.
Used solely by oracle/make_golden.py to import the unmodified reference in this container.
"""
from einops import rearrange


def check_shape(tensor, pattern, **sizes):
    # identity rearrange: raises when the pattern / fixed sizes do not match
    return rearrange(tensor, f"{pattern} -> {pattern}", **sizes)


def rearrange_many(tensors, pattern, **sizes):
    return tuple(rearrange(t, pattern, **sizes) for t in tensors)
