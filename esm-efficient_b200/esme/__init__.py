"""`esme` -- B200-native drop-in for the inference API of uci-cbcl/esm-efficient.

    from esme import ESM2, ESMC, ESM, tokenize
    from esme.alphabet import tokenize_unpad

Same classes, methods and tensor contracts as the reference package; the forward
pass runs in hand-written sm_100a CUDA kernels (libesmk.so, C ABI in
include/esmk.h).  CUDA-only: there is no CPU or eager fallback.
"""
from .alphabet import tokenize
from .esm import ESM, ESM1b, ESM1v, ESM2, ESMC


__all__ = ['ESM', 'ESMC', 'ESM2', 'ESM1b', 'ESM1v', 'tokenize']
