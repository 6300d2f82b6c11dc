"""physs_gp_b200 -- B200 (sm_100a) implementation of the state-space inference hot path of
jonathanfrennert/physs_gp, behind the reference's own filter / smoother / prior API.

The compute lives in libphyss_b200.so (hand-written CUDA, C ABI in include/physs_b200.h); this
package is the thin host-side mirror of the reference interface.  There is no CPU fallback.
"""
from . import settings  # noqa: F401
from .dispatch import dispatch, evoke  # noqa: F401

__all__ = ["settings", "dispatch", "evoke"]
