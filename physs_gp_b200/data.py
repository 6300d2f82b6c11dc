"""Input layout of the path -- mirror of the fields `filter_loop` reads from
stgp/data/data.py:341-719 (`X_time[Nt]`, `X_space[Ns, D]`, `Y_st[Nt, P, Ns]`, `Nt`, `Ns`, `P`),
plus an optional leading batch axis B on Y_st (the reference has none)."""
import numpy as np


class TemporalData:
    def __init__(self, X_time, Y_st, X_space=None):
        self.X_time = X_time
        self.X_space = X_space
        Y = Y_st
        nd = Y.ndim if hasattr(Y, "ndim") else Y.dim()
        if nd == 2:               # [Nt, P]  -> Ns = 1
            Y = Y[..., None]
            nd = 3
        if nd not in (3, 4):
            raise ValueError("Y_st must be [Nt, P, Ns] or [B, Nt, P, Ns]")
        self.Y_st = Y
        self.batched = (nd == 4)
        shp = tuple(Y.shape)
        self.B = shp[0] if self.batched else 1
        self.Nt, self.P, self.Ns = shp[-3], shp[-2], shp[-1]

    @property
    def N(self):
        return self.Nt * self.Ns


SpatioTemporalData = TemporalData
MultiOutputTemporalData = TemporalData
