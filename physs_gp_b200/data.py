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


# ------------------------------------------------------------------------------------------ sort / pad
class SequentialData:
    """Device-side mirror of `SequentialData.sort` / `.unsort` (stgp/data/data.py:353-415) and the numpy helpers
    behind it, `pad_with_nan_to_make_grid` and `order_sequentially_np` (stgp/data/sequential.py:9-144): the
    pre-step of the hot path that pads scattered observations (X [N, 1 + Ds], Y [N, P]) with NaN rows to the full
    time x space grid, sorts them into time-space order and returns Y in time-latent-space format
    [Nt, P, Ns] -- what `filter_loop` consumes.  The reference does this once per model on the host; here the
    same index algebra runs on whatever device X lives on (torch.unique = a device radix sort), so a batch of
    series that arrives on the GPU never goes back to the host.

    Same results as the reference: row order of `torch.unique(dim=0)` is lexicographic like `numpy.unique(axis=0)`,
    the FIRST occurrence of a duplicated location is kept (`return_index`), padded rows are appended behind the
    data so that real observations win over NaN padding."""

    def __init__(self):
        self.num_original_points = None
        self.num_points_added = None
        self.unique_idx = None
        self.reverse_unique_idx = None
        self.sort_idx = None

    @staticmethod
    def _unique_rows(X):
        """(unique rows sorted lexicographically, index of first occurrence, inverse) like
        numpy.unique(X, axis=0, return_index=True, return_inverse=True)."""
        import torch
        uniq, inv = torch.unique(X, dim=0, return_inverse=True)
        n = X.shape[0]
        first = torch.full((uniq.shape[0],), n, dtype=torch.long, device=X.device)
        first.scatter_reduce_(0, inv, torch.arange(n, device=X.device), reduce="amin")
        return uniq, first, inv

    @staticmethod
    def pad_with_nan_to_make_grid(X, Y):
        """sequential.py:9-71.  Returns (points_added, X_grid [N + added, D], Y_grid [N + added, P] or None)."""
        import torch
        N = X.shape[0]
        ut = torch.unique(X[:, 0])
        us = torch.unique(X[:, 1:], dim=0) if X.shape[1] > 1 else X.new_zeros((1, 0))
        Nt, Ns = ut.shape[0], us.shape[0]
        grid = torch.cat([ut[:, None, None].expand(Nt, Ns, 1), us[None].expand(Nt, Ns, us.shape[1])], dim=2)
        grid = grid.reshape(Nt * Ns, -1)
        both = torch.cat([X, grid], dim=0)
        _, first, _ = SequentialData._unique_rows(both)
        idx = first[first >= N]
        X_add = both[idx]
        X_grid = torch.cat([X, X_add], dim=0)
        Y_grid = None
        if Y is not None:
            Y_grid = torch.cat([Y, torch.full((idx.shape[0], Y.shape[1]), float("nan"), dtype=Y.dtype,
                                              device=Y.device)], dim=0)
        return int(idx.shape[0]), X_grid, Y_grid

    @staticmethod
    def order_sequentially(X, Y=None):
        """sequential.py:73-144.  Returns (unique_idx, reverse_idx, sort_idx, X [Nt, Ns, D], Y [Nt, Ns, P] or None)."""
        import torch
        uniq, first, inv = SequentialData._unique_rows(X)
        t0 = X[0, 0]
        grid_size = int((uniq[:, 0] == t0).sum())
        time_points = uniq.shape[0] // grid_size
        sort_idx = torch.arange(uniq.shape[0], device=X.device)       # unique() already sorted (sequential.py:121-126)
        Xs = uniq.reshape(time_points, grid_size, X.shape[1])
        Ys = None if Y is None else Y[first].reshape(time_points, grid_size, Y.shape[1])
        return first, inv, sort_idx, Xs, Ys

    def sort(self, X, Y):
        """data.py:353-390: pad to the grid, order in time-space, Y -> time-latent-space [Nt, P, Ns]."""
        self.num_original_points = X.shape[0]
        added, Xp, Yp = self.pad_with_nan_to_make_grid(X, Y)
        unique_idx, reverse_idx, sort_idx, Xs, Ys = self.order_sequentially(Xp, Yp)
        self.num_points_added = added
        self.unique_idx, self.reverse_unique_idx, self.sort_idx = unique_idx, reverse_idx, sort_idx
        if Ys is not None:
            Ys = Ys.transpose(1, 2)
        return Xs, Ys

    def unsort(self, A):
        """data.py:413-415."""
        return A[self.sort_idx][self.reverse_unique_idx][:self.num_original_points]
