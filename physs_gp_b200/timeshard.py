"""Time-sharded filter + smoother for long series: one contiguous time range per rank (GPU), one
all-gather of per-range scan summaries per pass (BASELINE config 3).

    rank r holds steps [t_r, t_{r+1}) of every series:  Y_r [B, T_r, m], dt_f_r, dt_s_r, R_r ...

  filter    1. local   : chunk summaries + prefix scan of the local range -> total_r  [B, 3d^2+2d]
            2. gather  : all_gather(total_r) over the ranks  (NCCL over NVLink; G * B * (3d^2+2d) * 8 bytes)
            3. fold    : start_r = (m0, P0) pushed through total_0 .. total_{r-1}      (<= G-1 combines)
            4. finish  : boundaries from start_r, concurrent replay of the local chunks, fix-up passes;
                         lml_r = local sum, all_reduce(SUM) for the series' log marginal likelihood
  smoother  the same backwards: total_r [B, 2d^2+d], fold of the terminal state of the LAST rank through
            total_{G-1} .. total_{r+1}, finish with that carried state.

The compute steps are the C-ABI calls physs_pscan_*_{local,fold,finish}_f64 (ops.pscan_*); this module is
only the collective plumbing, written against a small `comm` interface so that the same code runs over
torch.distributed (NCCL on GPUs; gloo in the CPU tests of the host logic) or a single process.

dt conventions per rank (global arrays dt_f = [0, diff(t)], dt_s = [diff(t), 0] sliced to the range):
    dt_f_r[0] is the gap to the previous rank's last step (0 on rank 0),
    dt_s_r[-1] is the gap to the next rank's first step (0 on the last rank).

Fix-up passes of the filter (jitter != 0) stay inside a rank: the first chunk of rank r > 0 starts from the
folded scan state, which is O(jitter) away from the jittered sequential recursion; its error decays over
that chunk exactly as for every other chunk boundary, but is not re-polished across the rank boundary.
`cross_rank_polish=True` (the default whenever jitter != 0) sends each rank's replayed last state to its
successor (one extra message of (d + d^2) * 8 bytes per series) and re-runs the finish step from it.

Collectives per call: one all-gather of the filter range summaries, one (optional) all-gather for the
cross-rank polish, one all-gather that carries the smoother range summaries together with every rank's last
filtered state, lml share and fix-up status (the status returned is the MAX over ranks).
"""
import torch

from . import settings


class SingleProcess:
    """comm for world_size 1."""
    rank, world = 0, 1

    def all_gather(self, x):
        return x[None]

    def gather_buffer(self, shape, device):
        """([world, *shape] buffer, this rank's slot): a kernel that writes its result into the slot needs no copy
        before all_gather_inplace."""
        buf = torch.empty((1,) + tuple(shape), dtype=torch.float64, device=device)
        return buf, buf[0]

    def all_gather_inplace(self, buf):
        return buf

    def shift_from_prev(self, x):
        return None


class TorchDist:
    """comm over torch.distributed (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, x):
        x = x.contiguous()
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        self.dist.all_gather_into_tensor(out, x, group=self.group)     # concatenated along dim 0
        return out.view((self.world,) + tuple(x.shape))

    def gather_buffer(self, shape, device):
        buf = torch.empty((self.world,) + tuple(shape), dtype=torch.float64, device=device)
        return buf, buf[self.rank]

    def all_gather_inplace(self, buf):
        """In-place all-gather: every rank's slot buf[rank] is already filled (NCCL: sendbuff == recvbuff + rank *
        count, no staging copy on either side)."""
        self.dist.all_gather_into_tensor(buf.view(-1), buf[self.rank].reshape(-1), group=self.group)
        return buf

    def shift_from_prev(self, x):
        """Every rank sends x to rank + 1; returns what rank - 1 sent (None on rank 0)."""
        gathered = self.all_gather(x)          # tiny messages: one collective beats G point-to-point pairs
        return gathered[self.rank - 1] if self.rank > 0 else None


def time_ranges(T, world):
    """Contiguous, near-equal split of T steps over `world` ranks: [(t0, t1), ...]."""
    base, rem = divmod(T, world)
    out, t0 = [], 0
    for r in range(world):
        t1 = t0 + base + (1 if r < rem else 0)
        out.append((t0, t1))
        t0 = t1
    return out


def filter_smooth(comm, ops, dt_f, dt_s, Y, R, H, m0, P0, disc_f, disc_s, chunk_len, jitter=None, polish=None,
                  Hout=None, cross_rank_polish=None, ws=None):
    """Filter + smoother of the local time range of a series sharded over comm.world ranks.

    All arguments are the LOCAL slices (see module docstring); disc_f / disc_s are ops.Disc for the local
    range.  Returns (lml [B] for the WHOLE series, mf, Pf, ms, Ps for the local range, status).
    """
    B, T = Y.shape[0], Y.shape[1]
    d = P0.shape[-1]
    if ws is None:
        ws = ops.pscan_workspace(B, T, d, chunk_len, Y.device)
    r, G = comm.rank, comm.world
    # ---------------------------------------------------------------- filter
    # the range summary is written by its kernel straight into this rank's slot of the all-gather buffer
    totals, slot = comm.gather_buffer((B, 3 * d * d + 2 * d), Y.device)
    ops.pscan_filter_local(dt_f, Y, R, H, m0, P0, disc_f, chunk_len, ws, jitter=jitter, out=slot)
    totals = comm.all_gather_inplace(totals)                          # [G, B, ne]
    start = None
    if r > 0:
        m0b = m0.expand(B, d) if m0.dim() == 2 else m0
        start = ops.pscan_filter_fold(totals[:r], m0b, P0.expand(B, d, d))
    out = ops.pscan_filter_finish(dt_f, Y, R, H, m0, P0, disc_f, chunk_len, ws, start=start, jitter=jitter,
                                  polish=polish)
    lml_loc, mf, Pf, status = out
    if cross_rank_polish is None:
        # with jitter != 0 the first chunk of every rank > 0 starts O(jitter) away from the jittered sequential
        # recursion and no rank-local fix-up pass re-visits it: polish across ranks by default
        cross_rank_polish = G > 1 and (settings.jitter if jitter is None else jitter) != 0.0
    if cross_rank_polish and G > 1:
        last = torch.cat([mf[:, -1, :], Pf[:, -1].reshape(B, d * d)], dim=1).contiguous()
        prev = comm.shift_from_prev(last)
        if r > 0:
            start = (prev[:, :d].contiguous(), prev[:, d:].reshape(B, d, d).contiguous())
            lml_loc, mf, Pf, status = ops.pscan_filter_finish(dt_f, Y, R, H, m0, P0, disc_f, chunk_len, ws,
                                                              start=start, jitter=jitter, polish=polish,
                                                              out=(mf, Pf))
    # -------------------------------------------------------------- smoother
    # ONE all-gather for the backward direction: every rank contributes
    #   [ range summary (E, L, g) | its last filtered state (m, P) | its share of the lml | its fix-up status ]
    # so the lml reduction, the terminal state of the last rank and the global status ride along with the
    # summaries instead of three more collectives (the ranks sum the lml shares in rank order: deterministic).
    stotal = ops.pscan_smooth_local(dt_s, mf, Pf, disc_s, chunk_len, ws, jitter=jitter)
    ns = stotal.shape[1]
    pack = torch.cat([stotal, mf[:, -1, :], Pf[:, -1].reshape(B, d * d), lml_loc.reshape(B, 1),
                      status.to(torch.float64).reshape(1, 1).expand(B, 1)], dim=1).contiguous()
    packs = comm.all_gather(pack)                                     # [G, B, ns + d + d*d + 2]
    lml = packs[:, :, ns + d + d * d].sum(dim=0)
    status = packs[:, 0, ns + d + d * d + 1].max().to(status.dtype).reshape(status.shape)
    sstart = None
    if G > 1 and r < G - 1:
        # terminal state = filtered state at the last step of the LAST rank
        m_end = packs[G - 1][:, ns:ns + d].contiguous()
        P_end = packs[G - 1][:, ns + d:ns + d + d * d].reshape(B, d, d).contiguous()
        sstart = ops.pscan_smooth_fold(packs[r + 1:, :, :ns].contiguous(), m_end, P_end)
    ms, Ps = ops.pscan_smooth_finish(dt_s, mf, Pf, disc_s, chunk_len, ws, start=sstart, Hout=Hout, jitter=jitter)
    return lml, mf, Pf, ms, Ps, status
