"""CPU oracle for the physs_gp state-space inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package (`physs_gp_b200/`)
may import, link or execute anything under `oracle/`.  The only legitimate
callers are `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs, and there only as the checker (or the
reported CPU baseline), never as the thing measured or shipped.

What it is: a plain numpy/scipy (and, for speed, plain C in `ssm_oracle.c`)
restatement of the reference's algorithm for this path, every function citing
the `/root/reference` file:line it follows (paths below are relative to
`/root/reference/src/lib/stgp/`).

Pinning status: the reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), and `jax`/`objax`/`chex`/`batchjax` are not installable
here, so the reference cannot run on its real numerical stack.  The oracle is
pinned in two ways instead:
  1. `tests/golden/make_golden.py` executes the reference's OWN unmodified
     source files from `/root/reference` on a small numpy-backed stand-in for
     the `jax` API (built inside `make_golden.py` itself), and commits the input/output
     vectors as fixtures; `tests/test_golden.py` checks the oracle
     against them.  This pins the algorithm (step order, jitter placement,
     masks, dt conventions) to the reference source, but on numpy/LAPACK
     arithmetic rather than XLA's.
  2. derived known-answer tests the reference supports by construction: dense
     GP marginal likelihood / posterior, scipy.linalg.expm, the Lyapunov
     equation, the CVI Gaussian fixed point, the Poisson closed-form ELL.
Rows without either (the Gauss-Hermite expected log-likelihood, whose
reference branch is dead code) are marked "parity unpinned" where they are
defined.
"""
