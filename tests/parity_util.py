"""Shared parity criteria of the GPU tests: element-wise tolerances per environment, and -- instead of a blanket
allowance of outlier environments -- a per-case argument: an environment outside the tolerance is accepted only when the
CHECKER's own result moves by a comparable amount under rounding-sized (1-4 ulp) perturbations of that environment's input (cond(H) ~ 1e4 and the
switching surfaces of the contact / limit model amplify rounding; two correct fp32 implementations cannot agree better)."""
import numpy as np


def elementwise_bad_envs(a, b, width, rtol, atol_frac=1e-6):
    """environments with a component outside |a - b| <= rtol |b| + rtol scale_env + atol_frac max|b| (scale_env = the
    environment's largest component: components that are differences of O(scale) terms carry the rounding of the scale)"""
    a = np.asarray(a, np.float64).reshape(-1, width); b = np.asarray(b, np.float64).reshape(-1, width)
    scale = np.abs(b).max(axis=1, keepdims=True)
    ok = np.abs(a - b) <= rtol * np.abs(b) + rtol * scale + atol_frac * np.abs(b).max()
    return np.nonzero(~ok.all(axis=1))[0]


def _probes(q_row, qd_row, n_random=10, seed=11):
    """perturbed copies of one environment's state: the two uniform one-ulp scalings, then random per-component 1-4 ulp ones
    (a switching surface is crossed by SOME rounding-sized change of the inputs, not necessarily by the uniform one)"""
    yield q_row * (1.0 + 1.2e-7), qd_row * (1.0 - 1.2e-7)
    yield q_row * (1.0 - 1.2e-7), qd_row * (1.0 + 1.2e-7)
    r = np.random.default_rng(seed)
    for _ in range(n_random):
        k = float(r.integers(1, 5))
        yield ((q_row * (1.0 + 1.2e-7 * k * r.choice([-1.0, 1.0], q_row.shape))).astype(q_row.dtype),
               (qd_row * (1.0 + 1.2e-7 * k * r.choice([-1.0, 1.0], qd_row.shape))).astype(qd_row.dtype))


def check_forward_against(forward_one, got_q, got_qd, ref_q, ref_qd, q0, qd0, Q, D, tol, what, max_bad=8, amplification=8.0):
    """got / ref: [n, Q], [n, D].  forward_one(i, q_row, qd_row) -> (q', qd') of the checker for environment i from a (perturbed)
    state.  Every environment outside `tol` must be ill-conditioned: error <= amplification x the checker's one-ulp sensitivity."""
    got_q, ref_q = np.asarray(got_q, np.float64).reshape(-1, Q), np.asarray(ref_q, np.float64).reshape(-1, Q)
    got_qd, ref_qd = np.asarray(got_qd, np.float64).reshape(-1, D), np.asarray(ref_qd, np.float64).reshape(-1, D)
    bad = sorted(set(elementwise_bad_envs(got_q, ref_q, Q, tol)) | set(elementwise_bad_envs(got_qd, ref_qd, D, tol)))
    assert len(bad) <= max_bad, (what, len(bad))
    sq, sqd = np.abs(ref_q).max(), np.abs(ref_qd).max()
    for i in bad:
        moved = 0.0
        for pq0, pqd0 in _probes(q0[i], qd0[i]):
            pq, pqd = forward_one(i, pq0, pqd0)
            moved = max(moved, np.abs(np.asarray(pq, np.float64).ravel() - ref_q[i]).max() / sq, np.abs(np.asarray(pqd, np.float64).ravel() - ref_qd[i]).max() / sqd)
        err = max(np.abs(got_q[i] - ref_q[i]).max() / sq, np.abs(got_qd[i] - ref_qd[i]).max() / sqd)
        assert err <= max(tol, amplification * moved), (what, "env %d: error %.2e, one-ulp sensitivity of the checker %.2e" % (i, err, moved))
    return bad


def check_gradients_against(grads_one, got, ref, widths, q0, qd0, tol, what, skip=(), max_bad=8, amplification=8.0):
    """got / ref: lists of [n, w] arrays (gq, gqd, gact[, gmusc]).  grads_one(i, q_row, qd_row) -> concatenated gradient of the
    checker for environment i.  An environment outside `tol` must sit on a switching surface: the checker's own gradient jumps by
    a comparable amount under rounding-sized perturbations of its input (_probes)."""
    n = np.asarray(ref[0]).reshape(-1, widths[0]).shape[0]
    suspects = set()
    for g, r, w in zip(got, ref, widths):
        suspects |= set(elementwise_bad_envs(np.asarray(g).reshape(n, w), np.asarray(r).reshape(n, w), w, tol))
    suspects -= set(skip)
    assert len(suspects) <= max_bad, (what, len(suspects))
    for i in sorted(suspects):
        base = np.asarray(grads_one(i, q0[i], qd0[i]), np.float64)
        scale = np.abs(base).max()
        jump = 0.0
        for pq0, pqd0 in _probes(q0[i], qd0[i]):
            jump = max(jump, np.abs(np.asarray(grads_one(i, pq0, pqd0), np.float64) - base).max() / scale)
        mine = np.concatenate([np.asarray(g, np.float64).reshape(n, w)[i] for g, w in zip(got, widths)])
        err = np.abs(mine - base).max() / scale
        assert err <= max(tol, amplification * jump), (what, "env %d: gradient error %.2e, one-ulp jump of the checker's own gradient %.2e" % (i, err, jump))
    return suspects
