"""CUDA kernels (through the C ABI) against the CPU oracle on fresh seeded states at a batch size the
oracle finishes in seconds, plus size-independent properties at the benchmark's full size."""
import os
import sys

import numpy as np
import pytest

from conftest import ENVS
from emu_util import load_golden
from tolerances import GRAD_RTOL, fwd_rtol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def _batch(name, N, seed):
    d, model = load_golden(name)
    from oracle import Oracle
    n0 = int(d["meta/num_envs"])
    o = Oracle.from_model(model, n0)
    Q, D, M = o.desc.Q, o.desc.D, o.desc.M
    rng = np.random.default_rng(seed)
    p = "case%d/" % (int(d["meta/num_cases"]) - 1)
    pick = rng.integers(0, n0, N)
    q0 = (d[p + "q0"].reshape(n0, Q)[pick] + 2e-3 * rng.standard_normal((N, Q))).astype(np.float32)
    qd0 = (d[p + "qd0"].reshape(n0, D)[pick] + 2e-2 * rng.standard_normal((N, D))).astype(np.float32)
    act = (d[p + "act"].reshape(n0, D)[pick] * rng.uniform(0.5, 1.5, (N, D))).astype(np.float32)
    musc = (d[p + "musc"].reshape(n0, M)[pick] * rng.uniform(0.5, 1.5, (N, M))).astype(np.float32) if M else None
    cfg = dict(S=int(d["meta/substeps"]), mm=int(d["meta/mass_matrix_freq"]), dt=float(d["meta/dt"]))
    return o, q0, qd0, act, musc, cfg


@pytest.mark.parametrize("name", ENVS)
def test_cuda_forward_matches_cpu_oracle(name):
    import torch
    from diffrl_b200.engine import ArticulationEngine
    N = 64
    o, q0, qd0, act, musc, c = _batch(name, N, 5)
    oq, oqd = o.forward(q0, qd0, act, musc, c["S"], c["mm"], c["dt"])
    eng = ArticulationEngine(o.desc, N, "cuda:0")
    t = lambda a: None if a is None else torch.tensor(a.ravel(), device="cuda:0")
    q, qd, _, _ = eng.forward(t(q0), t(qd0), t(act), t(musc), c["S"], c["mm"], c["dt"], want_tape=False)
    Q, D = o.desc.Q, o.desc.D
    from parity_util import check_forward_against

    def fwd_one(i, qi, qdi):
        return o.forward(qi[None], qdi[None], act[i:i + 1], None if musc is None else musc[i:i + 1], c["S"], c["mm"], c["dt"])

    # element-wise per environment; an environment outside the tolerance must be ill-conditioned (tests/parity_util.py)
    check_forward_against(fwd_one, q.cpu().numpy(), qd.cpu().numpy(), oq, oqd, q0, qd0, Q, D, fwd_rtol(name), name, max_bad=4)


@pytest.mark.parametrize("name", ["AntEnv", "CheetahEnv"])
def test_cuda_adjoint_matches_fp64_finite_differences(name):
    import torch
    from diffrl_b200.engine import ArticulationEngine
    N = 6
    o, q0, qd0, act, musc, c = _batch(name, N, 9)
    Q, D = o.desc.Q, o.desc.D
    rng = np.random.default_rng(3)
    gq_out, gqd_out = rng.standard_normal((N, Q)).astype(np.float32), rng.standard_normal((N, D)).astype(np.float32)
    eng = ArticulationEngine(o.desc, N, "cuda:0")
    t = lambda a: None if a is None else torch.tensor(a.ravel(), device="cuda:0")
    _, _, tape, _ = eng.forward(t(q0), t(qd0), t(act), t(musc), c["S"], c["mm"], c["dt"])
    gq, gqd, gact, _ = eng.backward(t(act), t(musc), tape, t(gq_out), t(gqd_out), c["S"], c["mm"], c["dt"])
    errs = []
    for e in range(N):
        f = o.fd_gradient(q0[e], qd0[e], act[e], None, gq_out[e], gqd_out[e], c["S"], c["mm"], c["dt"])
        scale = max(np.abs(f[0]).max(), np.abs(f[1]).max(), np.abs(f[2]).max())
        errs.append(max(np.abs(gq.cpu().numpy().reshape(N, Q)[e] - f[0]).max(), np.abs(gqd.cpu().numpy().reshape(N, D)[e] - f[1]).max(),
                        np.abs(gact.cpu().numpy().reshape(N, D)[e] - f[2]).max()) / scale)
    assert np.median(errs) < 1e-3, errs     # fp32 adjoint vs fp64 differences of a stiff contact model


def test_full_size_properties_ant_4096():
    """BASELINE.json configs[1] size: determinism, env-independence (permutation equivariance), the
    tape == state-entering-each-substep property, step-splitting, linearity of the adjoint."""
    import torch
    from diffrl_b200.engine import ArticulationEngine
    N = 4096
    o, q0, qd0, act, musc, c = _batch("AntEnv", N, 21)
    eng = ArticulationEngine(o.desc, N, "cuda:0")
    t = lambda a: torch.tensor(a.ravel(), device="cuda:0")
    q, qd, tape, _ = eng.forward(t(q0), t(qd0), t(act), None, c["S"], c["mm"], c["dt"])
    q2, qd2, tape2, _ = eng.forward(t(q0), t(qd0), t(act), None, c["S"], c["mm"], c["dt"])
    assert torch.equal(q, q2) and torch.equal(qd, qd2) and torch.equal(tape, tape2)          # deterministic (no atomics)
    perm = np.random.default_rng(0).permutation(N)
    qp, qdp, _, _ = eng.forward(t(q0[perm]), t(qd0[perm]), t(act[perm]), None, c["S"], c["mm"], c["dt"], want_tape=False)
    Q, D = o.desc.Q, o.desc.D
    assert torch.equal(qp.view(N, Q), q.view(N, Q)[torch.tensor(perm, device="cuda:0")])       # envs are independent
    first = eng.tape_rows(tape, 1)[0]            # rows of substep 0, env-major whatever the kernel family's layout
    assert torch.equal(first[:, :Q].reshape(-1), t(q0)) and torch.equal(first[:, Q:Q + D].reshape(-1), t(qd0))
    # one 16-substep call with the mass matrix refreshed every 8 == two chained 8-substep calls
    qa, qda, _, _ = eng.forward(t(q0), t(qd0), t(act), None, 16, 8, c["dt"], want_tape=False)
    qb, qdb, _, _ = eng.forward(t(q0), t(qd0), t(act), None, 8, 8, c["dt"] / 2, want_tape=False)
    qb, qdb, _, _ = eng.forward(qb, qdb, t(act), None, 8, 8, c["dt"] / 2, want_tape=False)
    assert torch.equal(qa, qb) and torch.equal(qda, qdb)
    gq, gqd, gact, _ = eng.backward(t(act), None, tape, torch.ones_like(q), torch.ones_like(qd), c["S"], c["mm"], c["dt"])
    assert torch.isfinite(gq).all() and torch.isfinite(gqd).all() and torch.isfinite(gact).all()
    g2 = eng.backward(t(act), None, tape, 2 * torch.ones_like(q), 2 * torch.ones_like(qd), c["S"], c["mm"], c["dt"])
    assert torch.allclose(g2[0], 2 * gq, rtol=1e-5, atol=1e-4 * float(gq.abs().max()))        # adjoint is linear
