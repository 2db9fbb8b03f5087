"""Edge cases of the C ABI on the GPU: batch sizes around the CTA tile, single substep, mass-matrix
period longer than the step, a model without contacts, argument validation, NULL gradient outputs."""
import ctypes

import numpy as np
import pytest

from emu_util import EmuSim, load_golden
from tolerances import GRAD_RTOL, fwd_rtol

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def _case(name, N, seed=0):
    d, model = load_golden(name)
    n0 = int(d["meta/num_envs"])
    emu = EmuSim(model, n0)
    Q, D = emu.desc.Q, emu.desc.D
    rng = np.random.default_rng(seed)
    p = "case%d/" % (int(d["meta/num_cases"]) - 1)
    pick = rng.integers(0, n0, N)
    q0 = d[p + "q0"].reshape(n0, Q)[pick].copy()
    qd0 = d[p + "qd0"].reshape(n0, D)[pick] * rng.uniform(0.8, 1.2, (N, 1)).astype(np.float32)
    act = d[p + "act"].reshape(n0, D)[pick] * rng.uniform(0.5, 1.5, (N, D)).astype(np.float32)
    emu.N = N
    return d, emu, q0.astype(np.float32), qd0.astype(np.float32), act.astype(np.float32)


@pytest.mark.parametrize("N", [1, 7, 8, 9, 33])
@pytest.mark.parametrize("cfg", [(16, 16), (1, 1), (5, 16), (16, 3)])
def test_batch_sizes_and_substep_patterns(N, cfg):
    import torch
    from diffrl_b200.engine import ArticulationEngine
    S, mm = cfg
    d, emu, q0, qd0, act = _case("AntEnv", N, seed=N)
    dt = float(d["meta/dt"]) * S / 16.0
    eq, eqd, etape, _ = emu.forward(q0.ravel(), qd0.ravel(), act.ravel(), None, S, mm, dt)
    gqo, gqdo = np.ones_like(q0).ravel(), np.ones_like(qd0).ravel()
    egq, egqd, egact, _ = emu.backward(act.ravel(), None, etape, gqo, gqdo, S, mm, dt)
    eng = ArticulationEngine(emu.desc, N, "cuda:0")
    t = lambda a: torch.tensor(a.ravel(), device="cuda:0")
    q, qd, tape, _ = eng.forward(t(q0), t(qd0), t(act), None, S, mm, dt)
    assert tape.numel() == eng.tape_floats(S, mm) == etape.size
    gq, gqd, gact, _ = eng.backward(t(act), None, tape, t(gqo), t(gqdo), S, mm, dt)
    assert rel(q.cpu().numpy(), eq) < fwd_rtol("AntEnv") and rel(qd.cpu().numpy(), eqd) < fwd_rtol("AntEnv")
    assert rel(gq.cpu().numpy(), egq) < 4 * GRAD_RTOL and rel(gact.cpu().numpy(), egact) < 4 * GRAD_RTOL


def test_null_gradient_outputs_and_cotangents():
    """NULL output pointers skip the write (reference: empty adjoint tensor => skip, adjoint.h:336-346);
    NULL cotangents mean zero."""
    import torch
    from diffrl_b200.engine import ArticulationEngine
    d, emu, q0, qd0, act = _case("CheetahEnv", 5)
    S, mm, dt = int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    eng = ArticulationEngine(emu.desc, 5, "cuda:0")
    t = lambda a: torch.tensor(a.ravel(), device="cuda:0")
    q, qd, tape, _ = eng.forward(t(q0), t(qd0), t(act), None, S, mm, dt)
    full = eng.backward(t(act), None, tape, torch.ones_like(q), None, S, mm, dt)
    only_act = eng.backward(t(act), None, tape, torch.ones_like(q), torch.zeros_like(qd), S, mm, dt, need=(False, False, True, False))
    assert only_act[0] is None and only_act[1] is None
    assert torch.allclose(full[2], only_act[2], rtol=1e-5, atol=1e-6 * float(full[2].abs().max()))


def test_argument_validation_returns_cuda_error_codes():
    import torch
    from diffrl_b200 import _capi
    from diffrl_b200.engine import ArticulationEngine
    d, emu, q0, qd0, act = _case("AntEnv", 4)
    eng = ArticulationEngine(emu.desc, 4, "cuda:0")
    lib = _capi.lib()
    z = ctypes.c_void_p(0)
    q = torch.tensor(q0.ravel(), device="cuda:0")
    p = ctypes.c_void_p(q.data_ptr())
    assert lib.dfx_step_forward(eng.pack, 4, 16, 16, 1.0 / 60, z, p, p, z, p, p, z, None, z) == 1      # cudaErrorInvalidValue
    assert lib.dfx_step_forward(eng.pack, 0, 16, 16, 1.0 / 60, p, p, p, z, p, p, z, None, z) == 1
    assert lib.dfx_step_backward(eng.pack, 4, 16, 16, 1.0 / 60, p, z, z, z, z, z, z, z, z, z) == 1      # no tape
    assert lib.dfx_set_group_size(5) != 0 and lib.dfx_set_group_size(0) == 0


def test_pack_rejects_malformed_models():
    from diffrl_b200 import _capi
    from diffrl_b200.engine import ArticulationEngine
    d, emu, *_ = _case("AntEnv", 2)
    bad = emu.desc
    bad.arrays["joint_parent"] = bad.arrays["joint_parent"].copy()
    bad.arrays["joint_parent"][1] = 5          # parent after child
    with pytest.raises(_capi.DfxError):
        ArticulationEngine(bad, 2, "cuda:0")


def test_cartpole_has_no_contacts_and_fixed_root():
    import torch
    from diffrl_b200.engine import ArticulationEngine
    d, emu, q0, qd0, act = _case("CartPoleSwingUpEnv", 6)
    S, mm, dt = int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    eq, eqd, etape, _ = emu.forward(q0.ravel(), qd0.ravel(), act.ravel(), None, S, mm, dt)
    eng = ArticulationEngine(emu.desc, 6, "cuda:0")
    assert eng.C == 0
    t = lambda a: torch.tensor(a.ravel(), device="cuda:0")
    q, qd, _, _ = eng.forward(t(q0), t(qd0), t(act), None, S, mm, dt, want_tape=False)
    assert rel(q.cpu().numpy(), eq) < 1e-5 and rel(qd.cpu().numpy(), eqd) < 1e-5
