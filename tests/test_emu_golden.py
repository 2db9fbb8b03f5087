"""CPU-side check of the device code's arithmetic: the host emulation (tests/host_emu, the same
phase/step headers the CUDA kernels compile) against golden vectors produced by the unmodified
reference (oracle/make_golden.py).  Forward tolerance is the north-star's 1e-5 relative."""
import numpy as np
import pytest

from conftest import ENVS
from emu_util import EmuSim, load_golden

from tolerances import FWD_RTOL, GRAD_RTOL, fwd_rtol


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


@pytest.mark.parametrize("variant", ["level", "tile"])
@pytest.mark.parametrize("name", ENVS)
def test_forward_and_adjoint_match_reference(name, variant):
    """variant "level": the lane-group kernels' formulation (contiguous scratch, level-by-level tree recursions);
    "tile": the tile kernels' (strided scratch, path / subtree passes)."""
    d, model = load_golden(name)
    N, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    sim = EmuSim(model, N) if variant == "level" else EmuSim(model, N, es=3, path_passes=True)
    for k in range(int(d["meta/num_cases"])):
        p = "case%d/" % k
        musc = d[p + "musc"] if (p + "musc") in d.files else None
        q, qd, tape, _ = sim.forward(d[p + "q0"], d[p + "qd0"], d[p + "act"], musc, S, mm, dt)
        assert rel(q, d[p + "traj_q"][-1]) < fwd_rtol(name), (name, k)
        assert rel(qd, d[p + "traj_qd"][-1]) < fwd_rtol(name), (name, k)
        # the tape holds the state entering every substep == the reference's per-substep trajectory
        QD = sim.lib.emu_pack_query(sim.pack, 8)      # DFX_QUERY_TAPE_ROW_FLOATS; a row starts with (q, qd)
        t = tape[: S * N * QD].reshape(S, N, QD)
        # (velocities right after a start from rest are ~0, and the fp32 solve of H q'' = tau is only
        #  good to cond(H) * eps ~ 1e-5 relative in the reference too -- hence the absolute floor)
        for s in (1, S // 2, S - 1):
            assert rel(t[s, :, : sim.desc.Q], d[p + "traj_q"][s - 1]) < fwd_rtol(name)
            ref_qd = d[p + "traj_qd"][s - 1]
            assert np.abs(t[s, :, sim.desc.Q:sim.desc.Q + sim.desc.D].ravel() - ref_qd).max() < fwd_rtol(name) * (1.0 + np.abs(ref_qd).max())
        gq, gqd, gact, gm = sim.backward(d[p + "act"], musc, tape, d[p + "gq_out"], d[p + "gqd_out"], S, mm, dt)
        assert rel(gq, d[p + "grad_q"]) < GRAD_RTOL, (name, k)
        assert rel(gqd, d[p + "grad_qd"]) < GRAD_RTOL, (name, k)
        assert rel(gact, d[p + "grad_act"]) < GRAD_RTOL, (name, k)
        if gm is not None:
            assert rel(gm, d[p + "grad_musc"]) < GRAD_RTOL, (name, k)


@pytest.mark.parametrize("name", ENVS)
def test_single_substep_derived_state(name):
    """Every intermediate of the first substep against the reference State tensors."""
    d, model = load_golden(name)
    N, S, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), float(d["meta/dt"])
    sim = EmuSim(model, N)
    p = "case%d/" % (int(d["meta/num_cases"]) - 1)
    musc = d[p + "musc"] if (p + "musc") in d.files else None
    _, _, _, dv = sim.forward(d[p + "q0"], d[p + "qd0"], d[p + "act"], musc, 1, 1, dt / S, tape=False, derived=True)
    for f in ("body_X_sc", "body_X_sm", "joint_S_s", "body_v_s", "body_a_s", "body_f_s", "body_ft_s", "joint_tau", "H", "L"):
        assert rel(dv[f], d[p + "first/" + f]) < 2e-5, (name, f)
    assert rel(dv["joint_qdd"], d[p + "first/joint_qdd"]) < 2e-4, name   # conditioned by H


def test_heterogeneous_batch_is_rejected():
    from diffrl_b200.modelpack import articulation_from_model
    d, model = load_golden("AntEnv")
    model = dict(model)
    model["joint_axis"] = model["joint_axis"].copy()
    model["joint_axis"][-1, 0] += 0.5
    with pytest.raises(ValueError):
        articulation_from_model(model, 2)


@pytest.mark.parametrize("name", ["HumanoidEnv", "SNUHumanoidEnv"])
def test_reference_solve_is_conditioning_limited(name):
    """Justifies tests/tolerances.py: the reference's own fp32 q'' deviates from an fp64 solve of its own
    H, tau by more than 1e-6 relative because cond(H + armature) is ~1e3..1e4."""
    d, model = load_golden(name)
    N = int(d["meta/num_envs"])
    D = int(d["meta/joint_dof_count"]) // N
    worst, cond = 0.0, 0.0
    for k in range(int(d["meta/num_cases"])):
        p = "case%d/" % k
        H = d[p + "first/H"].reshape(N, D, D).astype(np.float64)
        tau = d[p + "first/joint_tau"].reshape(N, D).astype(np.float64)
        arm = model["joint_armature"].reshape(N, D).astype(np.float64)
        ref = d[p + "first/joint_qdd"].reshape(N, D)
        for e in range(N):
            A = H[e] + np.diag(arm[e])
            A = 0.5 * (A + A.T)
            x = np.linalg.solve(A, tau[e])
            worst = max(worst, np.abs(ref[e] - x).max() / np.abs(x).max())
            cond = max(cond, np.linalg.cond(A))
    assert cond > 1e3 and worst > 1e-6, (cond, worst)


@pytest.mark.parametrize("name", ["AntEnv", "SNUHumanoidEnv", "CartPoleSwingUpEnv"])
def test_strided_scratch_is_bit_identical(name):
    """The tile kernels address the per-environment scratch with an element stride (structure-of-arrays over 32
    environments, csrc/dfx_math.h DFX_ES).  Built here with stride 3 on the CPU: every result must be bit-identical
    to the contiguous build -- an access that forgot the stride would read a neighbour's slot."""
    d, model = load_golden(name)
    n0 = int(d["meta/num_envs"])
    S, mm, dt = int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    p = "case0/"
    res = []
    for es in (1, 3):
        emu = EmuSim(model, n0, es=es)
        musc = d[p + "musc"] if emu.desc.M else None
        q, qd, tape, _ = emu.forward(d[p + "q0"], d[p + "qd0"], d[p + "act"], musc, S, mm, dt)
        g = emu.backward(d[p + "act"], musc, tape, d[p + "gq_out"], d[p + "gqd_out"], S, mm, dt)
        res.append([q, qd, tape] + [x for x in g if x is not None])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["HumanoidEnv", "SNUHumanoidEnv"])
def test_explicit_inverse_is_not_the_parity_floor(name):
    """is the explicit H^-1 (one mat-vec per substep) what keeps the two ill-conditioned models from the
    north-star's 1e-5?  A/B on the host emulation: q'' by the reference's two triangular sweeps with L (matnn.h:188-230)
    vs the H^-1 mat-vec, each with and without FMA contraction (what nvcc does on the GPU).  Both formulations are within
    1e-5 of the reference's final state without contraction and agree with each other to a few 1e-6; contraction, not the
    inverse, is what moves SNU to ~1.2e-5 on the GPU (tests/tolerances.py)."""
    d, model = load_golden(name)
    N, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    err = {}
    for tag, extra in (("hinv", ()), ("sweeps", ("-DDFX_EMU_SWEEPS",)), ("hinv_fma", ("-ffp-contract=fast", "-mfma")),
                       ("sweeps_fma", ("-DDFX_EMU_SWEEPS", "-ffp-contract=fast", "-mfma"))):
        sim = EmuSim(model, N, extra=extra)
        worst = 0.0
        for k in range(int(d["meta/num_cases"])):
            p = "case%d/" % k
            musc = d[p + "musc"] if (p + "musc") in d.files else None
            q, qd, _, _ = sim.forward(d[p + "q0"], d[p + "qd0"], d[p + "act"], musc, S, mm, dt, tape=False)
            worst = max(worst, rel(q, d[p + "traj_q"][-1]), rel(qd, d[p + "traj_qd"][-1]))
        err[tag] = worst
    assert err["hinv"] < 1e-5 and err["sweeps"] < 1e-5, err
    assert abs(err["hinv"] - err["sweeps"]) < 3e-6, err
    assert err["hinv_fma"] < 2e-5 and err["sweeps_fma"] < 2e-5, err


@pytest.mark.parametrize("flag", ["-DDFX_KIN_ADJ_CHAINS=1", "-DDFX_CRBA_ADJ_DIRECT=1"])
@pytest.mark.parametrize("name", ["AntEnv", "HumanoidEnv", "SNUHumanoidEnv"])
def test_restructured_adjoints_match_the_direct_formulations(name, flag):
    """Two adjoint phases are algebraic restructurings of what the reference's reversed tape performs (csrc/dfx_phases.h):
      * kin_adj accumulates the (v, a) and X_sc cotangents towards the root as subtree sums, quaternion parts in the world form,
        instead of the leaf -> root recursion (adjoint of sim.py:1668 / 1700-1720)             [A/B: -DDFX_KIN_ADJ_CHAINS=1]
      * crba_adj uses composite inertias and one moment matrix per link instead of the double sums over ancestor dofs
        (adjoint of eval_crba / eval_dense_gemm of the reference)                               [A/B: -DDFX_CRBA_ADJ_DIRECT=1]
    A/B on the host emulation, including states whose root quaternion is NOT unit (the world form is exact for those too):
    both formulations against the reference's gradients and against each other."""
    d, model = load_golden(name)
    N, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    new, direct = EmuSim(model, N), EmuSim(model, N, extra=(flag,))
    for k in range(int(d["meta/num_cases"])):
        p = "case%d/" % k
        musc = d[p + "musc"] if (p + "musc") in d.files else None
        for scale in (None, 1.07):
            q0 = d[p + "q0"].copy()
            if scale is not None:       # every free root: quaternion scaled away from unit length
                q0.reshape(N, -1)[:, 3:7] *= scale
            got = []
            for sim in (new, direct):
                _, _, tape, _ = sim.forward(q0, d[p + "qd0"], d[p + "act"], musc, S, mm, dt)
                got.append(sim.backward(d[p + "act"], musc, tape, d[p + "gq_out"], d[p + "gqd_out"], S, mm, dt))
            for a, b, key in zip(got[0], got[1], ("grad_q", "grad_qd", "grad_act", "grad_musc")):
                if a is None:
                    continue
                assert rel(a, b) < 6e-5, (name, k, scale, key, rel(a, b))     # (each is within GRAD_RTOL = 5e-5 of the reference)
                if scale is None:
                    assert rel(a, d[p + key]) < GRAD_RTOL and rel(b, d[p + key]) < GRAD_RTOL, (name, k, key)


@pytest.mark.parametrize("name", ["AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "CartPoleSwingUpEnv", "HopperEnv"])
def test_two_column_cholesky_is_bit_identical_to_column_by_column(name):
    """chol_inverse factorises two columns per barrier (csrc/dfx_phases.h): every entry of L is the same expression accumulated
    in the same order as column by column [A/B: -DDFX_CHOL_UNBLOCKED=1], so the final state, the whole tape (rows and the
    H^-1 blocks) and the dumped factor agree bit for bit -- odd and even numbers of dofs (14, 27, 24, 2, 6).  With FMA contraction
    left to the host compiler the two code shapes may be contracted differently (g++ vectorises the three diagonal-block chains):
    there the results agree to rounding only."""
    d, model = load_golden(name)
    N, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    for fma in ((), ("-ffp-contract=fast", "-mfma")):
        blocked, plain = EmuSim(model, N, extra=fma), EmuSim(model, N, extra=("-DDFX_CHOL_UNBLOCKED=1",) + fma)
        p = "case0/"
        musc = d[p + "musc"] if (p + "musc") in d.files else None
        out = []
        for sim in (blocked, plain):
            q, qd, tape, dumps = sim.forward(d[p + "q0"], d[p + "qd0"], d[p + "act"], musc, S, mm, dt, derived=True)
            out.append((q, qd, dumps["L"], dumps["H"], tape))
        for a, b in zip(*out):
            if not fma:
                assert np.array_equal(a, b)
        for a, b in zip(out[0][:3], out[1][:3]):
            assert rel(a, b) < 2e-6, rel(a, b)


@pytest.mark.parametrize("name", ["AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "CartPoleSwingUpEnv", "CheetahEnv"])
def test_merged_phases_are_bit_identical(name):
    """The tile kernels merge two pairs of phases to save CTA-wide barriers (Grp::kFusedPhases, csrc/dfx_phases.h): the
    joint-local transforms of the next substep are formed by the thread that integrates the link, and the torque adjoint
    recomputes the ancestors' direct wrench cotangents along each link's root path instead of staging them and summing in a
    second pass.  The same operations in the same order [A/B: -DDFX_EMU_FUSED_PHASES=1, path-pass build]: final state, tape and
    all gradients agree bit for bit."""
    d, model = load_golden(name)
    N, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
    plain, merged = EmuSim(model, N, es=3, path_passes=True), EmuSim(model, N, es=3, path_passes=True, extra=("-DDFX_EMU_FUSED_PHASES=1",))
    for k in range(int(d["meta/num_cases"])):
        p = "case%d/" % k
        musc = d[p + "musc"] if (p + "musc") in d.files else None
        out = []
        for sim in (plain, merged):
            q, qd, tape, _ = sim.forward(d[p + "q0"], d[p + "qd0"], d[p + "act"], musc, S, mm, dt)
            g = sim.backward(d[p + "act"], musc, tape, d[p + "gq_out"], d[p + "gqd_out"], S, mm, dt)
            out.append([q, qd, tape] + [x for x in g if x is not None])
        for a, b in zip(*out):
            assert np.array_equal(a, b)
