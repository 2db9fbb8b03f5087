"""Drive the REFERENCE'S OWN compiled CPU kernels without the reference's Python sources.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/README.md).  ``oracle/_ref/kernels/kernels.so`` is
the extension the unmodified reference generated and compiled for itself in the build container
(``oracle/ref_import.py``; g++ -O2, the arithmetic of dflex/dflex/sim.py + *.h exactly as shipped).  It
travels to the GPU box, but ``/root/reference`` does not, so this module re-states -- in our own words --
only the *launch schedule* around those kernels:

  * one substep = the launch sequence of ``SemiImplicitIntegrator._simulate`` (sim.py:2316-2601): fk, id,
    [contacts], [muscles], tau, [jacobian, mass, gemm P=MJ, gemm H=J^T P, cholesky], solve, integrate;
  * the tape = ``Tape.launch / replay`` (adjoint.py:2123-2199): launches recorded, replayed in reverse
    with zero-initialised adjoints for tensors that require grad and empty tensors ("skip") otherwise.

Kernel entry points follow ``<name>_cpu_forward(dim, *inputs, *outputs)`` /
``<name>_cpu_backward(dim, *inputs, *outputs, *adj_inputs, *adj_outputs)`` (adjoint.py:1268-1291).
"""
import importlib.machinery
import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS_SO = os.path.join(_HERE, "_ref", "kernels", "kernels.so")
_kernels = None


def available():
    return os.path.isfile(KERNELS_SO)


def kernels():
    global _kernels
    if _kernels is None:
        loader = importlib.machinery.ExtensionFileLoader("kernels", KERNELS_SO)
        spec = importlib.util.spec_from_file_location("kernels", KERNELS_SO, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _kernels = mod
    return _kernels


def _empty():
    return torch.FloatTensor()


class RefModel:
    """Reference-layout Model tensors (CPU) for ``num_envs`` copies of one articulation asset."""

    def __init__(self, arrays, num_envs, ground=True, gravity=(0.0, -9.81, 0.0)):
        import sys
        sys.path.insert(0, os.path.dirname(_HERE))
        from diffrl_b200.dflex_api.model import model_from_articulation
        m = model_from_articulation(arrays, num_envs, "cpu", ground=ground, gravity=gravity)
        self.m = m
        n = num_envs
        L = m.link_count // n
        D = m.joint_dof_count // n
        self.n, self.L, self.D, self.Q = n, L, D, m.joint_coord_count // n
        i32 = dict(dtype=torch.int32)
        # batched-GEMM bookkeeping (model.py:1745-1770)
        m.articulation_J_start = torch.arange(0, n, **i32) * (6 * L * D)
        m.articulation_M_start = torch.arange(0, n, **i32) * (36 * L * L)
        m.articulation_H_start = torch.arange(0, n, **i32) * (D * D)
        m.articulation_M_rows = torch.full((n,), 6 * L, **i32)
        m.articulation_H_rows = torch.full((n,), D, **i32)
        m.articulation_J_rows = torch.full((n,), 6 * L, **i32)
        m.articulation_J_cols = torch.full((n,), D, **i32)
        m.J_size, m.M_size, m.H_size = n * 6 * L * D, n * 36 * L * L, n * D * D
        m.gravity = torch.tensor(gravity, dtype=torch.float32)


class RefTape:
    def __init__(self):
        self.launches, self.adj, self.diff = [], {}, set()

    def mark(self, *tensors):
        for t in tensors:
            self.diff.add(id(t))
        return tensors[0] if len(tensors) == 1 else tensors

    def launch(self, name, dim, inputs, outputs):
        if dim <= 0:
            return
        getattr(kernels(), name + "_cpu_forward")(dim, *inputs, *outputs)
        self.launches.append((name, dim, inputs, outputs))

    def _adj(self, t):
        if not torch.is_tensor(t):
            return type(t)()
        key = id(t)
        if key in self.adj:
            return self.adj[key]
        if t.dtype == torch.float32 and key in self.diff:
            self.adj[key] = torch.zeros_like(t)
            return self.adj[key]
        return _empty()

    def replay(self):
        for name, dim, inputs, outputs in reversed(self.launches):
            adj_in = [self._adj(t) for t in inputs]
            adj_out = [self._adj(t) for t in outputs]
            getattr(kernels(), name + "_cpu_backward")(dim, *inputs, *outputs, *adj_in, *adj_out)


def _new_state(rm, tape):
    m, f32 = rm.m, dict(dtype=torch.float32)
    s = {}
    s["joint_qdd"] = torch.empty_like(m.joint_qd)
    s["joint_tau"] = torch.empty_like(m.joint_qd)
    s["joint_S_s"] = torch.empty((m.joint_dof_count, 6), **f32)
    s["body_X_sc"] = torch.empty((m.link_count, 7), **f32)
    s["body_X_sm"] = torch.empty((m.link_count, 7), **f32)
    s["body_I_s"] = torch.empty((m.link_count, 6, 6), **f32)
    s["body_v_s"] = torch.empty((m.link_count, 6), **f32)
    s["body_a_s"] = torch.empty((m.link_count, 6), **f32)
    s["body_f_s"] = torch.zeros((m.link_count, 6), **f32)
    s["body_ft_s"] = torch.zeros((m.link_count, 6), **f32)
    s["joint_q"] = torch.empty_like(m.joint_q)
    s["joint_qd"] = torch.empty_like(m.joint_qd)
    tape.mark(*s.values())
    return s


def _substep(rm, tape, q, qd, act, musc, dt, update):
    m = rm.m
    s = _new_state(rm, tape)
    n = m.articulation_count
    tape.launch("eval_rigid_fk", n, [m.articulation_joint_start, m.joint_type, m.joint_parent, m.joint_q_start,
                                    m.joint_qd_start, q, m.joint_X_pj, m.joint_X_cm, m.joint_axis],
                [s["body_X_sc"], s["body_X_sm"]])
    tape.launch("eval_rigid_id", n, [m.articulation_joint_start, m.joint_type, m.joint_parent, m.joint_q_start,
                                    m.joint_qd_start, q, qd, m.joint_axis, m.joint_target_ke, m.joint_target_kd,
                                    m.body_I_m, s["body_X_sc"], s["body_X_sm"], m.joint_X_pj, m.gravity],
                [s["joint_S_s"], s["body_I_s"], s["body_v_s"], s["body_f_s"], s["body_a_s"]])
    if m.ground and m.contact_count > 0:
        tape.launch("eval_rigid_contacts_art", m.contact_count,
                    [s["body_X_sc"], s["body_v_s"], m.contact_body0, m.contact_point0, m.contact_dist,
                     m.contact_material, m.shape_materials], [s["body_f_s"]])
    if m.muscle_count:
        tape.launch("eval_muscles", m.muscle_count,
                    [s["body_X_sc"], s["body_v_s"], m.muscle_start, m.muscle_params, m.muscle_links, m.muscle_points, musc],
                    [s["body_f_s"]])
    tape.launch("eval_rigid_tau", n, [m.articulation_joint_start, m.joint_type, m.joint_parent, m.joint_q_start,
                                     m.joint_qd_start, q, qd, act, m.joint_target, m.joint_target_ke, m.joint_target_kd,
                                     m.joint_limit_lower, m.joint_limit_upper, m.joint_limit_ke, m.joint_limit_kd,
                                     m.joint_axis, s["joint_S_s"], s["body_f_s"]], [s["body_ft_s"], s["joint_tau"]])
    if update:
        f32 = dict(dtype=torch.float32)
        rm.M, rm.J = torch.zeros(m.M_size, **f32), torch.zeros(m.J_size, **f32)
        rm.P, rm.H, rm.Lc = torch.empty(m.J_size, **f32), torch.empty(m.H_size, **f32), torch.zeros(m.H_size, **f32)
        tape.mark(rm.M, rm.J, rm.P, rm.H, rm.Lc)
        tape.launch("eval_rigid_jacobian", n, [m.articulation_joint_start, m.articulation_J_start, m.joint_parent,
                                              m.joint_qd_start, s["joint_S_s"]], [rm.J])
        tape.launch("eval_rigid_mass", n, [m.articulation_joint_start, m.articulation_M_start, s["body_I_s"]], [rm.M])
        tape.launch("eval_dense_gemm_batched", n, [m.articulation_M_rows, m.articulation_J_cols, m.articulation_J_rows, 0, 0,
                                                  m.articulation_M_start, m.articulation_J_start, m.articulation_J_start,
                                                  rm.M, rm.J], [rm.P])
        tape.launch("eval_dense_gemm_batched", n, [m.articulation_J_cols, m.articulation_J_cols, m.articulation_J_rows, 1, 0,
                                                  m.articulation_J_start, m.articulation_J_start, m.articulation_H_start,
                                                  rm.J, rm.P], [rm.H])
        tape.launch("eval_dense_cholesky_batched", n, [m.articulation_H_start, m.articulation_H_rows, rm.H, m.joint_armature],
                    [rm.Lc])
    tmp = torch.zeros_like(s["joint_tau"])
    tape.launch("eval_dense_solve_batched", n, [m.articulation_dof_start, m.articulation_H_start, m.articulation_H_rows,
                                               rm.H, rm.Lc, s["joint_tau"], tmp], [s["joint_qdd"]])
    tape.launch("eval_rigid_integrate", m.link_count, [m.joint_type, m.joint_q_start, m.joint_qd_start, q, qd,
                                                      s["joint_qdd"], dt], [s["joint_q"], s["joint_qd"]])
    return s


def env_step(rm, q, qd, act, musc, dt, substeps, mm_freq, gq_out=None, gqd_out=None):
    """One env-step on the reference kernels.  Returns (q', qd', grads or None, per-substep trajectory)."""
    tape = RefTape()
    q, qd, act = q.clone(), qd.clone(), act.clone()
    tape.mark(q, qd, act)
    if musc is not None:
        musc = musc.clone()
        tape.mark(musc)
    else:
        musc = rm.m.muscle_activation
    cq, cqd = q, qd
    traj = []
    sub_dt = float(dt) / float(substeps)
    for i in range(substeps):
        s = _substep(rm, tape, cq, cqd, act, musc, sub_dt, (i % mm_freq) == 0)
        cq, cqd = s["joint_q"], s["joint_qd"]
        traj.append((cq, cqd))
    grads = None
    if gq_out is not None:
        tape.adj[id(cq)] = gq_out.clone().contiguous()
        tape.adj[id(cqd)] = gqd_out.clone().contiguous()
        tape.replay()
        grads = (tape._adj(q), tape._adj(qd), tape._adj(act), tape._adj(musc) if id(musc) in tape.diff else None)
    return cq, cqd, grads, traj


def time_env_steps(arrays, num_envs, substeps, mm_freq, dt, steps, seed=0, ground=True):
    """Wall-clock a differentiable rollout (forward ``steps`` env-steps, then their adjoints) on the
    reference kernels; returns (seconds_forward, seconds_backward)."""
    import time
    rm = RefModel(arrays, num_envs, ground=ground)
    g = torch.Generator().manual_seed(seed)
    q, qd = rm.m.joint_q.clone(), rm.m.joint_qd.clone()
    tapes = []
    t0 = time.perf_counter()
    for _ in range(steps):
        act = (torch.rand(qd.shape, generator=g) * 2 - 1) * 100.0
        musc = torch.rand(rm.m.muscle_count, generator=g) * 50.0 if rm.m.muscle_count else None
        tape = RefTape()
        qi, qdi = q.clone(), qd.clone()
        tape.mark(qi, qdi, act)
        if musc is not None:
            tape.mark(musc)
        cq, cqd = qi, qdi
        for i in range(substeps):
            s = _substep(rm, tape, cq, cqd, act, musc if musc is not None else rm.m.muscle_activation,
                         float(dt) / substeps, (i % mm_freq) == 0)
            cq, cqd = s["joint_q"], s["joint_qd"]
        tapes.append((tape, qi, qdi, cq, cqd))
        q, qd = cq.detach(), cqd.detach()
    t1 = time.perf_counter()
    gq, gqd = torch.ones_like(q), torch.ones_like(qd)
    for tape, qi, qdi, cq, cqd in reversed(tapes):
        tape.adj[id(cq)], tape.adj[id(cqd)] = gq, gqd
        tape.replay()
        gq, gqd = tape._adj(qi), tape._adj(qdi)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1
