"""Generate the golden fixtures under ``tests/golden/`` by RUNNING THE REFERENCE ITSELF.

TEST INFRASTRUCTURE.  Run in the build container only (needs ``/root/reference``):

    python oracle/make_golden.py [EnvName ...]

The reference's own tests pin nothing for the articulated-rigid-body step (SURVEY.md section 4:
``dflex/tests/*.py`` contain zero asserts), so parity is pinned on outputs of the unmodified
reference CPU path (serial loop => deterministic), imported through ``oracle/ref_import.py``.

Per env (2 environments each) one ``tests/golden/<env>.npz`` holds

* ``model/<field>``       every tensor of the finalized reference ``Model`` (dflex/dflex/model.py:1646-1879)
                          plus the static contact list of ``Model.collide`` (model.py:424-515);
* ``meta/*``              dt, substeps, mass_matrix_freq, per-env counts;
* ``case<k>/...``         integrator-level cases: inputs (q0, qd0, act[, musc]), the (q, qd) after
                          EVERY substep of one ``SemiImplicitIntegrator.forward`` (sim.py:2182), the
                          derived State of the first and the last substep (+ H, L of the first),
                          random cotangents on (q', qd') and the resulting input gradients;
* ``rollout/...``         env-level: actions -> obs / rew / reset_buf for a few ``env.step`` calls from
                          reset and d(sum rew)/d(actions) (envs/ant.py:156-190 and siblings).
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref_import  # noqa: E402

ref_import.install()
import torch  # noqa: E402
import dflex as df  # noqa: E402  (the reference package)

GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")
NUM_ENVS = 2
ALL_ENVS = ["CartPoleSwingUpEnv", "AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "HopperEnv", "CheetahEnv"]
WARM_STEPS = {  # random-action steps rolled before a case is recorded (gets bodies into ground contact)
    "CartPoleSwingUpEnv": [0, 7],
    "AntEnv": [0, 14, 30],
    "HumanoidEnv": [0, 10, 22],
    "SNUHumanoidEnv": [0, 8],
    "HopperEnv": [0, 12, 25],
    "CheetahEnv": [0, 12, 25],
}
STATE_FIELDS = ["joint_q", "joint_qd", "joint_qdd", "joint_tau", "joint_S_s", "body_X_sc", "body_X_sm",
                "body_I_s", "body_v_s", "body_a_s", "body_f_s", "body_ft_s"]


def _np(t):
    return t.detach().cpu().numpy().copy()


def dump_model(env, out):
    model = env.model
    for key, value in model.__dict__.items():
        if torch.is_tensor(value) and key not in ("M", "J", "P", "H", "L"):
            out["model/" + key] = _np(value)
    for key in ("particle_count", "joint_coord_count", "joint_dof_count", "link_count", "shape_count",
                "articulation_count", "muscle_count", "contact_count", "J_size", "M_size", "H_size"):
        out["meta/" + key] = np.int64(getattr(model, key))
    out["meta/ground"] = np.int64(bool(model.ground))
    out["meta/num_envs"] = np.int64(env.num_envs)
    out["meta/dt"] = np.float64(env.sim_dt)
    out["meta/substeps"] = np.int64(env.sim_substeps)
    out["meta/mass_matrix_freq"] = np.int64(env.MM_caching_frequency)
    out["meta/num_obs"] = np.int64(env.num_obs)
    out["meta/num_act"] = np.int64(env.num_actions)


class SubstepRecorder:
    """Wraps SemiImplicitIntegrator._simulate (sim.py:2225) to snapshot every substep."""

    def __init__(self, integrator, model):
        self.integrator, self.model = integrator, model
        self.q, self.qd, self.first, self.last = [], [], None, None

    def __enter__(self):
        inner = type(self.integrator)._simulate
        rec = self

        def wrapped(self_, tape, model, state_in, state_out, dt, update_mass_matrix=True):
            ret = inner(self_, tape, model, state_in, state_out, dt, update_mass_matrix)
            rec.q.append(_np(state_out.joint_q))
            rec.qd.append(_np(state_out.joint_qd))
            snap = {f: _np(getattr(state_out, f)) for f in STATE_FIELDS}
            if rec.first is None:
                snap["H"], snap["L"] = _np(model.H), _np(model.L)
                rec.first = snap
            rec.last = snap
            return ret

        self._orig = inner
        type(self.integrator)._simulate = wrapped
        return self

    def __exit__(self, *exc):
        type(self.integrator)._simulate = self._orig


def run_case(env, q0, qd0, act, musc, rng, out, prefix):
    model = env.model
    state = model.state()
    state.joint_q = q0.clone().requires_grad_()
    state.joint_qd = qd0.clone().requires_grad_()
    state.joint_act = act.clone().requires_grad_()
    if musc is not None:
        model.muscle_activation = musc.clone().requires_grad_()
    with SubstepRecorder(env.integrator, model) as rec:
        new = env.integrator.forward(model, state, env.sim_dt, env.sim_substeps, env.MM_caching_frequency)
    gq = torch.tensor(rng.standard_normal(new.joint_q.shape), dtype=torch.float32)
    gqd = torch.tensor(rng.standard_normal(new.joint_qd.shape), dtype=torch.float32)
    loss = (new.joint_q * gq).sum() + (new.joint_qd * gqd).sum()
    loss.backward()
    out[prefix + "q0"], out[prefix + "qd0"], out[prefix + "act"] = _np(q0), _np(qd0), _np(act)
    out[prefix + "traj_q"], out[prefix + "traj_qd"] = np.stack(rec.q), np.stack(rec.qd)
    for key, value in rec.first.items():
        out[prefix + "first/" + key] = value
    for key, value in rec.last.items():
        out[prefix + "last/" + key] = value
    out[prefix + "gq_out"], out[prefix + "gqd_out"] = _np(gq), _np(gqd)
    out[prefix + "grad_q"], out[prefix + "grad_qd"] = _np(state.joint_q.grad), _np(state.joint_qd.grad)
    out[prefix + "grad_act"] = _np(state.joint_act.grad)
    if musc is not None:
        out[prefix + "musc"] = _np(musc)
        out[prefix + "grad_musc"] = _np(model.muscle_activation.grad)
        model.muscle_activation = model.muscle_activation.detach()


def act_from_actions(env, name, actions):
    """The joint_act / muscle_activation each env's step() derives from clipped actions."""
    n = env.num_envs
    act = torch.zeros(env.model.joint_dof_count)
    musc = None
    a = torch.clip(actions, -1.0, 1.0)
    if name == "CartPoleSwingUpEnv":
        act.view(n, -1)[:, 0:1] = a * env.action_strength
    elif name == "AntEnv":
        act.view(n, -1)[:, 6:] = a * env.action_strength
    elif name == "HumanoidEnv":
        act.view(n, -1)[:, 6:] = a * env.motor_scale * env.motor_strengths
    elif name == "SNUHumanoidEnv":
        musc = (a * 0.5 + 0.5).view(-1) * env.muscle_strengths
    elif name in ("HopperEnv", "CheetahEnv"):
        act.view(n, -1)[:, 3:] = a * env.action_strength
    return act, musc


def make_env_golden(name):
    rng = np.random.default_rng(1234)
    torch.manual_seed(0)
    out = {}
    env = ref_import.make_env(name, NUM_ENVS)
    dump_model(env, out)
    n, na = env.num_envs, env.num_actions

    # ---- env-level rollout from reset -------------------------------------------------
    env.clear_grad()
    env.reset()
    obs0 = env.initialize_trajectory()
    steps = 3
    actions = [torch.tensor(rng.uniform(-1, 1, (n, na)), dtype=torch.float32, requires_grad=True)
               for _ in range(steps)]
    obs_l, rew_l, done_l = [], [], []
    loss = 0.0
    for a in actions:
        obs, rew, done, _ = env.step(a)
        obs_l.append(_np(obs)); rew_l.append(_np(rew)); done_l.append(_np(done))
        loss = loss + rew.sum()
    loss.backward()
    out["rollout/obs0"] = _np(obs0)
    out["rollout/actions"] = np.stack([_np(a) for a in actions])
    out["rollout/obs"], out["rollout/rew"], out["rollout/done"] = np.stack(obs_l), np.stack(rew_l), np.stack(done_l)
    out["rollout/grad_actions"] = np.stack([_np(a.grad) for a in actions])
    out["rollout/final_q"], out["rollout/final_qd"] = _np(env.state.joint_q), _np(env.state.joint_qd)

    # ---- integrator-level cases -------------------------------------------------------
    env.clear_grad()
    env.reset()
    env.initialize_trajectory()
    rolled = 0
    for k, warm in enumerate(WARM_STEPS[name]):
        with torch.no_grad():
            pass
        while rolled < warm:
            a = torch.tensor(rng.uniform(-1, 1, (n, na)), dtype=torch.float32)
            env.step(a)
            env.clear_grad()
            rolled += 1
        q0 = env.state.joint_q.detach().clone()
        qd0 = env.state.joint_qd.detach().clone()
        a = torch.tensor(rng.uniform(-1, 1, (n, na)), dtype=torch.float32)
        act, musc = act_from_actions(env, name, a)
        run_case(env, q0, qd0, act, musc, rng, out, "case%d/" % k)
    out["meta/num_cases"] = np.int64(len(WARM_STEPS[name]))

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024.0), "cases", len(WARM_STEPS[name]))


if __name__ == "__main__":
    for env_name in (sys.argv[1:] or ALL_ENVS):
        make_env_golden(env_name)
