"""End-to-end through the reference-facing API (``envs.*Env.step`` -> ``dflex.sim.SemiImplicitIntegrator``
-> C ABI -> CUDA) against rollouts recorded from the unmodified reference (oracle/make_golden.py):
observations, rewards, resets and d(sum rewards)/d(actions) for 3 steps from reset."""
import numpy as np
import pytest

from conftest import ENVS
from emu_util import load_golden
from tolerances import GRAD_RTOL, fwd_rtol

pytestmark = pytest.mark.gpu

MM = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


@pytest.mark.parametrize("name", ENVS)
def test_env_rollout_matches_reference(name):
    import torch
    import diffrl_b200.envs as envs
    d, _ = load_golden(name)
    n = int(d["meta/num_envs"])
    env = getattr(envs, name)(num_envs=n, device="cuda:0", render=False, seed=0, stochastic_init=False, no_grad=False,
                              MM_caching_frequency=MM[name])
    env.clear_grad()
    env.reset()
    obs0 = env.initialize_trajectory()
    assert rel(obs0.detach().cpu().numpy(), d["rollout/obs0"]) < 1e-6
    actions = [torch.tensor(a, device="cuda:0", requires_grad=True) for a in d["rollout/actions"]]
    loss = 0.0
    tol = 5.0 * fwd_rtol(name)   # three env-steps chained
    for t, a in enumerate(actions):
        obs, rew, done, extras = env.step(a)
        assert rel(obs.detach().cpu().numpy(), d["rollout/obs"][t]) < tol, (name, t)
        assert np.abs(rew.detach().cpu().numpy() - d["rollout/rew"][t]).max() < tol * (1.0 + np.abs(d["rollout/rew"][t]).max()), (name, t)
        assert np.array_equal(done.cpu().numpy(), d["rollout/done"][t]), (name, t)
        assert "obs_before_reset" in extras and "episode_end" in extras
        loss = loss + rew.sum()
    loss.backward()
    got = np.stack([a.grad.cpu().numpy() for a in actions])
    assert rel(got, d["rollout/grad_actions"]) < 10 * GRAD_RTOL, name
    assert rel(env.state.joint_q.detach().cpu().numpy(), d["rollout/final_q"]) < tol


def test_no_grad_mode_updates_in_place():
    import torch
    import diffrl_b200.dflex_api as df
    import diffrl_b200.envs as envs
    env = envs.AntEnv(num_envs=8, device="cuda:0", no_grad=True, MM_caching_frequency=16)
    env.reset()
    q_before = env.state.joint_q.clone()
    state_obj = env.state
    a = torch.rand(8, 8, device="cuda:0") * 2 - 1
    env.step(a)
    assert env.state is state_obj and not torch.equal(q_before, env.state.joint_q)
    # same actions in autograd mode give the same trajectory
    env2 = envs.AntEnv(num_envs=8, device="cuda:0", no_grad=False, MM_caching_frequency=16)
    env2.clear_grad(); env2.reset(); env2.initialize_trajectory()
    env2.step(a)
    assert torch.allclose(env.state.joint_q, env2.state.joint_q.detach(), rtol=0, atol=0)
    df.config.no_grad = False


def test_derived_state_fields_on_demand():
    import torch
    import diffrl_b200.envs as envs
    d, _ = load_golden("AntEnv")
    env = envs.AntEnv(num_envs=2, device="cuda:0", no_grad=False, MM_caching_frequency=16)
    env.clear_grad(); env.reset(); env.initialize_trajectory()
    env.step(torch.zeros(2, 8, device="cuda:0"))
    X = env.state.body_X_sc
    assert X.shape == (18, 7) and torch.isfinite(X).all()


@pytest.mark.parametrize("name", ["AntEnv", "HumanoidEnv", "SNUHumanoidEnv"])
def test_fused_epilogue_equals_torch_ops(name):
    """The fused obs/reward/reset kernel and its adjoint against the plain PyTorch formulation."""
    import torch
    import diffrl_b200.envs as envs
    n = 64
    outs = []
    for fused in (True, False):
        torch.manual_seed(0)
        env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=MM[name])
        env.fused_epilogue = fused
        env.sync_free_reset = fused
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        g = torch.Generator(device="cuda:0").manual_seed(3)
        acts = [(torch.rand((n, env.num_actions), generator=g, device="cuda:0") * 2 - 1).requires_grad_() for _ in range(4)]
        loss, obs_l, rew_l = 0.0, [], []
        w = torch.linspace(0.5, 1.5, env.num_obs, device="cuda:0")
        for a in acts:
            obs, rew, done, _ = env.step(a)
            loss = loss + rew.sum() + (obs * w).sum() * 1e-2
            obs_l.append(obs.detach()); rew_l.append(rew.detach())
        loss.backward()
        outs.append((torch.stack(obs_l), torch.stack(rew_l), torch.stack([a.grad for a in acts])))
    (o1, r1, g1), (o2, r2, g2) = outs
    assert torch.allclose(o1, o2, rtol=1e-5, atol=1e-5), float((o1 - o2).abs().max())
    assert torch.allclose(r1, r2, rtol=1e-5, atol=1e-4), float((r1 - r2).abs().max())
    assert (g1 - g2).abs().max() <= 1e-4 * g2.abs().max() + 1e-6


@pytest.mark.parametrize("name", ENVS)
def test_fused_transition_equals_unfused_step(name):
    """env.step() as three launches (action map, simulation step, transition: dfx_action_map_* /
    dfx_walker_transition_* / dfx_planar_transition_*) against the op-by-op PyTorch step, all six envs: observations before and after the masked
    reset, rewards, flags, counters, next state, and the gradient of a loss through all of them -- with
    actions beyond the clip range and episodes short enough to terminate inside the window."""
    import torch
    import diffrl_b200.envs as envs
    n, T = 48, 7
    outs = []
    for mode in ("single", "fused", "epilogue", "torch"):
        torch.manual_seed(0)
        env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=MM[name], episode_length=3)
        env.single_launch_step = mode == "single"       # the transition as the epilogue of the simulation launch (dfx_env_step_*)
        env.fused_transition = mode in ("single", "fused")
        env.fused_epilogue = mode != "torch"
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        g = torch.Generator(device="cuda:0").manual_seed(11)
        acts = [((torch.rand((n, env.num_actions), generator=g, device="cuda:0") * 2 - 1) * 1.6).requires_grad_() for _ in range(T)]
        w = torch.linspace(0.5, 1.5, env.num_obs, device="cuda:0")
        loss, rec = 0.0, []
        for a in acts:
            obs, rew, done, extras = env.step(a)
            loss = loss + rew.sum() + (obs * w).sum() * 1e-2 + (extras["obs_before_reset"] * w).sum() * 3e-3
            rec.append((obs.detach().clone(), rew.detach().clone(), done.clone(), env.progress_buf.clone(),
                        env.state.joint_q.detach().clone(), env.state.joint_qd.detach().clone(), env.actions.detach().clone(),
                        extras["obs_before_reset"].detach().clone()))
        loss.backward()
        outs.append((rec, torch.stack([a.grad for a in acts]), float(loss.detach())))
    ref_rec, ref_grad, ref_loss = outs[3]
    assert any(bool(r[2].any()) for r in ref_rec), "the window must contain terminations"
    for rec, grad, loss in outs[:3]:
        for a, b in zip(rec, ref_rec):
            assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
            for x, y in zip(a[:2] + a[4:], b[:2] + b[4:]):
                assert torch.allclose(x, y, rtol=1e-5, atol=1e-4), float((x - y).abs().max())
        assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-3
        assert (grad - ref_grad).abs().max() <= 1e-4 * ref_grad.abs().max() + 1e-6


@pytest.mark.parametrize("name", ["AntEnv", "HopperEnv", "CheetahEnv", "CartPoleSwingUpEnv"])
def test_masked_reset_equals_indexed_reset(name):
    """Terminated environments re-initialised by mask (no host sync, graph-capturable) == the reference's
    reset_buf.nonzero() + indexed writes."""
    import torch
    import diffrl_b200.envs as envs
    n = 32
    res = []
    for masked in (True, False):
        env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=MM[name], episode_length=3)
        env.sync_free_reset = masked
        env.fused_transition = False
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        g = torch.Generator(device="cuda:0").manual_seed(5)
        traj = []
        for t in range(5):     # episode_length 3 forces resets at t = 2
            obs, rew, done, _ = env.step(torch.rand((n, env.num_actions), generator=g, device="cuda:0") * 2 - 1)
            traj.append((obs.detach().clone(), done.clone(), env.progress_buf.clone(), env.state.joint_q.detach().clone()))
        res.append(traj)
    assert any(bool(a[1].any()) for a in res[0])
    for a, b in zip(*res):
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
        assert torch.allclose(a[0], b[0], atol=1e-6) and torch.allclose(a[3], b[3], atol=1e-6)


@pytest.mark.parametrize("name", ["AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "HopperEnv", "CheetahEnv", "CartPoleSwingUpEnv"])
def test_graphed_rollout_equals_eager(name):
    """One CUDA graph for horizon env-steps + backward reproduces the eager rollout (loss, action gradients,
    final state) and chains across calls."""
    import torch
    import diffrl_b200.envs as envs
    from diffrl_b200.rollout import GraphedRollout
    n, T = 48, 6
    g = torch.Generator().manual_seed(2)
    num_act = {"AntEnv": 8, "HumanoidEnv": 21, "SNUHumanoidEnv": 152, "HopperEnv": 3, "CheetahEnv": 6, "CartPoleSwingUpEnv": 1}[name]
    acts = [torch.rand((T, n, num_act), generator=g) * 2 - 1 for _ in range(2)]

    def make():
        env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=MM[name], episode_length=9)
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        return env

    env = make()
    eager = []
    for a in acts:                      # two chained windows, eager
        env.initialize_trajectory()
        a_dev = a.to("cuda:0").requires_grad_()
        loss = 0.0
        for t in range(T):
            obs, rew, done, _ = env.step(a_dev[t])
            loss = loss + rew.sum()
        loss.backward()
        eager.append((float(loss), a_dev.grad.cpu(), env.state.joint_q.detach().cpu().clone()))
    env2 = make()
    roll = GraphedRollout(env2, T)       # (leaves the env where it was: the first roll() starts from the reset state)
    for k, a in enumerate(acts):
        loss, grad = roll(a)
        assert abs(float(loss) - eager[k][0]) <= 1e-4 * abs(eager[k][0]) + 1e-3
        assert (grad - eager[k][1]).abs().max() <= 2e-4 * eager[k][1].abs().max() + 1e-6
        assert torch.allclose(env2.state.joint_q.cpu(), eager[k][2], rtol=1e-5, atol=1e-5)


def test_humanoid_invalid_state_gives_zero_reward():
    """reference envs/humanoid.py:359-369: an environment whose state went NaN / Inf / > 1e6 is reset AND its reward is
    zeroed, so that sum(rew).backward() stays finite -- fused transition and op-by-op path alike."""
    import torch
    import diffrl_b200.envs as envs
    for fused in (True, False):
        env = envs.HumanoidEnv(num_envs=8, device="cuda:0", no_grad=False, MM_caching_frequency=48)
        env.fused_transition = fused
        env.fused_epilogue = fused
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        with torch.no_grad():
            q = env.state.joint_q.clone()
            q.view(8, -1)[3, 9] = float("nan")         # poison one joint angle of environment 3
            env.state.joint_q = q
        a = torch.zeros((8, env.num_actions), device="cuda:0", requires_grad=True)
        obs, rew, done, _ = env.step(a)
        assert bool(done[3]) and float(rew[3]) == 0.0, (fused, rew)
        assert torch.isfinite(rew).all()
        rew.sum().backward()
        assert torch.isfinite(a.grad[torch.arange(8) != 3]).all()


@pytest.mark.parametrize("name", ENVS)
def test_action_map_folded_into_the_step_equals_separate_launch(name):
    """env.step() with the action map inside the simulation launch (dfx_step_forward_mapped / _backward_mapped: 2 launches
    forward + 2 backward) against the 3 + 3 launch path: same arithmetic, so observations / rewards / states are
    bit-identical and the action gradients agree to rounding of the cotangent accumulation -- actions beyond the clip
    range included (zero gradient there)."""
    import torch
    import diffrl_b200.envs as envs
    from diffrl_b200 import _capi
    n, T = 40, 5
    outs = []
    for folded in (True, False):
        torch.manual_seed(0)
        env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=MM[name], episode_length=4)
        env.fused_action_map = folded
        env.single_launch_step = False
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        g = torch.Generator(device="cuda:0").manual_seed(13)
        acts = [((torch.rand((n, env.num_actions), generator=g, device="cuda:0") * 2 - 1) * 1.5).requires_grad_() for _ in range(T)]
        l0 = _capi.lib().dfx_launch_count()
        loss, rec = 0.0, []
        for a in acts:
            obs, rew, done, _ = env.step(a)
            loss = loss + rew.sum() + obs.sum() * 1e-2
            rec.append((obs.detach().clone(), rew.detach().clone(), done.clone(), env.state.joint_q.detach().clone(), env.actions.detach().clone()))
        launches_fwd = _capi.lib().dfx_launch_count() - l0
        loss.backward()
        outs.append((rec, torch.stack([a.grad for a in acts]), launches_fwd))
    (r1, g1, l1), (r2, g2, l2) = outs
    assert l1 == 2 * T and l2 == 3 * T, (l1, l2)
    for a, b in zip(r1, r2):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert (g1 - g2).abs().max() <= 1e-6 * g2.abs().max() + 1e-9
    assert bool((g1[(torch.stack([a.detach() for a in acts]).abs() > 1.0)] == 0).all())


@pytest.mark.parametrize("name", ENVS)
@pytest.mark.parametrize("n", [40, 67])
def test_single_launch_step_equals_the_two_launch_step(name, n):
    """env.step() as ONE launch forward and ONE backward (dfx_env_step_forward / _backward: the transition as the epilogue of
    the simulation launch, its adjoint as the prologue of the adjoint launch) against the two-launch path (dfx_step_*_mapped +
    dfx_*_transition_*): the same per-environment code on the same values -- flags, counters and states exactly equal,
    observations / rewards / action gradients to a few ulp -- with terminations inside the window, actions beyond the clip range,
    batch sizes that are not a multiple of the tile width, the humanoids' gradient guard, and the launch count checked through
    dfx_launch_count()."""
    import torch
    import diffrl_b200.envs as envs
    from diffrl_b200 import _capi
    T = 6
    outs = []
    for single in (True, False):
        torch.manual_seed(0)
        env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=MM[name], episode_length=3)
        env.single_launch_step = single
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        g = torch.Generator(device="cuda:0").manual_seed(17)
        acts = [((torch.rand((n, env.num_actions), generator=g, device="cuda:0") * 2 - 1) * 1.5).requires_grad_() for _ in range(T)]
        w = torch.linspace(0.5, 1.5, env.num_obs, device="cuda:0")
        l0 = _capi.lib().dfx_launch_count()
        loss, rec = 0.0, []
        for a in acts:
            obs, rew, done, extras = env.step(a)
            loss = loss + rew.sum() + (obs * w).sum() * 1e-2 + (extras["obs_before_reset"] * w).sum() * 3e-3
            rec.append((obs.detach().clone(), rew.detach().clone(), done.clone(), env.progress_buf.clone(), env.state.joint_q.detach().clone(),
                        env.state.joint_qd.detach().clone(), env.actions.detach().clone(), extras["obs_before_reset"].detach().clone()))
        l1 = _capi.lib().dfx_launch_count()
        loss.backward()
        l2 = _capi.lib().dfx_launch_count()
        outs.append((rec, torch.stack([a.grad for a in acts]), l1 - l0, l2 - l1))
    (r1, g1, f1, b1), (r2, g2, f2, b2) = outs
    assert (f1, b1) == (T, T) and (f2, b2) == (2 * T, 2 * T), (f1, b1, f2, b2)
    assert any(bool(r[2].any()) for r in r2), "the window must contain terminations"
    names = ("obs", "rew", "done", "progress", "joint_q", "joint_qd", "actions", "obs_before_reset")
    for t, (a, b) in enumerate(zip(r1, r2)):
        for key, x, y in zip(names, a, b):
            if key in ("done", "progress", "joint_q", "joint_qd", "actions"):
                # the simulation step is the same kernel code on the same inputs; flags, counters and the (copied or re-initialised) next state are exact
                assert torch.equal(x, y), (t, key, float((x.float() - y.float()).abs().max()))
            else:
                # the transition arithmetic is the same source compiled into two kernels: FMA contraction of a*b - c*d may differ -> a few ulp
                assert torch.allclose(x, y, rtol=2e-6, atol=2e-6), (t, key, float((x - y).abs().max()))
    # (gradients: the contact adjoint's float atomics are not order-deterministic from run to run, and ulp differences of the
    #  transition cotangents pass through the ill-conditioned humanoid steps)
    assert (g1 - g2).abs().max() <= 1e-4 * g2.abs().max() + 1e-6, float((g1 - g2).abs().max() / g2.abs().max())
