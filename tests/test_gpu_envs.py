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
