"""Bisect which part of an env rollout invalidates CUDA-graph capture (dev tool).
usage: python tools/debug_capture.py CartPoleSwingUpEnv"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffrl_b200.envs as envs

name = sys.argv[1]
mm = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}[name]
dev = torch.device("cuda:0")
env = getattr(envs, name)(num_envs=48, device="cuda:0", no_grad=False, MM_caching_frequency=mm, episode_length=9)
env.clear_grad(); env.reset(); env.initialize_trajectory()
n, a = env.num_envs, env.num_actions
q0, qd0 = env.state.joint_q.detach().clone(), env.state.joint_qd.detach().clone()
prog0, act0 = env.progress_buf.clone(), env.actions.detach().clone()
acts = torch.rand((3, n, a), device=dev, requires_grad=True)


def start():
    env.state = env.model.state()
    env.state.joint_q, env.state.joint_qd = q0.clone(), qd0.clone()
    env.progress_buf, env.actions = prog0.clone(), act0.clone()


def piece_obs():
    start(); env.calculateObservations()

def piece_apply():
    start(); env._apply_actions(torch.clip(acts[0], -1, 1))

def piece_sim():
    start(); env._apply_actions(torch.clip(acts[0], -1, 1))
    env.state = env.integrator.forward(env.model, env.state, env.sim_dt, env.sim_substeps, env.MM_caching_frequency)

def piece_reward():
    piece_sim(); env.progress_buf = env.progress_buf + 1; env.reset_buf = torch.zeros_like(env.reset_buf)
    env.actions = torch.clip(acts[0], -1, 1)
    env.calculateObservations(); env.calculateReward()

def piece_reset():
    piece_reward(); env._reset_masked(env.reset_buf)

def piece_step():
    start(); env.calculateObservations(); env.step(acts[0])

def piece_step_bwd():
    start(); env.calculateObservations()
    obs, rew, done, _ = env.step(acts[0])
    acts.grad = None
    rew.sum().backward()

for label, fn in [("observations", piece_obs), ("apply actions", piece_apply), ("simulation step", piece_sim),
                  ("obs + reward", piece_reward), ("masked reset", piece_reset), ("env.step", piece_step),
                  ("env.step + backward", piece_step_bwd)]:
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("capture ok   :", label, flush=True)
    except Exception as e:
        print("capture FAILS:", label, "--", repr(e)[:160].replace("\n", " "), flush=True)
        torch.cuda.synchronize()

# ---- multi-step variants through GraphedRollout itself
from diffrl_b200.rollout import GraphedRollout
for T in (1, 2, 3, 6):
    for ep_len in (9, 1000):
        try:
            e2 = getattr(envs, name)(num_envs=48, device="cuda:0", no_grad=False, MM_caching_frequency=mm, episode_length=ep_len)
            e2.clear_grad(); e2.reset(); e2.initialize_trajectory()
            GraphedRollout(e2, T)
            print("capture ok   : GraphedRollout T=%d episode_length=%d" % (T, ep_len), flush=True)
        except Exception as e:
            print("capture FAILS: GraphedRollout T=%d episode_length=%d -- %s" % (T, ep_len, repr(e)[:120].replace("\n", " ")), flush=True)
            torch.cuda.synchronize()
