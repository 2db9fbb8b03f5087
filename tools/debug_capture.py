"""Localise what invalidates the CUDA-graph capture of a CartPole rollout (dev tool): the test's sequence (an eager
rollout on one env, then GraphedRollout on a second one) in variations, and the capture error modes."""
import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffrl_b200.envs as envs
import diffrl_b200.rollout as rollout

name = sys.argv[1] if len(sys.argv) > 1 else "CartPoleSwingUpEnv"
mm = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}[name]
n, T = 48, 6


def make():
    env = getattr(envs, name)(num_envs=n, device="cuda:0", no_grad=False, MM_caching_frequency=mm, episode_length=9)
    env.clear_grad(); env.reset(); env.initialize_trajectory()
    return env


def eager_phase(env, windows=2):
    g = torch.Generator().manual_seed(2)
    for _ in range(windows):
        env.initialize_trajectory()
        a_dev = (torch.rand((T, n, env.num_actions), generator=g) * 2 - 1).to("cuda:0").requires_grad_()
        loss = 0.0
        for t in range(T):
            obs, rew, done, _ = env.step(a_dev[t])
            loss = loss + rew.sum()
        loss.backward()
        float(loss.detach()); a_dev.grad.cpu()


def try_capture(label, mode="global", eager=True, cleanup=False):
    try:
        if eager:
            e1 = make(); eager_phase(e1)
            if cleanup:
                del e1; gc.collect(); torch.cuda.synchronize(); torch.cuda.empty_cache()
        e2 = make()
        orig = torch.cuda.graph

        class G(orig):
            def __init__(self, *a, **k):
                k.setdefault("capture_error_mode", mode)
                super().__init__(*a, **k)
        torch.cuda.graph = G
        try:
            rollout.GraphedRollout(e2, T)
        finally:
            torch.cuda.graph = orig
        print("capture ok   :", label, flush=True)
    except Exception as e:
        print("capture FAILS:", label, "--", repr(e)[:140].replace("\n", " "), flush=True)
        torch.cuda.synchronize()


try_capture("no eager phase, global", eager=False)
try_capture("eager phase first, global")
try_capture("eager phase first, env deleted + gc + empty_cache, global", cleanup=True)
try_capture("eager phase first, thread_local", mode="thread_local")
try_capture("eager phase first, relaxed", mode="relaxed")
