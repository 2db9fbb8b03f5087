"""Find the call that invalidates CUDA-graph capture of an env rollout: run the body of GraphedRollout eagerly with
torch's sync-debug mode set to "error" (dev tool).  usage: python tools/debug_capture.py CartPoleSwingUpEnv"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffrl_b200.envs as envs
from diffrl_b200.rollout import GraphedRollout

name = sys.argv[1]
mm = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}[name]
env = getattr(envs, name)(num_envs=48, device="cuda:0", no_grad=False, MM_caching_frequency=mm, episode_length=9)
env.clear_grad(); env.reset(); env.initialize_trajectory()
roll = GraphedRollout.__new__(GraphedRollout)
# replicate __init__ without the capture
roll.env, roll.T, roll.device, roll.weight, roll.graph = env, 6, torch.device("cuda:0"), None, None
n, a = env.num_envs, env.num_actions
roll.actions = torch.zeros((6, n, a), device="cuda:0", requires_grad=True)
roll.q0, roll.qd0 = env.state.joint_q.detach().clone(), env.state.joint_qd.detach().clone()
roll.progress0, roll.prev_actions = env.progress_buf.clone(), env.actions.detach().clone()
roll.host_grad = torch.empty((6, n, a)).pin_memory(); roll.host_loss = torch.empty(()).pin_memory()
roll.host_actions = torch.rand((6, n, a)).pin_memory()
roll._body(); roll._body()
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("error")
try:
    roll._body()
    print("no synchronising call found in the body")
except Exception:
    traceback.print_exc()
torch.cuda.set_sync_debug_mode("default")
torch.cuda.synchronize()
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        roll._body()
    torch.cuda.current_stream().wait_stream(s)
    roll.actions.grad = None
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        roll._body()
    print("capture ok with capture_error_mode=thread_local")
except Exception as e:
    print("capture failed (thread_local):", repr(e)[:300])
