"""The drop-in claim on hardware: the reference's UNCHANGED ``envs/*.py`` (+ ``utils/load_utils.py``) and
``algorithms/shac.py`` run on THIS repo's ``dflex`` on ``cuda:0``.

The reference files come from the git-ignored install ``baseline/_ref`` (``oracle/install_reference.py``; it travels
to the GPU box, ``/root/reference`` does not).  Each case runs in its own interpreter with
``sys.path = [repo root (our dflex), oracle/refshim (gym / urdfpy / tensorboardX stand-ins), baseline/_ref]`` and is
compared with the rollouts recorded from the unmodified reference (``tests/golden/*.npz``, ``rollout/*``:
``/root/reference/envs/ant.py:165`` etc. -> ``dflex.sim.SemiImplicitIntegrator.forward`` -> ``dfx_step_forward``).
"""
import json
import os
import subprocess
import sys

import pytest

from tolerances import GRAD_RTOL, fwd_rtol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "envs")),
                                 reason="baseline/_ref not installed (python oracle/install_reference.py)")]

ENV_SCRIPT = r'''
import sys, json, numpy as np
np.Inf = np.inf
ROOT, REF, name = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path[:0] = [ROOT, ROOT + "/oracle/refshim", REF]
import torch
import dflex
assert dflex.__file__.startswith(ROOT + "/dflex"), dflex.__file__          # OUR dflex
import envs
assert envs.__file__.startswith(REF), envs.__file__                          # the REFERENCE's envs, unchanged
from diffrl_b200 import _capi
gold = np.load(ROOT + "/tests/golden/" + name + ".npz")
n = int(gold["meta/num_envs"])
mm = int(gold["meta/mass_matrix_freq"])
env = getattr(envs, name)(num_envs=n, device="cuda:0", render=False, seed=0, stochastic_init=False, no_grad=False,
                          MM_caching_frequency=mm)
l0 = _capi.lib().dfx_launch_count()
env.clear_grad(); env.reset()
obs0 = env.initialize_trajectory()
acts = [torch.tensor(a, device="cuda:0", requires_grad=True) for a in gold["rollout/actions"]]
obs, rew, done, loss = [], [], [], 0.0
for a in acts:
    o, r, d, _ = env.step(a)
    obs.append(o.detach().cpu().numpy()); rew.append(r.detach().cpu().numpy()); done.append(d.cpu().numpy())
    loss = loss + r.sum()
loss.backward()
def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
out = dict(obs0=rel(obs0.detach().cpu().numpy(), gold["rollout/obs0"]), obs=rel(np.stack(obs), gold["rollout/obs"]),
           rew=rel(np.stack(rew), gold["rollout/rew"]), done=bool(np.array_equal(np.stack(done), gold["rollout/done"])),
           grad=rel(np.stack([a.grad.cpu().numpy() for a in acts]), gold["rollout/grad_actions"]),
           final_q=rel(env.state.joint_q.detach().cpu().numpy(), gold["rollout/final_q"]),
           launches=int(_capi.lib().dfx_launch_count() - l0))
print("RESULT " + json.dumps(out))
'''

SHAC_SCRIPT = r'''
import sys, json, numpy as np, yaml
np.Inf = np.inf
ROOT, REF = sys.argv[1], sys.argv[2]
sys.path[:0] = [ROOT, ROOT + "/oracle/refshim", REF]
import torch
import torch_compat                                                         # oracle/refshim: torch >= 2 indexing compat, harness only
import dflex
assert dflex.__file__.startswith(ROOT + "/dflex"), dflex.__file__
import algorithms.shac as shac                                              # the reference's trainer, unchanged
assert shac.__file__.startswith(REF), shac.__file__
from diffrl_b200 import _capi
cfg = yaml.load(open(REF + "/examples/cfg/shac/ant.yaml"), Loader=yaml.SafeLoader)
cfg["params"]["general"] = dict(device="cuda:0", seed=0, render=False, logdir=sys.argv[3], train=True, checkpoint="Base", no_time_stamp=True)
cfg["params"]["config"]["max_epochs"] = 3
cfg["params"]["config"]["num_actors"] = 64
cfg["params"]["config"]["save_interval"] = 1000
cfg["params"]["diff_env"]["stochastic_env"] = True
l0 = _capi.lib().dfx_launch_count()
agent = shac.SHAC(cfg)
agent.train()
w = torch.cat([p.detach().flatten() for p in agent.actor.parameters()])
print("RESULT " + json.dumps(dict(finite=bool(torch.isfinite(w).all()), launches=int(_capi.lib().dfx_launch_count() - l0),
                                  epochs=int(agent.iter_count), steps=int(agent.step_count))))
'''


def _run(script, *argv, timeout=900):
    proc = subprocess.run([sys.executable, "-c", script, ROOT, REF] + list(argv), capture_output=True, text=True, timeout=timeout)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
    assert proc.returncode == 0 and lines, proc.stdout[-3000:] + proc.stderr[-3000:]
    return json.loads(lines[-1][len("RESULT "):])


@pytest.mark.parametrize("name", ["CartPoleSwingUpEnv", "AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "HopperEnv", "CheetahEnv"])
def test_unchanged_reference_env_steps_on_our_dflex(name):
    r = _run(ENV_SCRIPT, name)
    tol = fwd_rtol(name)
    assert r["launches"] >= 6, r                   # 3 forward + 3 adjoint launches of OUR kernels
    assert r["obs0"] < 1e-6 and r["done"], r
    assert r["obs"] < 10 * tol and r["rew"] < 10 * tol and r["final_q"] < 10 * tol, r      # 3 env-steps = 48..144 substeps
    assert r["grad"] < 10 * GRAD_RTOL, r


def test_unchanged_reference_shac_trains_on_our_dflex(tmp_path):
    r = _run(SHAC_SCRIPT, str(tmp_path))
    assert r["finite"] and r["epochs"] == 3 and r["launches"] >= 3 * 32 * 2, r
