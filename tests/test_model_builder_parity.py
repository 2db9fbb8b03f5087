"""The reference's UNCHANGED envs + asset loaders (imported from /root/reference, build container only)
run on OUR ``dflex`` package up to model construction on CPU; every finalized Model tensor must equal
the reference's bit for bit (they are the data ABI of the kernels).  Skipped where the reference tree
is absent (GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DIFFRL_REFERENCE", "/root/reference")

SCRIPT = r'''
import sys, numpy as np
np.Inf = np.inf
sys.path[:0] = [%(root)r, %(root)r + "/oracle/refshim", %(ref)r]
import torch
import dflex
assert "diffrl_b200" in dflex.__file__ or dflex.__file__.startswith(%(root)r), dflex.__file__
import envs
name = sys.argv[1]
mm = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}[name]
env = getattr(envs, name)(num_envs=2, device="cpu", render=False, seed=0, stochastic_init=False, no_grad=False, MM_caching_frequency=mm)
gold = np.load(%(root)r + "/tests/golden/" + name + ".npz")
bad = []
for key in gold.files:
    if not key.startswith("model/"):
        continue
    field = key[len("model/"):]
    ours = getattr(env.model, field, None)
    if ours is None:
        bad.append(field + ": missing"); continue
    ours = ours.detach().cpu().numpy()
    ref = gold[key]
    if ours.size == 0 and ref.size == 0:
        continue
    if ours.shape != ref.shape or ours.dtype != ref.dtype or not np.array_equal(ours, ref):
        bad.append("%%s: shape %%s vs %%s maxdiff %%s" %% (field, ours.shape, ref.shape,
                   np.abs(ours.astype(np.float64).ravel() - ref.astype(np.float64).ravel()).max() if ours.size == ref.size else "n/a"))
for cnt in ("link_count", "joint_coord_count", "joint_dof_count", "shape_count", "contact_count", "muscle_count", "articulation_count"):
    if int(getattr(env.model, cnt)) != int(gold["meta/" + cnt]):
        bad.append(cnt)
print("BAD" if bad else "OK", bad)
sys.exit(1 if bad else 0)
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "envs")), reason="reference tree not available")
@pytest.mark.parametrize("name", ["CartPoleSwingUpEnv", "AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "HopperEnv", "CheetahEnv"])
def test_reference_envs_build_identical_models_on_our_dflex(name):
    code = SCRIPT % dict(root=ROOT, ref=REF)
    proc = subprocess.run([sys.executable, "-c", code, name], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
