"""Write diffrl_b200/assets/<Env>.npz: the finalized Model tensors of ONE articulation (environment 0 of
the 2-env golden models produced by oracle/make_golden.py from the reference's own asset parsers),
indices made env-local.  These are what `diffrl_b200.envs` tile to N environments at start-up (the
reference re-parses the MJCF/URDF/SNU files once per environment instead)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from emu_util import load_golden
from diffrl_b200.modelpack import articulation_from_model

for name in ["CartPoleSwingUpEnv", "AntEnv", "HumanoidEnv", "SNUHumanoidEnv", "HopperEnv", "CheetahEnv"]:
    d, model = load_golden(name)
    n = int(d["meta/num_envs"])
    desc, _ = articulation_from_model(model, n)
    out = dict(desc.arrays)
    L, S = desc.L, desc.counts["shape_count"]
    Q, D = desc.Q, desc.D
    out["joint_q"] = model["joint_q"][:Q].astype(np.float32)
    out["joint_qd"] = model["joint_qd"][:D].astype(np.float32)
    out["shape_transform"] = model["shape_transform"][:S].astype(np.float32)
    out["shape_body"] = model["shape_body"][:S].astype(np.int32)
    out["shape_geo_type"] = model["shape_geo_type"][:S].astype(np.int32)
    out["shape_geo_scale"] = model["shape_geo_scale"][:S].astype(np.float32)
    M = desc.M
    out["muscle_params"] = (model["muscle_params"][:M].astype(np.float32) if M else np.zeros((0, 5), np.float32))
    out["gravity"] = model["gravity"].astype(np.float32)
    out["ground"] = np.int32(model["ground"])
    path = os.path.join(ROOT, "diffrl_b200", "assets", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
