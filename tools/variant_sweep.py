"""Parity + kernel-only timing of every kernel variant of an articulation (dev tool, GPU box).

    python tools/variant_sweep.py [--envs AntEnv,HumanoidEnv] [--n 4096] [--variants auto,tile8,...] [--reps 10] [--no-time]

Variants: auto (what a default pack picks), tileE (dfx_set_tile_envs(E): E-environment tile kernels, falls back to the
lane-group kernels when the articulation has none of that width -- reported as such), groupG (lane-group kernels with
G lanes per environment).  For each: max relative error against the reference goldens (forward state, all four
gradients), then forward / adjoint launch time at N environments (golden states tiled).  One JSON line per result.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from emu_util import load_golden
from diffrl_b200 import _capi
from diffrl_b200.engine import ArticulationEngine
from diffrl_b200.modelpack import articulation_from_model

DEFAULT_N = {"AntEnv": 4096, "HumanoidEnv": 8192, "SNUHumanoidEnv": 4096, "CartPoleSwingUpEnv": 4096, "HopperEnv": 4096, "CheetahEnv": 4096}


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def select(variant):
    lib = _capi.lib()
    lib.dfx_set_tile_envs(0); lib.dfx_set_group_size(0); lib.dfx_set_flags(9)
    if variant.startswith("tile"):
        lib.dfx_set_tile_envs(int(variant[4:].rstrip("LP")))
        if variant.endswith("L"):
            lib.dfx_set_flags(9 | 64)          # level-by-level tree recursions (alternative instantiation)
        if variant.endswith("P"):
            lib.dfx_set_flags(9 | 128)         # no L2 prefetch of the next tape row
    elif variant.startswith("group"):
        lib.dfx_set_flags(9 | 32); lib.dfx_set_group_size(int(variant[5:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="AntEnv,HumanoidEnv,SNUHumanoidEnv,HopperEnv,CheetahEnv,CartPoleSwingUpEnv")
    ap.add_argument("--variants", default="auto,tile8,tile16,tile32,group32")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-time", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    t = lambda x: None if x is None else torch.tensor(np.ascontiguousarray(x).ravel(), device=dev)
    for name in a.envs.split(","):
        d, model = load_golden(name)
        n0, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
        seen = set()
        for variant in a.variants.split(","):
            select(variant)
            try:
                eng = ArticulationEngine.from_model(model, dev, n0)
                tile = int(eng.lib.dfx_pack_query(eng.pack, 9))
                family = "tile%d" % tile if tile else "group"
                if variant != "auto" and variant.startswith("tile") and family != variant.rstrip("LP"):
                    print(json.dumps({"env": name, "variant": variant, "skipped": "no tile kernel of this width"}), flush=True)
                    continue
                err = {"q": 0.0, "gq": 0.0, "gact": 0.0, "gmusc": 0.0}
                for k in range(int(d["meta/num_cases"])):
                    p = "case%d/" % k
                    musc = t(d[p + "musc"]) if (p + "musc") in d.files else None
                    q, qd, tape, _ = eng.forward(t(d[p + "q0"]), t(d[p + "qd0"]), t(d[p + "act"]), musc, S, mm, dt)
                    gq, gqd, gact, gm = eng.backward(t(d[p + "act"]), musc, tape, t(d[p + "gq_out"]), t(d[p + "gqd_out"]), S, mm, dt)
                    torch.cuda.synchronize()
                    err["q"] = max(err["q"], rel(q.cpu(), d[p + "traj_q"][-1]), rel(qd.cpu(), d[p + "traj_qd"][-1]))
                    err["gq"] = max(err["gq"], rel(gq.cpu(), d[p + "grad_q"]), rel(gqd.cpu(), d[p + "grad_qd"]))
                    err["gact"] = max(err["gact"], rel(gact.cpu(), d[p + "grad_act"]))
                    if gm is not None:
                        err["gmusc"] = max(err["gmusc"], rel(gm.cpu(), d[p + "grad_musc"]))
                out = {"env": name, "variant": variant, "family": family, "err": err}
                plan = (ArticulationEngine.__dict__.get("launch_plan") or (lambda self, b: None))
                if not a.no_time:
                    N = a.n or DEFAULT_N[name]
                    desc, _ = articulation_from_model(model, n0)
                    engN = ArticulationEngine(desc, N, dev)
                    p = "case%d/" % (int(d["meta/num_cases"]) - 1)
                    pick = np.random.default_rng(0).integers(0, n0, N)
                    q0 = t(d[p + "q0"].reshape(n0, -1)[pick]); qd0 = t(d[p + "qd0"].reshape(n0, -1)[pick]); act = t(d[p + "act"].reshape(n0, -1)[pick])
                    musc = t(d[p + "musc"].reshape(n0, -1)[pick]) if desc.M else None
                    gq, gqd = torch.randn_like(q0), torch.randn_like(qd0)
                    for _ in range(3):
                        q, qd, tape, _x = engN.forward(q0, qd0, act, musc, S, mm, dt)
                        engN.backward(act, musc, tape, gq, gqd, S, mm, dt)
                    torch.cuda.synchronize()
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    ev[0].record()
                    for _ in range(a.reps):
                        q, qd, tape, _x = engN.forward(q0, qd0, act, musc, S, mm, dt)
                    ev[1].record()
                    for _ in range(a.reps):
                        engN.backward(act, musc, tape, gq, gqd, S, mm, dt)
                    ev[2].record()
                    torch.cuda.synchronize()
                    tf, tb = ev[0].elapsed_time(ev[1]) / a.reps, ev[1].elapsed_time(ev[2]) / a.reps
                    import ctypes
                    pl = [(ctypes.c_int * 6)(), (ctypes.c_int * 6)()]
                    engN.lib.dfx_launch_plan(engN.pack, 0, pl[0]); engN.lib.dfx_launch_plan(engN.pack, 1, pl[1])
                    out.update({"N": N, "fwd_ms": tf, "bwd_ms": tb, "env_steps_per_s": N / ((tf + tb) * 1e-3),
                                "smem_fwd": pl[0][3], "smem_bwd": pl[1][3], "envs_per_cta": pl[0][1], "finite": bool(torch.isfinite(q).all())})
                    del engN, tape
                print(json.dumps(out), flush=True)
            except Exception as exc:  # keep sweeping
                print(json.dumps({"env": name, "variant": variant, "error": repr(exc)[:300]}), flush=True)
                torch.cuda.synchronize() if False else None
    select("auto")


if __name__ == "__main__":
    main()
