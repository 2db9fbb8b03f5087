"""Kernel-only timing of forward / backward env-steps for a golden model tiled to N envs (dev tool)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from emu_util import load_golden
from diffrl_b200.modelpack import articulation_from_model
from diffrl_b200.engine import ArticulationEngine
from diffrl_b200 import _capi

name = sys.argv[1]; N = int(sys.argv[2]); groups = [int(g) for g in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
flag_list = [int(f) for f in sys.argv[4].split(",")] if len(sys.argv) > 4 else [3]
d, model = load_golden(name)
n0, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
desc, _ = articulation_from_model(model, n0)
p = "case%d/" % (int(d["meta/num_cases"]) - 1)
rng = np.random.default_rng(0)
pick = rng.integers(0, n0, N)
t = lambda a: torch.tensor(np.ascontiguousarray(a).ravel(), device="cuda:0")
q0 = t(d[p + "q0"].reshape(n0, -1)[pick]); qd0 = t(d[p + "qd0"].reshape(n0, -1)[pick]); act = t(d[p + "act"].reshape(n0, -1)[pick])
musc = t(d[p + "musc"].reshape(n0, -1)[pick]) if desc.M else None
gq = torch.randn_like(q0); gqd = torch.randn_like(qd0)
for grp, flags in [(g_, f_) for g_ in groups for f_ in flag_list]:
    _capi.lib().dfx_set_group_size(grp)
    _capi.lib().dfx_set_flags(flags)
    eng = ArticulationEngine(desc, N, "cuda:0")     # after the flags: bit 5 picks the kernel family per pack
    for _ in range(3):
        q, qd, tape, _x = eng.forward(q0, qd0, act, musc, S, mm, dt)
        eng.backward(act, musc, tape, gq, gqd, S, mm, dt)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 10
    e[0].record()
    for _ in range(reps): q, qd, tape, _x = eng.forward(q0, qd0, act, musc, S, mm, dt)
    e[1].record()
    for _ in range(reps): eng.backward(act, musc, tape, gq, gqd, S, mm, dt)
    e[2].record()
    torch.cuda.synchronize()
    tf, tb = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    print("flags=%d " % flags, end=""); print("%s N=%d G=%d: fwd %.3f ms  bwd %.3f ms  -> %.3g env-steps/s (kernel only, fwd+bwd)" % (name, N, grp, tf, tb, N / ((tf + tb) * 1e-3)), flush=True)
