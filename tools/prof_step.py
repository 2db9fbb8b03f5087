"""One forward + one adjoint env-step of an articulation inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_x python tools/prof_step.py HumanoidEnv 8192 [variant]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from emu_util import load_golden
from diffrl_b200.modelpack import articulation_from_model
from diffrl_b200.engine import ArticulationEngine
sys.path.insert(0, os.path.join(ROOT, "tools"))
from variant_sweep import select

name, N = sys.argv[1], int(sys.argv[2])
select(sys.argv[3] if len(sys.argv) > 3 else "auto")
d, model = load_golden(name)
n0, S, mm, dt = int(d["meta/num_envs"]), int(d["meta/substeps"]), int(d["meta/mass_matrix_freq"]), float(d["meta/dt"])
desc, _ = articulation_from_model(model, n0)
p = "case%d/" % (int(d["meta/num_cases"]) - 1)
pick = np.random.default_rng(0).integers(0, n0, N)
t = lambda a: torch.tensor(np.ascontiguousarray(a).ravel(), device="cuda:0")
q0 = t(d[p + "q0"].reshape(n0, -1)[pick]); qd0 = t(d[p + "qd0"].reshape(n0, -1)[pick]); act = t(d[p + "act"].reshape(n0, -1)[pick])
musc = t(d[p + "musc"].reshape(n0, -1)[pick]) if desc.M else None
gq, gqd = torch.randn_like(q0), torch.randn_like(qd0)
eng = ArticulationEngine(desc, N, "cuda:0")
for _ in range(2):
    q, qd, tape, _x = eng.forward(q0, qd0, act, musc, S, mm, dt)
    eng.backward(act, musc, tape, gq, gqd, S, mm, dt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
q, qd, tape, _x = eng.forward(q0, qd0, act, musc, S, mm, dt)
eng.backward(act, musc, tape, gq, gqd, S, mm, dt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", name, N)
