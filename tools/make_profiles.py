"""Turn the artefacts of a GPU evidence call (tools/gpu_call_s2*.sh: gpurun_out/<prefix>_*) into the committed summaries
under profiles/ (dev tool, build container: needs ncu + the object files the profiled library was built from).

    python tools/make_profiles.py <prefix> <objdir> [<tag>]

<objdir>: the directory with dfx_tile_e8.o / _e16.o / _e32.o of the PROFILED build (build/obj); <tag>: file-name stem (r02).
"""
import json, os, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix, objdir = sys.argv[1], sys.argv[2]
tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
py = sys.executable


def run(*a):
    return subprocess.run([py] + list(a), capture_output=True, text=True, cwd=ROOT).stdout


def one_obj(e):
    d = tempfile.mkdtemp()
    shutil.copy(os.path.join(objdir, "dfx_tile_e%d.o" % e), d)
    return d


CAPS = [("ant", "AntEnv 4096", 32, "32-environment tiles, 16 warps, 1 CTA per SM", "ILi16ELi1ELb%dELb1ELi0ELi9ELi14"),
        ("humanoid", "HumanoidEnv 8192", 8, "8-environment tiles, 8 warps, 2 CTAs per SM, compact layouts", "ILi8ELi2ELb%dELb1ELi3ELi22ELi27"),
        ("snu", "SNUHumanoidEnv 4096", 16, "16-environment tiles, 16 warps, 1 CTA per SM, compact layouts", "ILi16ELi1ELb%dELb1ELi3ELi11ELi24")]
for name, args, e, what, ksub in CAPS:
    rep = os.path.join(O, "prof_%s_%s.ncu-rep" % (prefix, name))
    if not os.path.exists(rep):
        continue
    d = one_obj(e)
    out = "# %s -- ncu --set full capture of the two %s kernels (%s)\n\n" % (tag, args.split()[0], what)
    out += ("Command: `ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dfx_ -o gpurun_out/prof_%s_%s "
            "python tools/prof_step.py %s` (one forward + one adjoint env-step after two warm-up steps; `tools/gpu_call_%s.sh`). "
            "Raw-page extract (`tools/ncu_summary.py raw`):\n\n" % (prefix, name, args, prefix))
    out += run("tools/ncu_summary.py", "raw", rep)
    for bwd, title in ((0, "forward"), (1, "adjoint")):
        txt = run("tools/ncu_by_line.py", rep, d, ksub % bwd, "12")
        keep = txt.split("== by function")[0]
        out += "\n## by phase, %s (`tools/ncu_by_line.py`: ncu source counters folded over the -lineinfo inline chains)\n```\n%s```\n" % (title, keep)
    open(os.path.join(P, "%s_%s_full.md" % (tag, name)), "w").write(out)
    print("wrote", name)

for src, dst, what in ((prefix + "_launches.csv", tag + "_launches", "kernel path (`value`): 32 forward + 32 adjoint launches through the C ABI on resident inputs (`--ncu-range`)"),
                       (prefix + "_e2e_launches.csv", tag + "_e2e_launches", "ONE end-to-end rollout (`e2e`): env.reset() + one CUDA graph = H2D actions, 32 x env.step, loss, backward, D2H (`--ncu-range-e2e`)")):
    f = os.path.join(O, src)
    if os.path.exists(f):
        shutil.copy(f, os.path.join(P, dst + ".csv"))
        body = run("tools/ncu_summary.py", "launches", f)
        open(os.path.join(P, dst + ".md"), "w").write(
            "# %s -- ncu launch list of %s\n\nCommand: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python bench.py "
            "--steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-configs --ncu-range[-e2e]` (`tools/gpu_call_%s.sh`); per-launch times under ncu are "
            "serialised and cold-cache: the SHARES are what is comparable with the bench line. Raw list: `profiles/%s.csv`.\n\n%s" % (tag, what, prefix, dst, body))
        print("wrote", dst)

for src, dst in ((prefix + "_bench.json", tag + "_bench_line.json"), (prefix + "_bench_reference.json", tag + "_bench_reference_line.json")):
    f = os.path.join(O, src)
    if os.path.exists(f):
        line = [l for l in open(f).read().splitlines() if l.startswith("{")][-1]
        json.loads(line)
        open(os.path.join(P, dst), "w").write(line + "\n")
        print("wrote", dst)
