"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line / function.

usage: python tools/ncu_by_line.py <report.ncu-rep> <lib.so | build/obj> <kernel-substring e.g. 'ILi16ELb0'> [top]
Needs the .so compiled with -lineinfo.  (ncu's own CSV export of the CUDA view carries no metrics.)
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# <lib.so> may also be a directory of object files (build/obj): dfx_tile.cu is compiled three times (one per tile width) and
# cuobjdump names extracted cubins after the SOURCE file, so the three would overwrite each other when taken from the .so
objs = [os.path.join(lib, f) for f in sorted(os.listdir(lib)) if f.endswith(".o")] if os.path.isdir(lib) else [lib]
dis = ""
for o in objs:
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(o)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in sorted(os.listdir(tmp)):
        if f.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
# ---- offset -> (file, line) of the innermost inlined function, and the whole inline chain, for the chosen kernel
off2line, off2chain, cur, chain, inside, fresh = {}, {}, None, [], False, True
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        inside = ksub in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)(?: inlined at "(.*?)", line (\d+))?', ln)
    if m:
        if fresh:
            chain, fresh = [], False
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        off2line[int(m.group(1), 16)] = cur
        off2chain[int(m.group(1), 16)] = chain
        fresh = True
# ---- ncu per-instruction rows
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# ksub is a mangled-name fragment such as ILi16ELb0 or ILi32ELb1ELi11; translate to the demangled fragment
import re as _re
nums = _re.findall(r"L([ib])(\d+)E", ksub)
want = ", ".join(("(int)%s" % v) if t == "i" else ("(bool)%s" % v) for t, v in nums)
i = 0
sect = None
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and want in rows[i][1]:
        sect = i
        break
    i += 1
assert sect is not None, "kernel not in report"
hdr = rows[sect + 1]
ia, ie, isamp, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
body = []
for r in rows[sect + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
base = int(body[0][ia], 16)
# ---- function ranges per file
def func_table(path):
    tbl = []
    for n, ln in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*\n)?(?:DFX_HD|inline|static|__device__|__global__)[^;(]*?\b(\w+)\(", ln)
        if m and not ln.startswith(" "):
            tbl.append((n, m.group(1)))
    return tbl
root = os.environ.get("DFX_SRC_ROOT") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffrl_b200", "csrc")   # the sources the profiled library was built from
ftab = {f: func_table(os.path.join(root, f)) for f in os.listdir(root)}
def func_of(file, line):
    best = "?"
    for n, name in ftab.get(file, []):
        if n <= line:
            best = name
        else:
            break
    return best
PHASES = ["kin_fwd", "body_force_fwd", "contact_fwd", "muscle_fwd", "wrench_collect", "tau_fwd", "crba_fwd", "chol_inverse",
          "solve_fwd", "integrate_fwd", "integrate_adj", "solve_adj", "crba_adj", "tau_adj", "muscle_adj", "contact_adj",
          "adj_scatter_scale", "adj_collect", "body_force_adj", "kin_adj", "zero_range", "dump_derived",
          "tile_transition_forward", "tile_transition_backward",      # env transition as epilogue / prologue (dfx_env_dev.h)
          # the items / tasks of the combined rigid-body + contact phases (cta_compact_with)
          "contact_point_fwd", "contact_point_adj", "body_force_link_fwd", "body_force_link_adj", "contact_penetrates",
          "fx_scatter", "cta_compact_with", "cta_compact", "block_in", "block_out", "row_in"]
def phase_of(chain):
    names = [func_of(*c) for c in chain]
    hit = None
    for nm in names:                      # innermost first; keep the OUTERMOST of a run of nested phase names,
        if nm in PHASES:                  # but let the combined-phase items (listed last in PHASES) win over their host
            if hit is None or PHASES.index(nm) < PHASES.index("contact_point_fwd"):
                hit = nm
    if hit is not None:
        return hit
    for nm in reversed(names):
        if nm.startswith("env_step") or nm.startswith("dfx_step_kernel") or nm in ("copy_row_async", "copy_row_out", "copy_wait_all"):
            return nm
    return names[-1] if names else "?"
def subphase_of(chain):
    """(phase, the function called directly from the phase body)"""
    names = [func_of(*c) for c in chain]
    for k in range(len(names) - 1, -1, -1):
        if names[k] in PHASES:
            child = names[k]
            for j in range(k - 1, -1, -1):
                if names[j] != names[k]:
                    child = names[j]
                    break
            return names[k] + " > " + child
    return phase_of(chain)
by_sub, samp_sub, thr_sub = collections.Counter(), collections.Counter(), collections.Counter()
by_phase, samp_phase, thr_phase = collections.Counter(), collections.Counter(), collections.Counter()
static_phase = collections.Counter()      # SASS instructions of the kernel that belong to the phase (16 bytes each): its instruction-cache footprint
stall_phase = collections.defaultdict(collections.Counter)
by_line, by_func = collections.Counter(), collections.Counter()
samp_line, samp_func, thr_func = collections.Counter(), collections.Counter(), collections.Counter()
stall_func = collections.defaultdict(collections.Counter)
tot = tots = 0
for r in body:
    off = int(r[ia], 16) - base
    loc = off2line.get(off) or ("?", 0)
    n, s, th = float(r[ie] or 0), float(r[isamp] or 0), float(r[ith] or 0)
    fn = func_of(*loc)
    ph = phase_of(off2chain.get(off) or [])
    by_phase[ph] += n; samp_phase[ph] += s; thr_phase[ph] += th; static_phase[ph] += 1
    sp = subphase_of(off2chain.get(off) or [])
    by_sub[sp] += n; samp_sub[sp] += s; thr_sub[sp] += th
    for c in stall_cols:
        stall_phase[ph][c] += float(r[hdr.index(c)] or 0)
    by_line[loc] += n; by_func[fn] += n; samp_line[loc] += s; samp_func[fn] += s; thr_func[fn] += th
    for c in stall_cols:
        stall_func[fn][c] += float(r[hdr.index(c)] or 0)
    tot += n; tots += s
print("kernel %s: %.3g warp-instructions, %d samples" % (want, tot, tots))
print("\n== by phase (outermost phase function on the inline chain): %inst  %samples  lanes/inst  top stalls")
for fn, n in by_phase.most_common(30):
    st = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, samp_phase[fn])) for k, v in stall_phase[fn].most_common(4))
    print("%6.2f%% %6.2f%%  %5.1f  %-26s %s" % (100 * n / tot, 100 * samp_phase[fn] / max(1, tots), thr_phase[fn] / max(1, n), fn, st))
print("\n== code size by phase (SASS bytes of the kernel, executed or not): KB")
for fn, n in static_phase.most_common(14):
    print("%7.1f  %s" % (n * 16 / 1024.0, fn))
print("%7.1f  (whole kernel)" % (sum(static_phase.values()) * 16 / 1024.0))
print("\n== by phase > callee: %inst  %samples  lanes/inst")
for fn, n in by_sub.most_common(40):
    print("%6.2f%% %6.2f%%  %5.1f  %s" % (100 * n / tot, 100 * samp_sub[fn] / max(1, tots), thr_sub[fn] / max(1, n), fn))
print("\n== by function: %inst  %samples  lanes/inst  top stalls")
for fn, n in by_func.most_common(30):
    st = ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, samp_func[fn])) for k, v in stall_func[fn].most_common(4))
    print("%6.2f%% %6.2f%%  %5.1f  %-26s %s" % (100 * n / tot, 100 * samp_func[fn] / max(1, tots), thr_func[fn] / max(1, n), fn, st))
print("\n== by line: %inst  %samples")
for loc, n in by_line.most_common(top):
    print("%6.2f%% %6.2f%%  %s:%d" % (100 * n / tot, 100 * samp_line[loc] / max(1, tots), loc[0], loc[1]))
