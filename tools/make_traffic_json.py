"""profiles/r02_traffic.json from a full ncu capture of tools/prof_step.py: per-launch DRAM bytes and executed fp32 FLOP of the
forward / adjoint kernel (what bench.py reports as roofline.traffic and roofline.fp32).
usage: python tools/make_traffic_json.py <Env>=<report.ncu-rep>:<num_envs> ... > profiles/r02_traffic.json"""
import csv, io, json, subprocess, sys

def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]

def num(row, hdr, units, key):
    i = hdr.index(key)
    v = float(row[i].replace(",", ""))
    u = units[i]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)

res = {}
for spec in sys.argv[1:]:
    env, rest = spec.split("=")
    rep, n = rest.rsplit(":", 1)
    hdr, units, body = rows(rep)
    kn = hdr.index("Kernel Name")
    d = {"num_envs": int(n), "source": "%s (ncu --set full --clock-control none, tools/prof_step.py %s %s): dram__bytes_read.sum + dram__bytes_write.sum; "
         "fp32 FLOP = smsp__sass_thread_inst_executed_op_{ffma x2, fmul, fadd}_pred_on.sum.per_cycle_elapsed x sm__cycles_elapsed.avg" % (rep.split("/")[-1], env, n)}
    for row in body:
        # template args: <NW, MINB, BWD, ...>
        args = row[kn].split("<")[1].split(",")
        tag = "bwd" if args[2].strip() in ("1", "(bool)1") else "fwd"
        d[tag + "_dram_bytes_per_launch"] = int(num(row, hdr, units, "dram__bytes_read.sum") + num(row, hdr, units, "dram__bytes_write.sum"))
        # (the set reports the op counters per elapsed cycle: x sm__cycles_elapsed.avg gives the totals)
        cyc = num(row, hdr, units, "sm__cycles_elapsed.avg")
        per = lambda op: num(row, hdr, units, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)
        flop = (2 * per("ffma") + per("fmul") + per("fadd")) * cyc
        d[tag + "_fp32_flop_per_launch"] = int(flop)
        d[tag + "_us_under_ncu"] = num(row, hdr, units, "gpu__time_duration.sum") * ({"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[hdr.index("gpu__time_duration.sum")], 1.0))
    res[env] = d
print(json.dumps(res, indent=1))
