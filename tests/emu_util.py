"""ctypes driver for the host emulation of the device code (tests/host_emu) -- CPU tests only."""
import ctypes
import os
import subprocess

import numpy as np

from diffrl_b200.modelpack import DfxDerived, DfxModelDesc, articulation_from_model

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_SRC = os.path.join(HERE, "host_emu", "dfx_emu.cpp")
EMU_LIB = os.path.join(HERE, "host_emu", "libdfx_emu.so")
GOLDEN = os.path.join(HERE, "golden")

_F = ctypes.POINTER(ctypes.c_float)


def build_emu(es=1, path_passes=False, layout_mode=0, extra=()):
    """es: element stride of the per-environment scratch (csrc/dfx_math.h DFX_ES).  1 is what the lane-group kernels
    use; any other value exercises the strided addressing of the 32-environment tile kernels on the CPU.
    path_passes: the tile kernels' path / subtree formulation of the tree passes (Grp::kPathPasses)."""
    lib_path = EMU_LIB
    if es != 1 or path_passes or layout_mode or extra:
        tag = "".join(c if c.isalnum() else "_" for c in "".join(extra))
        lib_path = EMU_LIB.replace(".so", "_es%d%s%s%s.so" % (es, "_path" if path_passes else "", "_lm%d" % layout_mode if layout_mode else "", tag))
    return _build_emu(lib_path, es, path_passes, layout_mode, extra)


def _build_emu(EMU_LIB, es, path_passes=False, layout_mode=0, extra=()):
    deps = [EMU_SRC] + [os.path.join(ROOT, "diffrl_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "diffrl_b200", "csrc")) if f.endswith(".h")]
    deps.append(os.path.join(ROOT, "include", "dfx.h"))
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17"] + ([] if any("contract" in x for x in extra) else ["-ffp-contract=off"]) + list(extra) + ["-DDFX_ES=%d" % es, "-DDFX_EMU_PATH_PASSES=%d" % int(path_passes)] + (["-DDFX_EMU_LAYOUT_MODE=%d" % layout_mode] if layout_mode else []) + ["-shared", "-fPIC",
                               "-o", EMU_LIB, EMU_SRC])
    lib = ctypes.CDLL(EMU_LIB)
    lib.emu_pack_create.restype = ctypes.c_void_p
    lib.emu_pack_create.argtypes = [ctypes.POINTER(DfxModelDesc), ctypes.c_char_p, ctypes.c_int]
    lib.emu_pack_destroy.argtypes = [ctypes.c_void_p]
    lib.emu_pack_query.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.emu_tape_floats.restype = ctypes.c_longlong
    lib.emu_tape_floats.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.emu_step_forward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                     _F, _F, _F, _F, _F, _F, _F, ctypes.POINTER(DfxDerived)]
    lib.emu_step_backward.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      _F, _F, _F, _F, _F, _F, _F, _F, _F]
    return lib


def fptr(a):
    return None if a is None else a.ctypes.data_as(_F)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = {k[len("model/"):]: d[k] for k in d.files if k.startswith("model/")}
    model["ground"] = bool(d["meta/ground"])
    return d, model


class EmuSim:
    """Host-emulated integrator for one golden model."""

    DERIVED = {"body_X_sc": ("L", 7), "body_X_sm": ("L", 7), "joint_S_s": ("D", 6), "body_v_s": ("L", 6),
               "body_a_s": ("L", 6), "body_f_s": ("L", 6), "body_ft_s": ("L", 6), "joint_tau": ("D", 1),
               "joint_qdd": ("D", 1), "H": ("DD", 1), "L": ("DD", 1)}

    def __init__(self, model, num_envs, es=1, path_passes=False, layout_mode=0, extra=()):
        """layout_mode: csrc/dfx_pack.h kLayoutCompact (1) | kLayoutHinvGlobal (2) -- the scratch layouts of the
        large-articulation tile kernels."""
        self.lib = build_emu(es, path_passes, layout_mode, extra)     # extra: more g++ flags (A/B builds)
        self.desc, self.N = articulation_from_model(model, num_envs)
        err = ctypes.create_string_buffer(256)
        st = self.desc.as_struct()
        self.pack = self.lib.emu_pack_create(ctypes.byref(st), err, 256)
        if not self.pack:
            raise RuntimeError(err.value.decode())

    tape_bf16 = False     # emulate the tile kernels' bf16 tape (the middle of every row rounded through bf16)

    def forward(self, q, qd, act, musc, substeps, mm_freq, dt, tape=True, derived=False):
        N, d = self.N, self.desc
        self.lib.emu_set_tape_bf16(int(self.tape_bf16))
        q = np.ascontiguousarray(q, np.float32); qd = np.ascontiguousarray(qd, np.float32)
        act = np.ascontiguousarray(act, np.float32)
        musc = None if musc is None else np.ascontiguousarray(musc, np.float32)
        q_out, qd_out = np.empty_like(q), np.empty_like(qd)
        tp = None
        if tape:
            tp = np.zeros(self.lib.emu_tape_floats(self.pack, N, substeps, mm_freq), np.float32)
        dv, bufs = None, {}
        if derived:
            dv = DfxDerived()
            dims = {"L": d.L, "D": d.D, "DD": d.D * d.D}
            for k, (a, b) in self.DERIVED.items():
                bufs[k] = np.zeros((N * dims[a], b), np.float32)
                setattr(dv, k, bufs[k].ctypes.data)
        self.lib.emu_step_forward(self.pack, N, substeps, mm_freq, float(dt), fptr(q), fptr(qd), fptr(act), fptr(musc),
                                  fptr(q_out), fptr(qd_out), fptr(tp), None if dv is None else ctypes.byref(dv))
        return q_out, qd_out, tp, bufs

    def backward(self, act, musc, tape, gq_out, gqd_out, substeps, mm_freq, dt):
        N, d = self.N, self.desc
        act = np.ascontiguousarray(act, np.float32)
        musc = None if musc is None else np.ascontiguousarray(musc, np.float32)
        gq_out = np.ascontiguousarray(gq_out, np.float32); gqd_out = np.ascontiguousarray(gqd_out, np.float32)
        gq, gqd, gact = np.zeros(N * d.Q, np.float32), np.zeros(N * d.D, np.float32), np.zeros(N * d.D, np.float32)
        gmusc = np.zeros(N * d.M, np.float32) if d.M else None
        self.lib.emu_step_backward(self.pack, N, substeps, mm_freq, float(dt), fptr(act), fptr(musc), fptr(tape),
                                   fptr(gq_out), fptr(gqd_out), fptr(gq), fptr(gqd), fptr(gact), fptr(gmusc))
        return gq, gqd, gact, gmusc
