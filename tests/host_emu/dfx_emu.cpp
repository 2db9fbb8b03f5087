// dfx_emu.cpp -- HOST EMULATION of the device code, for CPU-side unit tests only.
//
// Compiles the very same phase / step headers the CUDA kernels are built from
// (diffrl_b200/csrc/dfx_{math,phases,step}.h) with g++, one "lane" per environment, so that the
// arithmetic and the hand-derived adjoints can be checked against the reference's golden
// vectors in a container that has no GPU.  It is test infrastructure: nothing in the product
// (diffrl_b200/) loads it, and the product fails loudly when the CUDA library is missing.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../diffrl_b200/csrc/dfx_step.h"

using namespace dfx;

struct EmuPack {
    PackHost host;
    Pack pack;
};

static int g_tape_bf16 = 0;

extern "C" {

// round the middle of every tape row through bf16, as the tile kernels' bf16 tape does (dfx_set_tape_dtype)
void emu_set_tape_bf16(int on) { g_tape_bf16 = on ? 1 : 0; }

EmuPack* emu_pack_create(const DfxModelDesc* desc, char* err, int err_len) {
    EmuPack* p = new EmuPack();
    std::string msg;
    if (!build_pack(*desc, p->host, msg)) {
        if (err && err_len > 0) { strncpy(err, msg.c_str(), err_len - 1); err[err_len - 1] = 0; }
        delete p;
        return nullptr;
    }
    p->pack = p->host.bind(p->host.ints.data(), p->host.floats.data());
#ifdef DFX_EMU_LAYOUT_MODE
    p->host.set_layout_mode(DFX_EMU_LAYOUT_MODE);     // the compact layouts of the large-articulation tile kernels
#endif
    return p;
}
void emu_pack_destroy(EmuPack* p) { delete p; }
int emu_pack_query(const EmuPack* p, int what) {
    switch (what) {
        case DFX_QUERY_LINKS: return p->pack.L;
        case DFX_QUERY_DOFS: return p->pack.D;
        case DFX_QUERY_COORDS: return p->pack.Q;
        case DFX_QUERY_CONTACTS: return p->pack.C;
        case DFX_QUERY_MUSCLES: return p->pack.M;
        case DFX_QUERY_FWD_SCRATCH_FLOATS: return p->host.layout.fwd_size;
        case DFX_QUERY_BWD_SCRATCH_FLOATS: return p->host.layout_bwd.bwd_size;
        case DFX_QUERY_TAPE_ROW_FLOATS: return p->host.layout.tape_row;
        case DFX_QUERY_TREE_DEPTH: return p->pack.nlev;
        case DFX_QUERY_JOINT_MASK: return p->pack.jmask;
    }
    return -1;
}
int emu_pack_set_gravity(EmuPack* p, float gx, float gy, float gz, int ground) {
    p->pack.gx = gx; p->pack.gy = gy; p->pack.gz = gz;
    p->pack.ground = (ground && p->pack.C > 0) ? 1 : 0;
    return 0;
}
long long emu_tape_floats(const EmuPack* p, int n, int substeps, int mm_freq) {
    return tape_geom(p->pack.L, p->pack.Q, p->pack.D, n, substeps, mm_freq).total;
}

int emu_step_forward(const EmuPack* p, int n, int substeps, int mm_freq, double dt,
                     const float* q, const float* qd, const float* act, const float* musc,
                     float* q_out, float* qd_out, float* tape, const DfxDerived* derived) {
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.N = n; a.substeps = substeps; a.mm_freq = mm_freq;
    a.dt_sub = (float)(dt / (double)substeps);
    a.q = q; a.qd = qd; a.act = act; a.musc = musc; a.q_out = q_out; a.qd_out = qd_out; a.tape = tape;
    if (derived) { a.derived = *derived; a.has_derived = 1; }
    a.tape_bf16 = g_tape_bf16;
    a.hinv_base = tape_geom(p->pack.L, p->pack.Q, p->pack.D, n, substeps, mm_freq).hinv_base;
    // (filled with NaN: a field the layout overlays or drops must never be read before it is written)
    std::vector<float> scratch((size_t)(p->host.layout.fwd_size + 16) * DFX_ES, nanf(""));
    GroupSerial g{0};
    for (int env = 0; env < n; ++env) env_step_forward(p->pack, p->host.layout, SP{scratch.data()}, g, env, a);
    return 0;
}

int emu_step_backward(const EmuPack* p, int n, int substeps, int mm_freq, double dt,
                      const float* act, const float* musc, const float* tape,
                      const float* gq_out, const float* gqd_out,
                      float* gq, float* gqd, float* gact, float* gmusc) {
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.N = n; a.substeps = substeps; a.mm_freq = mm_freq;
    a.dt_sub = (float)(dt / (double)substeps);
    a.act = act; a.musc = musc; a.tape_in = tape; a.gq_out = gq_out; a.gqd_out = gqd_out;
    a.gq = gq; a.gqd = gqd; a.gact = gact; a.gmusc = gmusc;
    a.tape_bf16 = g_tape_bf16;
    a.hinv_base = tape_geom(p->pack.L, p->pack.Q, p->pack.D, n, substeps, mm_freq).hinv_base;
    std::vector<float> scratch((size_t)(p->host.layout_bwd.bwd_size + 16) * DFX_ES, nanf(""));
    GroupSerial g{0};
    for (int env = 0; env < n; ++env) env_step_backward(p->pack, p->host.layout_bwd, SP{scratch.data()}, g, env, a);
    return 0;
}

// ---- the fixed-point scatter-add of dfx_phases.h in isolation (tests/test_fixed_point.py): n contributions
// vals[k] are scattered in the order perm[k] into ONE accumulator with `scale`; returns the low / high words,
// the read-back value and whether the poison bit was raised
int emu_fx_accumulate(const float* vals, const int* perm, int n, float scale, int* lo_out, int* hi_out, float* value) {
    float lo_f[DFX_ES] = {0.0f};           // one strided slot each, as in the kernels (the low word lives in an fp32 array)
    float hi_f[DFX_ES] = {0.0f};
    float poison_f[DFX_ES] = {0.0f};
    const SP lo{lo_f}, hi{hi_f}, poison{poison_f};
    GroupSerial g{0};
    for (int k = 0; k < n; ++k) {
        const float w[1] = {vals[perm[k]]};
        fx_scatter(sp_int(lo), sp_int(hi), sp_uint(poison), 0, w, scale, g);
    }
    *lo_out = sp_int(lo)[0];
    *hi_out = sp_int(hi)[0];
    *value = fx_value(*lo_out, *hi_out, 1.0f / scale);
    return (int)(sp_uint(poison)[0] & 1u);
}

float emu_fx_pow2_scale(float m) { return fx_pow2_scale(m); }

}  // extern "C"
