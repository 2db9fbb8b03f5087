// dfx_step.h -- one environment's whole env-step (all substeps) forward and backward, for a
// group policy owning the environment's scratch.  Shared verbatim by the tile kernels (dfx_tile.cu), the
// lane-group kernels (dfx_kernels.cu) and the host emulation used by CPU-side unit tests.
//
// Tape (written by forward when taping, read by backward), all fp32, one block per substep / per update:
//   rows   per substep: the (q, qd) ENTERING it and the forward intermediates the adjoint
//          needs -- X_sc, X_sm, S, v, a, total link wrenches, q'' (32 L + 8 D + Q floats, padded to 4)
//   H^-1   D * D floats per mass-matrix update
// laid out [block][env][n] (lane-group kernels, host) or [block][tile of 32 envs][n][32] (tile kernels): the
// group policy's block_in / row_in / block_out hide the difference.
// The path is issue / latency bound with HBM idle, so the adjoint trades bandwidth for instructions: it reads these
// rows back (coalesced, tile-contiguous per substep) instead of re-running the forward dynamics.  Ant: 1.7 KB
// per env-substep vs the reference's ~3.7 KB (it keeps every State tensor alive).
#pragma once

#include "../../include/dfx.h"
#include "dfx_pack_build.h"
#include "dfx_phases.h"

namespace dfx {

struct StepArgs {
    int N, substeps, mm_freq;
    float dt_sub;
    const float* q; const float* qd; const float* act; const float* musc;
    float* q_out; float* qd_out;
    float* tape;
    DfxDerived derived; int has_derived;
    // backward
    const float* tape_in;
    const float* gq_out; const float* gqd_out;
    float* gq; float* gqd; float* gact; float* gmusc;
    long long hinv_base;
    int flags;   // bit 1: CTA-wide phase barriers
    int tape_bf16;   // the middle of every tape row (the forward intermediates) is stored as bf16 (tile kernels)
    // action map folded into the step (include/dfx.h dfx_step_*_mapped): raw != nullptr
    const float* raw; const float* strength; float* used; const float* g_used; float* g_raw;
    int map_num_act, map_offset, map_muscle;
    float map_pre_scale, map_pre_bias, map_drive_scale;
    // env transition folded into the step launch (include/dfx.h dfx_env_step_*; tile kernels only): env_kind != 0.  Forward:
    // `env` is read after the step (epilogue); backward: `env_adj` before it (prologue), writing gq_out / gqd_out / g_used.
    int env_kind;
    DfxEnvTransition env;
    DfxEnvTransitionAdj env_adj;
};

// derived State fields of the last substep (reference model.py:375-388); `late` = the ones that exist only after the solve
template <class Grp>
DFX_HD void dump_derived(const Pack& P, const Layout& Y, SP s, const DfxDerived& d, int env, bool late, const Grp& g) {
    const int L = P.L, D = P.D;
    if (late) {
        if (d.joint_qdd) DFX_FOR(i, D) d.joint_qdd[(long long)env * D + i] = s[Y.qdd + i];
        return;
    }
    if (d.body_X_sc) DFX_FOR(i, L * 7) d.body_X_sc[(long long)env * L * 7 + i] = s[Y.Xsc + i];
    if (d.body_X_sm) DFX_FOR(i, L * 7) d.body_X_sm[(long long)env * L * 7 + i] = s[Y.Xsm + i];
    if (d.joint_S_s) DFX_FOR(i, D * 6) d.joint_S_s[(long long)env * D * 6 + i] = s[Y.S + i];
    if (d.body_v_s) DFX_FOR(i, L * 6) d.body_v_s[(long long)env * L * 6 + i] = s[Y.v + i];
    if (d.body_a_s) DFX_FOR(i, L * 6) d.body_a_s[(long long)env * L * 6 + i] = s[Y.a + i];
    if (d.body_f_s) DFX_FOR(i, L * 6) d.body_f_s[(long long)env * L * 6 + i] = s[Y.f + i];
    if (d.body_ft_s) DFX_FOR(i, L * 6) d.body_ft_s[(long long)env * L * 6 + i] = s[Y.ft + i] - s[Y.f + i];
    if (d.joint_tau) DFX_FOR(i, D) d.joint_tau[(long long)env * D + i] = s[Y.tau + i];
}

// actuation of one environment: from the caller's arrays, or formed from the raw policy output (same arithmetic as
// dfx_action_map_forward: used = clip(raw, -1, 1) * pre_scale + pre_bias ; drive = (used * drive_scale) * strength)
template <class Grp>
DFX_HD void load_actuation(const Pack& P, const Layout& Y, SP s, const Grp& g, int env, const StepArgs& a, bool write_used) {
    const int D = P.D, M = P.M;
    if (!a.raw) {
        DFX_FOR(i, D) s[Y.act + i] = a.act[(long long)env * D + i];
        DFX_FOR(i, M) s[Y.musc + i] = a.musc[(long long)env * M + i];
        return;
    }
    const int A = a.map_num_act, off = a.map_offset;
    const float* re = a.raw + (long long)env * A;
    const int width = a.map_muscle ? M : D;
    const int dst = a.map_muscle ? Y.musc : Y.act;
    DFX_FOR(i, width) {
        const int j = i - off;
        float out = 0.0f;
        if (j >= 0 && j < A) {
            const float u = fminf(fmaxf(re[j], -1.0f), 1.0f) * a.map_pre_scale + a.map_pre_bias;
            out = (u * a.map_drive_scale) * a.strength[j];
        }
        s[dst + i] = out;
    }
    if (a.map_muscle) { DFX_FOR(i, D) s[Y.act + i] = a.act ? a.act[(long long)env * D + i] : 0.0f; }
    else { DFX_FOR(i, M) s[Y.musc + i] = a.musc ? a.musc[(long long)env * M + i] : 0.0f; }
    if (write_used && a.used) DFX_FOR(j, A) a.used[(long long)env * A + j] = fminf(fmaxf(re[j], -1.0f), 1.0f) * a.map_pre_scale + a.map_pre_bias;
}

template <class Grp>
DFX_HD void env_step_forward(const Pack& P, const Layout& Y, SP s, const Grp& g, int env, const StepArgs& a) {
    const int Q = P.Q, D = P.D, DD = D * D;
    const RowFmt rf = row_fmt(P, Y, a.tape_bf16 != 0);
    DFX_FOR(i, Q) s[Y.q + i] = a.q[(long long)env * Q + i];
    DFX_FOR(i, D) s[Y.qd + i] = a.qd[(long long)env * D + i];
    load_actuation(P, Y, s, g, env, a, true);
    for (int i = Y.qdd + D + g.lane; i < Y.q + Y.tape_row; i += Grp::G) s[i] = 0.0f;   // row padding
    if (g.lane == 0) s[Y.cmask] = 0.0f;
    DFX_FOR(i, P.L * 12) s[Y.fx + i] = 0.0f;   // fixed-point wrench accumulators (L x 6 low + high words)
    g.pre_store();
    g.sync();
    if constexpr (Grp::kFusedPhases) {      // K1 of the first substep (later ones: inside integrate_fwd)
        DFX_FOR(i, P.L) kin_local_fwd(P, Y, s, i);
        g.sync();
    }
    for (int sub = 0; sub < a.substeps; ++sub) {
        const bool upd = (sub % a.mm_freq) == 0;
        // (q, qd) ENTERING this substep: every thread fenced its writes before the barrier that closed the previous
        // integrate_fwd (or the loads above), so the store is issued without a barrier of its own
        if (a.tape) g.block_out_part(a.tape, sub, a.N, env, s + Y.q, rf, s + Y.stage, true, true);
        kin_fwd(P, Y, s, g);
        g.phase_sync();
        body_and_contact_fwd(P, Y, s, g);
        muscle_fwd(P, Y, s, g);
        wrench_collect(P, Y, s, g);
        g.phase_sync();
        tau_fwd(P, Y, s, g);
        g.phase_sync();
        if (a.has_derived && sub == a.substeps - 1) dump_derived(P, Y, s, a.derived, env, false, g);
        if (upd) {
            if (Y.overlay) g.sync();               // (Lm, Icmp) overlay (Xl, vj, f, fx): every reader of those is done
            crba_fwd(P, Y, s, g);
            if (a.has_derived && a.derived.H) DFX_FOR(e, DD) a.derived.H[(long long)env * DD + e] = s[Y.A + e];
            g.sync();
            chol_inverse(P, Y, s, g);
            if (a.has_derived && a.derived.L) DFX_FOR(e, DD) a.derived.L[(long long)env * DD + e] = s[Y.Lm + e];
            if (a.tape) g.block_out(a.tape + a.hinv_base, sub / a.mm_freq, a.N, env, s + Y.A, DD, false);
            if (Y.overlay) {                       // the fixed-point accumulators must read zero again
                g.sync();
                DFX_FOR(i, P.L * 12) s[Y.fx + i] = 0.0f;
            }
        }
        solve_fwd(P, Y, s, g);
        // the forward intermediates and q'': fenced before solve_fwd's closing barrier (a bf16 tape converts first and
        // fences again)
        if (a.tape) g.block_out_part(a.tape, sub, a.N, env, s + Y.q, rf, s + Y.stage, false, !rf.bf16);
        if (a.has_derived && sub == a.substeps - 1) dump_derived(P, Y, s, a.derived, env, true, g);
        g.phase_sync();
        g.store_sync();   // (synchronous tape stores / dumps above must finish before integrate_fwd overwrites q, qd)
        integrate_fwd(P, Y, s, a.dt_sub, g);
    }
    DFX_FOR(i, Q) a.q_out[(long long)env * Q + i] = s[Y.q + i];
    DFX_FOR(i, D) a.qd_out[(long long)env * D + i] = s[Y.qd + i];
}

template <class Grp>
DFX_HD void env_step_backward(const Pack& P, const Layout& Y, SP s, const Grp& g, int env, const StepArgs& a) {
    const int Q = P.Q, D = P.D, M = P.M, DD = D * D;
    const RowFmt rf = row_fmt(P, Y, a.tape_bf16 != 0);
    load_actuation(P, Y, s, g, env, a, false);
    DFX_FOR(i, D) s[Y.aact + i] = 0.0f;
    DFX_FOR(i, M) s[Y.amusc + i] = 0.0f;
    if (g.lane == 0) s[Y.cmask] = 0.0f;
    if (P.M > 0) DFX_FOR(i, P.L * 13) s[Y.fxH + i] = 0.0f;  // high words of the fixed-point cotangent accumulators
    DFX_FOR(i, Q) s[Y.aq + i] = a.gq_out ? a.gq_out[(long long)env * Q + i] : 0.0f;
    DFX_FOR(i, D) s[Y.aqd + i] = a.gqd_out ? a.gqd_out[(long long)env * D + i] : 0.0f;
    for (int sub = a.substeps - 1; sub >= 0; --sub) {
        const int seg = sub / a.mm_freq;
        const int s0 = seg * a.mm_freq;
        const bool seg_last = (sub == a.substeps - 1) || ((sub + 1) % a.mm_freq == 0);   // first visited of its segment
        // the row arrives in two asynchronous parts: (q, qd, q'') -- all that the first two adjoint phases read -- then
        // the rest, which substep_adj() waits for only before the phases that need it
        if constexpr (Grp::kBulkRows) {
            const bool want_hinv = seg_last && Y.A >= 0;
            g.rows_in(s + Y.q, a.tape_in, sub, rf, s + Y.stage, s + (want_hinv ? Y.A : 0),
                      want_hinv ? g.block_ptr(a.tape_in + a.hinv_base, seg, DD) : nullptr, DD);
            if (sub > 0 && !(a.flags & 128)) g.prefetch_row(a.tape_in, sub - 1, rf);
        } else {
            g.row_in(s + Y.q, a.tape_in, sub, a.N, env, rf.n, rf.early, rf.tail, true);
            if (seg_last && Y.A >= 0) g.block_in(s + Y.A, a.tape_in + a.hinv_base, seg, a.N, env, DD, false);
        }
        if (seg_last) DFX_FOR(e, dfx_sym_count(D)) s[Y.Lm + e] = 0.0f;      // symmetrised cotangent of H (packed)
        // layouts that do not stage H^-1 read its rows from the tape block in place (L2)
        const HinvView hv = (Y.A >= 0) ? HinvView{nullptr, 0} : g.hinv_view(a.tape_in + a.hinv_base, seg, a.N, env, DD);
        if constexpr (!Grp::kBulkRows) g.row_in(s + Y.q, a.tape_in, sub, a.N, env, rf.n, rf.early, rf.tail, false);
        g.copy_wait_first();
        g.sync();
        substep_adj(P, Y, s, a.dt_sub, sub == s0, hv, rf, g);
    }
    if (a.gq) DFX_FOR(i, Q) a.gq[(long long)env * Q + i] = s[Y.aq + i];
    if (a.gqd) DFX_FOR(i, D) a.gqd[(long long)env * D + i] = s[Y.aqd + i];
    if (a.gact) DFX_FOR(i, D) a.gact[(long long)env * D + i] = s[Y.aact + i];
    if (a.gmusc) DFX_FOR(i, M) a.gmusc[(long long)env * M + i] = s[Y.amusc + i];
    if (a.g_raw) {      // cotangent of the raw policy output (dfx_action_map_backward's arithmetic; clip passes the gradient on [-1, 1])
        const int A = a.map_num_act, src = (a.map_muscle ? Y.amusc : Y.aact) + a.map_offset;
        DFX_FOR(j, A) {
            const float r = a.raw[(long long)env * A + j];
            float gg = a.g_used ? a.g_used[(long long)env * A + j] : 0.0f;
            gg += (s[src + j] * a.strength[j]) * a.map_drive_scale;
            a.g_raw[(long long)env * A + j] = (r >= -1.0f && r <= 1.0f) ? gg * a.map_pre_scale : 0.0f;
        }
    }
}

}  // namespace dfx
