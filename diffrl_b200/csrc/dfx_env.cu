// dfx_env.cu -- fused observation / reward / termination epilogue of the free-root walker envs
// (Ant, Humanoid, SNU humanoid) and its adjoint: SURVEY.md section 8f row 1.
//
// In the reference these are ~35 small PyTorch ops per env.step() plus ~60 autograd ops in backward
// (envs/ant.py:266-307, envs/humanoid.py:314-368, envs/snu_humanoid.py:378-432): launch-bound once the
// simulation step itself takes a few hundred microseconds.  Here: one thread per environment, one
// launch forward, one launch backward, straight from / to the env-major (q, qd, actions) rows.
#include <cuda_runtime.h>

#include "../../include/dfx.h"
#include "dfx_env_dev.h"

using namespace dfx;

namespace {

__global__ void walker_forward_kernel(DfxWalkerParams p, int n, const float* __restrict__ q, const float* __restrict__ qd,
                                      const float* __restrict__ actions, const long long* __restrict__ progress,
                                      float* __restrict__ obs, float* __restrict__ rew, long long* __restrict__ reset) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float r = 0.0f;
    long long rs = 0;
    walker_eval(p, q + (size_t)e * p.num_q, qd + (size_t)e * p.num_qd, actions + (size_t)e * p.num_act,
                rew ? progress[e] : 0, rew != nullptr, obs + (size_t)e * p.num_obs, &r, &rs);
    if (rew) { rew[e] = r; reset[e] = rs; }
}

__global__ void walker_backward_kernel(DfxWalkerParams p, int n, const float* __restrict__ q, const float* __restrict__ qd,
                                       const float* __restrict__ actions, const float* __restrict__ g_obs,
                                       const float* __restrict__ g_rew, float* __restrict__ gq, float* __restrict__ gqd,
                                       float* __restrict__ gact) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    walker_eval_adj(p, q + (size_t)e * p.num_q, qd + (size_t)e * p.num_qd, actions + (size_t)e * p.num_act,
                    g_obs + (size_t)e * p.num_obs, g_obs != nullptr, g_obs, false, g_rew ? g_rew[e] : 0.0f, g_rew != nullptr,
                    gq + (size_t)e * p.num_q, gqd + (size_t)e * p.num_qd, gact + (size_t)e * p.num_act, gact != nullptr);
}

// ---- the whole env transition after the simulation step as its own launch, one thread per environment (the per-environment
// code is dfx_env_dev.h; the tile kernels run the same code inside the simulation launch: dfx_env_step_forward / _backward)
#define DFX_TRANSITION_KERNELS(name, Params)                                                                                      \
    __global__ void name##_transition_forward_kernel(Params p, int n, const float* __restrict__ q, const float* __restrict__ qd, \
                                                     const float* __restrict__ actions, const long long* __restrict__ progress,  \
                                                     const float* __restrict__ start_q, const float* __restrict__ start_qd,     \
                                                     float* __restrict__ obs_before, float* __restrict__ rew,                    \
                                                     long long* __restrict__ reset, float* __restrict__ q_next,                  \
                                                     float* __restrict__ qd_next, float* __restrict__ actions_next,              \
                                                     long long* __restrict__ progress_next, float* __restrict__ obs_next) {      \
        const int e = blockIdx.x * blockDim.x + threadIdx.x;                                                                      \
        if (e >= n) return;                                                                                                       \
        name##_transition_forward_env(p, e, q + (size_t)e * p.num_q, qd + (size_t)e * p.num_qd, actions, progress, start_q,      \
                                      start_qd, obs_before, rew, reset, q_next, qd_next, actions_next, progress_next, obs_next);  \
    }                                                                                                                             \
    __global__ void name##_transition_backward_kernel(Params p, int n, const float* __restrict__ q, const float* __restrict__ qd,\
                                                      const float* __restrict__ actions, const long long* __restrict__ reset,    \
                                                      const float* __restrict__ g_obs_before, const float* __restrict__ g_rew,   \
                                                      const float* __restrict__ g_q_next, const float* __restrict__ g_qd_next,   \
                                                      const float* __restrict__ g_actions_next,                                   \
                                                      const float* __restrict__ g_obs_next, float* __restrict__ gq,              \
                                                      float* __restrict__ gqd, float* __restrict__ gact) {                        \
        const int e = blockIdx.x * blockDim.x + threadIdx.x;                                                                      \
        if (e >= n) return;                                                                                                       \
        name##_transition_backward_env(p, e, q, qd, actions, reset, g_obs_before, g_rew, g_q_next, g_qd_next, g_actions_next,    \
                                       g_obs_next, gq, gqd, gact);                                                                \
    }
DFX_TRANSITION_KERNELS(walker, DfxWalkerParams)
DFX_TRANSITION_KERNELS(planar, DfxPlanarParams)
#undef DFX_TRANSITION_KERNELS

// ---- policy output -> actuation, one thread per (environment, action):
//   u = clip(a, -1, 1) * pre_scale + pre_bias          (what the env keeps as `actions`: observation + penalty)
//   drive[e, offset + j] = (u * drive_scale) * strength[j]   (joint_act row of width `width`, or the muscle activations;
//                                                        same association as the reference's actions * scale * strengths)
// Reference: torch.clip + per-env scaling + slice assignment, envs/ant.py:156-166, snu_humanoid.py:283-296.
__global__ void action_map_forward_kernel(int n, int num_act, int width, int offset, float pre_scale, float pre_bias, float drive_scale,
                                          const float* __restrict__ strength, const float* __restrict__ raw,
                                          float* __restrict__ used, float* __restrict__ drive) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    const int e = t / width, c = t - e * width, j = c - offset;
    float out = 0.0f;
    if (j >= 0 && j < num_act) {
        const float a = raw[(size_t)e * num_act + j];
        const float u = fminf(fmaxf(a, -1.0f), 1.0f) * pre_scale + pre_bias;
        used[(size_t)e * num_act + j] = u;
        out = (u * drive_scale) * strength[j];
    }
    drive[t] = out;
}
// g_used (nullable), g_drive (nullable) -> g_raw; torch.clip passes the gradient on [-1, 1] including the ends
__global__ void action_map_backward_kernel(int n, int num_act, int width, int offset, float pre_scale, float drive_scale,
                                           const float* __restrict__ strength, const float* __restrict__ raw,
                                           const float* __restrict__ g_used, const float* __restrict__ g_drive,
                                           float* __restrict__ g_raw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * num_act) return;
    const int e = t / num_act, j = t - e * num_act;
    const float a = raw[t];
    float g = g_used ? g_used[t] : 0.0f;
    if (g_drive) g += (g_drive[(size_t)e * width + offset + j] * strength[j]) * drive_scale;
    g_raw[t] = (a >= -1.0f && a <= 1.0f) ? g * pre_scale : 0.0f;
}

}  // namespace

extern long long dfx_count_launch(void);
// one warp per CTA: a thread walks ~100 row entries serially, so the kernels are latency bound and 4096 environments
// should cover 128 SMs rather than 32
static constexpr int kEnvThreads = 32;

extern "C" {

int dfx_walker_obs_forward(const DfxWalkerParams* p, int n, const float* q, const float* qd, const float* actions,
                           const long long* progress, float* obs, float* rew, long long* reset, void* stream) {
    if (!p || n <= 0 || !q || !qd || !actions || !obs || (rew && (!reset || !progress))) return (int)cudaErrorInvalidValue;
    walker_forward_kernel<<<(n + kEnvThreads - 1) / kEnvThreads, kEnvThreads, 0, (cudaStream_t)stream>>>(*p, n, q, qd, actions, progress, obs, rew, reset);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_walker_obs_backward(const DfxWalkerParams* p, int n, const float* q, const float* qd, const float* actions,
                            const float* g_obs, const float* g_rew, float* gq, float* gqd, float* gact, void* stream) {
    if (!p || n <= 0 || !q || !qd || !actions || !gq || !gqd) return (int)cudaErrorInvalidValue;
    walker_backward_kernel<<<(n + kEnvThreads - 1) / kEnvThreads, kEnvThreads, 0, (cudaStream_t)stream>>>(*p, n, q, qd, actions, g_obs, g_rew, gq, gqd, gact);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_walker_transition_forward(const DfxWalkerParams* p, int n, const float* q, const float* qd, const float* actions,
                                  const long long* progress, const float* start_q, const float* start_qd,
                                  float* obs_before, float* rew, long long* reset, float* q_next, float* qd_next,
                                  float* actions_next, long long* progress_next, float* obs_next, void* stream) {
    if (!p || n <= 0 || !q || !qd || !actions || !progress || !start_q || !start_qd || !obs_before || !rew || !reset ||
        !q_next || !qd_next || !actions_next || !progress_next || !obs_next || p->num_obs > kMaxObs)
        return (int)cudaErrorInvalidValue;
    walker_transition_forward_kernel<<<(n + kEnvThreads - 1) / kEnvThreads, kEnvThreads, 0, (cudaStream_t)stream>>>(
        *p, n, q, qd, actions, progress, start_q, start_qd, obs_before, rew, reset, q_next, qd_next, actions_next,
        progress_next, obs_next);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_walker_transition_backward(const DfxWalkerParams* p, int n, const float* q, const float* qd, const float* actions,
                                   const long long* reset, const float* g_obs_before, const float* g_rew,
                                   const float* g_q_next, const float* g_qd_next, const float* g_actions_next,
                                   const float* g_obs_next, float* gq, float* gqd, float* gact, void* stream) {
    if (!p || n <= 0 || !q || !qd || !actions || !reset || !gq || !gqd) return (int)cudaErrorInvalidValue;
    walker_transition_backward_kernel<<<(n + kEnvThreads - 1) / kEnvThreads, kEnvThreads, 0, (cudaStream_t)stream>>>(
        *p, n, q, qd, actions, reset, g_obs_before, g_rew, g_q_next, g_qd_next, g_actions_next, g_obs_next, gq, gqd, gact);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_planar_transition_forward(const DfxPlanarParams* p, int n, const float* q, const float* qd, const float* actions,
                                  const long long* progress, const float* start_q, const float* start_qd,
                                  float* obs_before, float* rew, long long* reset, float* q_next, float* qd_next,
                                  float* actions_next, long long* progress_next, float* obs_next, void* stream) {
    if (!p || n <= 0 || !q || !qd || !actions || !progress || !start_q || !start_qd || !obs_before || !rew || !reset ||
        !q_next || !qd_next || !actions_next || !progress_next || !obs_next || p->kind < 0 || p->kind > 2)
        return (int)cudaErrorInvalidValue;
    planar_transition_forward_kernel<<<(n + kEnvThreads - 1) / kEnvThreads, kEnvThreads, 0, (cudaStream_t)stream>>>(
        *p, n, q, qd, actions, progress, start_q, start_qd, obs_before, rew, reset, q_next, qd_next, actions_next,
        progress_next, obs_next);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_planar_transition_backward(const DfxPlanarParams* p, int n, const float* q, const float* qd, const float* actions,
                                   const long long* reset, const float* g_obs_before, const float* g_rew,
                                   const float* g_q_next, const float* g_qd_next, const float* g_actions_next,
                                   const float* g_obs_next, float* gq, float* gqd, float* gact, void* stream) {
    if (!p || n <= 0 || !q || !qd || !actions || !reset || !gq || !gqd || p->kind < 0 || p->kind > 2) return (int)cudaErrorInvalidValue;
    planar_transition_backward_kernel<<<(n + kEnvThreads - 1) / kEnvThreads, kEnvThreads, 0, (cudaStream_t)stream>>>(
        *p, n, q, qd, actions, reset, g_obs_before, g_rew, g_q_next, g_qd_next, g_actions_next, g_obs_next, gq, gqd, gact);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_action_map_forward(int n, int num_act, int width, int offset, float pre_scale, float pre_bias, float drive_scale,
                           const float* strength, const float* raw, float* used, float* drive, void* stream) {
    if (n <= 0 || num_act <= 0 || width < num_act || offset < 0 || offset + num_act > width || !strength || !raw || !used || !drive)
        return (int)cudaErrorInvalidValue;
    const long long total = (long long)n * width;
    action_map_forward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, num_act, width, offset, pre_scale, pre_bias, drive_scale, strength, raw, used, drive);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

int dfx_action_map_backward(int n, int num_act, int width, int offset, float pre_scale, float drive_scale, const float* strength,
                            const float* raw, const float* g_used, const float* g_drive, float* g_raw, void* stream) {
    if (n <= 0 || num_act <= 0 || width < num_act || offset < 0 || offset + num_act > width || !strength || !raw || !g_raw)
        return (int)cudaErrorInvalidValue;
    const long long total = (long long)n * num_act;
    action_map_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, num_act, width, offset, pre_scale, drive_scale, strength, raw, g_used, g_drive, g_raw);
    dfx_count_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
