// dfx_tile.cu -- 32-environment tile kernels of the differentiable articulated rigid-body step (sm_100a).
//
// Execution model (small articulations: Ant, Hopper, HalfCheetah, CartPole): one CTA owns a TILE of 32
// environments for the whole env-step.  LANE = environment, WARP = item: the link / degree of freedom / contact
// point / matrix entry loops of the phase code (dfx_phases.h `DFX_FOR`) are strided over the warps of the CTA, and
// every lane of a warp does the same item for its own environment.  Consequences:
//   * full SIMD efficiency in every phase and no divergence on the joint type (an item has ONE type for all 32 lanes);
//   * the per-environment scratch is a structure-of-arrays tile in shared memory (element i of environment e at
//     (i * 32 + e), DFX_ES = 32): consecutive lanes hit consecutive banks, every pack read is a warp-wide broadcast;
//   * the tape of a tile is [block][tile][row][32]: a substep's rows of the 32 environments are ONE contiguous
//     53 KB block (Ant) that moves between HBM and the scratch tile with flat 16-byte cp.async / float4 copies;
//   * tree recursions cost one warp per link instead of one (mostly idle) lane group per environment and level;
//   * sparse work (penetrating contact points) is compacted over the whole tile first (cta_compact).
// The price: every group barrier is a CTA barrier (16 warps).  The phase code is the same source as the lane-group
// kernels (dfx_kernels.cu) and the host emulation: only the Group policy and the scratch pointer type differ.
#define DFX_ES 32
#include "dfx_launch.h"

namespace dfx {

template <int NW>
struct GroupTile {
    static constexpr int G = NW;
    static constexpr bool kPathPasses = true;    // barriers are CTA-wide: path / subtree passes instead of per-level recursions
    int lane;            // warp index: the item slot of DFX_FOR
    int lane32;          // environment inside the tile
    int tile, ntiles;
    float* tile_base;    // scratch tile (element 0, environment 0)
    int* task_count;
    int* task_list;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ void phase_sync() const {}
    __device__ __forceinline__ void atomic_or(unsigned* p, unsigned v) const { atomicOr(p, v); }
    __device__ __forceinline__ void fx_add(int* p, int v) const { atomicAdd(p, v); }
    __device__ __forceinline__ void atomic_add(float* p, float v) const { atomicAdd(p, v); }
    // max over the items (warps) of each environment, through the environment's slot
    __device__ __forceinline__ float group_max(float v, SP slot) const {
        if (lane == 0) sp_int(slot)[0] = 0;
        __syncthreads();
        atomicMax(&sp_int(slot)[0], __float_as_int(fmaxf(v, 0.0f)));   // non-negative floats order like ints
        __syncthreads();
        const float m = slot[0];
        __syncthreads();
        return m;
    }
    template <class F>
    __device__ __forceinline__ void cta_tasks(SP s, int n, bool lead, F f) const {
        (void)lead;      // every phase ends with sync(), which is already CTA-wide here
        for (int k = lane; k < n; k += NW) f(s, k);
        __syncthreads();
    }
    // (k, environment) pairs with pred() true, compacted over the tile, then f() over the list with full warps
    template <class Pr, class F>
    __device__ __forceinline__ void cta_compact(SP s, int n, Pr pred, F f) const {
        // (*task_count is 0 on entry: cleared at kernel start and after every use; the preceding phase's closing
        //  barrier has published what pred() reads)
        for (int k0 = 0; k0 < n; k0 += NW) {
            const int k = k0 + lane;
            const bool hit = (k < n) && pred(s, k);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                int base = 0;
                if (lane32 == 0) base = atomicAdd(task_count, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) task_list[base + __popc(m & ((1u << lane32) - 1u))] = (k << 5) | lane32;
            }
        }
        __syncthreads();
        const int hits = *task_count;
        for (int i = threadIdx.x; i < hits; i += NW * 32) {
            const int t = task_list[i];
            f(SP{tile_base + (t & 31)}, t >> 5);
        }
        __syncthreads();
        if (threadIdx.x == 0) *task_count = 0;
    }
    // cta_compact() and m dense items item(s, k) as ONE phase: after the compaction every warp first does its items
    // (k = warp, warp + NW, ...), then all warps drain the task list 32 tasks at a time through a shared cursor, so the
    // warps without an item start on the tasks at once and nobody idles while work is left
    static constexpr bool kConcurrentItems = true;
    template <class Pr, class F, class H>
    __device__ __forceinline__ void cta_compact_with(SP s, int n, Pr pred, F f, int m, H item) const {
        int* cursor = task_count + 1;                  // 0 on entry, like *task_count
        for (int k0 = 0; k0 < n; k0 += NW) {
            const int k = k0 + lane;
            const bool hit = (k < n) && pred(s, k);
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (b) {
                int base = 0;
                if (lane32 == 0) base = atomicAdd(task_count, __popc(b));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (hit) task_list[base + __popc(b & ((1u << lane32) - 1u))] = (k << 5) | lane32;
            }
        }
        __syncthreads();
        const int hits = *task_count;
        for (int k = lane; k < m; k += NW) item(s, k);
        for (;;) {
            int base = 0;
            if (lane32 == 0) base = atomicAdd(cursor, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= hits) break;
            if (base + lane32 < hits) {
                const int t = task_list[base + lane32];
                f(SP{tile_base + (t & 31)}, t >> 5);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { *task_count = 0; *cursor = 0; }
    }
    // tape blocks [b][tile][n][32]: the scratch tile of the block is the same bytes, one flat copy by the whole CTA
    __device__ __forceinline__ void block_in(SP dst, const float* base, long long b, int N, int env, int n, bool rows) const {
        (void)N; (void)env; (void)rows;
        const float* src = base + ((b * ntiles + tile) * n) * 32;
        const unsigned d = (unsigned)__cvta_generic_to_shared(sp_raw(dst) - lane32);
        for (int i = threadIdx.x * 4; i < n * 32; i += NW * 32 * 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void block_out(float* base, long long b, int N, int env, SP src, int n, bool rows) const {
        (void)N; (void)env; (void)rows;
        float4* d4 = reinterpret_cast<float4*>(base + ((b * ntiles + tile) * n) * 32);
        const float4* s4 = reinterpret_cast<const float4*>(sp_raw(src) - lane32);
        for (int i = threadIdx.x; i < n * 8; i += NW * 32) d4[i] = s4[i];
    }
    // a row tile in two commit groups: the [0, head) and [tail, n) element ranges first, the middle second
    __device__ __forceinline__ void row_in(SP dst, const float* base, long long b, int N, int env, int n, int head, int tail, bool first) const {
        (void)N; (void)env;
        const float* src = base + ((b * ntiles + tile) * n) * 32;
        const unsigned d = (unsigned)__cvta_generic_to_shared(sp_raw(dst) - lane32);
        if (first) {
            for (int i = threadIdx.x * 4; i < head * 32; i += NW * 32 * 4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
            for (int i = tail * 32 + threadIdx.x * 4; i < n * 32; i += NW * 32 * 4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
        } else {
            for (int i = head * 32 + threadIdx.x * 4; i < tail * 32; i += NW * 32 * 4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4), "l"(src + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void copy_wait_first() const { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    __device__ __forceinline__ void copy_wait_all() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
};

template <int NW, bool BACKWARD, int SL, int SD, int SQ, int SC, int SM>
__global__ void __launch_bounds__(NW * 32) dfx_tile_kernel(const __grid_constant__ KernelArgs ka) {
    extern __shared__ __align__(128) float smem[];
    float* fpack = smem;
    int* ipack = reinterpret_cast<int*>(smem + ((ka.blob.n_floats + 3) & ~3));
    for (int i = threadIdx.x; i < ka.blob.n_floats; i += NW * 32) fpack[i] = ka.blob.floats[i];
    for (int i = threadIdx.x; i < ka.blob.n_ints; i += NW * 32) ipack[i] = ka.blob.ints[i];
    if (threadIdx.x < 4) reinterpret_cast<int*>(smem + ka.pack_smem_floats)[threadIdx.x] = 0;   // task counter / cursor of cta_compact
    __syncthreads();
    Pack P = bind_pack(ka.header, ka.blob, ipack, fpack);
    P.L = SL; P.D = SD; P.Q = SQ; P.C = SC; P.M = SM;
    constexpr Layout Y = make_layout(SL, SD, SQ, SC, SM);

    GroupTile<NW> g;
    g.lane = threadIdx.x >> 5;
    g.lane32 = threadIdx.x & 31;
    g.tile = blockIdx.x;
    g.ntiles = gridDim.x;
    g.task_count = reinterpret_cast<int*>(smem + ka.pack_smem_floats);
    g.task_list = g.task_count + 4;
    g.tile_base = smem + ka.pack_smem_floats + ka.cta_area_floats;
    // lanes past the end redo the last environment: identical values are stored twice, every thread reaches
    // every barrier, and the (padded) tape tile slots of those lanes are simply never read by anybody else
    int env = blockIdx.x * 32 + g.lane32;
    if (env >= ka.step.N) env = ka.step.N - 1;
    const SP s{g.tile_base + g.lane32};
    if (BACKWARD) env_step_backward(P, Y, s, g, env, ka.step);
    else env_step_forward(P, Y, s, g, env, ka.step);
}

// 128-byte aligned offsets so that the flat tile copies stay 16-byte aligned
static int tile_pack_floats(const KernelArgs& ka) {
    const int bytes = (((ka.blob.n_floats + 3) & ~3) + ((ka.blob.n_ints + 3) & ~3)) * 4;
    return ((bytes + 127) & ~127) / 4;
}
static int tile_area_floats(const KernelArgs& ka) { return ((16 + ka.header.C * 32 * 4 + 127) & ~127) / 4; }

template <int NW, bool BWD, int SL, int SD, int SQ, int SC, int SM>
static cudaError_t tile_launch_impl(KernelArgs& ka, cudaStream_t stream) {
    constexpr Layout Y = make_layout(SL, SD, SQ, SC, SM);
    const int per_env = ((BWD ? Y.bwd_size : Y.fwd_size) + 3) & ~3;
    ka.layout = Y;
    ka.pack_smem_floats = tile_pack_floats(ka);
    ka.cta_area_floats = tile_area_floats(ka);
    ka.scratch_stride = per_env;
    const size_t smem = (size_t)(ka.pack_smem_floats + ka.cta_area_floats + per_env * 32) * sizeof(float);
    auto kern = dfx_tile_kernel<NW, BWD, SL, SD, SQ, SC, SM>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int ntiles = (ka.step.N + 31) / 32;
    kern<<<ntiles, NW * 32, smem, stream>>>(ka);
    return cudaGetLastError();
}

}  // namespace dfx

using namespace dfx;

// the articulations with a tile kernel: (warps forward, warps backward, L, D, Q, C, M)
#ifndef DFX_TILE_NWF
#define DFX_TILE_NWF 16
#endif
#ifndef DFX_TILE_NWB
#define DFX_TILE_NWB 16
#endif
#ifdef DFX_TILE_ONLY_ANT      // (tuning builds)
#define DFX_TILE_MODELS(X) X(DFX_TILE_NWF, DFX_TILE_NWB, 9, 14, 15, 25, 0)
#else
#define DFX_TILE_MODELS(X)       \
    X(DFX_TILE_NWF, DFX_TILE_NWB, 9, 14, 15, 25, 0) /* Ant */         \
    X(8, 8, 3, 2, 2, 0, 0)     /* CartPole */    \
    X(16, 16, 6, 6, 6, 8, 0)    /* Hopper */      \
    X(16, 16, 9, 9, 9, 16, 0)   /* HalfCheetah */
#endif

bool dfx_tile_supported(int L, int D, int Q, int C, int M) {
#define X(nwf, nwb, l, d, q, c, m) if (L == l && D == d && Q == q && C == c && M == m) return true;
    DFX_TILE_MODELS(X)
#undef X
    return false;
}

size_t dfx_tile_smem(const KernelArgs& ka, bool backward) {
    const Layout Y = make_layout(ka.header.L, ka.header.D, ka.header.Q, ka.header.C, ka.header.M);
    const int per_env = ((backward ? Y.bwd_size : Y.fwd_size) + 3) & ~3;
    return (size_t)(tile_pack_floats(ka) + tile_area_floats(ka) + per_env * 32) * sizeof(float);
}

cudaError_t dfx_tile_launch(KernelArgs& ka, bool backward, cudaStream_t stream) {
    const Pack& h = ka.header;
#define X(nwf, nwb, l, d, q, c, m)                                                            \
    if (h.L == l && h.D == d && h.Q == q && h.C == c && h.M == m)                               \
        return backward ? tile_launch_impl<nwb, true, l, d, q, c, m>(ka, stream)              \
                        : tile_launch_impl<nwf, false, l, d, q, c, m>(ka, stream);
    DFX_TILE_MODELS(X)
#undef X
    return cudaErrorInvalidConfiguration;
}
