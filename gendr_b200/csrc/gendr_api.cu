// gendr_api.cu -- host side of libgendr_b200.so: the C ABI declared in include/gendr_b200.h.
// Derives the launch constants, owns the workspace layout, dispatches to the per-distribution kernel launchers
// (inst_dist.cu) and reports errors.  No torch, no CPU compute path: every entry point runs CUDA kernels.
#include "../../include/gendr_b200.h"
#include "render_kernels.cuh"
#include "scene_kernels.cuh"
#include "voxel_kernels.cuh"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>

namespace gendr {
#define GENDR_DECL(k) cudaError_t launch_render_dist_##k(const RenderParams&, const KernelIO&, const LaunchCfg&);
GENDR_DECL(0) GENDR_DECL(1) GENDR_DECL(2) GENDR_DECL(3) GENDR_DECL(4) GENDR_DECL(5) GENDR_DECL(6) GENDR_DECL(7) GENDR_DECL(8)
GENDR_DECL(9) GENDR_DECL(10) GENDR_DECL(11) GENDR_DECL(12) GENDR_DECL(13) GENDR_DECL(14) GENDR_DECL(15) GENDR_DECL(16) GENDR_DECL(17)
#undef GENDR_DECL
static const render_launch_fn kLaunchTable[D_COUNT] = {
    launch_render_dist_0,  launch_render_dist_1,  launch_render_dist_2,  launch_render_dist_3,  launch_render_dist_4,
    launch_render_dist_5,  launch_render_dist_6,  launch_render_dist_7,  launch_render_dist_8,  launch_render_dist_9,
    launch_render_dist_10, launch_render_dist_11, launch_render_dist_12, launch_render_dist_13, launch_render_dist_14,
    launch_render_dist_15, launch_render_dist_16, launch_render_dist_17};

// ---- longest-first CTA schedule (cta_to_tile, render_kernels.cuh) -----------------------------------------------------------------
// count_tiles: the face adds 1 to every tile its candidate rectangle overlaps -- the tile's cost estimate.  Faces that reach more
// than 128 tiles (dense regimes: the rectangle is the screen) are a uniform load and are left out (no 10^8 atomics in C4).
__device__ __forceinline__ void count_tiles(const RenderParams& P, unsigned* __restrict__ counts, long long b, uint32_t wA, uint32_t wB) {
    const int ix0 = (int)(wA & PIX_MASK), ix1 = (int)((wA >> 16) & PIX_MASK), iy0 = (int)(wB & PIX_MASK), iy1 = (int)((wB >> 16) & PIX_MASK);
    if (ix1 < ix0 || iy1 < iy0) return;
    const int tx0 = ix0 / TILE_W, tx1 = min(ix1 / TILE_W, P.tiles_x - 1), ty0 = iy0 / TILE_H, ty1 = min(iy1 / TILE_H, P.tiles_y - 1);
    if (tx1 < tx0 || ty1 < ty0 || (tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 128) return;
    unsigned* c = counts + b * (long long)(P.tiles_x * P.tiles_y);
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(c + ty * P.tiles_x + tx, 1u);
}
// tile_order_kernel: counting sort of the n = B * tiles (item, tile) pairs by (group of `group` batch items ascending, count
// descending), one CTA (n <= 2^18: a few microseconds).  Longest-first WITHIN a group of items rather than over the whole batch:
// the CTAs in flight then read the face records of one or two groups (tens of MB, L2-resident) instead of every item's -- sorted
// over the whole batch the C3 kernels re-read 1.4-1.7x more from DRAM -- and the kernel still ends on a group's lightest tiles.
// At most 32 groups x 256 count bins (exact below 128 faces, 16 faces per bin above); ties in arbitrary order.
constexpr int kOrderGroups = 32, kOrderBins = 256;
__device__ __forceinline__ unsigned tile_bin(unsigned c) {
    const unsigned k = c < 128u ? c : 128u + min((c - 128u) >> 4, 127u);
    return (unsigned)(kOrderBins - 1) - k;
}
__global__ void __launch_bounds__(1024) tile_order_kernel(const unsigned* __restrict__ counts, unsigned* __restrict__ order, int n, int tiles,
                                                          int group) {
    constexpr int NB = kOrderGroups * kOrderBins, PER = NB / 1024;
    __shared__ unsigned hist[NB];
    __shared__ unsigned warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int per_group = tiles * group;
#pragma unroll
    for (int k = 0; k < PER; ++k) hist[tid + 1024 * k] = 0u;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) atomicAdd(&hist[(i / per_group) * kOrderBins + tile_bin(counts[i])], 1u);
    __syncthreads();
    unsigned h[PER], mine = 0u;      // thread t owns bins PER*t .. PER*t + PER-1
#pragma unroll
    for (int k = 0; k < PER; ++k) { h[k] = hist[PER * tid + k]; mine += h[k]; }
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) warp_sum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const unsigned w = warp_sum[tid];
        unsigned x = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += u; }
        warp_sum[tid] = x - w;
    }
    __syncthreads();
    unsigned excl = warp_sum[tid >> 5] + incl - mine;
#pragma unroll
    for (int k = 0; k < PER; ++k) { hist[PER * tid + k] = excl; excl += h[k]; }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) order[atomicAdd(&hist[(i / per_group) * kOrderBins + tile_bin(counts[i])], 1u)] = (unsigned)i;
}

// ---- preprocessing kernel: one thread per (batch, face) -------------------------------------------------------
__global__ void __launch_bounds__(256) prep_kernel(const __grid_constant__ RenderParams P, const float* __restrict__ faces,
                                                   float* __restrict__ records, uint2* __restrict__ rects,
                                                   float* __restrict__ faces_info, unsigned* __restrict__ tile_counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)P.B * P.F) return;
    float v[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = __ldg(faces + i * 9 + k);
    float rec[REC_WORDS];
    prep_face_record(v, rec, faces_info ? faces_info + i * 27 : nullptr, P);
    float4* dst = reinterpret_cast<float4*>(records + i * REC_WORDS);
#pragma unroll
    for (int k = 0; k < REC_WORDS / 4; ++k) dst[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
    rects[i] = make_uint2(__float_as_uint(rec[R_PACK]), __float_as_uint(rec[R_PACK + 1]));
    if (tile_counts) count_tiles(P, tile_counts, i / P.F, __float_as_uint(rec[R_PACK]), __float_as_uint(rec[R_PACK + 1]));
}

// indexed variant: gathers the face's three vertices from vertices[B,V,3] through face_index (fused vertices[faces])
__global__ void __launch_bounds__(256) prep_indexed_kernel(const __grid_constant__ RenderParams P, const float* __restrict__ vertices,
                                                           const int* __restrict__ face_index, long long index_batch_stride, int V,
                                                           float* __restrict__ records, uint2* __restrict__ rects,
                                                           unsigned* __restrict__ tile_counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)P.B * P.F) return;
    const long long b = i / P.F, f = i - b * P.F;
    float v[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int vi = __ldg(face_index + b * index_batch_stride + f * 3 + k);
        vi = min(max(vi, 0), V - 1);
        const float* src = vertices + (b * V + vi) * 3;
        v[3 * k] = __ldg(src); v[3 * k + 1] = __ldg(src + 1); v[3 * k + 2] = __ldg(src + 2);
    }
    float rec[REC_WORDS];
    prep_face_record(v, rec, nullptr, P);
    float4* dst = reinterpret_cast<float4*>(records + i * REC_WORDS);
#pragma unroll
    for (int k = 0; k < REC_WORDS / 4; ++k) dst[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
    rects[i] = make_uint2(__float_as_uint(rec[R_PACK]), __float_as_uint(rec[R_PACK + 1]));
    if (tile_counts) count_tiles(P, tile_counts, b, __float_as_uint(rec[R_PACK]), __float_as_uint(rec[R_PACK + 1]));
}

// zero-fill of the two gradient buffers of a backward call in ONE launch (two cudaMemsetAsync cost two launches; tiny scenes are
// launch bound).  n_a / n_b in floats; either buffer may be null.
__global__ void __launch_bounds__(256) zero2_kernel(float* __restrict__ a, long long n_a, float* __restrict__ b, long long n_b) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_a + n_b; i += stride) {
        if (i < n_a) a[i] = 0.f; else b[i - n_a] = 0.f;
    }
}

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* what) {
    if (code >= GENDR_ERR_INVALID_ARGUMENT) snprintf(g_err, sizeof g_err, "gendr_b200: %s", what);
    else snprintf(g_err, sizeof g_err, "gendr_b200: %s: %s", what, cudaGetErrorString((cudaError_t)code));
    return code;
}
#define GENDR_CUDA(call, what)                                      \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return fail((int)e__, what);        \
    } while (0)

// NDC radius beyond which an OUTSIDE pixel (sign = -1) is guaranteed to fall under the 1e-6 probability threshold
// of K.cu:784 (in units of x; squared distances handled by the caller).  INF = never.  Constants are the analytic
// crossing points, nudged up where the fp32 evaluation is quantised near the threshold.
static double cull_x_over_tau(int dist, double shift) {
    const double INF = std::numeric_limits<double>::infinity();
    switch (dist) {
    case D_HARD: return 0.0;
    case D_UNIFORM: case D_CUBIC_HERMITE: case D_WIGNER: return 1.0;
    case D_GAUSSIAN: return 4.7535;
    case D_LAPLACE: return 13.1225;
    case D_LOGISTIC: return 13.8156;
    case D_GUDERMANNIAN: return 13.37;
    case D_CAUCHY: return 3.5e5;
    case D_RECIPROCAL: return 5.5e5;
    case D_GUMBEL_MAX: return 2.6259;
    case D_GUMBEL_MIN: return 13.9;
    case D_EXPONENTIAL: case D_GAMMA: case D_LEVY: return shift > 0 ? shift : 0.0;
    case D_EXPONENTIAL_REV: return (13.9 - shift) > 0 ? (13.9 - shift) : 0.0;
    case D_GAMMA_REV: return (15.01 - shift) > 0 ? (15.01 - shift) : 0.0;
    case D_LEVY_REV: return INF;
    }
    return INF;
}

static int make_params(RenderParams& P, int B, int F, int T, const gendr_render_params* u) {
    if (!u) return GENDR_ERR_INVALID_ARGUMENT;
    if (B < 0 || F < 0 || T < 1 || u->image_size < 1 || u->image_size > 16383) return GENDR_ERR_INVALID_ARGUMENT;
    if ((long long)B * F >= (1ll << 31) || (long long)B * u->image_size * u->image_size >= (1ll << 31)) return GENDR_ERR_INVALID_ARGUMENT;
    if (u->dist_func < 0 || u->dist_func >= D_COUNT) return GENDR_ERR_INVALID_ARGUMENT;
    if (u->aggr_alpha_func < 0 || u->aggr_alpha_func >= T_COUNT) return GENDR_ERR_INVALID_ARGUMENT;
    // aggr_rgb_func: 0 hard / 1 softmax; texture_type: 0 surface / 1 vertex (functional/renderer.py:81-83, renderer.py:39-42 raise
    // ValueError for anything else; a stray id would silently render the background here).  Vertex textures are exactly 3 colours
    // per face (K.cu:186-190 reads texture[0..8]).  Surface textures with a non-square T are accepted like the reference accepts
    // them (texture_res = int(sqrt(T)), K.cu:1098: T = 2, 3 behave as texture_res 1 with texel indices 0 and 1 inside the face).
    if (u->aggr_rgb_func < 0 || u->aggr_rgb_func > 1) return GENDR_ERR_INVALID_ARGUMENT;
    if (u->texture_type < 0 || u->texture_type > 1) return GENDR_ERR_INVALID_ARGUMENT;
    if (u->texture_type == 1 && T != 3) return GENDR_ERR_INVALID_ARGUMENT;
    memset(&P, 0, sizeof P);
    P.B = B; P.F = F; P.S = u->image_size; P.T = T;
    P.R = (int)sqrt((double)T);                          // K.cu:1098
    P.dist_func = u->dist_func; P.aggr_alpha_func = u->aggr_alpha_func; P.aggr_rgb_func = u->aggr_rgb_func;
    P.texture_type = u->texture_type; P.dist_squared = u->dist_squared ? 1 : 0; P.double_side = u->double_side ? 1 : 0;
    P.dist_scale = u->dist_scale; P.dist_shape = u->dist_shape; P.dist_shift = u->dist_shift; P.dist_eps = u->dist_eps;
    P.tcn_p = u->aggr_alpha_t_conorm_p; P.rgb_eps = u->aggr_rgb_eps; P.rgb_gamma = u->aggr_rgb_gamma;
    P.near_ = u->near_plane; P.far_ = u->far_plane;
    P.bg[0] = u->background[0]; P.bg[1] = u->background[1]; P.bg[2] = u->background[2];
    P.thr = u->dist_eps * u->dist_scale;                 // fp32 product, K.cu:725
    P.sqrt_thr = sqrtf(P.thr);                           // K.cu:747 (IEEE sqrt on both sides)
    double c = cull_x_over_tau(u->dist_func, (double)u->dist_shift);
    double r = c * (double)u->dist_scale;
    if (u->dist_squared) r = std::sqrt(r);
    if (!(u->dist_scale > 0.f) || !(r == r)) r = std::numeric_limits<double>::infinity();
    P.cull_radius = (float)r;
    P.cull_d2 = (float)((r * 1.01) * (r * 1.01) * 1.0001);
    P.gamma_kummer0 = (float)(1. / std::tgamma((double)u->dist_shape + 1.));
    P.gamma_lcoef = (float)((double)u->dist_shape * std::log(1. / (double)u->dist_scale) - std::lgamma((double)u->dist_shape));
    P.inv_tcn_p = (float)(1. / (double)u->aggr_alpha_t_conorm_p);
    P.zrange = P.far_ - P.near_;                         // fp32 subtraction, as the kernels' gd_sub(far, near) was
    P.y_tau = 1.0f / P.dist_scale; P.y_gamma = 1.0f / P.rgb_gamma; P.y_zrange = 1.0f / P.zrange;      // IEEE: correctly rounded
    P.k_zs = (float)(1.4426950408889634 / (double)P.rgb_gamma); P.k_cz = (float)(1.0 / ((double)P.rgb_gamma * (double)P.zrange));
    P.cta_group = 16;
    if (const char* ev = getenv("GENDR_B200_CTA_GROUP")) P.cta_group = std::max(1, atoi(ev));      // tuning experiments
    P.consts_ok = (rcp_in_range(P.dist_scale) && rcp_in_range(P.rgb_gamma) && rcp_in_range(P.zrange) && fabsf(P.near_) < 1e30f && fabsf(P.far_) < 1e30f) ? 1 : 0;
    P.tiles_x = (P.S + TILE_W - 1) / TILE_W; P.tiles_y = (P.S + TILE_H - 1) / TILE_H;
    return 0;
}

static size_t smem_bytes(const RenderParams&, bool face_stationary_backward) {
    return face_stationary_backward ? smem_bytes_total<BWD_WAVE, NPIX_BWD>() : smem_bytes_total<WAVE_FACES, 0>();
}

// Backward kernel choice.  The face-stationary kernel (one reduction per face per CTA, pixel state in shared memory, faces dealt
// round-robin to the warps) is the backward pass: measured on this round's final kernels it wins or ties in every regime --
// C4 (dense) 52.1 vs 58.7 ms at B = 8, C3 (sparse) 5.74 vs 5.86 ms at B = 64, C2 0.75 vs 0.76 ms (profiles/r2_kernel_ab_timings.jsonl).
// The pixel-stationary variant (pixel state in registers, one reduction per face per WARP) stays compiled for A/B measurements:
// GENDR_B200_BWD=ps selects it.
static int forced_backward_mode() {
    static const int forced = [] {
        const char* e = getenv("GENDR_B200_BWD");
        return (e && strcmp(e, "ps") == 0) ? 1 : ((e && strcmp(e, "fs") == 0) ? 0 : -1);
    }();
    return forced;
}
static int backward_mode(const RenderParams&) {
    const int forced = forced_backward_mode();
    return forced >= 0 ? forced : 0;
}

struct DeviceScope {   // run on the device that owns the data, restore the caller's device afterwards
    int prev = -1, cur = -1;
    cudaError_t enter(const void* p) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        cur = prev;
        cudaPointerAttributes a;
        if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) cur = a.device;
        else (void)cudaGetLastError();
        if (cur != prev) return cudaSetDevice(cur);
        return cudaSuccess;
    }
    ~DeviceScope() { if (prev >= 0 && cur != prev) cudaSetDevice(prev); }
};

static float* ws_records(void* ws) { return reinterpret_cast<float*>(ws); }
static uint2* ws_rects(void* ws, int B, int F) {
    size_t rec_bytes = ((size_t)B * F * REC_BYTES + 255) & ~(size_t)255;
    return reinterpret_cast<uint2*>(reinterpret_cast<char*>(ws) + rec_bytes);
}

// longest-first schedule: [counts B*tiles | order B*tiles] (uint32) behind the rectangles; 32 KB per batch item hold up to 4096
// tiles per image (image_size <= 1024).  Used when the grid is more than one wave of the machine (148 SMs x 4 CTAs) -- up to
// that every CTA is resident from the start and the order is irrelevant.
static const size_t kLptBytesPerItem = 32768;
static bool lpt_enabled(const RenderParams& P) {
    const long long tiles = (long long)P.tiles_x * P.tiles_y, n = tiles * P.B;
    return n > 148 * 4 && tiles * 8 <= (long long)kLptBytesPerItem && n <= (1ll << 18);
}
static unsigned* ws_tile_counts(void* ws, int B, int F) {
    size_t rct_bytes = ((size_t)B * F * sizeof(uint2) + 255) & ~(size_t)255;
    return reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ws_rects(ws, B, F)) + rct_bytes + 256);
}
static unsigned* ws_cta_order(void* ws, const RenderParams& P) { return ws_tile_counts(ws, P.B, P.F) + (size_t)P.B * P.tiles_x * P.tiles_y; }
static int lpt_begin(const RenderParams& P, void* ws, cudaStream_t st, unsigned** counts) {
    *counts = nullptr;
    if (!lpt_enabled(P)) return 0;
    *counts = ws_tile_counts(ws, P.B, P.F);
    GENDR_CUDA(cudaMemsetAsync(*counts, 0, (size_t)P.B * P.tiles_x * P.tiles_y * sizeof(unsigned), st), "clearing the tile counters");
    return 0;
}
static int lpt_finish(const RenderParams& P, void* ws, cudaStream_t st) {
    if (!lpt_enabled(P)) return 0;
    const int group = std::max(16, (P.B + kOrderGroups - 1) / kOrderGroups);      // items per group: >= 16, at most 32 groups
    tile_order_kernel<<<1, 1024, 0, st>>>(ws_tile_counts(ws, P.B, P.F), ws_cta_order(ws, P), P.B * P.tiles_x * P.tiles_y, P.tiles_x * P.tiles_y, group);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "tile_order_kernel launch");
    return 0;
}

static int zero_grads2(float* a, long long n_a, float* b, long long n_b, cudaStream_t st) {
    if (!a) n_a = 0;
    if (!b) n_b = 0;
    const long long n = n_a + n_b;
    if (n == 0) return 0;
    if (n > (1ll << 20) || n_a == 0 || n_b == 0) {      // large (or single) buffers: the copy engine's memset is the right tool
        if (n_a) GENDR_CUDA(cudaMemsetAsync(a, 0, (size_t)n_a * sizeof(float), st), "zero gradient buffer");
        if (n_b) GENDR_CUDA(cudaMemsetAsync(b, 0, (size_t)n_b * sizeof(float), st), "zero gradient buffer");
        return 0;
    }
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    zero2_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, n_a, b, n_b);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "zero2_kernel launch");
    return 0;
}

static int run_prep(const RenderParams& P, const float* faces, float* faces_info, void* ws, cudaStream_t st) {
    const long long n = (long long)P.B * P.F;
    unsigned* counts;
    if (int e = lpt_begin(P, ws, st, &counts)) return e;
    if (n == 0) return lpt_finish(P, ws, st);      // no faces: the render kernel still needs a valid CTA order
    prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, faces, ws_records(ws), ws_rects(ws, P.B, P.F), faces_info, counts);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "prep_kernel launch");
    return lpt_finish(P, ws, st);
}

static int run_render(const RenderParams& P, const KernelIO& io, bool backward, cudaStream_t st) {
    if (P.B == 0) return 0;
    LaunchCfg cfg;
    cfg.grid = dim3((unsigned)(P.B * P.tiles_x * P.tiles_y));
    cfg.device = 0;
    (void)cudaGetDevice(&cfg.device);
    cfg.bwd_mode = backward ? backward_mode(P) : 0;
    if (backward && forced_backward_mode() < 0) {
        // small problems: fewer CTAs than ~two waves of the machine (148 SMs x 4 CTAs).  The face-stationary kernel then also
        // splits the faces across gridDim.y CTAs per tile (render_bwd_fs_kernel), whatever the density regime.
        const long long ctas = (long long)P.B * P.tiles_x * P.tiles_y;
        int splits = (int)((2 * 148 * 4 + ctas - 1) / (ctas > 0 ? ctas : 1));
        splits = splits > 8 ? 8 : splits;
        while (splits > 1 && P.F / splits < 64) --splits;
        if (splits > 1) { cfg.bwd_mode = 0; cfg.grid.y = (unsigned)splits; }
    }
    cfg.smem = smem_bytes(P, backward && cfg.bwd_mode == 0);
    cfg.stream = st;
    cfg.backward = backward;
    const bool yager2 = (P.aggr_alpha_func == T_YAGER && P.tcn_p == 2.f);
    cfg.fast = (P.aggr_rgb_func == 1 && P.texture_type == 0 && P.T == 1 && !P.dist_squared &&
                (P.aggr_alpha_func == T_PROBABILISTIC || P.aggr_alpha_func == T_EINSTEIN || yager2));
    cfg.tcn_mode = cfg.fast ? (yager2 ? 4 : P.aggr_alpha_func) : (P.aggr_alpha_func >= T_HAMACHER ? 1 : 0);
    KernelIO io2 = io;      // (records = start of the workspace)
    io2.cta_order = lpt_enabled(P) ? ws_cta_order(const_cast<float*>(io.records), P) : nullptr;      // (indexed by blockIdx.x, whatever gridDim.y)
    cudaError_t e = kLaunchTable[P.dist_func](P, io2, cfg);
    g_launches++;
    if (e != cudaSuccess) return fail((int)e, backward ? "backward render_kernel launch" : "forward render_kernel launch");
    return 0;
}

// ---- scalar functions (K.cpp:195-236 -> K.cu:1230-1270): evaluated ON THE HOST by the host instantiation of the very templates
// the render kernels use (they are __host__ __device__, exactly like the reference's sigmoid_*_cuda / t_conorm_*_cuda) -- no
// device, no allocation, no launch.  scalar_kernel runs the same switch in one device thread; it exists for the self-test
// gendr_selftest_scalar_device (host and device instantiations must agree).
template <int D> struct DistEval {
    static __host__ __device__ float run(int id, bool pdf, float s, float x, const RenderParams& P, const Consts& K) {
#ifdef __CUDA_ARCH__
        if (id == D) return pdf ? dist_pdf<D>(s, x, P, K) : (P.aggr_alpha_func == T_MAX ? dist_cdf<D, true, false>(s, x, P, K) : dist_cdf<D, false, false>(s, x, P, K));
#else   // host: always the reference's own mixed fp32/fp64 expressions (the fp32 re-formulations rely on a fused multiply-add)
        if (id == D) return pdf ? dist_pdf<D>(s, x, P, K) : dist_cdf<D, true, false>(s, x, P, K);
#endif
        return DistEval<D + 1>::run(id, pdf, s, x, P, K);
    }
};
template <> struct DistEval<D_COUNT> {
    static __host__ __device__ float run(int, bool, float, float, const RenderParams&, const Consts&) { return gd_nan(); }
};
static __host__ __device__ float eval_scalar(const RenderParams& P, int what, int id, float a, float b) {
    const Consts K = make_consts(P);
    if (what == 0) return DistEval<0>::run(id, false, a, b, P, K);
    if (what == 1) return DistEval<0>::run(id, true, a, b, P, K);
    if (what == 2) return (id >= T_HAMACHER) ? tconorm_fold<true>(id, a, b, P) : tconorm_fold<false>(id, a, b, P);
    return (id >= T_HAMACHER) ? tconorm_dS<true>(id, a, b, P) : tconorm_dS<false>(id, a, b, P);
}
__global__ void scalar_kernel(const __grid_constant__ RenderParams P, int what, int id, float a, float b, float* out) {
    *out = eval_scalar(P, what, id, a, b);
}

static bool scalar_params(RenderParams& P, int what, int id, float scale, float shape, float shift, float p) {
    gendr_render_params u;
    memset(&u, 0, sizeof u);
    u.image_size = 1; u.dist_func = (what < 2 && id >= 0 && id < D_COUNT) ? id : 0; u.dist_scale = scale; u.dist_shape = shape;
    u.dist_shift = shift; u.dist_eps = 1.f; u.aggr_alpha_func = (what >= 2 && id >= 0 && id < T_COUNT) ? id : 0;
    u.aggr_alpha_t_conorm_p = p; u.aggr_rgb_func = 1; u.aggr_rgb_gamma = 1.f;
    if (make_params(P, 1, 1, 1, &u) != 0) return false;
    return !((what < 2 && (id < 0 || id >= D_COUNT)) || (what >= 2 && (id < 1 || id >= T_COUNT)));
}

static float run_scalar(int what, int id, float a, float b, float scale, float shape, float shift, float p) {
    RenderParams P;
    if (!scalar_params(P, what, id, scale, shape, shift, p)) return std::numeric_limits<float>::quiet_NaN();
    return eval_scalar(P, what, id, a, b);
}

static float run_scalar_device(int what, int id, float a, float b, float scale, float shape, float shift, float p) {
    RenderParams P;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if (!scalar_params(P, what, id, scale, shape, shift, p)) return nan;
    float* d = nullptr;
    if (cudaMalloc(&d, sizeof(float)) != cudaSuccess) { fail((int)cudaGetLastError(), "scalar cudaMalloc"); return nan; }
    scalar_kernel<<<1, 1>>>(P, what, id, a, b, d);
    g_launches++;
    float h = nan;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { fail((int)e, "scalar kernel"); return nan; }
    return h;
}

// ---- geometry probe -----------------------------------------------------------------------------------------
__global__ void probe_kernel(const __grid_constant__ RenderParams P, const float* faces, const float* xy, float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[9], rec[REC_WORDS];
    for (int k = 0; k < 9; ++k) v[k] = faces[(size_t)i * 9 + k];
    prep_face_record(v, rec, nullptr, P);
    PairGeom g;
    const float xp = xy[2 * i], yp = xy[2 * i + 1];
    pair_barycentric(g, rec, xp, yp);
    pair_project<false>(g, rec, xp, yp, __float_as_uint(rec[R_PACK]), __float_as_uint(rec[R_PACK + 1]));
    float* o = out + (size_t)i * 10;
    o[0] = g.w0; o[1] = g.w1; o[2] = g.w2; o[3] = g.t0; o[4] = g.t1; o[5] = g.t2; o[6] = g.dx; o[7] = g.dy; o[8] = g.sign;
    o[9] = sop2(g.dx, g.dx, g.dy, g.dy);
}

// ---- self-test: div_exact (shared-reciprocal division) must be bit-identical to __fdiv_rn ---------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void division_selftest_kernel(unsigned long long n, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t h1 = mix32((uint32_t)i * 2u + 1u), h2 = mix32((uint32_t)(i >> 7) * 2654435761u + 12345u);
        // divisor: random mantissa, exponent in [2^-20, 2^20], random sign; dividend: exponent in [2^-40, 2^20] or zero
        const float b = __uint_as_float((h1 & 0x807fffffu) | ((107u + (h1 >> 23) % 41u) << 23));
        float a = __uint_as_float((h2 & 0x807fffffu) | ((87u + (h2 >> 23) % 61u) << 23));
        if ((h2 & 0xff) == 0) a = 0.f;
        const Rcp r = make_rcp(b);
        const float q1 = div_exact(a, r), q2 = __fdiv_rn(a, b);
        if (__float_as_uint(q1) != __float_as_uint(q2)) ++bad;
        Rcp rr = r;                       // the flavour used for the launch constants: y = RN(1 / b), formed on the host
        rr.y = __frcp_rn(b);
        if (__float_as_uint(div_exact(a, rr)) != __float_as_uint(q2)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ---- cached device scratch for the host-buffer entry point -----------------------------------------------------
struct HostPathScratch {      // one per device (process-wide): a process may drive several GPUs, each call uses its current device's
    std::mutex mu;
    bool ready = false;
    size_t cap = 0;
    char* base = nullptr;
    cudaStream_t stream = nullptr;                 // compute, slice 0
    cudaStream_t extra[3] = {};                    // compute, slices 1..3: a slice's kernels fill the drain of the previous slice's
    cudaStream_t h2d = nullptr, d2h = nullptr;     // copy engines, overlapped with the kernels
    static const int MAX_CHUNKS = 8;
    cudaEvent_t in_ready[2 * MAX_CHUNKS] = {}, done[2 * MAX_CHUNKS] = {};      // [chunk] forward inputs / outputs, [MAX_CHUNKS + chunk] backward
};
static const int kMaxDevices = 64;
static HostPathScratch g_scratch_of[kMaxDevices];

}  // namespace gendr

using namespace gendr;

extern "C" {

const char* gendr_last_error(void) { return g_err; }
const char* gendr_version(void) { return "gendr_b200 0.1 (sm_100a)"; }
long long gendr_launch_count(void) { return g_launches.load(); }

size_t gendr_workspace_bytes(int batch, int faces) {
    if (batch < 0 || faces < 0) return 0;
    size_t rec = ((size_t)batch * faces * REC_BYTES + 255) & ~(size_t)255;
    size_t rct = ((size_t)batch * faces * sizeof(uint2) + 255) & ~(size_t)255;
    return rec + rct + 256 + (size_t)batch * kLptBytesPerItem;      // records | rectangles | slack | tile counters + CTA order
}

static int gendr_forward_render_chunk(const float* faces, const float* textures, long long tex_elems, float* aggrs_info, float* soft_colors,
                                      int batch, int num_faces, int texture_size, const gendr_render_params* params, void* workspace,
                                      size_t workspace_bytes, cudaStream_t st) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_render_forward_backward_host");
    (void)workspace_bytes;
    if (int e = run_prep(P, faces, nullptr, workspace, st)) return e;
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, batch, num_faces);
    io.textures = textures; io.tex_elems = tex_elems;
    io.soft_colors = soft_colors; io.aggrs = aggrs_info; io.bg_from_buffer = 0;
    return run_render(P, io, false, st);
}

static int gendr_backward_render_chunk(const float* faces, const float* textures, long long tex_elems, const float* soft_colors,
                                       const float* aggrs_info, float* grad_faces, float* grad_textures, const float* grad_soft_colors,
                                       int batch, int num_faces, int texture_size, const gendr_render_params* params, void* workspace,
                                       size_t workspace_bytes, cudaStream_t st) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_render_forward_backward_host");
    (void)workspace_bytes; (void)faces;
    if (int e = zero_grads2(grad_faces, (long long)batch * num_faces * 9, grad_textures, (long long)batch * num_faces * texture_size * 3, st)) return e;
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, batch, num_faces);
    io.textures = textures; io.tex_elems = tex_elems;
    io.soft_colors = const_cast<float*>(soft_colors); io.aggrs = const_cast<float*>(aggrs_info);
    io.grad_colors = grad_soft_colors; io.grad_faces = grad_faces; io.grad_textures = grad_textures;
    io.grad_batch_stride_f = (long long)num_faces * 9;
    return run_render(P, io, true, st);
}

int gendr_forward_render(const float* faces, const float* textures, float* faces_info, float* aggrs_info,
                         float* soft_colors, int batch, int num_faces, int texture_size,
                         const gendr_render_params* params, int background_prefilled,
                         void* workspace, size_t workspace_bytes, void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_forward_render");
    if (batch == 0) return 0;
    if (((!faces || !textures) && num_faces > 0) || !aggrs_info || !soft_colors || !workspace) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_forward_render");
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    DeviceScope dev;
    GENDR_CUDA(dev.enter(faces), "selecting the device that owns `faces`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (int e = run_prep(P, faces, faces_info, workspace, st)) return e;
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, batch, num_faces);
    io.textures = textures; io.tex_elems = (long long)batch * num_faces * texture_size * 3;
    io.soft_colors = soft_colors; io.aggrs = aggrs_info; io.bg_from_buffer = background_prefilled ? 1 : 0;
    return run_render(P, io, false, st);
}

static int backward_render_impl(const char* who, bool batchsum, const float* faces, const float* textures, const float* soft_colors,
                                const float* aggrs_info, float* grad_faces, float* grad_textures, const float* grad_soft_colors, int batch,
                                int num_faces, int texture_size, const gendr_render_params* params, int workspace_valid, int zero_grads,
                                void* workspace, size_t workspace_bytes, void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, who);
    if (batch == 0 || num_faces == 0) return 0;      // no faces: every gradient buffer is empty
    if (!faces || !textures || !soft_colors || !aggrs_info || !grad_faces || !grad_soft_colors || !workspace) return fail(GENDR_ERR_INVALID_ARGUMENT, who);
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    DeviceScope dev;
    GENDR_CUDA(dev.enter(faces), "selecting the device that owns `faces`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!workspace_valid) if (int e = run_prep(P, faces, nullptr, workspace, st)) return e;
    if (zero_grads) if (int e = zero_grads2(grad_faces, (long long)(batchsum ? 1 : batch) * num_faces * 9, grad_textures,
                                            (long long)batch * num_faces * texture_size * 3, st)) return e;
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, batch, num_faces);
    io.textures = textures; io.tex_elems = (long long)batch * num_faces * texture_size * 3;
    io.soft_colors = const_cast<float*>(soft_colors); io.aggrs = const_cast<float*>(aggrs_info);
    io.grad_colors = grad_soft_colors; io.grad_faces = grad_faces; io.grad_textures = grad_textures;
    io.grad_batch_stride_f = batchsum ? 0 : (long long)num_faces * 9;
    return run_render(P, io, true, st);
}

int gendr_backward_render(const float* faces, const float* textures, const float* soft_colors,
                          const float* aggrs_info, float* grad_faces, float* grad_textures,
                          const float* grad_soft_colors, int batch, int num_faces, int texture_size,
                          const gendr_render_params* params, int workspace_valid, int zero_grads,
                          void* workspace, size_t workspace_bytes, void* stream) {
    return backward_render_impl("invalid argument to gendr_backward_render", false, faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures,
                                grad_soft_colors, batch, num_faces, texture_size, params, workspace_valid, zero_grads, workspace, workspace_bytes, stream);
}

int gendr_backward_render_batchsum(const float* faces, const float* textures, const float* soft_colors,
                                   const float* aggrs_info, float* grad_faces_sum, float* grad_textures,
                                   const float* grad_soft_colors, int batch, int num_faces, int texture_size,
                                   const gendr_render_params* params, int workspace_valid, int zero_grads,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    return backward_render_impl("invalid argument to gendr_backward_render_batchsum", true, faces, textures, soft_colors, aggrs_info, grad_faces_sum,
                                grad_textures, grad_soft_colors, batch, num_faces, texture_size, params, workspace_valid, zero_grads, workspace,
                                workspace_bytes, stream);
}

int gendr_render_forward_backward_host(const float* h_faces, const float* h_textures, const float* h_grad_soft_colors,
                                       float* h_soft_colors, float* h_grad_faces, float* h_grad_textures,
                                       int batch, int num_faces, int texture_size, const gendr_render_params* params) {
    if (!params || !h_faces || !h_textures || !h_soft_colors) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_render_forward_backward_host");
    if (batch <= 0) return 0;
    const size_t S = (size_t)params->image_size;
    // per batch item
    const size_t e_faces = (size_t)num_faces * 9, e_tex = (size_t)num_faces * texture_size * 3, e_col = 4 * S * S, e_agg = 2 * S * S;
    const size_t n_faces = (size_t)batch * e_faces * 4, n_tex = (size_t)batch * e_tex * 4;
    const size_t n_col = (size_t)batch * e_col * 4, n_agg = (size_t)batch * e_agg * 4;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const bool bwd = h_grad_soft_colors && h_grad_faces;
    // The batch is cut into up to four slices that move through upload -> forward -> backward -> download as a pipeline: only the
    // first slice's geometry upload and the last slice's gradient download are exposed (one slice: ~0.5 ms each for C3, of a
    // 9 ms step).  Slices keep >= 8 items so that every kernel still fills the machine several times over.
    int n_chunks = std::max(1, std::min(4, batch / 8));
    if (const char* ev = getenv("GENDR_B200_HOST_CHUNKS")) n_chunks = std::max(1, std::min(std::min(4, batch), atoi(ev)));      // tuning experiments
    const int per = (batch + n_chunks - 1) / n_chunks;
    const size_t ws_chunk = al(gendr_workspace_bytes(per, num_faces));
    const size_t need = al(n_faces) * 2 + al(n_tex) * 2 + al(n_col) * 2 + al(n_agg) + ws_chunk * n_chunks;
    int dev = 0;
    GENDR_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) return fail(GENDR_ERR_INVALID_ARGUMENT, "device ordinal out of range for the host-buffer path");
    HostPathScratch& g_scratch = g_scratch_of[dev];
    std::lock_guard<std::mutex> lock(g_scratch.mu);
    if (!g_scratch.ready) {
        GENDR_CUDA(cudaStreamCreateWithFlags(&g_scratch.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < 3; ++i) GENDR_CUDA(cudaStreamCreateWithFlags(&g_scratch.extra[i], cudaStreamNonBlocking), "cudaStreamCreate");
        GENDR_CUDA(cudaStreamCreateWithFlags(&g_scratch.h2d, cudaStreamNonBlocking), "cudaStreamCreate");
        GENDR_CUDA(cudaStreamCreateWithFlags(&g_scratch.d2h, cudaStreamNonBlocking), "cudaStreamCreate");
        for (int i = 0; i < 2 * HostPathScratch::MAX_CHUNKS; ++i) {
            GENDR_CUDA(cudaEventCreateWithFlags(&g_scratch.in_ready[i], cudaEventDisableTiming), "cudaEventCreate");
            GENDR_CUDA(cudaEventCreateWithFlags(&g_scratch.done[i], cudaEventDisableTiming), "cudaEventCreate");
        }
        g_scratch.ready = true;
    }
    if (g_scratch.cap < need) {
        if (g_scratch.base) { cudaFree(g_scratch.base); g_scratch.base = nullptr; g_scratch.cap = 0; }
        GENDR_CUDA(cudaMalloc(&g_scratch.base, need), "cudaMalloc of host-path scratch");
        g_scratch.cap = need;
    }
    char* p = g_scratch.base;
    float* d_faces = (float*)p; p += al(n_faces);
    float* d_gfaces = (float*)p; p += al(n_faces);
    float* d_tex = (float*)p; p += al(n_tex);
    float* d_gtex = (float*)p; p += al(n_tex);
    float* d_col = (float*)p; p += al(n_col);
    float* d_gcol = (float*)p; p += al(n_col);
    float* d_agg = (float*)p; p += al(n_agg);
    char* d_ws = p;
    const int NC = HostPathScratch::MAX_CHUNKS;
    // one compute stream per slice: the slices' kernels are independent, so slice c + 1 fills the SMs that slice c's last CTAs leave idle
    auto cs = [&](int c) { return c == 0 ? g_scratch.stream : g_scratch.extra[c - 1]; };
    // Three streams: uploads / compute / downloads.  Uploads in the order the kernels need them: all textures (the forward kernel
    // reads one texel past its slice, quirk Q3), the slices' geometry, then the slices' cotangents (the largest input; overlaps
    // the forward kernels).  Image downloads overlap the remaining kernels.
    GENDR_CUDA(cudaMemcpyAsync(d_tex, h_textures, n_tex, cudaMemcpyHostToDevice, g_scratch.h2d), "H2D textures");
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = c * per, nb = std::min(per, batch - b0);
        if (nb <= 0) break;
        GENDR_CUDA(cudaMemcpyAsync(d_faces + (size_t)b0 * e_faces, h_faces + (size_t)b0 * e_faces, (size_t)nb * e_faces * 4, cudaMemcpyHostToDevice, g_scratch.h2d), "H2D faces");
        GENDR_CUDA(cudaEventRecord(g_scratch.in_ready[c], g_scratch.h2d), "event record");
    }
    if (bwd)
        for (int c = 0; c < n_chunks; ++c) {
            const int b0 = c * per, nb = std::min(per, batch - b0);
            if (nb <= 0) break;
            GENDR_CUDA(cudaMemcpyAsync(d_gcol + (size_t)b0 * e_col, h_grad_soft_colors + (size_t)b0 * e_col, (size_t)nb * e_col * 4, cudaMemcpyHostToDevice, g_scratch.h2d), "H2D grad_soft_colors");
            GENDR_CUDA(cudaEventRecord(g_scratch.in_ready[NC + c], g_scratch.h2d), "event record");
        }
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = c * per, nb = std::min(per, batch - b0);
        if (nb <= 0) break;
        GENDR_CUDA(cudaStreamWaitEvent(cs(c), g_scratch.in_ready[c], 0), "stream wait");
        if (int e = gendr_forward_render_chunk(d_faces + (size_t)b0 * e_faces, d_tex + (size_t)b0 * e_tex, (long long)(batch - b0) * (long long)e_tex,
                                               d_agg + (size_t)b0 * e_agg, d_col + (size_t)b0 * e_col, nb, num_faces, texture_size, params,
                                               d_ws + (size_t)c * ws_chunk, ws_chunk, cs(c))) return e;
        GENDR_CUDA(cudaEventRecord(g_scratch.done[c], cs(c)), "event record");
        GENDR_CUDA(cudaStreamWaitEvent(g_scratch.d2h, g_scratch.done[c], 0), "stream wait");
        GENDR_CUDA(cudaMemcpyAsync(h_soft_colors + (size_t)b0 * e_col, d_col + (size_t)b0 * e_col, (size_t)nb * e_col * 4, cudaMemcpyDeviceToHost, g_scratch.d2h), "D2H soft_colors");
    }
    if (bwd)
        for (int c = 0; c < n_chunks; ++c) {
            const int b0 = c * per, nb = std::min(per, batch - b0);
            if (nb <= 0) break;
            GENDR_CUDA(cudaStreamWaitEvent(cs(c), g_scratch.in_ready[NC + c], 0), "stream wait");
            if (int e = gendr_backward_render_chunk(d_faces + (size_t)b0 * e_faces, d_tex + (size_t)b0 * e_tex, (long long)(batch - b0) * (long long)e_tex,
                                                    d_col + (size_t)b0 * e_col, d_agg + (size_t)b0 * e_agg, d_gfaces + (size_t)b0 * e_faces,
                                                    h_grad_textures ? d_gtex + (size_t)b0 * e_tex : nullptr, d_gcol + (size_t)b0 * e_col, nb, num_faces,
                                                    texture_size, params, d_ws + (size_t)c * ws_chunk, ws_chunk, cs(c))) return e;
            GENDR_CUDA(cudaEventRecord(g_scratch.done[NC + c], cs(c)), "event record");
            GENDR_CUDA(cudaStreamWaitEvent(g_scratch.d2h, g_scratch.done[NC + c], 0), "stream wait");
            GENDR_CUDA(cudaMemcpyAsync(h_grad_faces + (size_t)b0 * e_faces, d_gfaces + (size_t)b0 * e_faces, (size_t)nb * e_faces * 4, cudaMemcpyDeviceToHost, g_scratch.d2h), "D2H grad_faces");
            if (h_grad_textures)
                GENDR_CUDA(cudaMemcpyAsync(h_grad_textures + (size_t)b0 * e_tex, d_gtex + (size_t)b0 * e_tex, (size_t)nb * e_tex * 4, cudaMemcpyDeviceToHost, g_scratch.d2h), "D2H grad_textures");
        }
    GENDR_CUDA(cudaStreamSynchronize(g_scratch.d2h), "host-path stream synchronize");
    GENDR_CUDA(cudaStreamSynchronize(g_scratch.stream), "host-path stream synchronize");
    for (int i = 0; i < 3; ++i) GENDR_CUDA(cudaStreamSynchronize(g_scratch.extra[i]), "host-path stream synchronize");
    return 0;
}

void gendr_release_host_scratch(void) {
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); return; }
    for (int d = 0; d < kMaxDevices; ++d) {
        HostPathScratch& sc = g_scratch_of[d];
        std::lock_guard<std::mutex> lock(sc.mu);
        if (!sc.ready && !sc.base) continue;
        if (cudaSetDevice(d) != cudaSuccess) { (void)cudaGetLastError(); continue; }
        if (sc.ready) {
            cudaStreamSynchronize(sc.stream); cudaStreamSynchronize(sc.h2d); cudaStreamSynchronize(sc.d2h);
            cudaStreamDestroy(sc.stream); cudaStreamDestroy(sc.h2d); cudaStreamDestroy(sc.d2h);
            for (int i = 0; i < 3; ++i) { cudaStreamSynchronize(sc.extra[i]); cudaStreamDestroy(sc.extra[i]); sc.extra[i] = nullptr; }
            for (int i = 0; i < 2 * HostPathScratch::MAX_CHUNKS; ++i) { cudaEventDestroy(sc.in_ready[i]); cudaEventDestroy(sc.done[i]); }
            sc.stream = sc.h2d = sc.d2h = nullptr;
            sc.ready = false;
        }
        if (sc.base) { cudaFree(sc.base); sc.base = nullptr; sc.cap = 0; }
    }
    cudaSetDevice(prev);
}

static int check_aa(const RenderParams& P, const void* pooled_or_flag) {
    if (pooled_or_flag && (P.S & 1)) return fail(GENDR_ERR_INVALID_ARGUMENT, "anti-aliasing needs an even (supersampled) image_size");
    return 0;
}

static int forward_indexed_impl(const RenderParams& P, const float* vertices, const int* face_index, int index_shared, const float* textures,
                                float* aggrs_info, float* soft_colors, float* pooled_colors, int num_vertices, void* workspace, cudaStream_t st) {
    const long long n = (long long)P.B * P.F;
    unsigned* counts;
    if (int e = lpt_begin(P, workspace, st, &counts)) return e;
    if (n == 0) { if (int e = lpt_finish(P, workspace, st)) return e; }
    if (n > 0) {
        prep_indexed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, vertices, face_index, index_shared ? 0 : (long long)P.F * 3, num_vertices,
                                                                       ws_records(workspace), ws_rects(workspace, P.B, P.F), counts);
        g_launches++;
        GENDR_CUDA(cudaGetLastError(), "prep_indexed_kernel launch");
        if (int e = lpt_finish(P, workspace, st)) return e;
    }
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, P.B, P.F);
    io.textures = textures; io.tex_elems = (long long)P.B * P.F * P.T * 3;
    io.soft_colors = soft_colors; io.aggrs = aggrs_info; io.pooled = pooled_colors;
    return run_render(P, io, false, st);
}

static int backward_indexed_impl(const RenderParams& P, const int* face_index, int index_shared, const float* textures, const float* soft_colors,
                                 const float* aggrs_info, float* grad_vertices, float* grad_textures, const float* grad_soft_colors,
                                 int grad_is_pooled, int num_vertices, int zero_grads, void* workspace, cudaStream_t st, bool batchsum = false) {
    if (zero_grads) {
        if (int e = zero_grads2(grad_vertices, (long long)(batchsum ? 1 : P.B) * num_vertices * 3, grad_textures, (long long)P.B * P.F * P.T * 3, st)) return e;
    }
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, P.B, P.F);    // left there by the indexed forward
    io.textures = textures; io.tex_elems = (long long)P.B * P.F * P.T * 3;
    io.soft_colors = const_cast<float*>(soft_colors); io.aggrs = const_cast<float*>(aggrs_info);
    io.grad_colors = grad_soft_colors; io.grad_textures = grad_textures; io.grad_pooled = grad_is_pooled ? 1 : 0;
    io.grad_vertices = grad_vertices; io.face_index = face_index; io.index_batch_stride = index_shared ? 0 : (long long)P.F * 3;
    io.num_vertices = num_vertices; io.grad_batch_stride_v = batchsum ? 0 : (long long)num_vertices * 3;
    return run_render(P, io, true, st);
}

int gendr_forward_render_indexed(const float* vertices, const int* face_index, int index_shared, const float* textures, float* aggrs_info,
                                 float* soft_colors, float* pooled_colors, int batch, int num_vertices, int num_faces, int texture_size,
                                 const gendr_render_params* params, void* workspace, size_t workspace_bytes, void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_forward_render_indexed");
    if (batch == 0) return 0;
    if (((!vertices || !face_index || !textures) && num_faces > 0) || !aggrs_info || !soft_colors || !workspace || num_vertices < 1)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_forward_render_indexed");
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    if (int e = check_aa(P, pooled_colors)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    return forward_indexed_impl(P, vertices, face_index, index_shared, textures, aggrs_info, soft_colors, pooled_colors, num_vertices, workspace,
                                reinterpret_cast<cudaStream_t>(stream));
}

int gendr_backward_render_indexed(const int* face_index, int index_shared, const float* textures, const float* soft_colors,
                                  const float* aggrs_info, float* grad_vertices, float* grad_textures, const float* grad_soft_colors,
                                  int grad_is_pooled, int batch, int num_vertices, int num_faces, int texture_size,
                                  const gendr_render_params* params, int zero_grads, void* workspace, size_t workspace_bytes, void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_backward_render_indexed");
    if (batch == 0) return 0;
    if (num_faces == 0) {      // no faces: the vertex gradient is zero
        if (!grad_vertices || num_vertices < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_backward_render_indexed");
        if (zero_grads) GENDR_CUDA(cudaMemsetAsync(grad_vertices, 0, (size_t)batch * num_vertices * 3 * sizeof(float), reinterpret_cast<cudaStream_t>(stream)), "zero gradient buffer");
        return 0;
    }
    if (!face_index || !textures || !soft_colors || !aggrs_info || !grad_vertices || !grad_soft_colors || !workspace || num_vertices < 1)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_backward_render_indexed");
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    if (int e = check_aa(P, grad_is_pooled ? grad_soft_colors : nullptr)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(grad_vertices), "selecting the device that owns `grad_vertices`");
    return backward_indexed_impl(P, face_index, index_shared, textures, soft_colors, aggrs_info, grad_vertices, grad_textures, grad_soft_colors,
                                 grad_is_pooled, num_vertices, zero_grads, workspace, reinterpret_cast<cudaStream_t>(stream));
}

int gendr_backward_render_indexed_batchsum(const int* face_index, int index_shared, const float* textures, const float* soft_colors,
                                           const float* aggrs_info, float* grad_vertices_sum, float* grad_textures, const float* grad_soft_colors,
                                           int grad_is_pooled, int batch, int num_vertices, int num_faces, int texture_size,
                                           const gendr_render_params* params, int zero_grads, void* workspace, size_t workspace_bytes, void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_backward_render_indexed_batchsum");
    if (batch == 0) return 0;
    if (num_faces == 0) {      // no faces: the vertex gradient is zero
        if (!grad_vertices_sum || num_vertices < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_backward_render_indexed_batchsum");
        if (zero_grads) GENDR_CUDA(cudaMemsetAsync(grad_vertices_sum, 0, (size_t)1 * num_vertices * 3 * sizeof(float), reinterpret_cast<cudaStream_t>(stream)), "zero gradient buffer");
        return 0;
    }
    if (!face_index || !textures || !soft_colors || !aggrs_info || !grad_vertices_sum || !grad_soft_colors || !workspace || num_vertices < 1)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_backward_render_indexed_batchsum");
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    if (int e = check_aa(P, grad_is_pooled ? grad_soft_colors : nullptr)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(grad_vertices_sum), "selecting the device that owns `grad_vertices_sum`");
    return backward_indexed_impl(P, face_index, index_shared, textures, soft_colors, aggrs_info, grad_vertices_sum, grad_textures, grad_soft_colors,
                                 grad_is_pooled, num_vertices, zero_grads, workspace, reinterpret_cast<cudaStream_t>(stream), true);
}

// ---- fused 2x anti-aliasing on the face-vertex path ----------------------------------------
int gendr_forward_render_aa(const float* faces, const float* textures, float* aggrs_info, float* soft_colors, float* pooled_colors, int batch,
                            int num_faces, int texture_size, const gendr_render_params* params, void* workspace, size_t workspace_bytes,
                            void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_forward_render_aa");
    if (batch == 0) return 0;
    if (((!faces || !textures) && num_faces > 0) || !aggrs_info || !soft_colors || !pooled_colors || !workspace)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_forward_render_aa");
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    if (int e = check_aa(P, pooled_colors)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(faces), "selecting the device that owns `faces`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (int e = run_prep(P, faces, nullptr, workspace, st)) return e;
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, batch, num_faces);
    io.textures = textures; io.tex_elems = (long long)batch * num_faces * texture_size * 3;
    io.soft_colors = soft_colors; io.aggrs = aggrs_info; io.pooled = pooled_colors;
    return run_render(P, io, false, st);
}

int gendr_backward_render_aa(const float* faces, const float* textures, const float* soft_colors, const float* aggrs_info, float* grad_faces,
                             float* grad_textures, const float* grad_pooled_colors, int batch, int num_faces, int texture_size,
                             const gendr_render_params* params, int workspace_valid, int zero_grads, void* workspace, size_t workspace_bytes,
                             void* stream) {
    RenderParams P;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_backward_render_aa");
    if (batch == 0 || num_faces == 0) return 0;
    if (!faces || !textures || !soft_colors || !aggrs_info || !grad_faces || !grad_pooled_colors || !workspace)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_backward_render_aa");
    if (workspace_bytes < gendr_workspace_bytes(batch, num_faces)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_workspace_bytes)");
    if (int e = check_aa(P, grad_pooled_colors)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(faces), "selecting the device that owns `faces`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!workspace_valid) if (int e = run_prep(P, faces, nullptr, workspace, st)) return e;
    if (zero_grads) {
        if (int e = zero_grads2(grad_faces, (long long)batch * num_faces * 9, grad_textures, (long long)batch * num_faces * texture_size * 3, st)) return e;
    }
    KernelIO io;
    memset(&io, 0, sizeof io);
    io.records = ws_records(workspace); io.rects = ws_rects(workspace, batch, num_faces);
    io.textures = textures; io.tex_elems = (long long)batch * num_faces * texture_size * 3;
    io.soft_colors = const_cast<float*>(soft_colors); io.aggrs = const_cast<float*>(aggrs_info);
    io.grad_colors = grad_pooled_colors; io.grad_pooled = 1; io.grad_faces = grad_faces; io.grad_textures = grad_textures;
    io.grad_batch_stride_f = (long long)num_faces * 9;
    return run_render(P, io, true, st);
}

// ---- camera transform + lighting ----------------------------------------------------------
static int make_camera(CameraParams& C, const gendr_camera_params* u, int eyes_batched) {
    if (!u || (u->mode != 0 && u->mode != 1)) return GENDR_ERR_INVALID_ARGUMENT;
    memset(&C, 0, sizeof C);
    C.mode = u->mode; C.perspective = u->perspective ? 1 : 0; C.eye_stride = eyes_batched ? 3 : 0;
    for (int k = 0; k < 3; ++k) { C.at_or_dir[k] = u->at_or_direction[k]; C.up[k] = u->up[k]; }
    // transform.py:20-22: torch.tan(torch.tensor(angle / 180 * math.pi, dtype=torch.float32)) -- the tan() itself runs in the kernels
    // (device libm, like torch's), the host only forms the fp32 angle
    C.angle_rad = (float)((double)u->viewing_angle / 180. * 3.14159265358979323846);
    C.scale = u->viewing_scale;
    return 0;
}
static void make_light(LightParams& L, const gendr_light_params* u) {
    for (int k = 0; k < 3; ++k) {
        L.ambient[k] = u->intensity_ambient * u->color_ambient[k];
        L.color_dir[k] = u->color_directional[k];
        L.direction[k] = u->direction[k];
    }
    L.intensity_dir = u->intensity_directional;
}
static unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

static int camera_forward_impl(const CameraParams& C, const float* vertices, long long vstride, const float* eyes, float* screen, int B, int V,
                               cudaStream_t st) {
    const long long n = (long long)B * V;
    if (n == 0) return 0;
    camera_forward_kernel<<<blocks_for(n), 256, 0, st>>>(C, vertices, vstride, eyes, screen, B, V);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "camera_forward_kernel launch");
    return 0;
}
// grad_eyes (may be null): [B,3], or [3] when one eye is shared by the batch; eye_acc: [B,12] scratch (required with grad_eyes)
static int camera_backward_impl(const CameraParams& C, const float* vertices, long long vstride, const float* eyes, const float* grad_screen,
                                float* grad_vertices, float* grad_eyes, float* eye_acc, int B, int V, cudaStream_t st) {
    const long long n = (long long)B * V;
    if (n == 0) return 0;
    if (grad_eyes) {
        GENDR_CUDA(cudaMemsetAsync(eye_acc, 0, (size_t)B * 12 * sizeof(float), st), "zero eye-gradient accumulators");
        if (C.eye_stride == 0) GENDR_CUDA(cudaMemsetAsync(grad_eyes, 0, 3 * sizeof(float), st), "zero grad_eyes");
    }
    camera_backward_kernel<<<blocks_for(n), 256, 0, st>>>(C, vertices, vstride, eyes, grad_screen, grad_vertices, grad_eyes ? eye_acc : nullptr, B, V);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "camera_backward_kernel launch");
    if (grad_eyes) {
        camera_eye_grad_kernel<<<(B + 127) / 128, 128, 0, st>>>(C, eyes, eye_acc, grad_eyes, B);
        g_launches++;
        GENDR_CUDA(cudaGetLastError(), "camera_eye_grad_kernel launch");
    }
    return 0;
}
static int lighting_forward_impl(const LightParams& L, const float* vertices, long long vstride, const int* face_index, int index_shared,
                                 const float* textures, float* lit, int B, int V, int F, int T, cudaStream_t st) {
    const long long n = (long long)B * F;
    if (n == 0) return 0;
    lighting_forward_kernel<<<blocks_for(n), 256, 0, st>>>(L, vertices, vstride, face_index, index_shared ? 0 : (long long)F * 3, textures, lit, B, V, F, T);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "lighting_forward_kernel launch");
    return 0;
}
static int lighting_backward_impl(const LightParams& L, const float* vertices, long long vstride, const int* face_index, int index_shared,
                                  const float* textures, const float* grad_lit, float* grad_textures, float* grad_vertices, int B, int V, int F, int T,
                                  cudaStream_t st) {
    const long long n = (long long)B * F;
    if (n == 0) return 0;
    lighting_backward_kernel<<<blocks_for(n), 256, 0, st>>>(L, vertices, vstride, face_index, index_shared ? 0 : (long long)F * 3, textures, grad_lit,
                                                            grad_textures, grad_vertices, B, V, F, T);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "lighting_backward_kernel launch");
    return 0;
}

int gendr_camera_forward(const float* vertices, const float* eyes, int eyes_batched, float* screen_vertices, int batch, int num_vertices,
                         const gendr_camera_params* camera, void* stream) {
    CameraParams C;
    if (make_camera(C, camera, eyes_batched) || batch < 0 || num_vertices < 0) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_camera_forward");
    if (batch == 0 || num_vertices == 0) return 0;
    if (!vertices || !eyes || !screen_vertices) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_camera_forward");
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    return camera_forward_impl(C, vertices, (long long)num_vertices * 3, eyes, screen_vertices, batch, num_vertices, reinterpret_cast<cudaStream_t>(stream));
}

int gendr_camera_backward(const float* vertices, const float* eyes, int eyes_batched, const float* grad_screen_vertices, float* grad_vertices,
                          float* grad_eyes, float* eye_scratch, int batch, int num_vertices, const gendr_camera_params* camera, void* stream) {
    CameraParams C;
    if (make_camera(C, camera, eyes_batched) || batch < 0 || num_vertices < 0) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_camera_backward");
    if (batch == 0 || num_vertices == 0) return 0;
    if (!vertices || !eyes || !grad_screen_vertices || !grad_vertices || (grad_eyes && !eye_scratch))
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_camera_backward");
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    return camera_backward_impl(C, vertices, (long long)num_vertices * 3, eyes, grad_screen_vertices, grad_vertices, grad_eyes, eye_scratch, batch,
                                num_vertices, reinterpret_cast<cudaStream_t>(stream));
}

int gendr_lighting_forward(const float* vertices, const int* face_index, int index_shared, const float* textures, float* lit_textures, int batch,
                           int num_vertices, int num_faces, int texture_size, const gendr_light_params* light, void* stream) {
    if (!light || batch < 0 || num_faces < 0 || num_vertices < 1 || texture_size < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_lighting_forward");
    if (batch == 0 || num_faces == 0) return 0;
    if (!vertices || !face_index || !textures || !lit_textures) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_lighting_forward");
    LightParams L;
    make_light(L, light);
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    return lighting_forward_impl(L, vertices, (long long)num_vertices * 3, face_index, index_shared, textures, lit_textures, batch, num_vertices, num_faces, texture_size,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int gendr_lighting_backward(const float* vertices, const int* face_index, int index_shared, const float* textures, const float* grad_lit_textures,
                            float* grad_textures, float* grad_vertices, int batch, int num_vertices, int num_faces, int texture_size,
                            const gendr_light_params* light, void* stream) {
    if (!light || batch < 0 || num_faces < 0 || num_vertices < 1 || texture_size < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_lighting_backward");
    if (batch == 0 || num_faces == 0) return 0;
    if (!vertices || !face_index || !textures || !grad_lit_textures) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_lighting_backward");
    LightParams L;
    make_light(L, light);
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    return lighting_backward_impl(L, vertices, (long long)num_vertices * 3, face_index, index_shared, textures, grad_lit_textures, grad_textures, grad_vertices, batch,
                                  num_vertices, num_faces, texture_size, reinterpret_cast<cudaStream_t>(stream));
}

// Lighting of VERTEX textures (gendr/lighting.py:60-66): two launches forward (corner cross products summed per vertex, then the
// per-vertex light), two backward.  normal_sums [B,V,3] is written by the forward call and read by the backward call.
int gendr_vertex_lighting_forward(const float* vertices, const int* face_index, int index_shared, const float* textures, float* lit_textures,
                                  float* normal_sums, int batch, int num_vertices, int num_faces, const gendr_light_params* light, void* stream) {
    if (!light || batch < 0 || num_faces < 0 || num_vertices < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_vertex_lighting_forward");
    if (batch == 0) return 0;
    if (!vertices || (!face_index && num_faces > 0) || !textures || !lit_textures || !normal_sums)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_vertex_lighting_forward");
    LightParams L;
    make_light(L, light);
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long nv = (long long)batch * num_vertices, nf = (long long)batch * num_faces;
    GENDR_CUDA(cudaMemsetAsync(normal_sums, 0, (size_t)nv * 12, st), "clearing the vertex normal sums");
    if (nf > 0) {
        vertex_normal_sums_kernel<<<blocks_for(nf), 256, 0, st>>>(vertices, face_index, index_shared ? 0 : (long long)num_faces * 3, normal_sums, batch,
                                                                  num_vertices, num_faces);
        g_launches++;
        GENDR_CUDA(cudaGetLastError(), "vertex_normal_sums_kernel launch");
    }
    vertex_lighting_forward_kernel<<<blocks_for(nv), 256, 0, st>>>(L, normal_sums, textures, lit_textures, nv);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "vertex_lighting_forward_kernel launch");
    return 0;
}

int gendr_vertex_lighting_backward(const float* vertices, const int* face_index, int index_shared, const float* textures, const float* normal_sums,
                                   const float* grad_lit_textures, float* grad_textures, float* grad_vertices, float* grad_sums_scratch, int batch,
                                   int num_vertices, int num_faces, const gendr_light_params* light, void* stream) {
    if (!light || batch < 0 || num_faces < 0 || num_vertices < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_vertex_lighting_backward");
    if (batch == 0) return 0;
    if (!vertices || (!face_index && num_faces > 0) || !textures || !normal_sums || !grad_lit_textures || (grad_vertices && !grad_sums_scratch))
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_vertex_lighting_backward");
    LightParams L;
    make_light(L, light);
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long nv = (long long)batch * num_vertices, nf = (long long)batch * num_faces;
    if (!grad_textures && !grad_vertices) return 0;
    vertex_lighting_backward_kernel<<<blocks_for(nv), 256, 0, st>>>(L, normal_sums, textures, grad_lit_textures, grad_textures,
                                                                    grad_vertices ? grad_sums_scratch : nullptr, nv);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "vertex_lighting_backward_kernel launch");
    if (grad_vertices && nf > 0) {
        vertex_normal_sums_backward_kernel<<<blocks_for(nf), 256, 0, st>>>(vertices, face_index, index_shared ? 0 : (long long)num_faces * 3, grad_sums_scratch,
                                                                           grad_vertices, batch, num_vertices, num_faces);
        g_launches++;
        GENDR_CUDA(cudaGetLastError(), "vertex_normal_sums_backward_kernel launch");
    }
    return 0;
}

// scene workspace: [render workspace | screen vertices B*V*3 | grad screen vertices B*V*3 | lit textures B*F*T*3 | grad lit B*F*T*3 |
//                   eye-gradient accumulators B*12]
struct SceneWs { void* render; float* screen; float* grad_screen; float* lit; float* grad_lit; float* eye_acc; };
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static SceneWs scene_ws(void* ws, int B, int V, int F, int T) {
    SceneWs w;
    char* p = reinterpret_cast<char*>(ws);
    w.render = p; p += al256(gendr_workspace_bytes(B, F));
    w.screen = reinterpret_cast<float*>(p); p += al256((size_t)B * V * 12);
    w.grad_screen = reinterpret_cast<float*>(p); p += al256((size_t)B * V * 12);
    w.lit = reinterpret_cast<float*>(p); p += al256((size_t)B * F * T * 12);
    w.grad_lit = reinterpret_cast<float*>(p); p += al256((size_t)B * F * T * 12);
    w.eye_acc = reinterpret_cast<float*>(p);
    return w;
}
size_t gendr_scene_workspace_bytes(int batch, int num_vertices, int num_faces, int texture_size) {
    if (batch < 0 || num_vertices < 0 || num_faces < 0 || texture_size < 1) return 0;
    return al256(gendr_workspace_bytes(batch, num_faces)) + 2 * al256((size_t)batch * num_vertices * 12) +
           2 * al256((size_t)batch * num_faces * texture_size * 12) + al256((size_t)batch * 48) + 256;
}

int gendr_scene_forward(const float* vertices, int vertices_shared, const int* face_index, int index_shared, const float* textures, const float* eyes, int eyes_batched,
                        const gendr_camera_params* camera, const gendr_light_params* light, float* aggrs_info, float* soft_colors,
                        float* pooled_colors, int batch, int num_vertices, int num_faces, int texture_size, const gendr_render_params* params,
                        void* workspace, size_t workspace_bytes, void* stream) {
    RenderParams P;
    CameraParams C;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_scene_forward");
    if (make_camera(C, camera, eyes_batched) || num_vertices < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid camera passed to gendr_scene_forward");
    if (P.texture_type != 0) return fail(GENDR_ERR_INVALID_ARGUMENT, "gendr_scene_forward supports surface textures only");
    if (batch == 0) return 0;
    if (!vertices || !eyes || ((!face_index || !textures) && num_faces > 0) || !aggrs_info || !soft_colors || !workspace)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_scene_forward");
    if (workspace_bytes < gendr_scene_workspace_bytes(batch, num_vertices, num_faces, texture_size))
        return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_scene_workspace_bytes)");
    if (int e = check_aa(P, pooled_colors)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const SceneWs w = scene_ws(workspace, batch, num_vertices, num_faces, texture_size);
    const long long vstride = vertices_shared ? 0 : (long long)num_vertices * 3;
    if (int e = camera_forward_impl(C, vertices, vstride, eyes, w.screen, batch, num_vertices, st)) return e;
    const float* tex = textures;
    if (light) {
        LightParams L;
        make_light(L, light);
        if (int e = lighting_forward_impl(L, vertices, vstride, face_index, index_shared, textures, w.lit, batch, num_vertices, num_faces, texture_size, st)) return e;
        tex = w.lit;
    }
    return forward_indexed_impl(P, w.screen, face_index, index_shared, tex, aggrs_info, soft_colors, pooled_colors, num_vertices, w.render, st);
}

int gendr_scene_backward(const float* vertices, int vertices_shared, const int* face_index, int index_shared, const float* textures, const float* eyes, int eyes_batched,
                         const gendr_camera_params* camera, const gendr_light_params* light, const float* soft_colors, const float* aggrs_info,
                         const float* grad_soft_colors, int grad_is_pooled, float* grad_vertices, float* grad_textures, float* grad_eyes,
                         int batch, int num_vertices, int num_faces, int texture_size, const gendr_render_params* params, void* workspace,
                         size_t workspace_bytes, void* stream) {
    RenderParams P;
    CameraParams C;
    if (int e = make_params(P, batch, num_faces, texture_size, params)) return fail(e, "invalid argument to gendr_scene_backward");
    if (make_camera(C, camera, eyes_batched) || num_vertices < 1) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid camera passed to gendr_scene_backward");
    if (P.texture_type != 0) return fail(GENDR_ERR_INVALID_ARGUMENT, "gendr_scene_backward supports surface textures only");
    if (batch == 0) return 0;
    if (!vertices || !eyes || !face_index || !textures || !soft_colors || !aggrs_info || !grad_soft_colors || !grad_vertices || !workspace)
        return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_scene_backward");
    if (workspace_bytes < gendr_scene_workspace_bytes(batch, num_vertices, num_faces, texture_size))
        return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_scene_workspace_bytes)");
    if (int e = check_aa(P, grad_is_pooled ? grad_soft_colors : nullptr)) return e;
    DeviceScope dev;
    GENDR_CUDA(dev.enter(vertices), "selecting the device that owns `vertices`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const SceneWs w = scene_ws(workspace, batch, num_vertices, num_faces, texture_size);
    // the rasterizer's texture gradient: w.r.t. the LIT textures when lighting is on (always needed then: it also feeds the
    // normals' gradient), else straight into the caller's grad_textures
    float* g_tex = light ? w.grad_lit : grad_textures;
    if (int e = zero_grads2(w.grad_screen, (long long)batch * num_vertices * 3, g_tex, (long long)batch * num_faces * texture_size * 3, st)) return e;
    if (int e = backward_indexed_impl(P, face_index, index_shared, light ? w.lit : textures, soft_colors, aggrs_info, w.grad_screen, g_tex,
                                      grad_soft_colors, grad_is_pooled, num_vertices, 0, w.render, st)) return e;
    const long long vstride = vertices_shared ? 0 : (long long)num_vertices * 3;
    if (vertices_shared) GENDR_CUDA(cudaMemsetAsync(grad_vertices, 0, (size_t)num_vertices * 12, st), "zero the batch-summed vertex gradient");
    if (int e = camera_backward_impl(C, vertices, vstride, eyes, w.grad_screen, grad_vertices, grad_eyes, w.eye_acc, batch, num_vertices, st)) return e;
    if (light) {
        LightParams L;
        make_light(L, light);
        if (int e = lighting_backward_impl(L, vertices, vstride, face_index, index_shared, textures, w.grad_lit, grad_textures, grad_vertices, batch,
                                           num_vertices, num_faces, texture_size, st)) return e;
    }
    return 0;
}

// ---- voxelizer ------------------------------------------------------------------------------
static const size_t kVoxelSmemLimit = 160 * 1024;      // the three bit masks of one batch item in shared memory up to here
static size_t voxel_words(int vs) { return (size_t)vs * vs * ((vs + 31) / 32); }

size_t gendr_voxelize_workspace_bytes(int batch, int voxel_size) {
    if (batch < 0 || voxel_size < 1) return 0;
    const size_t words = voxel_words(voxel_size);
    const bool smem = 3 * words * 4 <= kVoxelSmemLimit;
    return al256((size_t)batch * words * 4) + (smem ? 0 : al256((size_t)batch * 3 * words * 4)) + 256;
}

int gendr_voxelize(const float* faces, int* voxels, int batch, int num_faces, int voxel_size, void* workspace, size_t workspace_bytes,
                   void* stream) {
    if (batch < 0 || num_faces < 0 || voxel_size < 1 || voxel_size > 1024) return fail(GENDR_ERR_INVALID_ARGUMENT, "invalid argument to gendr_voxelize");
    if ((long long)batch * voxel_size * voxel_size * voxel_size >= (1ll << 31) || (long long)batch * num_faces * 9 >= (1ll << 31))
        return fail(GENDR_ERR_INVALID_ARGUMENT, "gendr_voxelize: problem too large for 32-bit indexing");
    if (batch == 0) return 0;
    if ((!faces && num_faces > 0) || !voxels || !workspace) return fail(GENDR_ERR_INVALID_ARGUMENT, "null pointer passed to gendr_voxelize");
    if (workspace_bytes < gendr_voxelize_workspace_bytes(batch, voxel_size)) return fail(GENDR_ERR_WORKSPACE_TOO_SMALL, "workspace too small (see gendr_voxelize_workspace_bytes)");
    DeviceScope dev;
    GENDR_CUDA(dev.enter(voxels), "selecting the device that owns `voxels`");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int vs = voxel_size, W = (vs + 31) / 32;
    const size_t words = voxel_words(vs);
    const bool smem = 3 * words * 4 <= kVoxelSmemLimit;
    uint32_t* mask = reinterpret_cast<uint32_t*>(workspace);
    uint32_t* scratch = smem ? nullptr : reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(workspace) + al256((size_t)batch * words * 4));
    GENDR_CUDA(cudaMemsetAsync(mask, 0, (size_t)batch * words * 4, st), "zero occupancy mask");
    if (num_faces > 0) {
        const long long n = (long long)batch * num_faces * 3;
        voxel_surface_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(faces, mask, batch, num_faces, vs, W);
        g_launches++;
        GENDR_CUDA(cudaGetLastError(), "voxel_surface_kernel launch");
    }
    const size_t smem_bytes = smem ? 3 * words * 4 : 0;
    if (smem_bytes > 48 * 1024)
        GENDR_CUDA(cudaFuncSetAttribute(voxel_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes), "voxel_fill_kernel smem");
    voxel_fill_kernel<<<(unsigned)batch, VOX_FILL_THREADS, smem_bytes, st>>>(faces, mask, scratch, voxels, num_faces, vs, W, smem ? 1 : 0);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "voxel_fill_kernel launch");
    return 0;
}

float gendr_sigmoid_forward(int id, float sign, float x, float scale, float shape, float shift) {
    return run_scalar(0, id, sign, x, scale, shape, shift, 0.f);
}
float gendr_sigmoid_backward(int id, float sign, float x, float scale, float shape, float shift) {
    return run_scalar(1, id, sign, x, scale, shape, shift, 0.f);
}
float gendr_t_conorm_forward(int id, float a_existing, float b_new, int face_id, float p) {
    (void)face_id;
    return run_scalar(2, id, a_existing, b_new, 1.f, 0.f, 0.f, p);
}
float gendr_t_conorm_backward(int id, float a_all, float b_current, int number_of_faces, float p) {
    (void)number_of_faces;
    return run_scalar(3, id, a_all, b_current, 1.f, 0.f, 0.f, p);
}

float gendr_selftest_scalar_device(int what, int id, float a, float b, float scale, float shape, float shift, float p) {
    return run_scalar_device(what, id, a, b, scale, shape, shift, p);
}

long long gendr_selftest_division(long long n) {
    unsigned long long* d = nullptr;
    if (cudaMalloc(&d, sizeof(unsigned long long)) != cudaSuccess) return -1;
    cudaMemset(d, 0, sizeof(unsigned long long));
    division_selftest_kernel<<<148 * 8, 256>>>((unsigned long long)n, d);
    g_launches++;
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { fail((int)e, "division self-test"); return -1; }
    return (long long)h;
}

int gendr_probe_pairs(const float* faces, const float* xy, float* out, int n, void* stream) {
    if (n <= 0) return 0;
    gendr_render_params u;
    memset(&u, 0, sizeof u);
    u.image_size = 256; u.dist_func = D_UNIFORM; u.dist_scale = 1e-2f; u.dist_eps = 1e4f; u.aggr_alpha_func = T_PROBABILISTIC;
    u.aggr_rgb_gamma = 1.f;
    RenderParams P;
    if (int e = make_params(P, 1, 1, 1, &u)) return fail(e, "probe params");
    DeviceScope dev;
    GENDR_CUDA(dev.enter(faces), "selecting device");
    probe_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(P, faces, xy, out, n);
    g_launches++;
    GENDR_CUDA(cudaGetLastError(), "probe_kernel launch");
    return 0;
}

}  // extern "C"
