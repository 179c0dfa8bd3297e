// scene_kernels.cuh -- the O(B*V) / O(B*F) steps on either side of the rasterizer, as CUDA kernels:
//
//   camera transform    look_at / look  +  perspective / orthogonal
//                       (/root/reference/gendr/functional/look_at.py:11-68, look.py:11-56, gendr/transform.py:14-44,109-168)
//   lighting            ambient + one directional light folded into surface textures (per face) or vertex textures (per vertex,
//                       vertex normals = gendr/functional/vertex_normals.py)
//                       (/root/reference/gendr/lighting.py:37-71, gendr/functional/lighting.py:12-48,
//                        surface normals gendr/mesh.py:104-108)
//   and their backward passes (the reference gets those from autograd over ~30 small torch kernels).
//
// Layout: one thread per (batch item, vertex) for the camera, one thread per (batch item, face) for the lighting.  The
// camera basis is recomputed per thread from the eye (about 40 flops) instead of being staged through memory by a
// separate launch: these kernels are launch-latency bound, not bandwidth bound.
//
// Arithmetic: the FORWARD kernels reproduce, operation for operation, what torch 2.x + cuBLAS compute for the reference's Python
// on a B200 -- identified by dumping every intermediate on the box (tools/gpu_scene_probe.py) and matching candidate orderings
// offline (tools/scene_probe_analyze.py: each rule below reproduces 100 % of 10^4..10^5 samples bit for bit):
//   torch.norm / torch.sum over a 3-vector   (t0 + t2) + t1 with individually rounded terms (thread 0 of the 2-thread reduction
//                                            owns elements 0 and 2, thread 1 element 1)
//   F.normalize                              IEEE division by max(norm, eps), per component
//   torch.cross                              fma(a1, b2, -(a2*b1)) and cyclic
//   torch.matmul(v - eye, R^T) (cuBLAS)      fma(d2, r2, fma(d1, r1, d0*r0))
//   perspective                              (x / z) / width, two IEEE divisions; width = tan() evaluated by the DEVICE libm
//   lighting                                 ambient + intensity * (colour * relu(cos)), every torch op rounded on its own
// Why it matters: the rasterizer's geometric stage amplifies 1-ulp differences of screen-space vertices into 1e-4-level image
// differences on sliver faces, so end-to-end parity with the reference package needs bit-identical vertices.
// The backward kernels are plain fp32 (gradients are sums; parity there is a few ulp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gendr {

struct CameraParams {
    int   mode;            // 0 = look_at (z axis = at - eye), 1 = look (z axis = direction)
    int   perspective;     // 1 = perspective(viewing_angle), 0 = orthogonal(viewing_scale)
    int   eye_stride;      // 3 = eyes [B,3], 0 = one eye shared by the batch
    float at_or_dir[3];    // look_at: the point looked at; look: the viewing direction
    float up[3];
    float angle_rad;       // (float)(viewing_angle / 180 * pi): transform.py:20 builds this fp32 tensor, torch.tan() runs on the device
    float scale;           // orthogonal scale
};

struct LightParams {
    float ambient[3];      // intensity_ambient * color_ambient (fp32 product, lighting.py:22)
    float intensity_dir;   // intensity_directionals
    float color_dir[3];    // color_directionals
    float direction[3];
};

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      // backward passes only
// torch.sum over a 3-vector of products: (t0 + t2) + t1, every term rounded on its own
__device__ __forceinline__ float sum3_torch(float t0, float t1, float t2) { return __fadd_rn(__fadd_rn(t0, t2), t1); }
// row of cuBLAS' [.,3] x [3,3] product: fma(d2, r2, fma(d1, r1, d0*r0))
__device__ __forceinline__ float dot3_gemm(f3 d, f3 r) { return __fmaf_rn(d.z, r.z, __fmaf_rn(d.y, r.y, __fmul_rn(d.x, r.x))); }
// torch.cross: fma(a1, b2, -(a2*b1)) and cyclic
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ f3 scale3(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
// F.normalize(v, eps): v / max(||v||_2, eps), torch.norm = sqrt((x^2 + z^2) + y^2)
__device__ __forceinline__ f3 normalize3(f3 a, float eps, float* norm_out = nullptr) {
    const float n = __fsqrt_rn(sum3_torch(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y), __fmul_rn(a.z, a.z)));
    if (norm_out) *norm_out = n;
    const float d = fmaxf(n, eps);
    return mk3(__fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d));
}
__device__ __forceinline__ f3 load3(const float* p) { return mk3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

struct CameraBasis { f3 x, y, z, eye; };
__device__ __forceinline__ CameraBasis camera_basis(const CameraParams& C, const float* __restrict__ eyes, int b) {
    CameraBasis K;
    K.eye = load3(eyes + (size_t)b * C.eye_stride);
    const f3 t = mk3(C.at_or_dir[0], C.at_or_dir[1], C.at_or_dir[2]);
    const f3 up = mk3(C.up[0], C.up[1], C.up[2]);
    K.z = normalize3(C.mode == 0 ? sub3(t, K.eye) : t, 1e-5f);       // look_at.py:54 / look.py:43
    K.x = normalize3(cross3(up, K.z), 1e-5f);                        // look_at.py:55
    K.y = normalize3(cross3(K.z, K.x), 1e-5f);                       // look_at.py:56
    return K;
}

// ---- forward: world -> screen space, one thread per vertex -----------------------------------------------------
// vstride: batch stride of `vertices` in floats -- V*3, or 0 when ONE mesh [V,3] is shared by the whole batch (the views differ
// only by their eyes: vertices.repeat(B,1,1) of experiments/opt_shape.py:86 without materialising the copies)
__global__ void __launch_bounds__(256) camera_forward_kernel(const __grid_constant__ CameraParams C, const float* __restrict__ vertices,
                                                             long long vstride, const float* __restrict__ eyes, float* __restrict__ screen,
                                                             int B, int V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * V) return;
    const int b = (int)(i / V);
    const long long v = i - (long long)b * V;
    const CameraBasis K = camera_basis(C, eyes, b);
    const f3 d = sub3(load3(vertices + b * vstride + v * 3), K.eye);               // look_at.py:64 (only_rotate = False)
    const float xc = dot3_gemm(d, K.x), yc = dot3_gemm(d, K.y), zc = dot3_gemm(d, K.z);   // look_at.py:66: matmul(v, r^T)
    float xs, ys;
    if (C.perspective) {                                                    // transform.py:20-27
        const float width = tanf(C.angle_rad);
        xs = __fdiv_rn(__fdiv_rn(xc, zc), width); ys = __fdiv_rn(__fdiv_rn(yc, zc), width);
    } else { xs = __fmul_rn(xc, C.scale); ys = __fmul_rn(yc, C.scale); }    // transform.py:41-42
    screen[i * 3 + 0] = xs; screen[i * 3 + 1] = ys; screen[i * 3 + 2] = zc;
}

// ---- backward: grad_screen [B,V,3] -> grad_vertices [B,V,3] (plain store: every thread owns its vertex), or, for a mesh shared
// by the batch (vstride == 0), red.add into the batch-summed grad_vertices [V,3] (zero-filled by the caller) --------------------
// eye_acc (may be null): [B][12] zero-filled accumulators for the gradient w.r.t. the camera position (experiments/opt_camera.py
// optimises the eye): sum_v grad_v (3) | sum_v gxc*d (3) | sum_v gyc*d (3) | sum_v gzc*d (3) with d = v - eye; camera_eye_grad_kernel
// turns them into d loss / d eye.
__global__ void __launch_bounds__(256) camera_backward_kernel(const __grid_constant__ CameraParams C, const float* __restrict__ vertices,
                                                              long long vstride, const float* __restrict__ eyes,
                                                              const float* __restrict__ grad_screen, float* __restrict__ grad_vertices,
                                                              float* __restrict__ eye_acc, int B, int V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i < (long long)B * V;
    if (!in_range && !eye_acc) return;
    const int b = in_range ? (int)(i / V) : B - 1;
    const long long v = in_range ? i - (long long)b * V : 0;
    const CameraBasis K = camera_basis(C, eyes, b);
    const f3 g = in_range ? load3(grad_screen + i * 3) : mk3(0.f, 0.f, 0.f);
    float gxc, gyc, gzc;
    const f3 d = sub3(load3(vertices + b * vstride + v * 3), K.eye);
    if (C.perspective) {
        const float xc = dot3_gemm(d, K.x), yc = dot3_gemm(d, K.y), zc = dot3_gemm(d, K.z);
        const float r = 1.f / (zc * tanf(C.angle_rad));              // d(xs)/d(xc)
        gxc = g.x * r; gyc = g.y * r;
        gzc = g.z - (g.x * xc + g.y * yc) * r / zc;                  // xs = xc / zc / width  =>  d(xs)/d(zc) = -xs / zc
    } else {
        gxc = g.x * C.scale; gyc = g.y * C.scale; gzc = g.z;
    }
    // v_cam = R (v - eye)  =>  grad_v = R^T grad_cam
    const float gx = gxc * K.x.x + gyc * K.y.x + gzc * K.z.x, gy = gxc * K.x.y + gyc * K.y.y + gzc * K.z.y,
                gz = gxc * K.x.z + gyc * K.y.z + gzc * K.z.z;
    if (in_range) {
        if (vstride != 0) {
            grad_vertices[i * 3 + 0] = gx; grad_vertices[i * 3 + 1] = gy; grad_vertices[i * 3 + 2] = gz;
        } else {
            atomicAdd(grad_vertices + v * 3 + 0, gx); atomicAdd(grad_vertices + v * 3 + 1, gy); atomicAdd(grad_vertices + v * 3 + 2, gz);
        }
    }
    if (eye_acc) {
        float t[12] = {gx, gy, gz, gxc * d.x, gxc * d.y, gxc * d.z, gyc * d.x, gyc * d.y, gyc * d.z, gzc * d.x, gzc * d.y, gzc * d.z};
        if (!in_range) {
#pragma unroll
            for (int k = 0; k < 12; ++k) t[k] = 0.f;
        }
        const unsigned full = 0xffffffffu;
        const int b0 = __shfl_sync(full, b, 0);
        if (__all_sync(full, b == b0)) {          // the usual case: the whole warp belongs to one batch item
#pragma unroll
            for (int k = 0; k < 12; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_xor_sync(full, t[k], o);
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 12; ++k) atomicAdd(eye_acc + (size_t)b * 12 + k, t[k]);
            }
        } else if (in_range) {
#pragma unroll
            for (int k = 0; k < 12; ++k) atomicAdd(eye_acc + (size_t)b * 12 + k, t[k]);
        }
    }
}

// F.normalize backward: u = v / max(|v|, eps); returns d loss / d v from d loss / d u
__device__ __forceinline__ f3 normalize3_bwd(f3 v, float eps, f3 gu) {
    const float n = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (!(n > eps)) return scale3(gu, 1.f / eps);
    const f3 u = scale3(v, 1.f / n);
    const float ug = u.x * gu.x + u.y * gu.y + u.z * gu.z;
    return scale3(mk3(gu.x - u.x * ug, gu.y - u.y * ug, gu.z - u.z * ug), 1.f / n);
}
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 cross3p(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// d loss / d eye from the accumulators of camera_backward_kernel: the translation term -sum_v grad_v, plus (look_at mode) the
// path through the rotation rows z = normalize(at - eye), x = normalize(up x z), y = normalize(z x x)   (look_at.py:52-66).
// One thread per batch item; grad_eyes is [B,3], or [3] accumulated over the batch when one eye is shared (eye_stride == 0).
__global__ void __launch_bounds__(128) camera_eye_grad_kernel(const __grid_constant__ CameraParams C, const float* __restrict__ eyes,
                                                              const float* __restrict__ eye_acc, float* __restrict__ grad_eyes, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* a = eye_acc + (size_t)b * 12;
    f3 ge = mk3(-a[0], -a[1], -a[2]);                                   // v - eye
    if (C.mode == 0) {
        const f3 eye = load3(eyes + (size_t)b * C.eye_stride);
        const f3 at = mk3(C.at_or_dir[0], C.at_or_dir[1], C.at_or_dir[2]), up = mk3(C.up[0], C.up[1], C.up[2]);
        const f3 q = sub3(at, eye);
        const f3 z = normalize3(q, 1e-5f);
        const f3 p = cross3(up, z);
        const f3 x = normalize3(p, 1e-5f);
        const f3 s2 = cross3(z, x);
        f3 gx = mk3(a[3], a[4], a[5]), gy = mk3(a[6], a[7], a[8]), gz = mk3(a[9], a[10], a[11]);
        const f3 gs = normalize3_bwd(s2, 1e-5f, gy);                    // y = normalize(z x x)
        gz = add3(gz, cross3p(x, gs));                                  // c = a x b: ga = b x gc, gb = gc x a
        gx = add3(gx, cross3p(gs, z));
        const f3 gp = normalize3_bwd(p, 1e-5f, gx);                     // x = normalize(up x z)
        gz = add3(gz, cross3p(gp, up));
        const f3 gq = normalize3_bwd(q, 1e-5f, gz);                     // z = normalize(at - eye)
        ge = sub3(ge, gq);
    }
    if (C.eye_stride != 0) {
        grad_eyes[(size_t)b * 3 + 0] = ge.x; grad_eyes[(size_t)b * 3 + 1] = ge.y; grad_eyes[(size_t)b * 3 + 2] = ge.z;
    } else {
        atomicAdd(grad_eyes + 0, ge.x); atomicAdd(grad_eyes + 1, ge.y); atomicAdd(grad_eyes + 2, ge.z);
    }
}

// ---- lighting of one face (surface textures) ---------------------------------------------------------------------
// returns light[3]; n / norm / cosine are kept for the backward pass
__device__ __forceinline__ void face_light(const LightParams& L, f3 v0, f3 v1, f3 v2, float light[3], f3* n_out = nullptr,
                                           float* norm_out = nullptr, float* cos_out = nullptr) {
    float norm;
    const f3 n = normalize3(cross3(sub3(v2, v1), sub3(v0, v1)), 1e-6f, &norm);     // mesh.py:105-108
    const float c = sum3_torch(__fmul_rn(n.x, L.direction[0]), __fmul_rn(n.y, L.direction[1]), __fmul_rn(n.z, L.direction[2]));
    const float cosine = fmaxf(c, 0.f);                                            // functional/lighting.py:46 (relu)
#pragma unroll
    for (int k = 0; k < 3; ++k)                                                    // :22 and :47, each torch op rounded on its own
        light[k] = __fadd_rn(L.ambient[k], __fmul_rn(L.intensity_dir, __fmul_rn(L.color_dir[k], cosine)));
    if (n_out) *n_out = n;
    if (norm_out) *norm_out = norm;
    if (cos_out) *cos_out = c;
}

__device__ __forceinline__ int clamp_index(int vi, int V) { return min(max(vi, 0), V - 1); }

// forward: lit_textures[b,f,t,:] = textures[b,f,t,:] * light(b,f)      (lighting.py:58)
__global__ void __launch_bounds__(256) lighting_forward_kernel(const __grid_constant__ LightParams L, const float* __restrict__ vertices,
                                                               long long vstride, const int* __restrict__ face_index, long long index_batch_stride,
                                                               const float* __restrict__ textures, float* __restrict__ lit, int B, int V,
                                                               int F, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * F) return;
    const long long b = i / F, f = i - b * F;
    const int* idx = face_index + b * index_batch_stride + f * 3;
    const float* vb = vertices + b * vstride;
    const f3 v0 = load3(vb + (size_t)clamp_index(__ldg(idx + 0), V) * 3), v1 = load3(vb + (size_t)clamp_index(__ldg(idx + 1), V) * 3),
             v2 = load3(vb + (size_t)clamp_index(__ldg(idx + 2), V) * 3);
    float light[3];
    face_light(L, v0, v1, v2, light);
    const float* src = textures + i * T * 3;
    float* dst = lit + i * T * 3;
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int k = 0; k < 3; ++k) dst[t * 3 + k] = __fmul_rn(__ldg(src + t * 3 + k), light[k]);
}

// backward: grad_lit [B,F,T,3] -> grad_textures [B,F,T,3] (store; may be null) and, through the surface normal, atomic
// adds into grad_vertices [B,V,3] (which already holds the camera path's gradient; may be null).
__global__ void __launch_bounds__(256) lighting_backward_kernel(const __grid_constant__ LightParams L, const float* __restrict__ vertices,
                                                                long long vstride, const int* __restrict__ face_index, long long index_batch_stride,
                                                                const float* __restrict__ textures, const float* __restrict__ grad_lit,
                                                                float* __restrict__ grad_textures, float* __restrict__ grad_vertices,
                                                                int B, int V, int F, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * F) return;
    const long long b = i / F, f = i - b * F;
    const int* idx = face_index + b * index_batch_stride + f * 3;
    const int i0 = clamp_index(__ldg(idx + 0), V), i1 = clamp_index(__ldg(idx + 1), V), i2 = clamp_index(__ldg(idx + 2), V);
    const float* vb = vertices + b * vstride;
    const f3 v0 = load3(vb + (size_t)i0 * 3), v1 = load3(vb + (size_t)i1 * 3), v2 = load3(vb + (size_t)i2 * 3);
    float light[3], norm, c;
    f3 n;
    face_light(L, v0, v1, v2, light, &n, &norm, &c);
    float G[3] = {0.f, 0.f, 0.f};                                    // d loss / d light
    const float* tex = textures + i * T * 3;
    const float* gl = grad_lit + i * T * 3;
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float g = __ldg(gl + t * 3 + k);
            G[k] = fmaf(g, __ldg(tex + t * 3 + k), G[k]);
            if (grad_textures) grad_textures[i * T * 3 + t * 3 + k] = g * light[k];
        }
    if (!grad_vertices || !(c > 0.f)) return;                        // relu: no gradient at or below 0
    const float gcos = L.intensity_dir * (L.color_dir[0] * G[0] + L.color_dir[1] * G[1] + L.color_dir[2] * G[2]);
    if (gcos == 0.f) return;
    const f3 gn = mk3(gcos * L.direction[0], gcos * L.direction[1], gcos * L.direction[2]);
    f3 gc;                                                            // gradient w.r.t. the un-normalised cross product
    if (norm > 1e-6f) gc = scale3(sub3(gn, scale3(n, dot3(n, gn))), 1.f / norm);
    else gc = scale3(gn, 1e6f);                                       // clamped norm: n = c / eps
    const f3 a = sub3(v2, v1), e = sub3(v0, v1);                      // c = a x e
    const f3 ga = cross3(e, gc), ge = cross3(gc, a);
    float* gv = grad_vertices + b * vstride;       // shared mesh (vstride == 0): straight into the batch-summed [V,3]
    atomicAdd(gv + (size_t)i2 * 3 + 0, ga.x); atomicAdd(gv + (size_t)i2 * 3 + 1, ga.y); atomicAdd(gv + (size_t)i2 * 3 + 2, ga.z);
    atomicAdd(gv + (size_t)i0 * 3 + 0, ge.x); atomicAdd(gv + (size_t)i0 * 3 + 1, ge.y); atomicAdd(gv + (size_t)i0 * 3 + 2, ge.z);
    atomicAdd(gv + (size_t)i1 * 3 + 0, -(ga.x + ge.x)); atomicAdd(gv + (size_t)i1 * 3 + 1, -(ga.y + ge.y));
    atomicAdd(gv + (size_t)i1 * 3 + 2, -(ga.z + ge.z));
}

// ---- lighting of vertex textures ------------------------------------------------------------------------------------
// gendr/lighting.py:60-66 with mesh.vertex_normals = gendr/functional/vertex_normals.py:11-49: every face adds the (un-normalised)
// cross product taken at each of its corners to that corner's vertex (three index_add_ calls = atomics in arbitrary order there
// as here), the sums are F.normalize'd, and the vertex colour is multiplied by ambient + directional light.
// normal_sums [B,V,3] is scratch the caller zero-fills (the API entry does); it is kept for the backward pass.
__device__ __forceinline__ void atomic_add3(float* p, f3 v) { atomicAdd(p + 0, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); }

__global__ void __launch_bounds__(256) vertex_normal_sums_kernel(const float* __restrict__ vertices, const int* __restrict__ face_index,
                                                                 long long index_batch_stride, float* __restrict__ normal_sums, int B, int V,
                                                                 int F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * F) return;
    const long long b = i / F, f = i - b * F;
    const int* idx = face_index + b * index_batch_stride + f * 3;
    const int i0 = clamp_index(__ldg(idx + 0), V), i1 = clamp_index(__ldg(idx + 1), V), i2 = clamp_index(__ldg(idx + 2), V);
    const float* vb = vertices + b * (long long)V * 3;
    const f3 v0 = load3(vb + (size_t)i0 * 3), v1 = load3(vb + (size_t)i1 * 3), v2 = load3(vb + (size_t)i2 * 3);
    float* nb = normal_sums + b * (long long)V * 3;
    atomic_add3(nb + (size_t)i1 * 3, cross3(sub3(v2, v1), sub3(v0, v1)));      // vertex_normals.py:33-35
    atomic_add3(nb + (size_t)i2 * 3, cross3(sub3(v0, v2), sub3(v1, v2)));      // :36-38
    atomic_add3(nb + (size_t)i0 * 3, cross3(sub3(v1, v0), sub3(v2, v0)));      // :39-41
}

__device__ __forceinline__ void vertex_light(const LightParams& L, f3 s, float light[3], f3* n_out = nullptr, float* norm_out = nullptr,
                                             float* cos_out = nullptr) {
    float norm;
    const f3 n = normalize3(s, 1e-6f, &norm);                                                      // vertex_normals.py:43
    const float c = sum3_torch(__fmul_rn(n.x, L.direction[0]), __fmul_rn(n.y, L.direction[1]), __fmul_rn(n.z, L.direction[2]));
    const float cosine = fmaxf(c, 0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) light[k] = __fadd_rn(L.ambient[k], __fmul_rn(L.intensity_dir, __fmul_rn(L.color_dir[k], cosine)));
    if (n_out) *n_out = n;
    if (norm_out) *norm_out = norm;
    if (cos_out) *cos_out = c;
}

// lit[b,v,:] = textures[b,v,:] * light(b,v)      (lighting.py:66)
__global__ void __launch_bounds__(256) vertex_lighting_forward_kernel(const __grid_constant__ LightParams L, const float* __restrict__ normal_sums,
                                                                      const float* __restrict__ textures, float* __restrict__ lit, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float light[3];
    vertex_light(L, load3(normal_sums + i * 3), light);
#pragma unroll
    for (int k = 0; k < 3; ++k) lit[i * 3 + k] = __fmul_rn(__ldg(textures + i * 3 + k), light[k]);
}

// backward, vertex part: grad_lit [B,V,3] -> grad_textures (store; may be null) and grad_sums [B,V,3] = d loss / d normal_sums (store)
__global__ void __launch_bounds__(256) vertex_lighting_backward_kernel(const __grid_constant__ LightParams L, const float* __restrict__ normal_sums,
                                                                       const float* __restrict__ textures, const float* __restrict__ grad_lit,
                                                                       float* __restrict__ grad_textures, float* __restrict__ grad_sums, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float light[3], norm, c;
    f3 nrm;
    vertex_light(L, load3(normal_sums + i * 3), light, &nrm, &norm, &c);
    float G[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float g = __ldg(grad_lit + i * 3 + k);
        G[k] = g * __ldg(textures + i * 3 + k);
        if (grad_textures) grad_textures[i * 3 + k] = g * light[k];
    }
    if (!grad_sums) return;
    f3 gs = mk3(0.f, 0.f, 0.f);
    if (c > 0.f) {                                                    // relu
        const float gcos = L.intensity_dir * (L.color_dir[0] * G[0] + L.color_dir[1] * G[1] + L.color_dir[2] * G[2]);
        const f3 gn = mk3(gcos * L.direction[0], gcos * L.direction[1], gcos * L.direction[2]);
        if (norm > 1e-6f) gs = scale3(sub3(gn, scale3(nrm, dot3(nrm, gn))), 1.f / norm);
        else gs = scale3(gn, 1e6f);
    }
    grad_sums[i * 3 + 0] = gs.x; grad_sums[i * 3 + 1] = gs.y; grad_sums[i * 3 + 2] = gs.z;
}

// backward, face part: the three corner cross products of every face -> atomic adds into grad_vertices [B,V,3]
__global__ void __launch_bounds__(256) vertex_normal_sums_backward_kernel(const float* __restrict__ vertices, const int* __restrict__ face_index,
                                                                          long long index_batch_stride, const float* __restrict__ grad_sums,
                                                                          float* __restrict__ grad_vertices, int B, int V, int F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * F) return;
    const long long b = i / F, f = i - b * F;
    const int* idx = face_index + b * index_batch_stride + f * 3;
    const int i0 = clamp_index(__ldg(idx + 0), V), i1 = clamp_index(__ldg(idx + 1), V), i2 = clamp_index(__ldg(idx + 2), V);
    const float* vb = vertices + b * (long long)V * 3;
    const float* gb = grad_sums + b * (long long)V * 3;
    const f3 v0 = load3(vb + (size_t)i0 * 3), v1 = load3(vb + (size_t)i1 * 3), v2 = load3(vb + (size_t)i2 * 3);
    const f3 g0 = load3(gb + (size_t)i0 * 3), g1 = load3(gb + (size_t)i1 * 3), g2 = load3(gb + (size_t)i2 * 3);
    // c = a x e  =>  d/da = e x gc,  d/de = gc x a
    f3 d0 = mk3(0.f, 0.f, 0.f), d1 = d0, d2 = d0;
    {   // at v1: a = v2 - v1, e = v0 - v1
        const f3 a = sub3(v2, v1), e = sub3(v0, v1), ga = cross3(e, g1), ge = cross3(g1, a);
        d2 = add3(d2, ga); d0 = add3(d0, ge); d1 = sub3(d1, add3(ga, ge));
    }
    {   // at v2: a = v0 - v2, e = v1 - v2
        const f3 a = sub3(v0, v2), e = sub3(v1, v2), ga = cross3(e, g2), ge = cross3(g2, a);
        d0 = add3(d0, ga); d1 = add3(d1, ge); d2 = sub3(d2, add3(ga, ge));
    }
    {   // at v0: a = v1 - v0, e = v2 - v0
        const f3 a = sub3(v1, v0), e = sub3(v2, v0), ga = cross3(e, g0), ge = cross3(g0, a);
        d1 = add3(d1, ga); d2 = add3(d2, ge); d0 = sub3(d0, add3(ga, ge));
    }
    float* gv = grad_vertices + b * (long long)V * 3;
    atomic_add3(gv + (size_t)i0 * 3, d0); atomic_add3(gv + (size_t)i1 * 3, d1); atomic_add3(gv + (size_t)i2 * 3, d2);
}

}  // namespace gendr
