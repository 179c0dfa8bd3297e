/* oracle/gendr_voxel_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Plain-C restatement of the reference voxelizer: gendr.functional.voxelization(faces, size, normalize=False)
 * (/root/reference/gendr/functional/voxelization.py:45-62) over the four kernels of
 * /root/reference/gendr/cuda/voxelization_cuda_kernel.cu (sub1 :36-93, sub2 :96-124, sub3 :126-149, sub4 :151-194).
 * Pinned: mode 0 is asserted bit-identical to the unmodified reference kernels run on the CPU (oracle/_ref/
 * libgendr_ref_voxel_cpu.so, built by oracle/Makefile where /root/reference exists) and to tests/golden/voxel_v1.npz
 * generated from that build (tests/test_voxel_cpu.py).
 *
 *   mode 0  C semantics of the source as written (no FMA contraction; what the CPU shim build computes)
 *   mode 1  the contraction nvcc 12.9 emitted for the reference on sm_100a, read from its SASS:
 *             det = fma(x1d, y2d, -(y1d*x2d));  t1 = fma(y2d, xpd, -(x2d*ypd)) / det;  t2 = fma(x1d, ypd, -(y1d*xpd)) / det
 *             z   = face[2] + fma(z1d, t1, z2d*t2)
 *           -- what the reference computes ON THE GPU, and what the CUDA product mirrors.
 *
 * The flood fill (sub3 + the sub4 loop) is a monotone fixed-point iteration whose result does not depend on the update
 * order: the empty voxels 6-connected to an empty boundary voxel.  It is computed here with an explicit queue. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int g_mode = 0;
void gendr_voxel_oracle_set_mode(int mode) { g_mode = mode; }

/* float -> int of `int zi = floor(x)`: mode 1 as the GPU's F2I.FLOOR does (saturating, NaN -> 0), mode 0 as x86's cvttss2si
 * does in the CPU shim build (out of range and NaN -> INT_MIN) */
static int f2i_floor(float x) {
    const float z = floorf(x);
    if (z != z) return g_mode == 0 ? (-2147483647 - 1) : 0;
    if (z >= 2147483648.f) return g_mode == 0 ? (-2147483647 - 1) : 2147483647;
    if (z < -2147483648.f) return -2147483647 - 1;
    return (int)z;
}

static void set_voxel(int32_t* vox, int vs, int yi, int xi, int zi) {
    if (0 <= yi && yi < vs && 0 <= xi && xi < vs && 0 <= zi && zi < vs) vox[((long)yi * vs + xi) * vs + zi] = 1;
}

/* one ray-casting pass (sub1, :36-93) along coordinate `ray`; `yr`, `xr` are the coordinates playing the kernel's y / x */
static void sub1_pass(const float* faces, int F, int vs, int yr, int xr, int ray, int32_t* occ) {
    for (int y = 0; y < vs; ++y) for (int x = 0; x < vs; ++x) for (int fn = 0; fn < F; ++fn) {
        const float* f = faces + (long)fn * 9;
        const float f0 = f[yr], f1 = f[xr], f2 = f[ray];
        const float y1d = f[3 + yr] - f0, x1d = f[3 + xr] - f1, z1d = f[3 + ray] - f2;
        const float y2d = f[6 + yr] - f0, x2d = f[6 + xr] - f1, z2d = f[6 + ray] - f2;
        const float ypd = (float)y - f0, xpd = (float)x - f1;
        float det, t1, t2, z;
        if (g_mode == 0) {
            det = x1d * y2d - x2d * y1d;
            if (det == 0) continue;
            t1 = (y2d * xpd - x2d * ypd) / det;
            t2 = (-y1d * xpd + x1d * ypd) / det;
        } else {
            det = fmaf(x1d, y2d, -(y1d * x2d));
            if (det == 0) continue;
            t1 = fmaf(y2d, xpd, -(x2d * ypd)) / det;
            t2 = fmaf(x1d, ypd, -(y1d * xpd)) / det;
        }
        if (t1 < 0) continue;
        if (t2 < 0) continue;
        if (1 < t1 + t2) continue;
        z = (g_mode == 0) ? (t1 * z1d + t2 * z2d + f2) : (f2 + fmaf(z1d, t1, z2d * t2));
        const int zi = f2i_floor(z);
        for (int dy = 0; dy < 2; ++dy) for (int dx = 0; dx < 2; ++dx) {
            int c[3];
            c[yr] = y - dy; c[xr] = x - dx; c[ray] = zi;
            if (0 <= y - dy && y - dy < vs && 0 <= x - dx && x - dx < vs) set_voxel(occ, vs, c[0], c[1], c[2]);
        }
    }
}

/* faces [B,F,3,3] as passed to voxelization(); voxels_out int32 [B,vs,vs,vs] */
int gendr_voxel_oracle(const float* faces_in, int32_t* voxels_out, int B, int F, int vs) {
    const long n = (long)vs * vs * vs;
    float* faces = (float*)malloc(sizeof(float) * (size_t)(F > 0 ? F : 1) * 9);
    int32_t* occ = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    int32_t* vis = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    long* queue = (long*)malloc(sizeof(long) * (size_t)n);
    for (int b = 0; b < B; ++b) {
        for (long i = 0; i < (long)F * 9; ++i) faces[i] = faces_in[(long)b * F * 9 + i] * (float)vs;   /* voxelization.py:50 */
        memset(occ, 0, sizeof(int32_t) * (size_t)n);
        sub1_pass(faces, F, vs, 2, 1, 0, occ);      /* dim 0: faces[..., [2,1,0]], result transposed back (voxelization.py:14-19) */
        sub1_pass(faces, F, vs, 0, 2, 1, occ);      /* dim 1: faces[..., [0,2,1]] */
        sub1_pass(faces, F, vs, 0, 1, 2, occ);      /* dim 2 */
        for (long v = 0; v < (long)F * 3; ++v) {    /* sub2 (:96-124): the voxel containing each vertex */
            const float* p = faces + v * 3;
            set_voxel(occ, vs, f2i_floor(p[0]), f2i_floor(p[1]), f2i_floor(p[2]));
        }
        /* sub3 + sub4 loop: empty voxels connected to an empty boundary voxel become visible */
        memset(vis, 0, sizeof(int32_t) * (size_t)n);
        long head = 0, tail = 0;
        for (int y = 0; y < vs; ++y) for (int x = 0; x < vs; ++x) for (int z = 0; z < vs; ++z) {
            const long i = ((long)y * vs + x) * vs + z;
            if ((y == 0 || y == vs - 1 || x == 0 || x == vs - 1 || z == 0 || z == vs - 1) && occ[i] == 0) { vis[i] = 1; queue[tail++] = i; }
        }
        while (head < tail) {
            const long i = queue[head++];
            const int z = (int)(i % vs), x = (int)((i / vs) % vs), y = (int)(i / ((long)vs * vs));
            const int dy[6] = {-1, 1, 0, 0, 0, 0}, dx[6] = {0, 0, -1, 1, 0, 0}, dz[6] = {0, 0, 0, 0, -1, 1};
            for (int k = 0; k < 6; ++k) {
                const int yy = y + dy[k], xx = x + dx[k], zz = z + dz[k];
                if (yy <= 0 || yy >= vs - 1 || xx <= 0 || xx >= vs - 1 || zz <= 0 || zz >= vs - 1) continue;   /* sub4 skips boundary voxels */
                const long j = ((long)yy * vs + xx) * vs + zz;
                if (occ[j] == 0 && vis[j] == 0) { vis[j] = 1; queue[tail++] = j; }
            }
        }
        for (long i = 0; i < n; ++i) voxels_out[(long)b * n + i] = 1 - vis[i];
    }
    free(faces); free(occ); free(vis); free(queue);
    return 0;
}
