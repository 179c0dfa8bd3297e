// oracle/ref_shim.h -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// A CPU execution shim for the *unmodified* reference CUDA translation unit
// (/root/reference/gendr/cuda/generalized_renderer_cuda_kernel.cu, lines 13-1068: the device helpers,
// the 18 distributions, the 9 t-conorms and the three __global__ kernels).  The reference text is never
// copied into this repository: oracle/Makefile streams that line range from /root/reference straight into
// the compiler's stdin, prefixed by this header and followed by ref_driver.inc.  The result
// (oracle/_ref/libgendr_ref_cpu.so, git-ignored) is "the reference itself, run here" and is what pins the
// C restatement in oracle/gendr_oracle.c.
//
// Semantics reproduced from CUDA C++:
//   * min/max overloads incl. the mixed float/double ones (promote to double), fmin/fmax NaN behaviour;
//   * exp/sqrt/pow/... overload sets visible at global scope (float args -> float versions, as in CUDA);
//   * normcdf(float|double) (CUDA-only function);
//   * blockIdx/blockDim/threadIdx as thread-local variables set by the driver loop;
//   * atomicAdd as an OpenMP atomic.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <limits>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__

struct ShimDim3 { unsigned x, y, z; };
static thread_local ShimDim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {256, 1, 1};

using std::exp; using std::sqrt; using std::pow; using std::log; using std::log1p; using std::atan;
using std::tanh; using std::cosh; using std::asin; using std::erfc; using std::tgamma; using std::copysign;

static inline float  max(float a, float b)   { return fmaxf(a, b); }
static inline float  min(float a, float b)   { return fminf(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(float a, double b)  { return fmax((double)a, b); }
static inline double max(double a, float b)  { return fmax(a, (double)b); }
static inline double min(float a, double b)  { return fmin((double)a, b); }
static inline double min(double a, float b)  { return fmin(a, (double)b); }
static inline int    max(int a, int b)       { return a > b ? a : b; }
static inline int    min(int a, int b)       { return a < b ? a : b; }

static inline float  normcdf(float x)  { return 0.5f * erfcf(-x * 0.70710678118654752440f); }
static inline double normcdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

template <typename T> static inline T atomicAdd(T* addr, T val) {
    T old;
#pragma omp atomic capture
    { old = *addr; *addr += val; }
    return old;
}
