// ps_kernels.cuh -- sm_100a kernels of the pictorial-structures message path.
//
// Parity rules shared by every kernel (DESIGN.md "Arithmetic contract"):
//   * grid data is fp32; a tap sum is acc = 0; acc = acc + x[k]*f[k] for k ascending with a separately
//     rounded multiply and add (__fmul_rn/__fadd_rn: never contracted to FMA) -- the order cblas_sdot
//     (Netlib) gives the reference (multi_array_filter.hpp:153,285,314).  Skipping terms whose data or tap
//     is exactly 0 is allowed (all data are >= +0, so acc + 0 == acc bit for bit); re-association is not.
//   * exp / log are evaluated in fp64 and narrowed (multi_array_op.hpp:165,177 call the double libm
//     routines on floats).
//   * coordinates of affine resampling are fp64 with separately rounded multiply/add
//     (homogeneous_coord.h:79-80), weights narrowed to fp32 (multi_array_transform.hpp:218-230).
// The file is compiled with -fmad=false as a second line of defence.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace psk {

constexpr float kLogZero = -1e6f;

// ---- monotone float <-> int encoding for atomicMax ------------------------------------------------
__host__ __device__ inline int enc_f(float f) {
#ifdef __CUDA_ARCH__
  int i = __float_as_int(f);
#else
  int i;
  memcpy(&i, &f, 4);
#endif
  return i >= 0 ? i : (i ^ 0x7fffffff);
}
__host__ __device__ inline float dec_f(int i) {
  int j = i >= 0 ? i : (i ^ 0x7fffffff);
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, 4);
  return f;
#endif
}
#define PS_ENC_NEG_INF ((int)0x807fffff) /* enc_f(-inf) = 0xff800000 ^ 0x7fffffff */

// ---- exp / log of the reference: fp64 libm routine, narrowed to fp32 -------------------------------------------
// Reference semantics: (float)exp((double)x), (float)log((double)x) (multi_array_op.hpp:165,177).
// *_slow call CUDA's fp64 libm.  *_fast use a short table-driven fp64 evaluation (relative error < 2^-50) and fall
// back to *_slow whenever the fp64 result lies within kGuard units (of 2^-29 fp32 ulp) of an fp32 rounding boundary,
// so fast == slow for every input (checked exhaustively by ps_selftest_math / tests/test_gpu_math.py).
// Tables are filled by the host at ps_create (ps_math_tables_init).
__device__ double2 d_log_tab[129];   // .x = 2^-23/F_j, .y = log(F_j) - (j >= 54 ? ln2 : 0),  F_j = 1 + j/128
__device__ double d_eln2_tab[512];   // [b] = RN((b - 127)*ln2), b = biased exponent (+1 when j >= 54); every 9-bit index is in range
__device__ double d_exp_tab[64];     // 2^(j/64)
constexpr int kGuard = 64;

__device__ __forceinline__ float exp_slow(float x) {
  if (x < -104.0f) return 0.0f;  // the double result is < 2^-150 and narrows to +0 exactly
  return (float)exp((double)x);
}
__device__ __forceinline__ float log_slow(float d) {  // computeLogGrid cell (multi_array_op.hpp:162-165)
  return d == 0.0f ? kLogZero : (float)log((double)d);
}

__device__ __forceinline__ bool near_f32_boundary(double v) {
  // bits below the fp32 mantissa: 29 of them; the round-to-nearest boundary is the pattern 0x10000000
  int lo = __double2loint(v) & 0x1fffffff;
  return abs(lo - 0x10000000) <= kGuard;
}

__device__ __forceinline__ float exp_fast(float x) {
  if (x < -104.0f) return 0.0f;
  if (!(x > -87.0f && x < 88.0f)) return (float)exp((double)x);  // fp32-subnormal results, overflow, NaN
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double xd = (double)x;
  const double t = __fma_rn(xd, 92.332482616893656877, MAGIC);  // 64/ln2
  const int k = __double2loint(t);
  const double kd = t - MAGIC;
  // ln2/64 = hi + lo; hi = 0x1.62e42fe8p-7 carries 30 significant bits, so kd*hi is exact for |k| < 2^22
  double r = __fma_rn(kd, -0x1.62e42fe800000p-7, xd);
  r = __fma_rn(kd, -0x1.e8e7bcd5e4f1ep-37, r);
  double p = __fma_rn(r, 8.3333333333333332177e-03, 4.1666666666666664354e-02);
  p = __fma_rn(r, p, 1.6666666666666665741e-01);
  p = __fma_rn(r, p, 0.5);
  p = __fma_rn(r, p, 1.0);
  p = __fma_rn(r, p, 1.0);
  double res = d_exp_tab[k & 63] * p;
  res = __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
  if (near_f32_boundary(res)) return (float)exp((double)x);
  return (float)res;
}

// Branch-free form of exp_fast (same arithmetic, same fallback conditions) for kernels that batch several cells per
// thread; `redo` asks the caller to replace the result by exp_slow_call(x).  Memory-safe for any bit pattern.
__device__ __noinline__ float exp_slow_call(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float exp_fast_nb(float x, bool &redo) {
  const double MAGIC = 6755399441055744.0;
  const double xd = (double)x;
  const double t = __fma_rn(xd, 92.332482616893656877, MAGIC);
  const int k = __double2loint(t);
  const double kd = t - MAGIC;
  double r = __fma_rn(kd, -0x1.62e42fe800000p-7, xd);
  r = __fma_rn(kd, -0x1.e8e7bcd5e4f1ep-37, r);
  double p = __fma_rn(r, 8.3333333333333332177e-03, 4.1666666666666664354e-02);
  p = __fma_rn(r, p, 1.6666666666666665741e-01);
  p = __fma_rn(r, p, 0.5);
  p = __fma_rn(r, p, 1.0);
  p = __fma_rn(r, p, 1.0);
  double res = d_exp_tab[k & 63] * p;
  res = __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
  const bool tiny = x < -104.0f;
  const bool inr = (x > -87.0f) & (x < 88.0f);
  redo = (!tiny) & ((!inr) | near_f32_boundary(res));
  return tiny ? 0.0f : (float)res;
}

// Table-driven fp64 log of a positive normal fp32 given by its bits.  Safe (if meaningless) for any bit pattern.
//   d = 2^e * m, m in [1,2);  F_j = nearest multiple of 1/128 to m;  log d = (e + c)*ln2 + (log F_j - c*ln2) + log1p(r),
//   r = (m - F_j)/F_j, |r| < 2^-8, c = (j >= 54) keeps the two big terms from cancelling near d = 1.
// m - F_j is an exact multiple of 2^-23, so r = RN(int(mant - j*2^16) * (2^-23/F_j)) equals RN((m - F_j) * (1/F_j)).
__device__ __forceinline__ double log_core(unsigned ib) {
  const unsigned mant = ib & 0x7fffffu;
  const unsigned j = (mant + 0x8000u) >> 16;  // 0..128
  const int diff = (int)mant - (int)(j << 16);
  const double2 t = d_log_tab[j];
  const double r = (double)diff * t.x;
  double p = __fma_rn(r, -1.6666666666666665741e-01, 0.2);
  p = __fma_rn(r, p, -0.25);
  p = __fma_rn(r, p, 3.3333333333333331483e-01);
  p = __fma_rn(r, p, -0.5);
  p = __fma_rn(r * r, p, r);
  // j >= 54  <=>  mant >= 0x358000: adding 0x800000 - 0x358000 carries exactly then into the exponent field
  return d_eln2_tab[(ib + 0x4a8000u) >> 23] + (t.y + p);
}

__device__ __forceinline__ float log_fast(float d) {
  if (d == 0.0f) return kLogZero;
  const unsigned ib = __float_as_uint(d);
  if (ib - 0x00800000u >= 0x7f000000u) return (float)log((double)d);  // subnormal, inf, NaN, negative
  const double res = log_core(ib);
  if (near_f32_boundary(res)) return (float)log((double)d);
  return (float)res;
}

// Branch-free form of log_fast for kernels that evaluate several cells per thread: always runs log_core and reports
// in `redo` whether the caller must replace the result by log_slow_call(d) -- exactly the conditions under which
// log_fast leaves its fast path.  d == 0 gives LOG_ZERO directly.
__device__ __noinline__ float log_slow_call(float d) { return (float)log((double)d); }
__device__ __forceinline__ float log_fast_nb(float d, bool &redo) {
  const unsigned ib = __float_as_uint(d);
  const double res = log_core(ib);
  const bool zero = d == 0.0f;
  redo = (!zero) & ((ib - 0x00800000u >= 0x7f000000u) | near_f32_boundary(res));
  return zero ? kLogZero : (float)res;
}

#ifndef PS_SLOW_MATH
__device__ __forceinline__ float exp_f64(float x) { return exp_fast(x); }
__device__ __forceinline__ float log_f64(float d) { return log_fast(d); }
#else
__device__ __forceinline__ float exp_f64(float x) { return exp_slow(x); }
__device__ __forceinline__ float log_f64(float d) { return log_slow(d); }
#endif

// out[i] = f(bits first + i): op 0 exp_f64 (the path's exp), 1 log_f64, 2 exp_slow, 3 log_slow.
__global__ void k_eval_math(int op, unsigned first, unsigned count, float *out) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned stride = gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    const float x = __uint_as_float(first + i);
    float r;
    if (op == 0) r = exp_f64(x);
    else if (op == 1) r = log_f64(x);
    else if (op == 2) r = exp_slow(x);
    else r = log_slow(x);
    out[i] = r;
  }
}

// Exhaustive self-test: every fp32 bit pattern in [first, first + count).
// out[0]: exp mismatches, out[1]: exp inputs taking the fast path, out[2]: log mismatches, out[3]: log fast path.
__global__ void k_selftest_math(unsigned first, unsigned long long count, unsigned long long *out) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long bad_e = 0, fast_e = 0, bad_l = 0, fast_l = 0;
  for (; i < count; i += stride) {
    const float x = __uint_as_float(first + (unsigned)i);
    if (x == x) {
      // both spellings of each fast routine -- the scalar one and the branch-free one of the batch kernels (with its
      // patch-up call) -- against the libm routine
      if (fabsf(x) <= 104.0f) {
        const float a = exp_fast(x), b = exp_slow(x);
        bool redo;
        float c = exp_fast_nb(x, redo);
        if (redo) c = exp_slow_call(x);
        if (__float_as_uint(a) != __float_as_uint(b) || __float_as_uint(c) != __float_as_uint(b)) ++bad_e;
        ++fast_e;
      }
      if (x >= 0.0f && x < INFINITY) {
        const float a = log_fast(x), b = log_slow(x);
        bool redo;
        float c = log_fast_nb(x, redo);
        if (redo) c = log_slow_call(x);
        if (__float_as_uint(a) != __float_as_uint(b) || __float_as_uint(c) != __float_as_uint(b)) ++bad_l;
        ++fast_l;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    bad_e += __shfl_xor_sync(0xffffffffu, bad_e, o);
    fast_e += __shfl_xor_sync(0xffffffffu, fast_e, o);
    bad_l += __shfl_xor_sync(0xffffffffu, bad_l, o);
    fast_l += __shfl_xor_sync(0xffffffffu, fast_l, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 0, bad_e);
    atomicAdd(out + 1, fast_e);
    atomicAdd(out + 2, bad_l);
    atomicAdd(out + 3, fast_l);
  }
}

// Operand of an integer max that reproduces the fmaxf folds: NaN never beats a number.
__device__ __forceinline__ int enc_max_operand(float v) { return v == v ? enc_f(v) : PS_ENC_NEG_INF; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide max folded into *dst (encoded) with one atomic per block. All threads must call.
__device__ __forceinline__ void block_max_to(float v, int *dst) {
  __shared__ float s_part[32];
  v = warp_max(v);
  const int tid_ = threadIdx.y * blockDim.x + threadIdx.x;
  int lane = tid_ & 31, w = tid_ >> 5;
  int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
  if (lane == 0) s_part[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? s_part[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) atomicMax(dst, enc_f(t));
  }
}

// First maximum in flat order (findrot.cpp:261-277) as a single 64-bit max: key = (ordered(value) << 32) | ~index
// picks the largest value and, among equals, the smallest index.  -0.0 and +0.0 compare equal in the reference, so
// -0.0 is canonicalised to +0.0 before encoding; NaN never wins a '>' comparison and is skipped by the callers.
__device__ __forceinline__ unsigned long long argmax_key(float v, unsigned idx) {
  if (v == 0.0f) v = 0.0f;
  unsigned e = (unsigned)enc_f(v) ^ 0x80000000u;  // order-preserving unsigned
  return ((unsigned long long)e << 32) | (unsigned)(~idx);
}
// Block-wide max of keys folded into *dst with one atomic per block. All threads must call.
__device__ __forceinline__ void block_key_max_to(unsigned long long best, unsigned long long *dst) {
  __shared__ unsigned long long s_keys[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t > best ? t : best;
  }
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int lane = tid & 31, w = tid >> 5;
  const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
  if (lane == 0) s_keys[w] = best;
  __syncthreads();
  if (w == 0) {
    unsigned long long t = lane < nw ? s_keys[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long u = __shfl_xor_sync(0xffffffffu, t, o);
      t = u > t ? u : t;
    }
    if (lane == 0 && t) atomicMax(dst, t);
  }
}

// Exact unsigned 32-bit division by a run-time constant (Granlund-Montgomery): 4 instructions instead of the
// ~25 of the software divide.  Valid for every n < 2^32, d >= 1.
struct FastDiv {
  unsigned d, m, s1, s2;
  FastDiv() : d(1), m(1), s1(0), s2(0) {}
  explicit FastDiv(unsigned div) : d(div) {
    unsigned l = 0;
    while ((1ull << l) < div) ++l;
    m = (unsigned)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
    s1 = l < 1 ? l : 1;
    s2 = l > 0 ? l - 1 : 0;
  }
#ifdef __CUDACC__
  __device__ __forceinline__ unsigned div(unsigned n) const {
    const unsigned t = __umulhi(m, n);
    return (t + ((n - t) >> s1)) >> s2;
  }
#endif
};

// ---- pointwise sweeps ------------------------------------------------------------------------------

// setGrid (multi_array_op.hpp:80-90) with 128-bit stores; optionally also resets a maximum slot (saves the separate
// one-thread launch in front of every ingest).
__global__ void __launch_bounds__(256) k_fill(float *__restrict__ p, size_t n, float v, int *slot = nullptr, int slot_init = 0) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if (slot && i == 0) *slot = slot_init;
  const size_t n4 = ((uintptr_t)p % 16 == 0) ? n / 4 : 0;
  float4 *p4 = reinterpret_cast<float4 *>(p);
  const float4 v4 = make_float4(v, v, v, v);
  for (size_t j = i; j < n4; j += stride) p4[j] = v4;
  for (size_t j = n4 * 4 + i; j < n; j += stride) p[j] = v;
}

__global__ void k_set_int(int *p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// clip_scores_fill (objectdetect_aux.hpp:42-59) + computeLogGrid (multi_array_op.hpp:154-167) of one cell
__device__ __forceinline__ float prepare_cell(float v) {
  if (v < 0.0f) v = (float)0.0001;
  return log_f64(v);
}
// In place over n floats (n4 = n / 4 handled as float4, the tail scalar); also folds max(result) into *dst if given,
// which is the getMinMax a leaf's upward message would otherwise spend a pass on.
__global__ void __launch_bounds__(256) k_prepare_unary(float *__restrict__ p, size_t n, int *dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = ((uintptr_t)p % 16 == 0) ? n / 4 : 0;
  float m = -INFINITY;
  float4 *p4 = reinterpret_cast<float4 *>(p);
  for (size_t j = i; j < n4; j += stride) {
    float4 v = p4[j];
    v.x = prepare_cell(v.x); v.y = prepare_cell(v.y); v.z = prepare_cell(v.z); v.w = prepare_cell(v.w);
    p4[j] = v;
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (size_t j = n4 * 4 + i; j < n; j += stride) {
    float v = prepare_cell(p[j]);
    p[j] = v;
    m = fmaxf(m, v);
  }
  if (dst) block_max_to(m, dst);
}

// computeExpGrid followed by computeLogGrid in place (what computePosJointMarginal leaves in its input,
// objectdetect_findpos.cpp:76,88)
__global__ void __launch_bounds__(256) k_exp_log(float *__restrict__ p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = log_f64(exp_f64(p[i]));
}

// getMinMax (multi_array_op.hpp:61-77), max only
__global__ void __launch_bounds__(256) k_grid_max(const float *__restrict__ p, size_t n, int *dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = ((uintptr_t)p % 16 == 0) ? n / 4 : 0;
  float m = -INFINITY;
  const float4 *p4 = reinterpret_cast<const float4 *>(p);
  for (size_t j = i; j < n4; j += stride) {
    float4 v = __ldg(p4 + j);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (size_t j = n4 * 4 + i; j < n; j += stride) m = fmaxf(m, p[j]);
  block_max_to(m, dst);
}


// ---- unary ingest from compact detector grids: PartApp::loadScoreGrid (libPartApp/partapp.cpp:830-903) ------------
// The reference scatters every evaluated cell of the compact grid (value != NO_CLASS_VALUE = 0) into a zeroed
// image-size grid with TM_DIRECT (transform.hpp:167-192; x1 outer, y1 inner, later writers overwrite), then applies
// clip_scores_fill + computeLogGrid (findrot.cpp:834-845).  Device version: (1) every evaluated compact cell does
// atomicMax(order key) on its target image cell, order = x1 * gh + y1 + 1, so the surviving key is the reference's
// last writer; (2) a sweep turns keys into log-unaries (0 -> LOG_ZERO) and folds their maximum.
constexpr int kMaxIngestRot = 64;
struct TigRows {
  double m[kMaxIngestRot * 6];  // passed by value as a kernel parameter: no staging copy, no host synchronisation
};
struct IngestArgs {
  const float *cells;   // [R][gh][gw]
  const double *Tig;    // [R][6]: rows 0,1 of the 3x3 image<-grid transform (Ti2 * T2g); null -> use the by-value rows
  int *keys;            // [R][H][W] scratch
  float *out;           // [R][H][W] log-domain unary
  int R, gh, gw, H, W;
  int raw;              // 1: stop after clip_scores_fill (scores stay in the probability domain, unevaluated cells 0):
                        // what findObjectRoiHelper takes its detection maxima from (objectdetect_roi.cpp:215-236)
};
__device__ __forceinline__ float ingest_cell(float v, int raw) {
  if (raw) return v < 0.0f ? (float)0.0001 : v;
  return prepare_cell(v);
}

__global__ void __launch_bounds__(256) k_ingest_scatter(IngestArgs a, const __grid_constant__ TigRows rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // flat over gh*gw, x fastest (coalesced read)
  const int r = blockIdx.y;
  if (i >= a.gh * a.gw) return;
  const int y1 = i / a.gw, x1 = i - y1 * a.gw;
  const float v = a.cells[(size_t)r * a.gh * a.gw + i];
  if (v == 0.0f) return;
  const double *T = a.Tig ? a.Tig + r * 6 : rows.m + r * 6;
  const double x3 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], (double)x1), __dmul_rn(T[1], (double)y1)), T[2]);
  const double y3 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], (double)x1), __dmul_rn(T[4], (double)y1)), T[5]);
  const int ix = (int)floor(__dadd_rn(x3, 0.5)), iy = (int)floor(__dadd_rn(y3, 0.5));
  if (ix >= 0 && ix < a.W && iy >= 0 && iy < a.H)
    atomicMax(&a.keys[(size_t)r * a.H * a.W + (size_t)iy * a.W + ix], x1 * a.gh + y1 + 1);
}

// Collision-free variant: when the host has proved that distinct grid cells cannot land on the same image cell
// (smallest singular value of the 2x2 part of every Tig > sqrt(2): two lattice points are then further apart than
// the diagonal of a cell), the order of the writers is irrelevant, so evaluated cells are written straight into a
// LOG_ZERO-filled unary -- one 23 MB fill instead of a key pass, a key sweep and their traffic.
__global__ void __launch_bounds__(256) k_ingest_scatter_direct(IngestArgs a, const __grid_constant__ TigRows rows,
                                                               int *max_dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  float m = a.raw ? 0.0f : kLogZero;  // unevaluated cells are LOG_ZERO (the fill)
  if (i < a.gh * a.gw) {
    const int y1 = i / a.gw, x1 = i - y1 * a.gw;
    const float v = a.cells[(size_t)r * a.gh * a.gw + i];
    if (v != 0.0f) {
      const double *T = a.Tig ? a.Tig + r * 6 : rows.m + r * 6;
      const double x3 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], (double)x1), __dmul_rn(T[1], (double)y1)), T[2]);
      const double y3 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], (double)x1), __dmul_rn(T[4], (double)y1)), T[5]);
      const int ix = (int)floor(__dadd_rn(x3, 0.5)), iy = (int)floor(__dadd_rn(y3, 0.5));
      if (ix >= 0 && ix < a.W && iy >= 0 && iy < a.H) {
        const float o = ingest_cell(v, a.raw);
        a.out[(size_t)r * a.H * a.W + (size_t)iy * a.W + ix] = o;
        m = fmaxf(m, o);
      }
    }
  }
  if (max_dst) block_max_to(m, max_dst);
}

// The collision-free scatter for several (part, scale) grids that share one lattice (same gh x gw and Tig: every part
// of an image is evaluated on the same detector grid): blockIdx.z selects the grid.  One launch per image instead of
// one per part.
constexpr int kMaxIngestBatch = 32;
struct IngestBatch {
  const float *cells[kMaxIngestBatch];
  float *out[kMaxIngestBatch];
  int *max_dst[kMaxIngestBatch];
};
// ALL: the destination already holds "fill + this lattice" (the image before used the same transforms): unevaluated
// cells write LOG_ZERO over whatever the last image left on their lattice point, and no fill pass is needed.
template <bool ALL>
__global__ void __launch_bounds__(256) k_ingest_scatter_direct_b(IngestArgs a, const __grid_constant__ IngestBatch b,
                                                                 const __grid_constant__ TigRows rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y, g = blockIdx.z;
  float m = kLogZero;
  if (i < a.gh * a.gw) {
    const int y1 = i / a.gw, x1 = i - y1 * a.gw;
    const float v = b.cells[g][(size_t)r * a.gh * a.gw + i];
    if (ALL || v != 0.0f) {
      const double *T = rows.m + r * 6;
      const double x3 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], (double)x1), __dmul_rn(T[1], (double)y1)), T[2]);
      const double y3 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], (double)x1), __dmul_rn(T[4], (double)y1)), T[5]);
      const int ix = (int)floor(__dadd_rn(x3, 0.5)), iy = (int)floor(__dadd_rn(y3, 0.5));
      if (ix >= 0 && ix < a.W && iy >= 0 && iy < a.H) {
        const float o = (!ALL || v != 0.0f) ? prepare_cell(v) : kLogZero;
        b.out[g][(size_t)r * a.H * a.W + (size_t)iy * a.W + ix] = o;
        m = fmaxf(m, o);
      }
    }
  }
  block_max_to(m, b.max_dst[g]);
}

__device__ __forceinline__ float bilinear_at(const float *__restrict__ s, int h, int w, int pitch, double x1, double y1);

// ExpParam.interpolate: TM_BILINEAR gather (transform.hpp:196-238) of every image cell through T13 = inverse(Tig),
// default NO_CLASS_VALUE = 0, followed by the unary prep.  `rows` / a.Tig hold rows 0,1 of T13 here.
__global__ void __launch_bounds__(256) k_ingest_bilinear(IngestArgs a, const __grid_constant__ TigRows rows, FastDiv Wdiv,
                                                         int *max_dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  float m = -INFINITY;
  if (i < a.H * a.W) {
    const int y3 = (int)Wdiv.div((unsigned)i), x3 = i - y3 * a.W;
    const double *T = a.Tig ? a.Tig + r * 6 : rows.m + r * 6;
    const double x1 = __dadd_rn(__dadd_rn(__dmul_rn(T[0], (double)x3), __dmul_rn(T[1], (double)y3)), T[2]);
    const double y1 = __dadd_rn(__dadd_rn(__dmul_rn(T[3], (double)x3), __dmul_rn(T[4], (double)y3)), T[5]);
    const float v = bilinear_at(a.cells + (size_t)r * a.gh * a.gw, a.gh, a.gw, a.gw, x1, y1);
    m = ingest_cell(v, a.raw);
    a.out[(size_t)r * a.H * a.W + i] = m;
  }
  if (max_dst) block_max_to(m, max_dst);
}

__global__ void __launch_bounds__(256) k_ingest_sweep(IngestArgs a, int *max_dst) {
  const size_t n = (size_t)a.R * a.H * a.W;
  const size_t HW = (size_t)a.H * a.W;
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  float m = -INFINITY;
  if (i < n) {
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = a.raw ? 0.0f : kLogZero;
      if (i + j < n) {
        const int k = a.keys[i + j];
        if (k > 0) {
          const int r = (int)((i + j) / HW);
          const int x1 = (k - 1) / a.gh, y1 = (k - 1) - x1 * a.gh;
          o[j] = ingest_cell(__ldg(&a.cells[(size_t)r * a.gh * a.gw + (size_t)y1 * a.gw + x1]), a.raw);
        }
        m = fmaxf(m, o[j]);
      }
    }
    if (i + 3 < n && ((uintptr_t)(a.out + i) & 15) == 0) *reinterpret_cast<float4 *>(a.out + i) = make_float4(o[0], o[1], o[2], o[3]);
    else
      for (int j = 0; j < 4 && i + j < n; ++j) a.out[i + j] = o[j];
  }
  if (max_dst) block_max_to(m, max_dst);
}

// Upright masking (findrot.cpp:509-523): slices flagged in mask[r] are set to LOG_ZERO.
__global__ void k_mask_slices(float *__restrict__ g, int R, size_t HW, const unsigned char *__restrict__ mask) {
  int r = blockIdx.y;
  if (!mask[r]) return;
  float *s = g + (size_t)r * HW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < HW; i += stride) s[i] = kLogZero;
}

// Border strip (findrot.cpp:537-549): columns [0,sw) and [W-sw,W) of every row of every slice.
__global__ void k_strip_border(float *__restrict__ g, int rows /* R*H */, int W, int sw) {
  int row = blockIdx.x * blockDim.y + threadIdx.y;
  if (row >= rows) return;
  float *p = g + (size_t)row * W;
  for (int i = threadIdx.x; i < sw; i += blockDim.x) {
    p[i] = kLogZero;
    p[W - sw + i] = kLogZero;
  }
}

// addExtraUnary with broadcast tables (objectdetect_icps.cpp:526-548, :183-190)
//   kind 0: g += w*table[r]; kind 1: g += w*table[y*W+x]; kind 2: g += table[y*W+x]
__global__ void k_add_table(float *__restrict__ g, int R, size_t HW, const float *__restrict__ table, int kind,
                            float w) {
  int r = blockIdx.y;
  float *s = g + (size_t)r * HW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  if (kind == 0) {
    float t = __fmul_rn(w, table[r]);
    for (; i < HW; i += stride) s[i] = __fadd_rn(s[i], t);
  } else if (kind == 1) {
    for (; i < HW; i += stride) s[i] = __fadd_rn(s[i], __fmul_rn(w, table[i]));
  } else {
    for (; i < HW; i += stride) s[i] = __fadd_rn(s[i], table[i]);
  }
}

// Several conditioning tables applied to one unary in ONE pass, in call order (the reference adds the rotation score,
// then the position score, then the torso prior -- findrot.cpp:913-949; every add is its own fp32 rounding, so the
// order is part of the result).  kinds as in k_add_table.
constexpr int kMaxTables = 4;
struct TableArgs {
  const float *table[kMaxTables];
  float weight[kMaxTables];
  int kind[kMaxTables];
  int n;
};
__global__ void k_add_tables(float *__restrict__ g, int R, size_t HW, const __grid_constant__ TableArgs a) {
  const int r = blockIdx.y;
  float *s = g + (size_t)r * HW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < HW; i += stride) {
    float v = s[i];
#pragma unroll
    for (int k = 0; k < kMaxTables; ++k)
      if (k < a.n) {
        if (a.kind[k] == 0) v = __fadd_rn(v, __fmul_rn(a.weight[k], a.table[k][r]));
        else if (a.kind[k] == 1) v = __fadd_rn(v, __fmul_rn(a.weight[k], a.table[k][i]));
        else v = __fadd_rn(v, a.table[k][i]);
      }
    s[i] = v;
  }
}

// DPM score fusion (objectdetect_icps.cpp:445-486 addLoadDPMScore, :488-524 addDPMScore): per-cell adds of a score
// grid g[nrot][H][W] (nrot = R, or 1 broadcast over rotations).
//   mode 0 (addDPMScore, grid already in the log domain):   u += w * g
//   mode 1 (addLoadDPMScore, raw DPM scores):                u += (g > 1e-4 ? w * logf(g) : log(1e-4))
// icps.cpp has "using namespace std", so log(float) is the fp32 logf there; the device evaluates the correctly
// rounded fp32 logarithm (fp64 log narrowed), which equals glibc's logf except on its rare non-correctly-rounded
// inputs (glibc documents < 1 ulp).  The else branch is the double constant log(1e-4) added in double and narrowed.
__global__ void k_add_grid(float *__restrict__ u, int R, size_t HW, const float *__restrict__ g, int nrot, int mode, float w) {
  const int r = blockIdx.y;
  float *s = u + (size_t)r * HW;
  const float *gs = g + (size_t)(nrot == R ? r : 0) * HW;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < HW; i += stride) {
    const float val = gs[i];
    if (mode == 0) {
      s[i] = __fadd_rn(s[i], __fmul_rn(w, val));
    } else if ((double)val > 1e-4) {
      s[i] = __fadd_rn(s[i], __fmul_rn(w, log_f64(val)));
    } else {
      s[i] = (float)__dadd_rn((double)s[i], -9.2103403719761836);  // log(1e-4)
    }
  }
}

// ---- message stage 1: shift + exp + circular rotation filter ------------------------------------------
// findrot.cpp:339-420.  One thread per pixel; the R shifted/exponentiated values of the pixel live in a
// private shared-memory column, then every output rotation is a sequential dot product over the taps.
constexpr int kMaxBatch = 8;  // messages of one tree level that share a launch (blockIdx.y / .z selects the message)
struct RotArgs {
  const float *in;      // [R][H][W] child belief (log domain)
  float *out;           // [R][H][W] rotation-filtered probabilities
  const int *xin;       // [R][W] source x or -1
  const int *yin;       // [R][H] source y or -1
  const int *shift_xy;  // [R][2] (dx, dy) when every table row is a pure shift, else null
  const float *taps;    // [len]
  const int *max_enc;   // encoded max of `in`
  int R, H, W;
  int shift;            // rot_mean_idx
  int mode;             // 0 copy, 1 filter, 2 zero
  int len;
};

struct RotBatch {
  RotArgs a[kMaxBatch];
};
constexpr int kRotThreads = 64;

__global__ void __launch_bounds__(kRotThreads) k_rotconv(RotArgs a) {
  extern __shared__ float s_col[];  // [R][kRotThreads]
  __shared__ float s_taps[1000];
  const int tid = threadIdx.x;
  for (int i = tid; i < a.len; i += kRotThreads) s_taps[i] = a.taps[i];
  __syncthreads();
  const size_t HW = (size_t)a.H * a.W;
  const size_t p = (size_t)blockIdx.x * kRotThreads + tid;
  if (p >= HW) return;
  const int y = (int)(p / a.W), x = (int)(p % a.W);
  const float negM = -dec_f(*a.max_enc);

  for (int ro = 0; ro < a.R; ++ro) {
    int r = ro - a.shift;
    float v = kLogZero;
    if (r >= 0 && r < a.R) {
      int ys = a.yin[r * a.H + y], xs = a.xin[r * a.W + x];
      if ((ys | xs) >= 0) v = __ldg(&a.in[(size_t)r * HW + (size_t)ys * a.W + xs]);
    }
    s_col[ro * kRotThreads + tid] = exp_f64(__fadd_rn(v, negM));
  }
  if (a.mode == 1) {
    const int n = (a.len - 1) / 2;
    for (int i = 0; i < a.R; ++i) {
      int src = (i - n) % a.R;
      if (src < 0) src += a.R;
      float acc = 0.0f;
      for (int k = 0; k < a.len; ++k) {
        acc = __fadd_rn(acc, __fmul_rn(s_col[src * kRotThreads + tid], s_taps[k]));
        if (++src == a.R) src = 0;
      }
      a.out[(size_t)i * HW + p] = acc;
    }
  } else if (a.mode == 0) {
    for (int i = 0; i < a.R; ++i) a.out[(size_t)i * HW + p] = s_col[i * kRotThreads + tid];
  } else {
    for (int i = 0; i < a.R; ++i) a.out[(size_t)i * HW + p] = 0.0f;
  }
}

// ---- message stage 2a: resample into the eigen-frame of the covariance ----------------------------------

// Builds, for every eigen-frame cell, the (<=2) image cells that the reference's TM_DIRECT forward scatter
// (multi_array_transform.hpp:167-192) maps onto it, ordered by scatter order (x1 outer, y1 inner): .x is the
// LAST writer, .y the one before it (or -1).  The gather "last non-zero writer wins" then reproduces the
// scatter without a race.  *overflow is set if a cell has more than two pre-images.
__global__ void k_build_direct_map(int2 *__restrict__ map, int EH, int EW, int H, int W, const double *__restrict__ T31,
                                   const double *__restrict__ T13, int *overflow) {
  int ix = blockIdx.x * blockDim.x + threadIdx.x;
  int iy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ix >= EW || iy >= EH) return;
  // approximate pre-image (only used to centre the 3x3 search window)
  double sx = T13[0] * ix + T13[1] * iy + T13[2];
  double sy = T13[3] * ix + T13[4] * iy + T13[5];
  int cx = (int)floor(sx + 0.5), cy = (int)floor(sy + 0.5);
  int best1 = -1, best2 = -1;
  long long key1 = -1, key2 = -1;
  int count = 0;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      int x1 = cx + dx, y1 = cy + dy;
      if (x1 < 0 || x1 >= W || y1 < 0 || y1 >= H) continue;
      // hc::map_point(T31, x1, y1): M00*x + M01*y + M02, separately rounded
      double x3 = __dadd_rn(__dadd_rn(__dmul_rn(T31[0], (double)x1), __dmul_rn(T31[1], (double)y1)), T31[2]);
      double y3 = __dadd_rn(__dadd_rn(__dmul_rn(T31[3], (double)x1), __dmul_rn(T31[4], (double)y1)), T31[5]);
      int jx = (int)floor(__dadd_rn(x3, 0.5)), jy = (int)floor(__dadd_rn(y3, 0.5));
      if (jx != ix || jy != iy) continue;
      ++count;
      long long key = (long long)x1 * H + y1;
      int idx = y1 * W + x1;
      if (key > key1) {
        key2 = key1; best2 = best1;
        key1 = key; best1 = idx;
      } else if (key > key2) {
        key2 = key; best2 = idx;
      }
    }
  if (count > 2) atomicExch(overflow, 1);
  map[(size_t)iy * EW + ix] = make_int2(best1, best2);
}

// TM_DIRECT as a gather through the map; all R slices per thread (the map is shared by every slice).
__global__ void k_warp_direct(const float *__restrict__ in, float *__restrict__ out, const int2 *__restrict__ map,
                              int R, size_t HW, int EH, int EW, int EP) {
  int ix = blockIdx.x * blockDim.x + threadIdx.x;
  int iy = blockIdx.y;
  if (ix >= EP) return;
  size_t eplane = (size_t)EH * EP;
  size_t o = (size_t)iy * EP + ix;
  if (ix >= EW) {
    for (int r = 0; r < R; ++r) out[r * eplane + o] = 0.0f;
    return;
  }
  int2 m = map[(size_t)iy * EW + ix];
  if (m.x < 0) {
    for (int r = 0; r < R; ++r) out[r * eplane + o] = 0.0f;
    return;
  }
  constexpr int UB = 8;  // slices in flight per thread
  for (int r0 = 0; r0 < R; r0 += UB) {
    float v[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u)
      if (r0 + u < R) v[u] = __ldg(in + (size_t)(r0 + u) * HW + m.x);
    if (m.y >= 0) {
#pragma unroll
      for (int u = 0; u < UB; ++u)
        if (r0 + u < R && v[u] == 0.0f) v[u] = __ldg(in + (size_t)(r0 + u) * HW + m.y);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u)
      if (r0 + u < R) out[(r0 + u) * eplane + o] = v[u];
  }
}

// One TM_BILINEAR sample (multi_array_transform.hpp:196-238), default value 0.
__device__ __forceinline__ float bilinear_at(const float *__restrict__ s, int h, int w, int pitch, double x1,
                                             double y1) {
  double fx = floor(x1), fy = floor(y1);
  int ix = (int)fx, iy = (int)fy;
  if (ix < 0 || ix >= w || iy < 0 || iy >= h) return 0.0f;
  float a = (float)__dsub_rn(x1, (double)ix);
  float b = (float)__dsub_rn(y1, (double)iy);
  const float eps10 = 10 * 1.1920928955078125e-07f;
  const float *p = s + (size_t)iy * pitch + ix;
  if (a < eps10 && b < eps10) return __ldg(p);
  if (ix < w - 1 && iy < h - 1) {
    float omb = __fsub_rn(1.0f, b), oma = __fsub_rn(1.0f, a);
    float t0 = __fmul_rn(__fmul_rn(omb, oma), __ldg(p));
    float t1 = __fmul_rn(__fmul_rn(omb, a), __ldg(p + 1));
    float t2 = __fmul_rn(__fmul_rn(b, oma), __ldg(p + pitch));
    float t3 = __fmul_rn(__fmul_rn(b, a), __ldg(p + pitch + 1));
    return __fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3);
  }
  return 0.0f;
}

struct Affine {
  double m[6];
};
__device__ __forceinline__ void affine_map(const Affine &T, double x, double y, double &ox, double &oy) {
  ox = __dadd_rn(__dadd_rn(__dmul_rn(T.m[0], x), __dmul_rn(T.m[1], y)), T.m[2]);
  oy = __dadd_rn(__dadd_rn(__dmul_rn(T.m[3], x), __dmul_rn(T.m[4], y)), T.m[5]);
}

// TM_BILINEAR into the eigen-frame (non-sparse messages), all R slices per thread.
__global__ void k_warp_bilinear(const float *__restrict__ in, float *__restrict__ out, Affine T13, int R, int H, int W,
                                int EH, int EW, int EP) {
  int ix = blockIdx.x * blockDim.x + threadIdx.x;
  int iy = blockIdx.y;
  if (ix >= EP) return;
  size_t eplane = (size_t)EH * EP, HW = (size_t)H * W;
  size_t o = (size_t)iy * EP + ix;
  if (ix >= EW) {
    for (int r = 0; r < R; ++r) out[r * eplane + o] = 0.0f;
    return;
  }
  double x1, y1;
  affine_map(T13, (double)ix, (double)iy, x1, y1);
  // the sample position is shared by all R slices: resolve the case and the weights once (transform.hpp:196-238)
  const double fx = floor(x1), fy = floor(y1);
  const int sx = (int)fx, sy = (int)fy;
  int mode = 0;  // 0: default (0), 1: exact hit, 2: four-point blend
  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
  if (sx >= 0 && sx < W && sy >= 0 && sy < H) {
    const float fa = (float)__dsub_rn(x1, (double)sx), fb = (float)__dsub_rn(y1, (double)sy);
    const float eps10 = 10 * 1.1920928955078125e-07f;
    if (fa < eps10 && fb < eps10) mode = 1;
    else if (sx < W - 1 && sy < H - 1) {
      mode = 2;
      const float omb = __fsub_rn(1.0f, fb), oma = __fsub_rn(1.0f, fa);
      w00 = __fmul_rn(omb, oma); w01 = __fmul_rn(omb, fa); w10 = __fmul_rn(fb, oma); w11 = __fmul_rn(fb, fa);
    }
  }
  const float *p = in + (size_t)sy * W + sx;
  if (mode == 0) {
    for (int r = 0; r < R; ++r) out[r * eplane + o] = 0.0f;
  } else if (mode == 1) {
    for (int r = 0; r < R; ++r) out[r * eplane + o] = __ldg(p + r * HW);
  } else {
    constexpr int UB = 4;
    for (int r0 = 0; r0 < R; r0 += UB) {
      float q[UB][4];
#pragma unroll
      for (int u = 0; u < UB; ++u)
        if (r0 + u < R) {
          const float *s = p + (size_t)(r0 + u) * HW;
          q[u][0] = __ldg(s); q[u][1] = __ldg(s + 1); q[u][2] = __ldg(s + W); q[u][3] = __ldg(s + W + 1);
        }
#pragma unroll
      for (int u = 0; u < UB; ++u)
        if (r0 + u < R) {
          float t0 = __fmul_rn(w00, q[u][0]), t1 = __fmul_rn(w01, q[u][1]);
          float t2 = __fmul_rn(w10, q[u][2]), t3 = __fmul_rn(w11, q[u][3]);
          out[(r0 + u) * eplane + o] = __fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3);
        }
    }
  }
}


// ---- resampling v2: one thread = one destination cell x a group of RG slices ----------------------------------
// The source position of a cell is the same for every rotation slice, so its fp64 coordinate math, case analysis and
// fp32 weights are done once per thread and reused for RG slices; RG < R keeps enough threads in flight (v1 looped
// over all R slices per thread and was latency-bound).

// Bilinear sample set-up shared by the slices (multi_array_transform.hpp:196-238, default value 0).
struct BilinearTap {
  int mode;      // 0: default, 1: exact hit, 2: four-point blend
  int off;       // iy * pitch + ix
  float w00, w01, w10, w11;
};
__device__ __forceinline__ BilinearTap bilinear_setup(int h, int w, int pitch, double x1, double y1) {
  BilinearTap t;
  t.mode = 0; t.off = 0; t.w00 = t.w01 = t.w10 = t.w11 = 0.0f;
  const double fx = floor(x1), fy = floor(y1);
  const int ix = (int)fx, iy = (int)fy;
  if (ix >= 0 && ix < w && iy >= 0 && iy < h) {
    const float a = (float)__dsub_rn(x1, fx), b = (float)__dsub_rn(y1, fy);
    const float eps10 = 10 * 1.1920928955078125e-07f;
    t.off = iy * pitch + ix;
    if (a < eps10 && b < eps10) t.mode = 1;
    else if (ix < w - 1 && iy < h - 1) {
      t.mode = 2;
      const float omb = __fsub_rn(1.0f, b), oma = __fsub_rn(1.0f, a);
      t.w00 = __fmul_rn(omb, oma); t.w01 = __fmul_rn(omb, a); t.w10 = __fmul_rn(b, oma); t.w11 = __fmul_rn(b, a);
    }
  }
  return t;
}

// dst[r][cell] = bilinear(src[r]) for r in the thread's slice group.  Used both ways: image -> eigen-frame
// (TM_BILINEAR of gaussFilter2dOffset, filter.hpp:359) and eigen-frame -> image (filter.hpp:367-368).
// tr != 0: the destination is stored transposed (dst[r][ix][iy], pitch dpitch over iy) and threadIdx.x walks iy.
// FULL: R is a multiple of RG, so no slice of the group needs a guard (ncu r01h: with the guards a warp issued 458
// instructions per 32 cells x 8 slices, 91 of them predicate logic and 126 address IMAD/LEA behind the predicates,
// against 102 loads, multiplies, adds and stores).
template <int RG, bool FULL>
__device__ __forceinline__ void resample_bilinear_block(const float *__restrict__ src, float *__restrict__ dst, const Affine &T,
                                                        int R, int sh, int sw, int spitch, size_t splane, int dh,
                                                        int dw, int dpitch, size_t dplane, int tr, int bx, int by, int bz) {
  // 2-D tiles: a tile's rotated footprint in the source stays compact, so the four taps of neighbouring cells
  // hit the same L1 lines
  // a warp owns an 8 x 4 patch of the 16 x 16 tile (8 along the contiguous destination axis): its rotated footprint in
  // the source spans ~7.6 rows on average over angles against ~11.5 for a 16 x 2 strip -- a third fewer L1 sectors
  // per gather request (ncu r01d: 11 sectors, 3.7 wavefronts per request with the strip)
  const int tid = threadIdx.y * 16 + threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int tu = (wrp & 1) * 8 + (lane & 7), tv = (wrp >> 1) * 4 + (lane >> 3);  // u: contiguous axis of dst
  const int ix = bx * 16 + (tr ? tv : tu);
  const int iy = by * 16 + (tr ? tu : tv);
  const int r0 = bz * RG;
  if (tr ? (ix >= dw || iy >= dh) : (ix >= dpitch || iy >= dh)) return;
  float *o = dst + (size_t)r0 * dplane + (tr ? (size_t)ix * dpitch + iy : (size_t)iy * dpitch + ix);
  BilinearTap t;
  t.mode = 0;
  if (ix < dw) {
    double x1, y1;
    affine_map(T, (double)ix, (double)iy, x1, y1);
    t = bilinear_setup(sh, sw, spitch, x1, y1);
  }
  const float *p = src + (size_t)r0 * splane + t.off;
  const int nu = FULL ? RG : min(RG, R - r0);
  if (t.mode == 2) {
    float q[RG][4];
    const float *s = p;
#pragma unroll
    for (int u = 0; u < RG; ++u, s += splane)
      if (FULL || u < nu) {
        const float *s1 = s + spitch;
        q[u][0] = __ldg(s); q[u][1] = __ldg(s + 1); q[u][2] = __ldg(s1); q[u][3] = __ldg(s1 + 1);
      }
#pragma unroll
    for (int u = 0; u < RG; ++u)
      if (FULL || u < nu) {
        float t0 = __fmul_rn(t.w00, q[u][0]), t1 = __fmul_rn(t.w01, q[u][1]);
        float t2 = __fmul_rn(t.w10, q[u][2]), t3 = __fmul_rn(t.w11, q[u][3]);
        o[(size_t)u * dplane] = __fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3);
      }
  } else if (t.mode == 1) {
#pragma unroll
    for (int u = 0; u < RG; ++u)
      if (FULL || u < nu) o[(size_t)u * dplane] = __ldg(p + (size_t)u * splane);
  } else {
#pragma unroll
    for (int u = 0; u < RG; ++u)
      if (FULL || u < nu) o[(size_t)u * dplane] = 0.0f;
  }
}


template <int RG, bool FULL>
__global__ void __launch_bounds__(256) k_resample_bilinear(const float *__restrict__ src, float *__restrict__ dst, Affine T,
                                                           int R, int sh, int sw, int spitch, size_t splane, int dh,
                                                           int dw, int dpitch, size_t dplane, int tr) {
  resample_bilinear_block<RG, FULL>(src, dst, T, R, sh, sw, spitch, splane, dh, dw, dpitch, dplane, tr, blockIdx.x, blockIdx.y,
                                    blockIdx.z);
}
// Several messages of one tree level per launch: blockIdx.z = message * (R / RG) + slice group; the grid spans the
// largest destination of the batch, blocks outside their own message's grid leave at once.
struct ResampleMsg {
  const float *src;
  float *dst;
  Affine T;
  size_t splane, dplane;
  int sh, sw, spitch, dh, dw, dpitch;
};
struct ResampleBatch {
  ResampleMsg m[kMaxBatch];
  int R, tr, zgroups;
};
template <int RG>
__global__ void __launch_bounds__(256) k_resample_bilinear_b(const __grid_constant__ ResampleBatch rb) {
  const int mi = blockIdx.z / rb.zgroups, zg = blockIdx.z - mi * rb.zgroups;
  const ResampleMsg &m = rb.m[mi];
  if ((int)blockIdx.x * 16 >= (rb.tr ? m.dw : m.dpitch) || (int)blockIdx.y * 16 >= m.dh) return;
  resample_bilinear_block<RG, true>(m.src, m.dst, m.T, rb.R, m.sh, m.sw, m.spitch, m.splane, m.dh, m.dw, m.dpitch, m.dplane,
                                    rb.tr, blockIdx.x, blockIdx.y, zg);
}

// (Round 1 also tried staging each tile's 28x28 source bounding box in shared memory -- the direct gathers are bound
// by the L1 data pipe at ~11 sectors per request -- but the staging loop plus 3x over-fetch made it 2x slower.)
// TM_DIRECT as a gather through the winner map, RG slices per thread.
// tr != 0: out is stored transposed ([r][ix][iy], pitch EP over iy, plane EW*EP) and threadIdx.x walks iy.
template <int RG, bool FULL>
__device__ __forceinline__ void warp_direct_block(const float *__restrict__ in, float *__restrict__ out,
                                                  const int2 *__restrict__ map, int R, size_t HW, int EH, int EW,
                                                  int EP, int tr, int bx, int by, int bz) {
  const int tid = threadIdx.y * 16 + threadIdx.x, lane = tid & 31, wrp = tid >> 5;  // 8 x 4 warp patches, see above
  const int tu = (wrp & 1) * 8 + (lane & 7), tv = (wrp >> 1) * 4 + (lane >> 3);
  const int ix = bx * 16 + (tr ? tv : tu);
  const int iy = by * 16 + (tr ? tu : tv);
  const int r0 = bz * RG;
  if (tr ? (ix >= EW || iy >= EH) : (ix >= EP || iy >= EH)) return;
  const size_t eplane = tr ? (size_t)EW * EP : (size_t)EH * EP;
  float *o = out + (size_t)r0 * eplane + (tr ? (size_t)ix * EP + iy : (size_t)iy * EP + ix);
  int2 m = make_int2(-1, -1);
  if (ix < EW) m = map[(size_t)iy * EW + ix];
  const int nu = FULL ? RG : min(RG, R - r0);  // FULL: R % RG == 0, no slice of the group needs a guard
  const float *base = in + (size_t)r0 * HW;
  float v[RG];
#pragma unroll
  for (int u = 0; u < RG; ++u) v[u] = 0.0f;
  if (m.x >= 0) {
    const float *s = base + m.x;
#pragma unroll
    for (int u = 0; u < RG; ++u, s += HW)
      if (FULL || u < nu) v[u] = __ldg(s);
  }
  if (m.y >= 0) {
    const float *s = base + m.y;
#pragma unroll
    for (int u = 0; u < RG; ++u, s += HW)
      if ((FULL || u < nu) && v[u] == 0.0f) v[u] = __ldg(s);
  }
#pragma unroll
  for (int u = 0; u < RG; ++u)
    if (FULL || u < nu) o[(size_t)u * eplane] = v[u];
}

template <int RG, bool FULL>
__global__ void __launch_bounds__(256) k_warp_direct2(const float *__restrict__ in, float *__restrict__ out,
                                                      const int2 *__restrict__ map, int R, size_t HW, int EH, int EW,
                                                      int EP, int tr) {
  warp_direct_block<RG, FULL>(in, out, map, R, HW, EH, EW, EP, tr, blockIdx.x, blockIdx.y, blockIdx.z);
}
struct DirectMsg {
  const float *in;
  float *out;
  const int2 *map;
  int EH, EW, EP;
};
struct DirectBatch {
  DirectMsg m[kMaxBatch];
  size_t HW;
  int R, tr, zgroups;
};
template <int RG>
__global__ void __launch_bounds__(256) k_warp_direct_b(const __grid_constant__ DirectBatch db) {
  const int mi = blockIdx.z / db.zgroups, zg = blockIdx.z - mi * db.zgroups;
  const DirectMsg &m = db.m[mi];
  if ((int)blockIdx.x * 16 >= (db.tr ? m.EW : m.EP) || (int)blockIdx.y * 16 >= m.EH) return;
  warp_direct_block<RG, true>(m.in, m.out, m.map, db.R, db.HW, m.EH, m.EW, m.EP, db.tr, blockIdx.x, blockIdx.y, zg);
}

// ---- message stage 2b: separable Gaussian, zero padded, unnormalised taps -------------------------------
// gaussFilterDiag2d (multi_array_filter.hpp:212-321).  Both kernels give each thread T consecutive outputs
// along the filtered axis and slide a T-wide register window of the data: per tap one new datum is loaded and
// T multiply-add pairs retire, in ascending tap order for every output.

struct ConvArgs {
  const float *in;
  float *out;
  const float *taps;
  int len;        // odd
  int rows, cols; // slice size (valid region)
  int pitch;      // row pitch in floats
  size_t plane;   // slice stride in floats
};

// Filter along y (column direction).  blockDim.x threads = adjacent columns (coalesced rows).
template <int T>
__global__ void __launch_bounds__(128) k_conv_cols(ConvArgs a) {
  __shared__ float s_taps[1000];
  for (int i = threadIdx.x; i < a.len; i += blockDim.x) s_taps[i] = a.taps[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= a.cols) return;
  const int y0 = blockIdx.y * T;
  const float *src = a.in + (size_t)blockIdx.z * a.plane + x;
  float *dst = a.out + (size_t)blockIdx.z * a.plane + x;
  const int n = (a.len - 1) / 2;

  float acc[T], d[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = 0.0f;
  // window rows y0-n .. y0-n+T-2 in d[0..T-2]
#pragma unroll
  for (int j = 0; j < T - 1; ++j) {
    int yy = y0 - n + j;
    d[j] = (yy >= 0 && yy < a.rows) ? __ldg(src + (size_t)yy * a.pitch) : 0.0f;
  }
  for (int kk = 0; kk < a.len; kk += T) {
#pragma unroll
    for (int u = 0; u < T; ++u) {
      int k = kk + u;
      if (k < a.len) {
        int yy = y0 - n + k + (T - 1);
        d[(u + T - 1) % T] = (yy >= 0 && yy < a.rows) ? __ldg(src + (size_t)yy * a.pitch) : 0.0f;
        float f = s_taps[k];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = __fadd_rn(acc[t], __fmul_rn(d[(u + t) % T], f));
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t)
    if (y0 + t < a.rows) dst[(size_t)(y0 + t) * a.pitch] = acc[t];
}

// Filter along x (row direction).  A block stages TY rows (+halo) in shared memory in a "transposed by T"
// layout (element i of a row lives at (i % T) * S + i / T) so that threads owning adjacent T-groups read
// adjacent words.
template <int T>
__global__ void __launch_bounds__(256) k_conv_rows(ConvArgs a, int TY, int S) {
  extern __shared__ float s_tile[];  // [TY][T*S]
  __shared__ float s_taps[1000];
  const int tid = threadIdx.x;
  for (int i = tid; i < a.len; i += blockDim.x) s_taps[i] = a.taps[i];
  const int n = (a.len - 1) / 2;
  const int y0 = blockIdx.x * TY;
  const float *src = a.in + (size_t)blockIdx.y * a.plane;
  float *dst = a.out + (size_t)blockIdx.y * a.plane;
  const int G = (a.cols + T - 1) / T;   // T-groups per row
  const int span = G * T + 2 * n;       // staged elements per row: x in [-n, G*T + n)
  const int rowsz = T * S;
  for (int ry = 0; ry < TY; ++ry) {
    int y = y0 + ry;
    const float *row = src + (size_t)y * a.pitch;
    for (int i = tid; i < span; i += blockDim.x) {
      int x = i - n;
      float v = (y < a.rows && x >= 0 && x < a.cols) ? __ldg(row + x) : 0.0f;
      s_tile[ry * rowsz + (i % T) * S + i / T] = v;
    }
  }
  __syncthreads();
  const int items = TY * G;
  for (int it = tid; it < items; it += blockDim.x) {
    const int ry = it / G, g = it % G;
    const int y = y0 + ry;
    if (y >= a.rows) continue;
    const float *tile = s_tile + ry * rowsz + g;
    float acc[T], d[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0f;
#pragma unroll
    for (int j = 0; j < T - 1; ++j) d[j] = tile[j * S];  // element g*T + j  ->  (j % T)*S + g
    for (int kk = 0; kk < a.len; kk += T) {
#pragma unroll
      for (int u = 0; u < T; ++u) {
        int k = kk + u;
        if (k < a.len) {
          // element g*T + k + T-1 = g*T + kk + (u+T-1)  ->  ((u+T-1)%T)*S + g + (kk + u + T-1)/T
          d[(u + T - 1) % T] = tile[((u + T - 1) % T) * S + (kk + u + T - 1) / T];
          float f = s_taps[k];
#pragma unroll
          for (int t = 0; t < T; ++t) acc[t] = __fadd_rn(acc[t], __fmul_rn(d[(u + t) % T], f));
        }
      }
    }
    float *o = dst + (size_t)y * a.pitch + g * T;
#pragma unroll
    for (int t = 0; t < T; ++t)
      if (g * T + t < a.cols) o[t] = acc[t];
  }
}


// =====================================================================================================
// v2 kernels: packed fp32x2 arithmetic.
//
// Measured on B200 (profiles/r01_fp32_pipe_microbench.txt): FFMA2/FADD2 retire two lanes per instruction but
// occupy the fp32 pipe for two cycles, so they do not raise the arithmetic ceiling (18.0e12 separately rounded
// multiply+add pairs per second) -- they halve the ISSUE slots per tap-output, which leaves the other half of
// the issue bandwidth for loads and address arithmetic.  The scalar kernels above were issue-bound.
//
// Parity: a tap-output is still acc = RN(acc + RN(x*f)).  ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// even with --fmad=false, so the multiply is spelled fma.rn.f32x2(x, f, -0.0) with the -0.0 pair passed as a
// kernel argument (x*f + -0 == RN(x*f) for every x*f, including +-0, inf, NaN; ptxas cannot fold a run-time
// addend), followed by add.rn.f32x2.  SASS: FFMA2 + FADD2 (checked by tests/test_build_sass.py).
// =====================================================================================================
typedef unsigned long long u64;
#define PS_NEGZERO2 0x8000000080000000ull

__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// RN(x.lo*f), RN(x.hi*f)
__device__ __forceinline__ u64 mul2_rn(u64 x, float f, u64 negzero2) {
  u64 r, ff = pk2(f, f);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(ff), "l"(negzero2));
  return r;
}
__device__ __forceinline__ u64 add2_rn(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// One filter tap on a packed pair.  FMA = false: the parity arithmetic (product and sum rounded separately, like the
// reference's sdot).  FMA = true (ps_config.fast_math): one fused multiply-add -- one rounding fewer per tap and half
// the fp32-pipe time; marginals then agree with the reference to ~1e-6 relative instead of bit for bit.
template <bool FMA>
__device__ __forceinline__ u64 tap2(u64 acc, u64 x, float f, u64 negzero2) {
  if (FMA) {
    u64 r, ff = pk2(f, f);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(ff), "l"(acc));
    return r;
  }
  return add2_rn(acc, mul2_rn(x, f, negzero2));
}

// ---- stage 1 v2: shift + exp + circular rotation filter, register resident ------------------------------------
// One thread owns two consecutive flat pixels; its 2 x R exponentiated values live in registers as R packed pairs.
// L is the compile-time (zero-padded, centred) tap count: padding taps are exactly 0, so they add +0.
template <int R, int L>
__global__ void __launch_bounds__(128) k_rotconv2(RotArgs a, u64 nz) {
  __shared__ float s_taps[L];
  const int tid = threadIdx.x;
  const int npad = (L - 1) / 2, n = (a.len - 1) / 2;
  if (tid < L) {
    int k = tid - (npad - n);
    s_taps[tid] = (a.mode == 1 && k >= 0 && k < a.len) ? a.taps[k] : 0.0f;
  }
  __syncthreads();
  const size_t HW = (size_t)a.H * a.W;
  const size_t p0 = ((size_t)blockIdx.x * blockDim.x + tid) * 2;
  if (p0 >= HW) return;
  const bool has1 = p0 + 1 < HW;
  const int y0 = (int)(p0 / a.W), x0 = (int)(p0 % a.W);
  int y1 = y0, x1 = x0 + 1;
  if (x1 == a.W) { x1 = 0; y1 = y0 + 1; }
  const float negM = -dec_f(*a.max_enc);

  u64 v[R];
#pragma unroll
  for (int ro = 0; ro < R; ++ro) {
    int r = ro - a.shift;
    float e0 = kLogZero, e1 = kLogZero;
    if (r >= 0 && r < R) {
      const float *s = a.in + (size_t)r * HW;
      int ys = a.yin[r * a.H + y0], xs = a.xin[r * a.W + x0];
      if ((ys | xs) >= 0) e0 = __ldg(&s[(size_t)ys * a.W + xs]);
      if (has1) {
        ys = a.yin[r * a.H + y1]; xs = a.xin[r * a.W + x1];
        if ((ys | xs) >= 0) e1 = __ldg(&s[(size_t)ys * a.W + xs]);
      }
    }
    v[ro] = pk2(exp_f64(__fadd_rn(e0, negM)), exp_f64(__fadd_rn(e1, negM)));
  }
  const bool vec = (HW & 1) == 0;
  if (a.mode == 1) {
    float tp[L];
#pragma unroll
    for (int k = 0; k < L; ++k) tp[k] = s_taps[k];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      u64 acc = pk2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const int src = ((i + k - npad) % R + R) % R;
        acc = add2_rn(acc, mul2_rn(v[src], tp[k], nz));
      }
      float lo, hi;
      upk2(acc, lo, hi);
      float *o = a.out + (size_t)i * HW + p0;
      if (vec) *reinterpret_cast<float2 *>(o) = make_float2(lo, hi);
      else { o[0] = lo; if (has1) o[1] = hi; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float lo = 0.0f, hi = 0.0f;
      if (a.mode == 0) upk2(v[i], lo, hi);
      float *o = a.out + (size_t)i * HW + p0;
      if (vec) *reinterpret_cast<float2 *>(o) = make_float2(lo, hi);
      else { o[0] = lo; if (has1) o[1] = hi; }
    }
  }
}


// ---- stage 1 v3: block-cooperative shift + exp, then rotation filter from shared memory ---------------------------
// k_rotconv2 gave every thread 2 pixels x R rotations of serial work (gather, fp64 exp, filter): only ~25 warps per
// SM existed and the kernel was latency-bound.  Here a block owns PX consecutive pixels: phase 1 spreads the R*PX
// gather+exp cells over all threads (coalesced along pixels); phase 2 gives each thread one pixel PAIR and a run of
// OUT consecutive output rotations, sliding a register window over the (circular) rotation axis.
template <int R, int L, int PX, int OUT>
__global__ void __launch_bounds__(256) k_rotconv3(RotArgs a, u64 nz) {
  static_assert(R % OUT == 0 && PX % 2 == 0, "tiling");
  __shared__ __align__(8) float s_e[R][PX];
  __shared__ float s_taps[L];
  const int tid = threadIdx.x;
  const int npad = (L - 1) / 2, n = (a.len - 1) / 2;
  if (tid < L) {
    int k = tid - (npad - n);
    s_taps[tid] = (a.mode == 1 && k >= 0 && k < a.len) ? a.taps[k] : 0.0f;
  }
  const size_t HW = (size_t)a.H * a.W;
  const size_t base = (size_t)blockIdx.x * PX;
  const float negM = -dec_f(*a.max_enc);
  // phase 1: A[ro][px] = exp(shifted child - M).  256 % PX == 0, so a thread keeps one pixel and walks rotations.
  static_assert(256 % PX == 0, "a thread must stay on one pixel");
  {
    const int px = tid % PX;
    const unsigned p = (unsigned)base + px;
    const bool okp = p < (unsigned)HW;
    const int y = okp ? (int)(p / (unsigned)a.W) : 0, x = okp ? (int)(p % (unsigned)a.W) : 0;
    constexpr int STEP = 256 / PX, NIT = (R + STEP - 1) / STEP;
    // Three batched rounds (tables, data, exp) instead of NIT dependent load chains.
    int off[NIT];
#pragma unroll
    for (int c = 0; c < NIT; ++c) {
      const int ro = tid / PX + c * STEP;
      const int r = ro - a.shift;
      off[c] = -1;
      if (okp && ro < R && r >= 0 && r < R) {
        if (a.shift_xy) {
          const int xs = x + a.shift_xy[2 * r], ys = y + a.shift_xy[2 * r + 1];
          if ((unsigned)xs < (unsigned)a.W && (unsigned)ys < (unsigned)a.H) off[c] = ys * a.W + xs;
        } else {
          const int ys = a.yin[r * a.H + y], xs = a.xin[r * a.W + x];
          if ((ys | xs) >= 0) off[c] = ys * a.W + xs;
        }
      }
    }
    float v[NIT];
#pragma unroll
    for (int c = 0; c < NIT; ++c) {
      const int ro = tid / PX + c * STEP;
      const int r = ro - a.shift;
      v[c] = off[c] >= 0 ? __ldg(&a.in[(size_t)r * HW + (unsigned)off[c]]) : kLogZero;
    }
#pragma unroll
    for (int c = 0; c < NIT; ++c) {
      const int ro = tid / PX + c * STEP;
      if (ro < R) s_e[ro][px] = okp ? exp_f64(__fadd_rn(v[c], negM)) : 0.0f;
    }
  }
  __syncthreads();
  // phase 2
  constexpr int NPAIR = PX / 2, NRUN = R / OUT;
  const bool vec = (HW & 1) == 0;
  for (int it = tid; it < NPAIR * NRUN; it += 256) {
    const int pp = it % NPAIR, run = it / NPAIR;
    const size_t p0 = base + 2 * pp;
    if (p0 >= HW) continue;
    const bool has1 = p0 + 1 < HW;
    const int i0 = run * OUT;
    const u64 *col = reinterpret_cast<const u64 *>(&s_e[0][0]) + pp;  // row stride PX/2 in u64 units
    float lo[OUT], hi[OUT];
    if (a.mode == 1) {
      // inputs (i0 - npad + m) mod R for m in [0, OUT + L - 1)
      u64 w[OUT + L - 1];
#pragma unroll
      for (int m = 0; m < OUT + L - 1; ++m) {
        int src = (i0 - npad + m) % R;
        if (src < 0) src += R;
        w[m] = col[src * (PX / 2)];
      }
      u64 acc[OUT];
#pragma unroll
      for (int o = 0; o < OUT; ++o) acc[o] = pk2(0.0f, 0.0f);
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const float f = s_taps[k];
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[o] = add2_rn(acc[o], mul2_rn(w[o + k], f, nz));
      }
#pragma unroll
      for (int o = 0; o < OUT; ++o) upk2(acc[o], lo[o], hi[o]);
    } else {
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        lo[o] = hi[o] = 0.0f;
        if (a.mode == 0) upk2(col[(i0 + o) * (PX / 2)], lo[o], hi[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      float *dst = a.out + (size_t)(i0 + o) * HW + p0;
      if (vec) *reinterpret_cast<float2 *>(dst) = make_float2(lo[o], hi[o]);
      else { dst[0] = lo[o]; if (has1) dst[1] = hi[o]; }
    }
  }
}

// ---- stage 1 v4: k_rotconv3 on an instruction diet (pure-shift tables only) ---------------------------------------
// ncu r01d: only 16 % of k_rotconv3's instructions were the FFMA2/FADD2 pairs; the rest was index arithmetic --
// a modulo per window element in phase 2, two table loads, a 64-bit address and three branches per cell in phase 1.
// Here the exp tile is stored circularly EXTENDED (row i = A[(i - npad) mod R], i < R + L - 1), so a thread's window
// is OUT + L - 1 loads at compile-time offsets from one address; the per-rotation shift records live in shared
// memory; a pixel's (x, y) comes from one FastDiv; exp runs branch-free behind one warp vote per cell.
// The filter of k_rotconv4 over the centred LE of the L staged taps (LE, L odd, LE <= L).
template <int LE, int L, int OUT, int NPAIR, bool FMA>
__device__ __forceinline__ void rot_filter(const u64 *col, const float *s_taps, u64 nz, float (&lo)[OUT], float (&hi)[OUT]) {
  constexpr int K0 = (L - (LE < L ? LE : L)) / 2, K1 = K0 + (LE < L ? LE : L);
  u64 acc[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) acc[o] = pk2(0.0f, 0.0f);
#pragma unroll
  for (int k = K0; k < K1; ++k) {
    const float f = s_taps[k];
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc[o] = tap2<FMA>(acc[o], col[(o + k) * NPAIR], f, nz);
  }
#pragma unroll
  for (int o = 0; o < OUT; ++o) upk2(acc[o], lo[o], hi[o]);
}
template <int R, int L, int PX, int OUT, bool FMA>
__global__ void __launch_bounds__(256) k_rotconv4(const __grid_constant__ RotBatch rb, FastDiv Wdiv, u64 nz) {
  static_assert(R % OUT == 0 && PX % 2 == 0 && 256 % PX == 0 && PX >= 32, "tiling");
  const RotArgs &a = rb.a[blockIdx.y];
  constexpr int NPAD = (L - 1) / 2, RX = R + L - 1;
  __shared__ __align__(8) float s_e[RX][PX];
  __shared__ float s_taps[L];
  __shared__ int4 s_sh[R];  // per OUTPUT rotation: (dx, dy, r*HW + dy*W + dx, source rotation or -1)
  const int tid = threadIdx.x;
  const int HW = a.H * a.W;
  const int n = (a.len - 1) / 2;
  if (tid < L) {
    const int k = tid - (NPAD - n);
    s_taps[tid] = (a.mode == 1 && k >= 0 && k < a.len) ? a.taps[k] : 0.0f;
  }
  if (tid < R) {
    const int r = tid - a.shift;
    int4 q = make_int4(0, 0, 0, -1);
    if (r >= 0 && r < R) {
      const int dx = a.shift_xy[2 * r], dy = a.shift_xy[2 * r + 1];
      q = make_int4(dx, dy, r * HW + dy * a.W + dx, r);
    }
    s_sh[tid] = q;
  }
  const float negM = -dec_f(*a.max_enc);
  const int base = blockIdx.x * PX;
  __syncthreads();
  {
    const int px = tid % PX, ro0 = tid / PX;
    const int p = base + px;
    const bool okp = p < HW;
    const int y = (int)Wdiv.div((unsigned)p), x = p - y * a.W;
    constexpr int STEP = 256 / PX, NIT = (R + STEP - 1) / STEP;
    float v[NIT];
#pragma unroll
    for (int c = 0; c < NIT; ++c) {
      const int ro = ro0 + c * STEP;
      v[c] = kLogZero;
      if (ro < R) {
        const int4 q = s_sh[ro];
        const bool ok = okp & (q.w >= 0) & ((unsigned)(x + q.x) < (unsigned)a.W) & ((unsigned)(y + q.y) < (unsigned)a.H);
        if (ok) v[c] = __ldg(a.in + (p + q.z));
      }
    }
#pragma unroll
    for (int c = 0; c < NIT; ++c) {
      const int ro = ro0 + c * STEP;
      if (ro < R) {
        const float xv = __fadd_rn(v[c], negM);
        float e = 0.0f;
        if (__any_sync(0xffffffffu, !(xv < -104.0f))) {  // sparse unaries: most warps see only LOG_ZERO cells
#ifndef PS_SLOW_MATH
          bool redo;
          e = exp_fast_nb(xv, redo);
          if (redo) e = exp_slow_call(xv);
#else
          e = exp_slow(xv);
#endif
        }
        s_e[ro + NPAD][px] = e;
        if (ro >= R - NPAD) s_e[ro + NPAD - R][px] = e;
        if (ro < NPAD) s_e[ro + NPAD + R][px] = e;
      }
    }
  }
  __syncthreads();
  constexpr int NPAIR = PX / 2, NRUN = R / OUT;
  const bool vec = (HW & 1) == 0;
  for (int it = tid; it < NPAIR * NRUN; it += 256) {
    const int pp = it % NPAIR, run = it / NPAIR;
    const int p0 = base + 2 * pp;
    if (p0 >= HW) continue;
    const bool has1 = p0 + 1 < HW;
    const int i0 = run * OUT;
    const u64 *col = reinterpret_cast<const u64 *>(&s_e[0][0]) + (i0 * NPAIR + pp);
    float lo[OUT], hi[OUT];
    if (a.mode == 1) {
      // L is the longest filter of the launch.  With many rotations (R = 48: up to 47 taps) a message whose own filter
      // is shorter runs only the centred LE <= L taps that can be non-zero -- the skipped ones are exactly 0 and add +0
      // to a non-negative sum.  Block-uniform.  (At R = 24 the gather + exp phase bounds the kernel: the same dispatch moved it by 1.5 %, measured.)
      if (R > 24 && L > 15 && a.len <= 15) rot_filter<15, L, OUT, NPAIR, FMA>(col, s_taps, nz, lo, hi);
      else if (R > 24 && L > 23 && a.len <= 23) rot_filter<23, L, OUT, NPAIR, FMA>(col, s_taps, nz, lo, hi);
      else if (R > 24 && L > 31 && a.len <= 31) rot_filter<31, L, OUT, NPAIR, FMA>(col, s_taps, nz, lo, hi);
      else if (R > 24 && L > 39 && a.len <= 39) rot_filter<39, L, OUT, NPAIR, FMA>(col, s_taps, nz, lo, hi);
      else rot_filter<L, L, OUT, NPAIR, FMA>(col, s_taps, nz, lo, hi);
    } else {
#pragma unroll
      for (int o = 0; o < OUT; ++o) {
        lo[o] = hi[o] = 0.0f;
        if (a.mode == 0) upk2(col[(o + NPAD) * NPAIR], lo[o], hi[o]);
      }
    }
    float *dst = a.out + (i0 * HW + p0);
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
      if (vec) *reinterpret_cast<float2 *>(dst + o * HW) = make_float2(lo[o], hi[o]);
      else { dst[o * HW] = lo[o]; if (has1) dst[o * HW + 1] = hi[o]; }
    }
  }
}

// ---- stage 2b v2: Gaussian along y, two adjacent columns per thread, row tile staged in shared memory -----------
// Block = 8 warps; warp w produces rows [y0 + w*T, y0 + w*T + T) of a 64-column strip.  The strip's
// (8T + 2n) input rows are staged once (coalesced float2 rows), so each input element is read from L2
// (8T+2n)/(8T) times instead of (len+T-1)/T times.  Requires an even pitch (8-byte aligned pairs).
template <int T>
__global__ void __launch_bounds__(256, 4) k_conv_cols2(ConvArgs a, u64 nz) {
  extern __shared__ float2 s_rows[];  // [8T + 2n][32]
  __shared__ float s_taps[1000];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < a.len; i += 256) s_taps[i] = a.taps[i];
  const int n = (a.len - 1) / 2;
  const int x = (blockIdx.x * 32 + lane) * 2;
  const int y0 = blockIdx.y * (8 * T);
  const float *src = a.in + (size_t)blockIdx.z * a.plane;
  float *dst = a.out + (size_t)blockIdx.z * a.plane;
  const int nrows = 8 * T + 2 * n;
  const bool in0 = x < a.cols, in1 = x + 1 < a.cols;
  for (int rr = w; rr < nrows; rr += 8) {
    int yy = y0 - n + rr;
    float2 v = make_float2(0.0f, 0.0f);
    if (in0 && yy >= 0 && yy < a.rows) {
      v = __ldg(reinterpret_cast<const float2 *>(src + (size_t)yy * a.pitch + x));
      if (!in1) v.y = 0.0f;
    }
    s_rows[rr * 32 + lane] = v;
  }
  __syncthreads();
  const u64 *win = reinterpret_cast<const u64 *>(s_rows) + (w * T) * 32 + lane;
  u64 acc[T], d[T];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t] = pk2(0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < T - 1; ++j) d[j] = win[j * 32];
  int kk = 0;
  for (; kk + T <= a.len; kk += T) {
#pragma unroll
    for (int u = 0; u < T; ++u) {
      d[(u + T - 1) % T] = win[(kk + u + T - 1) * 32];
      const float f = s_taps[kk + u];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = add2_rn(acc[t], mul2_rn(d[(u + t) % T], f, nz));
    }
  }
#pragma unroll
  for (int u = 0; u < T; ++u) {
    if (kk + u < a.len) {
      d[(u + T - 1) % T] = win[(kk + u + T - 1) * 32];
      const float f = s_taps[kk + u];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = add2_rn(acc[t], mul2_rn(d[(u + t) % T], f, nz));
    }
  }
  if (!in0) return;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    int y = y0 + w * T + t;
    if (y < a.rows) {
      float lo, hi;
      upk2(acc[t], lo, hi);
      float *o = dst + (size_t)y * a.pitch + x;
      if (in1) *reinterpret_cast<float2 *>(o) = make_float2(lo, hi);
      else o[0] = lo;
    }
  }
}


// ---- stage 2b, TMA plumbing: mbarrier and bulk-tensor wrappers used by k_conv_cols_tma2 below -------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2,
                                            unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

struct ColsTmaArgs {
  float *out;
  const float *taps;
  int len, rows, cols, pitch;  // pitch/plane describe the OUTPUT; the input layout lives in the tensor map
  size_t plane;
  int slices, ytiles, xtiles;  // tile grid: ytiles of 8T rows, xtiles of 64 columns
  int transpose_out;           // 1: out[z][col][row] (the input was stored transposed; this pass restores the layout)
  const int *tile_list;        // optional (first row, strip | groups << 12) pairs, same for every slice; null = all tiles
  int ntile_list;
};

// ---- stage 2b v4: Gaussian along y (and, on the transposed grid, along x) with TMA-fed tiles and a producer warp --
// Persistent blocks walk (slice, first row, column-strip) tiles of up to 8T rows.  The (8T + 2n) x 64 input box of a tile is fetched
// by cp.async.bulk.tensor.3d -- out-of-bounds rows/columns arrive as zeros = the reference's clipped window -- and
// completion is an mbarrier transaction count, so the fill spends no issue slots and no registers of the filter
// warps.  Arithmetic is identical to k_conv_cols2.  The first TMA version (one elected thread of warp 0 issued the
// box, a __syncthreads per tile) left 13 % of the warp samples in that barrier and 10 % behind the global loads of
// the tile list (ncu r01d).  Here warp 8 only feeds the pipeline (waits for a stage to be released, issues the
// next box), the eight filter warps hand a stage back through an `empty` mbarrier instead of a block barrier -- a
// warp that is done moves on to the next tile, whose box landed a tile ago -- and the tile list sits in shared memory.
// try_wait suspends the warp until the phase completes or a time limit passes; the default limit is short (ncu r01h:
// the producer warp came back ~900 times per tile while the filter warps worked, 1.8 M branches of a 26.6 M-instruction
// launch -- issue slots taken from the tap loops).  The hint raises the limit; completion still wakes the warp at once.
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(PS_MBAR_HINT_NS) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_h(unsigned long long *bar, unsigned parity, unsigned hint_ns) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int kMaxSmemTiles = 1024;
constexpr int kMaxTmaStages = 4;

template <int T, bool FMA>
__global__ void __launch_bounds__(288, PS_TMA_MINB) k_conv_cols_tma2(const __grid_constant__ CUtensorMap tmap, ColsTmaArgs a, u64 nz, int NS) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_full[kMaxTmaStages], s_empty[kMaxTmaStages];
  __shared__ float s_taps[1000];
  __shared__ ushort2 s_tiles[kMaxSmemTiles];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n = (a.len - 1) / 2;
  const int nrows = 8 * T + 2 * n;
  const unsigned stage_bytes = (unsigned)nrows * 64 * sizeof(float);
  const unsigned stage_stride = (stage_bytes + 127) & ~127u;
  for (int i = tid; i < a.len; i += 288) s_taps[i] = a.taps[i];
  const bool list_smem = a.tile_list && a.ntile_list <= kMaxSmemTiles;
  if (list_smem)
    for (int i = tid; i < a.ntile_list; i += 288)
      s_tiles[i] = make_ushort2((unsigned short)a.tile_list[2 * i], (unsigned short)a.tile_list[2 * i + 1]);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_slice = a.tile_list ? a.ntile_list : a.ytiles * a.xtiles;
  const int ntiles = a.slices * per_slice;
  // tile -> (first row, 64-column strip, 8-row groups to compute, slice); list entries are (row0, strip | groups << 12)
  auto decode = [&](int tile, int &xt, int &row0, int &ng, int &z) {
    z = tile / per_slice;
    const int i = tile - z * per_slice;
    if (a.tile_list) {
      int e0, e1;
      if (list_smem) {
        const ushort2 t = s_tiles[i];
        e0 = t.x;
        e1 = t.y;
      } else {
        e0 = a.tile_list[2 * i];
        e1 = a.tile_list[2 * i + 1];
      }
      row0 = e0;
      xt = e1 & 0xfff;
      ng = e1 >> 12;
    } else {
      const int yt = i / a.xtiles;
      xt = i - yt * a.xtiles;
      row0 = yt * 8 * T;
      ng = 8;
    }
  };
  if (w == 8) {  // producer
    if (lane == 0) {
      int s = 0, u = 0;  // stage, use count of that stage
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (u >= 1) {
          while (!mbar_try_wait(&s_empty[s], (unsigned)(u - 1) & 1u)) {}
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the stage -> async write
        }
        int xt, row0, ng, z;
        decode(tile, xt, row0, ng, z);
        mbar_expect_tx(&s_full[s], stage_bytes);
        tma_load_3d(s_raw + s * stage_stride, &tmap, xt * 64, row0 - n, z, &s_full[s]);
        if (++s == NS) {
          s = 0;
          ++u;
        }
      }
    }
    return;
  }
  int s = -1, u = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (++s == NS) {
      s = 0;
      ++u;
    }
    int xt, row0, ng, z;
    decode(tile, xt, row0, ng, z);
    while (!mbar_try_wait(&s_full[s], (unsigned)u & 1u)) {}
    if (w < ng && row0 + w * T < a.rows) {  // warp-uniform
      const u64 *win = reinterpret_cast<const u64 *>(s_raw + s * stage_stride) + (w * T) * 32 + lane;
      u64 acc[T], d[T];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = pk2(0.0f, 0.0f);
#pragma unroll
      for (int q = 0; q < T - 1; ++q) d[q] = win[q * 32];
      const u64 *wp = win + (T - 1) * 32;
      int kk = 0;
      for (; kk + T <= a.len; kk += T, wp += T * 32) {
#pragma unroll
        for (int u = 0; u < T; ++u) {
          d[(u + T - 1) % T] = wp[u * 32];
          const float f = s_taps[kk + u];
#pragma unroll
          for (int t = 0; t < T; ++t) acc[t] = tap2<FMA>(acc[t], d[(u + t) % T], f, nz);
        }
      }
#pragma unroll
      for (int u = 0; u < T; ++u) {
        if (kk + u < a.len) {
          d[(u + T - 1) % T] = wp[u * 32];
          const float f = s_taps[kk + u];
#pragma unroll
          for (int t = 0; t < T; ++t) acc[t] = tap2<FMA>(acc[t], d[(u + t) % T], f, nz);
        }
      }
      // the stage is no longer needed: release it before the stores
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);
      const int x = xt * 64 + lane * 2;
      if (x < a.cols) {
        const bool in1 = x + 1 < a.cols;
        if (a.transpose_out) {
          const int r0 = row0 + w * T;
          float lo[T], hi[T];
#pragma unroll
          for (int t = 0; t < T; ++t) upk2(acc[t], lo[t], hi[t]);
          float *o0 = a.out + (size_t)z * a.plane + (size_t)x * a.pitch + r0;
          if (r0 + T <= a.rows) {
#pragma unroll
            for (int t = 0; t < T; t += 4) *reinterpret_cast<float4 *>(o0 + t) = make_float4(lo[t], lo[t + 1], lo[t + 2], lo[t + 3]);
            if (in1) {
#pragma unroll
              for (int t = 0; t < T; t += 4)
                *reinterpret_cast<float4 *>(o0 + a.pitch + t) = make_float4(hi[t], hi[t + 1], hi[t + 2], hi[t + 3]);
            }
          } else {
#pragma unroll
            for (int t = 0; t < T; ++t)
              if (r0 + t < a.rows) {
                o0[t] = lo[t];
                if (in1) o0[a.pitch + t] = hi[t];
              }
          }
        } else {
          float *dst = a.out + (size_t)z * a.plane + x;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const int y = row0 + w * T + t;
            if (y < a.rows) {
              float lo, hi;
              upk2(acc[t], lo, hi);
              float *o = dst + (size_t)y * a.pitch;
              if (in1) *reinterpret_cast<float2 *>(o) = make_float2(lo, hi);
              else o[0] = lo;
            }
          }
        }
      }
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);
    }
  }
}

// ---- stage 2b v5: both Gaussian passes in ONE kernel, several messages per launch ---------------------------------
// gaussFilterDiag2d on the eigen-frame grid (multi_array_filter.hpp:277-317) without the intermediate grid: the
// x-filtered rows of a 64-column strip live in a shared-memory ring while a block walks down the strip.
//
// Input: the eigen-frame grid stored TRANSPOSED, UT[z][ex][ey] (the resampler writes it that way), so the x filter is a
// column filter over TMA boxes of (64 + 2 nx) ex-rows x 64 ey-columns, exactly like k_conv_cols_tma2.  Work item =
// (message, walk, slice): a walk is one 64-wide ex-strip and the interval of ey-rows some reader needs (MessagePlan::
// walks).  Step i of a walk:
//   X(i): box i (ey origin y0 - halo + 64 i) -> 64 x 64 x-filtered values, stored transposed into the ring as
//         ring[(64 i + c - dshift) mod cap][ex], c = ey column of the box; dshift = halo - ny makes the y windows start
//         on multiples of 8 ring rows, so the origin of every box stays 16-byte aligned AND an 8-row chunk of a y
//         window never straddles the wrap.  Warps whose 8 ex-rows no y output of the rectangle can reach (xmask) skip.
//   Y(i - K): 64 output rows (8 per warp) filtered along ey from the ring, stored to out[z][ey][ex].
// Arithmetic per output is the same ascending-tap RN(acc + RN(x f)) chain as the two-kernel route: bit-identical.
// Two work orders.  Fixed (DYN = false, the default): eight filter warps, 256 threads; every block walks its own list of
// the table the host dealt, and thread 0 requests the next box right behind the barrier that closes an x phase.
// Counter-drawn (DYN = true): a ninth warp is the producer -- it draws work items from an atomic counter (messages in
// launch order, longest walk first), publishes (message, walk, block, slice) next to each box and issues the TMA; the
// filter warps follow the published records and hand a stage back through its `empty` mbarrier right after the x phase.
// In both, the filter warps meet at a named barrier between the phases.

// The last REM (= taps mod 8, odd because every filter has 2 n + 1 taps) taps of a phase as straight-line code: every
// load of the tail is issued before its first use (ncu r02f: the per-tap branches of a predicated tail ran at half the
// multiply-add density of the main loop).  Row 0 of the tail is at w0, rows 1.. at w1 + (row - 1) * stride: the ring
// of the y phase may wrap between the two.
template <bool FMA, int REM>
__device__ __forceinline__ void gauss_tail(u64 (&acc)[8], u64 (&d)[8], const u64 *w0, const u64 *w1, int stride,
                                           const float *taps, u64 nz) {
  constexpr int T = 8;
  u64 nd[REM];
  float f[REM];
  nd[0] = w0[0];
#pragma unroll
  for (int uu = 1; uu < REM; ++uu) nd[uu] = w1[(uu - 1) * stride];
#pragma unroll
  for (int uu = 0; uu < REM; ++uu) f[uu] = taps[uu];
#pragma unroll
  for (int uu = 0; uu < REM; ++uu) {
    d[(uu + T - 1) % T] = nd[uu];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = tap2<FMA>(acc[t], d[(uu + t) % T], f[uu], nz);
  }
}
template <bool FMA>
__device__ __forceinline__ void gauss_tail_any(int rem, u64 (&acc)[8], u64 (&d)[8], const u64 *w0, const u64 *w1,
                                               int stride, const float *taps, u64 nz) {
  switch (rem) {  // warp-uniform
    case 1: gauss_tail<FMA, 1>(acc, d, w0, w1, stride, taps, nz); break;
    case 3: gauss_tail<FMA, 3>(acc, d, w0, w1, stride, taps, nz); break;
    case 5: gauss_tail<FMA, 5>(acc, d, w0, w1, stride, taps, nz); break;
    case 7: gauss_tail<FMA, 7>(acc, d, w0, w1, stride, taps, nz); break;
    default: break;  // even tap counts never reach this kernel (can_batch)
  }
}
constexpr int kRingPitch = 68;  // floats per ring row: 64 + 4 keeps rows 16-byte aligned and the transposed stores 2-way at worst
struct GaussMsg {
  float *out;                  // [z][ey][ex], pitch EP, plane oplane
  const float *taps_x, *taps_y;
  const int *walks;            // [nwalks][4] = strip, first row, 8-row groups, offset into masks
  const unsigned char *masks;  // per x block: which 8-column groups to filter
  size_t oplane;
  int len_x, len_y, EH, EW, EP, nwalks, lag, halo;
  int item0;                   // first work item of this message in the launch; items = nwalks * R, slice innermost
};
struct GaussBatch {
  GaussMsg m[kMaxBatch];
  unsigned *counter;           // zero at launch
  int nmsg, R, total_items;
  unsigned stage_stride;       // bytes between the TMA stages (the largest box of the batch, 128-byte multiple)
  int ring_rows;               // 64 * (largest lag of the batch + 1)
  int stages;                  // TMA stages, 1 or 2
  unsigned hint_ns;            // suspend-time hint of the mbarrier waits
  // fixed work order: block c processes items cta_items[cta_off[c] .. cta_off[c + 1]), each (message << 24 | walk << 12 |
  // slice); the host deals the message-major item list to the least loaded block (list scheduling on the tap counts)
  const int *cta_off, *cta_items;
};
constexpr int kMaxFusedTaps = 256;
struct alignas(64) TmapBatch {
  CUtensorMap t[kMaxBatch];
};
__device__ __forceinline__ void bar_sync_filter() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Register budget of k_gauss_xy: 64 per thread.  Two blocks of the fixed-order kernel (8 warps) then hold exactly half of
// the SM's 64 K registers and four 256-thread blocks of the 32-register kernels around it fit beside them (72 registers
// and a producer warp left room for two: A/B 464 -> 469 images/s at 64 registers although the kernel alone is 3 %
// slower; 56 registers spill 40 bytes and lose: 455).  PS_GAUSS_MINB (a __launch_bounds__ minimum) selects a budget for
// A/B builds instead.
#ifndef PS_GAUSS_MAXREG
#define PS_GAUSS_MAXREG 64
#endif
template <bool FMA, bool DYN>
#ifndef PS_GAUSS_MINB
__global__ void __maxnreg__(PS_GAUSS_MAXREG) k_gauss_xy(
#else
__global__ void __launch_bounds__(288, PS_GAUSS_MINB) k_gauss_xy(
#endif
    const __grid_constant__ TmapBatch tm, const __grid_constant__ GaussBatch b, u64 nz) {
  constexpr int T = 8;
  const int NS = b.stages;  // TMA stages: 1 (the next box is requested when the x phase ends and lands during the y phase) or 2
  extern __shared__ __align__(128) unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_full[2], s_empty[2];
  __shared__ int4 s_meta[2][2];  // [0] = (message or -1, strip, first row, groups)  [1] = (block index, slice, mask, 0)
  __shared__ float s_tx[kMaxFusedTaps], s_ty[kMaxFusedTaps];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float *ring = reinterpret_cast<float *>(s_raw + NS * b.stage_stride);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // Static order (DYN = false): 256 threads, no producer warp.  Thread 0 requests the block's boxes itself -- NS of them
  // up front, the next one each time the barrier behind an x phase says that every warp is done reading a stage -- from a
  // cursor over the block's item list that lives in shared memory (registers are the scarce resource of this kernel: 64
  // per thread, so that two blocks leave half of the register file to the other kernels of the image).
  __shared__ int s_pc[8];  // idx, step, steps of the walk, message, slice, strip, first row, stage
  auto issue_next_box = [&]() {  // thread 0 only
    int idx = s_pc[0], step = s_pc[1], nsteps = s_pc[2];
    if (step == nsteps) {
      if (idx >= b.cta_off[blockIdx.x + 1]) return;
      const int it = b.cta_items[idx++];
      const int mi = it >> 24;
      const GaussMsg &gg = b.m[mi];
      const int4 we = *reinterpret_cast<const int4 *>(gg.walks + 4 * ((it >> 12) & 0xfff));
      nsteps = (we.z + 7) / 8 + gg.lag;
      step = 0;
      s_pc[0] = idx; s_pc[2] = nsteps; s_pc[3] = mi; s_pc[4] = it & 0xfff; s_pc[5] = we.x; s_pc[6] = we.y;
    }
    const int mi = s_pc[3], st = s_pc[7];
    const GaussMsg &g = b.m[mi];
    const int nx = (g.len_x - 1) / 2;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the stage -> async write
    mbar_expect_tx(&s_full[st], (unsigned)(64 + 2 * nx) * 64u * sizeof(float));
    tma_load_3d(s_raw + st * b.stage_stride, &tm.t[mi], s_pc[6] - g.halo + 64 * step, s_pc[5] * 64 - nx, s_pc[4], &s_full[st]);
    s_pc[1] = step + 1;
    s_pc[7] = st + 1 == NS ? 0 : st + 1;
  };
  if (!DYN && tid == 0) {
    s_pc[0] = b.cta_off[blockIdx.x]; s_pc[1] = 0; s_pc[2] = 0; s_pc[7] = 0;
    for (int i = 0; i < NS; ++i) issue_next_box();
  }
  if (DYN && w == 8) {  // producer warp of the counter-drawn order (288 threads)
    if (lane == 0) {
      int s = 0, u = 0;
      auto next_stage = [&]() {
        if (++s == NS) {
          s = 0;
          ++u;
        }
      };
      auto wait_free = [&]() {
        if (u >= 1) {
          while (!mbar_try_wait_h(&s_empty[s], (unsigned)(u - 1) & 1u, b.hint_ns)) {}
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the stage -> async write
        }
      };
      // Work items in a fixed order: this block's list of the table the host dealt (cta_off / cta_items).  The filter
      // warps walk the same list on their own, so nothing but the boxes passes from the producer to them.
      // (DYN: items drawn from an atomic counter and published through s_meta instead -- same results, same speed;
      // racecheck cannot see the mbarrier that orders those shared-memory records and reports them.)
      int idx = DYN ? 0 : b.cta_off[blockIdx.x];
      const int idx_end = DYN ? 0 : b.cta_off[blockIdx.x + 1];
      for (;;) {
        int mi = 0, wk = 0, z = 0;
        if (DYN) {
          const int item = (int)atomicAdd(b.counter, 1u);
          if (item >= b.total_items) break;
          while (mi + 1 < b.nmsg && item >= b.m[mi + 1].item0) ++mi;
          const int local = item - b.m[mi].item0;
          wk = local / b.R;
          z = local - wk * b.R;
        } else {
          if (idx >= idx_end) break;
          const int it = b.cta_items[idx++];
          mi = it >> 24;
          wk = (it >> 12) & 0xfff;
          z = it & 0xfff;
        }
        const GaussMsg &g = b.m[mi];
        const int4 we = *reinterpret_cast<const int4 *>(g.walks + 4 * wk);
        const int nxb = (we.z + 7) / 8 + g.lag;
        const int nx = (g.len_x - 1) / 2;
        const unsigned bytes = (unsigned)(64 + 2 * nx) * 64u * sizeof(float);
        for (int i = 0; i < nxb; ++i) {
          wait_free();
          if (DYN) {
            s_meta[s][0] = make_int4(mi, we.x, we.y, we.z);
            s_meta[s][1] = make_int4(i, z, (int)g.masks[we.w + i], 0);
          }
          mbar_expect_tx(&s_full[s], bytes);
          tma_load_3d(s_raw + s * b.stage_stride, &tm.t[mi], we.y - g.halo + 64 * i, we.x * 64 - nx, z, &s_full[s]);
          next_stage();
        }
      }
      if (DYN) {
        wait_free();
        s_meta[s][0] = make_int4(-1, 0, 0, 0);
        mbar_arrive(&s_full[s]);
      }
    }
    return;
  }
  // the ring starts defined: skipped x groups leave stale cells that only feed outputs nobody reads
  {
    for (int i = tid; i < b.ring_rows * kRingPitch; i += 256) ring[i] = 0.0f;
  }
  bar_sync_filter();
  int s = -1, u = 0, cur = -1, slot = 0;
  // static order: the walk this block is in (see the producer)
  int idx = DYN ? 0 : b.cta_off[blockIdx.x], step = 0, nsteps = 0, s_mi = 0, s_z = 0, mask_cur = 0;
  const int idx_end = DYN ? 0 : b.cta_off[blockIdx.x + 1];
  int4 s_we = make_int4(0, 0, 0, 0);
  for (;;) {
    int4 m0, m1;
    if (!DYN) {
      if (step == nsteps) {
        if (idx >= idx_end) break;
        const int it = b.cta_items[idx++];
        s_mi = it >> 24;
        const int wk = (it >> 12) & 0xfff;
        s_z = it & 0xfff;
        const GaussMsg &gg = b.m[s_mi];
        s_we = *reinterpret_cast<const int4 *>(gg.walks + 4 * wk);
        nsteps = (s_we.z + 7) / 8 + gg.lag;
        step = 0;
        mask_cur = gg.masks[s_we.w];
      }
      m0 = make_int4(s_mi, s_we.x, s_we.y, s_we.z);
      m1 = make_int4(step, s_z, mask_cur, 0);
      ++step;
      if (step < nsteps) mask_cur = b.m[s_mi].masks[s_we.w + step];  // the next box's mask, in flight during this step
    }
    if (++s == NS) {
      s = 0;
      ++u;
    }
    while (!mbar_try_wait_h(&s_full[s], (unsigned)u & 1u, b.hint_ns)) {}
    if (DYN) {
      m0 = s_meta[s][0];
      m1 = s_meta[s][1];
      if (m0.x < 0) break;
    }
    const GaussMsg &g = b.m[m0.x];
    const int nx = (g.len_x - 1) / 2, ny = (g.len_y - 1) / 2;
    const int cap = 64 * (g.lag + 1), dshift = g.halo - ny;
    const int i = m1.x, z = m1.y;
    slot = (i == 0 || slot == g.lag) ? 0 : slot + 1;  // i mod (K + 1): a block takes the steps of a walk in order
    if (i == 0 && m0.x != cur) {  // every warp is past the previous walk's last barrier: the tap arrays are free
      cur = m0.x;
      for (int k = tid; k < g.len_x; k += 256) s_tx[k] = g.taps_x[k];
      for (int k = tid; k < g.len_y; k += 256) s_ty[k] = g.taps_y[k];
      bar_sync_filter();
    }
    // ---- X(i): filter along ex, store transposed into the ring ----
    if ((m1.z >> w) & 1) {
      const u64 *win = reinterpret_cast<const u64 *>(s_raw + s * b.stage_stride) + (w * T) * 32 + lane;
      u64 acc[T], d[T];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = pk2(0.0f, 0.0f);
#pragma unroll
      for (int q = 0; q < T - 1; ++q) d[q] = win[q * 32];
      const u64 *wp = win + (T - 1) * 32;
      int kk = 0;
      for (; kk + T <= g.len_x; kk += T, wp += T * 32) {
#pragma unroll
        for (int uu = 0; uu < T; ++uu) {
          d[(uu + T - 1) % T] = wp[uu * 32];
          const float f = s_tx[kk + uu];
#pragma unroll
          for (int t = 0; t < T; ++t) acc[t] = tap2<FMA>(acc[t], d[(uu + t) % T], f, nz);
        }
      }
      gauss_tail_any<FMA>(g.len_x - kk, acc, d, wp, wp + 32, 32, s_tx + kk, nz);
      float lo[T], hi[T];
#pragma unroll
      for (int t = 0; t < T; ++t) upk2(acc[t], lo[t], hi[t]);
      int q0 = 64 * slot + 2 * lane - dshift;  // ring row of ey column 2*lane of this box (slot = i mod (K + 1))
      if (q0 < 0) q0 += cap;
      int q1 = q0 + 1;
      if (q1 == cap) q1 = 0;
      float4 *r0 = reinterpret_cast<float4 *>(ring + q0 * kRingPitch + w * T);
      float4 *r1 = reinterpret_cast<float4 *>(ring + q1 * kRingPitch + w * T);
      r0[0] = make_float4(lo[0], lo[1], lo[2], lo[3]);
      r0[1] = make_float4(lo[4], lo[5], lo[6], lo[7]);
      r1[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
      r1[1] = make_float4(hi[4], hi[5], hi[6], hi[7]);
    }
    if (DYN) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);  // the box is consumed: the producer may refill the stage
    }
    bar_sync_filter();
    if (!DYN && tid == 0) issue_next_box();  // every warp is past its reads of this stage: refill it during the y phase
    // ---- Y(i - K): filter along ey from the ring ----
    const int ob = i - g.lag;
    if (ob >= 0) {
      const int row0 = m0.z + 64 * ob + w * T;
      if (w < m0.w - 8 * ob && row0 < g.EH) {  // warp-uniform
        // (64 ob + 8 w) mod cap without a division: ob = i - K is congruent to slot + 1 modulo K + 1
        int q = 64 * (slot == g.lag ? 0 : slot + 1) + w * T;  // multiple of 8; rows q .. q + 7 + 2 ny (mod cap) hold the window
        const u64 *rbase = reinterpret_cast<const u64 *>(ring) + lane;
        u64 acc[T], d[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = pk2(0.0f, 0.0f);
#pragma unroll
        for (int t = 0; t < T - 1; ++t) d[t] = rbase[(q + t) * (kRingPitch / 2)];
        q += T - 1;  // next row to load; q - (T-1) was a multiple of 8, so q + 1 may hit cap only after this row
        int kk = 0;
        for (; kk + T <= g.len_y; kk += T) {
          const u64 *wp = rbase + q * (kRingPitch / 2);  // rows q, then (q + 1 wrapped) .. : q % 8 == 7
          d[T - 1] = wp[0];
          q = q + 1 == cap ? 0 : q + 1;
          const u64 *wq = rbase + q * (kRingPitch / 2);  // q % 8 == 0: seven more rows without a wrap
          {
            const float f = s_ty[kk];
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = tap2<FMA>(acc[t], d[t % T], f, nz);
          }
#pragma unroll
          for (int uu = 1; uu < T; ++uu) {
            d[(uu + T - 1) % T] = wq[(uu - 1) * (kRingPitch / 2)];
            const float f = s_ty[kk + uu];
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = tap2<FMA>(acc[t], d[(uu + t) % T], f, nz);
          }
          q += T - 1;  // q % 8 == 7 again (never reaches cap: cap % 8 == 0)
        }
        {
          const int q1 = q + 1 == cap ? 0 : q + 1;  // q % 8 == 7: at most one wrap, right after the tail's first row
          gauss_tail_any<FMA>(g.len_y - kk, acc, d, rbase + q * (kRingPitch / 2), rbase + q1 * (kRingPitch / 2),
                              kRingPitch / 2, s_ty + kk, nz);
        }
        const int x = m0.y * 64 + lane * 2;
        if (x < g.EW) {
          const bool in1 = x + 1 < g.EW;
          float *dst = g.out + (size_t)z * g.oplane + x;
          if (in1 && row0 + T <= g.EH) {  // all eight rows and both columns inside: one running pointer, no tests
            float *o = dst + (size_t)row0 * g.EP;
#pragma unroll
            for (int t = 0; t < T; ++t, o += g.EP) {
              float lo, hi;
              upk2(acc[t], lo, hi);
              *reinterpret_cast<float2 *>(o) = make_float2(lo, hi);
            }
          } else {
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const int y = row0 + t;
              if (y < g.EH) {
                float lo, hi;
                upk2(acc[t], lo, hi);
                float *o = dst + (size_t)y * g.EP;
                if (in1) *reinterpret_cast<float2 *>(o) = make_float2(lo, hi);
                else o[0] = lo;
              }
            }
          }
        }
      }
      bar_sync_filter();  // the next x block overwrites ring rows this y block read
    }
  }
}

// ---- stage 2b v2: Gaussian along x, two adjacent rows per thread -------------------------------------------------
// A block stages PAIRS row pairs (+halo) as float2 (row 2q, row 2q+1) in the "transposed by T" layout
// (element i at (i % T) * S + i / T); a thread owns T consecutive outputs of both rows of a pair.
template <int T>
__global__ void __launch_bounds__(256, 4) k_conv_rows2(ConvArgs a, u64 nz, int PAIRS, int S) {
  extern __shared__ float2 s_pairs[];  // [PAIRS][T*S]
  __shared__ float s_taps[1000];
  const int tid = threadIdx.x;
  for (int i = tid; i < a.len; i += 256) s_taps[i] = a.taps[i];
  const int n = (a.len - 1) / 2;
  const int y0 = blockIdx.x * (2 * PAIRS);
  const float *src = a.in + (size_t)blockIdx.y * a.plane;
  float *dst = a.out + (size_t)blockIdx.y * a.plane;
  const int G = (a.cols + T - 1) / T;
  const int span = G * T + 2 * n;
  const int rowsz = T * S;
  for (int pr = 0; pr < PAIRS; ++pr) {
    const int ya = y0 + 2 * pr, yb = ya + 1;
    const float *ra = src + (size_t)ya * a.pitch, *rb = src + (size_t)yb * a.pitch;
    for (int i = tid; i < span; i += 256) {
      int x = i - n;
      bool okx = x >= 0 && x < a.cols;
      float va = (okx && ya < a.rows) ? __ldg(ra + x) : 0.0f;
      float vb = (okx && yb < a.rows) ? __ldg(rb + x) : 0.0f;
      s_pairs[pr * rowsz + (i % T) * S + i / T] = make_float2(va, vb);
    }
  }
  __syncthreads();
  const int items = PAIRS * G;
  for (int it = tid; it < items; it += 256) {
    const int pr = it / G, g = it % G;
    const int ya = y0 + 2 * pr;
    if (ya >= a.rows) continue;
    const u64 *tile = reinterpret_cast<const u64 *>(s_pairs) + pr * rowsz + g;
    u64 acc[T], d[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = pk2(0.0f, 0.0f);
#pragma unroll
    for (int j = 0; j < T - 1; ++j) d[j] = tile[j * S];
    int kk = 0;
    for (; kk + T <= a.len; kk += T) {
#pragma unroll
      for (int u = 0; u < T; ++u) {
        d[(u + T - 1) % T] = tile[((u + T - 1) % T) * S + (kk + u + T - 1) / T];
        const float f = s_taps[kk + u];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = add2_rn(acc[t], mul2_rn(d[(u + t) % T], f, nz));
      }
    }
#pragma unroll
    for (int u = 0; u < T; ++u) {
      if (kk + u < a.len) {
        d[(u + T - 1) % T] = tile[((u + T - 1) % T) * S + (kk + u + T - 1) / T];
        const float f = s_taps[kk + u];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = add2_rn(acc[t], mul2_rn(d[(u + t) % T], f, nz));
      }
    }
    float lo[T], hi[T];
#pragma unroll
    for (int t = 0; t < T; ++t) upk2(acc[t], lo[t], hi[t]);
    float *oa = dst + (size_t)ya * a.pitch + g * T;
    const bool full = g * T + T <= a.cols;
    const bool al = (a.pitch & 3) == 0;
    if (full && al) {
#pragma unroll
      for (int t = 0; t < T; t += 4) *reinterpret_cast<float4 *>(oa + t) = make_float4(lo[t], lo[t + 1], lo[t + 2], lo[t + 3]);
      if (ya + 1 < a.rows) {
        float *ob = oa + a.pitch;
#pragma unroll
        for (int t = 0; t < T; t += 4) *reinterpret_cast<float4 *>(ob + t) = make_float4(hi[t], hi[t + 1], hi[t + 2], hi[t + 3]);
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t)
        if (g * T + t < a.cols) {
          oa[t] = lo[t];
          if (ya + 1 < a.rows) oa[a.pitch + t] = hi[t];
        }
    }
  }
}


// ---- stage 2b v3: Gaussian along x, persistent blocks, cp.async double buffering ------------------------------------
// Same layout and arithmetic as k_conv_rows2, but a block walks (slice, row-tile) tiles and copies the NEXT tile's
// row pairs with 4-byte cp.async (LDGSTS, zero-filled outside the grid via src-size 0) straight into the
// "transposed by T" float2 layout while it filters the current one: no registers, no exposed load latency.
__device__ __forceinline__ void cp_async_f32(float *dst_smem, const float *src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 4 : 0)
               : "memory");
}

template <int T>
__global__ void __launch_bounds__(256, 2) k_conv_rows3(ConvArgs a, u64 nz, int PAIRS, int S, int slices, int ytiles) {
  extern __shared__ __align__(16) float2 s_pairs[];  // 2 stages x [PAIRS][T*S]
  __shared__ float s_taps[1000];
  const int tid = threadIdx.x;
  for (int i = tid; i < a.len; i += 256) s_taps[i] = a.taps[i];
  const int n = (a.len - 1) / 2;
  const int G = (a.cols + T - 1) / T;
  const int span = G * T + 2 * n;
  const int rowsz = T * S;
  const int stage_elems = PAIRS * rowsz;  // float2 units
  const int ntiles = slices * ytiles;

  auto fill = [&](int tile, int st) {
    const int z = tile / ytiles, y0 = (tile - z * ytiles) * (2 * PAIRS);
    const float *src = a.in + (size_t)z * a.plane;
    float *base = reinterpret_cast<float *>(s_pairs + (size_t)st * stage_elems);
    for (int row = 0; row < 2 * PAIRS; ++row) {
      const int y = y0 + row;
      const bool oky = y < a.rows;
      const float *rp = src + (size_t)(oky ? y : 0) * a.pitch;
      float *dst = base + ((row >> 1) * rowsz) * 2 + (row & 1);
      for (int i = tid; i < span; i += 256) {
        const int x = i - n;
        const bool ok = oky && x >= 0 && x < a.cols;
        cp_async_f32(dst + ((i % T) * S + i / T) * 2, ok ? rp + x : a.in, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = blockIdx.x;
  if (tile < ntiles) fill(tile, 0);
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int st = it & 1;
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      fill(next, st ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int z = tile / ytiles, y0 = (tile - z * ytiles) * (2 * PAIRS);
    float *dst = a.out + (size_t)z * a.plane;
    const u64 *stage = reinterpret_cast<const u64 *>(s_pairs) + (size_t)st * stage_elems;
    const int items = PAIRS * G;
    for (int itx = tid; itx < items; itx += 256) {
      const int pr = itx / G, g = itx - pr * G;
      const int ya = y0 + 2 * pr;
      if (ya >= a.rows) continue;
      const u64 *tile_p = stage + pr * rowsz + g;
      u64 acc[T], d[T];
#pragma unroll
      for (int t = 0; t < T; ++t) acc[t] = pk2(0.0f, 0.0f);
#pragma unroll
      for (int j = 0; j < T - 1; ++j) d[j] = tile_p[j * S];
      // window element m (m >= T-1) sits at ((m % T) * S + m / T); m = kk + u + T - 1 with kk a multiple of T
      const u64 *wp = tile_p;
      int kk = 0;
      for (; kk + T <= a.len; kk += T, ++wp) {
#pragma unroll
        for (int u = 0; u < T; ++u) {
          d[(u + T - 1) % T] = wp[((u + T - 1) % T) * S + (u + T - 1) / T];
          const float f = s_taps[kk + u];
#pragma unroll
          for (int t = 0; t < T; ++t) acc[t] = add2_rn(acc[t], mul2_rn(d[(u + t) % T], f, nz));
        }
      }
#pragma unroll
      for (int u = 0; u < T; ++u) {
        if (kk + u < a.len) {
          d[(u + T - 1) % T] = wp[((u + T - 1) % T) * S + (u + T - 1) / T];
          const float f = s_taps[kk + u];
#pragma unroll
          for (int t = 0; t < T; ++t) acc[t] = add2_rn(acc[t], mul2_rn(d[(u + t) % T], f, nz));
        }
      }
      float lo[T], hi[T];
#pragma unroll
      for (int t = 0; t < T; ++t) upk2(acc[t], lo[t], hi[t]);
      float *oa = dst + (size_t)ya * a.pitch + g * T;
      const bool full = g * T + T <= a.cols;
      const bool al = (a.pitch & 3) == 0;
      if (full && al) {
#pragma unroll
        for (int t = 0; t < T; t += 4) *reinterpret_cast<float4 *>(oa + t) = make_float4(lo[t], lo[t + 1], lo[t + 2], lo[t + 3]);
        if (ya + 1 < a.rows) {
          float *ob = oa + a.pitch;
#pragma unroll
          for (int t = 0; t < T; t += 4) *reinterpret_cast<float4 *>(ob + t) = make_float4(hi[t], hi[t + 1], hi[t + 2], hi[t + 3]);
        }
      } else {
#pragma unroll
        for (int t = 0; t < T; ++t)
          if (g * T + t < a.cols) {
            oa[t] = lo[t];
            if (ya + 1 < a.rows) oa[a.pitch + t] = hi[t];
          }
      }
    }
    __syncthreads();  // stage st may be refilled in the next iteration
  }
}

// ---- message stage 3: read back, log, +M, shift to the parent frame, combine ------------------------------
// findrot.cpp:431-448 plus the addGrid2 calls that consume the message (findrot.cpp:637-654, :201, :221-222).
struct EpiArgs {
  const float *src;     // diag: [R][H][W] filtered probabilities; general: eigen-frame [R][EH][EP]
  int general;
  Affine T34;
  int EH, EW, EP;
  const int *xout, *yout;  // [R][W], [R][H]
  const int *shift_xy;  // [R][2] (dx, dy) when every table row is a pure shift, else null
  const int *max_enc;   // M of the message input
  int R, H, W;
  // out0 = (acc0 ? acc0 + v : v) (+ add0);  out1 = add1 + v
  float *out0;
  const float *acc0;
  const float *add0;
  float *out1;
  const float *add1;
  int *max0;            // optional: max over out0
  int *max1;            // optional: max over out1
  unsigned long long *amax0;  // optional: first-maximum key of out0 (the part's final marginal), see argmax_key
};

__global__ void __launch_bounds__(256) k_epilogue(EpiArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int r = blockIdx.z;
  float m0 = -INFINITY, m1 = -INFINITY;
  unsigned long long key0 = 0;
  if (x < a.W) {
    const size_t HW = (size_t)a.H * a.W;
    const size_t cell = (size_t)r * HW + (size_t)y * a.W + x;
    const float M = dec_f(*a.max_enc);
    int xs = a.xout[r * a.W + x], ys = a.yout[r * a.H + y];
    float v = kLogZero;
    if ((xs | ys) >= 0) {
      float d;
      if (a.general) {
        double x1, y1;
        affine_map(a.T34, (double)xs, (double)ys, x1, y1);
        d = bilinear_at(a.src + (size_t)r * a.EH * a.EP, a.EH, a.EW, a.EP, x1, y1);
      } else {
        d = __ldg(&a.src[(size_t)r * HW + (size_t)ys * a.W + xs]);
      }
      v = __fadd_rn(log_f64(d), M);
    }
    if (a.out0) {
      float o = a.acc0 ? __fadd_rn(a.acc0[cell], v) : v;
      if (a.add0) o = __fadd_rn(o, a.add0[cell]);
      a.out0[cell] = o;
      m0 = o;
      if (a.amax0 && o == o) key0 = argmax_key(o, (unsigned)cell);
    }
    if (a.out1) {
      float o = __fadd_rn(a.add1[cell], v);
      a.out1[cell] = o;
      m1 = o;
    }
  }
  if (a.max0) block_max_to(m0, a.max0);
  if (a.max1) {
    __syncthreads();
    block_max_to(m1, a.max1);
  }
  if (a.amax0) block_key_max_to(key0, a.amax0);
}

// ---- message stage 3 v2: four consecutive cells per thread -------------------------------------------------------
// Same arithmetic as k_epilogue.  The per-row quantities (destination row -> source row, T34 row products, slice
// bases) are computed once per thread, table / accumulator / output traffic is 128-bit, and the block maximum is
// taken once over the four cells.  ncu r01a: k_epilogue spent 269 warp-instructions per cell, most of them
// parameter and address arithmetic.
template <bool GENERAL>
__global__ void __launch_bounds__(256) k_epilogue2(EpiArgs a, int XG /* ceil(W/4) */) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  float m0 = -INFINITY, m1 = -INFINITY;
  unsigned long long key0 = 0;
  if (it < XG * a.H) {
    const int y = it / XG, x0 = (it - y * XG) * 4;
    const size_t HW = (size_t)a.H * a.W;
    const size_t cell0 = (size_t)r * HW + (size_t)y * a.W + x0;
    const float M = dec_f(*a.max_enc);
    const int nx = min(4, a.W - x0);
    const bool vec = (a.W & 3) == 0;  // rows and slices are then 16-byte aligned
    int xs[4], ys;
    if (a.shift_xy) {
      const int dx = a.shift_xy[2 * r];
      ys = y + a.shift_xy[2 * r + 1];
      if ((unsigned)ys >= (unsigned)a.H) ys = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xs[j] = x0 + j + dx;
        if ((unsigned)xs[j] >= (unsigned)a.W || j >= nx) xs[j] = -1;
      }
    } else if (ys = a.yout[r * a.H + y], vec) {
      int4 t = *reinterpret_cast<const int4 *>(a.xout + r * a.W + x0);
      xs[0] = t.x; xs[1] = t.y; xs[2] = t.z; xs[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) xs[j] = j < nx ? a.xout[r * a.W + x0 + j] : -1;
    }
    float v[4];
    double ty1 = 0.0, ty4 = 0.0;
    const float *srcr;
    if (GENERAL) {
      ty1 = __dmul_rn(a.T34.m[1], (double)ys);
      ty4 = __dmul_rn(a.T34.m[4], (double)ys);
      srcr = a.src + (size_t)r * a.EH * a.EP;
    } else {
      srcr = a.src + (size_t)r * HW + (size_t)(ys < 0 ? 0 : ys) * a.W;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = kLogZero;
      if ((xs[j] | ys) >= 0) {
        float d;
        if (GENERAL) {
          const double xd = (double)xs[j];
          const double x1 = __dadd_rn(__dadd_rn(__dmul_rn(a.T34.m[0], xd), ty1), a.T34.m[2]);
          const double y1 = __dadd_rn(__dadd_rn(__dmul_rn(a.T34.m[3], xd), ty4), a.T34.m[5]);
          d = bilinear_at(srcr, a.EH, a.EW, a.EP, x1, y1);
        } else {
          d = __ldg(srcr + xs[j]);
        }
        v[j] = __fadd_rn(log_f64(d), M);
      }
    }
    if (a.out0) {
      float o[4] = {v[0], v[1], v[2], v[3]};
      if (a.acc0) {
        if (vec) {
          float4 t = *reinterpret_cast<const float4 *>(a.acc0 + cell0);
          o[0] = __fadd_rn(t.x, o[0]); o[1] = __fadd_rn(t.y, o[1]); o[2] = __fadd_rn(t.z, o[2]); o[3] = __fadd_rn(t.w, o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < nx) o[j] = __fadd_rn(a.acc0[cell0 + j], o[j]);
        }
      }
      if (a.add0) {
        if (vec) {
          float4 t = *reinterpret_cast<const float4 *>(a.add0 + cell0);
          o[0] = __fadd_rn(o[0], t.x); o[1] = __fadd_rn(o[1], t.y); o[2] = __fadd_rn(o[2], t.z); o[3] = __fadd_rn(o[3], t.w);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < nx) o[j] = __fadd_rn(o[j], a.add0[cell0 + j]);
        }
      }
      if (vec) *reinterpret_cast<float4 *>(a.out0 + cell0) = make_float4(o[0], o[1], o[2], o[3]);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nx) a.out0[cell0 + j] = o[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) if (j < nx) m0 = fmaxf(m0, o[j]);
      if (a.amax0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nx && o[j] == o[j]) {
            unsigned long long k = argmax_key(o[j], (unsigned)(cell0 + j));
            key0 = k > key0 ? k : key0;
          }
      }
    }
    if (a.out1) {
      float o[4];
      if (vec) {
        float4 t = *reinterpret_cast<const float4 *>(a.add1 + cell0);
        o[0] = __fadd_rn(t.x, v[0]); o[1] = __fadd_rn(t.y, v[1]); o[2] = __fadd_rn(t.z, v[2]); o[3] = __fadd_rn(t.w, v[3]);
        *reinterpret_cast<float4 *>(a.out1 + cell0) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = -INFINITY;
          if (j < nx) {
            o[j] = __fadd_rn(a.add1[cell0 + j], v[j]);
            a.out1[cell0 + j] = o[j];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) if (j < nx) m1 = fmaxf(m1, o[j]);
    }
  }
  if (a.max0) block_max_to(m0, a.max0);
  if (a.max1) {
    __syncthreads();
    block_max_to(m1, a.max1);
  }
  if (a.amax0) block_key_max_to(key0, a.amax0);
}


// ---- message stage 3 v3: lean variant of k_epilogue2<false> for the common case ------------------------------------
// Preconditions checked by the launcher: the translation is a pure shift (shift_xy), W % 4 == 0 and every grid is
// 16-byte aligned.  32-bit index arithmetic (R*H*W < 2^31 is enforced by ps_create), no per-cell branches: the four
// cells are evaluated unconditionally on clamped addresses and invalid ones are replaced by LOG_ZERO at the end.
// ncu r01c: k_epilogue2 spent 98 warp-instructions per cell, a third of them IMAD/ISETP/BRA bookkeeping.
struct EpiBatch {
  EpiArgs a[kMaxBatch];
};
__global__ void __launch_bounds__(256) k_epilogue3(const __grid_constant__ EpiBatch eb, FastDiv XG) {
  const EpiArgs &a = eb.a[blockIdx.z];
  // block-level folds: warp redux on the monotone integer encoding, one shared atomic per warp, one global per block
  __shared__ int s_m0, s_m1;
  __shared__ unsigned long long s_key;
  if (threadIdx.x == 0) {
    s_m0 = PS_ENC_NEG_INF;
    s_m1 = PS_ENC_NEG_INF;
    s_key = 0ull;
  }
  __syncthreads();
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  float m0 = -INFINITY, m1 = -INFINITY;
  unsigned long long key0 = 0;
  if (it < XG.d * a.H) {
    const int y = (int)XG.div((unsigned)it), x0 = (it - y * (int)XG.d) * 4;
    const int HW = a.H * a.W;
    const int cell0 = r * HW + y * a.W + x0;
    const float M = dec_f(*a.max_enc);
    const int dx = a.shift_xy[2 * r], ys = y + a.shift_xy[2 * r + 1];
    const bool rowok = (unsigned)ys < (unsigned)a.H;
    const int srow = r * HW + (rowok ? ys : 0) * a.W;  // one 32-bit index per load: R*H*W < 2^31
    float d[4], v[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xs = x0 + j + dx;
      ok[j] = rowok & ((unsigned)xs < (unsigned)a.W);
      d[j] = __ldg(a.src + (srow + (ok[j] ? xs : 0)));
    }
#ifndef PS_SLOW_MATH
    // four independent table-driven logs, no per-cell branches; the rare Ziv fallbacks are patched afterwards
    bool redo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = log_fast_nb(d[j], redo[j]);
    if (redo[0] | redo[1] | redo[2] | redo[3]) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (redo[j]) v[j] = log_slow_call(d[j]);
    }
#else
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = log_slow(d[j]);
#endif
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ok[j] ? __fadd_rn(v[j], M) : kLogZero;
    if (a.out0) {
      float o[4] = {v[0], v[1], v[2], v[3]};
      if (a.acc0) {
        const float4 t = *reinterpret_cast<const float4 *>(a.acc0 + cell0);
        o[0] = __fadd_rn(t.x, o[0]); o[1] = __fadd_rn(t.y, o[1]); o[2] = __fadd_rn(t.z, o[2]); o[3] = __fadd_rn(t.w, o[3]);
      }
      if (a.add0) {
        const float4 t = *reinterpret_cast<const float4 *>(a.add0 + cell0);
        o[0] = __fadd_rn(o[0], t.x); o[1] = __fadd_rn(o[1], t.y); o[2] = __fadd_rn(o[2], t.z); o[3] = __fadd_rn(o[3], t.w);
      }
      *reinterpret_cast<float4 *>(a.out0 + cell0) = make_float4(o[0], o[1], o[2], o[3]);
      m0 = fmaxf(fmaxf(o[0], o[1]), fmaxf(o[2], o[3]));
      if (a.amax0) {
        // first maximum of the four (ascending index, strict '>'), then one key
        float bv = o[0];
        int bj = 0;
        if (o[1] > bv) { bv = o[1]; bj = 1; }
        if (o[2] > bv) { bv = o[2]; bj = 2; }
        if (o[3] > bv) { bv = o[3]; bj = 3; }
        if (bv == bv) key0 = argmax_key(bv, (unsigned)(cell0 + bj));
      }
    }
    if (a.out1) {
      const float4 t = *reinterpret_cast<const float4 *>(a.add1 + cell0);
      const float4 o = make_float4(__fadd_rn(t.x, v[0]), __fadd_rn(t.y, v[1]), __fadd_rn(t.z, v[2]), __fadd_rn(t.w, v[3]));
      *reinterpret_cast<float4 *>(a.out1 + cell0) = o;
      m1 = fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w));
    }
  }
  const int lane = threadIdx.x & 31;
  if (a.max0) {
    const int e = __reduce_max_sync(0xffffffffu, enc_max_operand(m0));
    if (lane == 0) atomicMax(&s_m0, e);
  }
  if (a.max1) {
    const int e = __reduce_max_sync(0xffffffffu, enc_max_operand(m1));
    if (lane == 0) atomicMax(&s_m1, e);
  }
  if (a.amax0) {
    const unsigned hi = (unsigned)(key0 >> 32);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? (unsigned)key0 : 0u);
    if (lane == 0 && (mh | ml)) atomicMax(&s_key, ((unsigned long long)mh << 32) | ml);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (a.max0) atomicMax(a.max0, s_m0);
    if (a.max1) atomicMax(a.max1, s_m1);
    if (a.amax0 && s_key) atomicMax(a.amax0, s_key);
  }
}

// ---- root: combine the stored upward messages (findrot.cpp:637-654 and :169) -----------------------------
//   post[root]  = (((m_0 + m_1) + ...) + m_{n-1}) + unary[root]
//   fr_j        = (sum over i != j, ascending, starting from 0) + unary[root]     (written over m_j)
constexpr int kMaxRootChildren = 16;
struct RootArgs {
  float *m[kMaxRootChildren];
  int *fr_max[kMaxRootChildren];
  unsigned long long *amax;  // optional: first-maximum key of post[root]
  int n;
  const float *unary;   // may be null (root not is_detect: cannot happen, root needs is_detect)
  float *post;
  size_t N;
};

template <int NC>
__device__ __forceinline__ void root_combine_cell(const float *mv, float u, float &post, float *fr) {
  float tot = 0.0f;
#pragma unroll
  for (int j = 0; j < NC; ++j) tot = __fadd_rn(tot, mv[j]);
  post = __fadd_rn(tot, u);
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NC; ++k)
      if (k != j) s = __fadd_rn(s, mv[k]);
    fr[j] = __fadd_rn(s, u);
  }
}

// NC = number of root children (compile time, so the sums stay in registers); NV = 4: float4 path
// (N % 4 == 0, 16-byte aligned grids), NV = 1: scalar path.
template <int NC, int NV>
__global__ void __launch_bounds__(256) k_root_combine(RootArgs a) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * NV;
  float frmax[NC];
  unsigned long long key = 0;
#pragma unroll
  for (int j = 0; j < NC; ++j) frmax[j] = -INFINITY;
  if (i < a.N) {
    float u[NV], mv[NV][NC], post[NV], fr[NV][NC];
    if (NV == 4) {
      float4 t = *reinterpret_cast<const float4 *>(a.unary + i);
      u[0] = t.x; u[1 % NV] = t.y; u[2 % NV] = t.z; u[3 % NV] = t.w;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        float4 q = *reinterpret_cast<const float4 *>(a.m[j] + i);
        mv[0][j] = q.x; mv[1 % NV][j] = q.y; mv[2 % NV][j] = q.z; mv[3 % NV][j] = q.w;
      }
    } else {
      u[0] = a.unary[i];
#pragma unroll
      for (int j = 0; j < NC; ++j) mv[0][j] = a.m[j][i];
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) root_combine_cell<NC>(mv[v], u[v], post[v], fr[v]);
    if (a.amax) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (post[v] == post[v]) {
          unsigned long long k = argmax_key(post[v], (unsigned)(i + v));
          key = k > key ? k : key;
        }
    }
    if (NV == 4) {
      *reinterpret_cast<float4 *>(a.post + i) = make_float4(post[0], post[1 % NV], post[2 % NV], post[3 % NV]);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        *reinterpret_cast<float4 *>(a.m[j] + i) = make_float4(fr[0][j], fr[1 % NV][j], fr[2 % NV][j], fr[3 % NV][j]);
        frmax[j] = fmaxf(fmaxf(fr[0][j], fr[1 % NV][j]), fmaxf(fr[2 % NV][j], fr[3 % NV][j]));
      }
    } else {
      a.post[i] = post[0];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        a.m[j][i] = fr[0][j];
        frmax[j] = fr[0][j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    block_max_to(frmax[j], a.fr_max[j]);
    __syncthreads();
  }
  if (a.amax) block_key_max_to(key, a.amax);
}

// ---- root rotation-marginal (findrot.cpp:694-726) -----------------------------------------------------------
// rp = log( exp(post[v0]) + exp(post[v1]) + ... ) in list order, no max shift, 0 -> LOG_ZERO.
__global__ void k_root_marginal(const float *__restrict__ post, size_t HW, const int *__restrict__ valid, int nvalid,
                                float *__restrict__ rp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  float s = (float)exp((double)post[(size_t)valid[0] * HW + i]);
  for (int k = 1; k < nvalid; ++k) s = __fadd_rn(s, (float)exp((double)post[(size_t)valid[k] * HW + i]));
  rp[i] = log_f64(s);
}

// ---- legacy POS_GAUSSIAN path: mergeRotationsSum (objectdetect_findpos.cpp:92-116) ---------------------------------
// result[y][x] = (float) log( sum_r exp(g[r][y][x]) ).  findpos.cpp has `using namespace std`, so exp(float) is the
// FLOAT overload there: every term is rounded to fp32 before it is added (first GPU run of this path: summing the double
// exponentials instead moved 10 % of the root-posterior cells by an ulp or two).  The reference accumulates the terms in
// x87 long double (64-bit mantissa) and takes logl; here the fp32 terms are summed error-free (two-sum into a hi/lo pair,
// ~106 bits) and the logarithm is log(hi) + lo/hi, so the two agree except where the value sits within ~2^-52 of an fp32
// rounding boundary or glibc's expf (< 1 ulp, not correctly rounded) differs from the correctly rounded exponential used
// here.  A pixel whose slices are all LOG_ZERO gives log(0) = -inf, as there.
__global__ void __launch_bounds__(256) k_merge_rotations(const float *__restrict__ g, int R, size_t HW, float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
  double hi = 0.0, lo = 0.0;
  for (int r = 0; r < R; ++r) {
    const double e = (double)exp_f64(g[(size_t)r * HW + i]);
    const double s = __dadd_rn(hi, e);
    const double bb = __dsub_rn(s, hi);
    const double err = __dadd_rn(__dsub_rn(hi, __dsub_rn(s, bb)), __dsub_rn(e, bb));
    hi = s;
    lo = __dadd_rn(lo, err);
  }
  out[i] = hi > 0.0 ? (float)__dadd_rn(log(hi), lo / hi) : (float)log(hi);
}
// addGrid2 (multi_array_op.hpp:117-130): a += b
__global__ void __launch_bounds__(256) k_add2(float *__restrict__ a, const float *__restrict__ b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = __fadd_rn(a[i], b[i]);
}

// ---- readout --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_argmax(const float *__restrict__ g, size_t n, unsigned long long *dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = ((uintptr_t)g % 16 == 0) ? n / 4 : 0;
  // running (value, first index) per thread; a strictly greater value replaces, so the earliest index of the
  // thread's maximum is kept as long as the thread visits indices in ascending order
  float bv = -INFINITY;
  unsigned bi = 0xffffffffu;
  const float4 *g4 = reinterpret_cast<const float4 *>(g);
  for (size_t j = i; j < n4; j += stride) {
    float4 v = __ldg(g4 + j);
    unsigned b = (unsigned)(j * 4);
    if (v.x > bv) { bv = v.x; bi = b; }
    if (v.y > bv) { bv = v.y; bi = b + 1; }
    if (v.z > bv) { bv = v.z; bi = b + 2; }
    if (v.w > bv) { bv = v.w; bi = b + 3; }
  }
  for (size_t j = n4 * 4 + i; j < n; j += stride) {
    float v = g[j];
    if (v > bv) { bv = v; bi = (unsigned)j; }
  }
  unsigned long long best = bi == 0xffffffffu ? 0ull : argmax_key(bv, bi);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t > best ? t : best;
  }
  __shared__ unsigned long long s_best[8];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_best[w] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) best = s_best[k] > best ? s_best[k] : best;
    atomicMax(dst, best);
  }
}

// findLocalMax candidate test (objectdetect_aux.cpp:203-228): no 8-neighbour greater, strictly greater than
// both dim-0 neighbours (no wrap).  Survivors are appended as (score, scan-order key) pairs; scan-order key =
// (d0 * W + x) * H + y, the order the reference visits cells.
struct Cand {
  float score;
  unsigned key;
};

__global__ void __launch_bounds__(256) k_local_max(const float *__restrict__ g, int D0, int H, int W, Cand *out,
                                                   unsigned cap, unsigned *count) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  int s = blockIdx.z;
  if (x >= W) return;
  size_t HW = (size_t)H * W;
  const float *p = g + (size_t)s * HW;
  float c = p[(size_t)y * W + x];
  bool ok = true;
  for (int dy = -1; dy <= 1; ++dy) {
    int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      int xx = x + dx;
      if (xx < 0 || xx >= W || (dx == 0 && dy == 0)) continue;
      if (p[(size_t)yy * W + xx] > c) ok = false;
    }
  }
  if (ok && s > 0) ok = p[(size_t)y * W + x - HW] < c;
  if (ok && s < D0 - 1) ok = p[(size_t)y * W + x + HW] < c;
  if (ok) {
    unsigned slot = atomicAdd(count, 1u);
    if (slot < cap) {
      Cand cd;
      cd.score = c;
      cd.key = (unsigned)(((size_t)s * W + x) * H + y);
      out[slot] = cd;
    }
  }
}


// ---- top-K of the local-maximum candidates on the device (objectdetect_aux.cpp:233-258) -----------------------------
// The reference keeps the K highest-scoring candidates (std::sort on the fp32 score, ties unspecified).  Here the
// order is made total with the scan-order key: composite = (ordered(score) << 32) | ~scan_key, larger is better, all
// composites are distinct.  A 4-pass radix select (16-bit digits, most significant first) finds the K-th largest
// composite without sorting; a final pass compacts everything >= it.  Everything stays on the stream: no host
// round trip until the K winners are copied back.
struct TopKState {
  unsigned long long prefix;   // digits decided so far (high bits), rest 0
  unsigned long long mask;     // which bits of `prefix` are decided
  unsigned remaining;          // how many more items are needed from inside the current prefix bucket
  unsigned count;              // clamp(candidates, cap)
  unsigned k;                  // min(K, count)
  unsigned out_count;          // compaction cursor
};

__device__ __forceinline__ unsigned long long cand_composite(const Cand &c) {
  float v = c.score;
  if (v == 0.0f) v = 0.0f;
  unsigned e = (unsigned)enc_f(v) ^ 0x80000000u;
  return ((unsigned long long)e << 32) | (unsigned)(~c.key);
}

__global__ void k_topk_init(TopKState *st, const unsigned *count, unsigned cap, unsigned K, unsigned *hist) {
  for (int i = threadIdx.x; i < 65536; i += blockDim.x) hist[i] = 0;
  if (threadIdx.x == 0) {
    unsigned n = min(*count, cap);
    st->prefix = 0; st->mask = 0;
    st->count = n;
    st->k = min(K, n);
    st->remaining = st->k;
    st->out_count = 0;
  }
}

__global__ void __launch_bounds__(256) k_topk_hist(const Cand *cand, const TopKState *st, int shift, unsigned *hist) {
  const unsigned n = st->count;
  const unsigned long long prefix = st->prefix, mask = st->mask;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned long long ck = cand_composite(cand[i]);
    if ((ck & mask) == prefix) atomicAdd(&hist[(unsigned)(ck >> shift) & 0xffffu], 1u);
  }
}

// One block: walk the 65536 bins from the top until `remaining` items are covered; fix that digit; clear the bins.
__global__ void __launch_bounds__(1024) k_topk_scan(TopKState *st, int shift, unsigned *hist) {
  __shared__ unsigned s_sum[1024];
  __shared__ unsigned s_above;
  const int t = threadIdx.x;
  // thread t owns bins [65535 - 64t - 63, 65535 - 64t], i.e. chunk t counted from the top
  unsigned local = 0;
  for (int j = 0; j < 64; ++j) local += hist[65535 - (t * 64 + j)];
  s_sum[t] = local;
  __syncthreads();
  if (t == 0) {
    unsigned need = st->remaining, above = 0;
    int chunk = 0;
    while (chunk < 1023 && above + s_sum[chunk] < need) above += s_sum[chunk++];
    int bin = 65535 - chunk * 64;
    for (int j = 0; j < 64; ++j, --bin) {
      unsigned h = hist[bin];
      if (above + h >= need || j == 63) break;
      above += h;
    }
    if (st->k == 0) bin = 0;
    st->prefix |= (unsigned long long)(unsigned)bin << shift;
    st->mask |= 0xffffull << shift;
    st->remaining = need - above;
    s_above = above;
  }
  __syncthreads();
  for (int j = 0; j < 64; ++j) hist[t * 64 + j] = 0;
}

__global__ void __launch_bounds__(256) k_topk_compact(const Cand *cand, TopKState *st, Cand *out, unsigned out_cap) {
  const unsigned n = st->count;
  if (st->k == 0) return;
  const unsigned long long thr = st->prefix;  // all four digits decided: the K-th largest composite
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Cand c = cand[i];
    if (cand_composite(c) >= thr) {
      unsigned slot = atomicAdd(&st->out_count, 1u);
      if (slot < out_cap) out[slot] = c;
    }
  }
}

}  // namespace psk
