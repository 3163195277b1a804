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

#include <cuda_runtime.h>
#include <stdint.h>

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

// (float)exp((double)x).  Below -104 the double result is < 2^-150 and narrows to +0 exactly.
__device__ __forceinline__ float exp_f64(float x) {
  if (x < -104.0f) return 0.0f;
  return (float)exp((double)x);
}
// computeLogGrid cell (multi_array_op.hpp:162-165)
__device__ __forceinline__ float log_f64(float d) {
  return d == 0.0f ? kLogZero : (float)log((double)d);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide max folded into *dst (encoded) with one atomic per block. All threads must call.
__device__ __forceinline__ void block_max_to(float v, int *dst) {
  __shared__ float s_part[32];
  v = warp_max(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
  if (lane == 0) s_part[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? s_part[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) atomicMax(dst, enc_f(t));
  }
}

// ---- pointwise sweeps ------------------------------------------------------------------------------

__global__ void k_fill(float *__restrict__ p, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void k_set_int(int *p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// clip_scores_fill (objectdetect_aux.hpp:42-59) + computeLogGrid (multi_array_op.hpp:154-167)
__global__ void k_prepare_unary(float *__restrict__ p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = p[i];
    if (v < 0.0f) v = (float)0.0001;
    p[i] = log_f64(v);
  }
}

// getMinMax (multi_array_op.hpp:61-77), max only
__global__ void k_grid_max(const float *__restrict__ p, size_t n, int *dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float m = -INFINITY;
  for (; i < n; i += stride) m = fmaxf(m, p[i]);
  block_max_to(m, dst);
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

// ---- message stage 1: shift + exp + circular rotation filter ------------------------------------------
// findrot.cpp:339-420.  One thread per pixel; the R shifted/exponentiated values of the pixel live in a
// private shared-memory column, then every output rotation is a sequential dot product over the taps.
struct RotArgs {
  const float *in;      // [R][H][W] child belief (log domain)
  float *out;           // [R][H][W] rotation-filtered probabilities
  const int *xin;       // [R][W] source x or -1
  const int *yin;       // [R][H] source y or -1
  const float *taps;    // [len]
  const int *max_enc;   // encoded max of `in`
  int R, H, W;
  int shift;            // rot_mean_idx
  int mode;             // 0 copy, 1 filter, 2 zero
  int len;
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
  for (int r = 0; r < R; ++r) {
    const float *s = in + (size_t)r * HW;
    float v = 0.0f;
    if (m.x >= 0) {
      v = __ldg(&s[m.x]);
      if (v == 0.0f && m.y >= 0) v = __ldg(&s[m.y]);
    }
    out[r * eplane + o] = v;
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
  for (int r = 0; r < R; ++r) out[r * eplane + o] = bilinear_at(in + r * HW, H, W, W, x1, y1);
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

// ---- message stage 3: read back, log, +M, shift to the parent frame, combine ------------------------------
// findrot.cpp:431-448 plus the addGrid2 calls that consume the message (findrot.cpp:637-654, :201, :221-222).
struct EpiArgs {
  const float *src;     // diag: [R][H][W] filtered probabilities; general: eigen-frame [R][EH][EP]
  int general;
  Affine T34;
  int EH, EW, EP;
  const int *xout, *yout;  // [R][W], [R][H]
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
};

__global__ void __launch_bounds__(256) k_epilogue(EpiArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int r = blockIdx.z;
  float m0 = -INFINITY, m1 = -INFINITY;
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
}

// ---- root: combine the stored upward messages (findrot.cpp:637-654 and :169) -----------------------------
//   post[root]  = (((m_0 + m_1) + ...) + m_{n-1}) + unary[root]
//   fr_j        = (sum over i != j, ascending, starting from 0) + unary[root]     (written over m_j)
constexpr int kMaxRootChildren = 16;
struct RootArgs {
  float *m[kMaxRootChildren];
  int *fr_max[kMaxRootChildren];
  int n;
  const float *unary;   // may be null (root not is_detect: cannot happen, root needs is_detect)
  float *post;
  size_t N;
};

__global__ void __launch_bounds__(256) k_root_combine(RootArgs a) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  float mv[kMaxRootChildren];
  float frv[kMaxRootChildren];
#pragma unroll
  for (int j = 0; j < kMaxRootChildren; ++j) frv[j] = -INFINITY;
  if (i < a.N) {
    float u = a.unary[i];
    float tot = 0.0f;
#pragma unroll
    for (int j = 0; j < kMaxRootChildren; ++j)
      if (j < a.n) {
        mv[j] = a.m[j][i];
        tot = __fadd_rn(tot, mv[j]);
      }
    a.post[i] = __fadd_rn(tot, u);
#pragma unroll
    for (int j = 0; j < kMaxRootChildren; ++j)
      if (j < a.n) {
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxRootChildren; ++k)
          if (k < a.n && k != j) s = __fadd_rn(s, mv[k]);
        s = __fadd_rn(s, u);
        frv[j] = s;
        a.m[j][i] = s;
      }
  }
  for (int j = 0; j < a.n; ++j) {
    block_max_to(frv[j], a.fr_max[j]);
    __syncthreads();
  }
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

// ---- readout --------------------------------------------------------------------------------------------------
// First maximum in flat order (findrot.cpp:261-277): packed key = (enc(value) << 32) | ~index, max over keys
// picks the largest value and, among equals, the smallest index.  -0.0 and +0.0 compare equal in the reference,
// so -0.0 is canonicalised to +0.0 before encoding.
__device__ __forceinline__ unsigned long long argmax_key(float v, unsigned idx) {
  if (v == 0.0f) v = 0.0f;
  unsigned e = (unsigned)enc_f(v) ^ 0x80000000u;  // order-preserving unsigned
  return ((unsigned long long)e << 32) | (unsigned)(~idx);
}

__global__ void __launch_bounds__(256) k_argmax(const float *__restrict__ g, size_t n, unsigned long long *dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long best = 0;
  for (; i < n; i += stride) {
    float v = g[i];
    if (v != v) continue;  // NaN never wins a '>' comparison
    unsigned long long k = argmax_key(v, (unsigned)i);
    best = k > best ? k : best;
  }
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

}  // namespace psk
