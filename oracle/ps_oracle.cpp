// ps_oracle.cpp -- CPU ORACLE for the pictorial-structures inference hot path of partapp.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product (partapp_b200/,
// libpsinfer.so) never links, imports or calls anything in oracle/.
//
// PARITY STATUS: *pinned against the reference's own code, compiled here; unpinned only for BLAS order and libm*.
// The reference ships no golden vectors, known-answer tests or fixtures for this path, and its
// build needs Boost, Qt4, a BLAS, protoc output and MATLAB's libmat, none of which this image has.  Its SOURCES for
// the path nevertheless compile UNMODIFIED from /root/reference once the containers and interfaces they include exist:
// oracle/ref_shim/ holds stand-ins written for this repo (a dense boost::multi_array with views, an eager dense uBLAS,
// boost::lambda's bind / placeholders, QString, a Netlib-order cblas_sdot, the accessors of the protoc-generated
// messages, do-nothing libmat / libAnnotation / detector interfaces), and `make -C oracle ref` builds
//   oracle/_ref/libps_ref_core.so     libMultiArray/multi_array_{op,transform,filter}.hpp, libBoostMath/{boost_math,
//                                     homogeneous_coord}.cpp, libPartApp/partapp_aux.hpp, libPictStruct/objectdetect_aux.hpp
//   oracle/_ref/libps_ref_drivers.so  libPictStruct/objectdetect_findrot.cpp (computeRotJointMarginal,
//                                     computePartMarginals, computeRootPosteriorRot), objectdetect_aux.cpp
//                                     (findLocalMax, loadJoints) and objectdetect_icps.cpp (the conditioning adds)
// behind the C entry points of oracle/ref_core.cpp and oracle/ref_drivers.cpp.  tests/test_oracle_vs_ref.py holds this
// file to that code BIT FOR BIT -- grid primitives, single messages, whole inferences (argmax records, every marginal
// cell, root posterior, local-maximum lists, the in-place masking of the unaries), local maxima with the top-K cut, and
// the flip transform of the joints -- live where the reference tree exists and everywhere through the recorded outputs
// tests/golden/ref_core.npz and ref_drivers.npz (tests/golden/make_ref_golden.py).
// NOT pinned: the arithmetic that lives outside the reference tree --
//   * cblas_sdot  -> Netlib order: sequential ascending-index fp32 multiply-then-add, no FMA (the stand-in BLAS and this
//                    file share the convention; the authors' libblas is unknown);
//   * exp / log   -> evaluated in double by libm and narrowed to float (this machine's glibc);
// and the file-reading helpers (getRotParams, getPosParams, addLoadDPMScore's loader, PartApp::loadScoreGrid):
// those remain line-by-line restatements (every function below names the reference file:line it follows; paths are
// relative to /root/reference/src/libs).
// Build: g++ -O3 -ffp-contract=off (no -ffast-math, no -mfma) -- see oracle/Makefile.
//
// Loop nests are re-ordered where that cannot change any result (each output still sees its own
// taps in ascending order, in fp32, multiply then add), so the compiler may vectorise across
// independent outputs.

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace {

// libBoostMath/boost_math.h:23
const double LOG_ZERO = -1e6;

// libBoostMath/boost_math.hpp:35-37
inline int bm_round(double val) { return (int)floor(val + 0.5); }

// ---------------------------------------------------------------------------------------------
// 3x3 homogeneous matrices (libBoostMath/homogeneous_coord.cpp:38-156).  uBLAS prod() of small
// dense matrices is the plain triple loop: t = 0; t += a(i,k)*b(k,j) for k ascending.
// ---------------------------------------------------------------------------------------------
struct Mat3 {
  double m[3][3];
};

Mat3 mat3_zero() {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = 0.0;
  return r;
}

Mat3 mat3_identity() {
  Mat3 r = mat3_zero();
  r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
  return r;
}

Mat3 mat3_prod(const Mat3 &a, const Mat3 &b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = 0.0;
      for (int k = 0; k < 3; ++k) t += a.m[i][k] * b.m[k][j];
      r.m[i][j] = t;
    }
  return r;
}

void mat3_vec(const Mat3 &a, const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) {
    double t = 0.0;
    for (int k = 0; k < 3; ++k) t += a.m[i][k] * v[k];
    out[i] = t;
  }
}

// homogeneous_coord.cpp:38-47
Mat3 hc_homogeneous(const double R[2][2], double dx, double dy) {
  Mat3 T = mat3_zero();
  T.m[0][0] = R[0][0]; T.m[0][1] = R[0][1];
  T.m[1][0] = R[1][0]; T.m[1][1] = R[1][1];
  T.m[0][2] = dx;
  T.m[1][2] = dy;
  T.m[2][2] = 1;
  return T;
}

// homogeneous_coord.cpp:49-69
Mat3 hc_inverse(const Mat3 &T) {
  Mat3 inv;
  double D = T.m[0][0] * T.m[1][1] - T.m[1][0] * T.m[0][1];
  inv.m[0][0] = T.m[1][1] / D;
  inv.m[0][1] = -T.m[0][1] / D;
  inv.m[1][0] = -T.m[1][0] / D;
  inv.m[1][1] = T.m[0][0] / D;
  // t = -prod(inv[0:2,0:2], T[0:2,2:3])
  for (int i = 0; i < 2; ++i) {
    double t = 0.0;
    for (int k = 0; k < 2; ++k) t += inv.m[i][k] * T.m[k][2];
    inv.m[i][2] = -t;
  }
  inv.m[2][0] = 0;
  inv.m[2][1] = 0;
  inv.m[2][2] = 1;
  return inv;
}

// homogeneous_coord.cpp:71-78
Mat3 hc_scaling(double scale) {
  Mat3 R = mat3_identity();
  R.m[0][0] = scale;
  R.m[1][1] = scale;
  return R;
}

// homogeneous_coord.cpp:81-93
Mat3 hc_rotation(double rad) {
  Mat3 R = mat3_zero();
  double ca = cos(rad);
  double sa = sin(rad);
  R.m[0][0] = ca;
  R.m[0][1] = -sa;
  R.m[1][0] = sa;
  R.m[1][1] = ca;
  R.m[2][2] = 1;
  return R;
}

// homogeneous_coord.cpp:95-103
Mat3 hc_translation(double dx, double dy) {
  Mat3 T = mat3_identity();
  T.m[0][2] = dx;
  T.m[1][2] = dy;
  return T;
}

// homogeneous_coord.h:72-81
inline void hc_map_point(const Mat3 &M, double x, double y, double &ox, double &oy) {
  ox = M.m[0][0] * x + M.m[0][1] * y + M.m[0][2];
  oy = M.m[1][0] * x + M.m[1][1] * y + M.m[1][2];
}

// homogeneous_coord.cpp:139-156
void hc_transformed_bbox(const Mat3 &T21, int in_width, int in_height, double &minx, double &miny,
                         double &maxx, double &maxy) {
  double pts[4][3] = {{0, 0, 1},
                      {(double)(in_width - 1), 0, 1},
                      {0, (double)(in_height - 1), 1},
                      {(double)(in_width - 1), (double)(in_height - 1), 1}};
  double c[4][3];
  for (int i = 0; i < 4; ++i) mat3_vec(T21, pts[i], c[i]);
  minx = maxx = c[0][0];
  miny = maxy = c[0][1];
  for (int i = 1; i < 4; ++i) {
    minx = std::min(minx, c[i][0]);
    maxx = std::max(maxx, c[i][0]);
    miny = std::min(miny, c[i][1]);
    maxy = std::max(maxy, c[i][1]);
  }
}

// ---------------------------------------------------------------------------------------------
// libBoostMath/boost_math.cpp
// ---------------------------------------------------------------------------------------------

// boost_math.cpp:40-99.  V columns are eigenvectors, smallest eigenvalue first.
void eig2d(const double M[2][2], double V[2][2], double E[2][2]) {
  double m11 = M[0][0], m12 = M[0][1], m22 = M[1][1];
  double e1, e2, v11, v21;
  if (m12 != 0) {
    double sqrtD = sqrt((m11 - m22) * (m11 - m22) + 4 * m12 * m12);
    e1 = 0.5 * (m11 + m22 - sqrtD);
    e2 = 0.5 * (m11 + m22 + sqrtD);
    v11 = 0.5 * (m11 - m22 - sqrtD) / m12;
    v21 = 1;
  } else {
    if (m11 < m22) {
      e1 = m11; e2 = m22; v11 = 1; v21 = 0;
    } else {
      e1 = m22; e2 = m11; v11 = 0; v21 = 1;
    }
  }
  double norm_v1 = sqrt(v11 * v11 + v21 * v21);
  v11 /= norm_v1;
  v21 /= norm_v1;
  double v12 = -v21;
  double v22 = v11;
  E[0][0] = e1; E[0][1] = 0; E[1][0] = 0; E[1][1] = e2;
  V[0][0] = v11; V[1][0] = v21; V[0][1] = v12; V[1][1] = v22;
}

// boost_math.cpp:104-117 (bNormalize is false on every call of this path: findrot.cpp:370-371)
std::vector<double> get_gaussian_filter(double sigma) {
  int ksize = (int)floor(3 * sigma + 0.5);
  std::vector<double> f(2 * ksize + 1);
  f[ksize] = 1.0;
  for (int i = 1; i <= ksize; ++i) {
    f[ksize + i] = exp(-i * i / (2 * sigma * sigma));
    f[ksize - i] = f[ksize + i];
  }
  return f;
}

// ---------------------------------------------------------------------------------------------
// libPartApp/partapp_aux.hpp
// ---------------------------------------------------------------------------------------------
struct ExpParam {
  int num_rotation_steps;
  float min_part_rotation, max_part_rotation;
  int num_scale_steps;
  float min_object_scale, max_object_scale;
  float strip_border_detections;
  int roi_save_num_samples;
};

// partapp_aux.hpp:45-58
double value_from_index(double minval, double maxval, double num_steps, int idx) {
  if (minval == maxval) return minval;
  double step_size = (maxval - minval) / num_steps;
  return minval + step_size * (0.5 + idx);
}

// partapp_aux.hpp:25-43
int index_from_value(double minval, double maxval, double num_steps, double val) {
  if (minval == maxval) return 0;
  if (!(val >= minval && val < maxval)) return -1;  // reference asserts
  double step_size = (maxval - minval) / num_steps;
  return (int)(unsigned)floor((val - minval) / step_size);
}

// partapp_aux.hpp:123-129, :86-92, :94-100
double rot_from_index(const ExpParam &ep, int idx) {
  return value_from_index(ep.min_part_rotation, ep.max_part_rotation, ep.num_rotation_steps, idx);
}
double scale_from_index(const ExpParam &ep, int idx) {
  return value_from_index(ep.min_object_scale, ep.max_object_scale, ep.num_scale_steps, idx);
}
int index_from_rot(const ExpParam &ep, double rot) {
  return index_from_value(ep.min_part_rotation, ep.max_part_rotation, ep.num_rotation_steps, rot);
}

// ---------------------------------------------------------------------------------------------
// libMultiArray/multi_array_op.hpp -- flat pointwise sweeps
// ---------------------------------------------------------------------------------------------

// multi_array_op.hpp:61-77 (only the max is used by the path)
float grid_max(const float *p, size_t n) {
  float maxval = -std::numeric_limits<float>::infinity();
  for (size_t i = 0; i < n; ++i)
    if (p[i] > maxval) maxval = p[i];
  return maxval;
}

// multi_array_op.hpp:99-107
void add_grid1(float *p, size_t n, float num) {
  for (size_t i = 0; i < n; ++i) p[i] += num;
}

// multi_array_op.hpp:109-122
void add_grid2(float *a, const float *b, size_t n) {
  for (size_t i = 0; i < n; ++i) a[i] += b[i];
}

// multi_array_op.hpp:154-167.  log() is the double libm routine, narrowed.
void compute_log_grid(float *p, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    if (p[i] == 0)
      p[i] = (float)LOG_ZERO;
    else
      p[i] = (float)log((double)p[i]);
  }
}

// multi_array_op.hpp:170-180.  exp() is the double libm routine, narrowed.
void compute_exp_grid(float *p, size_t n) {
  for (size_t i = 0; i < n; ++i) p[i] = (float)exp((double)p[i]);
}

// ---------------------------------------------------------------------------------------------
// libMultiArray/multi_array_transform.hpp:123-242  transform_grid_helper
// ---------------------------------------------------------------------------------------------
enum { TM_NEAREST = 0, TM_BILINEAR = 1, TM_DIRECT = 2 };

inline bool check_bounds(int v, int lo, int hi) { return v >= lo && v < hi; }  // libMisc/misc.hpp:32-36

// Non-adaptive form used by transform_grid_fixed_size (:245-257): T23 = T32 = I.
// Adaptive form used by transform_grid_resize (:285-305): T23 = Trans(minx,miny); out grid is
// allocated by the caller to ceil(maxx-minx) x ceil(maxy-miny).
void transform_grid_helper(const float *in, int in_h, int in_w, float *out, int out_h, int out_w,
                           const Mat3 &T21, Mat3 &T23, float default_value, int method,
                           bool adaptive) {
  Mat3 T12 = hc_inverse(T21);
  Mat3 T32;
  if (adaptive) {
    double minx, miny, maxx, maxy;
    hc_transformed_bbox(T21, in_w, in_h, minx, miny, maxx, maxy);
    assert(out_w == (int)ceil(maxx - minx) && out_h == (int)ceil(maxy - miny));
    T23 = hc_translation(minx, miny);
    T32 = hc_translation(-minx, -miny);
  } else {
    T32 = mat3_identity();
    T23 = mat3_identity();
  }

  if (method == TM_DIRECT) {
    for (size_t i = 0; i < (size_t)out_h * out_w; ++i) out[i] = default_value;
    Mat3 T31 = mat3_prod(T32, T21);
    // x1 outer, y1 inner: later writers overwrite earlier ones (:176-190)
    for (int x1 = 0; x1 < in_w; ++x1)
      for (int y1 = 0; y1 < in_h; ++y1) {
        float v = in[(size_t)y1 * in_w + x1];
        if (v != default_value) {
          double x3, y3;
          hc_map_point(T31, (double)x1, (double)y1, x3, y3);
          int ix3 = (int)floor(x3 + 0.5);
          int iy3 = (int)floor(y3 + 0.5);
          if (check_bounds(ix3, 0, out_w) && check_bounds(iy3, 0, out_h))
            out[(size_t)iy3 * out_w + ix3] = v;
        }
      }
  } else {
    Mat3 T13 = mat3_prod(T12, T23);
    const float eps10 = 10 * std::numeric_limits<float>::epsilon();
    for (int y3 = 0; y3 < out_h; ++y3)
      for (int x3 = 0; x3 < out_w; ++x3) {
        float *o = &out[(size_t)y3 * out_w + x3];
        *o = default_value;
        double x1, y1;
        hc_map_point(T13, (double)x3, (double)y3, x1, y1);
        int ix1, iy1;
        if (method == TM_BILINEAR) {
          ix1 = (int)floor(x1);
          iy1 = (int)floor(y1);
        } else {
          ix1 = (int)floor(x1 + 0.5);
          iy1 = (int)floor(y1 + 0.5);
        }
        if (check_bounds(ix1, 0, in_w) && check_bounds(iy1, 0, in_h)) {
          if (method == TM_NEAREST) {
            *o = in[(size_t)iy1 * in_w + ix1];
          } else {
            float a = x1 - ix1;
            float b = y1 - iy1;
            if (a < eps10 && b < eps10)
              *o = in[(size_t)iy1 * in_w + ix1];
            else if (ix1 < in_w - 1 && iy1 < in_h - 1) {
              const float *p = &in[(size_t)iy1 * in_w + ix1];
              *o = (1.0f - b) * (1.0f - a) * p[0] + (1.0f - b) * a * p[1] +
                   b * (1.0f - a) * p[in_w] + b * a * p[in_w + 1];
            }
          }
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// libMultiArray/multi_array_filter.hpp
// ---------------------------------------------------------------------------------------------

// multi_array_filter.hpp:212-321  gaussFilterDiag2d.  Each output is
// cblas_sdot(len, in+n1, stride, f+(n1-(c-n)), 1) over the window clipped to the grid:
// Netlib order = ascending index, fp32 multiply then add, starting from 0.
void gauss_filter_diag2d(const float *in, float *out, int h, int w, double c00, double c11) {
  double sigma_x = sqrt(c00);
  double sigma_y = sqrt(c11);
  std::vector<double> _fx = get_gaussian_filter(sigma_x);
  std::vector<double> _fy = get_gaussian_filter(sigma_y);
  assert(_fx.size() < 1000 && _fy.size() < 1000);  // F_SIZE, :238-244
  std::vector<float> fx(_fx.size()), fy(_fy.size());
  for (size_t i = 0; i < _fx.size(); ++i) fx[i] = (float)_fx[i];
  for (size_t i = 0; i < _fy.size(); ++i) fy[i] = (float)_fy[i];
  int nx = ((int)fx.size() - 1) / 2;
  int ny = ((int)fy.size() - 1) / 2;

  std::vector<float> sm((size_t)h * w);
  // x direction (:277-288): out[y][x] = sum_{j=n1..n2} in[y][j] * fx[j-(x-nx)]
  for (int y = 0; y < h; ++y) {
    const float *row = in + (size_t)y * w;
    float *acc = sm.data() + (size_t)y * w;
    for (int x = 0; x < w; ++x) acc[x] = 0.0f;
    for (int k = 0; k < (int)fx.size(); ++k) {
      // input column j = x + k - nx must lie in [0, w)
      int x0 = std::max(0, nx - k);
      int x1 = std::min(w - 1, w - 1 + nx - k);
      float fk = fx[k];
      for (int x = x0; x <= x1; ++x) acc[x] += row[x + k - nx] * fk;
    }
  }
  // y direction (:304-317): out[y][x] = sum_{j=n1..n2} sm[j][x] * fy[j-(y-ny)]
  for (int y = 0; y < h; ++y) {
    float *acc = out + (size_t)y * w;
    for (int x = 0; x < w; ++x) acc[x] = 0.0f;
    int n1 = std::max(0, y - ny);
    int n2 = std::min(h - 1, y + ny);
    for (int j = n1; j <= n2; ++j) {
      const float *row = sm.data() + (size_t)j * w;
      float fk = fy[j - (y - ny)];
      for (int x = 0; x < w; ++x) acc[x] += row[x] * fk;
    }
  }
}

// multi_array_filter.hpp:335-369 gaussFilter2dOffset
void gauss_filter_2d_offset(const float *in, float *out, int h, int w, const double C[2][2], double offset0,
                            double offset1, bool is_sparse) {
  const float PADDING_VALUE = 0;
  double V[2][2], E[2][2];
  eig2d(C, V, E);
  double Vt[2][2] = {{V[0][0], V[1][0]}, {V[0][1], V[1][1]}};
  Mat3 T21 = hc_homogeneous(Vt, 0, 0);
  // transform_grid_resize (transform.hpp:285-305)
  double minx, miny, maxx, maxy;
  hc_transformed_bbox(T21, w, h, minx, miny, maxx, maxy);
  int ow = (int)ceil(maxx - minx);
  int oh = (int)ceil(maxy - miny);
  std::vector<float> tr((size_t)oh * ow), trs((size_t)oh * ow);
  Mat3 T23;
  transform_grid_helper(in, h, w, tr.data(), oh, ow, T21, T23, PADDING_VALUE,
                        is_sparse ? TM_DIRECT : TM_BILINEAR, true);
  gauss_filter_diag2d(tr.data(), trs.data(), oh, ow, E[0][0], E[1][1]);
  Mat3 T42 = hc_homogeneous(V, -offset0, -offset1);
  Mat3 T43 = mat3_prod(T42, T23);
  Mat3 dummy;
  transform_grid_helper(trs.data(), oh, ow, out, h, w, T43, dummy, PADDING_VALUE, TM_BILINEAR,
                        false);
}

// multi_array_filter.hpp:375-388 gaussFilter2d: the diagonal shortcut, else gaussFilter2dOffset with a zero offset
void gauss_filter_2d(const float *in, float *out, int h, int w, const double C[2][2],
                     bool is_sparse) {
  bool is_diag = (C[0][1] == 0 && C[1][0] == 0);
  if (is_diag) {
    gauss_filter_diag2d(in, out, h, w, C[0][0], C[1][1]);
    return;
  }
  gauss_filter_2d_offset(in, out, h, w, C, 0.0, 0.0, is_sparse);  // boost_math::double_zero_vector(2), :385
}

// libPictStruct/objectdetect_findpos.cpp:64-89 computePosJointMarginal (the legacy POS_GAUSSIAN message).
// Both grids are [H][W]; log_prob_child is rewritten with log(exp(child)) like the reference does (:76, :88).
void compute_pos_joint_marginal(float *log_prob_child, float *log_prob_parent, int H, int W, const double offset_in[2],
                                const double C_in[2][2], double scale, bool is_sparse) {
  assert(scale > 0);
  const size_t n = (size_t)H * W;
  const double offset[2] = {offset_in[0] * scale, offset_in[1] * scale};                     // offset *= scale
  const double s2 = scale * scale;                                                           // C *= square(scale)
  const double C[2][2] = {{C_in[0][0] * s2, C_in[0][1] * s2}, {C_in[1][0] * s2, C_in[1][1] * s2}};
  for (size_t i = 0; i < n; ++i) log_prob_child[i] = (float)exp((double)log_prob_child[i]);  // computeExpGrid
  gauss_filter_2d_offset(log_prob_child, log_prob_parent, H, W, C, offset[0], offset[1], is_sparse);
  compute_log_grid(log_prob_parent, n);
  compute_log_grid(log_prob_child, n);
}

// ---------------------------------------------------------------------------------------------
// libPictStruct/objectdetect_findrot.cpp:292-456  computeRotJointMarginal
// ---------------------------------------------------------------------------------------------
void compute_rot_joint_marginal(const ExpParam &ep, const float *log_prob_child,
                                float *log_prob_parent, int R, int H, int W,
                                const double _offset_c_10[2], const double _offset_p_01[2],
                                const double C[2][2], double rot_mean, double rot_sigma,
                                double scale, bool is_sparse, float *dbg_exp, float *dbg_rot,
                                float *dbg_spatial) {
  const size_t HW = (size_t)H * W;
  const size_t N = (size_t)R * HW;

  float in_M = grid_max(log_prob_child, N);  // :301-302

  double offset_c_10[3] = {_offset_c_10[0], _offset_c_10[1], 0};  // hc::get_vector, :308-309
  double offset_p_01[3] = {_offset_p_01[0], _offset_p_01[1], 0};

  // :319-326.  The subtraction and the division are float arithmetic (ExpParam floats, int count).
  double rot_step_size = (ep.max_part_rotation - ep.min_part_rotation) / ep.num_rotation_steps;
  rot_step_size *= M_PI / 180.0;
  double rot_sigma_idx = rot_sigma / rot_step_size;
  int rot_mean_idx = bm_round(-rot_mean / rot_step_size);

  std::vector<float> log_joint_10(N, (float)LOG_ZERO);  // :339-343
  std::vector<float> log_joint_01(N, (float)LOG_ZERO);

  for (int rotidx = 0; rotidx < R; ++rotidx) {  // :345-359
    int rotidx_out = rotidx + rot_mean_idx;
    if (rotidx_out >= 0 && rotidx_out < R) {
      float alpha = rot_from_index(ep, rotidx) * M_PI / 180.0;
      Mat3 Tgc = mat3_prod(hc_rotation(alpha), hc_scaling(scale));
      double offset_g_10[3];
      mat3_vec(Tgc, offset_c_10, offset_g_10);
      Mat3 Tjoint = hc_translation(offset_g_10[0], offset_g_10[1]);
      Mat3 dummy;
      transform_grid_helper(log_prob_child + rotidx * HW, H, W, log_joint_10.data() + rotidx_out * HW,
                            H, W, Tjoint, dummy, (float)LOG_ZERO, TM_NEAREST, false);
    }
  }

  add_grid1(log_joint_10.data(), N, -in_M);   // :362
  compute_exp_grid(log_joint_10.data(), N);   // :365
  if (dbg_exp) memcpy(dbg_exp, log_joint_10.data(), N * sizeof(float));

  std::vector<float> rot_filter_result(N, 0.0f);  // :373

  if (rot_sigma > 0) {  // :377-417
    std::vector<double> f_rot = get_gaussian_filter(rot_sigma_idx);
    int firstidx = 0;
    int f_rot_size = (int)f_rot.size();
    if (f_rot_size >= R) {  // clip kernel tails, :385-390
      int crot = (int)f_rot_size / 2;
      f_rot_size = (R % 2 == 1) ? R - 2 : R - 1;
      firstidx = crot - (int)f_rot_size / 2;
    }
    int lastidx = firstidx + f_rot_size;
    std::vector<float> f(f_rot_size);
    for (int idx = firstidx; idx < lastidx; ++idx) f[idx - firstidx] = (float)f_rot[idx];
    int nx = (f_rot_size - 1) / 2;
    // grid_filter_1d_blas_wraparound (filter.hpp:116-155) per (y,x) column:
    // out[i] = sdot(len, padded + i, f), padded[nx + j] = in[j] with circular wrap.
    for (int i = 0; i < R; ++i) {
      float *acc = rot_filter_result.data() + i * HW;
      for (int k = 0; k < f_rot_size; ++k) {
        int src = ((i + k - nx) % R + R) % R;
        const float *s = log_joint_10.data() + src * HW;
        float fk = f[k];
        for (size_t p = 0; p < HW; ++p) acc[p] += s[p] * fk;
      }
    }
  } else if (rot_sigma == 0) {
    rot_filter_result = log_joint_10;  // :418-420
  }
  if (dbg_rot) memcpy(dbg_rot, rot_filter_result.data(), N * sizeof(float));

  // :423-429
  double scaleC[2][2];
  double s2 = scale * scale;  // square(scale)
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) scaleC[i][j] = s2 * C[i][j];
  for (int rotidx = 0; rotidx < R; ++rotidx)
    gauss_filter_2d(rot_filter_result.data() + rotidx * HW, log_joint_01.data() + rotidx * HW, H, W,
                    scaleC, is_sparse);
  if (dbg_spatial) memcpy(dbg_spatial, log_joint_01.data(), N * sizeof(float));

  compute_log_grid(log_joint_01.data(), N);  // :432
  add_grid1(log_joint_01.data(), N, in_M);   // :435

  for (int rotidx = 0; rotidx < R; ++rotidx) {  // :438-448
    float alpha = rot_from_index(ep, rotidx) * M_PI / 180.0;
    Mat3 Tgo = mat3_prod(hc_rotation(alpha), hc_scaling(scale));
    double offset_g_01[3];
    mat3_vec(Tgo, offset_p_01, offset_g_01);
    Mat3 Tobject = hc_translation(-offset_g_01[0], -offset_g_01[1]);
    Mat3 dummy;
    transform_grid_helper(log_joint_01.data() + rotidx * HW, H, W, log_prob_parent + rotidx * HW, H,
                          W, Tobject, dummy, (float)LOG_ZERO, TM_NEAREST, false);
  }
}

// ---------------------------------------------------------------------------------------------
// Joints (libPictStruct/objectdetect.h:54-86)
// ---------------------------------------------------------------------------------------------
struct Joint {
  int type;  // POS_GAUSSIAN = 1, ROT_GAUSSIAN = 2
  int child_idx, parent_idx;
  double offset_c[2], offset_p[2];
  double C[4];  // row-major 2x2
  double rot_mean, rot_sigma;
};

// objectdetect_findrot.cpp:59-71
void get_incoming_joints(const std::vector<Joint> &joints, int curidx, std::vector<int> &all_children,
                         std::vector<int> &all_joints) {
  all_children.clear();
  all_joints.clear();
  for (size_t jidx = 0; jidx < joints.size(); ++jidx)
    if (joints[jidx].parent_idx == curidx) {
      all_children.push_back(joints[jidx].child_idx);
      all_joints.push_back((int)jidx);
    }
}

struct Hyp {  // objectdetect.h:88-195 PartHyp::toVect order
  float scaleidx, scale, rotidx, rot, x, y, score;
};

// objectdetect_aux.cpp:193-261.  dim0 is rotation for parts, scale for the root posterior.
// Emits (dim0, x, y, score) in scan order dim0 / x / y; if more than max_n, keeps the max_n best
// by fp32 score (std::sort is unstable: ties at the cut are implementation-defined).
struct LocalMax {
  int d0, x, y;
  float score;
};

void find_local_max(const float *g, int D0, int H, int W, std::vector<LocalMax> &local_max,
                    int max_hypothesis_number) {
  const size_t HW = (size_t)H * W;
  local_max.clear();
  for (int sidx = 0; sidx < D0; ++sidx) {
    const float *s = g + sidx * HW;
    for (int x = 0; x < W; ++x)
      for (int y = 0; y < H; ++y) {
        bool is_max = true;
        float c = s[(size_t)y * W + x];
        for (int dy = -1; dy <= 1 && is_max; ++dy)
          for (int dx = -1; dx <= 1 && is_max; ++dx) {
            int xpos = x + dx, ypos = y + dy;
            if (xpos >= 0 && xpos < W && ypos >= 0 && ypos < H && !(dx == 0 && dy == 0))
              if (s[(size_t)ypos * W + xpos] > c) is_max = false;
          }
        if (is_max && sidx > 0) is_max = g[(sidx - 1) * HW + (size_t)y * W + x] < c;
        if (is_max && sidx < D0 - 1) is_max = g[(sidx + 1) * HW + (size_t)y * W + x] < c;
        if (is_max) local_max.push_back(LocalMax{sidx, x, y, c});
      }
  }
  if ((int)local_max.size() > max_hypothesis_number) {
    std::vector<std::pair<float, int>> validx;
    validx.reserve(local_max.size());
    for (int idx = 0; idx < (int)local_max.size(); ++idx)
      validx.push_back(std::pair<float, int>(local_max[idx].score, idx));
    std::sort(validx.begin(), validx.end(),
              [](const std::pair<float, int> &a, const std::pair<float, int> &b) {
                return a.first > b.first;
              });
    std::vector<LocalMax> keep;
    for (int idx = 0; idx < max_hypothesis_number; ++idx) keep.push_back(local_max[validx[idx].second]);
    local_max = keep;
  }
}

Hyp make_part_hyp(const ExpParam &ep, int scaleidx, int rotidx, int x, int y, float score) {
  // objectdetect.h:91-97 (m_scale, m_rot are floats) and toVect :139-160
  Hyp h;
  h.scaleidx = (float)scaleidx;
  h.scale = (float)scale_from_index(ep, scaleidx);
  h.rotidx = (float)rotidx;
  h.rot = (float)rot_from_index(ep, rotidx);
  h.x = (float)x;
  h.y = (float)y;
  h.score = score;
  return h;
}

// Per-part readout, objectdetect_findrot.cpp:261-285 (and getMaxStates :88-109):
// entry 0 = global argmax (first strictly-greater in flat order), then <=K local maxima
// (findLocalMax wrapper aux.cpp:295-309 always tags them scaleidx 0).
void readout_part(const ExpParam &ep, const float *post, int R, int H, int W, int scaleidx,
                  std::vector<Hyp> &out) {
  float bestval = -DBL_MAX;  // float bestval = -DBL_MAX  -> -inf
  int bestidx = -1;
  int n = R * H * W;
  for (int idx = 0; idx < n; ++idx)
    if (post[idx] > bestval) {
      bestval = post[idx];
      bestidx = idx;
    }
  assert(bestidx >= 0);
  // disc_ps::index_from_flat3, libDiscPS/disc_sample.cpp:123-138
  int best_rotidx = bestidx / (H * W);
  int rem = bestidx % (H * W);
  int best_y = rem / W;
  int best_x = rem % W;
  out.clear();
  out.push_back(make_part_hyp(ep, scaleidx, best_rotidx, best_x, best_y, bestval));
  std::vector<LocalMax> lm;
  find_local_max(post, R, H, W, lm, ep.roi_save_num_samples);
  for (const LocalMax &m : lm) out.push_back(make_part_hyp(ep, 0, m.d0, m.x, m.y, m.score));
}

struct InferOut {
  float *marginals;     // optional [S][P][R][H][W]
  float *root_post;     // [S][H][W]
  std::vector<std::vector<Hyp>> best_part_hyp;  // describes the LAST scale only (findrot.cpp:257-259)
};

// objectdetect_findrot.cpp:124-286  computePartMarginals (downward pass + readout)
void compute_part_marginals(const ExpParam &ep, const std::vector<Joint> &joints, int P, int rootpart_idx,
                            int scaleidx, float *unaries /*[P][S][N]*/, int S, int R, int H, int W,
                            std::vector<std::vector<float>> &log_part_posterior,
                            std::vector<std::vector<float>> &log_from_root,
                            std::vector<std::vector<Hyp>> &best_part_hyp, bool do_readout) {
  const size_t N = (size_t)R * H * W;
  double scale = scale_from_index(ep, scaleidx);
  auto U = [&](int p) { return unaries + ((size_t)p * S + scaleidx) * N; };

  std::vector<bool> computed(P, false);
  std::vector<int> stack;
  std::vector<int> all_children, incoming_joints;
  get_incoming_joints(joints, rootpart_idx, all_children, incoming_joints);

  for (size_t i = 0; i < all_children.size(); ++i) {  // :165-186
    int child_idx = all_children[i];
    const Joint &j = joints[incoming_joints[i]];
    add_grid2(log_from_root[child_idx].data(), U(rootpart_idx), N);
    std::vector<float> tmpgrid2(N);
    double C[2][2] = {{j.C[0], j.C[1]}, {j.C[2], j.C[3]}};
    compute_rot_joint_marginal(ep, log_from_root[child_idx].data(), tmpgrid2.data(), R, H, W, j.offset_p,
                               j.offset_c, C, -j.rot_mean, j.rot_sigma, scale, false, 0, 0, 0);
    log_from_root[child_idx] = tmpgrid2;
    computed[child_idx] = true;
    stack.push_back(child_idx);
  }

  while (!stack.empty()) {  // :192-236
    int curidx = stack.back();
    stack.pop_back();
    assert(computed[curidx]);
    add_grid2(log_part_posterior[curidx].data(), log_from_root[curidx].data(), N);  // :201
    std::vector<int> ch, jn;
    get_incoming_joints(joints, curidx, ch, jn);
    assert(ch.size() <= 1);  // :210
    if (ch.size() == 1) {
      int child_idx = ch[0];
      const Joint &j = joints[jn[0]];
      std::vector<float> tmpgrid(U(curidx), U(curidx) + N);  // copy of the unary, :221
      add_grid2(tmpgrid.data(), log_from_root[curidx].data(), N);
      double C[2][2] = {{j.C[0], j.C[1]}, {j.C[2], j.C[3]}};
      compute_rot_joint_marginal(ep, tmpgrid.data(), log_from_root[child_idx].data(), R, H, W, j.offset_p,
                                 j.offset_c, C, -j.rot_mean, j.rot_sigma, scale, false, 0, 0, 0);
      computed[child_idx] = true;
      stack.push_back(child_idx);
    }
  }

  // :255-285
  best_part_hyp.clear();
  best_part_hyp.resize(P);
  if (do_readout)
    for (int pidx = 0; pidx < P; ++pidx)
      readout_part(ep, log_part_posterior[pidx].data(), R, H, W, scaleidx, best_part_hyp[pidx]);
}

// objectdetect_findrot.cpp:470-727  computeRootPosteriorRot
int compute_root_posterior_rot(const ExpParam &ep, int P, const int *is_detect, const int *is_upright,
                               int rootpart_idx, const std::vector<Joint> &joints, float *unaries, int H,
                               int W, bool is_sparse, bool do_readout, InferOut &out) {
  const int S = ep.num_scale_steps;
  const int R = ep.num_rotation_steps;
  const size_t HW = (size_t)H * W;
  const size_t N = (size_t)R * HW;
  auto U = [&](int p, int s) { return unaries + ((size_t)p * S + s) * N; };

  std::vector<float> root_full((size_t)S * N, (float)LOG_ZERO);  // :499-500

  for (int scaleidx = 0; scaleidx < S; ++scaleidx) {
    // enforce upright orientation, :509-523
    for (int pidx = 0; pidx < P; ++pidx)
      if (is_upright[pidx])
        for (int ridx = 0; ridx < R; ++ridx) {
          double cur_rot = rot_from_index(ep, ridx);
          if (!(std::abs(cur_rot) < 15.0)) {
            float *p = U(pidx, scaleidx) + ridx * HW;
            for (size_t i = 0; i < HW; ++i) p[i] = (float)LOG_ZERO;
          }
        }
    // strip border detections, :528-551 (inner loop shadows scaleidx: all scales, every iteration)
    if (ep.strip_border_detections > 0) {
      assert(ep.strip_border_detections < 0.5);
      int strip_width = (int)(ep.strip_border_detections * W);
      for (int s2 = 0; s2 < S; ++s2)
        for (int ridx = 0; ridx < R; ++ridx)
          for (int iy = 0; iy < H; ++iy) {
            float *row = U(rootpart_idx, s2) + ridx * HW + (size_t)iy * W;
            for (int ix = 0; ix < strip_width; ++ix) row[ix] = (float)LOG_ZERO;
            for (int ix = W - strip_width; ix < W; ++ix) row[ix] = (float)LOG_ZERO;
          }
    }

    double scale = scale_from_index(ep, scaleidx);

    std::vector<std::vector<float>> log_part_posterior(P, std::vector<float>(N, 0.0f));  // :560-561
    std::vector<std::vector<float>> log_from_root(P, std::vector<float>(N, 0.0f));       // :568-570
    std::vector<bool> computed(P, false);
    std::vector<int> compute_stack;
    compute_stack.push_back(rootpart_idx);

    while (!compute_stack.empty()) {  // :582-658
      bool can_compute = true;
      int curidx = compute_stack.back();
      compute_stack.pop_back();
      for (size_t jidx = 0; jidx < joints.size(); ++jidx)
        if (joints[jidx].parent_idx == curidx)
          if (!computed[joints[jidx].child_idx]) {
            can_compute = false;
            compute_stack.push_back(curidx);
            compute_stack.push_back(joints[jidx].child_idx);
            break;
          }
      if (can_compute) {
        std::vector<int> all_children, incoming_joints;
        get_incoming_joints(joints, curidx, all_children, incoming_joints);
        for (size_t i = 0; i < all_children.size(); ++i) {
          int child_idx = all_children[i];
          const Joint &j = joints[incoming_joints[i]];
          std::vector<float> from_child(N);
          double C[2][2] = {{j.C[0], j.C[1]}, {j.C[2], j.C[3]}};
          compute_rot_joint_marginal(ep, log_part_posterior[child_idx].data(), from_child.data(), R, H, W,
                                     j.offset_c, j.offset_p, C, j.rot_mean, j.rot_sigma, scale, is_sparse,
                                     0, 0, 0);
          add_grid2(log_part_posterior[curidx].data(), from_child.data(), N);  // :637
          if (curidx == rootpart_idx)                                           // :641-649
            for (size_t i2 = 0; i2 < all_children.size(); ++i2)
              if (i2 != i) add_grid2(log_from_root[all_children[i2]].data(), from_child.data(), N);
        }
        if (is_detect[curidx])  // :652-654
          add_grid2(log_part_posterior[curidx].data(), U(curidx, scaleidx), N);
        computed[curidx] = true;
      }
    }

    compute_part_marginals(ep, joints, P, rootpart_idx, scaleidx, unaries, S, R, H, W, log_part_posterior,
                           log_from_root, out.best_part_hyp, do_readout);  // :667

    memcpy(root_full.data() + (size_t)scaleidx * N, log_part_posterior[rootpart_idx].data(),
           N * sizeof(float));  // :676
    if (out.marginals)
      for (int p = 0; p < P; ++p)
        memcpy(out.marginals + ((size_t)scaleidx * P + p) * N, log_part_posterior[p].data(),
               N * sizeof(float));
  }

  // marginalise the root over valid rotations, :694-726
  std::vector<int> valid;
  if (is_upright[rootpart_idx]) {
    int k1 = index_from_rot(ep, -1e-6);
    int k2 = index_from_rot(ep, 1e-6);
    if (k1 < 0 || k2 < 0) return -1;
    valid.push_back(k1);
    if (k2 != k1) valid.push_back(k2);
  } else {
    for (int r = 0; r < R; ++r) valid.push_back(r);
  }
  for (int s = 0; s < S; ++s) {
    float *rp = out.root_post + (size_t)s * HW;
    memcpy(rp, root_full.data() + (size_t)s * N + valid[0] * HW, HW * sizeof(float));
  }
  compute_exp_grid(out.root_post, (size_t)S * HW);
  std::vector<float> tmp((size_t)S * HW);
  for (size_t idx = 1; idx < valid.size(); ++idx) {
    for (int s = 0; s < S; ++s)
      memcpy(tmp.data() + (size_t)s * HW, root_full.data() + (size_t)s * N + valid[idx] * HW,
             HW * sizeof(float));
    compute_exp_grid(tmp.data(), (size_t)S * HW);
    add_grid2(out.root_post, tmp.data(), (size_t)S * HW);
  }
  compute_log_grid(out.root_post, (size_t)S * HW);
  return 0;
}

}  // namespace

// =============================================================================================
// C entry points (loaded with ctypes by tests/ and bench.py's CPU-baseline legs)
// =============================================================================================
extern "C" {

struct orc_exp_param {
  int num_rotation_steps;
  float min_part_rotation, max_part_rotation;
  int num_scale_steps;
  float min_object_scale, max_object_scale;
  float strip_border_detections;
  int roi_save_num_samples;
};

struct orc_joint {
  int type, child_idx, parent_idx;
  double offset_c[2], offset_p[2];
  double C[4];
  double rot_mean, rot_sigma;
};

static ExpParam to_ep(const orc_exp_param *e) {
  ExpParam ep;
  ep.num_rotation_steps = e->num_rotation_steps;
  ep.min_part_rotation = e->min_part_rotation;
  ep.max_part_rotation = e->max_part_rotation;
  ep.num_scale_steps = e->num_scale_steps;
  ep.min_object_scale = e->min_object_scale;
  ep.max_object_scale = e->max_object_scale;
  ep.strip_border_detections = e->strip_border_detections;
  ep.roi_save_num_samples = e->roi_save_num_samples;
  return ep;
}

double orc_rot_from_index(const orc_exp_param *e, int idx) { return rot_from_index(to_ep(e), idx); }
double orc_scale_from_index(const orc_exp_param *e, int idx) { return scale_from_index(to_ep(e), idx); }
int orc_index_from_rot(const orc_exp_param *e, double rot) { return index_from_rot(to_ep(e), rot); }

// boost_math.cpp:104-117; returns the length, writes up to cap taps
int orc_gaussian_filter(double sigma, double *out, int cap) {
  std::vector<double> f = get_gaussian_filter(sigma);
  for (int i = 0; i < (int)f.size() && i < cap; ++i) out[i] = f[i];
  return (int)f.size();
}

void orc_eig2d(const double *M, double *V, double *E) {
  double m[2][2] = {{M[0], M[1]}, {M[2], M[3]}}, v[2][2], e[2][2];
  eig2d(m, v, e);
  V[0] = v[0][0]; V[1] = v[0][1]; V[2] = v[1][0]; V[3] = v[1][1];
  E[0] = e[0][0]; E[1] = e[0][1]; E[2] = e[1][0]; E[3] = e[1][1];
}

// gaussFilter2d on one H x W slice (filter.hpp:375-388)
void orc_gauss_filter_2d(const float *in, float *out, int H, int W, const double *C, int sparse) {
  double c[2][2] = {{C[0], C[1]}, {C[2], C[3]}};
  gauss_filter_2d(in, out, H, W, c, sparse != 0);
}

// gaussFilter2dOffset on one H x W slice (filter.hpp:335-369)
void orc_gauss_filter_2d_offset(const float *in, float *out, int H, int W, const double *C, const double *offset, int sparse) {
  double c[2][2] = {{C[0], C[1]}, {C[2], C[3]}};
  gauss_filter_2d_offset(in, out, H, W, c, offset[0], offset[1], sparse != 0);
}

// grid_filter_1d_blas_wraparound (filter.hpp:116-155) on one length-n column, the form compute_rot_joint_marginal
// applies to every pixel: out[i] = sdot over k of in[(i + k - nx) mod n] * f[k]
void orc_filter_1d_wraparound(const float *in, float *out, int n, const float *f, int f_len) {
  int nx = (f_len - 1) / 2;
  for (int i = 0; i < n; ++i) {
    float acc = 0.0f;
    for (int k = 0; k < f_len; ++k) acc += in[((i + k - nx) % n + n) % n] * f[k];
    out[i] = acc;
  }
}

// computeLogGrid (op 0) / computeExpGrid (op 1), multi_array_op.hpp:154-180
void orc_pointwise(int op, float *a, size_t n) {
  if (op == 0) compute_log_grid(a, n);
  else compute_exp_grid(a, n);
}

void orc_hc_inverse(const double *T, double *out) {
  Mat3 M;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M.m[i][j] = T[i * 3 + j];
  Mat3 I = hc_inverse(M);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i * 3 + j] = I.m[i][j];
}

void orc_transformed_bbox(const double *T, int w, int h, double *out4) {
  Mat3 M;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M.m[i][j] = T[i * 3 + j];
  hc_transformed_bbox(M, w, h, out4[0], out4[1], out4[2], out4[3]);
}

// size of the enlarged grid used by gaussFilter2dOffset for covariance C (transform.hpp:290-299)
void orc_enlarged_size(int H, int W, const double *C, int *oh, int *ow) {
  double c[2][2] = {{C[0], C[1]}, {C[2], C[3]}}, V[2][2], E[2][2];
  eig2d(c, V, E);
  double Vt[2][2] = {{V[0][0], V[1][0]}, {V[0][1], V[1][1]}};
  Mat3 T21 = hc_homogeneous(Vt, 0, 0);
  double minx, miny, maxx, maxy;
  hc_transformed_bbox(T21, W, H, minx, miny, maxx, maxy);
  *ow = (int)ceil(maxx - minx);
  *oh = (int)ceil(maxy - miny);
}

// transform_grid_fixed_size (transform.hpp:245-257); T is row-major 3x3
void orc_transform_fixed(const float *in, int in_h, int in_w, float *out, int out_h, int out_w,
                         const double *T, float default_value, int method) {
  Mat3 T21, T23;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T21.m[i][j] = T[i * 3 + j];
  transform_grid_helper(in, in_h, in_w, out, out_h, out_w, T21, T23, default_value, method, false);
}

// PartApp::loadScoreGrid's mapping of one (scale, rotation) compact grid to the image grid
// (libPartApp/partapp.cpp:874-896): Tig = prod(Ti2, T2g) is formed by the caller; transform_grid_fixed_size(...,
// Tig, NO_CLASS_VALUE = 0, TM_DIRECT or TM_BILINEAR).
void orc_load_score_grid(const float *cells, int R, int gh, int gw, const double *Tig /*[R][9]*/, int H, int W,
                         int interpolate, float *out /*[R][H][W]*/) {
  for (int r = 0; r < R; ++r) {
    Mat3 T21, T23;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) T21.m[i][j] = Tig[r * 9 + i * 3 + j];
    transform_grid_helper(cells + (size_t)r * gh * gw, gh, gw, out + (size_t)r * H * W, H, W, T21, T23, 0.0f,
                          interpolate ? TM_BILINEAR : TM_DIRECT, false);
  }
}

// computePosJointMarginal (findpos.cpp:64-89) on D independent [H][W] slices; child is rewritten like the reference.
void orc_pos_message(float *child, float *parent, int D, int H, int W, const double *offset, const double *C,
                     double scale, int sparse) {
  double c[2][2] = {{C[0], C[1]}, {C[2], C[3]}};
  for (int d = 0; d < D; ++d)
    compute_pos_joint_marginal(child + (size_t)d * H * W, parent + (size_t)d * H * W, H, W, offset, c, scale, sparse != 0);
}

// computeRotJointMarginal (findrot.cpp:292-456). dbg_* may be NULL.
void orc_message(const orc_exp_param *e, const float *child, float *parent, int R, int H, int W,
                 const double *off_in, const double *off_out, const double *C, double rot_mean,
                 double rot_sigma, double scale, int sparse, float *dbg_exp, float *dbg_rot,
                 float *dbg_spatial) {
  double c[2][2] = {{C[0], C[1]}, {C[2], C[3]}};
  compute_rot_joint_marginal(to_ep(e), child, parent, R, H, W, off_in, off_out, c, rot_mean, rot_sigma,
                             scale, sparse != 0, dbg_exp, dbg_rot, dbg_spatial);
}

// (float)exp((double)x) / (float)log((double)x) over fp32 bit patterns [first, first+count): the oracle's libm
// convention (multi_array_op.hpp:165,177), for comparison with the device functions.  Returns the number of
// positions where `dev` differs bitwise (NaNs compare equal to NaNs).
unsigned long long orc_compare_math(int op, unsigned first, unsigned count, const float *dev) {
  unsigned long long bad = 0;
  for (unsigned i = 0; i < count; ++i) {
    unsigned bits = first + i;
    float x;
    memcpy(&x, &bits, 4);
    float r = op == 0 ? (x < -104.0f ? 0.0f : (float)exp((double)x)) : (x == 0.0f ? (float)LOG_ZERO : (float)log((double)x));
    if (r != r && dev[i] != dev[i]) continue;
    if (memcmp(&r, &dev[i], 4) != 0) ++bad;
  }
  return bad;
}

// Unary prep, findrot.cpp:834-845: clip_scores_fill (aux.hpp:42-59) then computeLogGrid.
void orc_prepare_unary(float *g, size_t n) {
  for (size_t i = 0; i < n; ++i)
    if (g[i] < 0) g[i] = (float)0.0001;  // pData[i3] = min_val (double 1e-4 narrowed)
  compute_log_grid(g, n);
}

// loadJoints flip branch, objectdetect_aux.cpp:102-119: C <- T*(C*T), T = diag(-1,1)
void orc_flip_joint(orc_joint *j) {
  double T[2][2] = {{-1, 0}, {0, 1}};
  double C[2][2] = {{j->C[0], j->C[1]}, {j->C[2], j->C[3]}}, CT[2][2], TCT[2][2];
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 2; ++k) {
      double t = 0;
      for (int l = 0; l < 2; ++l) t += C[i][l] * T[l][k];
      CT[i][k] = t;
    }
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 2; ++k) {
      double t = 0;
      for (int l = 0; l < 2; ++l) t += T[i][l] * CT[l][k];
      TCT[i][k] = t;
    }
  double op[2], oc[2];
  for (int i = 0; i < 2; ++i) {
    double t = 0, u = 0;
    for (int l = 0; l < 2; ++l) {
      t += T[i][l] * j->offset_p[l];
      u += T[i][l] * j->offset_c[l];
    }
    op[i] = t;
    oc[i] = u;
  }
  j->C[0] = TCT[0][0]; j->C[1] = TCT[0][1]; j->C[2] = TCT[1][0]; j->C[3] = TCT[1][1];
  j->offset_p[0] = op[0]; j->offset_p[1] = op[1];
  j->offset_c[0] = oc[0]; j->offset_c[1] = oc[1];
  if (j->type == 2) j->rot_mean = -j->rot_mean;
}

// ---- conditioning tables (objectdetect_icps.cpp) ---------------------------------------------
// icps.cpp:29 has "using namespace std", so on floats log() is std::log(float) (= logf); exp() of the double
// expression is the double routine, narrowed on assignment to float.  pow(float, int) is DIALECT-DEPENDENT: the
// reference's toolchain (gcc 4.7, -std=gnu++98 by default) has the overload float pow(float, int) = __builtin_powif,
// i.e. x*x rounded to fp32; since C++11 (LWG 550) the call promotes both arguments and squares in double.  The default
// here is the authors' dialect.  oracle/_ref compiles the reference's sources as C++17 (the stand-in headers need it),
// so tests/test_oracle_vs_ref.py switches this restatement to the C++11 rule for the one comparison that involves
// pow() on non-integral values (getPosScoreGrid) -- which pins everything else about these routines.
static int g_pow_cxx11 = 0;
static inline double powf2(float t) { return g_pow_cxx11 ? (double)t * (double)t : (double)(t * t); }
void orc_set_pow_dialect(int cxx11) { g_pow_cxx11 = cxx11; }

// getRotScoreGrid :228-281: table[r] for one part; mu/var arrive as doubles and are narrowed.
void orc_rot_score_table(const orc_exp_param *e, double mu_d, double var_d, float *table) {
  ExpParam ep = to_ep(e);
  float mu = mu_d, var = var_d;
  for (int rotidx = 0; rotidx < ep.num_rotation_steps; ++rotidx) {
    float rot = rot_from_index(ep, rotidx) / 180 * M_PI;
    float score = 0.0;
    float d = rot - mu;           // square(rot-mu): template square<float>
    score = exp(-0.5 * (d * d) / var);
    if (score < 1e-4) score = 1e-4;
    score = logf(score);          // icps.cpp:29 "using namespace std": log(float) is std::log(float)
    double stored = score;        // pred_rot is a double_vector
    table[rotidx] = (float)stored;
  }
}

// getPosScoreGrid :366-423: table[y][x] for one non-root part relative to the detected root.
void orc_pos_score_table(int H, int W, double mu_x_d, double mu_y_d, double var_x_d, double var_y_d,
                         double root_x, double root_y, float *table) {
  float var_weight = 1.0;
  float mu_x = mu_x_d, mu_y = mu_y_d;
  float var_x = var_x_d * var_weight, var_y = var_y_d * var_weight;
  for (int iy = 0; iy < H; ++iy)
    for (int ix = 0; ix < W; ++ix) {
      float ix_rel = ix - root_x;
      float iy_rel = iy - root_y;
      float score_x = exp(-0.5 * powf2(ix_rel - mu_x) / var_x);
      float score_y = exp(-0.5 * powf2(iy_rel - mu_y) / var_y);
      float score = score_x * score_y;
      if (score < 1e-4) score = 1e-4;
      score = logf(score);
      table[(size_t)iy * W + ix] = (float)(double)score;
    }
}

// setTorsoPosPrior :137-191: table[y][x] (already weighted), added to every rotation/scale of root.
void orc_torso_prior_table(int H, int W, double mu_x_d, double mu_y_d, double var_x_d, double var_y_d,
                           float weight, float *table) {
  float mu_x = mu_x_d, mu_y = mu_y_d, var_x = var_x_d, var_y = var_y_d;
  float img_c_x = 0.5 * W;
  float img_c_y = 0.5 * H;
  float sigma_weight = 1.0;
  float var_weight = sigma_weight * sigma_weight;
  var_x *= var_weight;
  var_y *= var_weight;
  for (int iy = 0; iy < H; ++iy)
    for (int ix = 0; ix < W; ++ix) {
      float ix_rel = img_c_x - ix;
      float iy_rel = img_c_y - iy;
      float score_x = exp(-0.5 * powf2(ix_rel - mu_x) / var_x);
      float score_y = exp(-0.5 * powf2(iy_rel - mu_y) / var_y);
      float score = weight * score_x * score_y;
      if (score < 1e-4) score = 1e-4;
      score = logf(score);
      table[(size_t)iy * W + ix] = (float)(double)score;
    }
}

// addExtraUnary :526-548 with a per-rotation table broadcast over (y,x): g += weight*table[r]
void orc_add_rot_table(float *g, int R, int H, int W, const float *table, float weight) {
  size_t HW = (size_t)H * W;
  for (int r = 0; r < R; ++r)
    for (size_t i = 0; i < HW; ++i) g[r * HW + i] += weight * table[r];
}

// addExtraUnary with a per-position table broadcast over rotations: g += weight*table[y][x]
void orc_add_pos_table(float *g, int R, int H, int W, const float *table, float weight) {
  size_t HW = (size_t)H * W;
  for (int r = 0; r < R; ++r)
    for (size_t i = 0; i < HW; ++i) g[r * HW + i] += weight * table[i];
}

// setTorsoPosPrior's final add (:183-190): g += table[y][x] (float += double holding a float)
void orc_add_pos_table_unweighted(float *g, int R, int H, int W, const float *table) {
  size_t HW = (size_t)H * W;
  for (int r = 0; r < R; ++r)
    for (size_t i = 0; i < HW; ++i) g[r * HW + i] += table[i];
}

// addDPMScore (icps.cpp:488-524): g += dpm_weight * log_dpmPriorGrid[bRot ? r : 0]
void orc_add_dpm_score(float *g, int R, int H, int W, const float *grid, int nrot, float dpm_weight) {
  size_t HW = (size_t)H * W;
  for (int r = 0; r < R; ++r)
    for (size_t i = 0; i < HW; ++i) g[r * HW + i] += dpm_weight * grid[(nrot > 1 ? r : 0) * HW + i];
}

// addLoadDPMScore (icps.cpp:445-486): g += (val > 1e-4 ? dpm_weight*log(val) : log(1e-4)); log(float) is logf there
void orc_add_load_dpm_score(float *g, int R, int H, int W, const float *grid, int nrot, float dpm_weight) {
  size_t HW = (size_t)H * W;
  for (int r = 0; r < R; ++r) {
    int rd = (nrot == R ? r : 0);
    for (size_t i = 0; i < HW; ++i) {
      float val = grid[rd * HW + i];
      g[r * HW + i] += (val > 1e-4 ? dpm_weight * logf(val) : log(1e-4));
    }
  }
}

// ---- readout ---------------------------------------------------------------------------------

// findLocalMax core (aux.cpp:193-261). out: rows of (d0, x, y, score); returns the count (<= max_n).
int orc_find_local_max(const float *g, int D0, int H, int W, int max_n, float *out) {
  std::vector<LocalMax> lm;
  find_local_max(g, D0, H, W, lm, max_n);
  for (size_t i = 0; i < lm.size(); ++i) {
    out[4 * i + 0] = (float)lm[i].d0;
    out[4 * i + 1] = (float)lm[i].x;
    out[4 * i + 2] = (float)lm[i].y;
    out[4 * i + 3] = lm[i].score;
  }
  return (int)lm.size();
}

// argmax as in findrot.cpp:261-277: returns flat index of the first maximum
int orc_argmax(const float *g, int n, float *val) {
  float bestval = -DBL_MAX;
  int bestidx = -1;
  for (int idx = 0; idx < n; ++idx)
    if (g[idx] > bestval) {
      bestval = g[idx];
      bestidx = idx;
    }
  if (val) *val = bestval;
  return bestidx;
}

// computeRootPosteriorRot (findrot.cpp:470-727).
//   unaries   [P][S][R][H][W] log-domain, mutated in place (upright mask, border strip)
//   marginals optional [S][P][R][H][W]
//   root_post [S][H][W]
//   best_conf [P][7]   argmax record of every part for the LAST scale (toVect order)
//   part_hyps optional [P][cap][7], n_part_hyps [P]: argmax followed by <=K local maxima
// Returns 0 on success.
int orc_infer(const orc_exp_param *e, int P, const int *is_detect, const int *is_upright, int root_idx,
              const orc_joint *joints, int J, int H, int W, float *unaries, int sparse,
              float *marginals, float *root_post, float *best_conf, float *part_hyps, int cap,
              int *n_part_hyps) {
  ExpParam ep = to_ep(e);
  std::vector<Joint> js(J);
  for (int i = 0; i < J; ++i) {
    js[i].type = joints[i].type;
    js[i].child_idx = joints[i].child_idx;
    js[i].parent_idx = joints[i].parent_idx;
    for (int k = 0; k < 2; ++k) {
      js[i].offset_c[k] = joints[i].offset_c[k];
      js[i].offset_p[k] = joints[i].offset_p[k];
    }
    for (int k = 0; k < 4; ++k) js[i].C[k] = joints[i].C[k];
    js[i].rot_mean = joints[i].rot_mean;
    js[i].rot_sigma = joints[i].rot_sigma;
  }
  InferOut out;
  out.marginals = marginals;
  out.root_post = root_post;
  int rc = compute_root_posterior_rot(ep, P, is_detect, is_upright, root_idx, js, unaries, H, W,
                                      sparse != 0, best_conf != 0, out);
  if (rc) return rc;
  if (best_conf)
    for (int p = 0; p < P; ++p) {
      const Hyp &h = out.best_part_hyp[p][0];
      float *r = best_conf + 7 * p;
      r[0] = h.scaleidx; r[1] = h.scale; r[2] = h.rotidx; r[3] = h.rot; r[4] = h.x; r[5] = h.y;
      r[6] = h.score;
      if (part_hyps) {
        int n = std::min((int)out.best_part_hyp[p].size(), cap);
        n_part_hyps[p] = n;
        for (int i = 0; i < n; ++i) {
          const Hyp &q = out.best_part_hyp[p][i];
          float *d = part_hyps + ((size_t)p * cap + i) * 7;
          d[0] = q.scaleidx; d[1] = q.scale; d[2] = q.rotidx; d[3] = q.rot; d[4] = q.x; d[5] = q.y;
          d[6] = q.score;
        }
      }
    }
  return 0;
}

// getMaxStates (findrot.cpp:73-110): use_pairwise:false shortcut, scale 0 only.
void orc_get_max_states(const orc_exp_param *e, int P, int H, int W, const float *unaries,
                        float *best_conf) {
  ExpParam ep = to_ep(e);
  int S = ep.num_scale_steps, R = ep.num_rotation_steps;
  size_t N = (size_t)R * H * W;
  for (int p = 0; p < P; ++p) {
    std::vector<Hyp> hyps;
    ExpParam ep1 = ep;
    ep1.roi_save_num_samples = 0;
    readout_part(ep1, unaries + ((size_t)p * S + 0) * N, R, H, W, 0, hyps);
    const Hyp &h = hyps[0];
    float *r = best_conf + 7 * p;
    r[0] = h.scaleidx; r[1] = h.scale; r[2] = h.rotidx; r[3] = h.rot; r[4] = h.x; r[5] = h.y;
    r[6] = h.score;
  }
}

}  // extern "C"
