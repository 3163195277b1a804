// ref_core.cpp -- TEST INFRASTRUCTURE: C entry points over the reference's OWN numeric core, compiled unmodified from
// /root/reference/src/libs (libMultiArray/multi_array_{op,transform,filter}.hpp, libBoostMath/boost_math.cpp,
// libBoostMath/homogeneous_coord.cpp, libPartApp/partapp_aux.hpp, libPictStruct/objectdetect_aux.hpp) against the container stand-ins of oracle/ref_shim/ (Boost, Qt and a BLAS are
// not installed in this image).  Built by `make -C oracle ref` into oracle/_ref/libps_ref_core.so, only where
// /root/reference exists.  tests/test_oracle_vs_ref.py and tests/golden/make_ref_golden.py use it to PIN the oracle's
// restatement of these routines (except the BLAS order inside them) against the reference's code; nothing in partapp_b200/ may
// load it.  Not pinned by this: the BLAS summation order (cblas_sdot below is the Netlib order, a convention) and the
// libm of the authors' machine.
#include <cmath>
#include <cstring>

#include <libMultiArray/multi_array_op.hpp>
#include <libMultiArray/multi_array_transform.hpp>
#include <libMultiArray/multi_array_filter.hpp>
#include <libBoostMath/boost_math.hpp>
#include <libBoostMath/homogeneous_coord.h>
#include <libPartApp/partapp_aux.hpp>  // over ref_shim/libPartApp/ExpParam.pb.h
#include <libPictStruct/objectdetect_aux.hpp>

extern "C" float cblas_sdot(const int n, const float *x, const int incx, const float *y, const int incy) {
  float acc = 0.0f;  // Netlib sdot: sequential, ascending, multiply then add (its 5-way unroll keeps this order)
  for (int i = 0; i < n; ++i) acc = acc + x[(long)i * incx] * y[(long)i * incy];
  return acc;
}

namespace {
using boost_math::double_matrix;
using boost_math::double_vector;

double_matrix mat3(const double *m) {
  double_matrix M(3, 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M(i, j) = m[i * 3 + j];
  return M;
}
double_matrix mat2(const double *m) {
  double_matrix M(2, 2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) M(i, j) = m[i * 2 + j];
  return M;
}
void put(const double_matrix &M, double *out) {
  for (size_t i = 0; i < M.size1(); ++i)
    for (size_t j = 0; j < M.size2(); ++j) out[i * M.size2() + j] = M(i, j);
}
FloatGrid2 grid2(const float *p, int h, int w) {
  FloatGrid2 g(boost::extents[h][w]);
  memcpy(g.data(), p, sizeof(float) * (size_t)h * w);
  return g;
}
}  // namespace

extern "C" {

int ref_gaussian_filter(double sigma, double *out, int cap) {
  double_vector f;
  boost_math::get_gaussian_filter(f, sigma, false);
  if ((int)f.size() > cap) return -(int)f.size();
  for (size_t i = 0; i < f.size(); ++i) out[i] = f(i);
  return (int)f.size();
}

void ref_eig2d(const double *C, double *V, double *E) {
  double_matrix v(2, 2), e(2, 2);
  boost_math::eig2d(mat2(C), v, e);
  put(v, V);
  put(e, E);
}

void ref_hc_inverse(const double *T, double *out) { put(hc::inverse(mat3(T)), out); }

void ref_hc_compose(int kind, double a, double b, double *out) {  // 0 rotation(a), 1 scaling(a), 2 translation(a, b)
  put(kind == 0 ? hc::get_rotation_matrix(a) : kind == 1 ? hc::get_scaling_matrix(a) : hc::get_translation_matrix(a, b), out);
}

void ref_prod3(const double *A, const double *B, double *out) { put(prod(mat3(A), mat3(B)), out); }

void ref_map_point(const double *T, double x, double y, double *ox, double *oy) { hc::map_point(mat3(T), x, y, *ox, *oy); }

void ref_transformed_bbox(const double *T21, int w, int h, double *out4) {
  hc::get_transformed_bbox(mat3(T21), w, h, out4[0], out4[1], out4[2], out4[3]);
}

// transform_grid_fixed_size (multi_array_transform.hpp:245-257)
void ref_transform_fixed(const float *in, int ih, int iw, float *out, int oh, int ow, const double *T21, float default_value,
                         int method) {
  FloatGrid2 gi = grid2(in, ih, iw), go(boost::extents[oh][ow]);
  multi_array_op::transform_grid_fixed_size(gi, go, mat3(T21), default_value, (TransformationMethod)method);
  memcpy(out, go.data(), sizeof(float) * (size_t)oh * ow);
}

// transform_grid_resize (:285-305); returns 0 and fills out / T23 when cap is large enough, else the needed size
long ref_transform_resize(const float *in, int ih, int iw, const double *T21, float default_value, int method, float *out,
                          long cap, int *oh, int *ow, double *T23) {
  FloatGrid2 gi = grid2(in, ih, iw), go;
  double_matrix t23;
  multi_array_op::transform_grid_resize(gi, go, mat3(T21), t23, default_value, (TransformationMethod)method);
  *oh = (int)go.shape()[0];
  *ow = (int)go.shape()[1];
  const long n = (long)go.num_elements();
  if (n > cap) return n;
  memcpy(out, go.data(), sizeof(float) * n);
  put(t23, T23);
  return 0;
}

void ref_gauss_filter_diag2d(const float *in, float *out, int h, int w, double var_x, double var_y) {
  FloatGrid2 gi = grid2(in, h, w), go(boost::extents[h][w]);
  double_matrix C(2, 2);
  C(0, 0) = var_x; C(0, 1) = 0; C(1, 0) = 0; C(1, 1) = var_y;
  multi_array_op::gaussFilterDiag2d(gi, go, C, false);
  memcpy(out, go.data(), sizeof(float) * (size_t)h * w);
}

void ref_gauss_filter_2d(const float *in, float *out, int h, int w, const double *C, int sparse) {
  FloatGrid2 gi = grid2(in, h, w), go(boost::extents[h][w]);
  multi_array_op::gaussFilter2d(gi, go, mat2(C), false, sparse != 0);
  memcpy(out, go.data(), sizeof(float) * (size_t)h * w);
}

void ref_gauss_filter_2d_offset(const float *in, float *out, int h, int w, const double *C, const double *offset, int sparse) {
  FloatGrid2 gi = grid2(in, h, w), go(boost::extents[h][w]);
  double_vector off(2);
  off(0) = offset[0];
  off(1) = offset[1];
  multi_array_op::gaussFilter2dOffset(gi, go, mat2(C), off, false, sparse != 0);
  memcpy(out, go.data(), sizeof(float) * (size_t)h * w);
}

// grid_filter_1d_blas_wraparound (multi_array_filter.hpp:116-155)
void ref_filter_1d_wraparound(const float *in, float *out, int n, const float *f, int f_len) {
  FloatGrid1 gi(boost::extents[n]), go(boost::extents[n]);
  memcpy(gi.data(), in, sizeof(float) * n);
  multi_array_op::grid_filter_1d_blas_wraparound(gi, go, f, f_len);
  memcpy(out, go.data(), sizeof(float) * n);
}

// multi_array_op.hpp pointwise sweeps on a flat grid: 0 computeLogGrid, 1 computeExpGrid, 2 addGrid2 (a += b),
// 3 addGrid1 (a += scalar), 4 setGrid (a = scalar), 5 the unary prep of findrot.cpp:834-845 (clip_scores_fill from
// objectdetect_aux.hpp, then computeLogGrid)
void ref_pointwise(int op, float *a, const float *b, float scalar, long n) {
  FloatGrid1 ga(boost::extents[n]);
  memcpy(ga.data(), a, sizeof(float) * n);
  if (op == 0) multi_array_op::computeLogGrid(ga);
  else if (op == 1) multi_array_op::computeExpGrid(ga);
  else if (op == 2) {
    FloatGrid1 gb(boost::extents[n]);
    memcpy(gb.data(), b, sizeof(float) * n);
    multi_array_op::addGrid2(ga, gb);
  } else if (op == 3) multi_array_op::addGrid1(ga, scalar);
  else if (op == 4) multi_array_op::setGrid(ga, scalar);
  else if (op == 5) {
    object_detect::clip_scores_fill(ga);
    multi_array_op::computeLogGrid(ga);
  }
  memcpy(a, ga.data(), sizeof(float) * n);
}

// partapp_aux.hpp bin centres: what 0 rot_from_index, 1 scale_from_index (double result), 2 index_from_rot(value)
double ref_bins(int what, float mn, float mx, unsigned n, int idx, double value) {
  ExpParam e;
  e.min_part_rotation_ = e.min_object_scale_ = mn;
  e.max_part_rotation_ = e.max_object_scale_ = mx;
  e.num_rotation_steps_ = e.num_scale_steps_ = n;
  if (what == 0) return rot_from_index(e, idx);
  if (what == 1) return scale_from_index(e, idx);
  return (double)index_from_rot(e, value);
}

void ref_min_max(const float *a, long n, float *mn, float *mx) {
  FloatGrid1 ga(boost::extents[n]);
  memcpy(ga.data(), a, sizeof(float) * n);
  multi_array_op::getMinMax(ga, *mn, *mx);
}

}  // extern "C"
