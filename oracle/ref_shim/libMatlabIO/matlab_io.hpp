// Stand-in for libMatlabIO (MATLAB's libmat is not installed) -- TEST INFRASTRUCTURE.  File output is not part of what
// oracle/_ref pins: the savers do nothing, the loaders abort.
#pragma once
#include <QString>
#include <cstddef>
#include <cstdlib>
#include <libMultiArray/multi_array_def.h>
#include <libBoostMath/boost_math.h>
#include <libBoostMath/boost_math.hpp>  // the real header reaches it too; objectdetect_findrot.cpp relies on that
struct MATFile {};
namespace matlab_io {
inline MATFile *mat_open(QString, const char *) { static MATFile f; return &f; }
inline void mat_close(MATFile *) {}
// The reference hands its per-part marginals to mat_save_multi_array (objectdetect_findrot.cpp:239-253); the hook lets
// oracle/ref_drivers.cpp collect them instead of writing files.
typedef void (*capture_fn)(const char *file, const char *var, const void *data, size_t bytes);
extern capture_fn g_capture;
template <class A> bool mat_save_multi_array(QString file, QString var, const A &a) {
  if (g_capture) g_capture(file.toStdString().c_str(), var.toStdString().c_str(), a.data(), a.num_elements() * sizeof(typename A::element));
  return true;
}
template <class A> bool mat_save_multi_array(MATFile *, QString, const A &) { return true; }
// Loader hook (oracle/ref_eval.cpp): the evaluator reads pose_est_imgidx*.mat through mat_load_multi_array<FloatGrid2>;
// a registered provider hands it the matrix a test supplies instead of a file.  Without a provider the loaders abort.
typedef bool (*load2d_fn)(const char *file, const char *var, int *rows, int *cols, const float **data);
extern load2d_fn g_load2d;
template <class A> A mat_load_multi_array(QString file, QString var) {
  if constexpr (A::dimensionality == 2) {
    int rows = 0, cols = 0;
    const float *data = 0;
    if (!g_load2d || !g_load2d(file.toStdString().c_str(), var.toStdString().c_str(), &rows, &cols, &data)) abort();
    A a(boost::extents[rows][cols]);
    for (int i = 0; i < rows * cols; ++i) a.data()[i] = data[i];
    return a;
  } else {
    abort();
  }
}
template <class A> A mat_load_multi_array(MATFile *, QString) { abort(); }
inline bool mat_load_double_matrix(QString, QString, boost_math::double_matrix &) { abort(); }
inline bool mat_load_double_vector(QString, QString, boost_math::double_vector &) { abort(); }
inline bool mat_save_double_matrix(QString, QString, const boost_math::double_matrix &) { return true; }
inline bool mat_save_double_matrix(MATFile *, QString, const boost_math::double_matrix &) { return true; }
inline bool mat_save_double_vector(QString, QString, const boost_math::double_vector &) { return true; }
inline bool mat_save_double_vector(MATFile *, QString, const boost_math::double_vector &) { return true; }
inline bool mat_load_double_matrix(MATFile *, QString, boost_math::double_matrix &) { abort(); }
inline bool mat_load_double_vector(MATFile *, QString, boost_math::double_vector &) { abort(); }
template <class V> bool mat_load_multi_array_vec2(MATFile *, QString, V &) { abort(); }
template <class V> bool mat_save_multi_array_vec2(MATFile *, QString, const V &) { return true; }
template <class V> bool mat_save_multi_array_vec2(QString, QString, const V &) { return true; }
template <class V> bool mat_load_multi_array_vec2(QString, QString, V &) { abort(); }
template <class V> bool mat_load_stdcpp_vector(QString, QString, V &) { abort(); }
}  // namespace matlab_io
