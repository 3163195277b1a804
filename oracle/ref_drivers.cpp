// ref_drivers.cpp -- TEST INFRASTRUCTURE: C entry points over the reference's OWN inference drivers.
// libPictStruct/objectdetect_findrot.cpp (computeRotJointMarginal, computePartMarginals, computeRootPosteriorRot) is
// compiled UNMODIFIED from /root/reference next to this file (`make -C oracle ref`), against the stand-ins of
// oracle/ref_shim/ for everything the image lacks (Boost, Qt, protoc output, libmat, the detector libraries).  This file
// defines what that translation unit references but other, uncompilable translation units define.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <libPartApp/partapp_aux.hpp>
#include <libPictStruct/objectdetect.h>

extern "C" float cblas_sdot(const int n, const float *x, const int incx, const float *y, const int incy);

// ---- definitions for symbols of translation units that are not compiled here -------------------------------------------
namespace disc_ps {
// libDiscPS/disc_sample.cpp:123-138 pulls in the sampler's dependencies; this is the same C-order unravel
void index_from_flat3(int shape0, int shape1, int shape2, int flat_idx, int &idx1, int &idx2, int &idx3) {
  (void)shape0;
  idx1 = flat_idx / (shape1 * shape2);
  flat_idx = flat_idx % (shape1 * shape2);
  idx2 = flat_idx / shape2;
  idx3 = flat_idx % shape2;
}
}  // namespace disc_ps

void bbox_from_pos(const ExpParam &, const PartWindowParam::PartParam &, int, int, int, int, PartBBox &) { abort(); }
void bbox_from_pos(const PartWindowParam::PartParam &, double, double, int, int, PartBBox &) { abort(); }

#ifndef PS_REF_WITH_AUX
namespace object_detect {
// objectdetect_aux.cpp is not part of this build: the drivers are exercised with roi_save_num_samples = 0, where the
// local-maximum lists stay empty (the oracle's findLocalMax is checked elsewhere)
void findLocalMax(const ExpParam &, const FloatGrid3 &, std::vector<PartHyp> &part_hyp, int) { part_hyp.clear(); }
}  // namespace object_detect
#endif

namespace matlab_io {
capture_fn g_capture = 0;
}

namespace {
using boost_math::double_matrix;
using boost_math::double_vector;
using object_detect::Joint;

// marginals handed to mat_save_multi_array: log_part_posterior_final_imgidx<i>_scaleidx<s>_o<f>_pidx<p>.mat
float *g_marg = 0;
size_t g_marg_grid = 0;
int g_marg_P = 0;
void capture(const char *file, const char *var, const void *data, size_t bytes) {
  if (!g_marg || strcmp(var, "log_prob_grid") != 0 || bytes != g_marg_grid * sizeof(float)) return;
  const char *ps = strstr(file, "_scaleidx"), *pp = strstr(file, "_pidx");
  if (!ps || !pp) return;
  const int s = atoi(ps + 9), p = atoi(pp + 5);
  memcpy(g_marg + ((size_t)s * g_marg_P + p) * g_marg_grid, data, bytes);
}

ExpParam make_ep(const double *ep /* min_rot, max_rot, R, min_scale, max_scale, S, strip_border, K */) {
  ExpParam e;
  e.set_min_part_rotation((float)ep[0]);
  e.set_max_part_rotation((float)ep[1]);
  e.set_num_rotation_steps((uint32_t)ep[2]);
  e.set_min_object_scale((float)ep[3]);
  e.set_max_object_scale((float)ep[4]);
  e.set_num_scale_steps((uint32_t)ep[5]);
  e.set_strip_border_detections((float)ep[6]);
  e.set_roi_save_num_samples((float)ep[7]);
  return e;
}
FloatGrid3 grid3(const float *p, int R, int H, int W) {
  FloatGrid3 g(boost::extents[R][H][W]);
  memcpy(g.data(), p, sizeof(float) * (size_t)R * H * W);
  return g;
}
}  // namespace

extern "C" {

// object_detect::computeRotJointMarginal (objectdetect_findrot.cpp:292-456), the reference's code
void refd_message(const double *ep, const float *child, float *parent, int R, int H, int W, const double *off_c,
                  const double *off_p, const double *C, double rot_mean, double rot_sigma, double scale, int sparse) {
  ExpParam e = make_ep(ep);
  FloatGrid3 gc = grid3(child, R, H, W), gp(boost::extents[R][H][W]);
  double_vector oc(2), op(2);
  oc(0) = off_c[0]; oc(1) = off_c[1];
  op(0) = off_p[0]; op(1) = off_p[1];
  double_matrix Cm(2, 2);
  Cm(0, 0) = C[0]; Cm(0, 1) = C[1]; Cm(1, 0) = C[2]; Cm(1, 1) = C[3];
  object_detect::computeRotJointMarginal(e, gc, gp, oc, op, Cm, rot_mean, rot_sigma, scale, sparse != 0);
  memcpy(parent, gp.data(), sizeof(float) * (size_t)R * H * W);
}

// object_detect::computeRootPosteriorRot (:470-727) with computePartMarginals (:124-286), the reference's code.
// joints: rows of 13 doubles (type, child, parent, offset_c[2], offset_p[2], C[4], rot_mean, rot_sigma), 0-based ids.
// unaries [P][S][R][H][W] are masked in place like the reference does; best_conf [P][7] = best_part_hyp[p][0].toVect().
void refd_infer(const double *ep, int P, const int *is_detect, const int *is_upright, int rootpart_idx, const double *joints,
                int J, int H, int W, float *unaries, int sparse, float *root_post /*[S][H][W]*/, float *best_conf,
                float *marginals /*[S][P][R][H][W] or null*/) {
  PartApp app;
  app.m_exp_param = make_ep(ep);
  app.m_rootpart_idx = rootpart_idx;
  app.m_part_conf.parts_.resize(P);
  for (int p = 0; p < P; ++p) {
    app.m_part_conf.parts_[p].is_detect_ = is_detect[p] != 0;
    app.m_part_conf.parts_[p].is_upright_ = is_upright[p] != 0;
    app.m_part_conf.parts_[p].is_root_ = p == rootpart_idx;
  }
  const int S = (int)app.m_exp_param.num_scale_steps(), R = (int)app.m_exp_param.num_rotation_steps();
  const size_t G = (size_t)R * H * W;
  std::vector<std::vector<FloatGrid3> > det(P, std::vector<FloatGrid3>(S, FloatGrid3(boost::extents[R][H][W])));
  for (int p = 0; p < P; ++p)
    for (int s = 0; s < S; ++s) memcpy(det[p][s].data(), unaries + ((size_t)p * S + s) * G, sizeof(float) * G);
  std::vector<Joint> js(J);
  for (int j = 0; j < J; ++j) {
    const double *q = joints + (size_t)j * 13;
    js[j].type = (int)q[0];
    js[j].child_idx = (int)q[1];
    js[j].parent_idx = (int)q[2];
    js[j].offset_c.resize(2); js[j].offset_p.resize(2); js[j].C.resize(2, 2);
    js[j].offset_c(0) = q[3]; js[j].offset_c(1) = q[4];
    js[j].offset_p(0) = q[5]; js[j].offset_p(1) = q[6];
    js[j].C(0, 0) = q[7]; js[j].C(0, 1) = q[8]; js[j].C(1, 0) = q[9]; js[j].C(1, 1) = q[10];
    js[j].rot_mean = q[11];
    js[j].rot_sigma = q[12];
  }
  FloatGrid3 root;
  std::vector<std::vector<object_detect::PartHyp> > best;
  g_marg = marginals;
  g_marg_grid = G;
  g_marg_P = P;
  matlab_io::g_capture = marginals ? capture : 0;
  object_detect::computeRootPosteriorRot(app, det, root, rootpart_idx, js, false, sparse != 0, 0, best, marginals != 0);
  matlab_io::g_capture = 0;
  memcpy(root_post, root.data(), sizeof(float) * root.num_elements());
  for (int p = 0; p < P; ++p) {
    FloatGrid1 v = best[p][0].toVect();
    for (int k = 0; k < 7; ++k) best_conf[p * 7 + k] = v[k];
    for (int s = 0; s < S; ++s) memcpy(unaries + ((size_t)p * S + s) * G, det[p][s].data(), sizeof(float) * G);
  }
}

}  // extern "C"
