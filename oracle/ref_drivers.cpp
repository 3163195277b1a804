// ref_drivers.cpp -- TEST INFRASTRUCTURE: C entry points over the reference's OWN inference drivers.
// libPictStruct/objectdetect_findrot.cpp (computeRotJointMarginal, computePartMarginals, computeRootPosteriorRot) is
// libPictStruct/objectdetect_aux.cpp (findLocalMax, loadJoints), objectdetect_icps.cpp (the conditioning adds) and
// objectdetect_findpos.cpp (computePosJointMarginal, the legacy POS_GAUSSIAN message) are
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

// libPartDetect/partdef.cpp (ground-truth part boxes from annotations) is not compiled; objectdetect_icps.cpp only reaches
// these from its training-time helpers
bool annorect_has_part(const AnnoRect &, const PartDef &) { abort(); }
bool get_part_bbox(const AnnoRect &, const PartDef &, PartBBox &, double) { abort(); }
void bbox_from_pos(const ExpParam &, const PartWindowParam::PartParam &, int, int, int, int, PartBBox &) { abort(); }
void bbox_from_pos(const PartWindowParam::PartParam &, double, double, int, int, PartBBox &) { abort(); }

// libPictStruct/objectdetect_aux.cpp (findLocalMax, loadJoints, ...) is compiled next to this file as well; what it
// references from translation units that read files or run the detector is defined here.
namespace {
std::vector<object_detect::Joint> g_joint_table;  // what load_joint "reads": set by refd_load_joints
}
namespace object_detect {
// objectdetect_learnparam.cpp:92-179 reads joint_<c>_<p>.mat; the table stands in for the files
void load_joint(const PartApp &, int jidx, Joint &joint, int) { joint = g_joint_table.at((size_t)jidx); }
#ifdef PS_REF_WITHOUT_ICPS
int predictFactors(const PartApp &, int, int) { abort(); }
// fallback when objectdetect_icps.cpp is left out of the build
typedef std::vector<std::vector<FloatGrid3> > Grids;
void loadDPMScoreGrid(QString, int, std::vector<FloatGrid2> &, bool) { abort(); }
void getRotParams(const PartApp &, int, boost_math::double_matrix &, bool) { abort(); }
void getPosParams(const PartApp &, int, boost_math::double_matrix &, int, bool) { abort(); }
void getTorsoPosPriorParams(const PartApp &, int, int, boost_math::double_matrix &) { abort(); }
void addDPMScore(const PartApp &, Grids &, std::vector<FloatGrid2>, int, float) { abort(); }
void addLoadDPMScore(const PartApp &, Grids &, int, float, int, QString, bool, int) { abort(); }
void getRotScoreGrid(const PartApp &, Grids &, boost_math::double_matrix &) { abort(); }
void addExtraUnary(const PartApp &, Grids &, const Grids &, float) { abort(); }
void getRootPosDet(const PartApp &, int, int, boost_math::double_vector &, bool) { abort(); }
void getPosScoreGrid(const PartApp &, Grids &, int, boost_math::double_matrix &, int, boost_math::double_vector &) { abort(); }
void setTorsoPosPrior(const PartApp &, Grids &, boost_math::double_matrix &, int) { abort(); }
#endif
// objectdetect_findpos.cpp defines these but declares them in no header this file sees
void mergeRotations(const PartApp &part_app, const PartConfig &part_conf,
                    const std::vector<std::vector<std::vector<FloatGrid2> > > &part_score_grid_rotation,
                    std::vector<std::vector<FloatGrid2> > &log_part_detections);
void computeRootPosterior(const PartApp part_app, const std::vector<std::vector<FloatGrid2> > &log_part_detections,
                          FloatGrid3 &root_part_posterior, int rootpart_idx, std::vector<Joint> joints, bool flip,
                          QString qsDebugDir, bool bIsSparse);
void computePosJointMarginal(FloatGrid2 &log_prob_child, FloatGrid2 &log_prob_parent, boost_math::double_vector offset,
                             boost_math::double_matrix C, double scale, bool bIsSparse);
}  // namespace object_detect

namespace matlab_io {
capture_fn g_capture = 0;
load2d_fn g_load2d = 0;  // no file provider in this library: the loaders abort
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

namespace {
Joint joint_from_row(const double *q) {
  Joint j;
  j.type = (int)q[0];
  j.child_idx = (int)q[1];
  j.parent_idx = (int)q[2];
  j.offset_c.resize(2); j.offset_p.resize(2); j.C.resize(2, 2);
  j.offset_c(0) = q[3]; j.offset_c(1) = q[4];
  j.offset_p(0) = q[5]; j.offset_p(1) = q[6];
  j.C(0, 0) = q[7]; j.C(0, 1) = q[8]; j.C(1, 0) = q[9]; j.C(1, 1) = q[10];
  j.rot_mean = q[11];
  j.rot_sigma = q[12];
  return j;
}
void joint_to_row(const Joint &j, double *q) {
  q[0] = j.type; q[1] = j.child_idx; q[2] = j.parent_idx;
  q[3] = j.offset_c(0); q[4] = j.offset_c(1); q[5] = j.offset_p(0); q[6] = j.offset_p(1);
  q[7] = j.C(0, 0); q[8] = j.C(0, 1); q[9] = j.C(1, 0); q[10] = j.C(1, 1);
  q[11] = j.rot_mean; q[12] = j.rot_sigma;
}
}  // namespace

extern "C" {

// object_detect::findLocalMax (objectdetect_aux.cpp:193-261): rows of (dim0, x, y, score) as doubles
int refd_find_local_max(const float *grid, int D0, int H, int W, int max_n, double *out, int cap) {
  FloatGrid3 g = grid3(grid, D0, H, W);
  std::vector<double_vector> lm;
  object_detect::findLocalMax(g, lm, max_n);
  if ((int)lm.size() > cap) return -(int)lm.size();
  for (size_t i = 0; i < lm.size(); ++i)
    for (int k = 0; k < 4; ++k) out[i * 4 + k] = lm[i](k);
  return (int)lm.size();
}

// object_detect::loadJoints (objectdetect_aux.cpp:54-141) over a table of joints as load_joint would deliver them
// (1-based part ids in child/parent); rows of 13 doubles in and out, plus detC and invC (5 more doubles) out.
void refd_load_joints(int P, const double *table, int J, int flip, double *out13, double *out_det_inv5) {
  PartApp app;
  app.m_part_conf.parts_.resize(P);
  for (int p = 0; p < P; ++p) app.m_part_conf.parts_[p].part_id_ = p + 1;
  app.m_part_conf.joints_.resize(J);
  g_joint_table.clear();
  for (int j = 0; j < J; ++j) g_joint_table.push_back(joint_from_row(table + (size_t)j * 13));
  std::vector<Joint> joints;
  object_detect::loadJoints(app, joints, flip != 0, -1, true);
  for (int j = 0; j < J; ++j) {
    joint_to_row(joints[j], out13 + (size_t)j * 13);
    double *d = out_det_inv5 + (size_t)j * 5;
    d[0] = joints[j].detC;
    d[1] = joints[j].invC(0, 0); d[2] = joints[j].invC(0, 1); d[3] = joints[j].invC(1, 0); d[4] = joints[j].invC(1, 1);
  }
}

// The conditioning adds of objectdetect_icps.cpp on unaries [P][S][R][H][W] (in place), the reference's code:
//   kind 0: getRotScoreGrid (:228-281) with rot_params [P][2], then addExtraUnary (:526-548) with `weight`
//   kind 1: getPosScoreGrid (:366-423) with pos_params [P][4] and rootpos_det = params[4P .. 4P+1], then addExtraUnary
//   kind 2: setTorsoPosPrior (:137-191) with pos_prior_params [1][4]; `weight` = ExpParam.torso_pos_prior_weight
//   kind 3: addDPMScore (:488-524) on part `pidx` with n_dpm grids [n_dpm][H][W] and `weight`
void refd_condition(const double *ep, int P, const int *is_detect, int rootpart_idx, int H, int W, float *unaries, int kind,
                    const double *params, float weight, int pidx, const float *dpm, int n_dpm) {
  PartApp app;
  app.m_exp_param = make_ep(ep);
  app.m_exp_param.set_torso_pos_prior_weight(weight);
  app.m_rootpart_idx = rootpart_idx;
  app.m_part_conf.parts_.resize(P);
  for (int p = 0; p < P; ++p) app.m_part_conf.parts_[p].is_detect_ = is_detect[p] != 0;
  const int S = (int)app.m_exp_param.num_scale_steps(), R = (int)app.m_exp_param.num_rotation_steps();
  const size_t G = (size_t)R * H * W;
  typedef std::vector<std::vector<FloatGrid3> > Grids;
  Grids det(P, std::vector<FloatGrid3>(S, FloatGrid3(boost::extents[R][H][W])));
  for (int p = 0; p < P; ++p)
    for (int s = 0; s < S; ++s) memcpy(det[p][s].data(), unaries + ((size_t)p * S + s) * G, sizeof(float) * G);
  if (kind == 0 || kind == 1) {
    Grids extra(P, std::vector<FloatGrid3>(S, FloatGrid3(boost::extents[R][H][W])));  // zero-filled, findrot.cpp:917-932
    double_matrix prm(P, kind == 0 ? 2 : 4);
    for (int p = 0; p < P; ++p)
      for (size_t k = 0; k < prm.size2(); ++k) prm(p, k) = params[(size_t)p * prm.size2() + k];
    if (kind == 0) {
      object_detect::getRotScoreGrid(app, extra, prm);
    } else {
      double_vector root(2);
      root(0) = params[(size_t)P * 4];
      root(1) = params[(size_t)P * 4 + 1];
      object_detect::getPosScoreGrid(app, extra, 0, prm, rootpart_idx, root);
    }
    object_detect::addExtraUnary(app, det, extra, weight);
  } else if (kind == 2) {
    double_matrix prm(1, 4);
    for (int k = 0; k < 4; ++k) prm(0, k) = params[k];
    object_detect::setTorsoPosPrior(app, det, prm, rootpart_idx);
  } else {
    std::vector<FloatGrid2> g(n_dpm, FloatGrid2(boost::extents[H][W]));
    for (int i = 0; i < n_dpm; ++i) memcpy(g[i].data(), dpm + (size_t)i * H * W, sizeof(float) * (size_t)H * W);
    object_detect::addDPMScore(app, det, g, pidx, weight);
  }
  for (int p = 0; p < P; ++p)
    for (int s = 0; s < S; ++s) memcpy(unaries + ((size_t)p * S + s) * G, det[p][s].data(), sizeof(float) * G);
}

// object_detect::computePosJointMarginal (objectdetect_findpos.cpp:64-89), the reference's code, on one [H][W] grid;
// `child` comes back as the reference leaves it
void refd_pos_message(float *child, float *parent, int H, int W, const double *offset, const double *C, double scale, int sparse) {
  FloatGrid2 gc(boost::extents[H][W]), gp(boost::extents[H][W]);
  memcpy(gc.data(), child, sizeof(float) * (size_t)H * W);
  double_vector off(2);
  off(0) = offset[0]; off(1) = offset[1];
  double_matrix Cm(2, 2);
  Cm(0, 0) = C[0]; Cm(0, 1) = C[1]; Cm(1, 0) = C[2]; Cm(1, 1) = C[3];
  object_detect::computePosJointMarginal(gc, gp, off, Cm, scale, sparse != 0);
  memcpy(child, gc.data(), sizeof(float) * (size_t)H * W);
  memcpy(parent, gp.data(), sizeof(float) * (size_t)H * W);
}

// The legacy POS_GAUSSIAN driver (objectdetect_findpos.cpp): mergeRotations (:118-170) of log-domain unaries
// [P][S][R][H][W], then computeRootPosterior (:172-334).  Outputs: merged [P][S][H][W], root posterior [S][H][W].
void refd_root_posterior_pos(const double *ep, int P, const int *is_detect, int root_idx, const double *joints, int nj, int H,
                             int W, const float *unaries, int sparse, float *merged_out, float *root_post) {
  PartApp app;
  app.m_exp_param = make_ep(ep);
  const int S = (int)app.m_exp_param.num_scale_steps(), R = (int)app.m_exp_param.num_rotation_steps();
  for (int p = 0; p < P; ++p) {
    PartDef d;
    d.is_detect_ = is_detect[p] != 0;
    d.is_root_ = p == root_idx;
    app.m_part_conf.parts_.push_back(d);
  }
  std::vector<Joint> js;
  for (int j = 0; j < nj; ++j) js.push_back(joint_from_row(joints + 13 * j));
  std::vector<std::vector<std::vector<FloatGrid2> > > rot(P);
  for (int p = 0; p < P; ++p) {
    if (!is_detect[p]) continue;
    rot[p].resize(S);
    for (int s = 0; s < S; ++s)
      for (int r = 0; r < R; ++r) {
        FloatGrid2 g(boost::extents[H][W]);
        memcpy(g.data(), unaries + ((((size_t)p * S + s) * R + r) * H) * W, sizeof(float) * (size_t)H * W);
        rot[p][s].push_back(g);
      }
  }
  std::vector<std::vector<FloatGrid2> > merged(P);
  object_detect::mergeRotations(app, app.m_part_conf, rot, merged);
  for (int p = 0; p < P; ++p)
    for (int s = 0; s < S; ++s)
      memcpy(merged_out + ((size_t)p * S + s) * H * W, merged[p][s].data(), sizeof(float) * (size_t)H * W);
  FloatGrid3 rp;
  object_detect::computeRootPosterior(app, merged, rp, root_idx, js, false, QString("debug"), sparse != 0);
  memcpy(root_post, rp.data(), sizeof(float) * (size_t)S * H * W);
}

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
                float *marginals /*[S][P][R][H][W] or null*/, float *hyps /*[P][cap][7] or null*/, int cap, int *nhyps) {
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
  for (int j = 0; j < J; ++j) js[j] = joint_from_row(joints + (size_t)j * 13);
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
    if (hyps) {
      nhyps[p] = (int)std::min<size_t>(best[p].size(), (size_t)cap);
      for (int i = 0; i < nhyps[p]; ++i) {
        FloatGrid1 h = best[p][i].toVect();
        for (int k = 0; k < 7; ++k) hyps[((size_t)p * cap + i) * 7 + k] = h[k];
      }
    }
    for (int s = 0; s < S; ++s) memcpy(unaries + ((size_t)p * S + s) * G, det[p][s].data(), sizeof(float) * G);
  }
}

}  // extern "C"
