// ref_eval.cpp -- TEST INFRASTRUCTURE: C entry points over the reference's OWN evaluator code.
// libPartEval/parteval.cpp (get_bbox_endpoints, is_gt_match, the three bbox_merge variants, vis_eval_helper's model-part
// -> evaluation-part conversion, eval_segments) and libPartDetect/partdef.cpp (get_part_bbox and its helpers) are
// compiled UNMODIFIED from /root/reference next to this file (`make -C oracle ref`) against the stand-ins of
// oracle/ref_shim/ (Qt drawing classes that do nothing, plain-data protobuf messages, an annotation list, a matlab_io
// whose loader is a hook).  partapp_b200/parteval.py is held to them by tests/test_eval_vs_ref.py.
//
// Not compiled: libPartApp/partapp.cpp (its includes need the protobuf runtime).  The two small functions of it that the
// evaluator links against are defined here: bbox_from_pos (partapp.cpp:59-81, restated) and complete_relative_path
// (only reached from the visualisation entry points, which are never called).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <libPartApp/partapp.h>
#include <libPartApp/partapp_aux.hpp>
#include <libPartEval/parteval.h>
#include <libPictStruct/objectdetect.h>

// ---- declarations of parteval.cpp (its header is not included: the stand-in libPartEval/parteval.h is empty) -----------
const int EVAL_TYPE_PS = 1;  // parteval.h:30
void get_bbox_endpoints(PartBBox &bbox, boost_math::double_vector &endpoint_top, boost_math::double_vector &endpoint_bottom,
                        double &seg_len);
bool is_gt_match(PartBBox &gt_bbox, PartBBox &detect_bbox, bool match_x_axis, float factor);
PartBBox bbox_merge(const PartBBox &bbox1, const PartBBox &bbox2, const PartBBox &bbox3, const PartBBox &bbox4);
PartBBox bbox_merge(const PartBBox &bbox1, const PartBBox &bbox2);
PartBBox bbox_merge(const PartBBox &bbox1, const PartBBox &bbox2, const ExpParam &exp_param);
void vis_eval_helper(const PartApp &part_app, int imgidx, int scaleidx, const int eval_type, const PartConfig &part_conf_eval,
                     std::vector<PartBBox> &eval_bbox, QString qsHypDirName, int didx, bool bMerge);

// ---- symbols of translation units that are not compiled here -----------------------------------------------------------
// partapp.cpp:59-81 (restated: position, axes from the rotation, window box scaled and shifted by the part offset)
void bbox_from_pos(const PartWindowParam::PartParam &part_param, double scale, double rot, int ix, int iy, PartBBox &bbox) {
  bbox.part_pos(0) = ix;
  bbox.part_pos(1) = iy;
  bbox.part_x_axis(0) = cos(rot);
  bbox.part_x_axis(1) = sin(rot);
  bbox.part_y_axis(0) = -bbox.part_x_axis(1);
  bbox.part_y_axis(1) = bbox.part_x_axis(0);
  const double rect_width = scale * part_param.window_size_x(), rect_height = scale * part_param.window_size_y();
  bbox.min_proj_x = -scale * part_param.pos_offset_x();
  bbox.min_proj_y = -scale * part_param.pos_offset_y();
  bbox.max_proj_x = bbox.min_proj_x + rect_width;
  bbox.max_proj_y = bbox.min_proj_y + rect_height;
}
// partapp.cpp:47-57
void bbox_from_pos(const ExpParam &exp_param, const PartWindowParam::PartParam &part_param, int scaleidx, int rotidx, int ix,
                   int iy, PartBBox &bbox) {
  const double scale = scale_from_index(exp_param, scaleidx);
  const double rot = rot_from_index(exp_param, rotidx) / 180.0 * M_PI;
  bbox_from_pos(part_param, scale, rot, ix, iy, bbox);
}
QString complete_relative_path(QString, QString) { abort(); }

namespace matlab_io {
capture_fn g_capture = 0;
load2d_fn g_load2d = 0;
}  // namespace matlab_io

namespace {
std::vector<float> g_best_conf;
int g_rows = 0, g_cols = 0;
bool provide(const char *, const char *var, int *rows, int *cols, const float **data) {
  if (strcmp(var, "best_conf") != 0) return false;
  *rows = g_rows;
  *cols = g_cols;
  *data = g_best_conf.data();
  return true;
}
// bbox <-> 10 doubles: pos x y | x axis | y axis | min_proj_x max_proj_x min_proj_y max_proj_y
PartBBox unpack(const double *v) {
  PartBBox b;
  b.part_pos(0) = v[0]; b.part_pos(1) = v[1];
  b.part_x_axis(0) = v[2]; b.part_x_axis(1) = v[3];
  b.part_y_axis(0) = v[4]; b.part_y_axis(1) = v[5];
  b.min_proj_x = v[6]; b.max_proj_x = v[7]; b.min_proj_y = v[8]; b.max_proj_y = v[9];
  return b;
}
void pack(const PartBBox &b, double *v) {
  v[0] = b.part_pos(0); v[1] = b.part_pos(1);
  v[2] = b.part_x_axis(0); v[3] = b.part_x_axis(1);
  v[4] = b.part_y_axis(0); v[5] = b.part_y_axis(1);
  v[6] = b.min_proj_x; v[7] = b.max_proj_x; v[8] = b.min_proj_y; v[9] = b.max_proj_y;
}
// part definition from flat arrays: pos ids | from ids | to ids, then offset and the four extensions
PartDef make_partdef(const int *pos, int npos, const int *from, int nfrom, const int *to, int nto, const double *f5) {
  PartDef d;
  d.part_pos_.assign(pos, pos + npos);
  d.part_x_axis_from_.assign(from, from + nfrom);
  d.part_x_axis_to_.assign(to, to + nto);
  d.part_x_axis_offset_ = (float)f5[0];
  d.ext_x_pos_ = (float)f5[1]; d.ext_x_neg_ = (float)f5[2]; d.ext_y_pos_ = (float)f5[3]; d.ext_y_neg_ = (float)f5[4];
  return d;
}
}  // namespace

extern "C" {

// parteval.cpp:45-68 (use_endpoints == false): out = top x y | bottom x y | seg_len
void refe_bbox_endpoints(const double *bbox, double *out) {
  PartBBox b = unpack(bbox);
  boost_math::double_vector t, m;
  double len = 0;
  get_bbox_endpoints(b, t, m, len);
  out[0] = t(0); out[1] = t(1); out[2] = m(0); out[3] = m(1); out[4] = len;
}

// parteval.cpp:79-126
int refe_is_gt_match(const double *gt, const double *det, int match_x_axis, float factor) {
  PartBBox g = unpack(gt), d = unpack(det);
  return is_gt_match(g, d, match_x_axis != 0, factor) ? 1 : 0;
}

// parteval.cpp:208-306: kind 2 = two boxes, 4 = four boxes, 3 = two boxes with the rotation range of exp_param
void refe_bbox_merge(int kind, const double *boxes, float min_rot, float max_rot, int num_rot, double *out) {
  PartBBox r;
  if (kind == 4) r = bbox_merge(unpack(boxes), unpack(boxes + 10), unpack(boxes + 20), unpack(boxes + 30));
  else if (kind == 2) r = bbox_merge(unpack(boxes), unpack(boxes + 10));
  else {
    ExpParam ep;
    ep.set_min_part_rotation(min_rot);
    ep.set_max_part_rotation(max_rot);
    ep.set_num_rotation_steps((uint32_t)num_rot);
    r = bbox_merge(unpack(boxes), unpack(boxes + 10), ep);
  }
  pack(r, out);
}

// partdef.cpp:91-365: get_part_bbox of one annotated person.  Returns -1 if the rectangle lacks a point of the part
// (annorect_has_part), 0 if the axis is degenerate, 1 with the box in out.
int refe_get_part_bbox(const int *ids, const int *xs, const int *ys, int npts, const int *pos, int npos, const int *from, int nfrom,
                       const int *to, int nto, const double *f5, double scale, double *out) {
  AnnoRect rect;
  for (int i = 0; i < npts; ++i) {
    AnnoPoint p;
    p.id = ids[i]; p.x = xs[i]; p.y = ys[i];
    rect.m_vAnnoPoints.push_back(p);
  }
  const PartDef d = make_partdef(pos, npos, from, nfrom, to, nto, f5);
  if (!annorect_has_part(rect, d)) return -1;
  PartBBox b;
  if (!get_part_bbox(rect, d, b, scale)) return 0;
  pack(b, out);
  return 1;
}

// vis_eval_helper (parteval.cpp:312-853) for EVAL_TYPE_PS: best_conf [P][7] as findObjectImageRotJoints writes it ->
// evaluation boxes.  window [P][4] = window_size_x/y, pos_offset_x/y; ext [P][4] = ext_x_pos, ext_x_neg, ext_y_pos,
// ext_y_neg of the MODEL part_conf; ext_eval [Pe][4] the same of part_conf_eval.  Returns the number of boxes.
int refe_vis_eval_helper(const char *part_conf_type, const float *best_conf, int P, const int *window, const double *ext,
                         const double *ext_eval, int Pe, float min_rot, float max_rot, int num_rot, double *out, int cap) {
  PartApp app;
  app.m_exp_param.set_part_conf_type(part_conf_type);
  app.m_exp_param.set_min_part_rotation(min_rot);
  app.m_exp_param.set_max_part_rotation(max_rot);
  app.m_exp_param.set_num_rotation_steps((uint32_t)num_rot);
  app.m_exp_param.set_log_dir(".");
  app.m_exp_param.set_log_subdir("ref");
  for (int p = 0; p < P; ++p) {
    PartDef d;
    d.ext_x_pos_ = (float)ext[p * 4]; d.ext_x_neg_ = (float)ext[p * 4 + 1];
    d.ext_y_pos_ = (float)ext[p * 4 + 2]; d.ext_y_neg_ = (float)ext[p * 4 + 3];
    app.m_part_conf.parts_.push_back(d);
    PartWindowParam::PartParam w;
    w.window_size_x_ = window[p * 4]; w.window_size_y_ = window[p * 4 + 1];
    w.pos_offset_x_ = window[p * 4 + 2]; w.pos_offset_y_ = window[p * 4 + 3];
    app.m_window_param.parts_.push_back(w);
  }
  PartConfig eval_conf;
  for (int p = 0; p < Pe; ++p) {
    PartDef d;
    d.ext_x_pos_ = (float)ext_eval[p * 4]; d.ext_x_neg_ = (float)ext_eval[p * 4 + 1];
    d.ext_y_pos_ = (float)ext_eval[p * 4 + 2]; d.ext_y_neg_ = (float)ext_eval[p * 4 + 3];
    eval_conf.parts_.push_back(d);
  }
  g_best_conf.assign(best_conf, best_conf + (size_t)P * 7);
  g_rows = P;
  g_cols = 7;
  matlab_io::g_load2d = provide;
  std::vector<PartBBox> boxes;
  vis_eval_helper(app, 0, -1, EVAL_TYPE_PS, eval_conf, boxes, QString("hyp"), -1, true);
  matlab_io::g_load2d = 0;
  const int n = (int)boxes.size();
  for (int i = 0; i < n && i < cap; ++i) pack(boxes[(size_t)i], out + 10 * i);
  return n;
}

}  // extern "C"
