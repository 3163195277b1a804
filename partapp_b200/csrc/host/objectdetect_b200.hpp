// objectdetect_b200.hpp -- C++ host side of the `--find_obj` path, mirroring the reference's interface
// (src/libs/libPictStruct/objectdetect.h, src/libs/libPartApp/partapp.h) on top of the C ABI of include/psinfer.h.
//
// Same names, argument meaning and file layout as the reference; differences are listed where they occur:
//   * grids are `FloatGrid3` = contiguous C-order fp32 [rotation][y][x] (boost::multi_array<float,3> there);
//   * `assert`s become std::runtime_error;
//   * nothing is computed on the CPU: every grid operation goes through libpsinfer.so.
#pragma once

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/psinfer.h"

namespace object_detect {

struct FloatGrid3 {
  int R = 0, H = 0, W = 0;
  std::vector<float> v;
  FloatGrid3() {}
  FloatGrid3(int r, int h, int w) : R(r), H(h), W(w), v((size_t)r * h * w, 0.0f) {}
  float *data() { return v.data(); }
  const float *data() const { return v.data(); }
  size_t num_elements() const { return v.size(); }
};

// The ExpParam fields this path reads (libPartApp/ExpParam.proto); defaults are the proto's.
struct ExpParam {
  std::vector<std::string> test_dataset;
  std::string log_dir, log_subdir, class_dir, scoregrid_dir, spatial_dir, part_conf, pred_data_test_dir;
  float min_object_scale = 1, max_object_scale = 1;
  unsigned num_scale_steps = 1;
  float min_part_rotation = -180, max_part_rotation = 180;
  unsigned num_rotation_steps = 48;
  bool flip_orientation = false;
  int num_pose_samples = 0;
  float strip_border_detections = 0;
  float roi_save_num_samples = 1000;
  bool use_pairwise = true, save_part_marginals = false, save_part_marginals_local_max = false;
  bool save_part_detections_local_max = false, interpolate = false, force_recompute_scores = true;
  bool use_torso_pos_prior = false, save_root_marginal = false;
  float torso_pos_prior_weight = 1;
  // conditioning of the unaries (findrot.cpp:849-949).  The predictors themselves are MATLAB / DPM runs of the
  // reference (objectdetect_icps.cpp:608-625 spawns them); this host reads the files they write.
  bool pred_unary_rot = false, pred_unary_pos = false, use_dpm_torso = false, use_dpm_head = false, use_dpm_unary = false;
  bool do_dpm_rot = false, use_gt_torso = false;
  float pred_unary_rot_weight = 1, pred_unary_pos_weight = 1, dpm_torso_weight = 1, dpm_head_weight = 1, dpm_unary_weight = 1;
  unsigned rootidx_det = 1000;
  std::string torso_det_test_dir, test_dpm_torso_dir, test_dpm_unary_dir;
};

struct PartDef {   // libPartDetect/PartConfig.proto PartDef
  int part_id = 0;
  bool is_root = false, is_detect = true, is_upright = false;
  std::vector<int> part_pos;  // annopoint ids that define the part's position (read by use_gt_torso only)
};
// The annopoints of an image's FIRST annotated rectangle (libAnnotation AnnoRect::get_annopoint_by_id: first match wins)
struct AnnoPoint {
  int id = 0, x = 0, y = 0;
};
struct JointDef {  // PartConfig.proto Joint
  int child_idx = 0, parent_idx = 0;
  std::string type = "Gaussian";
  unsigned num_joint_types = 1;
};
struct PartConfig {
  std::vector<PartDef> part;
  std::vector<JointDef> joint;
};
struct PartWindowParam {  // only the fields findObjectImageRotJoints reads (findrot.cpp:1044-1045)
  double bbox_offset_x = 0, bbox_offset_y = 0;
};

// libPartApp/partapp.h PartApp, reduced to what --find_obj touches
struct PartApp {
  ExpParam m_exp_param;
  PartConfig m_part_conf;
  PartWindowParam m_window_param;
  std::vector<std::string> m_test_annolist;  // image file names (AnnotationList::imageName())
  std::vector<std::vector<AnnoPoint> > m_test_annopoints;  // per image: m_test_annolist[imgidx][0].m_vAnnoPoints (.al lists only)
  int m_rootpart_idx = -1;
  // PartApp::init (partapp.cpp:141) + init_setpath (:294): parse the expopt, resolve relative paths against it,
  // fill default directories, load part_conf, window_param.txt (if present) and the test image list.
  void init(const std::string &expopt_file);
  // partapp.cpp:792-799
  std::string getScoreGridFileName(int imgidx, int pidx, bool flip) const;
};

// objectdetect.h:54-86
struct Joint {
  enum { POS_GAUSSIAN = 1, ROT_GAUSSIAN = 2 };
  int type = 0, mix_comp_id = 0, child_idx = 0, parent_idx = 0;
  double offset_c[2] = {0, 0}, offset_p[2] = {0, 0};
  double C[2][2] = {{0, 0}, {0, 0}};
  double rot_mean = 0, rot_sigma = 0, detC = 0;
  double invC[2][2] = {{0, 0}, {0, 0}};
};

// objectdetect.h:88-195
struct PartHyp {
  int m_imgidx = -1, m_scaleidx = -1, m_rotidx = -1, m_x = -1, m_y = -1;
  float m_score = -1e6f, m_scale = 0, m_rot = 0;
  static unsigned vectSize() { return 7; }
  void toVect(float *r) const {
    r[0] = (float)m_scaleidx; r[1] = m_scale; r[2] = (float)m_rotidx; r[3] = m_rot;
    r[4] = (float)m_x; r[5] = (float)m_y; r[6] = m_score;
  }
  void fromVect(const float *r) {
    m_scaleidx = (int)r[0]; m_scale = r[1]; m_rotidx = (int)r[2]; m_rot = r[3];
    m_x = (int)r[4]; m_y = (int)r[5]; m_score = r[6];
  }
};

// HypothesisList.proto
struct ObjectHypothesis {
  float x = 0, y = 0, scale = 0, score = 0;
  bool flip = false;
};
struct HypothesisList {
  std::vector<ObjectHypothesis> hyp;
  std::string SerializeAsString() const;           // proto2 wire format
  static HypothesisList Parse(const std::string &);  // for tests
};

// partapp_aux.hpp
double rot_from_index(const ExpParam &, int rotidx);
double scale_from_index(const ExpParam &, int scaleidx);

// objectdetect_learnparam.cpp:92-179 / objectdetect_aux.cpp:54-141
void load_joint(const PartApp &, int jidx, Joint &, int tidx = -1);
void loadJoints(const PartApp &, std::vector<Joint> &, bool flip, int imgidx = -1);

// objectdetect_icps.cpp:193-226 / :326-363 / :283-324: the per-image predictor outputs.  rot_params [P][3] =
// (mu, var = sigma^2, cluster), pos_params [P][5] = (mu_x, mu_y, var_x, var_y, cluster); rows of parts that are not
// detected (and the root's position row) stay zero.
void getRotParams(const PartApp &, int imgidx, std::vector<double> &rot_params, bool bTest = true);
void getPosParams(const PartApp &, int imgidx, std::vector<double> &pos_params, int rootpart_idx, bool bTest = true);
void getRootPosDet(const PartApp &, int imgidx, int rootpart_idx, double rootpos_det[2], bool bTest = true);
// objectdetect_icps.cpp:550-581: <dir>/imgidx_%04d.mat (1-based), variable "scoregrid": one [H][W] grid (bIsCell false)
// or a cell array of them, one per DPM rotation (bIsCell true: `expected` grids must be there).  Grids are raw DPM
// scores, fp32, C order.
void loadDPMScoreGrid(const std::string &qsDPMdir, int imgidx, std::vector<std::vector<float> > &dpmPriorGrid, int H, int W,
                      bool bIsCell, int expected = 1);

// objectdetect_findrot.cpp:292-456
void computeRotJointMarginal(const ExpParam &, FloatGrid3 &log_prob_child, FloatGrid3 &log_prob_parent,
                             const double offset_c_10[2], const double offset_p_01[2], const double C[2][2],
                             double rot_mean, double rot_sigma, double scale, bool bIsSparse);

// objectdetect_findpos.cpp:64-89 (legacy POS_GAUSSIAN joints).  Grids are [D][H][W]; every [H][W] slice is one call of
// the reference, which works on 2-D grids.  Like there, log_prob_child comes back as log(exp(child)).
void computePosJointMarginal(const ExpParam &, FloatGrid3 &log_prob_child, FloatGrid3 &log_prob_parent,
                             const double offset[2], const double C[2][2], double scale, bool bIsSparse);

// objectdetect_findrot.cpp:470-727
void computeRootPosteriorRot(const PartApp &, std::vector<std::vector<FloatGrid3> > &log_part_detections,
                             FloatGrid3 &root_part_posterior, int rootpart_idx, std::vector<Joint> joints, bool flip,
                             bool bIsSparse, int imgidx, std::vector<std::vector<PartHyp> > &best_part_hyp,
                             bool bSaveMarginals);

// objectdetect_findpos.cpp:336-452 (legacy POS_GAUSSIAN joints): compact score grids -> unaries on the device ->
// mergeRotations (:118-170) -> computeRootPosterior (:172-334) -> findLocalMax(1000) into hypothesis_list; with
// save_root_marginal also root_part_posterior/root_part_posterior_imgidx<i>_o<f>.mat (:437-451).  ps_infer runs the whole
// chain when the joints are POS_GAUSSIAN.
void findObjectImagePosJoints(const PartApp &, int imgidx, bool flip, HypothesisList &hypothesis_list);

// objectdetect_findrot.cpp:729-1058.  Reads the compact score grids of the image straight into the device
// (PartApp::loadScoreGrid, partapp.cpp:830-903, runs as ps_set_unary_compact) and writes the reference's outputs.
void findObjectImageRotJoints(const PartApp &, int imgidx, bool flip, HypothesisList &hypothesis_list,
                              const std::string &qsPartMarginalsDir, const std::string &qsScoreGridDir,
                              const std::string &qsImgName);

// The detector responses of one part inside a region of interest: part_detect::ScoreGrid reduced to what
// findObjectRoiHelper reads (objectdetect_roi.cpp:195-207): the compact grids of every rotation and Tig = getTig().
struct ScoreGrid {
  int gh = 0, gw = 0;
  std::vector<float> cells;  // [R][gh][gw], 0 = not evaluated
  std::vector<double> Tig;   // [R][3][3]
};

// objectdetect_roi.cpp:45-278, the inference half: `roi` = (x1, y1, x2, y2) after the border extension and clamping
// (:141-148).  Computing the responses (computeDescriptorGridRoi / computeScoreGrid, :180-199) is the detector's job.
void findObjectRoiHelper(PartApp part_app, const int roi[4], double scale, const std::vector<ScoreGrid> &score_grid,
                         std::vector<Joint> joints, std::vector<std::vector<PartHyp> > &best_part_det,
                         std::vector<std::vector<PartHyp> > &best_part_hyp);

// objectdetect_aux.cpp:322-406.  Runs on `gpus` devices (contiguous image ranges per GPU, like the reference's
// --distribute shards of main.cpp:155-192) with `contexts_per_gpu` worker threads each; see set_parallelism.
void findObjectDataset(const PartApp &, int firstidx, int lastidx);
// Number of GPUs (devices PSINFER_DEVICE .. +gpus-1) and of worker threads / ps_ctx per GPU that findObjectDataset uses.
// 0 keeps the environment's PSINFER_GPUS / PSINFER_CTX_PER_GPU, else 1 x 4.
void set_parallelism(int gpus, int contexts_per_gpu);

// aux.cpp:311-320
std::string getObjectHypFilename(int imgidx, bool flip);

// image width/height from a PNG or JPEG header (the reference loads the whole image for this, findrot.cpp:752-760)
void image_size(const std::string &file, int &width, int &height);

}  // namespace object_detect
