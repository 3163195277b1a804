// Stand-in for the protoc-generated libPartApp/ExpParam.pb.h -- TEST INFRASTRUCTURE (protoc and the protobuf C++
// runtime are not installed here).  Plain data with the accessors (and field types: float ranges, uint32 counts) that
// the reference's partapp_aux.hpp and objectdetect_findrot.cpp read, so those files compile unmodified into oracle/_ref/.
#pragma once
#include <cassert>
#include <cstdint>
#include <string>
class ExpParam {
 public:
#define PS_FIELD(type, name, dflt)          \
  type name##_ = dflt;                      \
  bool has_##name##_ = false;               \
  type name() const { return name##_; }     \
  bool has_##name() const { return has_##name##_; } \
  void set_##name(type v) { name##_ = v; has_##name##_ = true; }
  PS_FIELD(float, min_object_scale, 1)
  PS_FIELD(float, max_object_scale, 1)
  PS_FIELD(uint32_t, num_scale_steps, 1)
  PS_FIELD(float, min_part_rotation, -180)
  PS_FIELD(float, max_part_rotation, 180)
  PS_FIELD(uint32_t, num_rotation_steps, 48)
  PS_FIELD(float, strip_border_detections, 0)
  PS_FIELD(float, roi_save_num_samples, 1000)
  PS_FIELD(int32_t, num_pose_samples, 0)
  PS_FIELD(bool, use_pairwise, true)
  PS_FIELD(bool, use_torso_pos_prior, false)
  PS_FIELD(bool, pred_unary_rot, false)
  PS_FIELD(bool, pred_unary_pos, false)
  PS_FIELD(float, pred_unary_rot_weight, 1)
  PS_FIELD(float, pred_unary_pos_weight, 1)
  PS_FIELD(bool, use_dpm_unary, false)
  PS_FIELD(bool, use_dpm_torso, false)
  PS_FIELD(bool, use_dpm_head, false)
  PS_FIELD(float, dpm_unary_weight, 1)
  PS_FIELD(float, dpm_torso_weight, 1)
  PS_FIELD(float, dpm_head_weight, 1)
  PS_FIELD(bool, do_dpm_rot, false)
  PS_FIELD(bool, save_part_marginals, false)
  PS_FIELD(bool, save_root_marginal, false)
  PS_FIELD(bool, save_part_marginals_local_max, false)
  PS_FIELD(bool, save_part_detections_local_max, false)
  PS_FIELD(bool, interpolate, false)
  PS_FIELD(std::string, log_dir, "")
  PS_FIELD(std::string, log_subdir, "")
  PS_FIELD(std::string, test_dpm_unary_dir, "")
  PS_FIELD(std::string, test_dpm_torso_dir, "")
  PS_FIELD(std::string, test_dpm_head_dir, "")
  PS_FIELD(std::string, part_conf_type, "")
  PS_FIELD(std::string, pred_data_test_dir, "")
  PS_FIELD(std::string, mix_dir, "")
  PS_FIELD(std::string, pred_data_train_dir, "")
  PS_FIELD(std::string, scoregrid_dir, "")
  PS_FIELD(std::string, spatial_dir, "")
  PS_FIELD(bool, flip_orientation, false)
  PS_FIELD(bool, force_recompute_scores, true)
  PS_FIELD(float, object_height_width_ratio, 2)
  PS_FIELD(std::string, class_dir, "")
  PS_FIELD(float, torso_pos_prior_weight, 1)
  PS_FIELD(std::string, pred_data_dir, "")
  PS_FIELD(int32_t, rootidx_det, -1)
  PS_FIELD(bool, use_gt_torso, false)
  PS_FIELD(std::string, torso_det_test_dir, "")
  PS_FIELD(std::string, torso_det_train_dir, "")
  PS_FIELD(std::string, poselet_resp_val_dir, "")
  PS_FIELD(std::string, poselet_resp_test_dir, "")
  PS_FIELD(std::string, poselet_resp_train_dir, "")
  PS_FIELD(std::string, poselet_strip, "")
  PS_FIELD(std::string, part_marginals_dir, "")
  PS_FIELD(int32_t, num_pred_part_types, 1)
  PS_FIELD(std::string, dai_samples_dir, "")
  PS_FIELD(std::string, part_conf_eval, "")
  PS_FIELD(std::string, roi_annolist, "")
  PS_FIELD(std::string, scoregrid_train_dir, "")
  PS_FIELD(bool, dai_bbox_prior, false)
#undef PS_FIELD
  std::string validation_dataset(int) const { return std::string(); }
  int validation_dataset_size() const { return 0; }
  std::string train_dataset(int) const { return std::string(); }
  int train_dataset_size() const { return 0; }
  std::string test_dataset(int) const { return std::string(); }
  int test_dataset_size() const { return 0; }
};
