// Stand-in for the protoc-generated PartConfig.pb.h -- TEST INFRASTRUCTURE: the accessors objectdetect_findrot.cpp reads.
#pragma once
#include <string>
#include <vector>
class PartDef {
 public:
  bool is_detect_ = true, is_upright_ = false, is_root_ = false;
  int part_id_ = 0;
  bool is_detect() const { return is_detect_; }
  bool is_upright() const { return is_upright_; }
  bool is_root() const { return is_root_; }
  int part_id() const { return part_id_; }
  std::vector<int> part_pos_;
  int part_pos(int i) const { return part_pos_.at((size_t)i); }
  int part_pos_size() const { return (int)part_pos_.size(); }
  // fields the evaluator reads (libPartDetect/partdef.cpp, libPartEval/parteval.cpp)
  std::vector<int> part_x_axis_from_, part_x_axis_to_;
  float part_x_axis_offset_ = 0, ext_x_pos_ = 0, ext_x_neg_ = 0, ext_y_pos_ = 0, ext_y_neg_ = 0;
  int part_x_axis_from(int i) const { return part_x_axis_from_.at((size_t)i); }
  int part_x_axis_from_size() const { return (int)part_x_axis_from_.size(); }
  int part_x_axis_to(int i) const { return part_x_axis_to_.at((size_t)i); }
  int part_x_axis_to_size() const { return (int)part_x_axis_to_.size(); }
  float part_x_axis_offset() const { return part_x_axis_offset_; }
  float ext_x_pos() const { return ext_x_pos_; }
  float ext_x_neg() const { return ext_x_neg_; }
  float ext_y_pos() const { return ext_y_pos_; }
  float ext_y_neg() const { return ext_y_neg_; }
  int num_pred_part_types() const { return 1; }
  int max_num_part_types() const { return 1; }
  bool has_mult_types() const { return false; }
};
class JointDef {
 public:
  int child_idx_ = 0, parent_idx_ = 0, num_joint_types_ = 1;
  std::string type_ = "RotGaussian";
  int child_idx() const { return child_idx_; }
  int parent_idx() const { return parent_idx_; }
  const std::string &type() const { return type_; }
  bool has_num_joint_types() const { return num_joint_types_ > 1; }
  int num_joint_types() const { return num_joint_types_; }
};
class PartConfig {
 public:
  std::vector<PartDef> parts_;
  std::vector<JointDef> joints_;
  const JointDef &joint(int i) const { return joints_.at((size_t)i); }
  int part_size() const { return (int)parts_.size(); }
  const PartDef &part(int i) const { return parts_.at((size_t)i); }
  int joint_size() const { return (int)joints_.size(); }
};
