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
};
class PartConfig {
 public:
  std::vector<PartDef> parts_;
  int part_size() const { return (int)parts_.size(); }
  const PartDef &part(int i) const { return parts_.at((size_t)i); }
  int joint_size() const { return 0; }
};
