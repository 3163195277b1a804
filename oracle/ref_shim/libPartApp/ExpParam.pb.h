// Stand-in for the protoc-generated libPartApp/ExpParam.pb.h -- TEST INFRASTRUCTURE (protoc and the protobuf C++
// runtime are not installed here).  Only the accessors the reference's libPartApp/partapp_aux.hpp reads, with the
// field types of ExpParam.proto (float ranges, uint32 counts), so that header compiles unmodified into oracle/_ref/.
#pragma once
#include <cassert>
#include <cstdint>
class ExpParam {
 public:
  float min_object_scale_ = 1, max_object_scale_ = 1, min_part_rotation_ = -180, max_part_rotation_ = 180;
  uint32_t num_scale_steps_ = 1, num_rotation_steps_ = 48;
  float min_object_scale() const { return min_object_scale_; }
  float max_object_scale() const { return max_object_scale_; }
  uint32_t num_scale_steps() const { return num_scale_steps_; }
  float min_part_rotation() const { return min_part_rotation_; }
  float max_part_rotation() const { return max_part_rotation_; }
  uint32_t num_rotation_steps() const { return num_rotation_steps_; }
};
