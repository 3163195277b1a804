// Stand-in for the protoc-generated PartWindowParam.pb.h -- TEST INFRASTRUCTURE.
#pragma once
#include <vector>
class PartWindowParam {
 public:
  class PartParam {
   public:
    int window_size_x_ = 0, window_size_y_ = 0, pos_offset_x_ = 0, pos_offset_y_ = 0;
    int window_size_x() const { return window_size_x_; }
    int window_size_y() const { return window_size_y_; }
    int pos_offset_x() const { return pos_offset_x_; }
    int pos_offset_y() const { return pos_offset_y_; }
  };
  std::vector<PartParam> parts_;
  int part_size() const { return (int)parts_.size(); }
  double bbox_offset_x_ = 0, bbox_offset_y_ = 0;
  double bbox_offset_x() const { return bbox_offset_x_; }
  double bbox_offset_y() const { return bbox_offset_y_; }
  double train_object_height() const { return 200; }
  const PartParam &part(int i) const {
    static PartParam p;
    return (size_t)i < parts_.size() ? parts_[(size_t)i] : p;
  }
};
