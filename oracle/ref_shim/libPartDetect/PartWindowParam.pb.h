// Stand-in for the protoc-generated PartWindowParam.pb.h -- TEST INFRASTRUCTURE.
#pragma once
class PartWindowParam {
 public:
  class PartParam {
   public:
    int window_size_x() const { return 0; }
    int window_size_y() const { return 0; }
    int pos_offset_x() const { return 0; }
    int pos_offset_y() const { return 0; }
  };
  double bbox_offset_x_ = 0, bbox_offset_y_ = 0;
  double bbox_offset_x() const { return bbox_offset_x_; }
  double bbox_offset_y() const { return bbox_offset_y_; }
  double train_object_height() const { return 200; }
  const PartParam &part(int) const { static PartParam p; return p; }
};
