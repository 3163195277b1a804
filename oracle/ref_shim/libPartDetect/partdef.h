// Stand-in for libPartDetect/partdef.h -- TEST INFRASTRUCTURE: PartBBox with the reference's fields.
#pragma once
#include <libBoostMath/boost_math.h>
struct PartBBox {
  PartBBox() : part_pos(2), part_x_axis(2), part_y_axis(2), use_endpoints(false) {}
  boost_math::double_vector part_pos, part_x_axis, part_y_axis;
  double max_proj_x, min_proj_x, max_proj_y, min_proj_y;
  float x1, x2, y1, y2;
  bool use_endpoints;
};

#include <libAnnotation/annotation.h>
#include <libPartDetect/PartConfig.pb.h>
bool annorect_has_part(const AnnoRect &annorect, const PartDef &partdef);
bool get_part_bbox(const AnnoRect &annorect, const PartDef &partdef, PartBBox &part_bbox, double scale = 1.0);
