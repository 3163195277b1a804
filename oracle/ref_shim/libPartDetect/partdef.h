// Stand-in for libPartDetect/partdef.h -- TEST INFRASTRUCTURE: PartBBox with the reference's fields.
#pragma once
#include <libBoostMath/boost_math.h>
struct PartBBox {
  PartBBox() : part_pos(2), part_x_axis(2), part_y_axis(2), use_endpoints(false) {}
  PartBBox(int ox, int oy, double xaxis_x, double xaxis_y, double _min_x, double _max_x, double _min_y, double _max_y);
  boost_math::double_vector part_pos, part_x_axis, part_y_axis;
  double max_proj_x, min_proj_x, max_proj_y, min_proj_y;
  float x1, x2, y1, y2;
  bool use_endpoints;
};

#include <libAnnotation/annotation.h>
#include <libPartDetect/PartConfig.pb.h>
#include <QPainter>
#include <libPartDetect/PartWindowParam.pb.h>
boost_math::double_vector get_part_position(const AnnoRect &annorect, const PartDef &partdef);
bool get_part_x_axis(const AnnoRect &annorect, const PartDef &partdef, boost_math::double_vector &part_x_axis);
bool annorect_has_part(const AnnoRect &annorect, const PartDef &partdef);
bool get_part_bbox(const AnnoRect &annorect, const PartDef &partdef, PartBBox &part_bbox, double scale = 1.0);
void draw_bbox(QPainter &painter, const PartBBox &part_bbox, int coloridx = 0, int pen_width = 1);
QImage visualize_parts(const PartConfig &conf, const PartWindowParam &window_param, const Annotation &annotation);
void get_part_polygon(PartBBox &part_bbox, QPolygonF &polygon);
void update_bbox_min_max_proj(PartBBox &part_bbox, std::vector<boost_math::double_vector> &corners);
void get_bbox_corners(const AnnoRect &annorect, const PartDef &partdef, std::vector<boost_math::double_vector> &corners);
