// Stand-in for libAnnotation -- TEST INFRASTRUCTURE.
#pragma once
#include <string>
#include <vector>
struct AnnoPoint {
  int id = -1, x = 0, y = 0;
  bool is_visible = true;
};
class AnnoRect {
 public:
  std::vector<AnnoPoint> m_vAnnoPoints;
  const AnnoPoint *get_annopoint_by_id(int id) const {
    for (const AnnoPoint &p : m_vAnnoPoints) if (p.id == id) return &p;
    return 0;
  }
  double scale() const { return -1; }
  int m_x1 = 0, m_y1 = 0, m_x2 = 0, m_y2 = 0;
  AnnoRect() {}
  AnnoRect(double, double, double, double) {}
  AnnoRect(float, float, float, float, double) {}
  AnnoRect(int, int, int, int, float, int, float) {}
  int silhouetteID() const { return -1; }
  int top() const { return m_y1; }
  int bottom() const { return m_y2; }
  int left() const { return m_x1; }
  int right() const { return m_x2; }
};
class Annotation {
 public:
  Annotation() {}
  explicit Annotation(const std::string &n) : name_(n) {}
  std::string name_;
  std::vector<AnnoRect> rects_;
  std::vector<AnnoRect> &m_vRects = rects_;
  Annotation(const Annotation &o) : name_(o.name_), rects_(o.rects_) {}
  Annotation &operator=(const Annotation &o) { name_ = o.name_; rects_ = o.rects_; return *this; }
  const std::string &imageName() const { return name_; }
  void addAnnoRect(const AnnoRect &r) const { const_cast<Annotation *>(this)->rects_.push_back(r); }
  size_t size() const { return rects_.size(); }
  const AnnoRect &operator[](int i) const { return rects_.at((size_t)i); }
};
