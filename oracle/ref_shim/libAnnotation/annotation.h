// Stand-in for libAnnotation -- TEST INFRASTRUCTURE.
#pragma once
#include <string>
#include <vector>
class AnnoRect {
 public:
  AnnoRect() {}
  AnnoRect(double, double, double, double) {}
  AnnoRect(int, int, int, int, float, int, float) {}
  int silhouetteID() const { return -1; }
};
class Annotation {
 public:
  Annotation() {}
  explicit Annotation(const std::string &n) : name_(n) {}
  std::string name_;
  std::vector<AnnoRect> rects_;
  const std::string &imageName() const { return name_; }
  void addAnnoRect(const AnnoRect &r) { rects_.push_back(r); }
  size_t size() const { return rects_.size(); }
  const AnnoRect &operator[](int i) const { return rects_.at((size_t)i); }
};
