// Stand-in for libAnnotation -- TEST INFRASTRUCTURE.
#pragma once
#include <string>
#include <vector>
class AnnoRect {
 public:
  AnnoRect() {}
  AnnoRect(double, double, double, double) {}
};
class Annotation {
 public:
  std::string name_;
  const std::string &imageName() const { return name_; }
};
