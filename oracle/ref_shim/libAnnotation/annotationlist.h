// Stand-in for libAnnotation -- TEST INFRASTRUCTURE.
#pragma once
#include <libAnnotation/annotation.h>
class AnnotationList {
 public:
  std::vector<Annotation> v_;
  size_t size() const { return v_.size(); }
  const Annotation &operator[](size_t i) const { return v_.at(i); }
};
