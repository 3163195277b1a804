// Stand-in for libAnnotation -- TEST INFRASTRUCTURE.
#pragma once
#include <libAnnotation/annotation.h>
class AnnotationList {
 public:
  std::vector<Annotation> v_;
  AnnotationList() {}
  explicit AnnotationList(const std::string &) {}
  size_t size() const { return v_.size(); }
  const Annotation &operator[](size_t i) const { return v_.at(i); }
  void addAnnotation(const Annotation &a) { v_.push_back(a); }
  void save(const std::string &, bool = false) const {}
  void saveIDL(const std::string &) const {}
  void load(const std::string &) {}
  Annotation &operator[](size_t i) { return v_.at(i); }
};
