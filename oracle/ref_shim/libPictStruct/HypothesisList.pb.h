// Stand-in for the protoc-generated HypothesisList.pb.h -- TEST INFRASTRUCTURE (HypothesisList.proto: repeated
// ObjectHypothesis hyp { x, y, scale, score, flip }).
#pragma once
#include <vector>
class HypothesisList {
 public:
  class ObjectHypothesis {
   public:
    float x_ = 0, y_ = 0, scale_ = 0, score_ = 0;
    bool flip_ = false;
    float x() const { return x_; }
    float y() const { return y_; }
    float scale() const { return scale_; }
    float score() const { return score_; }
    bool flip() const { return flip_; }
    void set_x(float v) { x_ = v; }
    void set_y(float v) { y_ = v; }
    void set_scale(float v) { scale_ = v; }
    void set_score(float v) { score_ = v; }
    void set_flip(bool v) { flip_ = v; }
  };
  std::vector<ObjectHypothesis> hyp_;
  int hyp_size() const { return (int)hyp_.size(); }
  const ObjectHypothesis &hyp(int i) const { return hyp_.at((size_t)i); }
  ObjectHypothesis *mutable_hyp(int i) { return &hyp_.at((size_t)i); }
  ObjectHypothesis *add_hyp() { hyp_.push_back(ObjectHypothesis()); return &hyp_.back(); }
  void Clear() { hyp_.clear(); }
};
