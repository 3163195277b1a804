// Stand-in for libKMA2 -- TEST INFRASTRUCTURE.
#pragma once
namespace kma {
class ImageContent {
 public:
  int x() const { return 0; }
  int y() const { return 0; }
};
inline ImageContent *load_convert_gray_image(const char *) { return 0; }
}  // namespace kma
