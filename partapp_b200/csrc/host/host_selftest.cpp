// psinfer_host_selftest -- CPU-only probes of the host formats, driven by tests/test_host_formats.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

#include "mat5.hpp"
#include "objectdetect_b200.hpp"

static void dump(const mat5::Var &v, int depth) {
  printf("%*sname %s class %d dims", depth, "", v.name.c_str(), v.cls);
  for (size_t d : v.dims) printf(" %zu", d);
  printf("\n");
  if (v.cls == mat5::mxCELL) {
    for (const mat5::Var &c : v.cells) dump(c, depth + 2);
  } else {
    for (size_t i = 0; i < v.numel(); ++i) printf("%*s%.9g\n", depth, "", v.at(i));
  }
}

int main(int argc, char **argv) {
  try {
    if (argc >= 4 && !strcmp(argv[1], "mat-dump")) {
      std::vector<mat5::Var> vars = mat5::load(argv[2]);
      dump(mat5::find(vars, argv[3], argv[2]), 0);
      return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "mat-write")) {
      mat5::Writer w(argv[2], argc < 4 || strcmp(argv[3], "raw") != 0);
      float a[24];
      for (int i = 0; i < 24; ++i) a[i] = (float)i * 0.5f;
      w.put("a", a, {2, 3, 4});
      double b[6] = {1.5, -2.5, 3.25, 4, 5, 6e10};
      w.put("b", b, {3, 2});
      double s = 7.5;
      w.put("s", &s, {1, 1});
      float v[3] = {1, 2, 3};
      w.put("v", v, {3});
      return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "pbuf-write")) {
      object_detect::HypothesisList l;
      for (int i = 0; i < 3; ++i) {
        object_detect::ObjectHypothesis h;
        h.x = 10.f + i; h.y = 20.f + i; h.scale = 1.0f + 0.1f * i; h.score = -3.5f * i; h.flip = i == 1;
        l.hyp.push_back(h);
      }
      std::string b = l.SerializeAsString();
      object_detect::HypothesisList back = object_detect::HypothesisList::Parse(b);
      if (back.hyp.size() != 3 || back.hyp[2].score != -7.0f || !back.hyp[1].flip) return 3;
      std::ofstream f(argv[2], std::ios::binary);
      f.write(b.data(), (std::streamsize)b.size());
      return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "expopt-dump")) {
      object_detect::PartApp app;
      app.init(argv[2]);
      const object_detect::ExpParam &e = app.m_exp_param;
      printf("log_dir %s\nlog_subdir %s\nclass_dir %s\nscoregrid_dir %s\nspatial_dir %s\n", e.log_dir.c_str(),
             e.log_subdir.c_str(), e.class_dir.c_str(), e.scoregrid_dir.c_str(), e.spatial_dir.c_str());
      printf("rot %u %g %g scale %u %g %g strip %g K %g flip %d\n", e.num_rotation_steps, e.min_part_rotation,
             e.max_part_rotation, e.num_scale_steps, e.min_object_scale, e.max_object_scale, e.strip_border_detections,
             e.roi_save_num_samples, (int)e.flip_orientation);
      printf("parts %zu root %d joints %zu\n", app.m_part_conf.part.size(), app.m_rootpart_idx, app.m_part_conf.joint.size());
      for (const auto &j : app.m_part_conf.joint) printf("joint %d %d %s %u\n", j.child_idx, j.parent_idx, j.type.c_str(), j.num_joint_types);
      printf("bbox_offset %g %g\n", app.m_window_param.bbox_offset_x, app.m_window_param.bbox_offset_y);
      for (const std::string &im : app.m_test_annolist) {
        int w, h;
        object_detect::image_size(im, w, h);
        printf("image %s %d %d\n", im.c_str(), w, h);
      }
      if (e.use_gt_torso)
        for (size_t i = 0; i < app.m_test_annolist.size(); ++i) {
          double rp[2];
          object_detect::getRootPosDet(app, (int)i, app.m_rootpart_idx, rp, true);
          printf("gt_torso %zu %g %g\n", i, rp[0], rp[1]);
        }
      if (argc >= 4) {
        std::vector<object_detect::Joint> joints;
        object_detect::loadJoints(app, joints, atoi(argv[3]) != 0, -1);
        for (const auto &j : joints)
          printf("J %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", j.type, j.child_idx, j.parent_idx,
                 j.offset_c[0], j.offset_c[1], j.offset_p[0], j.offset_p[1], j.C[0][0], j.C[0][1], j.C[1][0], j.C[1][1],
                 j.rot_mean, j.rot_sigma);
      }
      return 0;
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  fprintf(stderr, "usage: mat-dump <file> <var> | mat-write <file> [raw] | pbuf-write <file> | expopt-dump <expopt> [flip]\n");
  return 2;
}
