// psinfer_partapp -- the `--find_obj` verb of the reference CLI (reference src/apps/partapp/main.cpp:194-264, :774-783)
// on the B200 path.  Usage mirrors run_partapp.sh:
//
//   psinfer_partapp --expopt <exp.txt> --find_obj [--first <idx>] [--numimgs <n>]
//
// Only --expopt, --find_obj, --first and --numimgs exist here; every other verb of partapp (training, detection,
// evaluation, visualisation) stays on the reference binary.  Unary score grids must already be on disk in
// <scoregrid_dir> (the reference writes them with --part_detect; force_recompute_scores is ignored).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "objectdetect_b200.hpp"

int main(int argc, char **argv) {
  std::string expopt;
  bool find_obj = false;
  int first = -1, numimgs = -1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto value = [&](const char *name) -> std::string {
      std::string pre = std::string(name) + "=";
      if (a.compare(0, pre.size(), pre) == 0) return a.substr(pre.size());
      if (i + 1 >= argc) {
        fprintf(stderr, "%s needs a value\n", name);
        exit(2);
      }
      return argv[++i];
    };
    if (a == "--expopt" || a.compare(0, 9, "--expopt=") == 0) expopt = value("--expopt");
    else if (a == "--find_obj") find_obj = true;
    else if (a == "--first" || a.compare(0, 8, "--first=") == 0) first = atoi(value("--first").c_str());
    else if (a == "--numimgs" || a.compare(0, 10, "--numimgs=") == 0) numimgs = atoi(value("--numimgs").c_str());
    else if (a == "--help" || a == "-h") {
      printf("usage: %s --expopt <file> --find_obj [--first <idx>] [--numimgs <n>]\n", argv[0]);
      return 0;
    } else {
      fprintf(stderr, "unsupported option %s (only --expopt, --find_obj, --first, --numimgs run on this path)\n", a.c_str());
      return 2;
    }
  }
  if (expopt.empty() || !find_obj) {
    fprintf(stderr, "usage: %s --expopt <file> --find_obj [--first <idx>] [--numimgs <n>]\n", argv[0]);
    return 2;
  }
  try {
    object_detect::PartApp app;
    app.init(expopt);
    // init_firstidx_lastidx, main.cpp:155-192
    const int n = (int)app.m_test_annolist.size();
    int firstidx = first >= 0 ? first : 0;
    int lastidx = numimgs >= 0 ? std::min(n - 1, firstidx + numimgs - 1) : n - 1;
    object_detect::findObjectDataset(app, firstidx, lastidx);
    printf("find_obj: processed images %d..%d of %d\n", firstidx, lastidx, n);
  } catch (const std::exception &e) {
    fprintf(stderr, "psinfer_partapp: %s\n", e.what());
    return 1;
  }
  return 0;
}
