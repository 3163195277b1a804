// psinfer_partapp -- the `--find_obj` verb of the reference CLI (reference src/apps/partapp/main.cpp:194-264, :774-783)
// on the B200 path.  Usage mirrors run_partapp.sh:
//
//   psinfer_partapp --expopt <exp.txt> --find_obj [--first <idx>] [--numimgs <n>]
//                   [--distribute --ncpu <n> --batch_num <b>]      process-level shards, main.cpp:155-192
//                   [--gpus <n>] [--contexts <k>]                  GPUs of this process, worker threads per GPU
//
// Only these options exist here; every other verb of partapp (training, detection, evaluation, visualisation) stays
// on the reference binary.  Unary score grids must already be on disk in
// <scoregrid_dir> (the reference writes them with --part_detect; force_recompute_scores is ignored).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "objectdetect_b200.hpp"

int main(int argc, char **argv) {
  std::string expopt;
  bool find_obj = false;
  int first = -1, numimgs = -1, ncpu = -1, batch_num = -1, gpus = 0, contexts = 0;
  bool distribute = false;
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
    else if (a == "--distribute") distribute = true;
    else if (a == "--ncpu" || a.compare(0, 7, "--ncpu=") == 0) ncpu = atoi(value("--ncpu").c_str());
    else if (a == "--batch_num" || a.compare(0, 12, "--batch_num=") == 0) batch_num = atoi(value("--batch_num").c_str());
    else if (a == "--gpus" || a.compare(0, 7, "--gpus=") == 0) gpus = atoi(value("--gpus").c_str());
    else if (a == "--contexts" || a.compare(0, 11, "--contexts=") == 0) contexts = atoi(value("--contexts").c_str());
    else if (a == "--help" || a == "-h") {
      printf("usage: %s --expopt <file> --find_obj [--first <idx>] [--numimgs <n>] [--distribute --ncpu <n> --batch_num <b>] "
             "[--gpus <n>] [--contexts <k>]\n", argv[0]);
      return 0;
    } else {
      fprintf(stderr, "unsupported option %s (only --expopt, --find_obj, --first, --numimgs, --distribute, --ncpu, --batch_num, "
                      "--gpus, --contexts run on this path)\n", a.c_str());
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
    int lastidx = numimgs >= 0 ? firstidx + numimgs - 1 : n - 1;
    if (distribute) {
      if (ncpu < 1 || batch_num < 0) {
        fprintf(stderr, "--distribute needs --ncpu and --batch_num (main.cpp:175-176)\n");
        return 2;
      }
      const int num_per_cpu = (int)std::ceil((lastidx - firstidx + 1) / (float)ncpu);
      firstidx = firstidx + batch_num * num_per_cpu;
      lastidx = firstidx + num_per_cpu - 1;
    }
    // check_bounds_and_update: firstidx > lastidx is allowed, no image is processed then
    firstidx = std::max(0, std::min(firstidx, n));
    lastidx = std::max(-1, std::min(lastidx, n - 1));
    object_detect::set_parallelism(gpus, contexts);
    const auto t0 = std::chrono::steady_clock::now();
    if (firstidx <= lastidx) object_detect::findObjectDataset(app, firstidx, lastidx);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const int done = std::max(0, lastidx - firstidx + 1);
    printf("find_obj: processed images %d..%d of %d in %.3f s (%.2f images/s)\n", firstidx, lastidx, n, dt,
           dt > 0 ? done / dt : 0.0);
  } catch (const std::exception &e) {
    fprintf(stderr, "psinfer_partapp: %s\n", e.what());
    return 1;
  }
  return 0;
}
