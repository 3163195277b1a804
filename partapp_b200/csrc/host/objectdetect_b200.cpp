// objectdetect_b200.cpp -- see objectdetect_b200.hpp.  Reference line numbers are relative to
// /root/reference/src/libs/.
#include "objectdetect_b200.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>

#include "mat5.hpp"
#include "prototext.hpp"

namespace object_detect {

namespace {

void fail(const std::string &msg) { throw std::runtime_error(msg); }

bool file_exists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool dir_exists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
void make_dirs(const std::string &p) {  // filesys::create_dir, recursive
  std::string cur;
  for (size_t i = 0; i <= p.size(); ++i) {
    if (i == p.size() || p[i] == '/') {
      if (!cur.empty() && !dir_exists(cur) && mkdir(cur.c_str(), 0777) != 0 && !dir_exists(cur))
        fail("cannot create directory " + cur);
    }
    if (i < p.size()) cur += p[i];
  }
}
std::string dirname_of(const std::string &p) {
  size_t k = p.find_last_of('/');
  return k == std::string::npos ? "." : (k == 0 ? "/" : p.substr(0, k));
}
std::string basename_noext(const std::string &p) {
  size_t k = p.find_last_of('/');
  std::string b = k == std::string::npos ? p : p.substr(k + 1);
  size_t d = b.find_last_of('.');
  return d == std::string::npos ? b : b.substr(0, d);
}
// complete_relative_path, partapp.cpp:112-139
std::string complete_relative_path(std::string in, const std::string &reference_file) {
  while (!in.empty() && isspace((unsigned char)in.front())) in.erase(in.begin());
  while (!in.empty() && isspace((unsigned char)in.back())) in.pop_back();
  if (in.empty() || in[0] == '/') return in;
  if (in.compare(0, 2, "./") == 0) in = in.substr(2);
  return dirname_of(reference_file) + "/" + in;
}
std::string pad_zeros(int v, int n) {
  char buf[32];
  snprintf(buf, sizeof buf, "%0*d", n, v);
  return buf;
}

// ---- ps_ctx cache: one context per grid shape ----------------------------------------------------------------
struct CtxKey {
  int R, S, H, W, P, root, keep;
  float rmin, rmax, smin, smax, strip;
  int K;
  unsigned char flags[3 * PS_MAX_PARTS];
  bool operator==(const CtxKey &o) const { return memcmp(this, &o, sizeof(CtxKey)) == 0; }
};
struct CtxHolder {
  ps_ctx *ctx = nullptr;
  CtxKey key;
  ~CtxHolder() {
    if (ctx) ps_destroy(ctx);
  }
};
CtxHolder &holder() {
  static thread_local CtxHolder h;
  return h;
}

// PSINFER_HOST_TIMING=1: seconds spent per phase, summed over the worker threads, printed by findObjectDataset.
std::atomic<long long> g_ns_load(0), g_ns_ingest(0), g_ns_infer(0), g_ns_out(0), g_ns_joints(0);
struct PhaseTimer {
  std::atomic<long long> &acc;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(std::atomic<long long> &a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~PhaseTimer() { acc += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

// Device of the calling worker thread (findObjectDataset sets it); -1: PSINFER_DEVICE or device 0.
thread_local int t_device = -1;
int g_gpus = 0, g_ctx_per_gpu = 0;  // 0: take PSINFER_GPUS / PSINFER_CTX_PER_GPU, else 1 GPU x 4 contexts

void check(ps_ctx *ctx, int st, const char *what) {
  if (st != PS_OK) fail(std::string(what) + ": " + ps_last_error(ctx));
}

ps_config make_config(const PartApp &app, int H, int W, int root, bool keep_all) {
  const ExpParam &ep = app.m_exp_param;
  ps_config cfg;
  memset(&cfg, 0, sizeof cfg);
  const char *dev = getenv("PSINFER_DEVICE");
  cfg.device = t_device >= 0 ? t_device : (dev ? atoi(dev) : 0);
  cfg.num_parts = (int)app.m_part_conf.part.size();
  cfg.num_rotation_steps = (int)ep.num_rotation_steps;
  cfg.min_part_rotation = ep.min_part_rotation;
  cfg.max_part_rotation = ep.max_part_rotation;
  cfg.num_scale_steps = (int)ep.num_scale_steps;
  cfg.min_object_scale = ep.min_object_scale;
  cfg.max_object_scale = ep.max_object_scale;
  cfg.height = H;
  cfg.width = W;
  cfg.root_idx = root;
  if (cfg.num_parts > PS_MAX_PARTS) fail("too many parts");
  for (int p = 0; p < cfg.num_parts; ++p) {
    cfg.is_detect[p] = app.m_part_conf.part[p].is_detect;
    cfg.is_upright[p] = app.m_part_conf.part[p].is_upright;
    cfg.is_root[p] = app.m_part_conf.part[p].is_root;
  }
  cfg.strip_border_detections = ep.strip_border_detections;
  cfg.roi_save_num_samples = (int)ep.roi_save_num_samples;
  cfg.keep_all_scales = keep_all ? 1 : 0;
  cfg.interpolate = ep.interpolate ? 1 : 0;
  // arithmetic mode of the library (include/psinfer.h): parity by default; PSINFER_FAST_MATH=1 selects fused taps
  const char *fm = getenv("PSINFER_FAST_MATH");
  cfg.fast_math = (fm && fm[0] && fm[0] != '0') ? 1 : 0;
  return cfg;
}

ps_ctx *get_ctx(const PartApp &app, int H, int W, int root, bool keep_all) {
  ps_config cfg = make_config(app, H, W, root, keep_all);
  CtxKey k;
  memset(&k, 0, sizeof k);
  k.R = cfg.num_rotation_steps; k.S = cfg.num_scale_steps; k.H = H; k.W = W; k.P = cfg.num_parts; k.root = root;
  k.keep = cfg.keep_all_scales; k.rmin = cfg.min_part_rotation; k.rmax = cfg.max_part_rotation;
  k.smin = cfg.min_object_scale; k.smax = cfg.max_object_scale; k.strip = cfg.strip_border_detections;
  k.K = cfg.roi_save_num_samples * 4 + cfg.interpolate * 2 + cfg.fast_math + cfg.device * (1 << 24);
  memcpy(k.flags, cfg.is_detect, PS_MAX_PARTS);
  memcpy(k.flags + PS_MAX_PARTS, cfg.is_upright, PS_MAX_PARTS);
  memcpy(k.flags + 2 * PS_MAX_PARTS, cfg.is_root, PS_MAX_PARTS);
  CtxHolder &h = holder();
  if (h.ctx && h.key == k) return h.ctx;
  if (h.ctx) ps_destroy(h.ctx);
  h.ctx = nullptr;
  if (ps_create(&cfg, &h.ctx) != PS_OK) fail(std::string("ps_create: ") + ps_last_error(nullptr));
  h.key = k;
  return h.ctx;
}

std::vector<ps_joint> to_ps_joints(const std::vector<Joint> &joints) {
  std::vector<ps_joint> pj(joints.size());
  for (size_t j = 0; j < joints.size(); ++j) {
    pj[j].type = joints[j].type;
    pj[j].child_idx = joints[j].child_idx;
    pj[j].parent_idx = joints[j].parent_idx;
    for (int k = 0; k < 2; ++k) {
      pj[j].offset_c[k] = joints[j].offset_c[k];
      pj[j].offset_p[k] = joints[j].offset_p[k];
    }
    pj[j].C[0] = joints[j].C[0][0]; pj[j].C[1] = joints[j].C[0][1];
    pj[j].C[2] = joints[j].C[1][0]; pj[j].C[3] = joints[j].C[1][1];
    pj[j].rot_mean = joints[j].rot_mean;
    pj[j].rot_sigma = joints[j].rot_sigma;
  }
  return pj;
}

PartHyp hyp_from_row(const float *r) {
  PartHyp h;
  h.fromVect(r);
  return h;
}

void collect_results(const PartApp &app, ps_ctx *ctx, FloatGrid3 &root_part_posterior,
                     std::vector<std::vector<PartHyp> > &best_part_hyp, int H, int W) {
  const int P = (int)app.m_part_conf.part.size(), S = (int)app.m_exp_param.num_scale_steps;
  root_part_posterior = FloatGrid3(S, H, W);
  check(ctx, ps_get_root_posterior(ctx, root_part_posterior.data(), PS_MEM_HOST), "ps_get_root_posterior");
  best_part_hyp.assign(P, std::vector<PartHyp>());
  const int cap = (int)app.m_exp_param.roi_save_num_samples + 1;
  std::vector<float> rows((size_t)cap * PS_HYP_VEC);
  for (int p = 0; p < P; ++p) {
    int n = 0;
    check(ctx, ps_get_part_hyps(ctx, p, rows.data(), cap, &n), "ps_get_part_hyps");
    for (int i = 0; i < n; ++i) best_part_hyp[p].push_back(hyp_from_row(&rows[(size_t)i * PS_HYP_VEC]));
  }
}

void save_marginals(const PartApp &app, ps_ctx *ctx, int imgidx, bool flip, int H, int W) {
  // computePartMarginals :239-253: log_part_posterior_final_imgidx<i>_scaleidx<s>_o<f>_pidx<p>.mat, var log_prob_grid
  const ExpParam &ep = app.m_exp_param;
  const std::string dir = ep.log_dir + "/" + ep.log_subdir + "/part_marginals";
  make_dirs(dir);
  const int P = (int)app.m_part_conf.part.size(), S = (int)ep.num_scale_steps, R = (int)ep.num_rotation_steps;
  FloatGrid3 g(R, H, W);
  for (int s = 0; s < S; ++s)
    for (int p = 0; p < P; ++p) {
      check(ctx, ps_get_marginal(ctx, p, s, g.data(), PS_MEM_HOST), "ps_get_marginal");
      char name[256];
      snprintf(name, sizeof name, "/log_part_posterior_final_imgidx%d_scaleidx%d_o%d_pidx%d.mat", imgidx, s, (int)flip, p);
      mat5::Writer w(dir + name);
      w.put("log_prob_grid", g.data(), {(size_t)R, (size_t)H, (size_t)W});
    }
}

}  // namespace

// ---- small public helpers -------------------------------------------------------------------------------------------

double rot_from_index(const ExpParam &ep, int idx) {
  ps_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.num_rotation_steps = (int)ep.num_rotation_steps;
  cfg.min_part_rotation = ep.min_part_rotation;
  cfg.max_part_rotation = ep.max_part_rotation;
  return ps_rot_from_index(&cfg, idx);
}
double scale_from_index(const ExpParam &ep, int idx) {
  ps_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.num_scale_steps = (int)ep.num_scale_steps;
  cfg.min_object_scale = ep.min_object_scale;
  cfg.max_object_scale = ep.max_object_scale;
  return ps_scale_from_index(&cfg, idx);
}

std::string getObjectHypFilename(int imgidx, bool flip) {
  return "/object_hyp_imgidx" + std::to_string(imgidx) + "_o" + std::to_string((int)flip) + "_spmnone.pbuf";
}

void image_size(const std::string &file, int &width, int &height) {
  std::ifstream f(file.c_str(), std::ios::binary);
  if (!f) fail("cannot open image " + file);
  unsigned char h[32];
  f.read((char *)h, 32);
  if (f.gcount() >= 24 && h[0] == 0x89 && h[1] == 'P' && h[2] == 'N' && h[3] == 'G') {  // IHDR at offset 16
    width = (h[16] << 24) | (h[17] << 16) | (h[18] << 8) | h[19];
    height = (h[20] << 24) | (h[21] << 16) | (h[22] << 8) | h[23];
    return;
  }
  if (h[0] == 0xff && h[1] == 0xd8) {  // JPEG: walk the segments to a start-of-frame marker
    f.clear();
    f.seekg(2);
    for (;;) {
      unsigned char m[4];
      f.read((char *)m, 4);
      if (!f || m[0] != 0xff) break;
      int len = (m[2] << 8) | m[3];
      if (m[1] >= 0xc0 && m[1] <= 0xcf && m[1] != 0xc4 && m[1] != 0xc8 && m[1] != 0xcc) {
        unsigned char s[5];
        f.read((char *)s, 5);
        height = (s[1] << 8) | s[2];
        width = (s[3] << 8) | s[4];
        return;
      }
      f.seekg(len - 2, std::ios::cur);
    }
  }
  if (h[0] == 'P' && (h[1] == '5' || h[1] == '6' || h[1] == '2' || h[1] == '3')) {  // PNM
    f.clear();
    f.seekg(2);
    f >> width >> height;
    if (f) return;
  }
  fail("cannot read the size of image " + file + " (PNG, JPEG and PNM headers are understood)");
}

std::string HypothesisList::SerializeAsString() const {
  using namespace prototext::wire;
  std::string out;
  for (const ObjectHypothesis &h : hyp) {
    std::string m;
    f32(m, 1, h.x);
    f32(m, 2, h.y);
    f32(m, 3, h.scale);
    f32(m, 4, h.score);
    boolean(m, 5, h.flip);
    bytes(out, 1, m);
  }
  return out;
}

HypothesisList HypothesisList::Parse(const std::string &b) {
  HypothesisList l;
  size_t i = 0;
  auto varint = [&](size_t &k) {
    uint64_t v = 0;
    int shift = 0;
    while (k < b.size()) {
      unsigned char c = (unsigned char)b[k++];
      v |= (uint64_t)(c & 0x7f) << shift;
      if (!(c & 0x80)) break;
      shift += 7;
    }
    return v;
  };
  while (i < b.size()) {
    uint64_t key = varint(i);
    if (key != ((1 << 3) | 2)) fail("HypothesisList: unexpected field");
    size_t len = (size_t)varint(i), end = i + len;
    ObjectHypothesis h;
    while (i < end) {
      uint64_t k = varint(i);
      int field = (int)(k >> 3), type = (int)(k & 7);
      if (type == 5) {
        float v;
        memcpy(&v, b.data() + i, 4);
        i += 4;
        if (field == 1) h.x = v;
        if (field == 2) h.y = v;
        if (field == 3) h.scale = v;
        if (field == 4) h.score = v;
      } else if (type == 0) {
        uint64_t v = varint(i);
        if (field == 5) h.flip = v != 0;
      } else {
        fail("HypothesisList: unsupported wire type");
      }
    }
    l.hyp.push_back(h);
  }
  return l;
}

// ---- PartApp ---------------------------------------------------------------------------------------------------------

std::string PartApp::getScoreGridFileName(int imgidx, int pidx, bool flip) const {
  return m_exp_param.scoregrid_dir + "/imgidx" + std::to_string(imgidx) + "-pidx" + std::to_string(pidx) + "-o" +
         std::to_string((int)flip) + "-scoregrid.mat";
}

void PartApp::init(const std::string &expopt) {
  prototext::Node n = prototext::parse_file(expopt);
  ExpParam &e = m_exp_param;
  for (size_t i = 0; i < n.count("test_dataset"); ++i) e.test_dataset.push_back(n.get("test_dataset", i).scalar);
  e.log_dir = n.str("log_dir");
  e.part_conf = n.str("part_conf");
  e.min_object_scale = (float)n.num("min_object_scale", 1);
  e.max_object_scale = (float)n.num("max_object_scale", 1);
  e.num_scale_steps = (unsigned)n.num("num_scale_steps", 1);
  e.min_part_rotation = (float)n.num("min_part_rotation", -180);
  e.max_part_rotation = (float)n.num("max_part_rotation", 180);
  e.num_rotation_steps = (unsigned)n.num("num_rotation_steps", 48);
  e.flip_orientation = n.boolean("flip_orientation", false);
  e.num_pose_samples = (int)n.num("num_pose_samples", 0);
  e.strip_border_detections = (float)n.num("strip_border_detections", 0);
  e.roi_save_num_samples = (float)n.num("roi_save_num_samples", 1000);
  e.use_pairwise = n.boolean("use_pairwise", true);
  e.save_part_marginals = n.boolean("save_part_marginals", false);
  e.save_part_marginals_local_max = n.boolean("save_part_marginals_local_max", false);
  e.save_part_detections_local_max = n.boolean("save_part_detections_local_max", false);
  e.interpolate = n.boolean("interpolate", false);
  e.force_recompute_scores = n.boolean("force_recompute_scores", true);
  e.use_torso_pos_prior = n.boolean("use_torso_pos_prior", false);
  e.save_root_marginal = n.boolean("save_root_marginal", false);
  e.torso_pos_prior_weight = (float)n.num("torso_pos_prior_weight", 1);
  e.pred_unary_rot = n.boolean("pred_unary_rot", false);
  e.pred_unary_pos = n.boolean("pred_unary_pos", false);
  e.use_dpm_torso = n.boolean("use_dpm_torso", false);
  e.use_dpm_head = n.boolean("use_dpm_head", false);
  e.use_dpm_unary = n.boolean("use_dpm_unary", false);
  e.do_dpm_rot = n.boolean("do_dpm_rot", false);
  e.use_gt_torso = n.boolean("use_gt_torso", false);
  e.pred_unary_rot_weight = (float)n.num("pred_unary_rot_weight", 1);
  e.pred_unary_pos_weight = (float)n.num("pred_unary_pos_weight", 1);
  e.dpm_torso_weight = (float)n.num("dpm_torso_weight", 1);
  e.dpm_head_weight = (float)n.num("dpm_head_weight", 1);
  e.dpm_unary_weight = (float)n.num("dpm_unary_weight", 1);
  e.rootidx_det = (unsigned)n.num("rootidx_det", 1000);
  if (n.has("torso_det_test_dir")) e.torso_det_test_dir = complete_relative_path(n.str("torso_det_test_dir"), expopt);
  if (n.has("test_dpm_torso_dir")) e.test_dpm_torso_dir = complete_relative_path(n.str("test_dpm_torso_dir"), expopt);
  if (n.has("test_dpm_unary_dir")) e.test_dpm_unary_dir = complete_relative_path(n.str("test_dpm_unary_dir"), expopt);
  if (e.log_dir.empty()) fail("expopt: log_dir is not set");
  e.log_dir = complete_relative_path(e.log_dir, expopt);
  // init_setpath, partapp.cpp:313-441
  e.log_subdir = n.has("log_subdir") ? n.str("log_subdir") : basename_noext(expopt);
  const std::string base = e.log_dir + "/" + e.log_subdir;
  e.class_dir = n.has("class_dir") ? complete_relative_path(n.str("class_dir"), expopt) : base + "/class";
  e.scoregrid_dir = n.has("scoregrid_dir") ? complete_relative_path(n.str("scoregrid_dir"), expopt) : base + "/test_scoregrid";
  e.spatial_dir = n.has("spatial_dir") ? complete_relative_path(n.str("spatial_dir"), expopt) : base + "/spatial";
  e.pred_data_test_dir = n.has("pred_data_test_dir") ? complete_relative_path(n.str("pred_data_test_dir"), expopt)
                                                     : base + "/pred_data_test";
  if (e.num_pose_samples != 0) fail("num_pose_samples must be 0 (findrot.cpp:680-682 asserts)");

  // part_conf (PartConfig.proto)
  if (e.part_conf.empty()) fail("expopt: part_conf is not set");
  prototext::Node pc = prototext::parse_file(complete_relative_path(e.part_conf, expopt));
  for (size_t i = 0; i < pc.count("part"); ++i) {
    const prototext::Node &p = pc.msg("part", i);
    PartDef d;
    d.part_id = (int)p.num("part_id", 0);
    d.is_root = p.boolean("is_root", false);
    d.is_detect = p.boolean("is_detect", true);
    d.is_upright = p.boolean("is_upright", false);
    for (size_t k = 0; k < p.count("part_pos"); ++k) d.part_pos.push_back(atoi(p.get("part_pos", k).scalar.c_str()));
    m_part_conf.part.push_back(d);
  }
  for (size_t i = 0; i < pc.count("joint"); ++i) {
    const prototext::Node &j = pc.msg("joint", i);
    JointDef d;
    d.child_idx = (int)j.num("child_idx", 0);
    d.parent_idx = (int)j.num("parent_idx", 0);
    d.type = j.str("type", "Gaussian");
    d.num_joint_types = (unsigned)j.num("num_joint_types", 1);
    m_part_conf.joint.push_back(d);
  }
  for (size_t p = 0; p < m_part_conf.part.size(); ++p)  // aux.cpp:67-69
    if (m_part_conf.part[p].part_id != (int)p + 1) fail("part_conf: part_id must equal part index + 1 (aux.cpp:68)");
  m_rootpart_idx = -1;
  for (size_t p = 0; p < m_part_conf.part.size(); ++p)
    if (m_part_conf.part[p].is_detect && m_part_conf.part[p].is_root) {
      if (m_rootpart_idx != -1) fail("part_conf: more than one root part (findrot.cpp:786)");
      m_rootpart_idx = (int)p;
    }

  // window_param.txt (partapp.cpp:608-615): only the bbox offsets matter here; absent file -> defaults (0, 0)
  const std::string wp = e.class_dir + "/window_param.txt";
  if (file_exists(wp)) {
    prototext::Node w = prototext::parse_file(wp);
    m_window_param.bbox_offset_x = w.num("bbox_offset_x", 0);
    m_window_param.bbox_offset_y = w.num("bbox_offset_y", 0);
  }

  // test image list: .al (XML, <image><name>..</name>) or .idl ("file": ...;) annotation lists (libAnnotation)
  for (const std::string &ds : e.test_dataset) {
    const std::string path = complete_relative_path(ds, expopt);
    std::ifstream f(path.c_str());
    if (!f) fail("cannot open test_dataset " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    std::vector<std::string> names;
    std::vector<std::vector<AnnoPoint> > points;
    if (text.find("<name>") != std::string::npos) {
      // <annotation><image><name>..</name></image><annorect>..<annopoints><point><id/><x/><y/></point>..
      auto element = [](const std::string &t, const char *tag, size_t from, size_t to, size_t &next) -> std::string {
        const std::string open = std::string("<") + tag + ">", close = std::string("</") + tag + ">";
        const size_t a = t.find(open, from);
        if (a == std::string::npos || a >= to) return next = std::string::npos, std::string();
        const size_t b = t.find(close, a);
        if (b == std::string::npos || b > to) return next = std::string::npos, std::string();
        next = b + close.size();
        return t.substr(a + open.size(), b - a - open.size());
      };
      size_t pos = 0;
      while ((pos = text.find("<name>", pos)) != std::string::npos) {
        size_t end = text.find("</name>", pos);
        if (end == std::string::npos) break;
        names.push_back(text.substr(pos + 6, end - pos - 6));
        pos = end + 7;
        // the first annorect of this annotation, up to the next image name
        size_t stop = text.find("<name>", pos);
        if (stop == std::string::npos) stop = text.size();
        std::vector<AnnoPoint> pts;
        size_t nx;
        const size_t r0 = text.find("<annorect>", pos);
        if (r0 != std::string::npos && r0 < stop) {
          size_t r1 = text.find("</annorect>", r0);
          if (r1 == std::string::npos || r1 > stop) r1 = stop;
          size_t q = r0;
          for (;;) {
            const std::string pt = element(text, "point", q, r1, nx);
            if (nx == std::string::npos) break;
            q = nx;
            size_t dummy;
            AnnoPoint ap;  // getElementDataInt: atoi of the element text
            ap.id = atoi(element(pt, "id", 0, pt.size(), dummy).c_str());
            ap.x = atoi(element(pt, "x", 0, pt.size(), dummy).c_str());
            ap.y = atoi(element(pt, "y", 0, pt.size(), dummy).c_str());
            pts.push_back(ap);
          }
        }
        points.push_back(pts);
      }
    } else {
      std::istringstream ls(text);
      std::string line;
      while (std::getline(ls, line)) {
        size_t a = line.find('"'), b = a == std::string::npos ? a : line.find('"', a + 1);
        if (b != std::string::npos) names.push_back(line.substr(a + 1, b - a - 1));
      }
    }
    points.resize(names.size());
    for (size_t k = 0; k < names.size(); ++k) {  // convertFullPath, partapp.cpp:87-107
      std::string nm = names[k];
      if (!file_exists(nm)) nm = dirname_of(path) + "/" + nm;
      if (!file_exists(nm)) fail("image file not found: " + nm);
      m_test_annolist.push_back(nm);
      m_test_annopoints.push_back(points[k]);
    }
  }
}

// ---- joints ----------------------------------------------------------------------------------------------------------

static double mat_scalar(const std::vector<mat5::Var> &vars, const std::string &name, const std::string &path) {
  const mat5::Var &v = mat5::find(vars, name, path);
  if (v.numel() != 1) fail(path + ": '" + name + "' is not a scalar");
  return v.at(0);
}

void load_joint(const PartApp &app, int jidx, Joint &joint, int tidx) {
  const ExpParam &ep = app.m_exp_param;
  if (jidx < 0 || jidx >= (int)app.m_part_conf.joint.size()) fail("load_joint: joint index out of range");
  if (ep.spatial_dir.empty() || !dir_exists(ep.spatial_dir)) fail("spatial_dir does not exist: " + ep.spatial_dir);
  const int child = app.m_part_conf.joint[jidx].child_idx, parent = app.m_part_conf.joint[jidx].parent_idx;
  const std::string stem = ep.spatial_dir + "/joint_" + std::to_string(child) + "_" + std::to_string(parent);
  std::string file = stem + (tidx > -1 ? "_tidx_" + std::to_string(tidx) : "") + ".mat";
  if (!file_exists(file)) file = stem + ".mat";  // "Empty joint mixture component" fallback, learnparam.cpp:113-117
  std::vector<mat5::Var> vars = mat5::load(file);
  const double d_type = mat_scalar(vars, "type", file);
  if ((int)mat_scalar(vars, "child_idx", file) != child || (int)mat_scalar(vars, "parent_idx", file) != parent)
    fail(file + ": child_idx / parent_idx do not match part_conf (learnparam.cpp:133)");
  const mat5::Var &oc = mat5::find(vars, "offset_c", file), &op = mat5::find(vars, "offset_p", file);
  const mat5::Var &C = mat5::find(vars, "C", file);
  if (oc.numel() != 2 || op.numel() != 2) fail(file + ": offsets must have two elements");
  if (C.dims.size() != 2 || C.dims[0] != 2 || C.dims[1] != 2) fail(file + ": C must be 2x2");
  joint.type = (int)d_type;
  joint.mix_comp_id = tidx > -1 ? tidx : 0;
  joint.child_idx = child;
  joint.parent_idx = parent;
  for (int k = 0; k < 2; ++k) {
    joint.offset_c[k] = oc.at(k);
    joint.offset_p[k] = op.at(k);
  }
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 2; ++k) joint.C[i][k] = C.at(i * 2 + k);
  if (joint.type == Joint::ROT_GAUSSIAN) {
    joint.rot_sigma = mat_scalar(vars, "rot_sigma", file);
    joint.rot_mean = mat_scalar(vars, "rot_mean", file);
  }
}

void loadJoints(const PartApp &app, std::vector<Joint> &joints, bool flip, int imgidx) {
  const int nJoints = (int)app.m_part_conf.joint.size(), nParts = (int)app.m_part_conf.part.size();
  joints.assign(nJoints, Joint());
  std::vector<int> tidx(nJoints, -1);
  // mixtures of pairwise terms (aux.cpp:76-94).  The reference spawns the MATLAB predictor here (predictFactors);
  // this host reads its output file and fails if it is missing.
  if (imgidx > -1 && nJoints > 0 && app.m_part_conf.joint[0].num_joint_types > 1) {
    const std::string file = app.m_exp_param.pred_data_test_dir + "/testlist_pred_pwise_imgidx_" + std::to_string(imgidx) + ".mat";
    if (!file_exists(file))
      fail("poselet-conditioned joints need " + file + " (written by the MATLAB predictor, objectdetect_icps.cpp:608-625)");
    std::vector<mat5::Var> vars = mat5::load(file);
    const mat5::Var &cl = mat5::find(vars, "clusidx_test", file);
    if ((int)cl.numel() < nJoints) fail(file + ": clusidx_test is shorter than the joint list");
    for (int j = 0; j < nJoints; ++j) tidx[j] = (int)cl.at(j);
  }
  for (int j = 0; j < nJoints; ++j) {
    load_joint(app, j, joints[j], tidx[j]);
    if (flip) {  // aux.cpp:102-119, done by the library's ps_flip_joint
      ps_joint pj = to_ps_joints(std::vector<Joint>(1, joints[j]))[0];
      ps_flip_joint(&pj);
      for (int k = 0; k < 2; ++k) {
        joints[j].offset_c[k] = pj.offset_c[k];
        joints[j].offset_p[k] = pj.offset_p[k];
      }
      joints[j].C[0][0] = pj.C[0]; joints[j].C[0][1] = pj.C[1]; joints[j].C[1][0] = pj.C[2]; joints[j].C[1][1] = pj.C[3];
      joints[j].rot_mean = pj.rot_mean;
    }
    joints[j].parent_idx--;  // part ids -> indices (aux.cpp:123-124)
    joints[j].child_idx--;
    if (joints[j].child_idx < 0 || joints[j].child_idx >= nParts || joints[j].parent_idx < 0 || joints[j].parent_idx >= nParts)
      fail("joint refers to a part that does not exist (aux.cpp:126-127)");
    if (nJoints != nParts - 1) fail("need exactly num_parts-1 joints (aux.cpp:129)");
    const double (*C)[2] = joints[j].C;
    joints[j].detC = C[0][0] * C[1][1] - C[1][0] * C[0][1];
    if (!(joints[j].detC > 0)) fail("joint covariance must have a positive determinant (aux.cpp:133)");
    joints[j].invC[0][0] = C[1][1] / joints[j].detC; joints[j].invC[0][1] = -C[0][1] / joints[j].detC;
    joints[j].invC[1][0] = -C[1][0] / joints[j].detC; joints[j].invC[1][1] = C[0][0] / joints[j].detC;
  }
}

// ---- predictor outputs (objectdetect_icps.cpp) -------------------------------------------------------------------------

static const mat5::Var &load_matrix(std::vector<mat5::Var> &vars, const std::string &file, const std::string &name, size_t min_rows,
                                    size_t min_cols) {
  if (!file_exists(file)) fail("missing predictor output " + file + " (the reference's MATLAB side writes it, icps.cpp:608-625)");
  vars = mat5::load(file);
  const mat5::Var &m = mat5::find(vars, name, file);
  if (m.dims.size() != 2 || m.dims[0] < min_rows || m.dims[1] < min_cols)
    fail(file + ": '" + name + "' must be at least " + std::to_string(min_rows) + " x " + std::to_string(min_cols));
  return m;
}

void getRotParams(const PartApp &app, int imgidx, std::vector<double> &rot_params, bool bTest) {
  const int P = (int)app.m_part_conf.part.size();
  const std::string list = bTest ? "test" : "train", dir = app.m_exp_param.pred_data_test_dir;
  rot_params.assign((size_t)P * 3, 0.0);
  std::vector<mat5::Var> va, vb;
  const mat5::Var &all = load_matrix(va, dir + "/" + list + "list_params_rot_imgidx_" + std::to_string(imgidx) + ".mat", "rot_" + list, P, 2);
  const mat5::Var &clus = load_matrix(vb, dir + "/" + list + "list_pred_rot_imgidx_" + std::to_string(imgidx) + ".mat", "clusidx_" + list, P, 1);
  for (int p = 0; p < P; ++p)
    if (app.m_part_conf.part[p].is_detect) {
      rot_params[(size_t)p * 3 + 0] = all.at((size_t)p * all.dims[1] + 0);
      const double sd = all.at((size_t)p * all.dims[1] + 1);
      rot_params[(size_t)p * 3 + 1] = sd * sd;  // square(), icps.cpp:219
      rot_params[(size_t)p * 3 + 2] = clus.at((size_t)p * clus.dims[1]);
    }
}

void getPosParams(const PartApp &app, int imgidx, std::vector<double> &pos_params, int rootpart_idx, bool bTest) {
  const int P = (int)app.m_part_conf.part.size();
  const std::string list = bTest ? "test" : "train", dir = app.m_exp_param.pred_data_test_dir;
  pos_params.assign((size_t)P * 5, 0.0);
  std::vector<mat5::Var> va, vb;
  const mat5::Var &all = load_matrix(va, dir + "/" + list + "list_params_pos_imgidx_" + std::to_string(imgidx) + ".mat", "pos_" + list, P, 4);
  const mat5::Var &clus = load_matrix(vb, dir + "/" + list + "list_pred_pos_imgidx_" + std::to_string(imgidx) + ".mat", "clusidx_" + list, P, 1);
  for (int p = 0; p < P; ++p) {
    if (p == rootpart_idx || !app.m_part_conf.part[p].is_detect) continue;
    const size_t o = (size_t)p * all.dims[1];
    pos_params[(size_t)p * 5 + 0] = all.at(o + 0);
    pos_params[(size_t)p * 5 + 1] = all.at(o + 1);
    pos_params[(size_t)p * 5 + 2] = all.at(o + 2) * all.at(o + 2);
    pos_params[(size_t)p * 5 + 3] = all.at(o + 3) * all.at(o + 3);
    pos_params[(size_t)p * 5 + 4] = clus.at((size_t)p * clus.dims[1]);
  }
}

void getRootPosDet(const PartApp &app, int imgidx, int rootpart_idx, double rootpos_det[2], bool bTest) {
  const ExpParam &ep = app.m_exp_param;
  if (ep.use_gt_torso) {
    // icps.cpp:292-300: part_pos of get_part_bbox(m_test_annolist[imgidx][0], part(rootpart_idx)); the return value (an
    // invalid x axis) is ignored and part_pos is assigned before the axis either way (partdef.cpp:249, :315).
    // < 3 points: (sum of the points) * (1.0 / n) (partdef.cpp:140-158); else the centre of their bounding box
    // (:128-138).  `int root_pos_x = part_pos(0)`: truncation towards zero.
    const std::vector<int> &ids = app.m_part_conf.part.at((size_t)rootpart_idx).part_pos;
    if (ids.empty()) fail("use_gt_torso: the root part has no part_pos in part_conf");
    if ((size_t)imgidx >= app.m_test_annopoints.size()) fail("use_gt_torso: no annotation for image " + std::to_string(imgidx));
    const std::vector<AnnoPoint> &pts = app.m_test_annopoints[(size_t)imgidx];
    double sx = 0, sy = 0, x0 = INFINITY, x1 = -INFINITY, y0 = INFINITY, y1 = -INFINITY;
    for (int id : ids) {
      const AnnoPoint *p = nullptr;
      for (const AnnoPoint &q : pts)
        if (q.id == id) {
          p = &q;
          break;
        }
      if (!p) fail("use_gt_torso: image " + std::to_string(imgidx) + " has no annopoint " + std::to_string(id) + " (partdef.cpp:146)");
      sx += p->x; sy += p->y;
      x0 = std::min(x0, (double)p->x); x1 = std::max(x1, (double)p->x);
      y0 = std::min(y0, (double)p->y); y1 = std::max(y1, (double)p->y);
    }
    double px, py;
    if (ids.size() < 3) {
      const double inv = 1.0 / (double)ids.size();
      px = sx * inv; py = sy * inv;
    } else {
      px = 0.5 * (x0 + x1); py = 0.5 * (y0 + y1);
    }
    rootpos_det[0] = (double)(int)px;
    rootpos_det[1] = (double)(int)py;
    return;
  }
  if (!bTest) fail("getRootPosDet: only the test list is read on this path");
  if (ep.torso_det_test_dir.empty()) fail("pred_unary_pos needs torso_det_test_dir (icps.cpp:306)");
  std::vector<mat5::Var> vars;
  const mat5::Var &best = load_matrix(vars, ep.torso_det_test_dir + "/pose_est_imgidx" + pad_zeros(imgidx, 4) + ".mat", "best_conf",
                                      (size_t)ep.rootidx_det + 1, 6);
  // int root_pos_x = best_conf(rootidx_det, 4): truncation towards zero, icps.cpp:317-318
  rootpos_det[0] = (double)(int)best.at((size_t)ep.rootidx_det * best.dims[1] + 4);
  rootpos_det[1] = (double)(int)best.at((size_t)ep.rootidx_det * best.dims[1] + 5);
}

void loadDPMScoreGrid(const std::string &dir, int imgidx, std::vector<std::vector<float> > &grids, int H, int W, bool bIsCell,
                      int expected) {
  const std::string file = dir + "/imgidx_" + pad_zeros(imgidx + 1, 4) + ".mat";
  if (!file_exists(file)) fail("missing DPM score grid " + file + " (icps.cpp:555)");
  std::vector<mat5::Var> vars = mat5::load(file);
  const mat5::Var &sg = mat5::find(vars, "scoregrid", file);
  auto take = [&](const mat5::Var &g) {
    if (g.dims.size() != 2 || (int)g.dims[0] != H || (int)g.dims[1] != W)
      fail(file + ": scoregrid must be " + std::to_string(H) + " x " + std::to_string(W) + " like the image (icps.cpp:500-501)");
    std::vector<float> v((size_t)H * W);
    for (size_t i = 0; i < v.size(); ++i) v[i] = (float)g.at(i);
    grids.push_back(std::move(v));
  };
  grids.clear();
  if (bIsCell) {
    if (sg.cls != mat5::mxCELL) fail(file + ": scoregrid must be a cell array (icps.cpp:560-565)");
    if ((int)sg.cells.size() != expected) fail(file + ": scoregrid has " + std::to_string(sg.cells.size()) + " grids, expected " +
                                               std::to_string(expected) + " (icps.cpp:570)");
    for (const mat5::Var &g : sg.cells) take(g);
  } else {
    take(sg);
  }
}

// ---- inference ---------------------------------------------------------------------------------------------------------

void computeRotJointMarginal(const ExpParam &ep, FloatGrid3 &child, FloatGrid3 &parent, const double offset_c_10[2],
                             const double offset_p_01[2], const double C[2][2], double rot_mean, double rot_sigma,
                             double scale, bool bIsSparse) {
  PartApp app;
  app.m_exp_param = ep;
  app.m_part_conf.part.resize(2);
  app.m_part_conf.part[0].is_root = true;
  if (child.R != (int)ep.num_rotation_steps) fail("computeRotJointMarginal: rotation count mismatch (findrot.cpp:311)");
  ps_ctx *ctx = get_ctx(app, child.H, child.W, 0, false);
  parent = FloatGrid3(child.R, child.H, child.W);
  const double Cf[4] = {C[0][0], C[0][1], C[1][0], C[1][1]};
  check(ctx, ps_message(ctx, child.data(), parent.data(), PS_MEM_HOST, offset_c_10, offset_p_01, Cf, rot_mean, rot_sigma,
                        scale, bIsSparse ? 1 : 0), "ps_message");
}

void computePosJointMarginal(const ExpParam &ep, FloatGrid3 &child, FloatGrid3 &parent, const double offset[2],
                             const double C[2][2], double scale, bool bIsSparse) {
  PartApp app;
  app.m_exp_param = ep;
  app.m_exp_param.num_rotation_steps = (unsigned)child.R;  // the slice count of the grids handed in
  app.m_part_conf.part.resize(2);
  app.m_part_conf.part[0].is_root = true;
  ps_ctx *ctx = get_ctx(app, child.H, child.W, 0, false);
  parent = FloatGrid3(child.R, child.H, child.W);
  const double Cf[4] = {C[0][0], C[0][1], C[1][0], C[1][1]};
  check(ctx, ps_pos_message(ctx, child.data(), parent.data(), PS_MEM_HOST, offset, Cf, scale, bIsSparse ? 1 : 0),
        "ps_pos_message");
}

void computeRootPosteriorRot(const PartApp &app, std::vector<std::vector<FloatGrid3> > &log_part_detections,
                             FloatGrid3 &root_part_posterior, int rootpart_idx, std::vector<Joint> joints, bool flip,
                             bool bIsSparse, int imgidx, std::vector<std::vector<PartHyp> > &best_part_hyp,
                             bool bSaveMarginals) {
  const int P = (int)app.m_part_conf.part.size(), S = (int)app.m_exp_param.num_scale_steps;
  if ((int)log_part_detections.size() != P || (int)log_part_detections[0].size() != S)
    fail("computeRootPosteriorRot: log_part_detections must be [parts][scales] (findrot.cpp:490-492)");
  const int H = log_part_detections[0][0].H, W = log_part_detections[0][0].W;
  ps_ctx *ctx = get_ctx(app, H, W, rootpart_idx, bSaveMarginals);
  std::vector<ps_joint> pj = to_ps_joints(joints);
  check(ctx, ps_set_joints(ctx, pj.data(), (int)pj.size()), "ps_set_joints");
  for (int p = 0; p < P; ++p)
    for (int s = 0; s < S; ++s)
      check(ctx, ps_set_unary(ctx, p, s, log_part_detections[p][s].data(), PS_MEM_HOST, 0), "ps_set_unary");
  check(ctx, ps_infer(ctx, (bIsSparse ? PS_INFER_SPARSE : 0) | PS_INFER_LOCAL_MAX), "ps_infer");
  collect_results(app, ctx, root_part_posterior, best_part_hyp, H, W);
  for (int p = 0; p < P; ++p)  // the reference masks its argument in place (findrot.cpp:509-551)
    for (int s = 0; s < S; ++s)
      check(ctx, ps_get_unary(ctx, p, s, log_part_detections[p][s].data(), PS_MEM_HOST), "ps_get_unary");
  if (bSaveMarginals) save_marginals(app, ctx, imgidx, flip, H, W);
}

void findObjectRoiHelper(PartApp app, const int roi[4], double scale, const std::vector<ScoreGrid> &score_grid,
                         std::vector<Joint> joints, std::vector<std::vector<PartHyp> > &best_part_det,
                         std::vector<std::vector<PartHyp> > &best_part_hyp) {
  const int P = (int)app.m_part_conf.part.size(), R = (int)app.m_exp_param.num_rotation_steps;
  if ((int)score_grid.size() != P) fail("findObjectRoiHelper: one ScoreGrid per part");
  // "hacky way to set scale used during inference" (objectdetect_roi.cpp:82-85)
  app.m_exp_param.min_object_scale = app.m_exp_param.max_object_scale = (float)scale;
  app.m_exp_param.num_scale_steps = 1;
  const ExpParam &ep = app.m_exp_param;
  const int roi_x1 = roi[0], roi_y1 = roi[1];
  const int W = std::abs(roi[2] - roi[0]) + 1, H = std::abs(roi[3] - roi[1]) + 1;  // :150-151
  if (app.m_rootpart_idx < 0) fail("root part not found");
  ps_ctx *ctx = get_ctx(app, H, W, app.m_rootpart_idx, false);
  const int K = (int)ep.roi_save_num_samples;
  best_part_det.assign(P, std::vector<PartHyp>());
  std::vector<float> rows((size_t)std::max(K, 1) * 4);
  for (int p = 0; p < P; ++p) {
    const ScoreGrid &g = score_grid[p];
    if ((int)g.Tig.size() != R * 9 || g.cells.size() != (size_t)R * g.gh * g.gw) fail("findObjectRoiHelper: ScoreGrid shape");
    // transform_grid_fixed_size(TM_DIRECT) + clip_scores_fill (:205-215)
    check(ctx, ps_set_unary_compact_raw(ctx, p, 0, g.cells.data(), g.gh, g.gw, g.Tig.data(), PS_MEM_HOST), "ps_set_unary_compact_raw");
    int n = 0;  // maxima of the part detections, ROI offset added (:226-236)
    check(ctx, ps_unary_local_max(ctx, p, 0, K, rows.data(), &n), "ps_unary_local_max");
    for (int i = 0; i < n; ++i) {
      PartHyp h;
      h.m_scaleidx = 0;
      h.m_scale = (float)scale_from_index(ep, 0);
      h.m_rotidx = (int)rows[(size_t)i * 4];
      h.m_rot = (float)rot_from_index(ep, h.m_rotidx);
      h.m_x = (int)rows[(size_t)i * 4 + 1] + roi_x1;
      h.m_y = (int)rows[(size_t)i * 4 + 2] + roi_y1;
      h.m_score = rows[(size_t)i * 4 + 3];
      best_part_det[p].push_back(h);
    }
    check(ctx, ps_log_unary(ctx, p, 0), "ps_log_unary");  // computeLogGrid (:240-242)
  }
  std::vector<ps_joint> pj = to_ps_joints(joints);
  check(ctx, ps_set_joints(ctx, pj.data(), (int)pj.size()), "ps_set_joints");
  check(ctx, ps_infer(ctx, PS_INFER_SPARSE | PS_INFER_LOCAL_MAX), "ps_infer");  // bIsSparse = true, no marginals (:246-262)
  FloatGrid3 root_part_posterior;
  collect_results(app, ctx, root_part_posterior, best_part_hyp, H, W);
  for (std::vector<PartHyp> &v : best_part_hyp)  // :265-271
    for (PartHyp &h : v) {
      h.m_x += roi_x1;
      h.m_y += roi_y1;
    }
}

// loadScoreGrid (partapp.cpp:830-903) + unary prep (findrot.cpp:834-845 / findpos.cpp:377-384) of every detected part of
// one image, on the device.
static void load_image_unaries(const PartApp &app, ps_ctx *ctx, int imgidx, bool flip) {
  const ExpParam &ep = app.m_exp_param;
  const int P = (int)app.m_part_conf.part.size(), S = (int)ep.num_scale_steps, R = (int)ep.num_rotation_steps;
  std::vector<float> cells;
  std::vector<double> Tig((size_t)R * 9);
  for (int p = 0; p < P; ++p) {
    if (!app.m_part_conf.part[p].is_detect) continue;
    const std::string file = app.getScoreGridFileName(imgidx, p, flip);
    std::vector<mat5::Var> vars;
    {
      PhaseTimer t(g_ns_load);
      vars = mat5::load(file);
    }
    const mat5::Var &cg = mat5::find(vars, "cell_scoregrid", file);
    const mat5::Var &Ti2 = mat5::find(vars, "transform_Ti2", file), &T2g = mat5::find(vars, "transform_T2g", file);
    if (cg.cls != mat5::mxCELL || cg.dims.size() != 2 || (int)cg.dims[0] != S || (int)cg.dims[1] != R)
      fail(file + ": cell_scoregrid must be a [scales][rotations] cell array (findrot.cpp:811-812)");
    if (Ti2.numel() != (size_t)S * R * 9 || T2g.numel() != (size_t)S * R * 9) fail(file + ": transforms must be [S][R][3][3]");
    for (int s = 0; s < S; ++s) {
      const mat5::Var &c0 = cg.cells[(size_t)s * R];
      const size_t gh = c0.dims.at(0), gw = c0.dims.at(1);
      cells.assign((size_t)R * gh * gw, 0.0f);
      for (int r = 0; r < R; ++r) {
        const mat5::Var &c = cg.cells[(size_t)s * R + r];
        if (c.dims.size() != 2 || c.dims[0] != gh || c.dims[1] != gw) fail(file + ": score grids of one scale differ in size");
        if (c.cls == mat5::mxSINGLE && c.f32.size() == gh * gw)
          memcpy(&cells[(size_t)r * gh * gw], c.f32.data(), gh * gw * sizeof(float));
        else
          for (size_t i = 0; i < gh * gw; ++i) cells[(size_t)r * gh * gw + i] = (float)c.at(i);
        // Tig = prod(Ti2, T2g) in double (partapp.cpp:881-887; array_to_matrix widens the floats)
        const size_t o = ((size_t)s * R + r) * 9;
        for (int i = 0; i < 3; ++i)
          for (int k = 0; k < 3; ++k) {
            double t = 0;
            for (int l = 0; l < 3; ++l) t += Ti2.at(o + i * 3 + l) * T2g.at(o + l * 3 + k);
            Tig[(size_t)r * 9 + i * 3 + k] = t;
          }
      }
      // cudaMemcpyAsync from pageable memory returns once `cells` has been staged, so the buffer can be reused at once
      PhaseTimer t(g_ns_ingest);
      check(ctx, ps_set_unary_compact(ctx, p, s, cells.data(), (int)gh, (int)gw, Tig.data(), PS_MEM_HOST), "ps_set_unary_compact");
    }
  }
}

void findObjectImagePosJoints(const PartApp &app, int imgidx, bool flip, HypothesisList &hypothesis_list) {
  const ExpParam &ep = app.m_exp_param;
  const int S = (int)ep.num_scale_steps;
  int W = 0, H = 0;
  image_size(app.m_test_annolist[imgidx], W, H);
  std::vector<Joint> joints;
  loadJoints(app, joints, flip);  // no per-image joint types on this path (findpos.cpp:350)
  for (const Joint &j : joints)
    if (j.type != Joint::POS_GAUSSIAN) fail("findObjectImagePosJoints: POS_GAUSSIAN joints only (aux.cpp:346)");
  const int rootpart_idx = app.m_rootpart_idx;
  if (rootpart_idx < 0) fail("root part not found (findpos.cpp:359-362)");
  ps_ctx *ctx = get_ctx(app, H, W, rootpart_idx, false);
  std::vector<ps_joint> pj = to_ps_joints(joints);
  check(ctx, ps_set_joints(ctx, pj.data(), (int)pj.size()), "ps_set_joints");
  load_image_unaries(app, ctx, imgidx, flip);
  check(ctx, ps_infer(ctx, PS_INFER_SPARSE | PS_INFER_ROOT_HYPS), "ps_infer");  // bIsSparse = true (:342)
  hypothesis_list.hyp.clear();
  std::vector<float> rows(1000 * 4);
  int n = 0;
  check(ctx, ps_get_root_hyps(ctx, rows.data(), 1000, &n), "ps_get_root_hyps");
  for (int i = 0; i < n; ++i) {  // findLocalMax(exp_param, grid, hypothesis_list, n), aux.cpp:263-300
    ObjectHypothesis h;
    h.scale = (float)scale_from_index(ep, (int)rows[4 * i]);
    h.x = rows[4 * i + 1];  // no bounding-box offset on this path (aux.cpp:280-281)
    h.y = rows[4 * i + 2];
    h.score = rows[4 * i + 3];
    h.flip = flip;
    hypothesis_list.hyp.push_back(h);
  }
  if (ep.save_root_marginal) {  // :437-451
    const std::string dir = ep.log_dir + "/" + ep.log_subdir + "/root_part_posterior";
    make_dirs(dir);
    FloatGrid3 rp(S, H, W);
    check(ctx, ps_get_root_posterior(ctx, rp.data(), PS_MEM_HOST), "ps_get_root_posterior");
    mat5::Writer w(dir + "/root_part_posterior_imgidx" + std::to_string(imgidx) + "_o" + std::to_string((int)flip) + ".mat");
    w.put("root_part_posterior", rp.data(), {(size_t)S, (size_t)H, (size_t)W});
  }
}

void findObjectImageRotJoints(const PartApp &app, int imgidx, bool flip, HypothesisList &hypothesis_list,
                              const std::string &qsPartMarginalsDir, const std::string &qsScoreGridDir,
                              const std::string &qsImgName) {
  const ExpParam &ep = app.m_exp_param;
  const int P = (int)app.m_part_conf.part.size(), R = (int)ep.num_rotation_steps;
  (void)qsScoreGridDir;
  int W = 0, H = 0;
  image_size(qsImgName, W, H);  // findrot.cpp:752-760
  std::vector<Joint> joints;
  {
    PhaseTimer t(g_ns_joints);
    loadJoints(app, joints, flip, imgidx);
  }
  for (const Joint &j : joints)
    if (j.type != Joint::ROT_GAUSSIAN) fail("only ROT_GAUSSIAN joints are supported (findrot.cpp:766)");
  const int rootpart_idx = app.m_rootpart_idx;
  if (rootpart_idx < 0) fail("root part not found (findrot.cpp:831)");
  const bool bSaveMarginals = ep.save_part_marginals;
  ps_ctx *ctx = get_ctx(app, H, W, rootpart_idx, bSaveMarginals);
  std::vector<ps_joint> pj = to_ps_joints(joints);
  check(ctx, ps_set_joints(ctx, pj.data(), (int)pj.size()), "ps_set_joints");

  load_image_unaries(app, ctx, imgidx, flip);

  // ---- conditioning of the unaries, in the reference's order (findrot.cpp:849-949) ----
  // The predictors (MATLAB poselet classifiers, DPM detectors) are outside this path: their per-image outputs are read
  // from disk; the adds run on the device (ps_add_unary_grid / ps_add_unary_table).
  std::vector<std::vector<float> > dpm;
  if (ep.use_dpm_torso) {  // :851-854, :883-890: log of the torso DPM grid, added to the root with dpm_torso_weight
    loadDPMScoreGrid(ep.test_dpm_torso_dir, imgidx, dpm, H, W, false);
    for (float &v : dpm[0]) {
      if (v < 0) fail("DPM torso score grid holds a negative value (computeLogGrid asserts, multi_array_op.hpp:160)");
      v = v == 0 ? -1e6f : (float)std::log((double)v);
    }
    if (!(ep.dpm_torso_weight > 0)) fail("dpm_torso_weight must be > 0 (icps.cpp:497)");
    check(ctx, ps_add_unary_grid(ctx, rootpart_idx, dpm[0].data(), 1, 0, ep.dpm_torso_weight, PS_MEM_HOST), "ps_add_unary_grid");
  }
  auto add_load_dpm = [&](float weight, int nrot_dpm, const std::string &parent, bool use_pidx, int pidx_only) {  // icps.cpp:445-486
    std::vector<float> flat;
    for (int p = 0; p < P; ++p) {
      if (pidx_only > -1 && pidx_only < P && pidx_only != p) continue;
      if (!app.m_part_conf.part[p].is_detect) continue;
      loadDPMScoreGrid(parent + (use_pidx ? "/pidx_" + pad_zeros(p, 4) : ""), imgidx, dpm, H, W, true, nrot_dpm);
      flat.clear();
      for (const std::vector<float> &g : dpm) flat.insert(flat.end(), g.begin(), g.end());
      check(ctx, ps_add_unary_grid(ctx, p, flat.data(), nrot_dpm, 1, weight, PS_MEM_HOST), "ps_add_unary_grid");
    }
  };
  if (ep.use_dpm_head) {  // :893-904
    const int headpart_idx = P == 22 ? 11 : (P == 12 ? 1 : 5);
    add_load_dpm(ep.dpm_head_weight, 1, ep.test_dpm_unary_dir + "/head", false, headpart_idx);
  }
  if (ep.use_dpm_unary) {  // :907-914
    if (ep.test_dpm_unary_dir.empty()) fail("use_dpm_unary needs test_dpm_unary_dir (findrot.cpp:910)");
    add_load_dpm(ep.dpm_unary_weight, ep.do_dpm_rot ? R : 1, ep.test_dpm_unary_dir, true, -1);
  }
  if (ep.pred_unary_rot) {  // :862-868, :921-927: rotation score of every detected part
    std::vector<double> rot_params;
    getRotParams(app, imgidx, rot_params, true);
    ps_config cfg = make_config(app, H, W, rootpart_idx, bSaveMarginals);
    std::vector<float> table((size_t)R);
    for (int p = 0; p < P; ++p) {
      if (!app.m_part_conf.part[p].is_detect) continue;  // log_rot_scores stays 0 there: unary + weight * 0
      ps_rot_score_table(&cfg, rot_params[(size_t)p * 3], rot_params[(size_t)p * 3 + 1], table.data());
      check(ctx, ps_add_unary_table(ctx, p, table.data(), 0, ep.pred_unary_rot_weight), "ps_add_unary_table");
    }
  }
  if (ep.pred_unary_pos) {  // :872-878, :934-941: position score of every detected non-root part
    std::vector<double> pos_params;
    double rootpos_det[2];
    getPosParams(app, imgidx, pos_params, rootpart_idx, true);
    getRootPosDet(app, imgidx, rootpart_idx, rootpos_det, true);
    std::vector<float> table((size_t)H * W);
    for (int p = 0; p < P; ++p) {
      if (p == rootpart_idx || !app.m_part_conf.part[p].is_detect) continue;
      const double *q = &pos_params[(size_t)p * 5];
      ps_pos_score_table(H, W, q[0], q[1], q[2], q[3], rootpos_det[0], rootpos_det[1], table.data());
      check(ctx, ps_add_unary_table(ctx, p, table.data(), 1, ep.pred_unary_pos_weight), "ps_add_unary_table");
    }
  }
  // torso position prior (findrot.cpp:945-948, icps.cpp:137-191; params from <class_dir>/torso_pos_prior.mat :41-42)
  if (ep.use_torso_pos_prior) {
    const std::string file = ep.class_dir + "/torso_pos_prior.mat";
    std::vector<mat5::Var> vars = mat5::load(file);
    const mat5::Var &pr = mat5::find(vars, "params", file);
    if (pr.numel() < 4) fail(file + ": params must hold mu_x, mu_y, var_x, var_y");
    std::vector<float> table((size_t)H * W);
    ps_torso_prior_table(H, W, pr.at(0), pr.at(1), pr.at(2), pr.at(3), ep.torso_pos_prior_weight, table.data());
    check(ctx, ps_add_unary_table(ctx, rootpart_idx, table.data(), 2, 1.0f), "ps_add_unary_table");
  }

  int flags = PS_INFER_SPARSE | PS_INFER_ROOT_HYPS | PS_INFER_KEEP_UNARIES;
  if (ep.save_part_marginals_local_max) flags |= PS_INFER_LOCAL_MAX;
  std::vector<float> best_conf((size_t)P * PS_HYP_VEC);
  {
    PhaseTimer t(g_ns_infer);
    if (ep.use_pairwise) check(ctx, ps_infer(ctx, flags), "ps_infer");
    else check(ctx, ps_max_states(ctx, flags & PS_INFER_LOCAL_MAX), "ps_max_states");
    check(ctx, ps_get_best_conf(ctx, best_conf.data()), "ps_get_best_conf");
  }
  PhaseTimer t_out(g_ns_out);
  make_dirs(qsPartMarginalsDir);
  {  // findrot.cpp:1005-1011
    mat5::Writer w(qsPartMarginalsDir + "/pose_est_imgidx" + pad_zeros(imgidx, 4) + ".mat");
    w.put("best_conf", best_conf.data(), {(size_t)P, (size_t)PS_HYP_VEC});
  }
  if (ep.save_part_marginals_local_max) {  // findrot.cpp:1015-1034
    mat5::Writer w(qsPartMarginalsDir + "/part_post_imgidx" + pad_zeros(imgidx, 4) + ".mat");
    const int cap = (int)ep.roi_save_num_samples + 1;
    std::vector<float> rows((size_t)cap * PS_HYP_VEC);
    for (int p = 0; p < P; ++p) {
      int n = 0;
      check(ctx, ps_get_part_hyps(ctx, p, rows.data(), cap, &n), "ps_get_part_hyps");
      w.put("part" + std::to_string(p), rows.data(), {(size_t)n, (size_t)PS_HYP_VEC});
    }
  }
  if (bSaveMarginals && ep.use_pairwise) save_marginals(app, ctx, imgidx, flip, H, W);

  // root hypotheses (findrot.cpp:1037-1048)
  hypothesis_list.hyp.clear();
  if (ep.use_pairwise) {
    std::vector<float> rows(1000 * 4);
    int n = 0;
    check(ctx, ps_get_root_hyps(ctx, rows.data(), 1000, &n), "ps_get_root_hyps");
    for (int i = 0; i < n; ++i) {
      ObjectHypothesis h;
      h.scale = (float)scale_from_index(ep, (int)rows[4 * i]);
      h.x = (float)(int)(rows[4 * i + 1] + app.m_window_param.bbox_offset_x);
      h.y = (float)(int)(rows[4 * i + 2] + app.m_window_param.bbox_offset_y);
      h.score = rows[4 * i + 3];
      h.flip = flip;
      hypothesis_list.hyp.push_back(h);
    }
  }
}

void set_parallelism(int gpus, int contexts_per_gpu) {
  g_gpus = gpus;
  g_ctx_per_gpu = contexts_per_gpu;
}

// One image (both orientations): findObjectDataset's loop body, aux.cpp:344-402.
static void find_object_image(const PartApp &app, int imgidx, const std::string &qsHypDir, const std::string &qsPartMarginalsDir) {
  const ExpParam &ep = app.m_exp_param;
  const int flip_count = ep.flip_orientation ? 2 : 1;
  // "find out what type of joints are used in the spatial model" (aux.cpp:346-359): the first joint decides
  bool bFindObjectRot = false;
  if (!app.m_part_conf.joint.empty()) {
    Joint joint;
    load_joint(app, 0, joint, app.m_part_conf.joint[0].num_joint_types > 1 ? 0 : -1);
    bFindObjectRot = joint.type == Joint::ROT_GAUSSIAN;
  }
  for (int flip = 0; flip < flip_count; ++flip) {
    HypothesisList hypothesis_list;
    if (bFindObjectRot)
      findObjectImageRotJoints(app, imgidx, flip != 0, hypothesis_list, qsPartMarginalsDir, ep.scoregrid_dir,
                               app.m_test_annolist[imgidx]);
    else
      findObjectImagePosJoints(app, imgidx, flip != 0, hypothesis_list);
    const std::string file = qsHypDir + getObjectHypFilename(imgidx, flip != 0);
    std::ofstream f(file.c_str(), std::ios::binary);
    if (!f) fail("cannot write " + file);
    const std::string bytes = hypothesis_list.SerializeAsString();
    f.write(bytes.data(), (std::streamsize)bytes.size());
  }
}

// The reference shards the image range over PROCESSES (--distribute / --ncpu / --batch_num, main.cpp:155-192: contiguous
// ranges of ceil(n / ncpu) images).  Here one process drives every GPU the same way: GPU g gets the g-th contiguous
// range, and several worker threads per GPU -- each with its own ps_ctx and stream -- draw images from that range, so
// reading and inflating the score-grid files of one image, the inference of another and the output files of a third
// overlap.  Outputs are per image, so the files are the same whatever the number of GPUs or workers.
void findObjectDataset(const PartApp &app, int firstidx, int lastidx) {
  const ExpParam &ep = app.m_exp_param;
  if (firstidx < 0 || firstidx > (int)app.m_test_annolist.size() || lastidx >= (int)app.m_test_annolist.size())
    fail("image index range outside the test list (aux.cpp:328-329)");
  const std::string qsHypDir = ep.log_dir + "/" + ep.log_subdir + "/object_hyp";
  const std::string qsPartMarginalsDir = ep.log_dir + "/" + ep.log_subdir + "/part_marginals";
  make_dirs(qsHypDir);
  make_dirs(qsPartMarginalsDir);
  const char *eg = getenv("PSINFER_GPUS"), *ec = getenv("PSINFER_CTX_PER_GPU");
  const int gpus = std::max(1, g_gpus > 0 ? g_gpus : (eg ? atoi(eg) : 1));
  const int per_gpu = std::max(1, g_ctx_per_gpu > 0 ? g_ctx_per_gpu : (ec ? atoi(ec) : 4));
  const int n = lastidx - firstidx + 1;
  if (n <= 0) return;
  if (gpus == 1 && (per_gpu == 1 || n == 1)) {  // the reference's loop, in the calling thread
    for (int imgidx = firstidx; imgidx <= lastidx; ++imgidx) find_object_image(app, imgidx, qsHypDir, qsPartMarginalsDir);
    return;
  }
  const char *dev0 = getenv("PSINFER_DEVICE");
  const int base_dev = dev0 ? atoi(dev0) : 0;
  const int num_per_gpu = (n + gpus - 1) / gpus;  // (int)ceil(n / (float)ncpu), main.cpp:181
  std::vector<std::atomic<int> > next(gpus);
  for (int g = 0; g < gpus; ++g) next[g] = firstidx + g * num_per_gpu;
  std::mutex err_mutex;
  std::string first_error;
  std::atomic<bool> failed(false);
  std::vector<std::thread> workers;
  for (int g = 0; g < gpus; ++g) {
    const int range_end = std::min(lastidx, firstidx + (g + 1) * num_per_gpu - 1);
    for (int k = 0; k < per_gpu; ++k)
      workers.emplace_back([&, g, range_end]() {
        t_device = base_dev + g;
        try {
          for (;;) {
            const int imgidx = next[g].fetch_add(1);
            if (imgidx > range_end || failed.load()) break;
            find_object_image(app, imgidx, qsHypDir, qsPartMarginalsDir);
          }
        } catch (const std::exception &e) {
          std::lock_guard<std::mutex> lock(err_mutex);
          if (first_error.empty()) first_error = e.what();
          failed = true;
        }
        CtxHolder &h = holder();  // release the worker's context on its own thread
        if (h.ctx) ps_destroy(h.ctx);
        h.ctx = nullptr;
      });
  }
  for (std::thread &t : workers) t.join();
  if (failed) fail(first_error);
  if (getenv("PSINFER_HOST_TIMING"))
    fprintf(stderr, "host phases, seconds summed over %d threads: joints %.2f  load+inflate %.2f  ingest calls %.2f  infer+wait %.2f  outputs %.2f\n",
            gpus * per_gpu, g_ns_joints / 1e9, g_ns_load / 1e9, g_ns_ingest / 1e9, g_ns_infer / 1e9, g_ns_out / 1e9);
}

}  // namespace object_detect
