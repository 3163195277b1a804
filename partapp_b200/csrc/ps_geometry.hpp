// ps_geometry.hpp -- host-side, double-precision parameter math of one message.
//
// The reference evaluates all geometry of computeRotJointMarginal on the CPU in double precision
// (libBoostMath/homogeneous_coord.{h,cpp}, boost_math.cpp, libPartApp/partapp_aux.hpp) and only the
// grid sweeps are hot.  We keep that split: everything here runs once per (joint, direction, scale)
// when ps_set_joints is called and is turned into small device tables (integer shift tables, fp32
// taps, 2x3 affine rows); the kernels never redo this math in a different precision.
//
// Expression order follows the reference so the tables are identical to what its loops would see:
// 3x3 products are "t = 0; t += a(i,k)*b(k,j)" (uBLAS prod), map_point is
// M00*x + M01*y + M02 (homogeneous_coord.h:79-80), round is floor(v+0.5) (boost_math.hpp:35).
// This translation unit is compiled by the host compiler (nvcc -Xcompiler -ffp-contract=off),
// so no FMA contraction happens here.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace psg {

constexpr float kLogZero = -1e6f;  // LOG_ZERO, libBoostMath/boost_math.h:23

struct M3 {
  double a[3][3];
  static M3 zero() {
    M3 r;
    for (auto &row : r.a)
      for (double &v : row) v = 0.0;
    return r;
  }
  static M3 eye() {
    M3 r = zero();
    r.a[0][0] = r.a[1][1] = r.a[2][2] = 1.0;
    return r;
  }
};

inline M3 mul(const M3 &x, const M3 &y) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = 0.0;
      for (int k = 0; k < 3; ++k) t += x.a[i][k] * y.a[k][j];
      r.a[i][j] = t;
    }
  return r;
}

inline void mul(const M3 &x, const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) {
    double t = 0.0;
    for (int k = 0; k < 3; ++k) t += x.a[i][k] * v[k];
    out[i] = t;
  }
}

// hc::get_rotation_matrix / get_scaling_matrix / get_translation_matrix / get_homogeneous_matrix
// (homogeneous_coord.cpp:81-93, :71-78, :95-103, :38-47)
inline M3 rotation(double rad) {
  M3 r = M3::zero();
  double ca = std::cos(rad), sa = std::sin(rad);
  r.a[0][0] = ca; r.a[0][1] = -sa; r.a[1][0] = sa; r.a[1][1] = ca; r.a[2][2] = 1;
  return r;
}
inline M3 scaling(double s) {
  M3 r = M3::eye();
  r.a[0][0] = s; r.a[1][1] = s;
  return r;
}
inline M3 translation(double dx, double dy) {
  M3 r = M3::eye();
  r.a[0][2] = dx; r.a[1][2] = dy;
  return r;
}
inline M3 homogeneous(const double R[2][2], double dx, double dy) {
  M3 r = M3::zero();
  r.a[0][0] = R[0][0]; r.a[0][1] = R[0][1]; r.a[1][0] = R[1][0]; r.a[1][1] = R[1][1];
  r.a[0][2] = dx; r.a[1][2] = dy; r.a[2][2] = 1;
  return r;
}

// hc::inverse (homogeneous_coord.cpp:49-69): analytic affine inverse
inline M3 inverse(const M3 &T) {
  M3 inv;
  double D = T.a[0][0] * T.a[1][1] - T.a[1][0] * T.a[0][1];
  inv.a[0][0] = T.a[1][1] / D;
  inv.a[0][1] = -T.a[0][1] / D;
  inv.a[1][0] = -T.a[1][0] / D;
  inv.a[1][1] = T.a[0][0] / D;
  for (int i = 0; i < 2; ++i) {
    double t = 0.0;
    for (int k = 0; k < 2; ++k) t += inv.a[i][k] * T.a[k][2];
    inv.a[i][2] = -t;
  }
  inv.a[2][0] = 0; inv.a[2][1] = 0; inv.a[2][2] = 1;
  return inv;
}

// hc::map_point (homogeneous_coord.h:72-81)
inline void map_point(const M3 &M, double x, double y, double &ox, double &oy) {
  ox = M.a[0][0] * x + M.a[0][1] * y + M.a[0][2];
  oy = M.a[1][0] * x + M.a[1][1] * y + M.a[1][2];
}

// hc::get_transformed_bbox (homogeneous_coord.cpp:139-156)
inline void transformed_bbox(const M3 &T21, int w, int h, double &minx, double &miny, double &maxx,
                             double &maxy) {
  const double pts[4][3] = {{0, 0, 1}, {double(w - 1), 0, 1}, {0, double(h - 1), 1},
                            {double(w - 1), double(h - 1), 1}};
  for (int i = 0; i < 4; ++i) {
    double c[3];
    mul(T21, pts[i], c);
    if (i == 0) {
      minx = maxx = c[0];
      miny = maxy = c[1];
    } else {
      minx = std::fmin(minx, c[0]); maxx = std::fmax(maxx, c[0]);
      miny = std::fmin(miny, c[1]); maxy = std::fmax(maxy, c[1]);
    }
  }
}

// boost_math::eig2d (boost_math.cpp:40-99): smallest eigenvalue first, V = [v1 v2], v2 = (-v1y, v1x)
inline void eig2d(const double M[2][2], double V[2][2], double E[2]) {
  double m11 = M[0][0], m12 = M[0][1], m22 = M[1][1];
  double e1, e2, v11, v21;
  if (m12 != 0) {
    double sqrtD = std::sqrt((m11 - m22) * (m11 - m22) + 4 * m12 * m12);
    e1 = 0.5 * (m11 + m22 - sqrtD);
    e2 = 0.5 * (m11 + m22 + sqrtD);
    v11 = 0.5 * (m11 - m22 - sqrtD) / m12;
    v21 = 1;
  } else if (m11 < m22) {
    e1 = m11; e2 = m22; v11 = 1; v21 = 0;
  } else {
    e1 = m22; e2 = m11; v11 = 0; v21 = 1;
  }
  double nrm = std::sqrt(v11 * v11 + v21 * v21);
  v11 /= nrm;
  v21 /= nrm;
  V[0][0] = v11; V[1][0] = v21; V[0][1] = -v21; V[1][1] = v11;
  E[0] = e1; E[1] = e2;
}

// boost_math::get_gaussian_filter (boost_math.cpp:104-117), unnormalised (peak tap = 1)
inline std::vector<double> gaussian_taps(double sigma) {
  int k = (int)std::floor(3 * sigma + 0.5);
  std::vector<double> f(2 * k + 1);
  f[k] = 1.0;
  for (int i = 1; i <= k; ++i) {
    f[k + i] = std::exp(-i * i / (2 * sigma * sigma));
    f[k - i] = f[k + i];
  }
  return f;
}

inline std::vector<float> to_float(const std::vector<double> &d) {
  std::vector<float> f(d.size());
  for (size_t i = 0; i < d.size(); ++i) f[i] = (float)d[i];
  return f;
}

// partapp_aux.hpp:45-58 / :25-43
inline double value_from_index(double lo, double hi, double steps, int idx) {
  if (lo == hi) return lo;
  double step = (hi - lo) / steps;
  return lo + step * (0.5 + idx);
}
inline int index_from_value(double lo, double hi, double steps, double val) {
  if (lo == hi) return 0;
  if (!(val >= lo && val < hi)) return -1;
  double step = (hi - lo) / steps;
  return (int)(unsigned)std::floor((val - lo) / step);
}

struct Grid {
  int R, H, W;
  float min_rot, max_rot;  // degrees, ExpParam floats
};

inline double rot_deg(const Grid &g, int r) { return value_from_index(g.min_rot, g.max_rot, g.R, r); }

// Everything one computeRotJointMarginal call (findrot.cpp:292-456) derives from its scalar arguments.
struct MessagePlan {
  // rotation axis (:319-326, :377-420)
  int rot_shift = 0;            // rot_mean_idx: slice r is written to r + rot_shift, no wrap
  int rot_mode = 0;             // 0: rot_sigma == 0 (copy), 1: circular filter, 2: rot_sigma < 0 (all zero)
  std::vector<float> rot_taps;  // tail-clipped (:385-390), fp32

  // nearest-neighbour translations (:345-359 in, :438-448 out): source index per destination index, -1 = out of bounds
  std::vector<int> xin, yin;    // [R][W], [R][H]
  std::vector<int> xout, yout;
  // When every table row is a pure shift (src = dst + d, out of range -> -1) the kernels use the per-rotation
  // shifts instead of table look-ups; a row that is not (a rounding tie of x3 - t at .5) keeps the tables.
  bool in_pure = false, out_pure = false;
  std::vector<int> in_shift, out_shift;  // [R][2] = (dx, dy)

  // spatial filter (:423-429 -> multi_array_filter.hpp:375-388)
  bool diag = true;
  std::vector<float> fx, fy;    // taps along x / y of the (possibly rotated) frame
  // general covariance only (multi_array_filter.hpp:335-369):
  int EH = 0, EW = 0;           // size of the eigen-frame grid, transform.hpp:298-299
  double T31[6];                // image -> eigen-frame (TM_DIRECT forward scatter), rows 0,1 of the 3x3
  double T13[6];                // eigen-frame -> image (TM_BILINEAR gather into the eigen-frame)
  double T34[6];                // image -> eigen-frame coordinates for the bilinear read-back
  // Work lists for the tiled Gaussian passes.  The eigen-frame grid is the bounding box of the ROTATED image, so on
  // average ~40 % of its cells are never read by the bilinear read-back (it samples only inside the rotated image
  // rectangle, dilated by 2 cells for the 2x2 taps).  Entries are (first row, strip | groups << 12): a tile of up
  // to 64 rows (groups x 8) in a 64-cell strip; see plan_message for how they are cut.  Skipping the rest changes no
  // value that is ever used.
  std::vector<int> ytiles, xtiles;
  long long ycells = 0, xcells = 0;  // grid cells per slice the two lists cover
  // Work of the fused x+y Gaussian kernel (k_gauss_xy): one WALK per 64-column strip of the eigen-frame grid = the
  // interval of rows some reader needs, (strip, first row, 8-row groups, offset into xmasks), longest walk first.
  // The kernel walks down the strip in 64-row blocks; the x-filtered rows live in a shared-memory ring, so block i of
  // the x pass (rows first_row - halo + 64 i ...) runs `lag` blocks ahead of the y pass.  xmasks holds, per x block,
  // which of its eight 8-column groups some y output of the rectangle can reach (bit g = columns 8g .. 8g+7).
  // walks_all / xmasks_all: the same without skipping (every strip top to bottom) for A/B runs.
  std::vector<int> walks, walks_all;
  std::vector<unsigned char> xmasks, xmasks_all;
  int halo = 0;    // ny rounded up to 8: the x pass starts this many rows above a walk's first row
  int lag = 0;     // K: y block b needs x blocks b .. b + K;  ring = 64 (K + 1) rows
  long long fcells_x = 0, fcells_y = 0;  // cells per slice the fused kernel filters along x / along y
  std::string error;            // non-empty: the reference would have hit an assert
};

inline void rows01(const M3 &m, double out[6]) {
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) out[i * 3 + j] = m.a[i][j];
}

// Source-index tables of transform_grid_fixed_size(..., Trans(tx,ty), TM_NEAREST)
// (transform.hpp:194-216): src = floor(T13 * dst + 0.5) with T13 = inverse(T21) * I.
inline void nearest_tables(const M3 &T21, int W, int H, int *xt, int *yt) {
  M3 T13 = mul(inverse(T21), M3::eye());
  for (int x3 = 0; x3 < W; ++x3) {
    double x1, y1;
    map_point(T13, (double)x3, 0.0, x1, y1);
    int ix = (int)std::floor(x1 + 0.5);
    xt[x3] = (ix >= 0 && ix < W) ? ix : -1;
  }
  for (int y3 = 0; y3 < H; ++y3) {
    double x1, y1;
    map_point(T13, 0.0, (double)y3, x1, y1);
    int iy = (int)std::floor(y1 + 0.5);
    yt[y3] = (iy >= 0 && iy < H) ? iy : -1;
  }
}

// d such that t[i] == (0 <= i + d < n ? i + d : -1) for all i, if one exists.
inline bool pure_shift(const int *t, int n, int &d) {
  int found = 0;
  d = 2 * n;  // everything out of range
  for (int i = 0; i < n; ++i)
    if (t[i] >= 0) {
      d = t[i] - i;
      found = 1;
      break;
    }
  if (!found) return true;
  for (int i = 0; i < n; ++i) {
    int e = i + d;
    if (t[i] != ((e >= 0 && e < n) ? e : -1)) return false;
  }
  return true;
}

// pos_offset != null: the legacy POS_GAUSSIAN message (computePosJointMarginal, objectdetect_findpos.cpp:64-89):
// gaussFilter2dOffset is entered directly -- the eigen-frame route even for a diagonal covariance -- and the offset
// (scaled like the covariance, :72-73) rides on the back transform, T42 = [V | -offset] (multi_array_filter.hpp:362).
inline MessagePlan plan_message(const Grid &g, const double off_in[2], const double off_out[2],
                                const double C[4], double rot_mean, double rot_sigma, double scale,
                                const double *pos_offset = nullptr) {
  MessagePlan p;
  const int R = g.R, H = g.H, W = g.W;

  // findrot.cpp:319-326 -- the range and the division are float arithmetic (float ExpParam fields, uint32 count)
  double rot_step = (g.max_rot - g.min_rot) / (float)g.R;
  rot_step *= M_PI / 180.0;
  if (!(rot_step > 0)) {
    p.error = "rot_step_size must be > 0 (findrot.cpp:321)";
    return p;
  }
  double rot_sigma_idx = rot_sigma / rot_step;
  p.rot_shift = (int)std::floor(-rot_mean / rot_step + 0.5);

  if (rot_sigma > 0) {
    p.rot_mode = 1;
    std::vector<double> f = gaussian_taps(rot_sigma_idx);
    int first = 0, len = (int)f.size();
    if (len >= R) {  // clip kernel tails, :385-390
      int c = len / 2;
      len = (R % 2 == 1) ? R - 2 : R - 1;
      first = c - len / 2;
    }
    if (len < 1 || len > 1000) {
      p.error = "rotation filter length out of range (findrot.cpp:402)";
      return p;
    }
    p.rot_taps.resize(len);
    for (int i = 0; i < len; ++i) p.rot_taps[i] = (float)f[first + i];
  } else if (rot_sigma == 0) {
    p.rot_mode = 0;
  } else {
    p.rot_mode = 2;
  }

  // per-rotation translations
  p.xin.resize((size_t)R * W); p.yin.resize((size_t)R * H);
  p.xout.resize((size_t)R * W); p.yout.resize((size_t)R * H);
  const double vin[3] = {off_in[0], off_in[1], 0}, vout[3] = {off_out[0], off_out[1], 0};
  for (int r = 0; r < R; ++r) {
    float alpha = rot_deg(g, r) * M_PI / 180.0;  // narrowed to float, :348 / :439
    M3 Tg = mul(rotation(alpha), scaling(scale));
    double t[3], u[3];
    mul(Tg, vin, t);
    mul(Tg, vout, u);
    nearest_tables(translation(t[0], t[1]), W, H, &p.xin[(size_t)r * W], &p.yin[(size_t)r * H]);
    nearest_tables(translation(-u[0], -u[1]), W, H, &p.xout[(size_t)r * W], &p.yout[(size_t)r * H]);
  }

  p.in_shift.resize(2 * R); p.out_shift.resize(2 * R);
  p.in_pure = p.out_pure = true;
  for (int r = 0; r < R; ++r) {
    p.in_pure = p.in_pure && pure_shift(&p.xin[(size_t)r * W], W, p.in_shift[2 * r]) &&
                pure_shift(&p.yin[(size_t)r * H], H, p.in_shift[2 * r + 1]);
    p.out_pure = p.out_pure && pure_shift(&p.xout[(size_t)r * W], W, p.out_shift[2 * r]) &&
                 pure_shift(&p.yout[(size_t)r * H], H, p.out_shift[2 * r + 1]);
  }

  // spatial covariance, :423 scaleC = square(scale)*C
  double s2 = scale * scale;
  double Cs[2][2] = {{s2 * C[0], s2 * C[1]}, {s2 * C[2], s2 * C[3]}};
  p.diag = (Cs[0][1] == 0 && Cs[1][0] == 0) && !pos_offset;
  double var_x, var_y;
  if (p.diag) {
    var_x = Cs[0][0];
    var_y = Cs[1][1];
  } else {
    double V[2][2], E[2];
    eig2d(Cs, V, E);
    var_x = E[0];
    var_y = E[1];
    double Vt[2][2] = {{V[0][0], V[1][0]}, {V[0][1], V[1][1]}};
    M3 T21 = homogeneous(Vt, 0, 0);
    double minx, miny, maxx, maxy;
    transformed_bbox(T21, W, H, minx, miny, maxx, maxy);
    p.EW = (int)std::ceil(maxx - minx);
    p.EH = (int)std::ceil(maxy - miny);
    if (p.EW < 1 || p.EH < 1) {
      p.error = "degenerate eigen-frame grid";
      return p;
    }
    M3 T23 = translation(minx, miny), T32 = translation(-minx, -miny);
    rows01(mul(T32, T21), p.T31);
    rows01(mul(inverse(T21), T23), p.T13);
    const double bo0 = pos_offset ? pos_offset[0] * scale : 0.0, bo1 = pos_offset ? pos_offset[1] * scale : 0.0;
    M3 T43 = mul(homogeneous(V, -bo0, -bo1), T23);
    rows01(mul(inverse(T43), M3::eye()), p.T34);
  }
  if (!(var_x > 0 && var_y > 0)) {
    p.error = "covariance must be positive definite (multi_array_filter.hpp:219)";
    return p;
  }
  std::vector<double> fx = gaussian_taps(std::sqrt(var_x)), fy = gaussian_taps(std::sqrt(var_y));
  if (fx.size() >= 1000 || fy.size() >= 1000) {
    p.error = "spatial filter longer than F_SIZE (multi_array_filter.hpp:238-244)";
    return p;
  }
  p.fx = to_float(fx);
  p.fy = to_float(fy);
  if (!p.diag) {
    // rotated image rectangle in eigen-frame coordinates: centre and half-axes from the images of the pixel-centre corners
    auto map34 = [&](double x, double y, double &ox, double &oy) {
      ox = p.T34[0] * x + p.T34[1] * y + p.T34[2];
      oy = p.T34[3] * x + p.T34[4] * y + p.T34[5];
    };
    double ax, ay, bx, by, cx, cy;
    map34(0, 0, ax, ay);
    map34(W - 1, 0, bx, by);
    map34(0, H - 1, cx, cy);
    const double ux = (bx - ax) * 0.5, uy = (by - ay) * 0.5, vx = (cx - ax) * 0.5, vy = (cy - ay) * 0.5;  // half-edges
    const double mx = ax + ux + vx, my = ay + uy + vy;                                                  // centre
    const double lu = std::sqrt(ux * ux + uy * uy), lv = std::sqrt(vx * vx + vy * vy);
    const double eux = lu > 0 ? ux / lu : 1, euy = lu > 0 ? uy / lu : 0, evx = lv > 0 ? vx / lv : 0, evy = lv > 0 ? vy / lv : 1;
    const double margin = 2.5;
    auto rect_hits = [&](double x0, double y0, double x1, double y1) {  // separating-axis test, rectangle vs oriented box
      const double rcx = 0.5 * (x0 + x1), rcy = 0.5 * (y0 + y1), rhx = 0.5 * (x1 - x0), rhy = 0.5 * (y1 - y0);
      const double dx = rcx - mx, dy = rcy - my;
      const double hu = lu + margin, hv = lv + margin;
      // axes of the rectangle
      if (std::fabs(dx) > rhx + hu * std::fabs(eux) + hv * std::fabs(evx)) return false;
      if (std::fabs(dy) > rhy + hu * std::fabs(euy) + hv * std::fabs(evy)) return false;
      // axes of the oriented box
      if (std::fabs(dx * eux + dy * euy) > hu + rhx * std::fabs(eux) + rhy * std::fabs(euy)) return false;
      if (std::fabs(dx * evx + dy * evy) > hv + rhx * std::fabs(evx) + rhy * std::fabs(evy)) return false;
      return true;
    };
    // Work lists of the TMA column kernel: entries (row0, strip | groups << 12).  A strip is 64 cells of the axis the
    // filter does not run along; within a strip the cells some reader needs form one interval of the filter axis,
    // found by testing 8-row groups (one warp's share of a tile) against the rectangle.  The interval is cut into
    // tiles of 64 rows starting at its first group, so only the last tile of a strip is partial (groups < 8: the
    // remaining warps skip it).  y pass: strips walk x, rows walk y, readers = the read-back.  x pass (transposed
    // grid): strips walk y, rows walk x, readers = the y pass, which reaches ny rows beyond the rectangle -- a cell
    // (ex, ey) is needed iff some cell (ex, ey') of the rectangle has |ey' - ey| <= ny.  Cells outside the lists are
    // never computed and never feed a cell the read-back touches.
    const int TS = 64, SG = 8;
    const int ny = ((int)p.fy.size() - 1) / 2;
    auto build = [&](bool xpass, std::vector<int> &list, long long &cells) {
      const int strips = ((xpass ? p.EH : p.EW) + TS - 1) / TS;
      const int groups = ((xpass ? p.EW : p.EH) + SG - 1) / SG;
      cells = 0;
      if (strips > 0xfff) return;  // the strip index has 12 bits: no list, the kernel walks every tile
      for (int st = 0; st < strips; ++st) {
        int g0 = -1, g1 = -1;
        for (int g = 0; g < groups; ++g) {
          const bool hit = xpass ? rect_hits(g * SG - 0.5, st * TS - ny - 0.5, g * SG + SG - 0.5, st * TS + TS - 0.5 + ny)
                                 : rect_hits(st * TS - 0.5, g * SG - 0.5, st * TS + TS - 0.5, g * SG + SG - 0.5);
          if (hit) {
            if (g0 < 0) g0 = g;
            g1 = g;
          }
        }
        if (g0 < 0) continue;
        for (int g = g0; g <= g1; g += TS / SG) {
          const int ng = std::min(TS / SG, g1 - g + 1);
          list.push_back(g * SG);
          list.push_back(st | (ng << 12));
          cells += (long long)std::min(ng * SG, (xpass ? p.EW : p.EH) - g * SG) * std::min(TS, (xpass ? p.EH : p.EW) - st * TS);
        }
      }
    };
    build(false, p.ytiles, p.ycells);
    build(true, p.xtiles, p.xcells);
    // walks of the fused kernel: the y list's intervals, one entry per strip
    p.halo = (ny + 7) & ~7;
    p.lag = (ny + p.halo + 63) / 64;
    if (p.lag < 1) p.lag = 1;
    {
      const int strips = (p.EW + TS - 1) / TS, groups = (p.EH + SG - 1) / SG;
      struct W { int st, g0, ng; };
      std::vector<W> ws, wa;
      for (int st = 0; st < strips; ++st) {
        int g0 = -1, g1 = -1;
        for (int g = 0; g < groups; ++g)
          if (rect_hits(st * TS - 0.5, g * SG - 0.5, st * TS + TS - 0.5, g * SG + SG - 0.5)) {
            if (g0 < 0) g0 = g;
            g1 = g;
          }
        if (g0 >= 0) ws.push_back({st, g0, g1 - g0 + 1});
        wa.push_back({st, 0, groups});
      }
      auto emit = [&](std::vector<W> &v, bool all, std::vector<int> &walks, std::vector<unsigned char> &masks, bool count) {
        std::stable_sort(v.begin(), v.end(), [](const W &a, const W &b) { return a.ng > b.ng; });
        for (const W &w : v) {
          const int nb = (w.ng + 7) / 8, nxb = nb + p.lag;
          walks.push_back(w.st);
          walks.push_back(w.g0 * SG);
          walks.push_back(w.ng);
          walks.push_back((int)masks.size());
          const int yor = w.g0 * SG - p.halo;
          for (int i = 0; i < nxb; ++i) {
            unsigned m = 0;
            for (int g = 0; g < 8; ++g) {
              const int ex0 = w.st * TS + g * SG;
              if (ex0 >= p.EW) break;
              if (all || rect_hits(ex0 - 0.5, yor + 64 * i - ny - 0.5, ex0 + SG - 0.5, yor + 64 * i + 64 - 0.5 + ny)) m |= 1u << g;
            }
            masks.push_back((unsigned char)m);
            if (count) {  // cells inside the grid that this box filters along x
              const long long rows_in = std::max(0, std::min(p.EH, yor + 64 * i + 64) - std::max(0, yor + 64 * i));
              for (int g = 0; g < 8; ++g)
                if ((m >> g) & 1) p.fcells_x += rows_in * std::max(0, std::min(SG, p.EW - (w.st * TS + g * SG)));
            }
          }
          if (count) p.fcells_y += (long long)std::min(w.ng * SG, p.EH - w.g0 * SG) * std::min(TS, p.EW - w.st * TS);
        }
      };
      emit(ws, false, p.walks, p.xmasks, true);
      emit(wa, true, p.walks_all, p.xmasks_all, false);
    }
  }
  return p;
}

}  // namespace psg
