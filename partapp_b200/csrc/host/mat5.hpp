// mat5.hpp -- minimal MATLAB Level-5 MAT-file reader / writer (numeric and cell arrays, zlib-compressed elements).
//
// Replaces what the reference gets from MATLAB's proprietary libmat/libmx through libMatlabIO
// (reference src/libs/libMatlabIO/matlab_io.hpp:64-146, :154-215, matlab_cell_io.hpp:27-130):
//   * arrays keep their C dims and are stored column-major ("Matlab needs fortran storage order", matlab_io.hpp:72-75),
//     so element [i][j][k] of the C array is element (i+1, j+1, k+1) of the MATLAB variable;
//   * files are written compressed (mat_open(..., "wz")), one miCOMPRESSED element per variable;
//   * readers accept single or double storage (mxIsSingle branch) -- here any numeric storage type, because MATLAB
//     itself stores integer-valued doubles in the smallest integer type that fits.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mat5 {

enum { miINT8 = 1, miUINT8 = 2, miINT16 = 3, miUINT16 = 4, miINT32 = 5, miUINT32 = 6, miSINGLE = 7, miDOUBLE = 9,
       miINT64 = 12, miUINT64 = 13, miMATRIX = 14, miCOMPRESSED = 15, miUTF8 = 16 };
enum { mxCELL = 1, mxCHAR = 4, mxDOUBLE = 6, mxSINGLE = 7, mxINT8 = 8, mxUINT8 = 9, mxINT16 = 10, mxUINT16 = 11,
       mxINT32 = 12, mxUINT32 = 13, mxINT64 = 14, mxUINT64 = 15 };

// One variable.  Numeric data is held in C order (last index fastest) as float (single class) or double (others).
struct Var {
  std::string name;
  int cls = 0;
  std::vector<size_t> dims;
  std::vector<float> f32;    // cls == mxSINGLE
  std::vector<double> f64;   // every other numeric class
  std::vector<Var> cells;    // cls == mxCELL, C order over dims
  size_t numel() const {
    size_t n = 1;
    for (size_t d : dims) n *= d;
    return n;
  }
  double at(size_t i) const { return cls == mxSINGLE ? (double)f32[i] : f64[i]; }
};

namespace detail {

inline size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

struct Cursor {
  const uint8_t *p, *end;
  void need(size_t n) const {
    if ((size_t)(end - p) < n) throw std::runtime_error("mat5: truncated file");
  }
  uint32_t u32() {
    need(4);
    uint32_t v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
};

// Reads one tag; returns type, byte count and a pointer to the data; advances past the (padded) element.
inline void read_tag(Cursor &c, uint32_t &type, uint32_t &nbytes, const uint8_t *&data) {
  uint32_t w0 = c.u32();
  if (w0 >> 16) {  // small data element: bytes in the upper half, data in the next word
    type = w0 & 0xffff;
    nbytes = w0 >> 16;
    c.need(4);
    data = c.p;
    c.p += 4;
  } else {
    type = w0;
    nbytes = c.u32();
    c.need(nbytes);
    data = c.p;
    size_t adv = type == miCOMPRESSED ? nbytes : pad8(nbytes);
    if ((size_t)(c.end - c.p) < adv) adv = c.end - c.p;
    c.p += adv;
  }
}

template <typename T>
inline void widen(const uint8_t *data, size_t n, std::vector<double> &out) {
  out.resize(n);
  for (size_t i = 0; i < n; ++i) {
    T v;
    memcpy(&v, data + i * sizeof(T), sizeof(T));
    out[i] = (double)v;
  }
}

inline void numeric_to_double(uint32_t type, const uint8_t *data, uint32_t nbytes, std::vector<double> &out) {
  switch (type) {
    case miINT8: widen<int8_t>(data, nbytes, out); break;
    case miUINT8: case miUTF8: widen<uint8_t>(data, nbytes, out); break;
    case miINT16: widen<int16_t>(data, nbytes / 2, out); break;
    case miUINT16: widen<uint16_t>(data, nbytes / 2, out); break;
    case miINT32: widen<int32_t>(data, nbytes / 4, out); break;
    case miUINT32: widen<uint32_t>(data, nbytes / 4, out); break;
    case miSINGLE: widen<float>(data, nbytes / 4, out); break;
    case miDOUBLE: widen<double>(data, nbytes / 8, out); break;
    case miINT64: widen<int64_t>(data, nbytes / 8, out); break;
    case miUINT64: widen<uint64_t>(data, nbytes / 8, out); break;
    default: throw std::runtime_error("mat5: unsupported numeric storage type");
  }
}

// column-major (file) -> C order
template <typename T>
inline void fortran_to_c(const std::vector<T> &src, const std::vector<size_t> &dims, std::vector<T> &dst) {
  const size_t nd = dims.size(), n = src.size();
  dst.resize(n);
  if (nd == 2 && n == dims[0] * dims[1]) {  // the common case (score grids, parameter matrices): a blocked 2-D transpose
    const size_t R = dims[0], C = dims[1], B = 32;
    for (size_t r0 = 0; r0 < R; r0 += B)
      for (size_t c0 = 0; c0 < C; c0 += B)
        for (size_t c = c0; c < std::min(C, c0 + B); ++c)
          for (size_t r = r0; r < std::min(R, r0 + B); ++r) dst[r * C + c] = src[c * R + r];
    return;
  }
  std::vector<size_t> cstride(nd, 1), idx(nd, 0);
  for (size_t d = nd - 1; d-- > 0;) cstride[d] = cstride[d + 1] * dims[d + 1];
  for (size_t f = 0; f < n; ++f) {  // f walks column-major: first index fastest
    size_t c = 0;
    for (size_t d = 0; d < nd; ++d) c += idx[d] * cstride[d];
    dst[c] = src[f];
    for (size_t d = 0; d < nd; ++d) {
      if (++idx[d] < dims[d]) break;
      idx[d] = 0;
    }
  }
}
template <typename T>
inline void c_to_fortran(const T *src, const std::vector<size_t> &dims, std::vector<T> &dst) {
  size_t n = 1;
  for (size_t d : dims) n *= d;
  const size_t nd = dims.size();
  dst.resize(n);
  std::vector<size_t> cstride(nd, 1), idx(nd, 0);
  for (size_t d = nd - 1; d-- > 0;) cstride[d] = cstride[d + 1] * dims[d + 1];
  for (size_t f = 0; f < n; ++f) {
    size_t c = 0;
    for (size_t d = 0; d < nd; ++d) c += idx[d] * cstride[d];
    dst[f] = src[c];
    for (size_t d = 0; d < nd; ++d) {
      if (++idx[d] < dims[d]) break;
      idx[d] = 0;
    }
  }
}

inline Var parse_matrix(const uint8_t *data, uint32_t nbytes) {
  Var v;
  Cursor c{data, data + nbytes};
  uint32_t type, nb;
  const uint8_t *d;
  read_tag(c, type, nb, d);  // array flags
  if (type != miUINT32 || nb < 8) throw std::runtime_error("mat5: bad array flags");
  uint32_t flags;
  memcpy(&flags, d, 4);
  v.cls = flags & 0xff;
  if (flags & 0x0800) throw std::runtime_error("mat5: complex arrays are not supported");
  read_tag(c, type, nb, d);  // dimensions
  if (type != miINT32) throw std::runtime_error("mat5: bad dimensions");
  for (uint32_t i = 0; i < nb / 4; ++i) {
    int32_t dim;
    memcpy(&dim, d + 4 * i, 4);
    v.dims.push_back((size_t)dim);
  }
  read_tag(c, type, nb, d);  // name
  v.name.assign((const char *)d, nb);
  const size_t n = v.numel();
  if (v.cls == mxCELL) {
    std::vector<Var> col;
    for (size_t i = 0; i < n; ++i) {
      read_tag(c, type, nb, d);
      if (type != miMATRIX) throw std::runtime_error("mat5: bad cell element");
      col.push_back(nb ? parse_matrix(d, nb) : Var());
    }
    fortran_to_c(col, v.dims, v.cells);
  } else if (v.cls >= mxDOUBLE && v.cls <= mxUINT64) {
    if (n == 0) return v;
    read_tag(c, type, nb, d);  // real part
    if (v.cls == mxSINGLE && type == miSINGLE && nb / 4 >= n) {  // single stored as single: no detour through double
      std::vector<float> colf(n);
      memcpy(colf.data(), d, n * sizeof(float));
      fortran_to_c(colf, v.dims, v.f32);
      return v;
    }
    std::vector<double> col;
    numeric_to_double(type, d, nb, col);
    if (col.size() < n) throw std::runtime_error("mat5: short numeric data");
    col.resize(n);
    if (v.cls == mxSINGLE) {
      std::vector<float> colf(col.begin(), col.end());
      fortran_to_c(colf, v.dims, v.f32);
    } else {
      fortran_to_c(col, v.dims, v.f64);
    }
  } else if (v.cls == mxCHAR) {
    if (n) {
      read_tag(c, type, nb, d);
      std::vector<double> col;
      numeric_to_double(type, d, nb, col);
      col.resize(n);
      fortran_to_c(col, v.dims, v.f64);
    }
  } else {
    throw std::runtime_error("mat5: unsupported array class");
  }
  return v;
}

inline std::vector<uint8_t> inflate_all(const uint8_t *src, size_t n) {
  std::vector<uint8_t> out(std::max<size_t>(n * 4, 1024));
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit(&zs) != Z_OK) throw std::runtime_error("mat5: inflateInit failed");
  zs.next_in = const_cast<Bytef *>(src);
  zs.avail_in = (uInt)n;
  size_t have = 0;
  for (;;) {
    zs.next_out = out.data() + have;
    zs.avail_out = (uInt)(out.size() - have);
    int rc = inflate(&zs, Z_NO_FLUSH);
    have = out.size() - zs.avail_out;
    if (rc == Z_STREAM_END) break;
    if (rc != Z_OK) {
      inflateEnd(&zs);
      throw std::runtime_error("mat5: inflate failed");
    }
    if (zs.avail_out == 0) out.resize(out.size() * 2);
  }
  inflateEnd(&zs);
  out.resize(have);
  return out;
}

inline void put_u32(std::vector<uint8_t> &b, uint32_t v) {
  uint8_t t[4];
  memcpy(t, &v, 4);
  b.insert(b.end(), t, t + 4);
}
inline void put_element(std::vector<uint8_t> &b, uint32_t type, const void *data, uint32_t nbytes) {
  if (nbytes <= 4 && nbytes > 0) {  // small data element
    put_u32(b, (nbytes << 16) | type);
    uint8_t t[4] = {0, 0, 0, 0};
    memcpy(t, data, nbytes);
    b.insert(b.end(), t, t + 4);
    return;
  }
  put_u32(b, type);
  put_u32(b, nbytes);
  const uint8_t *p = (const uint8_t *)data;
  b.insert(b.end(), p, p + nbytes);
  b.resize(b.size() + (pad8(nbytes) - nbytes), 0);
}

}  // namespace detail

// ---- reading ------------------------------------------------------------------------------------------------------
inline std::vector<Var> load(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("mat5: cannot open " + path);
  std::vector<uint8_t> buf;
  uint8_t tmp[1 << 16];
  size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  if (buf.size() < 128) throw std::runtime_error("mat5: " + path + " is not a MAT-file");
  if (!(buf[126] == 'I' && buf[127] == 'M')) throw std::runtime_error("mat5: only little-endian MAT v5 files are supported");
  std::vector<Var> vars;
  detail::Cursor c{buf.data() + 128, buf.data() + buf.size()};
  while (c.end - c.p >= 8) {
    uint32_t type, nb;
    const uint8_t *d;
    detail::read_tag(c, type, nb, d);
    if (type == miCOMPRESSED) {
      std::vector<uint8_t> raw = detail::inflate_all(d, nb);
      detail::Cursor ic{raw.data(), raw.data() + raw.size()};
      uint32_t it, inb;
      const uint8_t *id;
      detail::read_tag(ic, it, inb, id);
      if (it == miMATRIX) vars.push_back(detail::parse_matrix(id, inb));
    } else if (type == miMATRIX) {
      vars.push_back(detail::parse_matrix(d, nb));
    }
  }
  return vars;
}

inline const Var &find(const std::vector<Var> &vars, const std::string &name, const std::string &path = "") {
  for (const Var &v : vars)
    if (v.name == name) return v;
  throw std::runtime_error("mat5: variable '" + name + "' not found" + (path.empty() ? "" : " in " + path));
}

// ---- writing ------------------------------------------------------------------------------------------------------
class Writer {
 public:
  explicit Writer(const std::string &path, bool compress = true) : compress_(compress) {
    f_ = fopen(path.c_str(), "wb");
    if (!f_) throw std::runtime_error("mat5: cannot create " + path);
    char hdr[128];
    memset(hdr, ' ', 116);
    const char *txt = "MATLAB 5.0 MAT-file, written by psinfer (partapp_b200)";
    memcpy(hdr, txt, strlen(txt));
    memset(hdr + 116, 0, 8);
    hdr[124] = 0x00; hdr[125] = 0x01;  // version 0x0100, little endian
    hdr[126] = 'I'; hdr[127] = 'M';
    fwrite(hdr, 1, 128, f_);
  }
  ~Writer() { close(); }
  void close() {
    if (f_) fclose(f_);
    f_ = nullptr;
  }
  // data in C order with the given dims (1-D arrays become n x 1 like mat_save_std_vector, matlab_io.cpp:158-166)
  void put(const std::string &name, const float *data, std::vector<size_t> dims) { put_t(name, data, dims, mxSINGLE, miSINGLE); }
  void put(const std::string &name, const double *data, std::vector<size_t> dims) { put_t(name, data, dims, mxDOUBLE, miDOUBLE); }

 private:
  template <typename T>
  void put_t(const std::string &name, const T *data, std::vector<size_t> dims, uint32_t cls, uint32_t mitype) {
    if (dims.size() == 1) dims.push_back(1);
    std::vector<T> col;
    detail::c_to_fortran(data, dims, col);
    std::vector<uint8_t> m;
    uint32_t flags[2] = {cls, 0};
    detail::put_element(m, miUINT32, flags, 8);
    std::vector<int32_t> d32(dims.begin(), dims.end());
    detail::put_element(m, miINT32, d32.data(), (uint32_t)(d32.size() * 4));
    detail::put_element(m, miINT8, name.data(), (uint32_t)name.size());
    if (!col.empty()) detail::put_element(m, mitype, col.data(), (uint32_t)(col.size() * sizeof(T)));
    std::vector<uint8_t> el;
    detail::put_u32(el, miMATRIX);
    detail::put_u32(el, (uint32_t)m.size());
    el.insert(el.end(), m.begin(), m.end());
    if (compress_) {
      uLongf bound = compressBound((uLong)el.size());
      std::vector<uint8_t> z(bound);
      if (compress2(z.data(), &bound, el.data(), (uLong)el.size(), Z_DEFAULT_COMPRESSION) != Z_OK)
        throw std::runtime_error("mat5: deflate failed");
      uint32_t tag[2] = {miCOMPRESSED, (uint32_t)bound};
      fwrite(tag, 4, 2, f_);
      fwrite(z.data(), 1, bound, f_);
    } else {
      fwrite(el.data(), 1, el.size(), f_);
    }
  }
  FILE *f_ = nullptr;
  bool compress_;
};

}  // namespace mat5
