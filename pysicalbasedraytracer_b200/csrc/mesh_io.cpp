// mesh_io.cpp - host-side scene ingestion: the reference's private ".3d" mesh text format and Radiance ".hdr"
// environment maps (below).  ".3d" mesh text format (Shape/plyRead.h:17-47,
// used for dragon.3d at Main/main.cpp:341-343):
//     vertex <N> face <M>            (two header records in either order)
//     N x  "<x> <y> <z>"             every vertex is multiplied by 20 in float (plyRead.h:39)
//     M x  "<count> <i0> <i1> <i2>"  the leading count token is read and ignored (plyRead.h:42)
// Same stream extraction (operator>> on std::fstream) as the reference, so the parsed floats are identical.
// This is scene ingestion on the wire side of the hot path; no GPU involved.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <utility>
#include <string>
#include <vector>

#include "../../include/pbrgpu.h"

extern "C" {

int pbrgpu_load_3d(const char* path, float scale, float** out_verts, int32_t* out_n_verts, int32_t** out_indices, int32_t* out_n_tris) {
  if (!path || !out_verts || !out_n_verts || !out_indices || !out_n_tris) return 1;
  *out_verts = nullptr; *out_indices = nullptr; *out_n_verts = 0; *out_n_tris = 0;
  std::fstream f(path);
  if (!f.is_open()) return 2;
  int nV = -1, nT = -1;
  std::string ed;
  for (int i = 0; i < 2; i++) {
    f >> ed;
    if (ed == "vertex") f >> nV;
    else if (ed == "face") f >> nT;
  }
  if (!f || nV <= 0 || nT <= 0 || nV > (1 << 28) || nT > (1 << 28)) return 3;
  float* v = (float*)malloc(sizeof(float) * 3 * (size_t)nV);
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)nT);
  if (!v || !idx) { free(v); free(idx); return 4; }
  for (int i = 0; i < nV; i++) {
    for (int k = 0; k < 3; ++k) {
      float x;
      f >> x;
      v[3 * (size_t)i + k] = x * scale;  // Point3::operator*=(float): one float multiply per component
    }
  }
  for (int i = 0; i < nT; i++) {
    f >> ed;
    f >> idx[3 * (size_t)i] >> idx[3 * (size_t)i + 1] >> idx[3 * (size_t)i + 2];
  }
  if (!f) { free(v); free(idx); return 5; }
  for (size_t i = 0; i < 3 * (size_t)nT; ++i)
    if (idx[i] < 0 || idx[i] >= nV) { free(v); free(idx); return 6; }
  *out_verts = v; *out_indices = idx; *out_n_verts = nV; *out_n_tris = nT;
  return 0;
}

// Radiance ".hdr" (RGBE) environment maps, the format InfiniteAreaLight / SkyBoxLight read through stbi_loadf
// (Light/InfiniteAreaLight.cpp:17-18, Light/SkyBoxLight.cpp:19-22).  Written from the format description:
//   "#?RADIANCE" | "#?RGBE" \n  header lines ... "FORMAT=32-bit_rle_rgbe" ... \n \n  "-Y <h> +X <w>" \n  pixel data
//   pixel data: flat RGBE quadruples, or per scanline  02 02 <w_hi> <w_lo>  followed by the four channels, each
//   run-length coded (count > 128: repeat the next byte count-128 times; else copy `count` literal bytes).
//   value = mantissa_byte * 2^(e - 136), exactly 0 when e == 0 - the same float operations as stb's conversion
//   (one ldexp, one multiply per channel), so the texels are bit-identical to what the reference loads.
// flip_vertically mirrors stbi_set_flip_vertically_on_load (SkyBoxLight sets it; InfiniteAreaLight inherits whatever
// was set last - SURVEY quirk Q6 - so the caller chooses).
namespace {
bool read_line(const std::vector<unsigned char>& buf, size_t* pos, std::string* line) {
  line->clear();
  if (*pos >= buf.size()) return false;
  while (*pos < buf.size() && buf[*pos] != '\n') line->push_back((char)buf[(*pos)++]);
  if (*pos < buf.size()) ++*pos;
  return true;
}
inline void rgbe_to_float(const unsigned char* in, float* out) {
  if (in[3] != 0) {
    const float f1 = (float)ldexp(1.0f, (int)in[3] - (128 + 8));
    out[0] = in[0] * f1; out[1] = in[1] * f1; out[2] = in[2] * f1;
  } else {
    out[0] = out[1] = out[2] = 0.f;
  }
}
}  // namespace

int pbrgpu_load_hdr(const char* path, int flip_vertically, float** out_rgb, int32_t* out_width, int32_t* out_height) {
  if (!path || !out_rgb || !out_width || !out_height) return 1;
  *out_rgb = nullptr; *out_width = 0; *out_height = 0;
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return 2;
  std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  size_t pos = 0;
  std::string line;
  if (!read_line(buf, &pos, &line) || (line != "#?RADIANCE" && line != "#?RGBE")) return 3;
  bool valid = false;
  while (read_line(buf, &pos, &line) && !line.empty())
    if (line == "FORMAT=32-bit_rle_rgbe") valid = true;
  if (!valid) return 3;
  if (!read_line(buf, &pos, &line)) return 3;
  int w = 0, h = 0;
  if (sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0 || w > (1 << 24) / h * 64) return 3;
  float* rgb = (float*)malloc(sizeof(float) * 3 * (size_t)w * h);
  if (!rgb) return 4;
  auto fail = [&](int code) { free(rgb); return code; };
  bool flat = (w < 8 || w >= 32768);
  std::vector<unsigned char> scan((size_t)w * 4);
  for (int j = 0; j < h && !flat; ++j) {
    if (pos + 4 > buf.size()) return fail(5);
    const unsigned char c1 = buf[pos], c2 = buf[pos + 1], l1 = buf[pos + 2], l2 = buf[pos + 3];
    if (c1 != 2 || c2 != 2 || (l1 & 0x80)) {
      if (j != 0) return fail(5);
      flat = true;  // not run-length coded: the four bytes just read are the first pixel
      break;
    }
    pos += 4;
    if (((int)l1 << 8 | (int)l2) != w) return fail(5);
    for (int k = 0; k < 4; ++k) {
      int i = 0;
      while (i < w) {
        if (pos >= buf.size()) return fail(5);
        int count = buf[pos++];
        if (count > 128) {
          count -= 128;
          if (count > w - i || pos >= buf.size()) return fail(5);
          const unsigned char v = buf[pos++];
          for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = v;
        } else {
          if (count == 0 || count > w - i || pos + (size_t)count > buf.size()) return fail(5);
          for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = buf[pos++];
        }
      }
    }
    for (int i = 0; i < w; ++i) rgbe_to_float(&scan[(size_t)i * 4], rgb + 3 * ((size_t)j * w + i));
  }
  if (flat) {
    if (pos + (size_t)w * h * 4 > buf.size()) return fail(5);
    for (size_t i = 0; i < (size_t)w * h; ++i) rgbe_to_float(&buf[pos + 4 * i], rgb + 3 * i);
  }
  if (flip_vertically)
    for (int j = 0; j < h / 2; ++j)
      for (size_t i = 0; i < (size_t)w * 3; ++i) std::swap(rgb[(size_t)j * w * 3 + i], rgb[(size_t)(h - 1 - j) * w * 3 + i]);
  *out_rgb = rgb; *out_width = w; *out_height = h;
  return 0;
}

void pbrgpu_free(void* p) { free(p); }

}  // extern "C"
