// mesh_io.cpp - host-side loader for the reference's private ".3d" mesh text format (Shape/plyRead.h:17-47,
// used for dragon.3d at Main/main.cpp:341-343):
//     vertex <N> face <M>            (two header records in either order)
//     N x  "<x> <y> <z>"             every vertex is multiplied by 20 in float (plyRead.h:39)
//     M x  "<count> <i0> <i1> <i2>"  the leading count token is read and ignored (plyRead.h:42)
// Same stream extraction (operator>> on std::fstream) as the reference, so the parsed floats are identical.
// This is scene ingestion on the wire side of the hot path; no GPU involved.
#include <cstdlib>
#include <cstring>
#include <fstream>
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

void pbrgpu_free(void* p) { free(p); }

}  // extern "C"
