// texture.cuh - ImageTexture<RGBSpectrum, Spectrum> on the device: UVMapping2D + MIPMap lookups.
//   ImageTexture::Evaluate                 Texture/ImageTexture.h:53-61
//   UVMapping2D::Map                       Texture/Texture.cpp:4-10
//   MIPMap::{Lookup(st,width), Lookup(st,dst0,dst1), triangle, EWA, Texel}   Texture/MIPMap.h:183-341
//   SurfaceInteraction::ComputeDifferentials   Core/Interaction.cpp:48-99, SolveLinearSystem2x2 Core/Transform.cpp:9-17
//   camera ray differentials               Camera/Perspective.cpp:44-80, Camera/Orthographic.cpp:27-63,
//                                          RayDifferential::ScaleDifferentials Core/Geometry.h:1395-1400
// The MIP pyramid (2x2 box filter per level, MIPMap.h:146-168) and the EWA weight table (:170-176, glibc expf) are
// built on the host; texel sizes must be powers of two (the reference Lanczos-resamples anything else).
#pragma once
#include "device_types.cuh"

namespace pbr {

constexpr int kWeightLutSize = 128;  // MIPMap::WeightLUTSize

PBR_D int tex_mod(int a, int b) {  // Mod<int> (Core/PBR.h:193-197)
  int r = a - (a / b) * b;
  return r < 0 ? r + b : r;
}

PBR_D Spec tex_texel(const TexDev& T, int level, int s, int t) {
  const int w = max(1, T.width >> level), h = max(1, T.height >> level);
  switch (T.wrap) {
    case PBRGPU_WRAP_REPEAT: s = tex_mod(s, w); t = tex_mod(t, h); break;
    case PBRGPU_WRAP_CLAMP: s = min(max(s, 0), w - 1); t = min(max(t, 0), h - 1); break;
    default:
      if (s < 0 || s >= w || t < 0 || t >= h) return Spec(0.f);
      break;
  }
  const float4 v = __ldg(T.texels + T.levelOffset[level] + (size_t)t * w + s);
  return Spec(v.x, v.y, v.z);
}

PBR_D Spec tex_triangle(const TexDev& T, int level, float s_, float t_) {
  level = min(max(level, 0), T.nLevels - 1);
  const int w = max(1, T.width >> level), h = max(1, T.height >> level);
  float s = s_ * w - 0.5f;
  float t = t_ * h - 0.5f;
  int s0 = (int)floorf(s), t0 = (int)floorf(t);
  float ds = s - s0, dt = t - t0;
  return (1 - ds) * (1 - dt) * tex_texel(T, level, s0, t0) + (1 - ds) * dt * tex_texel(T, level, s0, t0 + 1) +
         ds * (1 - dt) * tex_texel(T, level, s0 + 1, t0) + ds * dt * tex_texel(T, level, s0 + 1, t0 + 1);
}

PBR_D float tex_log2(float x) { return logf(x) * 1.442695040888963387004650940071f; }
PBR_D Spec spec_lerp(float t, Spec a, Spec b) { return (1 - t) * a + t * b; }

PBR_D Spec tex_ewa(const TexDev& T, const float* __restrict__ lut, int level, float st0, float st1, float d00, float d01, float d10, float d11) {
  if (level >= T.nLevels) return tex_texel(T, T.nLevels - 1, 0, 0);
  const int w = max(1, T.width >> level), h = max(1, T.height >> level);
  st0 = st0 * w - 0.5f;
  st1 = st1 * h - 0.5f;
  d00 *= w; d01 *= h; d10 *= w; d11 *= h;
  float A = d01 * d01 + d11 * d11 + 1;
  float B = -2 * (d00 * d01 + d10 * d11);
  float C = d00 * d00 + d10 * d10 + 1;
  float invF = 1 / (A * C - B * B * 0.25f);
  A *= invF; B *= invF; C *= invF;
  float det = -B * B + 4 * A * C;
  float invDet = 1 / det;
  float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
  int s0 = (int)ceilf(st0 - 2 * invDet * uSqrt);
  int s1 = (int)floorf(st0 + 2 * invDet * uSqrt);
  int t0 = (int)ceilf(st1 - 2 * invDet * vSqrt);
  int t1 = (int)floorf(st1 + 2 * invDet * vSqrt);
  Spec sum(0.f);
  float sumWts = 0;
  for (int it = t0; it <= t1; ++it) {
    float tt = it - st1;
    for (int is = s0; is <= s1; ++is) {
      float ss = is - st0;
      float r2 = A * ss * ss + B * ss * tt + C * tt * tt;
      if (r2 < 1) {
        int index = min((int)(r2 * kWeightLutSize), kWeightLutSize - 1);
        float weight = __ldg(lut + index);
        sum = sum + tex_texel(T, level, is, it) * weight;
        sumWts += weight;
      }
    }
  }
  return sum / sumWts;
}

// MIPMap::Lookup(st, dst0, dst1)
PBR_D Spec tex_lookup(const TexDev& T, const float* lut, float s, float t, float d00, float d01, float d10, float d11) {
  if (T.doTrilinear) {
    float width = stdmax(stdmax(fabsf(d00), fabsf(d01)), stdmax(fabsf(d10), fabsf(d11)));
    float level = T.nLevels - 1 + tex_log2(stdmax(width, (float)1e-8));
    if (level < 0) return tex_triangle(T, 0, s, t);
    else if (level >= T.nLevels - 1) return tex_texel(T, T.nLevels - 1, 0, 0);
    int iLevel = (int)floorf(level);
    float delta = level - iLevel;
    return spec_lerp(delta, tex_triangle(T, iLevel, s, t), tex_triangle(T, iLevel + 1, s, t));
  }
  if (d00 * d00 + d01 * d01 < d10 * d10 + d11 * d11) {
    float a = d00, b = d01;
    d00 = d10; d01 = d11; d10 = a; d11 = b;
  }
  float majorLength = sqrtf(d00 * d00 + d01 * d01);
  float minorLength = sqrtf(d10 * d10 + d11 * d11);
  if (minorLength * T.maxAnisotropy < majorLength && minorLength > 0) {
    float scale = majorLength / (minorLength * T.maxAnisotropy);
    d10 *= scale; d11 *= scale;
    minorLength *= scale;
  }
  if (minorLength == 0) return tex_triangle(T, 0, s, t);
  float lod = stdmax((float)0, T.nLevels - (float)1 + tex_log2(minorLength));
  int ilod = (int)floorf(lod);
  return spec_lerp(lod - ilod, tex_ewa(T, lut, ilod, s, t, d00, d01, d10, d11), tex_ewa(T, lut, ilod + 1, s, t, d00, d01, d10, d11));
}

// du/dx, dv/dx, du/dy, dv/dy of a surface hit
struct UvDiff {
  float dudx, dvdx, dudy, dvdy;
};

PBR_D bool solve2x2(float a00, float a01, float a10, float a11, float b0, float b1, float* x0, float* x1) {
  float det = a00 * a11 - a01 * a10;
  if (fabsf(det) < 1e-10f) return false;
  *x0 = (a11 * b0 - a01 * b1) / det;
  *x1 = (a00 * b1 - a10 * b0) / det;
  if (isnan(*x0) || isnan(*x1)) return false;
  return true;
}

PBR_D float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// SurfaceInteraction::ComputeDifferentials for a ray WITH differentials (rx / ry = the offset rays of the camera sample)
PBR_D UvDiff compute_differentials(V3 p, V3 n, V3 dpdu, V3 dpdv, V3 rxO, V3 rxD, V3 ryO, V3 ryD) {
  UvDiff r{0.f, 0.f, 0.f, 0.f};
  float d = dot(n, p);
  float tx = -(dot(n, rxO) - d) / dot(n, rxD);
  if (isinf(tx) || isnan(tx)) return r;
  V3 px = rxO + tx * rxD;
  float ty = -(dot(n, ryO) - d) / dot(n, ryD);
  if (isinf(ty) || isnan(ty)) return r;
  V3 py = ryO + ty * ryD;
  int d0, d1;
  if (fabsf(n.x) > fabsf(n.y) && fabsf(n.x) > fabsf(n.z)) { d0 = 1; d1 = 2; }
  else if (fabsf(n.y) > fabsf(n.z)) { d0 = 0; d1 = 2; }
  else { d0 = 0; d1 = 1; }
  const float a00 = comp(dpdu, d0), a01 = comp(dpdv, d0), a10 = comp(dpdu, d1), a11 = comp(dpdv, d1);
  const float bx0 = comp(px, d0) - comp(p, d0), bx1 = comp(px, d1) - comp(p, d1);
  const float by0 = comp(py, d0) - comp(p, d0), by1 = comp(py, d1) - comp(p, d1);
  if (!solve2x2(a00, a01, a10, a11, bx0, bx1, &r.dudx, &r.dvdx)) r.dudx = r.dvdx = 0;
  if (!solve2x2(a00, a01, a10, a11, by0, by1, &r.dudy, &r.dvdy)) r.dudy = r.dvdy = 0;
  return r;
}

// ImageTexture::Evaluate with a UVMapping2D
PBR_D Spec tex_evaluate(const TexDev& T, const float* lut, float u, float v, UvDiff dd) {
  const float d00 = T.su * dd.dudx, d01 = T.sv * dd.dvdx;  // dstdx
  const float d10 = T.su * dd.dudy, d11 = T.sv * dd.dvdy;  // dstdy
  return tex_lookup(T, lut, T.su * u + T.du, T.sv * v + T.dv, d00, d01, d10, d11);
}

}  // namespace pbr
