// vecmath.cuh - fp32 vector / spectrum helpers whose operation ORDER follows the reference's
// Core/Geometry.h and Core/Spectrum.h so that results are bit-identical to the CPU reference built
// with -ffp-contract=off.  This translation unit is compiled with -fmad=false: every * and + below
// is an individually rounded IEEE operation; FMAs appear only where written explicitly (fmaf).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PBR_HD __host__ __device__ __forceinline__
#define PBR_D __device__ __forceinline__

namespace pbr {

// Core/PBR.h:12-18 (float constants)
constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInv2Pi = 0.15915494309189533577f;
constexpr float kInv4Pi = 0.07957747154594766788f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kOneMinusEpsilon = 0.99999994f;  // Sampler/RNG.h:13
constexpr float kShadowEpsilon = 0.0001f;        // Core/PBR.h:45
constexpr float kInfinity = __builtin_huge_valf();

// gamma(n) = (n*eps)/(1-n*eps) evaluated in DOUBLE and narrowed (Core/PBR.h:21-24)
PBR_HD constexpr float gamma_f(int n) {
  return (float)((n * (1.1920928955078125e-07 * 0.5)) / (1 - n * (1.1920928955078125e-07 * 0.5)));
}

struct V3 {
  float x, y, z;
  PBR_HD V3() : x(0), y(0), z(0) {}
  PBR_HD V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  PBR_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
PBR_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
PBR_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
PBR_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
PBR_HD V3 operator*(V3 a, float s) { return V3(s * a.x, s * a.y, s * a.z); }
PBR_HD V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
PBR_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
// Vector3::operator/(f): multiply by (float)1/f  (Geometry.h:223-228)
PBR_HD V3 div_by(V3 a, float f) {
  float inv = 1.0f / f;
  return V3(a.x * inv, a.y * inv, a.z * inv);
}
PBR_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PBR_HD float absdot(V3 a, V3 b) { return fabsf(dot(a, b)); }
PBR_HD float length_sq(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PBR_HD float length(V3 a) { return sqrtf(length_sq(a)); }
PBR_HD V3 normalize(V3 a) { return div_by(a, length(a)); }
PBR_HD V3 vabs(V3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// Cross(): products and differences in double, narrowed to float (Geometry.h:706-734)
PBR_HD V3 cross(V3 a, V3 b) {
  double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
  return V3((float)((ay * bz) - (az * by)), (float)((az * bx) - (ax * bz)), (float)((ax * by) - (ay * bx)));
}
PBR_HD float dist_sq(V3 a, V3 b) { return length_sq(a - b); }
PBR_HD V3 faceforward(V3 n, V3 v) { return (dot(n, v) < 0.f) ? -n : n; }
// std::min / std::max semantics (second argument wins only on strict compare; NaN-transparent like the STL)
PBR_HD float stdmin(float a, float b) { return (b < a) ? b : a; }
PBR_HD float stdmax(float a, float b) { return (a < b) ? b : a; }
PBR_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
PBR_HD float max_component(V3 a) { return stdmax(a.x, stdmax(a.y, a.z)); }
PBR_HD int max_dimension(V3 v) { return (v.x > v.y) ? ((v.x > v.z) ? 0 : 2) : ((v.y > v.z) ? 1 : 2); }
// CoordinateSystem (Geometry.h:779-787)
PBR_HD void coordinate_system(V3 v1, V3* v2, V3* v3) {
  if (fabsf(v1.x) > fabsf(v1.y))
    *v2 = div_by(V3(-v1.z, 0, v1.x), sqrtf(v1.x * v1.x + v1.z * v1.z));
  else
    *v2 = div_by(V3(0, v1.z, -v1.y), sqrtf(v1.y * v1.y + v1.z * v1.z));
  *v3 = cross(v1, *v2);
}

// NextFloatUp / NextFloatDown (Core/PBR.h:142-165)
PBR_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PBR_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
PBR_HD float next_float_up(float v) {
  if (isinf(v) && v > 0.f) return v;
  if (v == -0.f) v = 0.f;
  uint32_t ui = f2u(v);
  if (v >= 0) ++ui; else --ui;
  return u2f(ui);
}
PBR_HD float next_float_down(float v) {
  if (isinf(v) && v < 0.f) return v;
  if (v == 0.f) v = -0.f;
  uint32_t ui = f2u(v);
  if (v > 0) --ui; else ++ui;
  return u2f(ui);
}
// OffsetRayOrigin (Geometry.h:1470-1484)
PBR_HD V3 offset_ray_origin(V3 p, V3 pError, V3 n, V3 w) {
  float d = dot(vabs(n), pError);
  V3 offset = d * n;
  if (dot(w, n) < 0) offset = -offset;
  V3 po = p + offset;
  if (offset.x > 0) po.x = next_float_up(po.x); else if (offset.x < 0) po.x = next_float_down(po.x);
  if (offset.y > 0) po.y = next_float_up(po.y); else if (offset.y < 0) po.y = next_float_down(po.y);
  if (offset.z > 0) po.z = next_float_up(po.z); else if (offset.z < 0) po.z = next_float_down(po.z);
  return po;
}

// RGBSpectrum (Core/Spectrum.h): 3 floats with per-channel arithmetic; `/ float` is a true divide.
struct Spec {
  float r, g, b;
  PBR_HD Spec() : r(0), g(0), b(0) {}
  PBR_HD explicit Spec(float v) : r(v), g(v), b(v) {}
  PBR_HD Spec(float r_, float g_, float b_) : r(r_), g(g_), b(b_) {}
  PBR_HD bool is_black() const { return !(r != 0.f) && !(g != 0.f) && !(b != 0.f); }
  PBR_HD float max_value() const { return stdmax(stdmax(r, g), b); }
  PBR_HD float y() const { return 0.212671f * r + 0.715160f * g + 0.072169f * b; }
};
PBR_HD Spec operator+(Spec a, Spec b) { return Spec(a.r + b.r, a.g + b.g, a.b + b.b); }
PBR_HD Spec operator-(Spec a, Spec b) { return Spec(a.r - b.r, a.g - b.g, a.b - b.b); }
PBR_HD Spec operator*(Spec a, Spec b) { return Spec(a.r * b.r, a.g * b.g, a.b * b.b); }
PBR_HD Spec operator*(Spec a, float s) { return Spec(a.r * s, a.g * s, a.b * s); }
PBR_HD Spec operator*(float s, Spec a) { return Spec(a.r * s, a.g * s, a.b * s); }
// a / s, bit-identical to IEEE division.  On the device `0 / s` sends the division through its slow path
// (FCHK flags a zero dividend); spectra have exactly-zero channels all the time (a green Matte is (0,1,0)), and
// ncu showed that slow path costing ~7 % of the shading kernel.  For a = +-0 and s a normal finite number,
// a * s gives the same signed zero as a / s, so those cases skip the division.
PBR_HD float div_f(float a, float s) {
#ifdef __CUDA_ARCH__
  const float as = fabsf(s);
  if (a == 0.f && as >= 1.17549435e-38f && as <= 3.40282347e+38f) return a * s;
#endif
  return a / s;
}
PBR_HD Spec operator/(Spec a, float s) { return Spec(div_f(a.r, s), div_f(a.g, s), div_f(a.b, s)); }
PBR_HD Spec operator/(Spec a, Spec b) { return Spec(div_f(a.r, b.r), div_f(a.g, b.g), div_f(a.b, b.b)); }
PBR_HD Spec spec_sqrt(Spec a) { return Spec(sqrtf(a.r), sqrtf(a.g), sqrtf(a.b)); }
PBR_HD Spec spec_clamp0(Spec a) {  // Spectrum::Clamp(0, Infinity)
  return Spec(clampf(a.r, 0.f, kInfinity), clampf(a.g, 0.f, kInfinity), clampf(a.b, 0.f, kInfinity));
}

}  // namespace pbr
