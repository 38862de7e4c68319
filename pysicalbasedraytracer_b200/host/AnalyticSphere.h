// AnalyticSphere.h - an analytic sphere for the reference renderer: the CPU twin of the GPU sphere primitive (csrc/sphere.cuh).
//
// The reference's own PBR::Sphere (Shape/Sphere.h, Shape/Sphere.cpp:8-37) is a stub: Intersect computes the discriminant of
// the ray / sphere quadratic and returns `discriminant > 0` without filling tHit or the SurfaceInteraction, IntersectP returns
// false, Sample returns an empty Interaction - which is why every scene of the reference tessellates its spheres.
// Shape/README.md section 3 documents what the class is meant to do:
//   1. ray -> object space (WorldToObject);  2. A = d.d, B = 2 (d.o), C = o.o - r^2;  3. discriminant = B^2 - 4AC, miss if < 0;
//   4. of the roots t0 <= t1 the nearer one inside (0, ray.tMax) is tHit;  5. pHit = ray(tHit), n = pHit / r, uv = spherical
//   coordinates (phi, theta), dpdu / dpdv, everything back to world space.
// PBR::AnalyticSphere implements exactly that as a subclass of the reference's Shape, so it runs inside the reference's own
// BVHAccel / GeometricPrimitive / integrators; the GPU library implements the same arithmetic (pbrgpu_sphere).
//
// PARITY UNPINNED: nothing in the reference pins these values (its Sphere never produces a hit record).  The only pins are
// CPU twin == GPU (t bit-exact, radiance to 1e-3) and the known-answer rows of tests/test_gpu_sphere.py.  Choices made where
// the README is silent, all stated here:
//   * the quadratic is evaluated in DOUBLE from the float ray (the stub's float evaluation loses the ray-origin offset of a
//     spawned ray to cancellation in C = o.o - r^2, i.e. self-intersection; the reference has no EFloat);
//   * a hit needs discriminant > 0 (the stub's condition), t > 0 and t < ray.tMax;
//   * pHit is re-projected onto the sphere (pHit *= r / |pHit|), pError = gamma(5) (|pHit_object| + |pHit_world|);
//   * the parameterisation (phi = atan2(y, x) in [0, 2pi), theta = acos(z / r), u = phi / 2pi, v = (theta - pi) / (-pi),
//     dpdu = (-2pi y, 2pi x, 0), dpdv = -pi (z cos phi, z sin phi, -r sin theta), sin theta = sqrt(1 - (z/r)^2)) and the
//     area-light sampling (uniform cone outside, uniform area inside) are pbrt-v3's published Sphere;
//   * ObjectToWorld must be a pure translation (the GPU record is centre + radius).
#pragma once
#include <algorithm>
#include <cmath>

#include "Core/Interaction.h"
#include "Core/Transform.h"
#include "Sampler/Sampling.h"
#include "Shape/Shape.h"

namespace PBR {

class AnalyticSphere : public Shape {
 public:
  AnalyticSphere(const Transform* ObjectToWorld, const Transform* WorldToObject, bool reverseOrientation, float radius)
      : Shape(ObjectToWorld, WorldToObject, reverseOrientation), radius(radius) {}
  Bounds3f ObjectBound() const override { return Bounds3f(Point3f(-radius, -radius, -radius), Point3f(radius, radius, radius)); }
  float Area() const override { return 4 * Pi * radius * radius; }
  Point3f Center() const { return (*ObjectToWorld)(Point3f(0, 0, 0)); }

  // roots of A t^2 + B t + C = 0 in double; false when discriminant <= 0 or no root lies in (0, tMax)
  static bool SolveHit(const Ray& ray, float radius, float* tHit) {
    const double ox = ray.o.x, oy = ray.o.y, oz = ray.o.z, dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    const double a = dx * dx + dy * dy + dz * dz;
    const double b = 2.0 * (ox * dx + oy * dy + oz * dz);
    const double c = (ox * ox + oy * oy + oz * oz) - (double)radius * (double)radius;
    const double discriminant = b * b - 4.0 * a * c;
    if (!(discriminant > 0)) return false;
    const double root = std::sqrt(discriminant);
    const double t0 = (-b - root) / (2.0 * a), t1 = (-b + root) / (2.0 * a);
    const float f0 = (float)t0, f1 = (float)t1;
    if (f0 > 0 && f0 < ray.tMax) { *tHit = f0; return true; }
    if (f1 > 0 && f1 < ray.tMax) { *tHit = f1; return true; }
    return false;
  }

  bool Intersect(const Ray& r, float* tHit, SurfaceInteraction* isect, bool testAlphaTexture) const override {
    Ray ray = (*WorldToObject)(r);
    float t;
    if (!SolveHit(ray, radius, &t)) return false;
    if (tHit) *tHit = t;
    if (!isect) return true;
    Point3f pHit = ray(t);
    pHit *= radius / Distance(pHit, Point3f(0, 0, 0));
    if (pHit.x == 0 && pHit.y == 0) pHit.x = 1e-5f * radius;
    float phi = std::atan2(pHit.y, pHit.x);
    if (phi < 0) phi += 2 * Pi;
    const float thetaMin = Pi, thetaMax = 0;
    const float u = phi / (2 * Pi);
    const float cosTheta = Clamp(pHit.z / radius, -1, 1);
    const float theta = std::acos(cosTheta);
    const float sinTheta = std::sqrt(std::max(0.f, 1 - cosTheta * cosTheta));  // no libm in the normal / frame: bit-exact on the GPU
    const float v = (theta - thetaMin) / (thetaMax - thetaMin);
    const float zRadius = std::sqrt(pHit.x * pHit.x + pHit.y * pHit.y);
    const float invZRadius = 1 / zRadius;
    const float cosPhi = pHit.x * invZRadius, sinPhi = pHit.y * invZRadius;
    const Vector3f dpdu(-(2 * Pi) * pHit.y, (2 * Pi) * pHit.x, 0);
    const Vector3f dpdv = (thetaMax - thetaMin) * Vector3f(pHit.z * cosPhi, pHit.z * sinPhi, -radius * sinTheta);
    const Point3f pWorld = (*ObjectToWorld)(pHit);
    const Vector3f pError = gamma(5) * (Abs((Vector3f)pHit) + Abs((Vector3f)pWorld));
    // translation only: vectors are unchanged by ObjectToWorld
    *isect = SurfaceInteraction(pWorld, pError, Point2f(u, v), -r.d, dpdu, dpdv, Normal3f(0, 0, 0), Normal3f(0, 0, 0), r.time, this);
    return true;
  }
  bool IntersectP(const Ray& r, bool testAlphaTexture) const override {
    Ray ray = (*WorldToObject)(r);
    float t;
    return SolveHit(ray, radius, &t);
  }

  // uniform by area
  Interaction Sample(const Point2f& u, float* pdf) const override {
    Point3f pObj = Point3f(0, 0, 0) + radius * UniformSampleSphere(u);
    Interaction it;
    it.n = Normalize(Normal3f(pObj.x, pObj.y, pObj.z));
    if (reverseOrientation) it.n *= -1;
    pObj *= radius / Distance(pObj, Point3f(0, 0, 0));
    it.p = (*ObjectToWorld)(pObj);
    it.pError = gamma(5) * (Abs((Vector3f)pObj) + Abs((Vector3f)it.p));
    *pdf = 1 / Area();
    return it;
  }
  // by solid angle from `ref`: the cone the sphere subtends (outside), area sampling (inside)
  Interaction Sample(const Interaction& ref, const Point2f& u, float* pdf) const override {
    const Point3f pCenter = Center();
    const Point3f pOrigin = OffsetRayOrigin(ref.p, ref.pError, ref.n, pCenter - ref.p);
    if (DistanceSquared(pOrigin, pCenter) <= radius * radius) {
      Interaction intr = Sample(u, pdf);
      Vector3f wi = intr.p - ref.p;
      if (wi.LengthSquared() == 0) *pdf = 0;
      else {
        wi = Normalize(wi);
        *pdf *= DistanceSquared(ref.p, intr.p) / AbsDot(intr.n, -wi);
      }
      if (std::isinf(*pdf)) *pdf = 0.f;
      return intr;
    }
    const float dc = Distance(ref.p, pCenter);
    const float invDc = 1 / dc;
    const Vector3f wc = (pCenter - ref.p) * invDc;
    Vector3f wcX, wcY;
    CoordinateSystem(wc, &wcX, &wcY);
    const float sinThetaMax = radius * invDc;
    const float sinThetaMax2 = sinThetaMax * sinThetaMax;
    const float invSinThetaMax = 1 / sinThetaMax;
    const float cosThetaMax = std::sqrt(std::max((float)0.f, 1 - sinThetaMax2));
    float cosTheta = (cosThetaMax - 1) * u[0] + 1;
    float sinTheta2 = 1 - cosTheta * cosTheta;
    if (sinThetaMax2 < 0.00068523f /* sin^2(1.5 deg) */) {
      sinTheta2 = sinThetaMax2 * u[0];
      cosTheta = std::sqrt(1 - sinTheta2);
    }
    const float cosAlpha = sinTheta2 * invSinThetaMax +
                           cosTheta * std::sqrt(std::max((float)0.f, 1.f - sinTheta2 * invSinThetaMax * invSinThetaMax));
    const float sinAlpha = std::sqrt(std::max((float)0.f, 1.f - cosAlpha * cosAlpha));
    const float phi = u[1] * 2 * Pi;
    const Vector3f nWorld = SphericalDirection(sinAlpha, cosAlpha, phi, -wcX, -wcY, -wc);
    const Point3f pWorld = pCenter + radius * Point3f(nWorld.x, nWorld.y, nWorld.z);
    Interaction it;
    it.p = pWorld;
    it.pError = gamma(5) * Abs((Vector3f)pWorld);
    it.n = Normal3f(nWorld);
    if (reverseOrientation) it.n *= -1;
    *pdf = 1 / (2 * Pi * (1 - cosThetaMax));
    return it;
  }
  float Pdf(const Interaction& ref, const Vector3f& wi) const override {
    const Point3f pCenter = Center();
    const Point3f pOrigin = OffsetRayOrigin(ref.p, ref.pError, ref.n, pCenter - ref.p);
    if (DistanceSquared(pOrigin, pCenter) <= radius * radius) return Shape::Pdf(ref, wi);
    const float sinThetaMax2 = radius * radius / DistanceSquared(ref.p, pCenter);
    const float cosThetaMax = std::sqrt(std::max((float)0, 1 - sinThetaMax2));
    return UniformConePdf(cosThetaMax);
  }

  const float radius;
};

}  // namespace PBR
