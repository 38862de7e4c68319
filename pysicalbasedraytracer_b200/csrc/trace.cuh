// trace.cuh - ray / triangle and ray / BVH queries.
//   Triangle::Intersect hit test   Shape/Triangle.cpp:66-145   (identical arithmetic, bit-exact t,b0,b1,b2)
//   Triangle::IntersectP           Shape/Triangle.cpp:262-350
//   BVHAccel::Intersect/IntersectP Accelerator/BVHAccel.cpp:285-366 (replaced: our own BVH + conservative
//   slab test; the reference's slab test Geometry.h:1437-1468 has no gamma widening, ours does, so the
//   set of triangles we test is a superset of what a watertight query needs).
#pragma once
#include "device_types.cuh"

namespace pbr {

// Ray-dependent part of the watertight test, hoisted out of the per-triangle loop.  The reference
// recomputes these per triangle from the same inputs (Triangle.cpp:74-86) - same values.
struct RayPrep {
  V3 o;
  int kx, ky, kz;
  float Sx, Sy, Sz;
};

PBR_D RayPrep prep_ray(V3 o, V3 d) {
  RayPrep rp;
  rp.o = o;
  int kz = max_dimension(vabs(d));
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float dx = d[kx], dy = d[ky], dz = d[kz];
  rp.kx = kx; rp.ky = ky; rp.kz = kz;
  rp.Sx = -dx / dz;
  rp.Sy = -dy / dz;
  rp.Sz = 1.f / dz;
  return rp;
}

// Returns true on hit and writes t, b0, b1, b2.  tMax is ray.tMax at the time of the call.
PBR_D bool intersect_triangle(const RayPrep& rp, float tMax, V3 p0, V3 p1, V3 p2, float* tOut, float* b0Out,
                              float* b1Out, float* b2Out) {
  V3 q0 = p0 - rp.o, q1 = p1 - rp.o, q2 = p2 - rp.o;
  V3 p0t(q0[rp.kx], q0[rp.ky], q0[rp.kz]);
  V3 p1t(q1[rp.kx], q1[rp.ky], q1[rp.kz]);
  V3 p2t(q2[rp.kx], q2[rp.ky], q2[rp.kz]);
  p0t.x += rp.Sx * p0t.z;
  p0t.y += rp.Sy * p0t.z;
  p1t.x += rp.Sx * p1t.z;
  p1t.y += rp.Sy * p1t.z;
  p2t.x += rp.Sx * p2t.z;
  p2t.y += rp.Sy * p2t.z;
  float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
  float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
  float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // fp64 fallback (Triangle.cpp:98-109)
    double p2txp1ty = (double)p2t.x * (double)p1t.y;
    double p2typ1tx = (double)p2t.y * (double)p1t.x;
    e0 = (float)(p2typ1tx - p2txp1ty);
    double p0txp2ty = (double)p0t.x * (double)p2t.y;
    double p0typ2tx = (double)p0t.y * (double)p2t.x;
    e1 = (float)(p0typ2tx - p0txp2ty);
    double p1txp0ty = (double)p1t.x * (double)p0t.y;
    double p1typ0tx = (double)p1t.y * (double)p0t.x;
    e2 = (float)(p1typ0tx - p1txp0ty);
  }
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
  float det = e0 + e1 + e2;
  if (det == 0) return false;
  p0t.z *= rp.Sz;
  p1t.z *= rp.Sz;
  p2t.z *= rp.Sz;
  float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
  if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return false;
  else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return false;
  float invDet = 1 / det;
  float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
  float t = tScaled * invDet;
  // conservative bound on the error of t (Triangle.cpp:133-145)
  float maxZt = max_component(vabs(V3(p0t.z, p1t.z, p2t.z)));
  float deltaZ = gamma_f(3) * maxZt;
  float maxXt = max_component(vabs(V3(p0t.x, p1t.x, p2t.x)));
  float maxYt = max_component(vabs(V3(p0t.y, p1t.y, p2t.y)));
  float deltaX = gamma_f(5) * (maxXt + maxZt);
  float deltaY = gamma_f(5) * (maxYt + maxZt);
  float deltaE = 2 * (gamma_f(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
  float maxE = max_component(vabs(V3(e0, e1, e2)));
  float deltaT = 3 * (gamma_f(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
  if (t <= deltaT) return false;
  *tOut = t; *b0Out = b0; *b1Out = b1; *b2Out = b2;
  return true;
}

struct HitRec {
  int tri;  // leaf-order triangle index, -1 = miss
  float t, b0, b1, b2;
};

PBR_D V3 ld3(const float4* p, int i) {
  float4 v = __ldg(p + i);
  return V3(v.x, v.y, v.z);
}

// Conservative slab test of both children of a node.  FMAs/fast min-max allowed here: the only
// requirement is that no box actually pierced by the ray inside (0, tMax] is rejected.
PBR_D void test_children(const Bvh2Node* __restrict__ nodes, int node, V3 o, V3 inv, float tMax, bool* h0, bool* h1,
                         float* t0, float* t1, int* c0, int* c1) {
  const float4* np = reinterpret_cast<const float4*>(nodes + node);
  float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
  int4 d = __ldg(reinterpret_cast<const int4*>(np + 3));
  const float widen = 1.0000004f;  // >= 1 + 2*gamma(3)
  float x0 = (a.x - o.x) * inv.x, x1 = (a.y - o.x) * inv.x;
  float y0 = (a.z - o.y) * inv.y, y1 = (a.w - o.y) * inv.y;
  float z0 = (c.x - o.z) * inv.z, z1 = (c.y - o.z) * inv.z;
  float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
  float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * widen;
  *h0 = (tn <= fminf(tf, tMax));
  *t0 = tn;
  x0 = (b.x - o.x) * inv.x; x1 = (b.y - o.x) * inv.x;
  y0 = (b.z - o.y) * inv.y; y1 = (b.w - o.y) * inv.y;
  z0 = (c.z - o.z) * inv.z; z1 = (c.w - o.z) * inv.z;
  tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
  tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * widen;
  *h1 = (tn <= fminf(tf, tMax));
  *t1 = tn;
  *c0 = d.x;
  *c1 = d.y;
}

template <bool ANY_HIT, bool COUNT>
PBR_D bool traverse(const SceneDev& sc, V3 o, V3 d, float tMax, HitRec* hit, unsigned* nNodes, unsigned* nTris) {
  RayPrep rp = prep_ray(o, d);
  V3 inv(1.f / d.x, 1.f / d.y, 1.f / d.z);
  int stack[64];
  int sp = 0;
  int cur = 0;
  hit->tri = -1;
  hit->t = tMax;
  bool found = false;
  while (true) {
    if (cur >= 0) {
      bool h0, h1;
      float t0, t1;
      int c0, c1;
      test_children(sc.nodes, cur, o, inv, tMax, &h0, &h1, &t0, &t1, &c0, &c1);
      if (COUNT) (*nNodes)++;
      if (h0 && h1) {
        if (t1 < t0) { int tmp = c0; c0 = c1; c1 = tmp; }
        stack[sp++] = c1;
        cur = c0;
        continue;
      } else if (h0) { cur = c0; continue; }
      else if (h1) { cur = c1; continue; }
    } else {
      unsigned code = (unsigned)(~cur);
      int first = (int)(code >> 2), cnt = (int)(code & 3u) + 1;
      for (int i = 0; i < cnt; ++i) {
        int ti = first + i;
        V3 p0 = ld3(sc.tv0, ti), p1 = ld3(sc.tv1, ti), p2 = ld3(sc.tv2, ti);
        float t, b0, b1, b2;
        if (COUNT) (*nTris)++;
        if (intersect_triangle(rp, tMax, p0, p1, p2, &t, &b0, &b1, &b2)) {
          if (ANY_HIT) return true;
          tMax = t;
          hit->tri = ti; hit->t = t; hit->b0 = b0; hit->b1 = b1; hit->b2 = b2;
          found = true;
        }
      }
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  return found;
}

}  // namespace pbr
