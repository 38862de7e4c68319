/*
 * TEST INFRASTRUCTURE ONLY - never linked into, imported by or executed from the product path.
 *
 * restate.c - plain-C CPU restatement of the bit-exact parts of the reference's PathIntegrator hot
 * path (G0T-cha/PysicalBasedRaytracer).  It exists so that parity can be checked on a machine where
 * the compiled reference (oracle/_ref/libpbr_ref.so) is not available, and as an independent second
 * opinion on the sampler / camera / triangle arithmetic.  Every function cites the reference
 * file:line it follows.  Compile with -ffp-contract=off (no FMA contraction), as the reference
 * oracle is.
 *
 * Pinning: tests/test_oracle_restate.py checks this file against (a) the golden vectors of
 * SURVEY.md 8(c) (hex bit patterns extracted from the compiled reference), (b) the committed
 * fixtures under tests/golden/ (generated from the compiled reference by tests/golden/make_golden.py)
 * and (c) oracle/_ref itself wherever it is present.
 *
 * Scope: Halton sampler (integer radical inverse, PCG32 permutations), camera ray generation,
 * watertight ray/triangle test with brute-force closest / any hit over a triangle soup (no BVH: the
 * answer of a ray query does not depend on the acceleration structure), film tone map.  Radiance
 * (BSDF / light sampling) is checked against oracle/_ref and the golden images, not restated here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Sampler/LowDiscrepancy.cpp:9-208: Primes / PrimeSums (recomputed: first 1000 primes) ---- */
#define PRIME_TABLE_SIZE 1000
static int g_primes[PRIME_TABLE_SIZE], g_prime_sums[PRIME_TABLE_SIZE];
static uint16_t* g_perms = 0;
static int64_t g_nperms = 0;

static void init_primes(void) {
  if (g_primes[0]) return;
  static char sieve[8000];
  memset(sieve, 1, sizeof(sieve));
  int n = 0, s = 0;
  for (int i = 2; n < PRIME_TABLE_SIZE; ++i) {
    if (!sieve[i]) continue;
    g_primes[n] = i;
    g_prime_sums[n] = s;
    s += i;
    ++n;
    for (int j = i * i; j < (int)sizeof(sieve); j += i) sieve[j] = 0;
  }
}
int rs_prime(int i) { init_primes(); return g_primes[i]; }
int rs_prime_sum(int i) { init_primes(); return g_prime_sums[i]; }

/* ---- Sampler/RNG.h:93-108 (PCG32, default state/stream) and :34-40 (UniformUInt32(b)) ---- */
typedef struct { uint64_t state, inc; } pcg32;
static uint32_t pcg_next(pcg32* r) {
  uint64_t old = r->state;
  r->state = old * 0x5851f42d4c957f2dULL + r->inc;
  uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  uint32_t rot = (uint32_t)(old >> 59u);
  return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}
static uint32_t pcg_bounded(pcg32* r, uint32_t b) {
  uint32_t threshold = (~b + 1u) % b;
  for (;;) {
    uint32_t v = pcg_next(r);
    if (v >= threshold) return v % b;
  }
}

/* ---- Sampler/LowDiscrepancy.cpp:2284-2298 ComputeRadicalInversePermutations; Shuffle Sampler/Sampling.h:47-54 ---- */
int64_t rs_perms(uint16_t* out, int64_t cap) {
  init_primes();
  if (!g_perms) {
    int64_t total = 0;
    for (int i = 0; i < PRIME_TABLE_SIZE; ++i) total += g_primes[i];
    g_perms = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)total);
    g_nperms = total;
    pcg32 rng = {0x853c49e6748fea9bULL, 0xda3e39cb94b95bdbULL};
    uint16_t* p = g_perms;
    for (int i = 0; i < PRIME_TABLE_SIZE; ++i) {
      int b = g_primes[i];
      for (int j = 0; j < b; ++j) p[j] = (uint16_t)j;
      for (int j = 0; j < b; ++j) {
        int other = j + (int)pcg_bounded(&rng, (uint32_t)(b - j));
        uint16_t t = p[j]; p[j] = p[other]; p[other] = t;
      }
      p += b;
    }
  }
  if (out) memcpy(out, g_perms, sizeof(uint16_t) * (size_t)(cap < g_nperms ? cap : g_nperms));
  return g_nperms;
}

static const float kOneMinusEpsilon = 0.99999994f; /* Sampler/RNG.h:13 */

/* ---- Sampler/LowDiscrepancy.h:15-27 ---- */
static uint32_t reverse_bits32(uint32_t n) {
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
  n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
  n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
  n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
  return n;
}
static uint64_t reverse_bits64(uint64_t n) {
  uint64_t n0 = reverse_bits32((uint32_t)n), n1 = reverse_bits32((uint32_t)(n >> 32));
  return (n0 << 32) | n1;
}

/* ---- Sampler/LowDiscrepancy.cpp:211-232 RadicalInverse (the 1024-way switch restated with a runtime base) ---- */
float rs_radical_inverse(int baseIndex, uint64_t a) {
  init_primes();
  if (baseIndex == 0) return (float)((double)reverse_bits64(a) * 5.4210108624275222e-20);
  if (baseIndex >= PRIME_TABLE_SIZE) return 0.f; /* :2279-2280 */
  const uint64_t base = (uint64_t)g_primes[baseIndex];
  const float invBase = (float)1 / (float)base;
  uint64_t rev = 0;
  float invBaseN = 1;
  while (a) {
    uint64_t next = a / base, digit = a - next * base;
    rev = rev * base + digit;
    invBaseN *= invBase;
    a = next;
  }
  float v = (float)rev * invBaseN;
  return v < kOneMinusEpsilon ? v : kOneMinusEpsilon;
}

/* ---- Sampler/LowDiscrepancy.cpp:2300-2316 ScrambledRadicalInverse ---- */
float rs_scrambled_radical_inverse(int baseIndex, uint64_t a) {
  rs_perms(0, 0);
  if (baseIndex >= PRIME_TABLE_SIZE) return 0.f;
  const uint64_t base = (uint64_t)g_primes[baseIndex];
  const uint16_t* perm = g_perms + g_prime_sums[baseIndex];
  const float invBase = (float)1 / (float)base;
  uint64_t rev = 0;
  float invBaseN = 1;
  while (a) {
    uint64_t next = a / base, digit = a - next * base;
    rev = rev * base + perm[digit];
    invBaseN *= invBase;
    a = next;
  }
  float v = invBaseN * ((float)rev + invBase * (float)(int)perm[0] / (1 - invBase));
  return v < kOneMinusEpsilon ? v : kOneMinusEpsilon;
}

/* ---- Sampler/Halton.cpp:9-28,40-57: ctor: baseScales/Exponents, stride, multiplicative inverses ---- */
static void extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {
  if (b == 0) { *x = 1; *y = 0; return; }
  int64_t d = (int64_t)(a / b), xp, yp;
  extended_gcd(b, a % b, &xp, &yp);
  *x = yp;
  *y = xp - (d * yp);
}
static int64_t mod64(int64_t a, int64_t b) { int64_t r = a - (a / b) * b; return r < 0 ? r + b : r; }
void rs_halton_params(const int32_t bounds[4], int32_t out7[7]) {
  int res[2] = {bounds[2] - bounds[0], bounds[3] - bounds[1]};
  for (int i = 0; i < 2; ++i) {
    int base = i == 0 ? 2 : 3, scale = 1, e = 0;
    int lim = res[i] < 128 ? res[i] : 128; /* kMaxResolution, Halton.cpp:7 */
    while (scale < lim) { scale *= base; ++e; }
    out7[i] = scale;
    out7[2 + i] = e;
  }
  out7[4] = out7[0] * out7[1];
  int64_t x, y;
  extended_gcd((uint64_t)out7[1], (uint64_t)out7[0], &x, &y);
  out7[5] = (int32_t)mod64(x, out7[0]);
  extended_gcd((uint64_t)out7[0], (uint64_t)out7[1], &x, &y);
  out7[6] = (int32_t)mod64(x, out7[1]);
}

/* ---- Sampler/Halton.cpp:61-81 GetIndexForSample; LowDiscrepancy.h:28-37 InverseRadicalInverse ---- */
static uint64_t inverse_radical_inverse(int base, uint64_t inverse, int nDigits) {
  uint64_t index = 0;
  for (int i = 0; i < nDigits; ++i) {
    uint64_t digit = inverse % (uint64_t)base;
    inverse /= (uint64_t)base;
    index = index * (uint64_t)base + digit;
  }
  return index;
}
int64_t rs_halton_index(const int32_t p7[7], int px, int py, int64_t sampleNum) {
  int64_t offset = 0;
  const int stride = p7[4];
  if (stride > 1) {
    int pm[2] = {(int)mod64(px, 128), (int)mod64(py, 128)};
    for (int i = 0; i < 2; ++i) {
      uint64_t dimOffset = inverse_radical_inverse(i == 0 ? 2 : 3, (uint64_t)pm[i], p7[2 + i]);
      offset += (int64_t)(dimOffset * (uint64_t)(stride / p7[i]) * (uint64_t)(int64_t)p7[5 + i]);
    }
    offset %= stride;
  }
  return offset + sampleNum * stride;
}

/* ---- Sampler/Halton.cpp:83-92 SampleDimension ---- */
float rs_halton_sample(const int32_t p7[7], int64_t index, int dim) {
  if (dim == 0) return rs_radical_inverse(0, (uint64_t)(index >> p7[2]));
  if (dim == 1) return rs_radical_inverse(1, (uint64_t)(index / p7[1]));
  return rs_scrambled_radical_inverse(dim, (uint64_t)index);
}

/* out[pixel][sample][dim] - the stream Sampler::Get1D() yields with no sample arrays (Sampler.cpp:131-143) */
void rs_sampler_dump(const int32_t bounds[4], int x0, int y0, int x1, int y1, int64_t s0, int64_t s1, int nDims, float* out) {
  int32_t p7[7];
  rs_halton_params(bounds, p7);
  size_t k = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x)
      for (int64_t s = s0; s < s1; ++s) {
        int64_t idx = rs_halton_index(p7, x, y, s);
        for (int d = 0; d < nDims; ++d) out[k++] = rs_halton_sample(p7, idx, d);
      }
}

/* ---- Core/Transform.h:138-170 point / vector transforms ---- */
static void xf_point(const float* m, const float p[3], float out[3]) {
  float x = p[0], y = p[1], z = p[2];
  float xp = m[0] * x + m[1] * y + m[2] * z + m[3];
  float yp = m[4] * x + m[5] * y + m[6] * z + m[7];
  float zp = m[8] * x + m[9] * y + m[10] * z + m[11];
  float wp = m[12] * x + m[13] * y + m[14] * z + m[15];
  if (wp == 1) { out[0] = xp; out[1] = yp; out[2] = zp; return; }
  float inv = 1.f / wp; /* Point3::operator/ multiplies by the reciprocal, Geometry.h:486-491 */
  out[0] = inv * xp; out[1] = inv * yp; out[2] = inv * zp;
}
static void xf_vec(const float* m, const float v[3], float out[3]) {
  out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  out[1] = m[4] * v[0] + m[5] * v[1] + m[6] * v[2];
  out[2] = m[8] * v[0] + m[9] * v[1] + m[10] * v[2];
}

/* ---- Camera/Perspective.cpp:44-80, Camera/Orthographic.cpp:7-25 with lensRadius == 0; Sampler.cpp:10-21 ----
 * out[8] = o.xyz, tMax(inf), d.xyz, 0 */
void rs_primary_ray(int camType, const float* r2c, const float* c2w, const int32_t p7[7], int px, int py, int64_t s, float out[8]) {
  int64_t idx = rs_halton_index(p7, px, py, s);
  float pFilm[3] = {(float)px + rs_halton_sample(p7, idx, 0), (float)py + rs_halton_sample(p7, idx, 1), 0.f};
  float pc[3], o[3], d[3];
  xf_point(r2c, pFilm, pc);
  if (camType == 0) {
    o[0] = o[1] = o[2] = 0.f;
    float len = sqrtf(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
    float inv = 1.f / len; /* Normalize = v / Length(), operator/ multiplies by 1/f */
    d[0] = pc[0] * inv; d[1] = pc[1] * inv; d[2] = pc[2] * inv;
  } else {
    o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2];
    d[0] = 0; d[1] = 0; d[2] = 1;
  }
  xf_point(c2w, o, out);
  out[3] = INFINITY;
  xf_vec(c2w, d, out + 4);
  out[7] = 0.f;
}

/* gamma(n), Core/PBR.h:21-24: evaluated in double, narrowed */
static float gamma_n(int n) {
  const double eps = 1.1920928955078125e-07 * 0.5;
  return (float)((n * eps) / (1 - n * eps));
}

/* ---- Shape/Triangle.cpp:62-145 watertight hit test.  Returns 1 and t, b0,b1,b2 on hit. ---- */
static int tri_hit(const float o[3], const float d[3], float tMax, const float* p0, const float* p1, const float* p2, float* tOut, float b[3]) {
  float ad[3] = {fabsf(d[0]), fabsf(d[1]), fabsf(d[2])};
  int kz = (ad[0] > ad[1]) ? ((ad[0] > ad[2]) ? 0 : 2) : ((ad[1] > ad[2]) ? 1 : 2);
  int kx = kz + 1; if (kx == 3) kx = 0;
  int ky = kx + 1; if (ky == 3) ky = 0;
  float q0[3] = {p0[0] - o[0], p0[1] - o[1], p0[2] - o[2]};
  float q1[3] = {p1[0] - o[0], p1[1] - o[1], p1[2] - o[2]};
  float q2[3] = {p2[0] - o[0], p2[1] - o[1], p2[2] - o[2]};
  float dx = d[kx], dy = d[ky], dz = d[kz];
  float p0t[3] = {q0[kx], q0[ky], q0[kz]}, p1t[3] = {q1[kx], q1[ky], q1[kz]}, p2t[3] = {q2[kx], q2[ky], q2[kz]};
  float Sx = -dx / dz, Sy = -dy / dz, Sz = 1.f / dz;
  p0t[0] += Sx * p0t[2]; p0t[1] += Sy * p0t[2];
  p1t[0] += Sx * p1t[2]; p1t[1] += Sy * p1t[2];
  p2t[0] += Sx * p2t[2]; p2t[1] += Sy * p2t[2];
  float e0 = p1t[0] * p2t[1] - p1t[1] * p2t[0];
  float e1 = p2t[0] * p0t[1] - p2t[1] * p0t[0];
  float e2 = p0t[0] * p1t[1] - p0t[1] * p1t[0];
  if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
    double a = (double)p2t[0] * (double)p1t[1], c = (double)p2t[1] * (double)p1t[0];
    e0 = (float)(c - a);
    a = (double)p0t[0] * (double)p2t[1]; c = (double)p0t[1] * (double)p2t[0];
    e1 = (float)(c - a);
    a = (double)p1t[0] * (double)p0t[1]; c = (double)p1t[1] * (double)p0t[0];
    e2 = (float)(c - a);
  }
  if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return 0;
  float det = e0 + e1 + e2;
  if (det == 0) return 0;
  p0t[2] *= Sz; p1t[2] *= Sz; p2t[2] *= Sz;
  float tScaled = e0 * p0t[2] + e1 * p1t[2] + e2 * p2t[2];
  if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return 0;
  else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return 0;
  float invDet = 1 / det;
  b[0] = e0 * invDet; b[1] = e1 * invDet; b[2] = e2 * invDet;
  float t = tScaled * invDet;
#define MAX3ABS(a, b_, c) fmaxf(fabsf(a), fmaxf(fabsf(b_), fabsf(c)))
  float maxZt = MAX3ABS(p0t[2], p1t[2], p2t[2]);
  float deltaZ = gamma_n(3) * maxZt;
  float maxXt = MAX3ABS(p0t[0], p1t[0], p2t[0]);
  float maxYt = MAX3ABS(p0t[1], p1t[1], p2t[1]);
  float deltaX = gamma_n(5) * (maxXt + maxZt);
  float deltaY = gamma_n(5) * (maxYt + maxZt);
  float deltaE = 2 * (gamma_n(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
  float maxE = MAX3ABS(e0, e1, e2);
  float deltaT = 3 * (gamma_n(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
  if (t <= deltaT) return 0;
  *tOut = t;
  return 1;
}

/* Brute-force Scene::Intersect / IntersectP over a triangle soup.  The reference's BVH keeps the LAST
 * accepted hit in its traversal order at equal t (BVHAccel.cpp:285-328 with the non-strict tMax test);
 * a brute-force loop cannot know that order, so ties are reported through *nTies and the caller
 * excludes them (they are the "ties" of the north_star's counted mismatch fraction).
 * rays float[n][8] (o,tMax,d,pad); out int32[n][4] = prim, t bits, 0, 0   (any_hit: out int32[n]). */
void rs_trace(const float* verts, const int32_t* indices, int nTris, const float* rays, int64_t n, int anyHit, int32_t* out, int64_t* nTies) {
  int64_t ties = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* r = rays + 8 * i;
    float tMax = r[3];
    int best = -1, tie = 0;
    for (int k = 0; k < nTris; ++k) {
      float t, b[3];
      if (tri_hit(r, r + 4, tMax, verts + 3 * indices[3 * k], verts + 3 * indices[3 * k + 1], verts + 3 * indices[3 * k + 2], &t, b)) {
        if (anyHit) { best = k; break; }
        tie = (best >= 0 && t == tMax);
        tMax = t;
        best = k;
      }
    }
    if (anyHit) out[i] = best >= 0;
    else {
      out[4 * i] = best;
      memcpy(&out[4 * i + 1], &tMax, 4);
      out[4 * i + 2] = tie;
      out[4 * i + 3] = 0;
      ties += tie;
    }
  }
  if (nTies) *nTies = ties;
}

/* ---- Integrator/Integrator.cpp:327-344 + Core/Spectrum.h:10-25: mean RGB -> XYZ -> RGB -> gamma -> 8 bit ---- */
void rs_tonemap(const float* rgbMean, int n, unsigned char* rgba8) {
  for (int i = 0; i < n; ++i) {
    const float* c = rgbMean + 3 * i;
    float xyz[3], rgb[3];
    xyz[0] = 0.412453f * c[0] + 0.357580f * c[1] + 0.180423f * c[2];
    xyz[1] = 0.212671f * c[0] + 0.715160f * c[1] + 0.072169f * c[2];
    xyz[2] = 0.019334f * c[0] + 0.119193f * c[1] + 0.950227f * c[2];
    rgb[0] = 3.240479f * xyz[0] - 1.537150f * xyz[1] - 0.498535f * xyz[2];
    rgb[1] = -0.969256f * xyz[0] + 1.875991f * xyz[1] + 0.041556f * xyz[2];
    rgb[2] = 0.055648f * xyz[0] - 0.204043f * xyz[1] + 1.057311f * xyz[2];
    for (int k = 0; k < 3; ++k) {
      float v = rgb[k];
      float g = (v <= 0.0031308f) ? 12.92f * v : 1.055f * powf(v, (1.f / 2.4f)) - 0.055f;
      float q = 255.f * g + 0.5f;
      q = q < 0.f ? 0.f : (q > 255.f ? 255.f : q);
      rgba8[4 * i + k] = (unsigned char)q;
    }
    rgba8[4 * i + 3] = 255;
  }
}

/* ---- participating media: HomogeneousMedium::{Sample,Tr} (Media/HomogeneousMedium.cpp:10-45),
 *      PhaseHG / HenyeyGreenstein::{p,Sample_p} (Media/Medium.h:21-24, Media/Medium.cpp:7-28),
 *      CoordinateSystem (Core/Geometry.h), SphericalDirection (Core/Geometry.h:1510-1515).
 *      Same argument layout as ref_medium_kat / ref_hg_kat in ref_harness.cpp. ---- */
static float vlen3(const float v[3]) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
#define RS_MAXFLOAT 3.402823466e+38f
void rs_medium_kat(const float sigma_a[3], const float sigma_s[3], float g, const float* in, int n, float* out) {
  (void)g;
  float sigma_t[3];
  for (int c = 0; c < 3; ++c) sigma_t[c] = sigma_s[c] + sigma_a[c];
  for (int i = 0; i < n; ++i) {
    const float* r = in + 8 * i;
    float* o = out + 8 * i;
    const float dLen = vlen3(r), tMax = r[3], u0 = r[4], u1 = r[5];
    int channel = (int)(u0 * 3);
    if (channel > 2) channel = 2;
    float dist = -logf(1 - u1) / sigma_t[channel];
    float t = dist / dLen;
    if (tMax < t) t = tMax;
    int sampled = t < tMax;
    float Tr[3], density[3], pdf = 0;
    float tc = RS_MAXFLOAT < t ? RS_MAXFLOAT : t;  /* std::min(t, MaxFloat): a NaN t stays NaN */
    for (int c = 0; c < 3; ++c) {
      Tr[c] = expf(-sigma_t[c] * tc * dLen);
      density[c] = sampled ? sigma_t[c] * Tr[c] : Tr[c];
    }
    for (int c = 0; c < 3; ++c) pdf += density[c];
    pdf *= 1 / (float)3;
    if (pdf == 0) pdf = 1;
    for (int c = 0; c < 3; ++c) o[c] = sampled ? Tr[c] * sigma_s[c] / pdf : Tr[c] / pdf;
    int ax = fabsf(r[0]) > fabsf(r[1]) ? (fabsf(r[0]) > fabsf(r[2]) ? 0 : 2) : (fabsf(r[1]) > fabsf(r[2]) ? 1 : 2);
    o[3] = sampled ? 0.f + r[ax] * t : tMax;
    o[4] = sampled ? 1.f : 0.f;
    float tl = tMax * dLen;
    if (RS_MAXFLOAT < tl) tl = RS_MAXFLOAT;
    for (int c = 0; c < 3; ++c) o[5 + c] = expf(-sigma_t[c] * tl);
  }
}
static float phase_hg(float cosTheta, float g) {
  float denom = 1 + g * g + 2 * g * cosTheta;
  return 0.07957747154594766788f * (1 - g * g) / (denom * sqrtf(denom));
}
void rs_hg_kat(float g, const float* in, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    const float* r = in + 8 * i;
    float* o = out + 8 * i;
    const float wo[3] = {r[0], r[1], r[2]}, u0 = r[3], u1 = r[4];
    float cosTheta;
    if (fabsf(g) < 1e-3)
      cosTheta = 1 - 2 * u0;
    else {
      float sqrTerm = (1 - g * g) / (1 + g - 2 * g * u0);
      cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g);
    }
    float s2 = 1 - cosTheta * cosTheta;
    float sinTheta = sqrtf(s2 > 0 ? s2 : 0.f);
    float phi = 2 * 3.14159265358979323846f * u1;
    float v1[3], v2[3];
    if (fabsf(wo[0]) > fabsf(wo[1])) {  /* CoordinateSystem; Vector3::operator/ multiplies by the reciprocal */
      float inv = (float)1 / sqrtf(wo[0] * wo[0] + wo[2] * wo[2]);
      v1[0] = -wo[2] * inv; v1[1] = 0 * inv; v1[2] = wo[0] * inv;
    } else {
      float inv = (float)1 / sqrtf(wo[1] * wo[1] + wo[2] * wo[2]);
      v1[0] = 0 * inv; v1[1] = wo[2] * inv; v1[2] = -wo[1] * inv;
    }
    /* Cross(v1, v2 = wo x v1) in double like Core/Geometry.h */
    {
      double ax = wo[0], ay = wo[1], az = wo[2], bx = v1[0], by = v1[1], bz = v1[2];
      v2[0] = (float)((ay * bz) - (az * by)); v2[1] = (float)((az * bx) - (ax * bz)); v2[2] = (float)((ax * by) - (ay * bx));
    }
    float a = sinTheta * cosf(phi), b = sinTheta * sinf(phi);
    for (int c = 0; c < 3; ++c) o[c] = (a * v1[c] + b * v2[c]) + cosTheta * wo[c];
    o[3] = phase_hg(cosTheta, g);
    o[4] = phase_hg(wo[0] * r[5] + wo[1] * r[6] + wo[2] * r[7], g);
    o[5] = o[6] = o[7] = 0.f;
  }
}
