// host_tables.h - host-side integer tables and small fp32 helpers the flattener needs.
//   Primes / PrimeSums                        Sampler/LowDiscrepancy.cpp:9-208 (recomputed by a sieve)
//   PCG32 RNG                                 Sampler/RNG.h:93-108
//   ComputeRadicalInversePermutations         Sampler/LowDiscrepancy.cpp:2284-2298, Shuffle Sampler/Sampling.h:47-54
//   HaltonSampler ctor                        Sampler/Halton.cpp:30-58
//   Distribution1D ctor                       Sampler/Sampling.h:79-90
//   TrowbridgeReitzDistribution::RoughnessToAlpha   Material/Microfacet.h:78-83
// Everything here runs once per scene / sampler on the host; there is no per-sample CPU work.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace pbr {
namespace host {

constexpr int kPrimeTableSize = 1000;

inline void prime_tables(std::vector<int32_t>* primes, std::vector<int32_t>* sums) {
  primes->clear();
  sums->clear();
  std::vector<char> sieve(8000, 1);
  for (int i = 2; (int)primes->size() < kPrimeTableSize; ++i) {
    if (!sieve[i]) continue;
    primes->push_back(i);
    for (int j = i * i; j < (int)sieve.size(); j += i) sieve[j] = 0;
  }
  int s = 0;
  for (int p : *primes) { sums->push_back(s); s += p; }
}

// (m, s): floor(a / p) == (uint32)(((uint64)a * m) >> 32) >> s for every a < 2^31
inline void prime_magic(const std::vector<int32_t>& primes, std::vector<uint32_t>* out) {
  out->clear();
  for (int p : primes) {
    int L = 1;
    while ((1 << L) < p) ++L;
    uint64_t num = 1ull << (31 + L);
    uint64_t m = (num + (uint64_t)p - 1) / (uint64_t)p;
    out->push_back((uint32_t)m);
    out->push_back((uint32_t)(L - 1));
  }
}

struct Pcg32 {
  uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
  uint32_t next() {
    uint64_t old = state;
    state = old * 0x5851f42d4c957f2dULL + inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31));
  }
  uint32_t next(uint32_t b) {
    uint32_t threshold = (~b + 1u) % b;
    for (;;) {
      uint32_t r = next();
      if (r >= threshold) return r % b;
    }
  }
};

inline std::vector<uint16_t> radical_inverse_permutations(const std::vector<int32_t>& primes) {
  Pcg32 rng;
  size_t total = 0;
  for (int p : primes) total += (size_t)p;
  std::vector<uint16_t> perms(total);
  uint16_t* p = perms.data();
  for (int b : primes) {
    for (int j = 0; j < b; ++j) p[j] = (uint16_t)j;
    for (int i = 0; i < b; ++i) {
      int other = i + (int)rng.next((uint32_t)(b - i));
      std::swap(p[i], p[other]);
    }
    p += b;
  }
  return perms;
}

inline void extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {
  if (b == 0) { *x = 1; *y = 0; return; }
  int64_t d = (int64_t)(a / b), xp, yp;
  extended_gcd(b, a % b, &xp, &yp);
  *x = yp;
  *y = xp - (d * yp);
}
inline int64_t mod_i64(int64_t a, int64_t b) { int64_t r = a - (a / b) * b; return r < 0 ? r + b : r; }
inline uint64_t multiplicative_inverse(int64_t a, int64_t n) {
  int64_t x, y;
  extended_gcd((uint64_t)a, (uint64_t)n, &x, &y);
  return (uint64_t)mod_i64(x, n);
}

// out7 = baseScales[2], baseExponents[2], sampleStride, multInverse[2]
inline void halton_params(const int32_t bounds[4], int32_t out7[7]) {
  int res[2] = {bounds[2] - bounds[0], bounds[3] - bounds[1]};
  for (int i = 0; i < 2; ++i) {
    int base = (i == 0) ? 2 : 3, scale = 1, e = 0;
    while (scale < std::min(res[i], 128)) { scale *= base; ++e; }
    out7[i] = scale;
    out7[2 + i] = e;
  }
  out7[4] = out7[0] * out7[1];
  out7[5] = (int32_t)multiplicative_inverse(out7[1], out7[0]);
  out7[6] = (int32_t)multiplicative_inverse(out7[0], out7[1]);
}

struct Dist1D {
  std::vector<float> func, cdf;
  float funcInt = 0;
  void build(const float* f, int n) {
    func.assign(f, f + n);
    cdf.assign((size_t)n + 1, 0.f);
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] / n;
    funcInt = cdf[n];
    if (funcInt == 0) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= funcInt;
  }
};

inline float roughness_to_alpha(float roughness) {
  roughness = std::max(roughness, (float)1e-3);
  float x = std::log(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

// MIPMap's resampling of an image whose sides are not powers of two (Texture/MIPMap.h:45-64,93-144): separable 4-tap Lanczos
// (tau = 2, Texture/Texture.cpp:16-24) up to RoundUpPow2 in s, then in t, texels fetched through the wrap mode, result clamped
// to [0, inf).  RGB texels; the float operations and their order are the reference's, std::sin is the host libm on both sides.
// wrap: 0 = Repeat, 1 = Black, 2 = Clamp (PBRGPU_WRAP_*).
inline float lanczos(float x, float tau = 2.f) {
  x = std::abs(x);
  if (x < 1e-5f) return 1;
  if (x > 1.f) return 0;
  x *= 3.14159265358979323846f;
  float s = std::sin(x * tau) / (x * tau);
  float lz = std::sin(x) / x;
  return s * lz;
}
struct ResampleWeight {
  int firstTexel;
  float weight[4];
};
inline std::vector<ResampleWeight> resample_weights(int oldRes, int newRes) {
  std::vector<ResampleWeight> wt((size_t)newRes);
  const float filterwidth = 2.f;
  for (int i = 0; i < newRes; ++i) {
    float center = (i + .5f) * oldRes / newRes;
    wt[i].firstTexel = (int)std::floor((center - filterwidth) + 0.5f);
    for (int j = 0; j < 4; ++j) {
      float pos = wt[i].firstTexel + j + .5f;
      wt[i].weight[j] = lanczos((pos - center) / filterwidth);
    }
    float invSumWts = 1 / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
    for (int j = 0; j < 4; ++j) wt[i].weight[j] *= invSumWts;
  }
  return wt;
}
inline int round_up_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline int wrap_index(int i, int n, int wrap) {  // -1 = outside (contributes nothing)
  if (wrap == 0) { int r = i - (i / n) * n; i = r < 0 ? r + n : r; }
  else if (wrap == 2) i = std::max(0, std::min(n - 1, i));
  return (i >= 0 && i < n) ? i : -1;
}
// in: w*h*3 floats -> out: (*W2)*(*H2)*3 floats
inline std::vector<float> resample_to_pow2(const float* img, int w, int h, int wrap, int* W2, int* H2) {
  const int w2 = round_up_pow2(w), h2 = round_up_pow2(h);
  *W2 = w2; *H2 = h2;
  std::vector<float> res((size_t)w2 * h2 * 3, 0.f);
  const std::vector<ResampleWeight> sW = resample_weights(w, w2);
  for (int t = 0; t < h; ++t)
    for (int s = 0; s < w2; ++s) {
      float* dst = &res[3 * ((size_t)t * w2 + s)];
      dst[0] = dst[1] = dst[2] = 0.f;
      for (int j = 0; j < 4; ++j) {
        const int o = wrap_index(sW[s].firstTexel + j, w, wrap);
        if (o >= 0)
          for (int c = 0; c < 3; ++c) dst[c] += sW[s].weight[j] * img[3 * ((size_t)t * w + o) + c];
      }
    }
  const std::vector<ResampleWeight> tW = resample_weights(h, h2);
  std::vector<float> work((size_t)h2 * 3);
  for (int s = 0; s < w2; ++s) {
    for (int t = 0; t < h2; ++t) {
      float* wd = &work[3 * (size_t)t];
      wd[0] = wd[1] = wd[2] = 0.f;
      for (int j = 0; j < 4; ++j) {
        const int o = wrap_index(tW[t].firstTexel + j, h, wrap);
        if (o >= 0)
          for (int c = 0; c < 3; ++c) wd[c] += tW[t].weight[j] * res[3 * ((size_t)o * w2 + s) + c];
      }
    }
    for (int t = 0; t < h2; ++t)
      for (int c = 0; c < 3; ++c) res[3 * ((size_t)t * w2 + s) + c] = std::max(0.f, work[3 * (size_t)t + c]);  // Clamp(v, 0, Infinity)
  }
  return res;
}

}  // namespace host
}  // namespace pbr
