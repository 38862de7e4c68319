// sampler.cuh - Halton sampler: integer radical inverse, bit-exact with the reference.
//   RadicalInverseSpecialized<b>            Sampler/LowDiscrepancy.cpp:211-225
//   RadicalInverse (base 2 via double)      Sampler/LowDiscrepancy.cpp:227-232
//   ScrambledRadicalInverseSpecialized<b>   Sampler/LowDiscrepancy.cpp:2300-2316
//   HaltonSampler::GetIndexForSample        Sampler/Halton.cpp:61-81
//   HaltonSampler::SampleDimension          Sampler/Halton.cpp:83-92
// The reference dispatches over 1024 template instantiations; here the base is a runtime value read
// from the prime table (same arithmetic, the digits and the fp32 product chain are identical).
#pragma once
#include "../../include/pbrgpu.h"
#include "vecmath.cuh"

namespace pbr {

struct HaltonParams {
  int32_t baseScales[2];
  int32_t baseExponents[2];
  int32_t sampleStride;
  int32_t multInverse[2];
  int32_t sampleAtPixelCenter;
  const uint16_t* perms;   // radicalInversePermutations (device)
  const int32_t* primes;   // Primes[1000]
  const int32_t* primeSums;
  const uint2* primeMagic; // per prime: (m, s) with floor(a / p) == __umulhi(a, m) >> s for every a < 2^31
  const float2* dimConst;  // per dimension: (invBase = 1/(float)p, tail = invBase * perm[0] / (1 - invBase)), host-computed
                           // with the same IEEE float operations the reference performs per call
  int64_t nPerms;
  // ---- Sobol' (type == PBRGPU_SAMPLER_SOBOL; host/SobolSampler.h is the CPU twin) ----
  int32_t type;                  // PBRGPU_SAMPLER_HALTON | PBRGPU_SAMPLER_SOBOL
  int32_t sobolDims;             // rows available in sobolMatrices32
  int32_t resolution, log2Resolution;
  int32_t boundsMinX, boundsMinY;
  const uint32_t* sobolMatrices32;  // [sobolDims][52]
  const uint64_t* sobolVdc;         // VdCSobolMatrices[log2Resolution-1][52]
  const uint64_t* sobolVdcInv;      // VdCSobolMatricesInv[log2Resolution-1][52]
};
constexpr int kSobolMatrixSize = 52;

// SobolIntervalToIndex / SobolSampleFloat: see host/SobolSampler.h (same integer arithmetic)
PBR_D int64_t sobol_index_for_sample(int px, int py, int64_t sampleNum, const HaltonParams& hp) {
  const uint32_t m = (uint32_t)hp.log2Resolution;
  if (m == 0) return 0;
  uint64_t frame = (uint64_t)sampleNum;
  uint64_t index = frame << (m << 1);
  uint64_t delta = 0;
  for (int c = 0; frame; frame >>= 1, ++c)
    if (frame & 1) delta ^= __ldg(hp.sobolVdc + c);
  uint64_t b = ((((uint64_t)(uint32_t)(px - hp.boundsMinX)) << m) | (uint64_t)(uint32_t)(py - hp.boundsMinY)) ^ delta;
  for (int c = 0; b; b >>= 1, ++c)
    if (b & 1) index ^= __ldg(hp.sobolVdcInv + c);
  return (int64_t)index;
}
PBR_D float sobol_sample(int64_t index, int dim, const HaltonParams& hp) {
  if (dim >= hp.sobolDims) return 0.f;
  uint32_t v = 0;
  uint64_t a = (uint64_t)index;
  const uint32_t* row = hp.sobolMatrices32 + dim * kSobolMatrixSize;
  for (int i = 0; a != 0; a >>= 1, ++i)
    if (a & 1) v ^= __ldg(row + i);
  return stdmin((float)v * 2.3283064365386963e-10f, kOneMinusEpsilon);
}
// dims 0 / 1 are remapped to the pixel (needs the pixel coordinate of that dimension)
PBR_D float sobol_film_sample(int64_t index, int dim, int pixelCoord, const HaltonParams& hp) {
  float s = sobol_sample(index, dim, hp);
  s = s * hp.resolution + (dim == 0 ? hp.boundsMinX : hp.boundsMinY);
  return stdmin(stdmax(s - pixelCoord, 0.f), kOneMinusEpsilon);
}

constexpr int kPrimeTableSize = 1000;
constexpr int kMaxResolution = 128;  // Halton.cpp:7

// digits of `a` in base `base`, reversed;  returns reversedDigits, *invBaseN = invBase^nDigits (fp32 chain).
// The reference divides uint64 by a compile-time base (1024 template instantiations).  Here the base is
// a run-time prime, so for a < 2^31 (every index of a <= 4096-spp 4K render) the quotient comes from a
// precomputed multiply-shift (Granlund-Montgomery: exact for all a < 2^31) instead of a hardware-less
// integer division; larger indices take the plain 64-bit path.  Integer results are identical.
PBR_D uint64_t reverse_digits(uint64_t a, uint32_t base, uint2 magic, float invBase, float* invBaseN, const uint16_t* perm) {
  uint64_t rev = 0;
  float ibn = 1;
  if ((a >> 31) == 0) {
    // all digits but the last accumulate in 32 bits (rev < base^(n-1) <= a < 2^31); only the final
    // rev * base + digit can exceed 32 bits
    uint32_t a32 = (uint32_t)a;
    uint32_t rev32 = 0;
    uint32_t next = __umulhi(a32, magic.x) >> magic.y;
    while (next) {
      uint32_t digit = a32 - next * base;
      rev32 = rev32 * base + (perm ? (uint32_t)__ldg(perm + digit) : digit);
      ibn *= invBase;
      a32 = next;
      next = __umulhi(a32, magic.x) >> magic.y;
    }
    if (a32) {  // most significant digit (a32 < base)
      rev = (uint64_t)rev32 * base + (perm ? (uint32_t)__ldg(perm + a32) : a32);
      ibn *= invBase;
    }
  } else {
    while (a) {
      uint64_t next = a / base;
      uint64_t digit = a - next * base;
      rev = rev * base + (perm ? (uint64_t)perm[digit] : digit);
      ibn *= invBase;
      a = next;
    }
  }
  *invBaseN = ibn;
  return rev;
}

PBR_D float radical_inverse(int baseIndex, uint64_t a, const HaltonParams& hp) {
  if (baseIndex == 0) {
    // ReverseBits64(a) * 5.4210108624275222e-20 : uint64 -> double product, narrowed on return
    return (float)((double)__brevll(a) * 5.4210108624275222e-20);
  }
  // LowDiscrepancy.cpp:2279-2280 returns 0 from 1024 on, but Primes[] / PrimeSums[] have 1000 entries: dimensions 1000..1023 read
  // past the tables in the reference (undefined).  Here every dimension beyond the tables yields 0.
  if (baseIndex >= kPrimeTableSize) return 0.f;
  uint32_t base = (uint32_t)hp.primes[baseIndex];
  float invBase = __ldg(hp.dimConst + baseIndex).x;
  float invBaseN;
  uint64_t rev = reverse_digits(a, base, __ldg(hp.primeMagic + baseIndex), invBase, &invBaseN, nullptr);
  return stdmin((float)rev * invBaseN, kOneMinusEpsilon);
}

PBR_D float scrambled_radical_inverse(int baseIndex, uint64_t a, const HaltonParams& hp) {
  if (baseIndex >= kPrimeTableSize) return 0.f;  // see radical_inverse
  uint32_t base = (uint32_t)hp.primes[baseIndex];
  const uint16_t* perm = hp.perms + hp.primeSums[baseIndex];
  const float2 dc = __ldg(hp.dimConst + baseIndex);  // invBase, invBase * perm[0] / (1 - invBase)
  float invBaseN;
  uint64_t rev = reverse_digits(a, base, __ldg(hp.primeMagic + baseIndex), dc.x, &invBaseN, perm);
  return stdmin(invBaseN * ((float)rev + dc.y), kOneMinusEpsilon);
}

PBR_D float halton_sample_dimension(int64_t index, int dim, const HaltonParams& hp) {
  if (hp.type == PBRGPU_SAMPLER_SOBOL) return sobol_sample(index, dim, hp);  // dims >= 2 (film dims: sampler_film_sample)
  if (hp.sampleAtPixelCenter && (dim == 0 || dim == 1)) return 0.5f;
  if (dim == 0) return radical_inverse(0, (uint64_t)(index >> hp.baseExponents[0]), hp);
  if (dim == 1) return radical_inverse(1, (uint64_t)(index / hp.baseScales[1]), hp);
  return scrambled_radical_inverse(dim, (uint64_t)index, hp);
}

template <int base>
PBR_D uint64_t inverse_radical_inverse(uint64_t inverse, int nDigits) {  // LowDiscrepancy.h:28-37
  uint64_t index = 0;
  for (int i = 0; i < nDigits; ++i) {
    uint64_t digit = inverse % base;
    inverse /= base;
    index = index * base + digit;
  }
  return index;
}

// Halton index of sample `sampleNum` of pixel (px,py)   (Halton.cpp:61-81)
PBR_D int64_t halton_index_for_sample(int px, int py, int64_t sampleNum, const HaltonParams& hp) {
  int64_t offset = 0;
  if (hp.sampleStride > 1) {
    int pmx = px % kMaxResolution, pmy = py % kMaxResolution;
    if (pmx < 0) pmx += kMaxResolution;
    if (pmy < 0) pmy += kMaxResolution;
    uint64_t d0 = inverse_radical_inverse<2>((uint64_t)pmx, hp.baseExponents[0]);
    uint64_t d1 = inverse_radical_inverse<3>((uint64_t)pmy, hp.baseExponents[1]);
    offset += (int64_t)(d0 * (uint64_t)(hp.sampleStride / hp.baseScales[0]) * (uint64_t)(int64_t)hp.multInverse[0]);
    offset += (int64_t)(d1 * (uint64_t)(hp.sampleStride / hp.baseScales[1]) * (uint64_t)(int64_t)hp.multInverse[1]);
    offset %= hp.sampleStride;
  }
  return offset + sampleNum * hp.sampleStride;
}

// Sampler-agnostic entry points used by the kernels
PBR_D int64_t sampler_index_for_sample(int px, int py, int64_t sampleNum, const HaltonParams& hp) {
  return hp.type == PBRGPU_SAMPLER_SOBOL ? sobol_index_for_sample(px, py, sampleNum, hp) : halton_index_for_sample(px, py, sampleNum, hp);
}
PBR_D float sampler_film_sample(int64_t index, int dim, int pixelCoord, const HaltonParams& hp) {
  return hp.type == PBRGPU_SAMPLER_SOBOL ? sobol_film_sample(index, dim, pixelCoord, hp) : halton_sample_dimension(index, dim, hp);
}

// Per-path sampler cursor: (index, dim) - Get1D/Get2D of GlobalSampler (Sampler.cpp:131-143) with
// arrayEndDim == arrayStartDim (no sample arrays requested by the Path integrator).
struct SamplerCursor {
  int64_t index;
  int dim;
  int arrayEndDim;  // == 5 unless DirectLighting/UniformSampleAll requested arrays
  PBR_D float get1D(const HaltonParams& hp) {
    if (dim >= 5 && dim < arrayEndDim) dim = arrayEndDim;
    return halton_sample_dimension(index, dim++, hp);
  }
  PBR_D void get2D(const HaltonParams& hp, float* u0, float* u1) {
    if (dim + 1 >= 5 && dim < arrayEndDim) dim = arrayEndDim;
    *u0 = halton_sample_dimension(index, dim, hp);
    *u1 = halton_sample_dimension(index, dim + 1, hp);
    dim += 2;
  }
};

}  // namespace pbr
