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

constexpr int kSobolMatrixSize = 52;

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
  // v = XOR of the matrix columns selected by the set bits of the index (SobolSampleFloat's loop visits every bit up to the
  // highest one and branches on it; XOR does not care about the order, so only the set bits are visited here: about half the
  // iterations, no data-dependent branch inside, independent loads)
  uint32_t v = 0;
  uint32_t lo = (uint32_t)(uint64_t)index, hi = (uint32_t)((uint64_t)index >> 32);
  const uint32_t* row = hp.sobolMatrices32 + dim * kSobolMatrixSize;
  while (lo) { v ^= __ldg(row + (__ffs((int)lo) - 1)); lo &= lo - 1u; }
  hi &= (1u << (kSobolMatrixSize - 32)) - 1u;  // an index has at most kSobolMatrixSize bits
  while (hi) { v ^= __ldg(row + 32 + (__ffs((int)hi) - 1)); hi &= hi - 1u; }
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
// indices of 2^31 and more (beyond every BASELINE configuration: 4096 spp at 4K stay below 2^27): plain 64-bit division, out
// of line - inlined, its division code was ~650 of k_shade<Matte>'s 10 400 instructions, in a kernel that already stalls on
// instruction fetch
static __device__ __noinline__ uint64_t reverse_digits_wide(uint64_t a, uint32_t base, float invBase, float* invBaseN, const uint16_t* perm) {
  uint64_t rev = 0;
  float ibn = 1;
  while (a) {
    uint64_t next = a / base;
    uint64_t digit = a - next * base;
    rev = rev * base + (perm ? (uint64_t)perm[digit] : digit);
    ibn *= invBase;
    a = next;
  }
  *invBaseN = ibn;
  return rev;
}
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
    return reverse_digits_wide(a, base, invBase, invBaseN, perm);
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

// Four scrambled radical inverses of the SAME index (dimensions dim0 .. dim0+3), their digit loops advanced in lock step.
// One digit chain is a serial dependency (quotient -> digit -> permutation lookup -> accumulate): evaluated one dimension
// after the other, the dimensions of a scattering event run at the latency of that chain.  Interleaved, four independent
// chains are in flight.  Integer digits and the fp32 product
// chain are exactly those of reverse_digits / scrambled_radical_inverse (LowDiscrepancy.cpp:2300-2316).
// Preconditions (the caller falls back to halton_sample_dimension otherwise): a < 2^31, 2 <= dim0, dim0 + 4 <= kPrimeTableSize.
// Chains k >= nValid are idle (their out[k] is meaningless).
PBR_D void scrambled_radical_inverse_x4(int dim0, int nValid, uint32_t a, const HaltonParams& hp, float out[4]) {
  uint32_t base[4], mm[4], ms[4], a32[4], rev32[4], poff[4];
  float ib[4], tail[4], ibn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int bi = dim0 + k;
    base[k] = (uint32_t)__ldg(hp.primes + bi);
    poff[k] = (uint32_t)__ldg(hp.primeSums + bi);
    const uint2 mg = __ldg(hp.primeMagic + bi);
    mm[k] = mg.x; ms[k] = mg.y;
    const float2 dc = __ldg(hp.dimConst + bi);
    ib[k] = dc.x; tail[k] = dc.y;
    a32[k] = k < nValid ? a : 0u;
    rev32[k] = 0u;
    ibn[k] = 1.f;
  }
  bool more;
  do {
    more = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t next = __umulhi(a32[k], mm[k]) >> ms[k];
      if (next) {
        const uint32_t digit = a32[k] - next * base[k];
        rev32[k] = rev32[k] * base[k] + (uint32_t)__ldg(hp.perms + poff[k] + digit);
        ibn[k] *= ib[k];
        a32[k] = next;
        more = true;
      }
    }
  } while (more);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint64_t rev = rev32[k];
    float n = ibn[k];
    if (a32[k]) {  // most significant digit: only this last step can leave 32 bits
      rev = (uint64_t)rev32[k] * base[k] + (uint32_t)__ldg(hp.perms + poff[k] + a32[k]);
      n *= ib[k];
    }
    out[k] = stdmin(n * ((float)rev + tail[k]), kOneMinusEpsilon);
  }
}

// Out-of-line copy of halton_sample_dimension for the kernels that normally read precomputed values (SamplerCursorX4::prefetch).
static __device__ __noinline__ float sample_dimension_outofline(int64_t index, int dim, const HaltonParams* hp) {
  return halton_sample_dimension(index, dim, *hp);
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

// The same cursor with a window of prefetched dimensions: prefetch(hp, n) evaluates the next n <= 8 dimensions at once
// (interleaved digit loops, see scrambled_radical_inverse_x4); get1D / get2D then return those values, and anything outside
// the window goes through the out-of-line evaluation.  Used by the medium kernels (k_medium_presample, k_shade_medium:
// -6 % on the smoke volume).  Measured on k_shade<Matte> as well (one window of 8, windows of 3 + 4, 3 + 5): 3-7 % SLOWER
// there - the window's registers and the call cost more than the interleaving gains in a kernel that already spills.
struct SamplerCursorX4 : SamplerCursor {
  int preDim0 = 0, preN = 0;
  float pre[8];
  PBR_D explicit SamplerCursorX4(const SamplerCursor& c) : SamplerCursor(c) {}
  PBR_D void prefetch(const HaltonParams& hp, int n) {
    preN = 0;
    if (hp.type != PBRGPU_SAMPLER_HALTON || ((uint64_t)index >> 31) != 0 || dim < 5 || dim < arrayEndDim || dim + 8 > kPrimeTableSize) return;
    preDim0 = dim;
    preN = n;
    scrambled_radical_inverse_x4(dim, n < 4 ? n : 4, (uint32_t)index, hp, pre);
    if (n > 4) scrambled_radical_inverse_x4(dim + 4, n - 4, (uint32_t)index, hp, pre + 4);
  }
  // make sure the next n dimensions are inside the prefetched window
  PBR_D void ensure(const HaltonParams& hp, int n) {
    const unsigned k = (unsigned)(dim - preDim0);
    if (!(k < (unsigned)preN && k + (unsigned)n <= (unsigned)preN)) prefetch(hp, n);
  }
  PBR_D float prefetched(int k) const {  // pre[k] without run-time indexing (the array stays in registers)
    float r = pre[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) r = k == j ? pre[j] : r;
    return r;
  }
  PBR_D float value(const HaltonParams& hp, int d) const {
    const unsigned k = (unsigned)(d - preDim0);
    return k < (unsigned)preN ? prefetched((int)k) : sample_dimension_outofline(index, d, &hp);
  }
  PBR_D float get1D(const HaltonParams& hp) {
    if (dim >= 5 && dim < arrayEndDim) dim = arrayEndDim;
    return value(hp, dim++);
  }
  PBR_D void get2D(const HaltonParams& hp, float* u0, float* u1) {
    if (dim + 1 >= 5 && dim < arrayEndDim) dim = arrayEndDim;
    *u0 = value(hp, dim);
    *u1 = value(hp, dim + 1);
    dim += 2;
  }
};

}  // namespace pbr
