// SobolSampler.h - a Sobol' (0,2)-sequence GlobalSampler for the reference renderer.
//
// The reference ships the generator matrices (Sampler/SobolMatrices.h:42-47: SobolMatrices32,
// VdCSobolMatrices, VdCSobolMatricesInv) but NO code that consumes them - there is no SobolSampler class
// (SURVEY.md 8(c)).  This header adds one, as a subclass of the reference's own GlobalSampler
// (Sampler/Sampler.h:61-77: GetIndexForSample / SampleDimension), following the published construction
// (Gruenschloss et al., "(t,m,s)-nets and maximized minimum distance", the interval-to-index mapping and
// the 32-bit matrix product used by pbrt-v3's SobolSampler).  It is the CPU twin of the GPU sampler in
// csrc/sampler.cuh; the two are tested bit-for-bit against each other.
// PARITY UNPINNED: nothing in the reference pins these values; the only pins are CPU twin == GPU and the
// reference's own table bytes.
#pragma once
#include <algorithm>
#include <cstdint>
#include <memory>

#include "Sampler/Sampler.h"
#include "Sampler/SobolMatrices.h"
#include "Sampler/RNG.h"

namespace PBR {

inline uint64_t SobolIntervalToIndex(const uint32_t m, uint64_t frame, int px, int py) {
  if (m == 0) return 0;
  const uint32_t m2 = m << 1;
  uint64_t index = uint64_t(frame) << m2;
  uint64_t delta = 0;
  for (int c = 0; frame; frame >>= 1, ++c)
    if (frame & 1) delta ^= VdCSobolMatrices[m - 1][c];
  uint64_t b = (((uint64_t)((uint32_t)px) << m) | ((uint32_t)py)) ^ delta;
  for (int c = 0; b; b >>= 1, ++c)
    if (b & 1) index ^= VdCSobolMatricesInv[m - 1][c];
  return index;
}

inline float SobolSampleFloat(int64_t a, int dimension, uint32_t scramble = 0) {
  uint32_t v = scramble;
  for (int i = dimension * SobolMatrixSize; a != 0; a >>= 1, i++)
    if (a & 1) v ^= SobolMatrices32[i];
  return std::min(v * 2.3283064365386963e-10f /* 2^-32 */, OneMinusEpsilon);
}

class SobolSampler : public GlobalSampler {
 public:
  SobolSampler(int64_t samplesPerPixel, const Bounds2i& sampleBounds)
      : GlobalSampler(roundUpPow2(samplesPerPixel)), sampleBounds(sampleBounds) {
    Vector2i d = sampleBounds.pMax - sampleBounds.pMin;
    resolution = (int)roundUpPow2(std::max(d.x, d.y));
    log2Resolution = 0;
    while ((1 << log2Resolution) < resolution) ++log2Resolution;
  }
  int64_t GetIndexForSample(int64_t sampleNum) const override {
    return (int64_t)SobolIntervalToIndex((uint32_t)log2Resolution, (uint64_t)sampleNum, currentPixel.x - sampleBounds.pMin.x,
                                         currentPixel.y - sampleBounds.pMin.y);
  }
  float SampleDimension(int64_t index, int dim) const override {
    if (dim >= NumSobolDimensions) return 0.f;
    float s = SobolSampleFloat(index, dim);
    if (dim == 0 || dim == 1) {  // remap the first two dimensions to the pixel
      s = s * resolution + sampleBounds.pMin[dim];
      s = std::min(std::max(s - currentPixel[dim], 0.f), OneMinusEpsilon);
    }
    return s;
  }
  std::unique_ptr<Sampler> Clone(int seed) override { return std::unique_ptr<Sampler>(new SobolSampler(*this)); }
  static int64_t roundUpPow2(int64_t v) {
    int64_t r = 1;
    while (r < v) r <<= 1;
    return r;
  }
  const Bounds2i sampleBounds;
  int resolution, log2Resolution;
};

}  // namespace PBR
