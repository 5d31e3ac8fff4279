// quantize() for device code (src/word2bits.cpp:73-108) — shared by the training kernels and
// the evaluator.
#pragma once
#include <cuda_runtime.h>

namespace w2b {

constexpr unsigned kFull = 0xffffffffu;

// quantize(), :73-108.  BM = compile-time bitlevel for 0/1/2, 9 = decide at run time.
struct QParams {
  int bits;
  float seg;  // 2^(bits-1) for bits >= 4
};
template <int BM>
__device__ __forceinline__ float quant(float x, const QParams &q) {
  const int b = (BM == 9) ? q.bits : BM;
  if (b == 0) return x;
  if (b == 1) return x < 0.f ? -0.33333334f : 0.33333334f;  // sign/3; -0.0 and NaN take +
  if (b == 2) {
    float lv = (fabsf(x) <= 0.5f) ? 0.25f : 0.75f;  // NaN fails the test -> .75 as in :93-94
    return x < 0.f ? -lv : lv;
  }
  if (b >= 4) {
    int k = __float2int_rz(__fadd_rn(__fmul_rn(fabsf(x), q.seg), 0.5f));
    int segi = (int)q.seg;
    k = k > segi ? segi : k;
    float lv = __fdiv_rn((float)k, q.seg);
    return x < 0.f ? -lv : lv;
  }
  return x < 0.f ? -0.0f : 0.0f;  // bitlevel 3 (and < 0): no branch of :86-105 matches
}


}  // namespace w2b
