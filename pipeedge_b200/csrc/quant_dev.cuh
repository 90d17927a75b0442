// Device-side pieces of QuantPipe shared by quant.cu (stand-alone encode / decode) and link.cu (quantisation fused into
// the inter-stage send / receive kernels). Everything here is written so that both users produce the same bits.
#pragma once
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace pe {

constexpr int kQMaxChunks = 64;      // partial-reduction chunks per item
constexpr int kQPartialDoubles = 5;  // min, max, sum, sumsq, sumsq of fp32-rounded squares

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// basic_op.py:127-130 (`_quant_op`): clamp, (x - shift) / scale, * (2^bit - 1), np.around, astype(uint32).
// IEEE round-to-nearest sub / div / mul (no FMA contraction), rintf = round-half-to-even.
__device__ __forceinline__ uint32_t quant_code(float x, float alpha, float shift, float scale, float levels) {
  const float xc = fminf(fmaxf(x, -alpha), alpha);
  const float r = __fdiv_rn(__fsub_rn(xc, shift), scale);
  return static_cast<uint32_t>(rintf(__fmul_rn(levels, r)));
}

// clamp_op.py:11-33: the Banner-2019 threshold from whole-tensor statistics accumulated in fp64.
//   Laplace (min < 0.2): torch.var(x, unbiased=False) -> fp32; GeLU: 2 * sum(x^2) / numel with fp32 squares.
__device__ __forceinline__ float clamp_alpha(int clamp, double gmin, double gs, double gss, double gss32, double cnt,
                                             float factor_laplace, float factor_gelu) {
  if (clamp == 0) return INFINITY;   // PE_CLAMP_NONE
  float variance, factor;
  const bool laplace = clamp == 2 || (clamp == 1 && gmin < 0.2);   // PE_CLAMP_LAPLACE / PE_CLAMP_AUTO
  if (laplace) {
    const double mean = gs / cnt;
    double var_d = gss / cnt - mean * mean;
    if (var_d < 0.0) var_d = 0.0;
    variance = static_cast<float>(var_d);
    factor = factor_laplace;
  } else {
    variance = __fdiv_rn(__fmul_rn(2.0f, static_cast<float>(gss32)), static_cast<float>(cnt));
    factor = factor_gelu;
  }
  return __fmul_rn(factor, __fsqrt_rn(__fmul_rn(0.5f, variance)));
}

// `_intmap2float` + `tensor_decode` (basic_op.py:146-163): float32(code / (2^bit - 1)) with a float64 divide, then
// * scale + shift as two fp32 roundings.
__device__ __forceinline__ float dequant_unit(uint32_t code, double levels) {
  return static_cast<float>(static_cast<double>(code) / levels);
}
__device__ __forceinline__ float dequant_value(float unit, float scale, float shift) {
  return __fadd_rn(__fmul_rn(unit, scale), shift);
}

}  // namespace pe
