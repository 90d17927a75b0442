// QuantPipe on the device: Banner-2019 clamp + per-item affine quantise + LSB-first bit-pack, and the
// inverse. HBM-bound byte/integer work: two streaming passes over the fp32 activation (statistics,
// then quantise+pack) and one over the codes. Bit-exactness against the NumPy reference comes from
// using IEEE round-to-nearest sub/div/mul intrinsics (no FMA contraction), rintf (half-to-even) and
// fp64 accumulation for the variance that sets the clamp threshold.
// Replaces runtime.py:73-119 + quantization/basic_op.py + clamp_op.py (see the header for the mapping).
#include <math.h>

#include "../../include/pipeedge_b200.h"
#include "common.cuh"
#include "quant_dev.cuh"

namespace pe {

void count_launches(int n);

constexpr int kQStatThreads = 256;

// ------------------------------------------------------------------ Lambert W (host, fp64)
static double lambert_w0(double z) {
  // principal branch for z >= 0: Halley iterations from the asymptotic guess
  if (z == 0.0) return 0.0;
  double w = z < 3.0 ? 0.5 * z / (1.0 + 0.5 * z) + 0.3 : log(z) - log(log(z));
  for (int it = 0; it < 100; ++it) {
    const double ew = exp(w);
    const double f = w * ew - z;
    const double step = f / (ew * (w + 1.0) - (w + 2.0) * f / (2.0 * w + 2.0));
    w -= step;
    if (fabs(step) <= 1e-16 * fabs(w)) break;
  }
  return w;
}

float clamp_factor(int bit, int gelu) {
  // clamp_op.py:6-8,22-24: lambertw(3 * 4^bit) (Laplace) or lambertw(3 * 4^(bit+1)) (GeLU), cast to f32
  return static_cast<float>(lambert_w0(3.0 * pow(4.0, static_cast<double>(bit + (gelu ? 1 : 0)))));
}

// ------------------------------------------------------------------ pass 1: statistics
__global__ void __launch_bounds__(kQStatThreads)
quant_stats_kernel(const float* __restrict__ x, size_t n, int chunks, double* __restrict__ partials) {
  const int item = blockIdx.y, chunk = blockIdx.x;
  const float* xi = x + static_cast<size_t>(item) * n;
  // chunk boundaries on multiples of 4 elements so the body can use float4 when xi is 16-byte aligned
  const size_t per = ((n + chunks - 1) / chunks + 3) & ~static_cast<size_t>(3);
  const size_t begin = static_cast<size_t>(chunk) * per;
  const size_t end = begin + per < n ? begin + per : n;
  float mn = INFINITY, mx = -INFINITY;
  double s = 0.0, ss = 0.0, ss32 = 0.0;
  const bool vec = ((reinterpret_cast<uintptr_t>(xi) & 15) == 0);
  if (begin < end) {
    if (vec) {
      const size_t nv = (end - begin) >> 2;
      const float4* x4 = reinterpret_cast<const float4*>(xi + begin);
      for (size_t i = threadIdx.x; i < nv; i += kQStatThreads) {
        const float4 v = x4[i];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mn = fminf(mn, e[j]);
          mx = fmaxf(mx, e[j]);
          const double d = static_cast<double>(e[j]);
          s += d;
          ss += d * d;
          ss32 += static_cast<double>(__fmul_rn(e[j], e[j]));
        }
      }
      for (size_t i = begin + (nv << 2) + threadIdx.x; i < end; i += kQStatThreads) {
        const float e = xi[i];
        mn = fminf(mn, e); mx = fmaxf(mx, e);
        const double d = static_cast<double>(e);
        s += d; ss += d * d; ss32 += static_cast<double>(__fmul_rn(e, e));
      }
    } else {
      for (size_t i = begin + threadIdx.x; i < end; i += kQStatThreads) {
        const float e = xi[i];
        mn = fminf(mn, e); mx = fmaxf(mx, e);
        const double d = static_cast<double>(e);
        s += d; ss += d * d; ss32 += static_cast<double>(__fmul_rn(e, e));
      }
    }
  }
  mn = warp_reduce(mn, [](float a, float b) { return fminf(a, b); });
  mx = warp_max(mx);
  s = warp_sum_d(s); ss = warp_sum_d(ss); ss32 = warp_sum_d(ss32);
  __shared__ double red[kQStatThreads / 32][kQPartialDoubles];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = mn; red[warp][1] = mx; red[warp][2] = s; red[warp][3] = ss; red[warp][4] = ss32;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double o0 = red[0][0], o1 = red[0][1], o2 = red[0][2], o3 = red[0][3], o4 = red[0][4];
    for (int w = 1; w < kQStatThreads / 32; ++w) {
      o0 = fmin(o0, red[w][0]); o1 = fmax(o1, red[w][1]);
      o2 += red[w][2]; o3 += red[w][3]; o4 += red[w][4];
    }
    double* p = partials + (static_cast<size_t>(item) * chunks + chunk) * kQPartialDoubles;
    p[0] = o0; p[1] = o1; p[2] = o2; p[3] = o3; p[4] = o4;
  }
}

// ------------------------------------------------------------------ pass 1b: thresholds (one block)
struct QuantHeader {  // lives at the start of the workspace; read by the pack kernel
  float alpha;
  float pad[3];
};

__global__ void quant_finalize_kernel(double* partials, int items, int chunks, size_t n, int clamp,
                                      float factor_laplace, float factor_gelu, QuantHeader* hdr, float* item_min,
                                      float* item_max, float* scale, float* shift, float* alpha_out) {
  // fixed-order reductions -> run-to-run deterministic: one thread per item over its chunks, then
  // thread 0 over the items (the per-item sums are parked in the partials' first chunk slot)
  __shared__ float s_alpha;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    double mn = INFINITY, mx = -INFINITY, s = 0.0, ss = 0.0, ss32 = 0.0;
    for (int c = 0; c < chunks; ++c) {
      const double* p = partials + (static_cast<size_t>(i) * chunks + c) * kQPartialDoubles;
      mn = fmin(mn, p[0]); mx = fmax(mx, p[1]);
      s += p[2]; ss += p[3]; ss32 += p[4];
    }
    item_min[i] = static_cast<float>(mn);
    item_max[i] = static_cast<float>(mx);
    double* p0 = partials + static_cast<size_t>(i) * chunks * kQPartialDoubles;
    p0[0] = mn; p0[2] = s; p0[3] = ss; p0[4] = ss32;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double gmin = INFINITY, gs = 0.0, gss = 0.0, gss32 = 0.0;
    for (int i = 0; i < items; ++i) {
      const double* p0 = partials + static_cast<size_t>(i) * chunks * kQPartialDoubles;
      gmin = fmin(gmin, p0[0]);
      gs += p0[2]; gss += p0[3]; gss32 += p0[4];
    }
    // clamp_op.py:11-33 (see clamp_alpha in quant_dev.cuh; the fused send kernel of link.cu shares it)
    const float alpha = clamp_alpha(clamp, gmin, gs, gss, gss32, static_cast<double>(items) * static_cast<double>(n),
                                    factor_laplace, factor_gelu);
    s_alpha = alpha;
    hdr->alpha = alpha;
    if (alpha_out != nullptr) *alpha_out = alpha;
  }
  __syncthreads();
  const float alpha = s_alpha;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    // clamp is monotonic: min/max of the clamped item = clamped min/max (basic_op.py:127-129)
    const float lo = fminf(fmaxf(item_min[i], -alpha), alpha);
    const float hi = fminf(fmaxf(item_max[i], -alpha), alpha);
    shift[i] = lo;
    scale[i] = __fsub_rn(hi, lo);
  }
}

// ------------------------------------------------------------------ pass 2: quantise + pack
// Fast path: bit in {2,4,8,16} and n % 16 == 0. One thread = 16 consecutive elements (4 x float4 in,
// 16*bit/32 words out as one vector store).
template <int BIT>
__global__ void __launch_bounds__(256)
quant_pack16_kernel(const float* __restrict__ x, size_t n, const QuantHeader* __restrict__ hdr,
                    const float* __restrict__ scale, const float* __restrict__ shift, uint32_t* __restrict__ codes,
                    size_t words_per_item) {
  constexpr int kWords = 16 * BIT / 32;
  const int item = blockIdx.y;
  const size_t groups = n >> 4;
  const float alpha = hdr->alpha;
  const float sh = shift[item], sc = scale[item];
  const float levels = static_cast<float>((1u << BIT) - 1u);
  const float4* xi = reinterpret_cast<const float4*>(x + static_cast<size_t>(item) * n);
  uint32_t* ci = codes + static_cast<size_t>(item) * words_per_item;
  for (size_t g = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; g < groups;
       g += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = xi[g * 4 + j];
    uint32_t q[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[4 * j + 0] = quant_code(v[j].x, alpha, sh, sc, levels);
      q[4 * j + 1] = quant_code(v[j].y, alpha, sh, sc, levels);
      q[4 * j + 2] = quant_code(v[j].z, alpha, sh, sc, levels);
      q[4 * j + 3] = quant_code(v[j].w, alpha, sh, sc, levels);
    }
    uint32_t w[kWords];
    constexpr int kRatio = 32 / BIT;
#pragma unroll
    for (int k = 0; k < kWords; ++k) {
      uint32_t acc = 0;
#pragma unroll
      for (int j = 0; j < kRatio; ++j) acc |= q[k * kRatio + j] << (j * BIT);
      w[k] = acc;
    }
    uint32_t* dst = ci + g * kWords;
    if (kWords == 1) dst[0] = w[0];
    else if (kWords == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    else {
#pragma unroll
      for (int k = 0; k < kWords; k += 4) *reinterpret_cast<uint4*>(dst + k) = make_uint4(w[k], w[k + 1], w[k + 2], w[k + 3]);
    }
  }
}

// Generic path: any bit in [1,16], any n. One thread = one output word (floor(32/bit) codes).
__global__ void __launch_bounds__(256)
quant_pack_generic_kernel(const float* __restrict__ x, size_t n, int bit, const QuantHeader* __restrict__ hdr,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          uint32_t* __restrict__ codes, size_t words_per_item) {
  const int item = blockIdx.y;
  const int ratio = 32 / bit;
  const float alpha = hdr->alpha;
  const float sh = shift[item], sc = scale[item];
  const float levels = static_cast<float>((1u << bit) - 1u);
  const float* xi = x + static_cast<size_t>(item) * n;
  uint32_t* ci = codes + static_cast<size_t>(item) * words_per_item;
  for (size_t w = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; w < words_per_item;
       w += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint32_t acc = 0;
    const size_t e0 = w * ratio;
    for (int j = 0; j < ratio; ++j) {
      const size_t e = e0 + j;
      if (e < n) acc |= quant_code(xi[e], alpha, sh, sc, levels) << (j * bit);  // zero-padded tail
    }
    ci[w] = acc;
  }
}

// ------------------------------------------------------------------ decode
__global__ void __launch_bounds__(256)
quant_decode_kernel(const uint32_t* __restrict__ codes, size_t n, int bit, size_t words_per_item,
                    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out) {
  // _intmap2float: float32(code / (2^bit - 1)) with a float64 divide. A shared-memory table of the
  // 2^bit possible values replaces the divide for bit <= 12.
  extern __shared__ float lut[];
  const int item = blockIdx.y;
  const int ratio = 32 / bit;
  const uint32_t mask = (1u << bit) - 1u;
  const double levels = static_cast<double>(mask);
  const bool use_lut = bit <= 12;
  if (use_lut) {
    for (uint32_t c = threadIdx.x; c <= mask; c += blockDim.x) lut[c] = dequant_unit(c, levels);
    __syncthreads();
  }
  const float sc = scale[item], sh = shift[item];
  const uint32_t* ci = codes + static_cast<size_t>(item) * words_per_item;
  float* oi = out + static_cast<size_t>(item) * n;
  for (size_t w = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; w < words_per_item;
       w += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t word = ci[w];
    const size_t e0 = w * ratio;
    for (int j = 0; j < ratio; ++j) {
      const size_t e = e0 + j;
      if (e >= n) break;
      const uint32_t c = (word >> (j * bit)) & mask;
      const float v = use_lut ? lut[c] : dequant_unit(c, levels);
      oi[e] = dequant_value(v, sc, sh);  // basic_op.py:163: two fp32 roundings
    }
  }
}

// ------------------------------------------------------------------ host entry points
static int pick_chunks(int items, size_t n) {
  long want = (2L * kNumSMs + items - 1) / items;
  long by_size = static_cast<long>((n + 4095) / 4096);
  long c = want < by_size ? want : by_size;
  if (c < 1) c = 1;
  if (c > kQMaxChunks) c = kQMaxChunks;
  return static_cast<int>(c);
}

size_t quant_words(size_t n, int bit) {
  if (bit < 1 || bit > 32) return 0;
  const size_t ratio = static_cast<size_t>(32 / bit);
  return (n + ratio - 1) / ratio;
}

size_t quant_workspace_bytes(int items, size_t /*n*/) {
  // header + per-item min/max + partials
  return 256 + static_cast<size_t>(items) * 2 * sizeof(float) + 256 +
         static_cast<size_t>(items) * kQMaxChunks * kQPartialDoubles * sizeof(double);
}

// statistics + thresholds: fills scale/shift (and *alpha) and leaves the QuantHeader at the start of `work`
int quant_stats_impl(const void* x, int items, size_t n, int bit, int clamp, void* scale, void* shift, void* alpha,
                     void* work, cudaStream_t stream) {
  PE_REQUIRE(x && scale && shift && work, "pe_quant: null pointer");
  PE_REQUIRE(items > 0 && n > 0, "pe_quant: empty tensor");
  PE_REQUIRE(bit >= 1 && bit <= 16, "pe_quant: bit=%d outside [1,16]", bit);
  PE_REQUIRE(clamp >= PE_CLAMP_NONE && clamp <= PE_CLAMP_GELU, "pe_quant: bad clamp mode %d", clamp);
  PE_REQUIRE((reinterpret_cast<uintptr_t>(work) & 255) == 0, "pe_quant: workspace must be 256-byte aligned");
  uint8_t* wsp = static_cast<uint8_t*>(work);
  QuantHeader* hdr = reinterpret_cast<QuantHeader*>(wsp);
  float* item_min = reinterpret_cast<float*>(wsp + 256);
  float* item_max = item_min + items;
  const size_t off = (256 + static_cast<size_t>(items) * 2 * sizeof(float) + 255) & ~static_cast<size_t>(255);
  double* partials = reinterpret_cast<double*>(wsp + off);
  const int chunks = pick_chunks(items, n);
  const float* xf = static_cast<const float*>(x);

  quant_stats_kernel<<<dim3(chunks, items), kQStatThreads, 0, stream>>>(xf, n, chunks, partials);
  PE_CUDA(cudaGetLastError());
  quant_finalize_kernel<<<1, 128, 0, stream>>>(partials, items, chunks, n, clamp, clamp_factor(bit, 0),
                                               clamp_factor(bit, 1), hdr, item_min, item_max,
                                               static_cast<float*>(scale), static_cast<float*>(shift),
                                               static_cast<float*>(alpha));
  PE_CUDA(cudaGetLastError());
  count_launches(2);
  return PE_OK;
}

int quant_encode_impl(const void* x, int items, size_t n, int bit, int clamp, void* codes, void* scale, void* shift,
                      void* alpha, void* work, cudaStream_t stream) {
  PE_REQUIRE(codes != nullptr, "pe_quant_encode: null pointer");
  PE_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0, "pe_quant_encode: codes must be 16-byte aligned");
  const int rc = quant_stats_impl(x, items, n, bit, clamp, scale, shift, alpha, work, stream);
  if (rc != PE_OK) return rc;
  const QuantHeader* hdr = reinterpret_cast<const QuantHeader*>(work);
  const float* xf = static_cast<const float*>(x);
  const size_t words = quant_words(n, bit);
  uint32_t* cw = static_cast<uint32_t*>(codes);
  const bool fast = (n % 16 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                    (bit == 2 || bit == 4 || bit == 8 || bit == 16);
  if (fast) {
    const size_t groups = n / 16;
    size_t bx = (groups + 255) / 256;
    const size_t cap = static_cast<size_t>(kNumSMs) * 8 / static_cast<size_t>(items) + 1;
    if (bx > cap) bx = cap;
    const dim3 grid(static_cast<unsigned>(bx), items);
    const float* sc = static_cast<const float*>(scale);
    const float* sh = static_cast<const float*>(shift);
    switch (bit) {
      case 2: quant_pack16_kernel<2><<<grid, 256, 0, stream>>>(xf, n, hdr, sc, sh, cw, words); break;
      case 4: quant_pack16_kernel<4><<<grid, 256, 0, stream>>>(xf, n, hdr, sc, sh, cw, words); break;
      case 8: quant_pack16_kernel<8><<<grid, 256, 0, stream>>>(xf, n, hdr, sc, sh, cw, words); break;
      default: quant_pack16_kernel<16><<<grid, 256, 0, stream>>>(xf, n, hdr, sc, sh, cw, words); break;
    }
  } else {
    size_t bx = (words + 255) / 256;
    const size_t cap = static_cast<size_t>(kNumSMs) * 8 / static_cast<size_t>(items) + 1;
    if (bx > cap) bx = cap;
    quant_pack_generic_kernel<<<dim3(static_cast<unsigned>(bx), items), 256, 0, stream>>>(
        xf, n, bit, hdr, static_cast<const float*>(scale), static_cast<const float*>(shift), cw, words);
  }
  PE_CUDA(cudaGetLastError());
  count_launches(1);
  return PE_OK;
}

int quant_decode_impl(const void* codes, int items, size_t n, int bit, const void* scale, const void* shift, void* out,
                      cudaStream_t stream) {
  PE_REQUIRE(codes && scale && shift && out, "pe_quant_decode: null pointer");
  PE_REQUIRE(items > 0 && n > 0, "pe_quant_decode: empty tensor");
  PE_REQUIRE(bit >= 1 && bit <= 16, "pe_quant_decode: bit=%d outside [1,16]", bit);
  const size_t words = quant_words(n, bit);
  size_t bx = (words + 255) / 256;
  const size_t cap = static_cast<size_t>(kNumSMs) * 8 / static_cast<size_t>(items) + 1;
  if (bx > cap) bx = cap;
  const size_t smem = bit <= 12 ? (static_cast<size_t>(1) << bit) * sizeof(float) : 0;
  quant_decode_kernel<<<dim3(static_cast<unsigned>(bx), items), 256, smem, stream>>>(
      static_cast<const uint32_t*>(codes), n, bit, words, static_cast<const float*>(scale),
      static_cast<const float*>(shift), static_cast<float*>(out));
  PE_CUDA(cudaGetLastError());
  count_launches(1);
  return PE_OK;
}

}  // namespace pe
