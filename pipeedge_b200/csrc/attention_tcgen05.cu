// Unmasked multi-head self-attention on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 64, S <= 256 keys
// (PE_ATTN_TCGEN05=0 selects the mma.sync kernel of attention.cu instead). DESIGN.md section 4.
//
// One CTA per (128 query rows, head, item); 9 warps: warps 0-7 do the softmax and the epilogue - TWO threads per query
// row (warps w and w + 4 may touch the same 32 TMEM lanes; each takes half of the key chunks and half of the output
// columns: with one thread per row the ~200 dependent exp2 / pack steps of a row were the kernel's critical path),
// warp 8 issues the TMA loads and the MMAs.
//   1. Q [128 x 64], K [kpad x 64], V [kpad x 64] arrive as three TMA boxes of 128-byte rows (SWIZZLE_128B) through a
//      3-D tensor map [item][row][3H]: rows past the item's S are out of bounds inside the item -> zero-filled.
//   2. S = Q K^T: 4 x tcgen05.mma (M=128, N=kpad, K=16), A and B K-major from shared memory, fp32 in TMEM columns
//      [0, kpad).
//   3. softmax: thread r reads its row with tcgen05.ld (32 columns at a time), two passes (max; exp2 with the scale
//      folded in + fp32 sum), keys >= S masked; P is packed to fp16 pairs and written back over S with tcgen05.st
//      (column j of P = keys 2j, 2j+1; chunk c writes columns [16c, 16c+16) after reading [32c, 32c+32): in place).
//   4. O += P V: kpad/16 x tcgen05.mma with A = P read FROM TMEM and B = V read MN-major from the same shared-memory
//      image TMA wrote (instruction descriptor b_major = 1; 8-key groups 1024 B apart, +2048 B per K=16 step): no
//      transpose of V anywhere. O: TMEM columns [128, 192) - S columns the softmax has finished with (P only occupies
//      [0, kpad/2)), so a CTA needs 256 TMEM columns and two CTAs share an SM: one's softmax overlaps the other's
//      loads and MMAs.
//   5. epilogue: tcgen05.ld O, scale by 1 / row sum, fp16, 128-byte row segments to ctx.
#include "../../include/pipeedge_b200.h"
#include "common.cuh"

namespace pe {

void count_launches(int n);

namespace {

constexpr int kD = 64;
constexpr int kRows = 128;          // query rows per CTA = TMEM lanes
constexpr int kOCol = 128;          // first TMEM column of O (>= kpad / 2: past the packed P)
constexpr int kTmem = 256;
constexpr int kSoftmaxWarps = 8;    // two threads per query row: warps w and w + 4 share a TMEM lane quarter
constexpr int kIssueWarp = kSoftmaxWarps;
constexpr int kThreads = (kSoftmaxWarps + 1) * 32;
constexpr int kMaxChunksPerThread = 4;   // kpad <= 256 -> 8 chunks of 32 keys, split over the two halves

__device__ __forceinline__ void tma_load_3d_addr(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_addr, int c0,
                                                 int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0],"
      " {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// named barrier among the softmax warps only (the issue warp never joins it)
__device__ __forceinline__ void softmax_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxWarps * 32) : "memory"); }

__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// 2^x for x <= 0 on the SFU (2 ulp): the probabilities are rounded to fp16 right after
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One 32-key chunk of a score row out of TMEM (the last chunk may hold only 16 keys: kpad % 32 == 16).
__device__ __forceinline__ void load_scores(uint32_t taddr, bool full, uint32_t pad_bits, uint32_t (&r)[32]) {
  if (full) {
    tmem_ld_32x32b_x32(taddr, r);
  } else {
    uint32_t h[16];
    tmem_ld_32x32b_x16(taddr, h);
#pragma unroll
    for (int i = 0; i < 16; ++i) { r[i] = h[i]; r[16 + i] = pad_bits; }
  }
  tmem_wait_ld();
}

__global__ void __launch_bounds__(kThreads, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                         __half* __restrict__ ctx, int tokens, int heads, int kpad, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_loaded, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_max[2][kRows], s_sum[2][kRows];
  pdl_launch_dependents();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kRows, head = blockIdx.y, item = blockIdx.z;
  const int hidden = heads * kD;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sq = smem_base;                                        // 128 rows x 128 B
  const uint32_t sk = sq + kRows * 128;                                 // kpad rows x 128 B
  const uint32_t sv = sk + static_cast<uint32_t>((kpad * 128 + 1023) & ~1023);

  if (threadIdx.x == kIssueWarp * 32) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(&bar_loaded, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, kSoftmaxWarps);   // one arrival per softmax warp
    mbar_init(&bar_o, 1);
    fence_barrier_init();
  }
  if (warp == kIssueWarp) {
    tmem_alloc(&tmem_base_smem, kTmem);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();                                                           // qkv is the predecessor's output

  if (warp == kIssueWarp) {
    // ------------------------------------------------------------ loads + MMA issue (converged, one lane issues)
    if (elect_one()) {
      const uint32_t bytes = static_cast<uint32_t>(kRows + 2 * kpad) * 128u;
      mbar_arrive_expect_tx_addr(smem_u32(&bar_loaded), bytes);
      tma_load_3d_addr(sq, &tm_q, smem_u32(&bar_loaded), head * kD, q0, item);
      tma_load_3d_addr(sk, &tm_kv, smem_u32(&bar_loaded), hidden + head * kD, 0, item);
      tma_load_3d_addr(sv, &tm_kv, smem_u32(&bar_loaded), 2 * hidden + head * kD, 0, item);
    }
    __syncwarp();
    mbar_wait(&bar_loaded, 0);
    tcgen05_fence_after();
    if (elect_one()) {
      const uint32_t idesc_s = idesc_f16(kRows, kpad, 0);
      const uint64_t dq = umma_desc_kmajor_sw128(sq), dk = umma_desc_kmajor_sw128(sk);
#pragma unroll
      for (int k = 0; k < kD / 16; ++k)     // +32 bytes inside the swizzled row per K = 16 step
        umma_f16_ss(tmem_base, dq + static_cast<uint64_t>(k * 2), dk + static_cast<uint64_t>(k * 2), idesc_s, k != 0);
      umma_commit(&bar_s);
    }
    __syncwarp();
    mbar_wait(&bar_p, 0);                                               // P is in TMEM
    tcgen05_fence_after();
    if (elect_one()) {
      const uint32_t idesc_o = idesc_f16(kRows, kD, 1);                 // B = V, MN-major
      const uint64_t dv = umma_desc_kmajor_sw128(sv);                   // same fields; the major lives in idesc
      const int ksteps = kpad >> 4;
      for (int k = 0; k < ksteps; ++k)      // A: 8 packed columns per 16 keys; B: 16 key rows = 2048 bytes
        umma_f16_ts(tmem_base + kOCol, tmem_base + static_cast<uint32_t>(k * 8), dv + static_cast<uint64_t>(k * 128),
                    idesc_o, k != 0);
      umma_commit(&bar_o);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ softmax + epilogue: two threads per query row
    // Warps w and w + 4 own the same 32 TMEM lanes (rows); `half` picks which key chunks (and which 32 output columns)
    // a thread handles. The row maximum and sum are combined through shared memory.
    const int quarter = warp & 3, half = warp >> 2;
    const int r_local = quarter * 32 + lane;
    const int row = q0 + r_local;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int nchunks = (kpad + 31) >> 5;                               // the last chunk may be half (kpad % 32 == 16)
    const int c_split = (nchunks + 1) >> 1;
    const int c_begin = half ? c_split : 0, c_end = half ? nchunks : c_split;
    mbar_wait(&bar_s, 0);
    tcgen05_fence_after();
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxChunksPerThread; ++i) {
      const int c = c_begin + i;
      if (c < c_end) {
        uint32_t r[32];
        load_scores(t_row + static_cast<uint32_t>(c * 32), c * 32 + 32 <= kpad, 0xff800000u, r);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
    }
    s_max[half][r_local] = mx;
    softmax_barrier();
    mx = fmaxf(s_max[0][r_local], s_max[1][r_local]);
    const float msc = mx * scale_log2e;
    float sum = 0.f;
    uint32_t pk[kMaxChunksPerThread][16];
#pragma unroll
    for (int i = 0; i < kMaxChunksPerThread; ++i) {
      const int c = c_begin + i;
      if (c < c_end) {
        uint32_t r[32];
        load_scores(t_row + static_cast<uint32_t>(c * 32), c * 32 + 32 <= kpad, 0u, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int key = c * 32 + 2 * j;
          const float p0 = key < tokens ? fast_exp2(__uint_as_float(r[2 * j]) * scale_log2e - msc) : 0.f;
          const float p1 = key + 1 < tokens ? fast_exp2(__uint_as_float(r[2 * j + 1]) * scale_log2e - msc) : 0.f;
          const __half2 h2 = __floats2half2_rn(p0, p1);
          sum += p0 + p1;
          pk[i][j] = *reinterpret_cast<const uint32_t*>(&h2);
        }
      }
    }
    s_sum[half][r_local] = sum;
    // P overwrites S in place (packed pairs: chunk c -> columns [16c, 16c + 16)): every thread must have finished
    // READING its score columns before anyone writes
    softmax_barrier();
#pragma unroll
    for (int i = 0; i < kMaxChunksPerThread; ++i) {
      const int c = c_begin + i;
      if (c < c_end) tmem_st_32x32b_x16(t_row + static_cast<uint32_t>(c * 16), pk[i]);
    }
    tmem_wait_st();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_p);
    mbar_wait(&bar_o, 0);
    tcgen05_fence_after();
    const float inv = 1.0f / (s_sum[0][r_local] + s_sum[1][r_local]);
    __half* out = ctx + (static_cast<size_t>(item) * tokens + row) * hidden + static_cast<size_t>(head) * kD + half * 32;
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(kOCol + half * 32), r);
    tmem_wait_ld();
    if (row < tokens) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 v;
        __half2* h2 = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          h2[i] = __floats2half2_rn(__uint_as_float(r[8 * j + 2 * i]) * inv, __uint_as_float(r[8 * j + 2 * i + 1]) * inv);
        *reinterpret_cast<uint4*>(out + j * 8) = v;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kIssueWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmem);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// qkv viewed as [item][row][3H] fp16; box = 64 columns x `box_rows` rows of one item.
int encode_qkv_3d(CUtensorMap* map, const void* qkv, int batch, int tokens, int hidden3, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is unavailable");
    return PE_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(hidden3), static_cast<cuuint64_t>(tokens),
                              static_cast<cuuint64_t>(batch)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(hidden3) * 2,
                                 static_cast<cuuint64_t>(hidden3) * 2 * static_cast<cuuint64_t>(tokens)};
  const cuuint32_t box[3] = {64u, static_cast<cuuint32_t>(box_rows), 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (qkv) failed (CUresult %d)", static_cast<int>(rc));
    return PE_ERR_CUDA;
  }
  return PE_OK;
}

}  // namespace

// Returns PE_ERR_INVALID (without an error message change) when the shape is outside what this kernel handles, so the
// caller can fall back to the mma.sync kernel.
int attention_tcgen05_impl(const void* qkv, void* ctx, int batch, int tokens, int heads, int head_dim,
                           cudaStream_t stream) {
  if (head_dim != kD || tokens > 256 || tokens < 1 || ((heads * kD * 3 * 2) & 15) != 0) return PE_ERR_INVALID;
  const int kpad = (tokens + 15) & ~15;
  CUtensorMap tm_q, tm_kv;
  int rc = encode_qkv_3d(&tm_q, qkv, batch, tokens, 3 * heads * kD, kRows);
  if (rc != PE_OK) return rc;
  rc = encode_qkv_3d(&tm_kv, qkv, batch, tokens, 3 * heads * kD, kpad);
  if (rc != PE_OK) return rc;
  const size_t smem = static_cast<size_t>(kRows) * 128 + 2 * static_cast<size_t>((kpad * 128 + 1023) & ~1023) + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    PE_CUDA(cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    configured = smem;
  }
  const dim3 grid((tokens + kRows - 1) / kRows, heads, batch);
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(kD));
  PE_CUDA(launch_pdl(attention_tcgen05_kernel, grid, dim3(kThreads), smem, stream, tm_q, tm_kv,
                     static_cast<__half*>(ctx), tokens, heads, kpad, scale_log2e));
  count_launches(1);
  return PE_OK;
}

}  // namespace pe
