// out = epilogue(A @ W^T + bias) on sm_100a tensor cores.
//
// One persistent CTA per SM, warp-specialised (Blackwell GEMM anatomy):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D boxes of A [128 x 64] and W [BN x 64] fp16 into a
//               ring of 128B-swizzled shared-memory stages, completion on `full` mbarriers
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage, accumulating
//               in TMEM; tcgen05.commit releases the stage (`empty`) and, after the last K block,
//               publishes the accumulator (`tmem_full`)
//   warps 2..9  epilogue: tcgen05.ld the 128 x BN fp32 accumulator out of TMEM (each warp owns a
//               32-lane quarter and half of the columns), apply bias / GELU / residual, store to HBM,
//               then hand the accumulator back (`tmem_empty`)
// Two accumulator stages (2 x 256 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// These GEMMs (M = 1.5k..6k tokens, N, K <= 4k) are L2-bandwidth bound with one 128 x BN tile per CTA: every
// SM re-reads A and W slices out of L2 (r01a ncu: tensor pipe 9-27 % active, ~5 TB/s of L2->SM traffic). So
// CTAs are launched as thread-block CLUSTERS of CM x CN: the CM CTAs that share a W tile each load 1/CM of
// it and TMA-MULTICAST it to the others, likewise the CN CTAs sharing an A tile, cutting L2 reads by up to
// CM (W) and CN (A). Every CTA's `full` barrier sees the bytes of its whole stage whoever issued them; a
// stage is released to all of its writers at once (tcgen05.commit multicast onto their `empty` barriers).
// CM, CN and BN (a runtime multiple of 16 in [16, 256]) are chosen per problem by a small cost model:
// waves x max(tensor time, L2 time).
//
// Replaces every nn.Linear on the reference path (see include/pipeedge_b200.h: pe_linear).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/pipeedge_b200.h"
#include "common.cuh"

namespace pe {

void count_launches(int n);
int require_sm100();

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 bytes = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = (2 + kNumEpiWarps) * 32;
constexpr int kMaxStages = 8;
constexpr int kPipeSmemBudget = 200 * 1024;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB

struct GemmParams {
  const float* bias;
  const float* resid;
  void* out;
  int m, n, k;
  int block_n;
  int stages;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int cm, cn;                      // cluster shape: CM CTAs along M share a W tile, CN along N share an A tile
  int num_super_m, num_super_n;    // cluster-level tiles (CM*128 x CN*BN)
  // Optional output-row remap (patch embedding writes token rows 1.. of each item and adds a
  // position table that repeats per item). rows_per_item == 0 disables it.
  int rows_per_item;   // GEMM rows per item
  int out_item_rows;   // output rows per item
  int out_row_offset;  // first output row of an item that GEMM row 0 maps to
  int resid_per_item;  // 1: resid is [out_item_rows, n] shared by all items
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int col0, const uint32_t (&r)[16]) {
  int out_row = row, resid_row = row;
  if (p.rows_per_item > 0) {
    const int item = row / p.rows_per_item;
    const int in_item = row - item * p.rows_per_item + p.out_row_offset;
    out_row = item * p.out_item_rows + in_item;
    resid_row = p.resid_per_item ? in_item : out_row;
  }
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
  const bool full = (col0 + 16 <= p.n) && ((p.n & 7) == 0);
  if (full) {
    if (p.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = __ldg(b4 + i);
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    }
    if (EPI == PE_EPI_GELU_F16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
    }
    if (EPI == PE_EPI_TANH_F32) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
    }
    if (EPI == PE_EPI_RESID_F32) {
      const float4* r4 = reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(resid_row) * p.n + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = r4[i];
        v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      }
    }
    if (EPI == PE_EPI_F16 || EPI == PE_EPI_GELU_F16) {
      __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(out_row) * p.n + col0;
      uint4 pk[2];
      __half2* h2 = reinterpret_cast<__half2*>(pk);
#pragma unroll
      for (int i = 0; i < 8; ++i) h2[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      reinterpret_cast<uint4*>(o)[0] = pk[0];
      reinterpret_cast<uint4*>(o)[1] = pk[1];
    } else {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(out_row) * p.n + col0);
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  } else {
    // ragged right edge or an N that breaks 16-byte alignment: guarded scalar path
    for (int i = 0; i < 16; ++i) {
      const int col = col0 + i;
      if (col >= p.n) break;
      float x = v[i];
      if (p.bias != nullptr) x += __ldg(p.bias + col);
      if (EPI == PE_EPI_GELU_F16) x = gelu_erf(x);
      if (EPI == PE_EPI_TANH_F32) x = tanhf(x);
      if (EPI == PE_EPI_RESID_F32) x += p.resid[static_cast<size_t>(resid_row) * p.n + col];
      if (EPI == PE_EPI_F16 || EPI == PE_EPI_GELU_F16) {
        reinterpret_cast<__half*>(p.out)[static_cast<size_t>(out_row) * p.n + col] = __float2half_rn(x);
      } else {
        reinterpret_cast<float*>(p.out)[static_cast<size_t>(out_row) * p.n + col] = x;
      }
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // SWIZZLE_128B tiles must start on 1024-byte boundaries
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.block_n) * (kBlockK * 2);

  // position inside the cluster: rank = m_rank + CM * n_rank
  const int csize = p.cm * p.cn;
  const uint32_t crank = csize > 1 ? cluster_ctarank() : 0u;
  const int m_rank = static_cast<int>(crank) % p.cm;
  const int n_rank = static_cast<int>(crank) / p.cm;
  uint16_t row_mask = 0, col_mask = 0;   // CTAs sharing my A tile (same m_rank) / my W tile (same n_rank)
  for (int c = 0; c < p.cn; ++c) row_mask |= static_cast<uint16_t>(1u << (m_rank + p.cm * c));
  for (int m = 0; m < p.cm; ++m) col_mask |= static_cast<uint16_t>(1u << (m + p.cm * n_rank));
  const uint16_t peers_mask = row_mask | col_mask;   // everyone who writes into my stages == everyone I write into

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], static_cast<uint32_t>(p.cm + p.cn - 1));
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();   // peers must not multicast into barriers that are not initialised yet
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int num_supers = p.num_super_m * p.num_super_n;
  const int cluster_id = static_cast<int>(blockIdx.x) / csize;
  const int num_clusters = static_cast<int>(gridDim.x) / csize;
  const int a_slice_rows = kBlockM / p.cn;       // my share of the A tile
  const int b_slice_rows = p.block_n / p.cm;     // my share of the W tile

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
        const int m_blk = (sup % p.num_super_m) * p.cm + m_rank;
        const int n_blk = (sup / p.num_super_m) * p.cn + n_rank;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);   // every CTA I multicast into has drained this stage
          uint8_t* sa = smem_gen + static_cast<size_t>(stage) * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          uint8_t* a_dst = sa + static_cast<size_t>(n_rank) * a_slice_rows * (kBlockK * 2);
          uint8_t* b_dst = sa + kABytes + static_cast<size_t>(m_rank) * b_slice_rows * (kBlockK * 2);
          const int a_row = m_blk * kBlockM + n_rank * a_slice_rows;
          const int b_row = n_blk * p.block_n + m_rank * b_slice_rows;
          if (p.cn > 1) tma_load_2d_multicast(a_dst, &tm_a, &full_bar[stage], kb * kBlockK, a_row, row_mask);
          else tma_load_2d(a_dst, &tm_a, &full_bar[stage], kb * kBlockK, a_row);
          if (p.cm > 1) tma_load_2d_multicast(b_dst, &tm_b, &full_bar[stage], kb * kBlockK, b_row, col_mask);
          else tma_load_2d(b_dst, &tm_b, &full_bar[stage], kb * kBlockK, b_row);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------ MMA issuer (single thread)
      const uint32_t idesc = umma_idesc_f16(kBlockM, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kAccStride);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_base + static_cast<uint32_t>(stage) * stage_bytes;
          const uint64_t desc_a = umma_desc_kmajor_sw128(a_addr);
          const uint64_t desc_b = umma_desc_kmajor_sw128(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advancing K by 16 fp16 = 32 bytes inside the swizzled row: +2 in 16-byte units
            umma_f16_ss(d_tmem, desc_a + static_cast<uint64_t>(k * 2), desc_b + static_cast<uint64_t>(k * 2), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          }
          // release the stage to every CTA that wrote part of it (incl. this one)
          if (csize > 1) umma_commit_multicast(&empty_bar[stage], peers_mask);
          else umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // -------------------------------------------------------------- epilogue warps
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    const int half = (warp - 2) >> 2;      // which half of the tile's 16-column chunks
    const int chunks = p.block_n >> 4;
    const int c_mid = (chunks + 1) >> 1;
    const int c_begin = half == 0 ? 0 : c_mid;
    const int c_end = half == 0 ? c_mid : chunks;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
      const int m_blk = (sup % p.num_super_m) * p.cm + m_rank;
      const int n_blk = (sup / p.num_super_m) * p.cn + n_rank;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const int row = m_blk * kBlockM + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * kAccStride);
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + static_cast<uint32_t>(c * 16), r);
        tmem_wait_ld();
        const int col0 = n_blk * p.block_n + c * 16;
        if (row < p.m && col0 < p.n) epilogue_chunk<EPI>(p, row, col0, r);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  // no CTA may leave while a peer can still multicast into its shared memory or arrive on its barriers
  if (csize > 1) cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || ptr == nullptr) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Row-major fp16 [rows, cols] -> boxes of [box_rows, 64] with the 128-byte swizzle; OOB reads give 0.
static int encode_f16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return PE_ERR_CUDA;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {kBlockK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for [%llu x %llu] box %u", static_cast<int>(rc),
              static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols), box_rows);
    return PE_ERR_CUDA;
  }
  return PE_OK;
}

struct GemmPlan {
  int cm, cn, bn;
};

static int max_clusters(int csize) {
  // co-scheduled clusters of 1 CTA/SM kernels: 148 SMs in GPCs of 16-20 (clusters never straddle a GPC)
  return csize == 1 ? kNumSMs : (csize == 2 ? kNumSMs / 2 : 132 / csize);
}

// Cost model (SM cycles): waves x max(tensor time of one tile, L2->SM time of one wave's operand reads) + launch.
// Tensor: 128 x BN x 16 MMA = BN/2 cycles. L2: ~3300 bytes per SM-cycle chip-wide (~6.4 TB/s), the rate the
// r01a kernels saturated at. A cluster reads (CM*128 + CN*BN) * K * 2 bytes per cluster tile.
GemmPlan plan_gemm(int m, int n, int k) {
  int forced[3] = {0, 0, 0};
  const char* env = getenv("PE_GEMM_FORCE");   // "CM,CN,BN": tuning / debugging only
  if (env != nullptr && sscanf(env, "%d,%d,%d", &forced[0], &forced[1], &forced[2]) == 3 && forced[0] > 0)
    return {forced[0], forced[1], forced[2]};
  GemmPlan best = {1, 1, 16};
  double best_cost = 1e30;
  const int shapes[5][2] = {{1, 1}, {2, 1}, {1, 2}, {2, 2}, {4, 1}};
  for (const auto& sh : shapes) {
    const int cm = sh[0], cn = sh[1], csize = cm * cn;
    for (int bn = 256; bn >= 16; bn -= 16) {
      if ((bn / cm) % 8 != 0 || (bn % cm) != 0) continue;   // W slices must keep whole 8-row swizzle atoms
      const long sm_ = (m + kBlockM * cm - 1) / (kBlockM * cm);
      const long sn_ = (n + bn * cn - 1) / (bn * cn);
      const long supers = sm_ * sn_;
      const long avail = max_clusters(csize);
      const long waves = (supers + avail - 1) / avail;
      const long active = supers < avail ? supers : avail;
      const double tile_cycles = static_cast<double>(k) * bn / 32.0;
      const double l2_cycles = static_cast<double>(cm * kBlockM + cn * bn) * k * 2.0 * active / 3300.0;
      const double cost = waves * (tile_cycles > l2_cycles ? tile_cycles : l2_cycles) + 2500.0 + 150.0 * csize;
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        best = {cm, cn, bn};
      }
    }
  }
  return best;
}

template <int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  // the attribute bounds DYNAMIC shared memory; static barriers live outside it (227 KiB total per CTA)
  const int max_smem = kPipeSmemBudget + 2048;
  if (!configured) {
    PE_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.block_n) * (kBlockK * 2);
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 1024;
  const int csize = p.cm * p.cn;
  const int supers = p.num_super_m * p.num_super_n;
  const int avail = max_clusters(csize);
  const int clusters = supers < avail ? supers : avail;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * csize));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(csize);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PE_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<EPI>, ta, tb, p));
  count_launches(1);
  return PE_OK;
}

int linear_impl(const void* a, const void* w, const void* bias, const void* resid, void* out, int m, int n, int k,
                int epilogue, int rows_per_item, int out_item_rows, int out_row_offset, int resid_per_item,
                cudaStream_t stream) {
  PE_REQUIRE(a && w && out, "pe_linear: null pointer");
  PE_REQUIRE(m > 0 && n > 0 && k > 0, "pe_linear: bad shape m=%d n=%d k=%d", m, n, k);
  PE_REQUIRE((k & 7) == 0, "pe_linear: k=%d must be a multiple of 8 (16-byte TMA row pitch)", k);
  PE_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
             "pe_linear: operands must be 16-byte aligned");
  PE_REQUIRE(epilogue != PE_EPI_RESID_F32 || resid != nullptr, "pe_linear: PE_EPI_RESID_F32 needs resid");
  int rc = require_sm100();
  if (rc != PE_OK) return rc;

  const GemmPlan plan = plan_gemm(m, n, k);
  PE_REQUIRE(plan.bn >= 16 && plan.bn <= 256 && plan.bn % 16 == 0 && plan.cm >= 1 && plan.cn >= 1 &&
                 plan.cm * plan.cn <= 8 && kBlockM % plan.cn == 0 && (plan.bn / plan.cm) % 8 == 0,
             "pe_linear: bad tile plan cm=%d cn=%d bn=%d", plan.cm, plan.cn, plan.bn);
  GemmParams p;
  p.bias = static_cast<const float*>(bias);
  p.resid = static_cast<const float*>(resid);
  p.out = out;
  p.m = m; p.n = n; p.k = k;
  p.block_n = plan.bn;
  p.cm = plan.cm;
  p.cn = plan.cn;
  const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.block_n) * (kBlockK * 2);
  int stages = kPipeSmemBudget / static_cast<int>(stage_bytes);
  p.stages = stages > kMaxStages ? kMaxStages : (stages < 2 ? 2 : stages);
  p.num_m_blocks = (m + kBlockM - 1) / kBlockM;
  p.num_n_blocks = (n + p.block_n - 1) / p.block_n;
  p.num_k_blocks = (k + kBlockK - 1) / kBlockK;
  p.num_super_m = (p.num_m_blocks + p.cm - 1) / p.cm;
  p.num_super_n = (p.num_n_blocks + p.cn - 1) / p.cn;
  p.rows_per_item = rows_per_item;
  p.out_item_rows = out_item_rows;
  p.out_row_offset = out_row_offset;
  p.resid_per_item = resid_per_item;

  CUtensorMap ta, tb;   // each CTA loads its SLICE of a tile: 128/CN rows of A, BN/CM rows of W
  rc = encode_f16_2d(&ta, a, static_cast<uint64_t>(m), static_cast<uint64_t>(k), static_cast<uint32_t>(kBlockM / p.cn));
  if (rc != PE_OK) return rc;
  rc = encode_f16_2d(&tb, w, static_cast<uint64_t>(n), static_cast<uint64_t>(k), static_cast<uint32_t>(p.block_n / p.cm));
  if (rc != PE_OK) return rc;

  switch (epilogue) {
    case PE_EPI_F16: return launch_gemm<PE_EPI_F16>(ta, tb, p, stream);
    case PE_EPI_GELU_F16: return launch_gemm<PE_EPI_GELU_F16>(ta, tb, p, stream);
    case PE_EPI_RESID_F32: return launch_gemm<PE_EPI_RESID_F32>(ta, tb, p, stream);
    case PE_EPI_F32: return launch_gemm<PE_EPI_F32>(ta, tb, p, stream);
    case PE_EPI_TANH_F32: return launch_gemm<PE_EPI_TANH_F32>(ta, tb, p, stream);
    default: set_error("pe_linear: unknown epilogue %d", epilogue); return PE_ERR_INVALID;
  }
}

// ------------------------------------------------------------------- debug reference (CUDA cores)
template <int EPI>
__global__ void gemm_simt_kernel(const __half* __restrict__ a, const __half* __restrict__ w,
                                 const float* __restrict__ bias, const float* resid, void* out, int m, int n, int k) {
  __shared__ float sa[16][17];
  __shared__ float sw[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < k; k0 += 16) {
    const int ar = blockIdx.y * 16 + ty, ac = k0 + tx;
    sa[ty][tx] = (ar < m && ac < k) ? __half2float(a[static_cast<size_t>(ar) * k + ac]) : 0.f;
    const int wr = blockIdx.x * 16 + ty, wc = k0 + tx;
    sw[ty][tx] = (wr < n && wc < k) ? __half2float(w[static_cast<size_t>(wr) * k + wc]) : 0.f;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) acc += sa[ty][kk] * sw[tx][kk];
    __syncthreads();
  }
  if (row >= m || col >= n) return;
  if (bias != nullptr) acc += bias[col];
  if (EPI == PE_EPI_GELU_F16) acc = gelu_erf(acc);
  if (EPI == PE_EPI_TANH_F32) acc = tanhf(acc);
  if (EPI == PE_EPI_RESID_F32) acc += resid[static_cast<size_t>(row) * n + col];
  if (EPI == PE_EPI_F16 || EPI == PE_EPI_GELU_F16) {
    reinterpret_cast<__half*>(out)[static_cast<size_t>(row) * n + col] = __float2half_rn(acc);
  } else {
    reinterpret_cast<float*>(out)[static_cast<size_t>(row) * n + col] = acc;
  }
}

int linear_simt_impl(const void* a, const void* w, const void* bias, const void* resid, void* out, int m, int n, int k,
                     int epilogue, cudaStream_t stream) {
  PE_REQUIRE(a && w && out && m > 0 && n > 0 && k > 0, "pe_debug_linear_simt: bad arguments");
  const dim3 grid((n + 15) / 16, (m + 15) / 16), block(16, 16);
  const __half* ha = static_cast<const __half*>(a);
  const __half* hw = static_cast<const __half*>(w);
  const float* fb = static_cast<const float*>(bias);
  const float* fr = static_cast<const float*>(resid);
  switch (epilogue) {
    case PE_EPI_F16: gemm_simt_kernel<PE_EPI_F16><<<grid, block, 0, stream>>>(ha, hw, fb, fr, out, m, n, k); break;
    case PE_EPI_GELU_F16: gemm_simt_kernel<PE_EPI_GELU_F16><<<grid, block, 0, stream>>>(ha, hw, fb, fr, out, m, n, k); break;
    case PE_EPI_RESID_F32: gemm_simt_kernel<PE_EPI_RESID_F32><<<grid, block, 0, stream>>>(ha, hw, fb, fr, out, m, n, k); break;
    case PE_EPI_F32: gemm_simt_kernel<PE_EPI_F32><<<grid, block, 0, stream>>>(ha, hw, fb, fr, out, m, n, k); break;
    case PE_EPI_TANH_F32: gemm_simt_kernel<PE_EPI_TANH_F32><<<grid, block, 0, stream>>>(ha, hw, fb, fr, out, m, n, k); break;
    default: set_error("pe_debug_linear_simt: unknown epilogue %d", epilogue); return PE_ERR_INVALID;
  }
  PE_CUDA(cudaGetLastError());
  count_launches(1);
  return PE_OK;
}

}  // namespace pe
