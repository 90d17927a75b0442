// out = epilogue(A @ W^T + bias) on sm_100a tensor cores.
//
// One persistent CTA per SM, warp-specialised (Blackwell GEMM anatomy):
//   warp 16     TMA producer: cp.async.bulk.tensor 2D boxes of A [128 x 64] and W [BN x 64] fp16 into a
//               ring of 128B-swizzled shared-memory stages, completion on `full` mbarriers
//   warp 17     MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage, accumulating
//               in TMEM; tcgen05.commit releases the stage (`empty`) and, after the last K block,
//               publishes the accumulator (`tmem_full`)
//   warps 0..15 epilogue: tcgen05.ld the 128 x BN fp32 accumulator out of TMEM (each warp owns a
//               32-lane quarter and half of the columns), apply bias / GELU / residual, store to HBM,
//               then hand the accumulator back (`tmem_empty`)
// The two single-thread roles get the HIGHEST warp ids: the warp scheduler favours higher ids, and with the
// roles on warps 0/1 the waiting epilogue warps starved them (r01a: ~650 cycles per K block whatever the tile).
// Two accumulator stages (2 x 256 TMEM columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// These GEMMs (M = 1.5k..6k tokens, N, K <= 4k) are L2-bandwidth bound with one 128 x BN tile per CTA: every
// SM re-reads A and W slices out of L2 (r01a ncu: tensor pipe 9-27 % active, ~5 TB/s of L2->SM traffic). So
// CTAs are launched as thread-block CLUSTERS of CM x CN: the CM CTAs that share a W tile each load 1/CM of
// it and TMA-MULTICAST it to the others, likewise the CN CTAs sharing an A tile, cutting L2 reads by up to
// CM (W) and CN (A). Every CTA's `full` barrier sees the bytes of its whole stage whoever issued them; a
// stage is released to all of its writers at once (tcgen05.commit multicast onto their `empty` barriers).
// CM, CN and BN (a runtime multiple of 32 in [32, 256]) are chosen per problem by a small cost model:
// waves x max(tensor time, L2 time).
//
// Replaces every nn.Linear on the reference path (see include/pipeedge_b200.h: pe_linear).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/pipeedge_b200.h"
#include "common.cuh"
#include "ln_dev.cuh"

namespace pe {

void count_launches(int n);
int require_sm100();

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 fp16 = 128 bytes = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kNumEpiWarps = 16;   // 4 TMEM lane quarters x 4 column groups: the epilogue is issue-latency bound
constexpr int kGemmThreads = (3 + kNumEpiWarps) * 32;
constexpr int kProducerWarp = kNumEpiWarps;      // warp 16: activations (A) + arms the full barriers
constexpr int kMmaWarp = kNumEpiWarps + 1;       // warp 17
constexpr int kWeightWarp = kNumEpiWarps + 2;    // warp 18: weights (W); never needs the predecessor's data
constexpr int kMaxStages = 8;
constexpr int kTraceSlots = 32;   // pe_debug_gemm_trace: clock64 stamps per CTA
constexpr int kPipeSmemBudget = 144 * 1024;   // operand ring; the epilogue staging (80 KiB) follows it
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB

struct GemmParams {
  const float* bias;
  const float* resid;
  void* out;
  int m, n, k;
  int block_n;
  int stages;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  long long* trace;                // debug: per-CTA clock64 timeline (8 slots) or nullptr
  int tma_store;                   // 1: epilogue stages 32-row boxes in swizzled smem and TMA-stores them
  int debug_mode;                  // debug (wrong results!): bit0 = skip the MMAs, bit1 = skip all TMA loads
  int cm, cn;                      // cluster shape: CM CTAs along M share a W tile, CN along N share an A tile
  int num_super_m, num_super_n;    // cluster-level tiles (CM*128 x CN*BN)
  // Optional output-row remap (patch embedding writes token rows 1.. of each item and adds a
  // position table that repeats per item). rows_per_item == 0 disables it.
  int rows_per_item;   // GEMM rows per item
  int out_item_rows;   // output rows per item
  int out_row_offset;  // first output row of an item that GEMM row 0 maps to
  int resid_per_item;  // 1: resid is [out_item_rows, n] shared by all items
  int stg_offset;      // byte offset of the epilogue staging from the ring base: 0 = aliases the ring (one tile per CTA)
  int static_w;        // 1: W is not written by anything still pending on the stream -> may be read before pdl_wait()
  // PE_EPI_RESID_LN: v = A W^T + bias + resid, then LayerNorm(v) over the full row (the CN CTAs of a cluster own a row)
  const float* ln_gamma;
  const float* ln_beta;
  float ln_eps;
  int ln_off;          // byte offset (from the ring base) of the row-statistics exchange area
  int f32_is_ln;       // the fp32 output (tm_out) carries LayerNorm(v) (post-LN models) instead of v (pre-LN models)
  int has_f32, has_f16;
};

constexpr int kLnAreaBytes = 8192;   // [4][128] float2 chunk statistics + [2][128] float2 CTA statistics (double-buffered)

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// Epilogue GELU (exact-erf form, as the reference's nn.GELU): |error| <= 7e-7 absolute, 4e-6 relative - far below the
// fp16 rounding of the value it feeds (tests/test_kernels_gpu.py compares with torch's erf GELU).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  // gelu(x) = x Phi(x); with E(a) = Phi(-a) = erfc(a / sqrt2) / 2 (a = |x|):  gelu(x) = max(x, 0) - |x| E(|x|).
  // log2 E is smooth: a degree-7 fit on [0, 6] (Chebyshev nodes, scripts/fit_gelu.py) is within 5.7e-6, i.e. E is
  // relatively accurate to 4e-6 INCLUDING the tail that the negative half of GELU lives on. One MUFU (ex2) and 11
  // FMA-pipe instructions; the A&S 7.1.26 form used before needed rcp + ex2 and 13 (the epilogue is MUFU/issue bound).
  const float a = fminf(fabsf(x), 6.0f);
  float pl = fmaf(-1.889626233e-06f, a, 6.268140101e-05f);
  pl = fmaf(pl, a, -9.388679333e-04f);
  pl = fmaf(pl, a, 8.539461332e-03f);
  pl = fmaf(pl, a, -5.402068712e-02f);
  pl = fmaf(pl, a, -4.584097768e-01f);
  pl = fmaf(pl, a, -1.151269147e+00f);
  pl = fmaf(pl, a, -9.999943403e-01f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(pl));
  return fmaxf(x, 0.0f) - fabsf(x * e);
}

constexpr int kStgLd = 36;                        // floats per staged row: 144 B keeps 16-byte accesses conflict-free
constexpr int kStgFloatsPerWarp = 1280;           // 5 KiB per epilogue warp: a 32 x 36 fp32 chunk, or a 1 KiB-aligned 4 KiB TMA box
constexpr int kStgBytes = kNumEpiWarps * kStgFloatsPerWarp * 4;

__device__ __forceinline__ void map_rows(const GemmParams& p, int row, int& out_row, int& resid_row) {
  out_row = row;
  resid_row = row;
  if (p.rows_per_item > 0) {
    const int item = row / p.rows_per_item;
    const int in_item = row - item * p.rows_per_item + p.out_row_offset;
    out_row = item * p.out_item_rows + in_item;
    resid_row = p.resid_per_item ? in_item : out_row;
  }
}

template <int EPI>
__device__ __forceinline__ float epi_act(float x) {
  if (EPI == PE_EPI_GELU_F16) return gelu_erf_fast(x);
  if (EPI == PE_EPI_TANH_F32) return tanhf(x);
  return x;
}

// One 32-row x 32-column chunk of the accumulator, already staged row-major in `stg` (this warp's private
// shared memory): re-read it so that consecutive lanes cover consecutive columns of a row, then apply the
// epilogue and store. TMEM hands each lane one ROW (tcgen05.ld 32x32b), which would make every global access
// touch 32 different lines (r01a: the epilogue took longer than the main loop); after the transpose each
// instruction covers whole 64-128 byte row segments and the residual loads of a chunk are all in flight at once.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk32(const GemmParams& p, const float* stg, int row0, int col0, int lane) {
  constexpr bool kHalfOut = (EPI == PE_EPI_F16 || EPI == PE_EPI_GELU_F16);
  const bool vec_ok = (p.n & 7) == 0;
  if (kHalfOut) {
    // lane -> (row = it*8 + lane/4, 8 columns starting at (lane%4)*8): 16-byte fp16 stores, 64 B per row
    const int cc = (lane & 3) * 8;
    const int col = col0 + cc;
    const bool fast = vec_ok && col + 8 <= p.n;
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (p.bias != nullptr && fast) {   // this lane's 8 columns are the same for all of its rows
      b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int rr = it * 8 + (lane >> 2);
      const int row = row0 + rr;
      if (row >= p.m || col >= p.n) continue;
      int out_row, resid_row;
      map_rows(p, row, out_row, resid_row);
      const float4 a = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cc);
      const float4 b = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cc + 4);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(out_row) * p.n + col;
      if (fast) {
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        uint4 pk;
        __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) h2[i] = __floats2half2_rn(epi_act<EPI>(v[2 * i]), epi_act<EPI>(v[2 * i + 1]));
        *reinterpret_cast<uint4*>(o) = pk;
      } else {
        for (int i = 0; i < 8 && col + i < p.n; ++i) {
          float x = v[i];
          if (p.bias != nullptr) x += __ldg(p.bias + col + i);
          o[i] = __float2half_rn(epi_act<EPI>(x));
        }
      }
    }
  } else {
    // lane -> (row = it*4 + lane/8, 4 columns starting at (lane%8)*4): 16-byte fp32 accesses, 128 B per row
    const int cc = (lane & 7) * 4;
    const int col = col0 + cc;
    const bool col_ok = col < p.n;
    const bool fast = vec_ok && col + 4 <= p.n;
    float4 res[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == PE_EPI_RESID_F32) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {   // all residual loads of the chunk are issued before any is used
        const int row = row0 + it * 4 + (lane >> 3);
        if (row < p.m && fast) {
          int out_row, resid_row;
          map_rows(p, row, out_row, resid_row);
          res[it] = *reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(resid_row) * p.n + col);
        }
      }
    }
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias != nullptr && fast) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const int row = row0 + rr;
      if (row >= p.m || !col_ok) continue;
      int out_row, resid_row;
      map_rows(p, row, out_row, resid_row);
      const float4 a = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cc);
      float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(out_row) * p.n + col;
      if (fast) {
        float4 r4;
        r4.x = epi_act<EPI>(a.x + bias4.x) + res[it].x;
        r4.y = epi_act<EPI>(a.y + bias4.y) + res[it].y;
        r4.z = epi_act<EPI>(a.z + bias4.z) + res[it].z;
        r4.w = epi_act<EPI>(a.w + bias4.w) + res[it].w;
        *reinterpret_cast<float4*>(o) = r4;
      } else {
        const float v[4] = {a.x, a.y, a.z, a.w};
        for (int i = 0; i < 4 && col + i < p.n; ++i) {
          float x = v[i];
          if (p.bias != nullptr) x += __ldg(p.bias + col + i);
          x = epi_act<EPI>(x);
          if (EPI == PE_EPI_RESID_F32) x += p.resid[static_cast<size_t>(resid_row) * p.n + col + i];
          o[i] = x;
        }
      }
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_out2,
                    const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t ln_bar[2];     // PE_EPI_RESID_LN: "every CTA of the row has published its statistics"
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* trace = p.trace != nullptr ? p.trace + static_cast<size_t>(blockIdx.x) * kTraceSlots : nullptr;
  if (trace != nullptr && threadIdx.x == 0) trace[0] = clock64();
  pdl_launch_dependents();   // the next kernel may become resident as SMs free up; it blocks in its own pdl_wait()
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

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], static_cast<uint32_t>(p.cm + p.cn - 1));
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpiWarps);
      mbar_init(&ln_bar[s], static_cast<uint32_t>(p.cn));
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(&tmem_base_smem, kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();   // peers must not multicast into barriers that are not initialised yet
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // Everything above touched only this CTA's shared memory / TMEM and kernel parameters. The predecessor's output
  // (A via TMA, resid) may be read, and buffers it may still be reading overwritten, only after pdl_wait(): the
  // producer warp first pulls in what does NOT depend on the predecessor (the weights), everyone else waits here.
  if (warp != kWeightWarp || p.static_w == 0) pdl_wait();
  if (trace != nullptr && threadIdx.x == 0) trace[1] = clock64();
  const int num_supers = p.num_super_m * p.num_super_n;
  const int cluster_id = static_cast<int>(blockIdx.x) / csize;
  const int num_clusters = static_cast<int>(gridDim.x) / csize;
  const int a_slice_rows = kBlockM / p.cn;       // my share of the A tile
  const int b_slice_rows = p.block_n / p.cm;     // my share of the W tile

  const uint32_t full_addr0 = keep_in_register(smem_u32(&full_bar[0]));
  const uint32_t empty_addr0 = keep_in_register(smem_u32(&empty_bar[0]));
  const int num_k_blocks = p.num_k_blocks, num_stages = p.stages;
  const bool pair_ok = num_stages >= 4;   // pairing needs two free stages per trip: with a 3-deep ring it starves the pipe
  const int dbg = p.debug_mode;
  // The three single-issuer loops below (A loads, W loads, MMA) are latency chains: barrier wait -> elect -> issue ->
  // loop. Measured per k-block (profiles/r01d_gemm_loops.txt): a wait costs ~100-170 cycles even when the phase is
  // already complete, expect_tx + two TMA issues ~200, four MMAs ~80, commit ~95, loop-back ~150: ~450 in all, more
  // than the tensor time of a 128 x 96..160 tile (2*BN cycles). Hence A and W are issued from different warps, and
  // every loop handles TWO k-blocks per trip: both barrier waits are issued back to back (latencies overlap) and one
  // elected region issues both k-blocks' work.
  if (warp == kProducerWarp) {
    // ---------------------------------------------------------------- A producer (warp converged, one lane issues)
    const int cm = p.cm, cn = p.cn, nsm = p.num_super_m;
    const uint32_t a_off = static_cast<uint32_t>(n_rank * a_slice_rows) * (kBlockK * 2);
    int stage = 0;
    uint32_t phase = 0;
    for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
      const int a_row = ((sup % nsm) * cm + m_rank) * kBlockM + n_rank * a_slice_rows;
      for (int kb = 0; kb < num_k_blocks;) {
        const bool two = pair_ok && kb + 1 < num_k_blocks;
        int stage1 = stage + 1;
        uint32_t phase1 = phase;
        if (stage1 == num_stages) { stage1 = 0; phase1 ^= 1u; }
        const uint32_t e0 = empty_addr0 + static_cast<uint32_t>(stage) * 8u;
        const uint32_t e1 = empty_addr0 + static_cast<uint32_t>(stage1) * 8u;
        // all my destinations drained the stage(s)
        if (!two) mbar_wait_addr(e0, phase ^ 1u);
        else if (!mbar_try2_addr(e0, phase ^ 1u, e1, phase1 ^ 1u)) { mbar_wait_addr(e0, phase ^ 1u); mbar_wait_addr(e1, phase1 ^ 1u); }
        if (elect_one()) {
          const uint32_t f0 = full_addr0 + static_cast<uint32_t>(stage) * 8u;
          const uint32_t sa0 = smem_base + static_cast<uint32_t>(stage) * stage_bytes + a_off;
          if (dbg & 2) mbar_arrive_addr(f0);                     // debug: no loads at all, operands are garbage
          else {
          mbar_arrive_expect_tx_addr(f0, stage_bytes);           // A and W bytes; W's complete_tx may already be in
          if (cn > 1) tma_load_2d_multicast_addr(sa0, &tm_a, f0, kb * kBlockK, a_row, row_mask);
          else tma_load_2d_addr(sa0, &tm_a, f0, kb * kBlockK, a_row);
          }
          if (two) {
            const uint32_t f1 = full_addr0 + static_cast<uint32_t>(stage1) * 8u;
            const uint32_t sa1 = smem_base + static_cast<uint32_t>(stage1) * stage_bytes + a_off;
            if (dbg & 2) mbar_arrive_addr(f1);
            else {
            mbar_arrive_expect_tx_addr(f1, stage_bytes);
            if (cn > 1) tma_load_2d_multicast_addr(sa1, &tm_a, f1, (kb + 1) * kBlockK, a_row, row_mask);
            else tma_load_2d_addr(sa1, &tm_a, f1, (kb + 1) * kBlockK, a_row);
            }
          }
        }
        __syncwarp();
        if (two) { stage = stage1; phase = phase1; }
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        kb += two ? 2 : 1;
      }
    }
  } else if (warp == kWeightWarp) {
    // ---------------------------------------------------------------- W producer (warp converged, one lane issues)
    const int cm = p.cm, cn = p.cn, nsm = p.num_super_m, bn = p.block_n;
    const uint32_t b_off = kABytes + static_cast<uint32_t>(m_rank * b_slice_rows) * (kBlockK * 2);
    // Static weights do not depend on the predecessor kernel: this warp skipped pdl_wait(), so while the predecessor
    // drains it already streams the first stages' weights into shared memory and pulls the rest of its weight panel
    // into L2 (shared out among the CTAs that read the same panel).
    if (p.static_w != 0 && cluster_id < num_supers && elect_one()) {
      const int b_row = ((cluster_id / nsm) * cn + n_rank) * bn + m_rank * b_slice_rows;
      const int sharers = nsm * cm;
      const int me = (cluster_id % nsm) * cm + m_rank;
      for (int kb = num_stages + me; kb < num_k_blocks; kb += sharers) tma_prefetch_l2_2d(&tm_b, kb * kBlockK, b_row);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
      const int b_row = ((sup / nsm) * cn + n_rank) * bn + m_rank * b_slice_rows;
      for (int kb = 0; kb < num_k_blocks;) {
        const bool two = pair_ok && kb + 1 < num_k_blocks;
        int stage1 = stage + 1;
        uint32_t phase1 = phase;
        if (stage1 == num_stages) { stage1 = 0; phase1 ^= 1u; }
        const uint32_t e0 = empty_addr0 + static_cast<uint32_t>(stage) * 8u;
        const uint32_t e1 = empty_addr0 + static_cast<uint32_t>(stage1) * 8u;
        if (!two) mbar_wait_addr(e0, phase ^ 1u);
        else if (!mbar_try2_addr(e0, phase ^ 1u, e1, phase1 ^ 1u)) { mbar_wait_addr(e0, phase ^ 1u); mbar_wait_addr(e1, phase1 ^ 1u); }
        if (elect_one()) {
          const uint32_t f0 = full_addr0 + static_cast<uint32_t>(stage) * 8u;
          const uint32_t sb0 = smem_base + static_cast<uint32_t>(stage) * stage_bytes + b_off;
          if (dbg & 2) {}
          else if (cm > 1) tma_load_2d_multicast_addr(sb0, &tm_b, f0, kb * kBlockK, b_row, col_mask);
          else tma_load_2d_addr(sb0, &tm_b, f0, kb * kBlockK, b_row);
          if (two) {
            const uint32_t f1 = full_addr0 + static_cast<uint32_t>(stage1) * 8u;
            const uint32_t sb1 = smem_base + static_cast<uint32_t>(stage1) * stage_bytes + b_off;
            if (dbg & 2) {}
            else if (cm > 1) tma_load_2d_multicast_addr(sb1, &tm_b, f1, (kb + 1) * kBlockK, b_row, col_mask);
            else tma_load_2d_addr(sb1, &tm_b, f1, (kb + 1) * kBlockK, b_row);
          }
        }
        __syncwarp();
        if (two) { stage = stage1; phase = phase1; }
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        kb += two ? 2 : 1;
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issuer (warp converged, one lane issues)
    const uint32_t idesc = umma_idesc_f16(kBlockM, p.block_n);
    const uint32_t a_lo0 = umma_desc_lo(smem_base);
    const uint32_t stage_lo = stage_bytes >> 4;
    const uint32_t tmem_full_addr0 = keep_in_register(smem_u32(&tmem_full_bar[0]));
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kAccStride);
      for (int kb = 0; kb < num_k_blocks;) {
        const bool two = pair_ok && kb + 1 < num_k_blocks;
        int stage1 = stage + 1;
        uint32_t phase1 = phase;
        if (stage1 == num_stages) { stage1 = 0; phase1 ^= 1u; }
        const uint32_t f0 = full_addr0 + static_cast<uint32_t>(stage) * 8u;
        const uint32_t f1 = full_addr0 + static_cast<uint32_t>(stage1) * 8u;
        if (!two) mbar_wait_addr(f0, phase);
        else if (!mbar_try2_addr(f0, phase, f1, phase1)) { mbar_wait_addr(f0, phase); mbar_wait_addr(f1, phase1); }
        tcgen05_fence_after();
        if (trace != nullptr && sup == cluster_id && kb == 0 && lane == 0) trace[2] = clock64();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(stage) * stage_lo;
          const uint32_t b_lo = a_lo + (kABytes >> 4);
          if (!(dbg & 1)) {
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advancing K by 16 fp16 = 32 bytes inside the swizzled row: +2 in 16-byte units
            umma_f16_ss_lohi(d_tmem, a_lo + k * 2, b_lo + k * 2, kUmmaDescHiSw128, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          }
          // release the stage to every CTA that wrote part of it (incl. this one)
          const uint32_t e0 = empty_addr0 + static_cast<uint32_t>(stage) * 8u;
          if (csize > 1) umma_commit_multicast_addr(e0, peers_mask);
          else umma_commit_addr(e0);
          if (two) {
            const uint32_t a_lo1 = a_lo0 + static_cast<uint32_t>(stage1) * stage_lo;
            const uint32_t b_lo1 = a_lo1 + (kABytes >> 4);
            if (!(dbg & 1)) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16_ss_lohi(d_tmem, a_lo1 + k * 2, b_lo1 + k * 2, kUmmaDescHiSw128, idesc, 1u);
            }
            const uint32_t e1 = empty_addr0 + static_cast<uint32_t>(stage1) * 8u;
            if (csize > 1) umma_commit_multicast_addr(e1, peers_mask);
            else umma_commit_addr(e1);
          }
          if (kb + (two ? 2 : 1) >= num_k_blocks) umma_commit_addr(tmem_full_addr0 + static_cast<uint32_t>(acc) * 8u);
        }
        __syncwarp();
        if (two) { stage = stage1; phase = phase1; }
        if (++stage == num_stages) { stage = 0; phase ^= 1u; }
        kb += two ? 2 : 1;
      }
      if (trace != nullptr && sup == cluster_id && lane == 0) trace[3] = clock64();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // -------------------------------------------------------------- epilogue warps
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32) are this warp's
    const int group = warp >> 2;           // 32-column chunks c = group, group + 4, ... are this warp's
    const int chunks = p.block_n >> 5;     // block_n is a multiple of 32
    float* stg = reinterpret_cast<float*>(smem_gen + p.stg_offset) + warp * kStgFloatsPerWarp;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (EPI == PE_EPI_RESID_LN) {
      // ---- projection + residual add + LayerNorm in one epilogue. The CN CTAs of a cluster hold the CN column slices
      // (BN <= 128: one 32-column chunk per warp) of the same 128 rows. Per row: every thread reduces its 32 values to
      // (mean, M2); chunks are merged per CTA, CTAs exchange (mean, M2) through distributed shared memory (one remote
      // mbarrier arrival per peer), merged in rank order with Chan's update - every CTA derives bit-identical row
      // statistics. Then v (or LayerNorm(v)) goes out as fp32 and LayerNorm(v) as fp16 through the TMA-store boxes.
      float2* ln_part = reinterpret_cast<float2*>(smem_gen + p.ln_off);   // [4 chunk groups][128 rows]
      float2* ln_cta = ln_part + 4 * kBlockM;                             // [2][128], double-buffered by tile parity
      const bool active = group < chunks;
      const int r_local = quarter * 32 + lane;
      const float inv_n = 1.0f / static_cast<float>(p.n);
      uint8_t* box = reinterpret_cast<uint8_t*>(stg);
      const uint32_t box_addr = smem_u32(box);
      int par = 0;
      uint32_t ln_phase[2] = {0u, 0u};
      for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
        const int m_blk = (sup % p.num_super_m) * p.cm + m_rank;
        const int n_blk = (sup / p.num_super_m) * p.cn + n_rank;
        mbar_wait_relaxed(&tmem_full_bar[acc], acc_phase);
        tcgen05_fence_after();
        const int row0 = m_blk * kBlockM + quarter * 32;
        const int row = row0 + lane;
        const int gcol = n_blk * p.block_n + group * 32;
        float v[32];
        if (active) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>(acc * kAccStride + group * 32), r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);   // the accumulator lives in registers from here on
        float mean_c = 0.f, m2_c = 0.f;
        if (active) {
          if (row < p.m) {
            const float4* r4 = reinterpret_cast<const float4*>(p.resid + static_cast<size_t>(row) * p.n + gcol);
#pragma unroll
            for (int h = 0; h < 2; ++h) {     // two batches of four independent 16-byte residual loads
              float4 rv[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) rv[i] = r4[4 * h + i];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c = 4 * (4 * h + i);
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol) + 4 * h + i);
                v[c] = __fadd_rn(__fadd_rn(v[c], b4.x), rv[i].x); v[c + 1] = __fadd_rn(__fadd_rn(v[c + 1], b4.y), rv[i].y);
                v[c + 2] = __fadd_rn(__fadd_rn(v[c + 2], b4.z), rv[i].z); v[c + 3] = __fadd_rn(__fadd_rn(v[c + 3], b4.w), rv[i].w);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          ln_chunk32(v, mean_c, m2_c);
        }
        ln_part[group * kBlockM + r_local] = make_float2(mean_c, m2_c);
        named_barrier<2>(kNumEpiWarps * 32);
        if (group == 0) {   // this CTA's slice of the row: merge its chunks in column order
          float cnt = 32.f, mean = mean_c, m2 = m2_c;
          for (int c = 1; c < chunks; ++c) {
            const float2 q = ln_part[c * kBlockM + r_local];
            ln_merge(cnt, mean, m2, 32.f, q.x, q.y);
          }
          ln_cta[par * kBlockM + r_local] = make_float2(mean, m2);
        }
        named_barrier<3>(kNumEpiWarps * 32);
        if (warp == 0 && lane < p.cn)   // tell every CTA of the row (incl. this one) that my statistics are published
          mbar_arrive_cluster(mapa_shared(smem_u32(&ln_bar[par]), static_cast<uint32_t>(lane)));
        mbar_wait_cluster(&ln_bar[par], ln_phase[par]);
        ln_phase[par] ^= 1u;
        float cnt = 0.f, mean = 0.f, m2 = 0.f;
        {
          const uint32_t mine = smem_u32(&ln_cta[par * kBlockM + r_local]);
          const float bn = static_cast<float>(p.block_n);
          for (int j = 0; j < p.cn; ++j) {     // slices in column order (cluster rank = slice index)
            const float2 q = ld_dsmem_f2(mapa_shared(mine, static_cast<uint32_t>(j)));
            ln_merge(cnt, mean, m2, bn, q.x, q.y);
          }
        }
        const float rstd = ln_rstd(m2, inv_n, p.ln_eps);
        if (active) {
          uint8_t* my_row128 = box + lane * 128;
          uint8_t* my_row64 = box + lane * 64;
          const int sw128 = lane & 7, sw64 = (lane >> 1) & 3;
          // normalised values are recomputed where they are needed instead of being kept beside v (register budget: 96)
          auto norm4 = [&](int i) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + gcol) + i);
            const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + gcol) + i);
            return make_float4(ln_apply(v[4 * i], mean, rstd, g4.x, e4.x), ln_apply(v[4 * i + 1], mean, rstd, g4.y, e4.y),
                               ln_apply(v[4 * i + 2], mean, rstd, g4.z, e4.z), ln_apply(v[4 * i + 3], mean, rstd, g4.w, e4.w));
          };
          if (p.has_f32) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 o = p.f32_is_ln ? norm4(j) : make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              *reinterpret_cast<float4*>(my_row128 + ((j ^ sw128) << 4)) = o;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.m) {
              tma_store_2d(&tm_out, box_addr, gcol, row0);
              tma_store_commit();
            }
          }
          if (p.has_f16) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 lo = norm4(2 * j), hi = norm4(2 * j + 1);
              uint4 pk;
              __half2* h2 = reinterpret_cast<__half2*>(&pk);
              h2[0] = __floats2half2_rn(lo.x, lo.y); h2[1] = __floats2half2_rn(lo.z, lo.w);
              h2[2] = __floats2half2_rn(hi.x, hi.y); h2[3] = __floats2half2_rn(hi.z, hi.w);
              *reinterpret_cast<uint4*>(my_row64 + ((j ^ sw64) << 4)) = pk;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.m) {
              tma_store_2d(&tm_out2, box_addr, gcol, row0);
              tma_store_commit();
            }
          }
        }
        par ^= 1;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    } else
    for (int sup = cluster_id; sup < num_supers; sup += num_clusters) {
      const int m_blk = (sup % p.num_super_m) * p.cm + m_rank;
      const int n_blk = (sup / p.num_super_m) * p.cn + n_rank;
      mbar_wait_relaxed(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      if (trace != nullptr && sup == cluster_id && warp == 0 && lane == 0) trace[4] = clock64();
      const int row0 = m_blk * kBlockM + quarter * 32;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * kAccStride);
      if (p.tma_store) {
        // ---- TMEM -> registers -> (bias, activation, convert) -> swizzled shared box -> TMA store.
        // One box = this warp's 32 rows x 32 columns: 64-byte rows (fp16, SWIZZLE_64B) or 128-byte rows (fp32,
        // SWIZZLE_128B), so any block_n that is a multiple of 32 tiles exactly; TMA clips the M / N tails.
        constexpr bool kHalfOut = (EPI == PE_EPI_F16 || EPI == PE_EPI_GELU_F16);
        uint8_t* box = reinterpret_cast<uint8_t*>(stg);            // <= 4 KiB, 1024-byte aligned
        const uint32_t box_addr = smem_u32(box);
        uint8_t* my_row = box + lane * (kHalfOut ? 64 : 128);
        // 16-byte chunk j of row r lives at j ^ (r & 7) (128B swizzle) or j ^ ((r >> 1) & 3) (64B swizzle)
        const int sw = kHalfOut ? ((lane >> 1) & 3) : (lane & 7);
        for (int c = group; c < chunks; c += kNumEpiWarps / 4) {
          const int gcol = n_blk * p.block_n + c * 32;
          if (lane == 0) tma_store_wait_read();                    // the previous box has left shared memory
          __syncwarp();
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), r);
          tmem_wait_ld();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (p.bias != nullptr) {
            if (gcol + 32 <= p.n) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gcol) + i);
                v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (gcol + i < p.n) v[i] += __ldg(p.bias + gcol + i);
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = epi_act<EPI>(v[i]);
          if (kHalfOut) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                          // 4 chunks of 8 halves
              uint4 pk;
              __half2* h2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
              for (int i = 0; i < 4; ++i) h2[i] = __floats2half2_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
              *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)                            // 8 chunks of 4 floats
              *reinterpret_cast<float4*>(my_row + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          fence_proxy_async_smem();                                // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0 && row0 < p.m && gcol < p.n) {
            tma_store_2d(&tm_out, box_addr, gcol, row0);
            tma_store_commit();
          }
        }
      } else
      for (int c = group; c < chunks; c += kNumEpiWarps / 4) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(c * 32), r);
        tmem_wait_ld();
        if (trace != nullptr && sup == cluster_id && warp == 0 && lane == 0 && c == group) trace[7] = clock64();
        float4* dst = reinterpret_cast<float4*>(stg + lane * kStgLd);   // lane = accumulator row
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        if (trace != nullptr && sup == cluster_id && warp == 0 && lane == 0 && c == group) trace[8] = clock64();
        const int col0 = n_blk * p.block_n + c * 32;
        if (row0 < p.m && col0 < p.n) epilogue_chunk32<EPI>(p, stg, row0, col0, lane);
        __syncwarp();
        if (trace != nullptr && sup == cluster_id && warp == 0 && lane == 0 && c == group) trace[9] = clock64();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (trace != nullptr && sup == cluster_id && warp == 0 && lane == 0) trace[5] = clock64();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  if (warp < kNumEpiWarps && lane == 0) tma_store_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  // no CTA may leave while a peer can still multicast into its shared memory or arrive on its barriers
  if (csize > 1) cluster_sync_all();
  if (trace != nullptr && threadIdx.x == 0) trace[6] = clock64();
  if (warp == kMmaWarp) {
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

static long long* g_gemm_trace = nullptr;   // set by pe_debug_gemm_trace
void set_gemm_trace(void* buf) { g_gemm_trace = static_cast<long long*>(buf); }

// Output [rows, cols] (fp16 or fp32) -> boxes of [32 rows, 128 bytes] with the 128-byte swizzle for TMA stores.
static int encode_out_2d(CUtensorMap* map, void* ptr, uint64_t rows, uint64_t cols, bool half_out) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return PE_ERR_CUDA;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * (half_out ? 2u : 4u)};
  const cuuint32_t box[2] = {32u, 32u};   // 64-byte (fp16) or 128-byte (fp32) rows
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, half_out ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         half_out ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed (CUresult %d)", static_cast<int>(rc));
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

// Cost model in SM cycles, fitted on B200 to in-graph launch times of the ViT-B shapes under forced plans with weights
// streaming from HBM (profiles/r01d_plan_sweep.txt; +-1k cycles over 24 shape x plan points):
//   fixed ~7000: launch gap, set-up, first operands landing, drain;
//   main loop: K/64 blocks of ~(415 + BN) cycles each - a per-block issue/latency cost plus the tile's shared-memory
//     ingest, NOT the 2*BN tensor cycles: one CTA per SM cannot keep tcgen05 busy at these tile sizes;
//   epilogue of a 128 x BN tile ~1200 + c*BN, c = 6 (fp16), 8 (fp32), 24 (fp16 + erf GELU: MUFU / issue bound);
//   tiles are dealt round-robin to <=148 CTAs: a CTA with r tiles hides each epilogue but the last under the next
//     tile's main loop (two TMEM accumulators), unless the epilogue is the longer of the two.
// Clusters (TMA multicast) cut L2 reads, but these shapes are not L2-bound: they only pay the cluster launch cost.
GemmPlan plan_gemm(int m, int n, int k, int epilogue) {
  int forced[3] = {0, 0, 0};
  const char* env = getenv("PE_GEMM_FORCE");   // "CM,CN,BN": tuning / debugging only
  if (env != nullptr && sscanf(env, "%d,%d,%d", &forced[0], &forced[1], &forced[2]) == 3 && forced[0] > 0)
    return {forced[0], forced[1], forced[2]};
  GemmPlan best = {1, 1, 32};
  double best_cost = 1e30;
  const int shapes[4][2] = {{1, 1}, {2, 1}, {1, 2}, {2, 2}};
  const double kb = (k + kBlockK - 1) / kBlockK;
  const double epi_per_col = epilogue == PE_EPI_GELU_F16 ? 24.0 : (epilogue == PE_EPI_F16 ? 6.0 : (epilogue == PE_EPI_TANH_F32 ? 30.0 : 8.0));
  for (const auto& sh : shapes) {
    const int cm = sh[0], cn = sh[1], csize = cm * cn;
    for (int bn = 256; bn >= 32; bn -= 32) {
      if ((bn / cm) % 8 != 0 || (bn % cm) != 0) continue;   // W slices must keep whole 8-row swizzle atoms
      const long sm_ = (m + kBlockM * cm - 1) / (kBlockM * cm);
      const long sn_ = (n + bn * cn - 1) / (bn * cn);
      const long supers = sm_ * sn_;
      const long avail = max_clusters(csize);
      const long rounds = (supers + avail - 1) / avail;       // tiles of the busiest CTA
      const double loop = kb * (415.0 + bn);
      const double epi = 1200.0 + epi_per_col * bn;
      const double cost = 7000.0 + (csize > 1 ? 800.0 : 0.0) + loop + (rounds - 1) * (loop > epi ? loop : epi) + epi;
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        best = {cm, cn, bn};
      }
    }
  }
  return best;
}

template <int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const CUtensorMap& tout2,
                       const GemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  // the attribute bounds DYNAMIC shared memory; static barriers live outside it (227 KiB total per CTA)
  const int max_smem = kPipeSmemBudget + kStgBytes + 1024;
  if (!configured) {
    PE_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    configured = true;
  }
  const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.block_n) * (kBlockK * 2);
  const size_t ring = static_cast<size_t>(p.stages) * stage_bytes, stg_end = static_cast<size_t>(p.stg_offset) + kStgBytes;
  size_t smem = (ring > stg_end ? ring : stg_end) + 1024;
  if (EPI == PE_EPI_RESID_LN) smem = static_cast<size_t>(p.ln_off) + kLnAreaBytes + 1024;
  const int csize = p.cm * p.cn;
  const int supers = p.num_super_m * p.num_super_n;
  const int avail = max_clusters(csize);
  const int clusters = supers < avail ? supers : avail;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * csize));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(csize);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  PE_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<EPI>, ta, tb, tout, tout2, p));
  count_launches(1);
  return PE_OK;
}

// Tile grid, ring depth and staging placement that follow from a plan (shared by the launcher and pe_debug_gemm_plan).
static void fill_geometry(GemmParams& p, const GemmPlan& plan, int m, int n, int k, int reserve = 0) {
  p.m = m; p.n = n; p.k = k;
  p.block_n = plan.bn;
  p.cm = plan.cm;
  p.cn = plan.cn;
  const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.block_n) * (kBlockK * 2);
  p.num_m_blocks = (m + kBlockM - 1) / kBlockM;
  p.num_n_blocks = (n + p.block_n - 1) / p.block_n;
  p.num_super_m = (p.num_m_blocks + p.cm - 1) / p.cm;
  p.num_super_n = (p.num_n_blocks + p.cn - 1) / p.cn;
  // One tile per CTA: the epilogue starts only after the last MMA has read the ring, so its staging may alias the
  // ring and the whole 224 KiB go to pipeline depth (BN=256: 4 stages instead of 3). With several tiles per CTA the
  // next tile's loads overlap the epilogue and the two need separate space.
  const bool one_round = p.num_super_m * p.num_super_n <= max_clusters(p.cm * p.cn);
  int stages = ((one_round ? kPipeSmemBudget + kStgBytes : kPipeSmemBudget) - reserve) / static_cast<int>(stage_bytes);
  p.stages = stages > kMaxStages ? kMaxStages : (stages < 2 ? 2 : stages);
  if (const char* cap = getenv("PE_GEMM_STAGES")) {   // tuning scripts only
    const int c = atoi(cap);
    if (c >= 2 && c < p.stages) p.stages = c;
  }
  p.stg_offset = one_round ? 0 : p.stages * static_cast<int>(stage_bytes);
  p.num_k_blocks = (k + kBlockK - 1) / kBlockK;
}

int linear_impl(const void* a, const void* w, const void* bias, const void* resid, void* out, int m, int n, int k,
                int epilogue, int rows_per_item, int out_item_rows, int out_row_offset, int resid_per_item,
                int static_w, cudaStream_t stream) {
  PE_REQUIRE(a && w && out, "pe_linear: null pointer");
  PE_REQUIRE(m > 0 && n > 0 && k > 0, "pe_linear: bad shape m=%d n=%d k=%d", m, n, k);
  PE_REQUIRE((k & 7) == 0, "pe_linear: k=%d must be a multiple of 8 (16-byte TMA row pitch)", k);
  PE_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
             "pe_linear: operands must be 16-byte aligned");
  PE_REQUIRE(epilogue != PE_EPI_RESID_F32 || resid != nullptr, "pe_linear: PE_EPI_RESID_F32 needs resid");
  int rc = require_sm100();
  if (rc != PE_OK) return rc;

  const GemmPlan plan = plan_gemm(m, n, k, epilogue);
  PE_REQUIRE(plan.bn >= 32 && plan.bn <= 256 && plan.bn % 32 == 0 && plan.cm >= 1 && plan.cn >= 1 &&
                 plan.cm * plan.cn <= 8 && kBlockM % plan.cn == 0 && (plan.bn / plan.cm) % 8 == 0,
             "pe_linear: bad tile plan cm=%d cn=%d bn=%d", plan.cm, plan.cn, plan.bn);
  GemmParams p;
  p.tma_store = 0;
  p.bias = static_cast<const float*>(bias);
  p.resid = static_cast<const float*>(resid);
  p.out = out;
  p.m = m; p.n = n; p.k = k;
  fill_geometry(p, plan, m, n, k);
  p.trace = g_gemm_trace;
  p.debug_mode = 0;
  if (const char* dm = getenv("PE_GEMM_DEBUG_MODE")) p.debug_mode = atoi(dm);   // timing experiments only: results are wrong
  p.rows_per_item = rows_per_item;
  p.out_item_rows = out_item_rows;
  p.out_row_offset = out_row_offset;
  p.resid_per_item = resid_per_item;
  p.static_w = static_w != 0 ? 1 : 0;

  CUtensorMap ta, tb;   // each CTA loads its SLICE of a tile: 128/CN rows of A, BN/CM rows of W
  rc = encode_f16_2d(&ta, a, static_cast<uint64_t>(m), static_cast<uint64_t>(k), static_cast<uint32_t>(kBlockM / p.cn));
  if (rc != PE_OK) return rc;
  rc = encode_f16_2d(&tb, w, static_cast<uint64_t>(n), static_cast<uint64_t>(k), static_cast<uint32_t>(p.block_n / p.cm));
  if (rc != PE_OK) return rc;

  // Output through TMA whenever the row pitch allows it; the residual / row-remapping epilogues keep per-thread stores.
  const bool half_out = epilogue == PE_EPI_F16 || epilogue == PE_EPI_GELU_F16;
  const size_t out_pitch = static_cast<size_t>(n) * (half_out ? 2 : 4);
  CUtensorMap tout = ta;   // placeholder when unused
  p.tma_store = (epilogue != PE_EPI_RESID_F32 && rows_per_item == 0 && (out_pitch & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
  if (getenv("PE_GEMM_NO_TMA_STORE") != nullptr) p.tma_store = 0;
  if (p.tma_store) {
    rc = encode_out_2d(&tout, out, static_cast<uint64_t>(m), static_cast<uint64_t>(n), half_out);
    if (rc != PE_OK) return rc;
  }
  switch (epilogue) {
    case PE_EPI_F16: return launch_gemm<PE_EPI_F16>(ta, tb, tout, tout, p, stream);
    case PE_EPI_GELU_F16: return launch_gemm<PE_EPI_GELU_F16>(ta, tb, tout, tout, p, stream);
    case PE_EPI_RESID_F32: return launch_gemm<PE_EPI_RESID_F32>(ta, tb, tout, tout, p, stream);
    case PE_EPI_F32: return launch_gemm<PE_EPI_F32>(ta, tb, tout, tout, p, stream);
    case PE_EPI_TANH_F32: return launch_gemm<PE_EPI_TANH_F32>(ta, tb, tout, tout, p, stream);
    default: set_error("pe_linear: unknown epilogue %d", epilogue); return PE_ERR_INVALID;
  }
}

// Cluster width for the fused projection + residual + LayerNorm: the CN CTAs of a cluster own the CN column slices of a
// row, each at most 128 columns wide (one 32-column chunk per epilogue warp). 0 = this width is not supported.
int linear_ln_cluster(int n) {
  for (int cn = 8; cn >= 1; cn >>= 1)
    if (n % cn == 0 && (n / cn) % 32 == 0 && n / cn <= 128 && kBlockM % cn == 0) return cn;
  return 0;
}

// out_f32 (nullable) = v or LayerNorm(v) (`f32_is_ln`), out_f16 (nullable) = LayerNorm(v), v = a @ w^T + bias + resid.
int linear_ln_impl(const void* a, const void* w, const void* bias, const void* resid, const void* gamma, const void* beta,
                   float eps, void* out_f32, int f32_is_ln, void* out_f16, int m, int n, int k, int static_w,
                   cudaStream_t stream) {
  PE_REQUIRE(a && w && bias && resid && gamma && beta && (out_f32 || out_f16), "pe_linear_residual_layernorm: null pointer");
  PE_REQUIRE(m > 0 && n > 0 && k > 0 && (k & 7) == 0, "pe_linear_residual_layernorm: bad shape m=%d n=%d k=%d", m, n, k);
  const int cn = linear_ln_cluster(n);
  PE_REQUIRE(cn > 0, "pe_linear_residual_layernorm: n=%d is not supported (needs n = cn * bn, cn in {1,2,4,8}, bn %% 32 == 0, bn <= 128)", n);
  PE_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(resid) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15) == 0 &&
                 (out_f32 == nullptr || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) &&
                 (out_f16 == nullptr || (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0),
             "pe_linear_residual_layernorm: operands must be 16-byte aligned");
  int rc = require_sm100();
  if (rc != PE_OK) return rc;
  const GemmPlan plan = {1, cn, n / cn};
  GemmParams p = {};
  p.bias = static_cast<const float*>(bias);
  p.resid = static_cast<const float*>(resid);
  p.out = out_f32;
  fill_geometry(p, plan, m, n, k, kLnAreaBytes);
  p.trace = nullptr;
  p.debug_mode = 0;
  p.tma_store = 1;
  p.static_w = static_w != 0 ? 1 : 0;
  p.ln_gamma = static_cast<const float*>(gamma);
  p.ln_beta = static_cast<const float*>(beta);
  p.ln_eps = eps;
  p.f32_is_ln = f32_is_ln;
  p.has_f32 = out_f32 != nullptr;
  p.has_f16 = out_f16 != nullptr;
  {
    const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(p.block_n) * (kBlockK * 2);
    const size_t ring = static_cast<size_t>(p.stages) * stage_bytes, stg_end = static_cast<size_t>(p.stg_offset) + kStgBytes;
    p.ln_off = static_cast<int>(((ring > stg_end ? ring : stg_end) + 1023) / 1024 * 1024);
  }
  CUtensorMap ta, tb;
  rc = encode_f16_2d(&ta, a, static_cast<uint64_t>(m), static_cast<uint64_t>(k), static_cast<uint32_t>(kBlockM / p.cn));
  if (rc != PE_OK) return rc;
  rc = encode_f16_2d(&tb, w, static_cast<uint64_t>(n), static_cast<uint64_t>(k), static_cast<uint32_t>(p.block_n));
  if (rc != PE_OK) return rc;
  CUtensorMap t32 = ta, t16 = ta;
  if (out_f32 != nullptr) {
    rc = encode_out_2d(&t32, out_f32, static_cast<uint64_t>(m), static_cast<uint64_t>(n), false);
    if (rc != PE_OK) return rc;
  }
  if (out_f16 != nullptr) {
    rc = encode_out_2d(&t16, out_f16, static_cast<uint64_t>(m), static_cast<uint64_t>(n), true);
    if (rc != PE_OK) return rc;
  }
  return launch_gemm<PE_EPI_RESID_LN>(ta, tb, t32, t16, p, stream);
}

// Host-only: the plan and launch geometry pe_linear would use for this shape (no device needed).
// out6 = {cm, cn, block_n, stages, tiles, ctas}.
int gemm_plan_query(int m, int n, int k, int epilogue, int* out6) {
  PE_REQUIRE(out6 != nullptr && m > 0 && n > 0 && k > 0, "pe_debug_gemm_plan: bad arguments");
  PE_REQUIRE(epilogue >= PE_EPI_F16 && epilogue <= PE_EPI_TANH_F32, "pe_debug_gemm_plan: unknown epilogue %d", epilogue);
  const GemmPlan plan = plan_gemm(m, n, k, epilogue);
  GemmParams p = {};
  fill_geometry(p, plan, m, n, k);
  const int supers = p.num_super_m * p.num_super_n, avail = max_clusters(p.cm * p.cn);
  out6[0] = p.cm; out6[1] = p.cn; out6[2] = p.block_n; out6[3] = p.stages;
  out6[4] = p.num_m_blocks * p.num_n_blocks;
  out6[5] = (supers < avail ? supers : avail) * p.cm * p.cn;
  return PE_OK;
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
