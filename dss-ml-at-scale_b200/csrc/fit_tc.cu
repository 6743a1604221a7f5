// fit_tc.cu -- the sm_100a fast path: TMA-staged series tiles, tcgen05 moment GEMM with the
// series tile as the TMEM A operand, TMEM accumulators, per-row epilogue.
//
// What it computes (reference 02_Fine_Grained_Demand_Forecasting.py:435-494 for every group of the
// applyInPandas fan-out 02:523-528, in the whitened calendar basis of DESIGN.md section 2):
//     b_i  = A_fit^T (y_i - c_i)          [128 series x T] x [T x 16]  -> tcgen05.mma kind::tf32
//     yhat = c_i + A_pred b_i             (fully observed series: G_i = I, gamma_i = b_i)
// fp32-grade accuracy on the tf32 tensor path comes from an exact 2-term split of both operands
//     y - c = hi + lo,  A = A_hi + A_lo  ->  D = hi*[A_hi | A_lo] (N = 32)  +  lo*A_hi (N = 16)
// Rows that contain a non-finite value poison only their own accumulator row (NaN/Inf); the epilogue
// detects that, marks the row MMF_STATUS_PENDING and the masked warp kernel finishes them.
//
// CTA = 6 warps, 1 CTA per SM, persistent over 128-series tiles:
//   warps 0-3  transform + epilogue: thread r owns series row r of the tile == TMEM lane r.
//              LDS.128 its 128-B row of the TMA-swizzled stage, centre, split hi/lo, tcgen05.st the
//              two 32-column A operands; at tile end tcgen05.ld the accumulators and write forecasts.
//   warp 4     TMA producer: y box {32 t x 128 series} + design box {32 t x 32 (hi|lo columns)} per stage.
//   warp 5     MMA issuer (one thread) + TMEM allocation.
// Algorithmic HBM bytes per series: 4*t_fit read + 4*n_pred written.
#include "mmf_internal.cuh"
#include "sm100_ptx.cuh"

namespace mmf {
namespace {

using namespace sm100;

constexpr int TILE_M = 128;            // series per tile == TMEM lanes
constexpr int KC = 32;                 // time steps per stage == one 128-B swizzle row
#ifndef MMF_TC_STAGES
#define MMF_TC_STAGES 8
#endif
#ifndef MMF_TC_ASLOTS
#define MMF_TC_ASLOTS 6
#endif
constexpr int STAGES = MMF_TC_STAGES;  // shared-memory ring (20 KB per stage)
constexpr int ASLOTS = MMF_TC_ASLOTS;  // TMEM A-operand ring: covers the tcgen05.st -> mma -> commit round trip
constexpr int MAX_PRED = 64;           // forecast rows the epilogue supports
constexpr int Y_STAGE_BYTES = TILE_M * KC * 4;      // 16384
constexpr int AT_STAGE_BYTES = 2 * P * KC * 4;      // 4096
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = (32 + 64 * ASLOTS) <= 256 ? 256 : 512;
static_assert(32 + 64 * ASLOTS <= 512, "TMEM holds 512 columns");
constexpr uint32_t ACC_COL = 0;                     // 32 accumulator columns
constexpr uint32_t ASLOT_COL0 = 32;                 // slot j: hi at 32+64j, lo at 32+64j+32

struct SmemLayout {
  // offsets from the 1024-aligned base
  static constexpr int y = 0;
  static constexpr int at = y + STAGES * Y_STAGE_BYTES;
  static constexpr int apred = at + STAGES * AT_STAGE_BYTES;
  static constexpr int bars = apred + MAX_PRED * P * 4;
  static constexpr int n_bars = 2 * STAGES + 2 * ASLOTS + 1;
  static constexpr int tmem_ptr = bars + n_bars * 8;
  static constexpr int total = tmem_ptr + 16;
};

__global__ void __launch_bounds__(THREADS, 1)
fit_tc_kernel(const __grid_constant__ TcLaunch tl, const DesignView d, const FitArgs a,
              uint32_t* __restrict__ pending_count, const int n_tiles, const int n_chunks) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t s_y = sbase + SmemLayout::y;
  const uint32_t s_at = sbase + SmemLayout::at;
  float* s_apred = reinterpret_cast<float*>(smem + SmemLayout::apred);
  const uint32_t s_bars = sbase + SmemLayout::bars;
  auto bar_full = [&](int s) { return s_bars + 8u * s; };
  auto bar_empty = [&](int s) { return s_bars + 8u * (STAGES + s); };
  auto bar_afull = [&](int j) { return s_bars + 8u * (2 * STAGES + j); };
  auto bar_aempty = [&](int j) { return s_bars + 8u * (2 * STAGES + ASLOTS + j); };
  const uint32_t bar_acc = s_bars + 8u * (2 * STAGES + 2 * ASLOTS);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SmemLayout::tmem_ptr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- one-time setup
  if (warp == 5) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(bar_full(s), 1);        // producer's expect_tx arrive (+ TMA bytes)
        mbar_init(bar_empty(s), 5);       // 4 transform warps + 1 tcgen05.commit
      }
      for (int j = 0; j < ASLOTS; ++j) {
        mbar_init(bar_afull(j), 4);       // 4 transform warps
        mbar_init(bar_aempty(j), 1);      // tcgen05.commit
      }
      mbar_init(bar_acc, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
    tmem_relinquish();
  } else if (warp == 4) {
    if (lane == 0) {
      prefetch_tensormap(tl.tmap_y);
      prefetch_tensormap(tl.tmap_at);
    }
  } else {
    // prediction rows of the whitened design -> shared (broadcast-read in the epilogue)
    for (int i = threadIdx.x; i < a.n_pred * P; i += 128)
      s_apred[i] = __ldg(d.apred + (size_t)a.pred_start * P + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 4) {
    // =========================== TMA producer ===========================
    // whole warp converged, one elected lane issues (keeps the tensor-map / barrier operands uniform)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(bar_empty(stage), phase ^ 1u);
        tma_load_2d_x2_elect(bar_full(stage), Y_STAGE_BYTES + AT_STAGE_BYTES,
                             s_y + stage * Y_STAGE_BYTES, tl.tmap_y, ch * KC, tile * TILE_M, L2_EVICT_FIRST,
                             s_at + stage * AT_STAGE_BYTES, tl.tmap_at, ch * KC, 0, L2_EVICT_LAST);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // =========================== MMA issuer ===========================
    // whole warp converged; tcgen05.mma / commit are predicated on elect.sync inside the wrappers
    constexpr uint32_t IDESC_N32 = umma_idesc_tf32(TILE_M, 2 * P);
    constexpr uint32_t IDESC_N16 = umma_idesc_tf32(TILE_M, P);
    int stage = 0, aslot = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(bar_full(stage), phase);          // design chunk landed (same barrier as the y box)
        mbar_wait(bar_afull(aslot), aphase);        // A operand written to TMEM by the transform warps
        tc_fence_after();
        const uint64_t bdesc0 = umma_desc_k_sw128(s_at + stage * AT_STAGE_BYTES);
        const uint32_t a_hi = tmem_base + ASLOT_COL0 + aslot * 64;
        const uint32_t a_lo = a_hi + 32;
#pragma unroll
        for (int k = 0; k < KC / 8; ++k) {
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(k * 2);       // +32 B (16-B units)
          umma_tf32_ts_elect(tmem_base + ACC_COL, a_hi + k * 8, bdesc, IDESC_N32, (ch | k) != 0 ? 1u : 0u);
          umma_tf32_ts_elect(tmem_base + ACC_COL, a_lo + k * 8, bdesc, IDESC_N16, 1u);
        }
        umma_commit_elect(bar_aempty(aslot));
        umma_commit_elect(bar_empty(stage));
        if (ch == n_chunks - 1) umma_commit_elect(bar_acc);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (++aslot == ASLOTS) { aslot = 0; aphase ^= 1u; }
      }
    }
  } else {
    // =========================== transform + epilogue (warps 0-3) ===========================
    const int r = threadIdx.x;                          // row in tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t row_off = static_cast<uint32_t>(r) * 128u;
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    int stage = 0, aslot = 0;
    uint32_t phase = 0, aphase = 0, accphase = 0;
    const bool vec_out = (a.n_pred % 4 == 0) && (a.ld_out % 4 == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      float c = 0.f;
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(bar_full(stage), phase);
        const uint32_t rowp = s_y + stage * Y_STAGE_BYTES + row_off;
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = lds128(rowp + ((static_cast<uint32_t>(q) ^ sw) << 4));
        if (ch == 0) c = d.has_constant ? v[0].x : 0.f;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float e[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const float rr = e[w] - c;
            const uint32_t h = __float_as_uint(rr) & 0xFFFFE000u;
            hi[q * 4 + w] = h;
            lo[q * 4 + w] = __float_as_uint(rr - __uint_as_float(h));
          }
        }
        mbar_wait(bar_aempty(aslot), aphase ^ 1u);      // MMAs that read this A slot have retired
        tc_fence_after();
        const uint32_t a_hi = tmem_base + lane_addr + ASLOT_COL0 + aslot * 64;
        tmem_st_32x32b_x32(a_hi, hi);
        tmem_st_32x32b_x32(a_hi + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_afull(aslot));
          mbar_arrive(bar_empty(stage));                // this warp's smem reads of the stage are done
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (++aslot == ASLOTS) { aslot = 0; aphase ^= 1u; }
      }
      // ---- epilogue: this thread's 16 moments -> forecast row
      mbar_wait(bar_acc, accphase);
      accphase ^= 1u;
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld_32x32b_x32(tmem_base + lane_addr + ACC_COL, acc);
      tmem_wait_ld();
      tc_fence_before();
      float g[P];
      bool finite = true;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        g[p] = __uint_as_float(acc[p]) + __uint_as_float(acc[P + p]);
        finite = finite && ((__float_as_uint(g[p]) & 0x7f800000u) != 0x7f800000u);
        if (!((d.kept_mask >> p) & 1u)) g[p] = 0.f;
      }
      const int64_t row = (int64_t)tile * TILE_M + r;
      const bool live = row < a.n;
      const bool pend = live && !finite;
      const unsigned pm = __ballot_sync(0xffffffffu, pend);
      if (lane == 0 && pm != 0u) atomicAdd(pending_count, __popc(pm));
      if (live) {
        if (finite) {
          float* __restrict__ outr = a.out + row * a.ld_out;
          if (vec_out) {
            for (int k = 0; k < a.n_pred; k += 4) {
              float o[4];
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const float4* ap = reinterpret_cast<const float4*>(s_apred + (k + w) * P);
                const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];
                float s = c;
                s = fmaf(a0.x, g[0], s);  s = fmaf(a0.y, g[1], s);  s = fmaf(a0.z, g[2], s);  s = fmaf(a0.w, g[3], s);
                s = fmaf(a1.x, g[4], s);  s = fmaf(a1.y, g[5], s);  s = fmaf(a1.z, g[6], s);  s = fmaf(a1.w, g[7], s);
                s = fmaf(a2.x, g[8], s);  s = fmaf(a2.y, g[9], s);  s = fmaf(a2.z, g[10], s); s = fmaf(a2.w, g[11], s);
                s = fmaf(a3.x, g[12], s); s = fmaf(a3.y, g[13], s); s = fmaf(a3.z, g[14], s); s = fmaf(a3.w, g[15], s);
                o[w] = s;
              }
              __stcs(reinterpret_cast<float4*>(outr + k), make_float4(o[0], o[1], o[2], o[3]));
            }
          } else {
            for (int k = 0; k < a.n_pred; ++k) {
              const float4* ap = reinterpret_cast<const float4*>(s_apred + k * P);
              const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];
              float s = c;
              s = fmaf(a0.x, g[0], s);  s = fmaf(a0.y, g[1], s);  s = fmaf(a0.z, g[2], s);  s = fmaf(a0.w, g[3], s);
              s = fmaf(a1.x, g[4], s);  s = fmaf(a1.y, g[5], s);  s = fmaf(a1.z, g[6], s);  s = fmaf(a1.w, g[7], s);
              s = fmaf(a2.x, g[8], s);  s = fmaf(a2.y, g[9], s);  s = fmaf(a2.z, g[10], s); s = fmaf(a2.w, g[11], s);
              s = fmaf(a3.x, g[12], s); s = fmaf(a3.y, g[13], s); s = fmaf(a3.z, g[14], s); s = fmaf(a3.w, g[15], s);
              outr[k] = s;
            }
          }
          if (a.out_beta != nullptr) {
            float* __restrict__ br = a.out_beta + row * P;
            for (int p = 0; p < P; ++p) {
              float s = (p == 0 && d.has_constant) ? c : 0.f;
#pragma unroll
              for (int q = 0; q < P; ++q) s = fmaf(__ldg(d.w + p * P + q), g[q], s);
              br[p] = s;
            }
          }
          a.status[row] = MMF_STATUS_OK;
        } else {
          a.status[row] = MMF_STATUS_PENDING;
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

bool fit_tc_supported(const DesignView& d, const FitArgs& a, const char** why) {
  const char* w = nullptr;
  if (a.n_pred > MAX_PRED) w = "n_pred > 64 (holdout/fitted mode uses the warp kernel)";
  else if (a.n_pred < 1) w = "n_pred < 1";
  else if (a.ld_y % 4 != 0) w = "ld_y not a multiple of 4 floats (TMA needs 16-B row pitch)";
  else if ((reinterpret_cast<uintptr_t>(a.y) & 15u) != 0) w = "y not 16-B aligned";
  else if (d.t_fit < 1) w = "t_fit < 1";
  else if (a.n > (int64_t)0x7fffffff - TILE_M) w = "n too large for 32-bit TMA coordinates";
  if (why) *why = w;
  return w == nullptr;
}

cudaError_t launch_fit_tc(const DesignView& d, const FitArgs& a, const TcLaunch& tl, uint32_t* pending_count,
                          int sm_count, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  const int n_tiles = (int)((a.n + TILE_M - 1) / TILE_M);
  const int n_chunks = d.t_pad / KC;
  const size_t smem = SmemLayout::total + 1024;
  cudaError_t e = cudaFuncSetAttribute(fit_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  fit_tc_kernel<<<grid, THREADS, smem, s>>>(tl, d, a, pending_count, n_tiles, n_chunks);
  return cudaGetLastError();
}

}  // namespace mmf
