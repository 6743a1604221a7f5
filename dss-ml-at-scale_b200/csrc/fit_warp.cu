// fit_warp.cu -- warp-per-series fit + forecast on CUDA cores (general path).
//
// One warp owns one series (one (Product,SKU) group of the reference fan-out,
// group_apply/02_Fine_Grained_Demand_Forecasting.py:523-528) and does what the
// reference UDF does for it (02:435-494) in the whitened calendar basis:
//   b   = sum_{t observed} a_t (y_t - c)            lanes stride over t, coalesced 128 B reads
//   D   = sum_{t missing}  a_t a_t^T                warp-cooperative, only for NaN positions
//   G_i = diag(kept) - D ; in-order Cholesky of G_i in shared memory with pivot dropping
//   gamma = G_i^-1 b ;  yhat_t = c + a_t . gamma     for the requested rows
// It handles everything (NaN masks, any leading dimension, any number of
// prediction rows) and is also the masked fix-up pass behind the tcgen05 kernel.
// Bound: HBM for fully observed data; see DESIGN.md section 4.
#include "mmf_internal.cuh"

namespace mmf {
namespace {

constexpr int WARPS = 12;
constexpr int THREADS = WARPS * 32;
constexpr int U = 8;                       // independent 128-B row segments in flight per warp
constexpr int DPL = (NPAIR + 31) / 32;     // packed Gram entries per lane (5)

struct WarpScratch {
  float G[P][P + 1];
  float diag0[P];
  float b[P];
};

__device__ __forceinline__ bool is_finite_bits(float v) {
  return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct ARows {
  const float4* s;    // shared copy [4][srows]
  const float4* g;    // global      [4][n_rows_pad]
  int srows;
  int grows;
  __device__ __forceinline__ float4 vec(int j, int t) const {
    return (t < srows) ? s[j * srows + t] : __ldg(&g[(size_t)j * grows + t]);
  }
  __device__ __forceinline__ float elem(int t, int i) const {
    const float* base = (t < srows) ? reinterpret_cast<const float*>(s + (i >> 2) * srows + t)
                                    : reinterpret_cast<const float*>(g + (size_t)(i >> 2) * grows + t);
    return base[i & 3];
  }
};

__global__ void __launch_bounds__(THREADS, 2)
fit_warp_kernel(const DesignView d, const FitArgs a, const int smem_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (a.pending_count != nullptr && *a.pending_count == 0u) return;   // grid-uniform early exit

  float4* s_a4 = reinterpret_cast<float4*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  WarpScratch& scr = reinterpret_cast<WarpScratch*>(smem_raw + (size_t)4 * smem_rows * sizeof(float4))[warp];

  for (int i = threadIdx.x; i < 4 * smem_rows; i += THREADS) {
    const int j = i / smem_rows, t = i - j * smem_rows;
    s_a4[i] = d.a4[(size_t)j * d.n_rows_pad + t];
  }
  __syncthreads();

  const ARows A{s_a4, d.a4, smem_rows, d.n_rows_pad};
  const int t_fit = d.t_fit;

  // packed lower-triangular entries owned by this lane: e = lane + 32k -> (i >= j)
  int pi[DPL], pj[DPL];
#pragma unroll
  for (int k = 0; k < DPL; ++k) {
    const int e = lane + 32 * k;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    pi[k] = (e < NPAIR) ? i : 0;
    pj[k] = (e < NPAIR) ? e - i * (i + 1) / 2 : 0;
  }

  const int64_t warps_total = (int64_t)gridDim.x * WARPS;
  for (int64_t row = (int64_t)blockIdx.x * WARPS + warp; row < a.n; row += warps_total) {
    if (a.only_pending && a.status[row] != MMF_STATUS_PENDING) continue;
    const float* __restrict__ yr = a.y + row * a.ld_y;
    float* __restrict__ outr = a.out + row * a.ld_out;

    // ---- centring constant: first observed value (exact shift-equivariance needs X[:,0]==1)
    float c = 0.f;
    bool any = false;
    for (int t0 = 0; t0 < t_fit; t0 += 32) {
      const int t = t0 + lane;
      const float v = (t < t_fit) ? __ldg(yr + t) : __int_as_float(0x7fc00000);
      const unsigned m = __ballot_sync(0xffffffffu, is_finite_bits(v));
      if (m) {
        c = __shfl_sync(0xffffffffu, v, __ffs(m) - 1);
        any = true;
        break;
      }
    }
    if (!any) {                                   // no observed fit row
      const float qnan = __int_as_float(0x7fc00000);
      for (int k = lane; k < a.n_pred; k += 32) outr[k] = qnan;
      if (a.out_beta != nullptr && lane < P) a.out_beta[row * P + lane] = qnan;
      if (lane == 0) a.status[row] = MMF_STATUS_EMPTY;
      continue;
    }
    if (!d.has_constant) c = 0.f;

    // ---- moments
    float acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = 0.f;
    float dacc[DPL];
#pragma unroll
    for (int k = 0; k < DPL; ++k) dacc[k] = 0.f;
    unsigned anymiss = 0u;
    int nmiss = 0;

    for (int t0 = 0; t0 < t_fit; t0 += 32 * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + u * 32 + lane;
        v[u] = (t < t_fit) ? __ldcs(yr + t) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int tb = t0 + u * 32;
        if (tb < t_fit) {                          // warp-uniform
          const int t = tb + lane;
          const bool inr = t < t_fit;
          const bool fin = is_finite_bits(v[u]);
          const float r = (inr && fin) ? v[u] - c : 0.f;
          const float4 a0 = A.vec(0, t), a1 = A.vec(1, t), a2 = A.vec(2, t), a3 = A.vec(3, t);
          acc[0] = fmaf(a0.x, r, acc[0]);   acc[1] = fmaf(a0.y, r, acc[1]);
          acc[2] = fmaf(a0.z, r, acc[2]);   acc[3] = fmaf(a0.w, r, acc[3]);
          acc[4] = fmaf(a1.x, r, acc[4]);   acc[5] = fmaf(a1.y, r, acc[5]);
          acc[6] = fmaf(a1.z, r, acc[6]);   acc[7] = fmaf(a1.w, r, acc[7]);
          acc[8] = fmaf(a2.x, r, acc[8]);   acc[9] = fmaf(a2.y, r, acc[9]);
          acc[10] = fmaf(a2.z, r, acc[10]); acc[11] = fmaf(a2.w, r, acc[11]);
          acc[12] = fmaf(a3.x, r, acc[12]); acc[13] = fmaf(a3.y, r, acc[13]);
          acc[14] = fmaf(a3.z, r, acc[14]); acc[15] = fmaf(a3.w, r, acc[15]);
          unsigned mm = __ballot_sync(0xffffffffu, inr && !fin);
          anymiss |= mm;
          nmiss += __popc(mm);
          while (mm) {                             // rare: Gram downdate for each missing t
            const int tt = tb + __ffs(mm) - 1;
            mm &= mm - 1;
#pragma unroll
            for (int k = 0; k < DPL; ++k)
              dacc[k] = fmaf(A.elem(tt, pi[k]), A.elem(tt, pj[k]), dacc[k]);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) acc[p] = warp_sum(acc[p]);

    float g[P];
    int st = MMF_STATUS_OK;
    if (anymiss == 0u) {
      // fully observed: G_i = I on the kept columns, gamma = b
#pragma unroll
      for (int p = 0; p < P; ++p) g[p] = ((d.kept_mask >> p) & 1u) ? acc[p] : 0.f;
    } else {
      // ---- per-series normal equations in shared memory
      // Mostly-missing rows: G_i = I - D would cancel catastrophically in fp32, so re-accumulate the
      // Gram directly over the (few) observed rows instead of downdating over the (many) missing ones.
      const bool direct = 2 * nmiss > t_fit;
      if (direct) {
#pragma unroll
        for (int k = 0; k < DPL; ++k) dacc[k] = 0.f;
        for (int t0 = 0; t0 < t_fit; t0 += 32) {
          const int t = t0 + lane;
          const float v = (t < t_fit) ? __ldg(yr + t) : __int_as_float(0x7fc00000);
          unsigned mm = __ballot_sync(0xffffffffu, is_finite_bits(v));
          while (mm) {
            const int tt = t0 + __ffs(mm) - 1;
            mm &= mm - 1;
#pragma unroll
            for (int k = 0; k < DPL; ++k)
              dacc[k] = fmaf(A.elem(tt, pi[k]), A.elem(tt, pj[k]), dacc[k]);
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < DPL; ++k) {
        const int e = lane + 32 * k;
        if (e < NPAIR) {
          const int i = pi[k], j = pj[k];
          const float full = (i == j && ((d.kept_mask >> i) & 1u)) ? 1.f : 0.f;
          const float gij = direct ? dacc[k] : full - dacc[k];
          scr.G[i][j] = gij;
          scr.G[j][i] = gij;
          if (i == j) scr.diag0[i] = gij;
        }
      }
      if (lane < P) {
        float bl = 0.f;
#pragma unroll
        for (int p = 0; p < P; ++p) bl = (lane == p) ? acc[p] : bl;
        scr.b[lane] = bl;
      }
      __syncwarp();
      // right-looking Cholesky, lane -> (row i = lane&15, column half h = lane>>4)
      const int ri = lane & 15, ch = lane >> 4;
      unsigned dropped = 0u;
      for (int j = 0; j < P; ++j) {
        const float dj = scr.G[j][j];
        const float d0 = scr.diag0[j];
        const bool globally_out = !((d.kept_mask >> j) & 1u);
        const bool keep = !globally_out && d0 > 0.f && dj > MMF_PIVOT_TOL * d0;
        __syncwarp();
        if (keep) {
          const float inv = rsqrtf(dj);
          const float lij = (ri > j) ? scr.G[ri][j] * inv : 0.f;     // column j of L, row ri
          __syncwarp();
          if (ch == 0) {
            if (ri > j) scr.G[ri][j] = lij;
            if (ri == j) scr.G[j][j] = dj * inv;                      // sqrt(dj)
          }
          // trailing update G[ri][k] -= L[ri][j] * L[k][j], k in this lane's column half
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int k = ch * 8 + kk;
            const float lkj = __shfl_sync(0xffffffffu, lij, k);       // lane k (half 0) holds L[k][j]
            if (ri > j && k > j) scr.G[ri][k] = fmaf(-lij, lkj, scr.G[ri][k]);
          }
        } else {
          if (!globally_out && d0 > 0.f) dropped |= 1u << j;
          if (ch == 0) {
            if (ri > j) scr.G[ri][j] = 0.f;
            if (ri == j) scr.G[j][j] = 1.f;
          }
          if (lane == 0) scr.b[j] = 0.f;                              // gamma_j = 0
          // row j of the remaining matrix must not feed later columns: zero G[j][k>j] is not
          // needed (only the lower triangle / column j is read below).
        }
        __syncwarp();
      }
      const unsigned outmask = dropped | ~d.kept_mask;
      // forward solve L z = b (column oriented), then backward L^T gamma = z
      for (int j = 0; j < P; ++j) {
        float zj = 0.f;
        if (!((outmask >> j) & 1u)) zj = scr.b[j] / scr.G[j][j];
        __syncwarp();
        if (lane == j) scr.b[j] = zj;
        if (lane > j && lane < P) scr.b[lane] = fmaf(-scr.G[lane][j], zj, scr.b[lane]);
        __syncwarp();
      }
      for (int j = P - 1; j >= 0; --j) {
        float gj = 0.f;
        if (!((outmask >> j) & 1u)) gj = scr.b[j] / scr.G[j][j];
        __syncwarp();
        if (lane == j) scr.b[j] = gj;
        if (lane < j) scr.b[lane] = fmaf(-scr.G[j][lane], gj, scr.b[lane]);
        __syncwarp();
      }
#pragma unroll
      for (int p = 0; p < P; ++p) g[p] = ((outmask >> p) & 1u) ? 0.f : scr.b[p];
      if (dropped) st = MMF_STATUS_RANKDEF;
      __syncwarp();
    }

    // ---- predictions for rows [pred_start, pred_start + n_pred)
    for (int k = lane; k < a.n_pred; k += 32) {
      const int t = a.pred_start + k;
      const float4 a0 = A.vec(0, t), a1 = A.vec(1, t), a2 = A.vec(2, t), a3 = A.vec(3, t);
      float s = c;
      s = fmaf(a0.x, g[0], s);  s = fmaf(a0.y, g[1], s);  s = fmaf(a0.z, g[2], s);  s = fmaf(a0.w, g[3], s);
      s = fmaf(a1.x, g[4], s);  s = fmaf(a1.y, g[5], s);  s = fmaf(a1.z, g[6], s);  s = fmaf(a1.w, g[7], s);
      s = fmaf(a2.x, g[8], s);  s = fmaf(a2.y, g[9], s);  s = fmaf(a2.z, g[10], s); s = fmaf(a2.w, g[11], s);
      s = fmaf(a3.x, g[12], s); s = fmaf(a3.y, g[13], s); s = fmaf(a3.z, g[14], s); s = fmaf(a3.w, g[15], s);
      __stcs(outr + k, s);
    }
    if (a.out_beta != nullptr && lane < P) {       // beta = W gamma (+ c on the intercept)
      float s = (lane == 0 && d.has_constant) ? c : 0.f;
#pragma unroll
      for (int q = 0; q < P; ++q) s = fmaf(__ldg(d.w + lane * P + q), g[q], s);
      a.out_beta[row * P + lane] = s;
    }
    if (lane == 0) a.status[row] = st;
  }
}

}  // namespace

size_t fit_warp_smem_bytes(const DesignView& d, int* smem_rows) {
  // keep as many design rows resident as fit beside the per-warp scratch (<= ~100 KB so 2 CTAs/SM fit)
  const size_t scratch = sizeof(WarpScratch) * WARPS;
  const size_t budget = 100 * 1024 - scratch;
  int rows = d.n_rows_pad;
  const int max_rows = (int)(budget / (4 * sizeof(float4))) & ~31;
  if (rows > max_rows) rows = max_rows;
  *smem_rows = rows;
  return (size_t)rows * 4 * sizeof(float4) + scratch;
}

cudaError_t launch_fit_warp(const DesignView& d, const FitArgs& a, int sm_count, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  int smem_rows = 0;
  const size_t smem = fit_warp_smem_bytes(d, &smem_rows);
  cudaError_t e = cudaFuncSetAttribute(fit_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_warp_kernel, THREADS, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = (a.n + WARPS - 1) / WARPS;
  const int64_t cap = (int64_t)sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  fit_warp_kernel<<<(unsigned)blocks, THREADS, smem, s>>>(d, a, smem_rows);
  return cudaGetLastError();
}

}  // namespace mmf
