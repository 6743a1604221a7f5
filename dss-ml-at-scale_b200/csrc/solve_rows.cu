// solve_rows.cu -- thread-per-series normal equations for series with gaps.
//
// The streaming kernel (fit_warp.cu) leaves, for every masked series, a 256-B record: the moments
// b = A_fit^T (y - c) over its observed rows, the centring constant and the grid positions of its missing
// rows.  Here ONE THREAD owns one series and does the rest of build_tune_and_score_model's fit/predict
// (reference 02:435-494) for it without touching shared memory or a barrier:
//   G_i = diag(kept) - sum_{t missing} a_t a_t^T        136 packed entries, in registers
//   in-order Cholesky of G_i with relative pivot dropping (MMF_PIVOT_TOL), fully unrolled
//   L z = b, L^T gamma = z ; yhat_t = c + a_t . gamma ; beta = W gamma on request
// A warp therefore factors 32 different 16x16 systems at once at full lane utilisation -- two orders of
// magnitude fewer issue slots per series than a warp-cooperative Cholesky with a barrier per column.
#include "mmf_internal.cuh"

namespace mmf {
namespace {

constexpr int THREADS = 128;

__device__ __forceinline__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

__device__ __forceinline__ void ldg256_nc(const float* p, float4& lo, float4& hi) {      // p: 32-B aligned
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
               : "l"(p));
}

// both 128-B lines of a record into L1: the record is read piecemeal (moments, then one gap position at a time),
// and every new 32-B sector would otherwise be its own trip to DRAM in the middle of the dependent chain
__device__ __forceinline__ void prefetch_rec(const SolveRec* r) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(r));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(r) + 128));
}

__global__ void __launch_bounds__(THREADS, 2)
solve_rows_kernel(const DesignView d, const FitArgs a) {
  asm volatile("griddepcontrol.wait;" ::: "memory");    // programmatic dependent launch: the producer kernels are done
  const uint32_t count = min(*a.rec_count, a.rec_cap);
  const uint32_t stride = gridDim.x * THREADS;
  bool vec_out = (a.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0);
  for (int q = 0; q + 1 < a.n_out; ++q) vec_out = vec_out && ((reinterpret_cast<uintptr_t>(a.out_more[q]) & 15u) == 0);
  uint32_t i = blockIdx.x * THREADS + threadIdx.x;
  int64_t row_next = i < count ? a.rec_rows[i] : 0;
  if (i < count) prefetch_rec(a.recs + row_next);
  for (; i < count; i += stride) {
    const int64_t row = row_next;
    const SolveRec& rec = a.recs[row];
    if (i + stride < count) {                           // the next series' record arrives while this one is solved
      row_next = a.rec_rows[i + stride];
      prefetch_rec(a.recs + row_next);
    }
    float b[P];
    {
      const float4* bp = reinterpret_cast<const float4*>(rec.b);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = bp[q];
        b[4 * q] = v.x; b[4 * q + 1] = v.y; b[4 * q + 2] = v.z; b[4 * q + 3] = v.w;
      }
    }
    const float c = rec.c;
    const int nm0 = rec.nm[0], nm1 = rec.nm[1];

    // ---- G_i = diag(kept) - sum over the missing rows of a_t a_t^T
    float G[NPAIR];
#pragma unroll
    for (int e = 0; e < NPAIR; ++e) G[e] = 0.f;
#pragma unroll
    for (int j = 0; j < P; ++j) G[tri(j, j)] = ((d.kept_mask >> j) & 1u) ? 1.f : 0.f;
    // software-pipelined: gap positions arrive four at a time (one 8-B load, two groups ahead) and the design row
    // of the next gap is in flight while the 136 FMAs of the current one issue -- the load -> address -> load ->
    // FMA chain and the L1 data pipe (every lane gathers its own addresses) were this kernel's limits
    // every lane gathers a different 64-B design row: two 256-bit loads (one 32-B sector each) instead of four LDG.128
    auto design_row = [&](int t, float4& r0, float4& r1, float4& r2, float4& r3) {
      const float* ap = d.apred + (size_t)t * P;
      ldg256_nc(ap, r0, r1);
      ldg256_nc(ap + 8, r2, r3);
    };
    // (one flat loop over both segments, to pay the warp's max-over-lanes trip count once, measured slower:
    //  the segment-switch bookkeeping costs more issue slots than the shorter trip count saves)
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int cnt = seg ? nm1 : nm0;
      if (cnt == 0) continue;
      const unsigned long long* grp = reinterpret_cast<const unsigned long long*>(rec.miss_t + seg * SOLVE_SEG);
      const int n_grp = (cnt + 3) >> 2;
      unsigned long long cur = grp[0];
      unsigned long long nxt = n_grp > 1 ? grp[1] : 0ull;
      float4 n0, n1, n2, n3;
      design_row((int)(cur & 0xffffull), n0, n1, n2, n3);
#pragma unroll 1
      for (int m = 0; m < cnt; ++m) {
        const float4 a0 = n0, a1 = n1, a2 = n2, a3 = n3;
        const int k1 = (m + 1) & 3;
        if (k1 == 0) {
          cur = nxt;
          const int gi = ((m + 1) >> 2) + 1;
          nxt = gi < n_grp ? grp[gi] : 0ull;
        }
        // past the end: row 0 is a harmless filler (loaded, never used)
        design_row(m + 1 < cnt ? (int)((cur >> (16 * k1)) & 0xffffull) : 0, n0, n1, n2, n3);
        const float av[P] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) G[tri(r, q)] = fmaf(-av[r], av[q], G[tri(r, q)]);
      }
    }

    // ---- in-order right-looking Cholesky with pivot dropping (dropped column: L_jj = 1, rest 0)
    unsigned outmask = ~d.kept_mask & 0xFFFFu;
    unsigned dropped = 0u;
    float diag0[P];
#pragma unroll
    for (int j = 0; j < P; ++j) diag0[j] = G[tri(j, j)];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const float dj = G[tri(j, j)];
      const bool kept_cal = (d.kept_mask >> j) & 1u;
      const bool keep = kept_cal && diag0[j] > 0.f && dj > MMF_PIVOT_TOL * diag0[j];
      if (!keep) {
        outmask |= 1u << j;
        if (kept_cal && diag0[j] > 0.f) dropped |= 1u << j;
      }
      const float inv = keep ? rsqrtf(dj) : 0.f;
      G[tri(j, j)] = keep ? dj * inv : 1.f;
#pragma unroll
      for (int r = j + 1; r < P; ++r) G[tri(r, j)] *= inv;             // column j of L (zero when dropped)
#pragma unroll
      for (int r = j + 1; r < P; ++r)
#pragma unroll
        for (int q = j + 1; q <= r; ++q) G[tri(r, q)] = fmaf(-G[tri(r, j)], G[tri(q, j)], G[tri(r, q)]);
    }
    // ---- L z = b, L^T gamma = z (dropped columns pinned to 0)
#pragma unroll
    for (int j = 0; j < P; ++j) {
      float s = b[j];
#pragma unroll
      for (int q = 0; q < j; ++q) s = fmaf(-G[tri(j, q)], b[q], s);
      b[j] = ((outmask >> j) & 1u) ? 0.f : s / G[tri(j, j)];
    }
#pragma unroll
    for (int j = P - 1; j >= 0; --j) {
      float s = b[j];
#pragma unroll
      for (int r = j + 1; r < P; ++r) s = fmaf(-G[tri(r, j)], b[r], s);
      b[j] = ((outmask >> j) & 1u) ? 0.f : s / G[tri(j, j)];
    }

    if (a.out_gamma != nullptr) {
      float4* gp = reinterpret_cast<float4*>(a.out_gamma + row * P);
      gp[0] = make_float4(b[0], b[1], b[2], b[3]);    gp[1] = make_float4(b[4], b[5], b[6], b[7]);
      gp[2] = make_float4(b[8], b[9], b[10], b[11]);  gp[3] = make_float4(b[12], b[13], b[14], b[15]);
      a.out_c[row] = c;
    }
    // ---- forecasts (16-B stores when the table allows it: a thread owns a whole row of it)
    const int64_t off = row * a.ld_out;
    auto predict = [&](int k) -> float {
      const float4* ap = reinterpret_cast<const float4*>(d.apred + (size_t)(a.pred_start + k) * P);
      const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1), a2 = __ldg(ap + 2), a3 = __ldg(ap + 3);
      float s = c;
      s = fmaf(a0.x, b[0], s);  s = fmaf(a0.y, b[1], s);  s = fmaf(a0.z, b[2], s);  s = fmaf(a0.w, b[3], s);
      s = fmaf(a1.x, b[4], s);  s = fmaf(a1.y, b[5], s);  s = fmaf(a1.z, b[6], s);  s = fmaf(a1.w, b[7], s);
      s = fmaf(a2.x, b[8], s);  s = fmaf(a2.y, b[9], s);  s = fmaf(a2.z, b[10], s); s = fmaf(a2.w, b[11], s);
      s = fmaf(a3.x, b[12], s); s = fmaf(a3.y, b[13], s); s = fmaf(a3.z, b[14], s); s = fmaf(a3.w, b[15], s);
      return s;
    };
    const int n_pred = a.skip_pred ? 0 : a.n_pred;
    int k = 0;
    if (vec_out) {
#pragma unroll 1
      for (; k + 4 <= n_pred; k += 4)
        store_out4(a, off + k, make_float4(predict(k), predict(k + 1), predict(k + 2), predict(k + 3)));
    }
#pragma unroll 1
    for (; k < n_pred; ++k) store_out1(a, off + k, predict(k));
    if (a.out_beta != nullptr) {
#pragma unroll 1
      for (int p = 0; p < P; ++p) {
        float s = (p == 0 && d.has_constant) ? c : 0.f;
#pragma unroll
        for (int q = 0; q < P; ++q) s = fmaf(__ldg(d.w + p * P + q), b[q], s);
        a.out_beta[row * P + p] = s;
      }
    }
    a.status[row] = dropped ? MMF_STATUS_RANKDEF : MMF_STATUS_OK;
  }
}

}  // namespace

cudaError_t launch_solve_rows(const DesignView& d, const FitArgs& a, int sm_count, cudaStream_t s) {
  if (a.recs == nullptr || a.rec_cap == 0) return cudaSuccess;
  const int64_t want = ((int64_t)a.rec_cap + THREADS - 1) / THREADS;
  const int64_t cap = (int64_t)sm_count * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // launch latency hides under the producer
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, solve_rows_kernel, d, a);
}

}  // namespace mmf
