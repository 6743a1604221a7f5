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
#include "solve_math.cuh"

namespace mmf {
namespace {

constexpr int THREADS = 128;

// both 128-B lines of a record into L1: the record is read piecemeal (moments, then one gap position at a time),
// and every new 32-B sector would otherwise be its own trip to DRAM in the middle of the dependent chain
__device__ __forceinline__ void prefetch_rec(const SolveRec* r) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(r));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(r) + 128));
}

// One queued series: Gram downdate, Cholesky with pivot dropping, both solves, forecasts, status.
template <bool MULTI>
__device__ __forceinline__ void solve_one(const DesignView& d0, const FitArgs& a, const CalMeta* __restrict__ cals,
                                          const int64_t row, const bool vec_out) {
  const SolveRec& rec = a.recs[row];
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
  // ragged launches: the record names its calendar; the stacked design tables are indexed from that calendar's row 0
  DesignView d = d0;
  int pred_start = a.pred_start;
  if (MULTI) {
    const int4* cp = reinterpret_cast<const int4*>(cals + rec.cal);
    const int4 m0 = __ldg(cp), m1 = __ldg(cp + 1);
    d.t_fit = m0.x;
    d.kept_mask = static_cast<uint32_t>(m0.w);
    d.apred = d0.apred + (size_t)m1.x * P;
    pred_start = m1.y;
  }

  // ---- G_i = I - sum over the missing rows of a_t a_t^T, Cholesky with pivot dropping, both solves (solve_math.cuh)
  const unsigned dropped = masked_solve(d, b, nm0, nm1, [&](int seg, int gi) {
    return reinterpret_cast<const unsigned long long*>(rec.miss_t + seg * SOLVE_SEG)[gi];
  });

  if (a.out_gamma != nullptr) {
    float4* gp = reinterpret_cast<float4*>(a.out_gamma + row * P);
    gp[0] = make_float4(b[0], b[1], b[2], b[3]);    gp[1] = make_float4(b[4], b[5], b[6], b[7]);
    gp[2] = make_float4(b[8], b[9], b[10], b[11]);  gp[3] = make_float4(b[12], b[13], b[14], b[15]);
    a.out_c[row] = c;
  }
  // ---- forecasts (16-B stores when the table allows it: a thread owns a whole row of it)
  const int64_t off = row * a.ld_out;
  auto predict = [&](int k) -> float {
    const float4* ap = reinterpret_cast<const float4*>(d.apred + (size_t)(pred_start + k) * P);
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

__device__ __forceinline__ bool vec_out_ok(const FitArgs& a) {
  bool v = (a.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0);
  for (int q = 0; q + 1 < a.n_out; ++q) v = v && ((reinterpret_cast<uintptr_t>(a.out_more[q]) & 15u) == 0);
  return v;
}

// The pass over the whole work list, after the producers have finished.  A record whose series is no longer
// MMF_STATUS_DEFERRED was already solved by the streaming consumer below.
template <bool MULTI>
__global__ void __launch_bounds__(THREADS, 2)
solve_rows_kernel(const DesignView d0, const FitArgs a, const CalMeta* __restrict__ cals) {
  asm volatile("griddepcontrol.wait;" ::: "memory");    // programmatic dependent launch: the producer kernels are done
  const uint32_t count = min(*a.rec_count, a.rec_cap);
  const uint32_t stride = gridDim.x * THREADS;
  const bool vec_out = vec_out_ok(a);
  uint32_t i = blockIdx.x * THREADS + threadIdx.x;
  int64_t row_next = i < count ? a.rec_rows[i] : 0;
  if (i < count) prefetch_rec(a.recs + row_next);
  for (; i < count; i += stride) {
    const int64_t row = row_next;
    if (i + stride < count) {                           // the next series' record arrives while this one is solved
      row_next = a.rec_rows[i + stride];
      prefetch_rec(a.recs + row_next);
    }
    if (a.stream_ctl != nullptr) {
      a.rec_rows[i] = -1;                               // the work list of a streaming call is left as it was found: all -1
      if (a.status[row] != MMF_STATUS_DEFERRED) continue;
    }
    solve_one<MULTI>(d0, a, cals, row, vec_out);
  }
}

// The same solve as a CONSUMER running beside fit_tc_kernel.  A warp claims 32 consecutive work-list slots and every
// lane waits for its slot to be published (the producer stores the row index last, with release semantics; the list
// starts out as -1).  When the producer's last CTA has raised `done`, the count of queued records is final: a lane
// whose slot lies beyond it -- and with it every later slot -- has nothing left to do.
__device__ __forceinline__ uint64_t timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int64_t ld_acquire_s64(const int64_t* p) {
  int64_t v;
  asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(THREADS, 2)
solve_stream_kernel(const DesignView d0, const FitArgs a) {
  const int lane = threadIdx.x & 31;
  const bool vec_out = vec_out_ok(a);
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.stream_ctl, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t idx = base + lane;
    int64_t row = -1;
    bool exhausted = idx >= a.rec_cap;
    if (!exhausted) {
      uint64_t t0 = 0;
      for (unsigned spins = 0;; ++spins) {
        row = ld_acquire_s64(a.rec_rows + idx);
        if (row >= 0) break;
        if (ld_acquire_u32(a.stream_ctl + 2) != 0u) {            // every producer CTA has finished
          if (idx >= *reinterpret_cast<volatile uint32_t*>(a.rec_count)) { exhausted = true; break; }
          row = ld_acquire_s64(a.rec_rows + idx);                 // queued before `done`: visible now
          if (row >= 0) break;
        }
        __nanosleep(spins < 64 ? 100 : 1000);
        if ((spins & 1023u) == 1023u) {                           // a protocol bug traps instead of hanging the GPU
          if (t0 == 0) t0 = timer_ns();
          else if (timer_ns() - t0 > 4000000000ull) __trap();
        }
      }
    }
    if (row >= 0) solve_one<false>(d0, a, nullptr, row, vec_out);
    // (the slot is reset to -1 by the closing solve_rows pass, which walks the whole list once more)
    if (__any_sync(0xffffffffu, exhausted)) break;                // slots are claimed in order: nothing lies beyond
  }
}

}  // namespace

cudaError_t launch_solve_rows(const DesignView& d, const FitArgs& a, int sm_count, cudaStream_t s, const CalMeta* cals) {
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
  return cals != nullptr ? cudaLaunchKernelEx(&cfg, solve_rows_kernel<true>, d, a, cals)
                         : cudaLaunchKernelEx(&cfg, solve_rows_kernel<false>, d, a, cals);
}

cudaError_t launch_solve_stream(const DesignView& d, const FitArgs& a, int sm_count, cudaStream_t s) {
  if (a.recs == nullptr || a.rec_cap == 0 || a.stream_ctl == nullptr) return cudaSuccess;
  // two blocks per SM in the grid: one fits beside a resident fit_tc CTA, the second takes the SM over when that CTA
  // retires and helps drain what is left
  const int64_t want = ((int64_t)a.rec_cap + THREADS - 1) / THREADS;
  const int64_t cap = (int64_t)sm_count * 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(want < cap ? want : cap));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // starts when every fit_tc CTA is resident
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, solve_stream_kernel, d, a);
}

}  // namespace mmf
