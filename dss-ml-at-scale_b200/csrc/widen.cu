// widen.cu -- integer demand columns -> the float32 series rows the fit kernels read.
//
// The reference's demand is integer valued (01-data-generator.py:304 `round`) but travels as float32
// (enriched_schema, 02:360-370).  With host-resident input the whole path is PCIe-bound, so an int16 / uint16
// (or int32) value column halves (keeps) the bytes that cross the link; this kernel widens a staged chunk on the
// device: y32[i, t] = (float) y_int[i, t], the type's sentinel -> NaN (= missing, like a float NaN).
// Exact for every representable value (|v| < 2^24 for int32), so the forecasts are bit-equal to a float32 ingest.
// Bound: HBM, 2 (4) B read + 4 B written per value -- noise next to the PCIe copy that feeds it.
#include "mmf_internal.cuh"

namespace mmf {
namespace {

constexpr int TPB = 256;

template <typename T> struct Sentinel;
template <> struct Sentinel<int16_t>  { static constexpr int32_t v = -32768; };
template <> struct Sentinel<uint16_t> { static constexpr int32_t v = 65535; };
template <> struct Sentinel<int32_t>  { static constexpr int32_t v = INT32_MIN; };

template <typename T>
__device__ __forceinline__ float widen_one(T x) {
  return static_cast<int32_t>(x) == Sentinel<T>::v ? __int_as_float(0x7fc00000) : static_cast<float>(x);
}

// 8 values per 16-B (int16) / 32-B (int32) load, two 16-B stores; every thread keeps UNROLL independent loads in
// flight (a single load per trip left the kernel latency-bound at 1.3 TB/s: profiles/r02/ncu_widen.txt).  Rows are
// walked with a pitch, so a group of 8 never straddles a row; the tail of a row (t % 8) takes the scalar path.
constexpr int UNROLL = 4;

template <typename T>
__global__ void __launch_bounds__(TPB)
widen_kernel(const T* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst, int64_t n, int32_t t,
             int vec_ok) {
  const int groups = (t + 7) >> 3;
  const int64_t total = n * groups;
  const int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t i0 = (int64_t)blockIdx.x * TPB + threadIdx.x; i0 < total; i0 += stride * UNROLL) {
    T v[UNROLL][8];
    int64_t row[UNROLL];
    int c0[UNROLL];
    bool vec[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      row[u] = i < total ? i / groups : -1;
      c0[u] = row[u] >= 0 ? (int)(i - row[u] * groups) << 3 : 0;
      vec[u] = row[u] >= 0 && vec_ok && c0[u] + 8 <= t;
      if (vec[u]) {
        const T* __restrict__ s = src + row[u] * ld_src + c0[u];
        if (sizeof(T) == 2) {
          *reinterpret_cast<uint4*>(v[u]) = __ldcs(reinterpret_cast<const uint4*>(s));
        } else {
          reinterpret_cast<uint4*>(v[u])[0] = __ldcs(reinterpret_cast<const uint4*>(s));
          reinterpret_cast<uint4*>(v[u])[1] = __ldcs(reinterpret_cast<const uint4*>(s) + 1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (row[u] < 0) continue;
      float* __restrict__ d = dst + row[u] * ld_dst + c0[u];
      if (vec[u]) {
        __stcs(reinterpret_cast<float4*>(d), make_float4(widen_one(v[u][0]), widen_one(v[u][1]), widen_one(v[u][2]), widen_one(v[u][3])));
        __stcs(reinterpret_cast<float4*>(d + 4), make_float4(widen_one(v[u][4]), widen_one(v[u][5]), widen_one(v[u][6]), widen_one(v[u][7])));
      } else {
        const T* __restrict__ s = src + row[u] * ld_src + c0[u];
        for (int k = 0; k < 8 && c0[u] + k < t; ++k) d[k] = widen_one(s[k]);
      }
    }
  }
}

template <typename T>
cudaError_t launch(const void* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t n, int32_t t, int sm, cudaStream_t s) {
  if (n <= 0 || t <= 0) return cudaSuccess;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (ld_src * sizeof(T)) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && ld_dst % 4 == 0) ? 1 : 0;
  const int64_t total = n * ((t + 7) >> 3);
  const int64_t want = (total + (int64_t)TPB * UNROLL - 1) / ((int64_t)TPB * UNROLL), cap = (int64_t)sm * 16;
  widen_kernel<T><<<(unsigned)(want < cap ? want : cap), TPB, 0, s>>>(static_cast<const T*>(src), ld_src, dst, ld_dst, n, t, vec_ok);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_widen(int dtype, const void* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t n, int32_t t,
                         int sm_count, cudaStream_t s) {
  switch (dtype) {
    case MMF_DT_I16: return launch<int16_t>(src, ld_src, dst, ld_dst, n, t, sm_count, s);
    case MMF_DT_U16: return launch<uint16_t>(src, ld_src, dst, ld_dst, n, t, sm_count, s);
    case MMF_DT_I32: return launch<int32_t>(src, ld_src, dst, ld_dst, n, t, sm_count, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mmf
