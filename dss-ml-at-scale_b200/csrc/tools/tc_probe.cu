// tc_probe.cu -- bring-up probe for the Blackwell building blocks of fit_tc.cu, one CTA, tiny shapes.
// Each mode isolates one assumption so a failure on the GPU box points at one thing:
//   mode 3  TMA 128-B swizzle formula  : LDS with (chunk ^ (row & 7)) must read back the source tile
//   mode 0  SS tcgen05.mma kind::tf32  : smem descriptors (K-major SW128), instruction descriptor, commit, TMEM ld
//   mode 1  TS tcgen05.mma             : A operand written with tcgen05.st 32x32b (lane = row, column = k)
//   mode 2  production inner sequence  : hi/lo split, N=32 then N=16 accumulating into the same columns
// Build: make tc_probe ; run on a B200:  ./tc_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../sm100_ptx.cuh"

using namespace sm100;

struct Maps {
  alignas(64) unsigned char a[128];
  alignas(64) unsigned char b[128];
};

__global__ void __launch_bounds__(160, 1) probe_kernel(const __grid_constant__ Maps maps, float* __restrict__ out,
                                                      int mode) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t s_a = smem_u32(smem);              // 128 x 128 B
  const uint32_t s_b = s_a + 16384;                 // 32 x 128 B
  const uint32_t bars = s_b + 4096;
  const uint32_t bar_full = bars, bar_mma = bars + 8, bar_a = bars + 16;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + 16384 + 4096 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4) {
    if (lane == 0) {
      mbar_init(bar_full, 1);
      mbar_init(bar_mma, 1);
      mbar_init(bar_a, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 4 && lane == 0) {
    mbar_expect_tx(bar_full, 16384 + 4096);
    tma_load_2d(s_a, maps.a, bar_full, 0, 0, L2_EVICT_NORMAL);
    tma_load_2d(s_b, maps.b, bar_full, 0, 0, L2_EVICT_NORMAL);
  }
  mbar_wait(bar_full, 0);

  constexpr uint32_t IDN32 = umma_idesc_tf32(128, 32);
  constexpr uint32_t IDN16 = umma_idesc_tf32(128, 16);

  if (mode == 3) {
    if (warp < 4) {
      const int r = threadIdx.x;
      for (int q = 0; q < 8; ++q) {
        const float4 v = lds128(s_a + r * 128 + ((q ^ (r & 7)) << 4));
        out[r * 32 + q * 4 + 0] = v.x; out[r * 32 + q * 4 + 1] = v.y;
        out[r * 32 + q * 4 + 2] = v.z; out[r * 32 + q * 4 + 3] = v.w;
      }
    }
  } else {
    if (warp < 4 && mode != 0) {
      const int r = threadIdx.x;
      uint32_t hi[32], lo[32];
      for (int q = 0; q < 8; ++q) {
        const float4 v = lds128(s_a + r * 128 + ((q ^ (r & 7)) << 4));
        const float e[4] = {v.x, v.y, v.z, v.w};
        for (int w = 0; w < 4; ++w) {
          if (mode == 1) {
            hi[q * 4 + w] = __float_as_uint(e[w]);
            lo[q * 4 + w] = 0u;
          } else {
            const uint32_t h = __float_as_uint(e[w]) & 0xFFFFE000u;
            hi[q * 4 + w] = h;
            lo[q * 4 + w] = __float_as_uint(e[w] - __uint_as_float(h));
          }
        }
      }
      const uint32_t la = static_cast<uint32_t>(warp * 32) << 16;
      tmem_st_32x32b_x32(tmem + la + 32, hi);
      if (mode == 2) {
        // reuse columns 32..63 for hi, and the upper half of the accumulator allocation is not available
        // for lo in a 64-column allocation -> store lo after the hi MMAs were issued (see issuer below)
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a);
      if (mode == 2) {
        // second phase: wait for the hi MMAs to retire, then overwrite the slot with lo
        mbar_wait(bar_mma, 0);
        tc_fence_after();
        tmem_st_32x32b_x32(tmem + la + 32, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a);
      }
    }
    if (warp == 4 && lane == 0) {
      if (mode == 0) {
        for (int k = 0; k < 4; ++k)
          umma_tf32_ss(tmem, umma_desc_k_sw128(s_a + k * 32), umma_desc_k_sw128(s_b + k * 32), IDN32, k ? 1u : 0u);
        umma_commit(bar_mma);
      } else if (mode == 1) {
        mbar_wait(bar_a, 0);
        tc_fence_after();
        for (int k = 0; k < 4; ++k)
          umma_tf32_ts(tmem, tmem + 32 + k * 8, umma_desc_k_sw128(s_b + k * 32), IDN32, k ? 1u : 0u);
        umma_commit(bar_mma);
      } else {
        mbar_wait(bar_a, 0);
        tc_fence_after();
        for (int k = 0; k < 4; ++k)
          umma_tf32_ts(tmem, tmem + 32 + k * 8, umma_desc_k_sw128(s_b + k * 32), IDN32, k ? 1u : 0u);
        umma_commit(bar_mma);          // phase 0: hi MMAs done
        mbar_wait(bar_a, 1);
        tc_fence_after();
        for (int k = 0; k < 4; ++k)
          umma_tf32_ts(tmem, tmem + 32 + k * 8, umma_desc_k_sw128(s_b + k * 32), IDN16, 1u);
        umma_commit(bar_mma);          // phase 1: lo MMAs done
      }
    }
    if (warp < 4) {
      mbar_wait(bar_mma, mode == 2 ? 1 : 0);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), acc);
      tmem_wait_ld();
      tc_fence_before();
      for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = __uint_as_float(acc[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}


// ---- timing probe: cycles per tcgen05.mma (M=128, K=8 tf32) for dependent vs independent accumulators,
// issued the production way: warp-converged, one elected lane, descriptors in uniform registers.
template <int NCOLS, int NACC, bool SS>
__global__ void __launch_bounds__(160, 1) timing_kernel(long long* __restrict__ cycles, int niter) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const uint32_t s_a = smem_u32(smem);
  const uint32_t s_b = s_a + 16384;
  const uint32_t bar = s_b + 8192;
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + 16384 + 8192 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 8192) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 4) {
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (warp < 4) {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t la = static_cast<uint32_t>(warp * 32) << 16;
    for (int c = 0; c < 512; c += 32) tmem_st_32x32b_x32(tmem + la + c, z);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, NCOLS);
    constexpr int STRIDE = NCOLS > 32 ? 64 : 32;
    const uint64_t bd0 = umma_desc_k_sw128(s_b);
    const uint64_t ad0 = umma_desc_k_sw128(s_a);
    const long long t0 = clock64();
    for (int it = 0; it < niter; it += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t d = tmem + (j % NACC) * STRIDE;
        const uint64_t bd = bd0 + static_cast<uint64_t>((j & 3) * 2);      // +32 B in 16-B units
        if (SS) umma_tf32_ss_elect(d, ad0 + static_cast<uint64_t>((j & 3) * 2), bd, idesc, 1u);
        else umma_tf32_ts_elect(d, tmem + 256 + (j & 3) * 8, bd, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    umma_commit_elect(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NCOLS, int NACC, bool SS>
int run_timing(long long* dcy, int niter) {
  long long h[2];
  cudaFuncSetAttribute(timing_kernel<NCOLS, NACC, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  timing_kernel<NCOLS, NACC, SS><<<1, 160, 40960>>>(dcy, niter);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("timing kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
  cudaMemcpy(h, dcy, 16, cudaMemcpyDeviceToHost);
  printf("  %s N=%3d accumulators=%d : %6.1f | %6.1f\n", SS ? "SS" : "TS", NCOLS, NACC, (double)h[0] / niter, (double)h[1] / niter);
  return 0;
}


// ---- read-bandwidth probes: what can a pure read stream reach on this part? ---------------------------
// (a) TMA: per CTA one producer warp streams {32 x 128} fp32 boxes through an 8-stage ring, a consumer warp
//     only releases the stages.  (b) LDG.128: grid-stride streaming loads, 8 in flight per thread.
__global__ void __launch_bounds__(64, 1) tma_read_kernel(const __grid_constant__ Maps maps, int n_tiles, int n_chunks,
                                                        float* __restrict__ sink) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  constexpr int ST = 8;
  const uint32_t s_y = smem_u32(smem);
  const uint32_t bars = s_y + ST * 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (ST + i), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  int stage = 0; uint32_t phase = 0;
  if (warp == 0) {
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(bars + 8 * (ST + stage), phase ^ 1u);
        tma_load_2d_elect(bars + 8 * stage, 16384, s_y + stage * 16384, maps.a, ch * 32, tile * 128, L2_EVICT_FIRST);
        if (++stage == ST) { stage = 0; phase ^= 1u; }
      }
  } else {
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int ch = 0; ch < n_chunks; ++ch) {
        mbar_wait(bars + 8 * stage, phase);
        acc += lds128(s_y + stage * 16384 + lane * 16).x;
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (ST + stage));
        if (++stage == ST) { stage = 0; phase ^= 1u; }
      }
    if (acc == 123.456f) sink[0] = acc;
  }
}

// same, but each stage takes TWO adjacent boxes of the same rows (64 t = 256 contiguous bytes per row)
__global__ void __launch_bounds__(64, 1) tma_read2_kernel(const __grid_constant__ Maps maps, int n_tiles, int n_chunks2,
                                                         float* __restrict__ sink) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  constexpr int ST = 5;
  const uint32_t s_y = smem_u32(smem);
  const uint32_t bars = s_y + ST * 32768;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (ST + i), 1); }
    fence_mbar_init();
  }
  __syncthreads();
  int stage = 0; uint32_t phase = 0;
  if (warp == 0) {
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int ch = 0; ch < n_chunks2; ++ch) {
        mbar_wait(bars + 8 * (ST + stage), phase ^ 1u);
        tma_load_2d_x2_elect(bars + 8 * stage, 32768, s_y + stage * 32768, maps.a, ch * 64, tile * 128, L2_EVICT_FIRST,
                             s_y + stage * 32768 + 16384, maps.a, ch * 64 + 32, tile * 128, L2_EVICT_FIRST);
        if (++stage == ST) { stage = 0; phase ^= 1u; }
      }
  } else {
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      for (int ch = 0; ch < n_chunks2; ++ch) {
        mbar_wait(bars + 8 * stage, phase);
        acc += lds128(s_y + stage * 32768 + lane * 16).x;
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (ST + stage));
        if (++stage == ST) { stage = 0; phase ^= 1u; }
      }
    if (acc == 123.456f) sink[0] = acc;
  }
}

__global__ void __launch_bounds__(512) ldg_read_kernel(const float4* __restrict__ p, size_t n4, float* __restrict__ sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n4; i += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (; i < n4; i += stride) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
  if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32_trunc(float v) {
  uint32_t b; memcpy(&b, &v, 4); b &= 0xFFFFE000u; memcpy(&v, &b, 4); return v;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaFree(0));
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const int M = 128, N = 32, K = 32;
  std::vector<float> A(M * K), B(N * K), Araw(M * K);
  srand(7);
  for (auto& v : Araw) v = 1000.f * ((rand() / (float)RAND_MAX) - 0.5f);
  for (int i = 0; i < M * K; ++i) A[i] = tf32_trunc(Araw[i]);
  for (auto& v : B) v = tf32_trunc(2.f * (rand() / (float)RAND_MAX) - 1.f);
  float *dA, *dB, *dOut;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dOut, M * 32 * 4));
  Maps maps;
  auto encode = [&](void* out, void* g, uint64_t inner, uint64_t outer, uint32_t bi, uint32_t bo) {
    cuuint64_t dims[2] = {inner, outer}; cuuint64_t str[1] = {inner * 4}; cuuint32_t box[2] = {bi, bo}; cuuint32_t es[2] = {1, 1};
    return enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, str, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  if (encode(maps.a, dA, K, M, 32, 128) != CUDA_SUCCESS || encode(maps.b, dB, K, N, 32, 32) != CUDA_SUCCESS) {
    printf("tensor map encode failed\n"); return 2;
  }
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  int bad_total = 0;
  const int modes[4] = {3, 0, 1, 2};
  for (int mi = 0; mi < 4; ++mi) {
    const int mode = modes[mi];
    const std::vector<float>& Asrc = (mode == 2) ? Araw : A;
    CK(cudaMemcpy(dA, Asrc.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0xff, M * 32 * 4));
    probe_kernel<<<1, 160, 32768>>>(maps, dOut, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 3; }
    std::vector<float> out(M * 32);
    CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int j = 0; j < 32; ++j) {
        double ref;
        if (mode == 3) ref = Asrc[m * K + j];
        else if (mode == 2) {
          // cols 0..15: full-precision dot with B rows 0..15 (hi+lo); cols 16..31: hi part only with B rows 16..31
          double s = 0;
          for (int k = 0; k < K; ++k) {
            const double a = (j < 16) ? (double)Asrc[m * K + k] : (double)tf32_trunc(Asrc[m * K + k]);
            s += a * B[j * K + k];
          }
          ref = s;
        } else {
          double s = 0;
          for (int k = 0; k < K; ++k) s += (double)Asrc[m * K + k] * B[j * K + k];
          ref = s;
        }
        const double err = fabs(out[m * 32 + j] - ref);
        maxerr = fmax(maxerr, err); maxref = fmax(maxref, fabs(ref));
        const double tol = (mode == 3) ? 0.0 : 1e-5 * 1000.0 * 8;
        if (!(err <= tol)) { if (bad < 4) printf("  mode %d mismatch at [%d][%d]: got %.6f want %.6f\n", mode, m, j, out[m * 32 + j], ref); ++bad; }
      }
    printf("mode %d: max|err| = %.3e (max|ref| = %.3e)  %s\n", mode, maxerr, maxref, bad ? "FAIL" : "ok");
    bad_total += bad;
  }
  {
    long long* dcy;
    CK(cudaMalloc(&dcy, 16));
    const int niter = 4096;
    printf("tcgen05.mma timing, M=128 K=8 tf32, %d MMAs, elected lane of a converged warp (issue cyc/MMA | issue+retire cyc/MMA)\n", niter);
    run_timing<16, 1, false>(dcy, niter); run_timing<16, 2, false>(dcy, niter); run_timing<16, 4, false>(dcy, niter); run_timing<16, 8, false>(dcy, niter);
    run_timing<32, 1, false>(dcy, niter); run_timing<32, 2, false>(dcy, niter); run_timing<32, 4, false>(dcy, niter); run_timing<32, 8, false>(dcy, niter);
    run_timing<64, 1, false>(dcy, niter); run_timing<64, 4, false>(dcy, niter);
    run_timing<128, 1, false>(dcy, niter);
    run_timing<16, 1, true>(dcy, niter); run_timing<16, 8, true>(dcy, niter);
    run_timing<32, 1, true>(dcy, niter); run_timing<32, 8, true>(dcy, niter);
    run_timing<64, 1, true>(dcy, niter);
  }
  {
    // pure read streams over a 1M x 1096 fp32 buffer (4.38 GB)
    const size_t n = 1000000, ld = 1096;
    float* big; float* sink;
    CK(cudaMalloc(&big, n * ld * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(big, 0, n * ld * 4));
    Maps m2;
    {
      cuuint64_t dims[2] = {1095, n}; cuuint64_t str[1] = {ld * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
      if (enc(reinterpret_cast<CUtensorMap*>(m2.a), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, big, dims, str, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 2; }
      memcpy(m2.b, m2.a, 128);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(tma_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 2048));
    for (int grid : {148, 296}) {
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        tma_read_kernel<<<grid, 64, 8 * 16384 + 2048>>>(m2, (int)((n + 127) / 128), 35, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tma_read_kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("TMA read stream, grid %d: %.3f ms, %.0f GB/s (4*1095 B per row counted)\n", grid, ms, n * 1095.0 * 4 / ms / 1e6);
      }
    }
    CK(cudaFuncSetAttribute(tma_read2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 32768 + 2048));
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      tma_read2_kernel<<<148, 64, 5 * 32768 + 2048>>>(m2, (int)((n + 127) / 128), 18, sink);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("tma_read2_kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("TMA read stream, 2 adjacent boxes per stage (64 t), grid 148: %.3f ms, %.0f GB/s\n", ms, n * 1095.0 * 4 / ms / 1e6);
    }
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      ldg_read_kernel<<<148 * 4, 512>>>(reinterpret_cast<const float4*>(big), n * ld / 4, sink);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("LDG.128 read stream: %.3f ms, %.0f GB/s (all %zu bytes counted)\n", ms, n * ld * 4.0 / ms / 1e6, n * ld * 4);
    }
    CK(cudaMemcpyAsync(big, big + n * ld / 2, n * ld * 2, cudaMemcpyDeviceToDevice));
    cudaEventRecord(e0);
    CK(cudaMemcpyAsync(big, big + n * ld / 2, n * ld * 2, cudaMemcpyDeviceToDevice));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    { float ms; cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemcpy D2D 2.19 GB: %.3f ms, %.0f GB/s (read+write)\n", ms, 2.0 * n * ld * 2 / ms / 1e6); }
  }
  printf(bad_total ? "PROBE FAIL\n" : "PROBE OK\n");
  return bad_total ? 1 : 0;
}
