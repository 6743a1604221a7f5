// pack.cu -- device-side packer: long-format (key, Date, Demand) rows -> padded series y[N, T] on the GPU.
//
// Replaces, for all groups at once, what Spark + pandas do before the per-group fit in the reference:
// the hash shuffle of `repartition(n_tasks, "Product", "SKU")` + `groupBy` (02:525-526) and the per-group
// `sort_values("Date")` + `set_index("Date").asfreq(freq)` (02:422-423).  Inputs are Arrow column buffers
// copied to the device as they are (utf8 offsets + bytes or dictionary indices for the keys, date32 days,
// float32 values); nothing is sorted by the caller.
//   1. hash_utf8 / hash_i32      64-bit FNV-1a of each row's key columns (chained across columns)
//   2. group codes               cub radix sort of (hash,row) -> segment heads -> dense codes 0..G-1,
//                                first row of every group (to recover its key strings on the host)
//   3. minmax                    per-group first / last day (warp-aggregated atomics)
//   4. scatter                   y[row_of_group[g], (day - start_g) / step] = value   (NaN-prefilled)
// Every step is HBM-bound integer / byte work: coalesced streaming reads, one scattered 4-B write per row.
#include <cub/cub.cuh>

#include "mmf_internal.cuh"

namespace mmf {
namespace {

constexpr int TPB = 256;
constexpr unsigned PACK_FILL_BITS = 0x7fc00000u;   // the quiet NaN fill_nan_kernel writes: a cell no row has written yet

__device__ __forceinline__ uint64_t fnv1a_byte(uint64_t h, uint32_t b) { return (h ^ b) * 1099511628211ull; }

// first == 1: the standard FNV offset basis; first > 1: a different basis per value, for the re-hash after a
// detected collision (verify_* below)
__device__ __forceinline__ uint64_t fnv_basis(int first) {
  return 14695981039346656037ull ^ ((uint64_t)(first - 1) * 0x9E3779B97F4A7C15ull);
}

__global__ void hash_utf8_kernel(const int32_t* __restrict__ offsets, const uint8_t* __restrict__ data, int64_t n,
                                 uint64_t* __restrict__ h, int first) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
    uint64_t x = first ? fnv_basis(first) : h[i];
    const int32_t lo = offsets[i], hi = offsets[i + 1];
    for (int32_t k = lo; k < hi; ++k) x = fnv1a_byte(x, data[k]);
    x = fnv1a_byte(x, 0xffu);                       // column separator: ("ab","c") != ("a","bc")
    h[i] = x;
  }
}

__global__ void hash_i32_kernel(const int32_t* __restrict__ v, int64_t n, uint64_t* __restrict__ h, int first) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
    uint64_t x = first ? fnv_basis(first) : h[i];
    const uint32_t u = static_cast<uint32_t>(v[i]);
    x = fnv1a_byte(x, u & 0xffu); x = fnv1a_byte(x, (u >> 8) & 0xffu);
    x = fnv1a_byte(x, (u >> 16) & 0xffu); x = fnv1a_byte(x, u >> 24);
    x = fnv1a_byte(x, 0xffu);
    h[i] = x;
  }
}

__global__ void iota_kernel(int32_t* __restrict__ v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) v[i] = (int32_t)i;
}

__global__ void heads_kernel(const uint64_t* __restrict__ sorted_h, int64_t n, int32_t* __restrict__ head) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB)
    head[i] = (i == 0 || sorted_h[i] != sorted_h[i - 1]) ? 1 : 0;
}

// code of sorted position i = (inclusive scan of heads)[i] - 1 ; scatter back to the original row order
__global__ void codes_kernel(const int32_t* __restrict__ scan, const int32_t* __restrict__ head,
                             const int32_t* __restrict__ sorted_row, int64_t n, int32_t* __restrict__ gid,
                             int32_t* __restrict__ first_row) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
    const int32_t g = scan[i] - 1;
    gid[sorted_row[i]] = g;
    if (head[i]) first_row[g] = sorted_row[i];
  }
}

// rows grouped by hash only: check every row's key against the key of its group's first row (a 64-bit collision
// would silently merge two series).  One coalesced pass over the key column + a gather of the group head's key.
__global__ void verify_utf8_kernel(const int32_t* __restrict__ offsets, const uint8_t* __restrict__ data, int64_t n,
                                   const int32_t* __restrict__ gid, const int32_t* __restrict__ first_row,
                                   unsigned long long* __restrict__ mismatches) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
    const int32_t j = first_row[gid[i]];
    if (j == i) continue;
    const int32_t lo = offsets[i], len = offsets[i + 1] - lo, lo2 = offsets[j];
    bool same = (offsets[j + 1] - lo2) == len;
    for (int32_t k = 0; same && k < len; ++k) same = data[lo + k] == data[lo2 + k];
    if (!same) atomicAdd(mismatches, 1ull);
  }
}

__global__ void verify_i32_kernel(const int32_t* __restrict__ v, int64_t n, const int32_t* __restrict__ gid,
                                  const int32_t* __restrict__ first_row, unsigned long long* __restrict__ mismatches) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB)
    if (v[i] != v[first_row[gid[i]]]) atomicAdd(mismatches, 1ull);
}

__global__ void minmax_kernel(const int32_t* __restrict__ gid, const int32_t* __restrict__ day, int64_t n,
                              int32_t* __restrict__ gmin, int32_t* __restrict__ gmax) {
  const int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t i0 = (int64_t)blockIdx.x * TPB; i0 < n; i0 += stride) {
    const int64_t i = i0 + threadIdx.x;
    const bool live = i < n;
    const int32_t g = live ? gid[i] : -1;
    const int32_t d = live ? day[i] : 0;
    // rows of one group are usually adjacent: one atomic per (warp, group) instead of one per row
    const unsigned peers = __match_any_sync(0xffffffffu, g);
    const int32_t lo = __reduce_min_sync(peers, d);
    const int32_t hi = __reduce_max_sync(peers, d);
    if (live && (threadIdx.x & 31) == (__ffs(peers) - 1)) {
      atomicMin(gmin + g, lo);
      atomicMax(gmax + g, hi);
    }
  }
}

__global__ void fill_nan_kernel(float4* __restrict__ y, int64_t n4) {
  const float qn = __int_as_float(0x7fc00000);
  const float4 v = make_float4(qn, qn, qn, qn);
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n4; i += (int64_t)gridDim.x * TPB) __stcs(y + i, v);
}

__global__ void scatter_kernel(const int32_t* __restrict__ gid, const int32_t* __restrict__ day,
                               const float* __restrict__ val, int64_t n, const int64_t* __restrict__ row_of_group,
                               const int32_t* __restrict__ gstart, int32_t step, float* __restrict__ y, int64_t ld_y,
                               int32_t t_len, unsigned long long* __restrict__ dups) {
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB) {
    const int32_t g = gid[i];
    const int64_t r = row_of_group[g];
    if (r < 0) continue;                            // group belongs to another calendar bucket
    const int32_t off = day[i] - gstart[g];
    if (off < 0 || off % step != 0) continue;       // off-grid rows vanish, like asfreq (02:423)
    const int32_t t = off / step;
    if (t >= t_len) continue;
    if (dups == nullptr) { y[r * ld_y + t] = val[i]; continue; }
    // duplicate (key, date) rows: the reference's asfreq raises on them (02:423).  Cells start as the fill NaN; an
    // exchange that returns anything else means another row already landed here (a first row whose own value is that
    // NaN is the one case this cannot see -- and there both rows mean "missing or overwritten" anyway).
    const unsigned old = atomicExch(reinterpret_cast<unsigned*>(y + r * ld_y + t), __float_as_uint(val[i]));
    if (old != PACK_FILL_BITS) atomicAdd(dups, 1ull);
  }
}

unsigned grid_for(int64_t n, int sm_count) {
  const int64_t want = (n + TPB - 1) / TPB;
  const int64_t cap = (int64_t)sm_count * 16;
  return (unsigned)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace

cudaError_t pack_hash_utf8(const int32_t* offsets, const uint8_t* data, int64_t n, uint64_t* h, int first, int sm,
                           cudaStream_t s) {
  if (n > 0) hash_utf8_kernel<<<grid_for(n, sm), TPB, 0, s>>>(offsets, data, n, h, first);
  return cudaGetLastError();
}
cudaError_t pack_hash_i32(const int32_t* v, int64_t n, uint64_t* h, int first, int sm, cudaStream_t s) {
  if (n > 0) hash_i32_kernel<<<grid_for(n, sm), TPB, 0, s>>>(v, n, h, first);
  return cudaGetLastError();
}

cudaError_t pack_verify_utf8(const int32_t* offsets, const uint8_t* data, int64_t n, const int32_t* gid,
                             const int32_t* first_row, uint64_t* mismatches, int sm, cudaStream_t s) {
  if (n > 0) verify_utf8_kernel<<<grid_for(n, sm), TPB, 0, s>>>(offsets, data, n, gid, first_row,
                                                                reinterpret_cast<unsigned long long*>(mismatches));
  return cudaGetLastError();
}
cudaError_t pack_verify_i32(const int32_t* v, int64_t n, const int32_t* gid, const int32_t* first_row,
                            uint64_t* mismatches, int sm, cudaStream_t s) {
  if (n > 0) verify_i32_kernel<<<grid_for(n, sm), TPB, 0, s>>>(v, n, gid, first_row,
                                                               reinterpret_cast<unsigned long long*>(mismatches));
  return cudaGetLastError();
}

// Scratch layout of pack_group_codes (the caller owns the buffer: no cudaMalloc / cudaFree on the packing path).
namespace {
struct CodesScratch {
  size_t h_sorted, rows, rows_sorted, head, scan, tmp, tmp_bytes, total;
};
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
cudaError_t codes_scratch_layout(int64_t n, CodesScratch* L) {
  size_t tmp_sort = 0, tmp_scan = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 64);
  if (e != cudaSuccess) return e;
  e = cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n);
  if (e != cudaSuccess) return e;
  size_t off = 0;
  L->h_sorted = off;    off += align256((size_t)n * sizeof(uint64_t));
  L->rows = off;        off += align256((size_t)n * sizeof(int32_t));
  L->rows_sorted = off; off += align256((size_t)n * sizeof(int32_t));
  L->head = off;        off += align256((size_t)n * sizeof(int32_t));
  L->scan = off;        off += align256((size_t)n * sizeof(int32_t));
  L->tmp = off;
  L->tmp_bytes = std::max(tmp_sort, tmp_scan);
  L->total = off + align256(L->tmp_bytes);
  return cudaSuccess;
}
}  // namespace

cudaError_t pack_group_codes_scratch_bytes(int64_t n, size_t* bytes) {
  CodesScratch L{};
  *bytes = 0;
  if (n <= 0) return cudaSuccess;
  cudaError_t e = codes_scratch_layout(n, &L);
  *bytes = L.total;
  return e;
}

// hash[n] -> gid[n] (dense codes in hash order), first_row[>= n_groups], *n_groups.  Synchronises once (to read G).
// `scratch` holds pack_group_codes_scratch_bytes(n) bytes of device memory.
cudaError_t pack_group_codes(const uint64_t* h, int64_t n, int32_t* gid, int32_t* first_row, int32_t* n_groups_host,
                             void* scratch, int sm, cudaStream_t s) {
  *n_groups_host = 0;
  if (n <= 0) return cudaSuccess;
  CodesScratch L{};
  cudaError_t e = codes_scratch_layout(n, &L);
  if (e != cudaSuccess) return e;
  char* base = static_cast<char*>(scratch);
  uint64_t* h_sorted = reinterpret_cast<uint64_t*>(base + L.h_sorted);
  int32_t* rows = reinterpret_cast<int32_t*>(base + L.rows);
  int32_t* rows_sorted = reinterpret_cast<int32_t*>(base + L.rows_sorted);
  int32_t* head = reinterpret_cast<int32_t*>(base + L.head);
  int32_t* scan = reinterpret_cast<int32_t*>(base + L.scan);
  void* tmp = base + L.tmp;
#define PK_TRY(x) do { e = (x); if (e != cudaSuccess) return e; } while (0)
  iota_kernel<<<grid_for(n, sm), TPB, 0, s>>>(rows, n);
  size_t t1 = L.tmp_bytes;
  PK_TRY(cub::DeviceRadixSort::SortPairs(tmp, t1, h, h_sorted, rows, rows_sorted, (int)n, 0, 64, s));
  heads_kernel<<<grid_for(n, sm), TPB, 0, s>>>(h_sorted, n, head);
  size_t t2 = L.tmp_bytes;
  PK_TRY(cub::DeviceScan::InclusiveSum(tmp, t2, head, scan, (int)n, s));
  codes_kernel<<<grid_for(n, sm), TPB, 0, s>>>(scan, head, rows_sorted, n, gid, first_row);
  PK_TRY(cudaGetLastError());
  PK_TRY(cudaMemcpyAsync(n_groups_host, scan + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  PK_TRY(cudaStreamSynchronize(s));
#undef PK_TRY
  return cudaSuccess;
}

cudaError_t pack_minmax(const int32_t* gid, const int32_t* day, int64_t n, int32_t n_groups, int32_t* gmin,
                        int32_t* gmax, int sm, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(gmin, 0x7f, (size_t)n_groups * sizeof(int32_t), s);     // 0x7f7f7f7f: > any date32
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(gmax, 0x80, (size_t)n_groups * sizeof(int32_t), s);                 // 0x80808080: < any date32
  if (e != cudaSuccess) return e;
  if (n > 0) minmax_kernel<<<grid_for(n, sm), TPB, 0, s>>>(gid, day, n, gmin, gmax);
  return cudaGetLastError();
}

cudaError_t pack_scatter(const int32_t* gid, const int32_t* day, const float* val, int64_t n,
                         const int64_t* row_of_group, const int32_t* gstart, int32_t step, float* y, int64_t n_rows,
                         int64_t ld_y, int32_t t_len, unsigned long long* dups, int sm, cudaStream_t s) {
  const int64_t n4 = n_rows * ld_y / 4;             // ld_y is a multiple of 4 floats and y is 16-B aligned
  if (n4 > 0) fill_nan_kernel<<<grid_for(n4, sm), TPB, 0, s>>>(reinterpret_cast<float4*>(y), n4);
  if (n > 0) scatter_kernel<<<grid_for(n, sm), TPB, 0, s>>>(gid, day, val, n, row_of_group, gstart, step, y, ld_y, t_len, dups);
  return cudaGetLastError();
}

}  // namespace mmf
