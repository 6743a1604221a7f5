// host_narrow.cpp -- lossless narrowing of float32 demand on the host, so that half the bytes cross PCIe.
//
// The reference's demand is integer valued (01-data-generator.py:304 `round`) but its schema carries it as float32
// (enriched_schema, 02:360-370).  With host-resident input the whole path is bound by the PCIe link (55 GB/s against
// 6 TB/s of HBM), so the host-buffer path of mmf_fit_forecast_f32 narrows each chunk to uint16 on a few host threads
// WHILE the previous chunk's copy is in flight, ships 2 B per value, and widens on the device (widen.cu) into the
// same float32 staging rows the kernels read.  The narrowing is exact or it is not used: a chunk in which any finite
// value is not an integer in [0, 65534] is sent as float32 like before.  NaN / Inf (missing) -> 65535.
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace mmf {

namespace {

// rows [r0, r1): returns false as soon as a value cannot be carried exactly
bool narrow_rows(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int64_t r0, int64_t r1, int32_t t,
                 bool stream_stores) {
  for (int64_t r = r0; r < r1; ++r) {
    const float* s = src + r * ld_src;
    uint16_t* d = dst + r * ld_dst;
    int32_t k = 0;
#if defined(__AVX2__)
    const __m256i expm = _mm256_set1_epi32(0x7f800000);
    const __m256i maxv = _mm256_set1_epi32(65534);
    const __m256i miss = _mm256_set1_epi32(65535);
    __m256i bad = _mm256_setzero_si256();
    // Two regimes for the destination (a page-locked slot the copy engine reads next):
    //  * small slots that stay in the last-level cache: ordinary stores -- the copy engine's reads are then served
    //    from the cache and the only DRAM traffic of the whole pass is the 4 B per value read from the caller's buffer;
    //  * slots larger than the cache: streaming stores, no read-for-ownership traffic.
    // The source is read once: non-temporal prefetch keeps it from pushing the slots out of the cache.
    const bool nt = stream_stores && (reinterpret_cast<uintptr_t>(d) & 31u) == 0;
    for (; k + 16 <= t; k += 16) {
      _mm_prefetch(reinterpret_cast<const char*>(s + k + 256), _MM_HINT_NTA);     // 1 KB ahead in the row stream
      const __m256 v0 = _mm256_loadu_ps(s + k), v1 = _mm256_loadu_ps(s + k + 8);
      const __m256i b0 = _mm256_castps_si256(v0), b1 = _mm256_castps_si256(v1);
      const __m256i nf0 = _mm256_cmpeq_epi32(_mm256_and_si256(b0, expm), expm);      // NaN / Inf
      const __m256i nf1 = _mm256_cmpeq_epi32(_mm256_and_si256(b1, expm), expm);
      const __m256i i0 = _mm256_cvttps_epi32(v0), i1 = _mm256_cvttps_epi32(v1);
      const __m256i ex0 = _mm256_castps_si256(_mm256_cmp_ps(_mm256_cvtepi32_ps(i0), v0, _CMP_EQ_OQ));
      const __m256i ex1 = _mm256_castps_si256(_mm256_cmp_ps(_mm256_cvtepi32_ps(i1), v1, _CMP_EQ_OQ));
      // in range: 0 <= i <= 65534  <=>  (unsigned) i <= 65534; min_epu32(i, max) == i
      const __m256i in0 = _mm256_cmpeq_epi32(_mm256_min_epu32(i0, maxv), i0);
      const __m256i in1 = _mm256_cmpeq_epi32(_mm256_min_epu32(i1, maxv), i1);
      const __m256i ok0 = _mm256_or_si256(nf0, _mm256_and_si256(ex0, in0));
      const __m256i ok1 = _mm256_or_si256(nf1, _mm256_and_si256(ex1, in1));
      bad = _mm256_or_si256(bad, _mm256_andnot_si256(_mm256_and_si256(ok0, ok1), _mm256_set1_epi32(-1)));
      const __m256i o0 = _mm256_blendv_epi8(i0, miss, nf0), o1 = _mm256_blendv_epi8(i1, miss, nf1);
      // packus works per 128-bit lane: {o0.lo, o1.lo | o0.hi, o1.hi} -> restore the order with a 64-bit permute
      const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(o0, o1), 0xD8);
      if (nt) _mm256_stream_si256(reinterpret_cast<__m256i*>(d + k), p);
      else _mm256_storeu_si256(reinterpret_cast<__m256i*>(d + k), p);
    }
    if (!_mm256_testz_si256(bad, bad)) return false;
#endif
    for (; k < t; ++k) {
      const float v = s[k];
      uint32_t bits;
      memcpy(&bits, &v, 4);
      if ((bits & 0x7f800000u) == 0x7f800000u) { d[k] = 65535; continue; }
      if (!(v >= 0.f && v <= 65534.f)) return false;
      const int32_t i = (int32_t)v;
      if ((float)i != v) return false;
      d[k] = (uint16_t)i;
    }
  }
  return true;
}

}  // namespace

// A small persistent pool: the narrowing of one chunk is a parallel-for over row blocks.
class NarrowPool {
 public:
  NarrowPool(int n_threads, bool pin) : stop_(false), gen_(0), pending_(0) {
    // Workers are pinned (one per CPU of the process's affinity mask, in mask order, the first CPU left to the caller)
    // only when the caller says this process has the host to itself (pin == true: one GPU visible); several processes
    // -- one per GPU -- share a NUMA node's cores and would all pick the same ones.
    std::vector<int> cpus;
    if (pin) {
      cpu_set_t set;
      if (sched_getaffinity(0, sizeof(set), &set) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
          if (CPU_ISSET(c, &set)) cpus.push_back(c);
    }
    for (int i = 0; i < n_threads; ++i) {
      const int cpu = (int)cpus.size() > i + 1 ? cpus[i + 1] : -1;
      workers_.emplace_back([this, cpu] {
        if (cpu >= 0) {
          cpu_set_t one;
          CPU_ZERO(&one);
          CPU_SET(cpu, &one);
          pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
        }
        loop();
      });
    }
  }
  ~NarrowPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_.store(true);
      gen_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return (int)workers_.size(); }

  // float32 rows -> uint16 rows; false (and dst unspecified) when some value is not exactly representable
  // between begin_call() and end_call() the workers spin on the generation counter instead of sleeping on the
  // condition variable: a call narrows hundreds of small sub-chunks back to back and a futex wake per sub-chunk
  // and worker would cost more than the narrowing itself
  void begin_call() { hot_.store(true, std::memory_order_release); { std::lock_guard<std::mutex> lk(mu_); } cv_.notify_all(); }
  void end_call() { hot_.store(false, std::memory_order_release); }

  bool run(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int64_t n, int32_t t, bool stream_stores) {
    job_ = Job{src, ld_src, dst, ld_dst, n, t, stream_stores};
    next_.store(0, std::memory_order_relaxed);
    ok_.store(true, std::memory_order_relaxed);
    pending_.store((int)workers_.size(), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);
      gen_.fetch_add(1, std::memory_order_release);
    }
    if (!hot_.load(std::memory_order_relaxed)) cv_.notify_all();
    work();                                        // the calling thread helps
    while (pending_.load(std::memory_order_acquire) != 0) _mm_pause();
    return ok_.load(std::memory_order_relaxed);
  }

 private:
  struct Job { const float* src; int64_t ld_src; uint16_t* dst; int64_t ld_dst; int64_t n; int32_t t; bool stream_stores; };
  static constexpr int64_t BLOCK = 64;             // rows per grab: ~280 KB of float32 at T = 1,095

  void work() {
    const Job j = job_;
    for (;;) {
      const int64_t r0 = next_.fetch_add(BLOCK, std::memory_order_relaxed);
      if (r0 >= j.n || !ok_.load(std::memory_order_relaxed)) break;
      const int64_t r1 = r0 + BLOCK < j.n ? r0 + BLOCK : j.n;
      if (!narrow_rows(j.src, j.ld_src, j.dst, j.ld_dst, r0, r1, j.t, j.stream_stores)) ok_.store(false, std::memory_order_relaxed);
    }
    _mm_sfence();                                  // streaming stores are globally visible before the copy is enqueued
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      // hot: spin (a call is in progress, the next sub-chunk is microseconds away); cold: sleep
      int spins = 0;
      while (gen_.load(std::memory_order_acquire) == seen) {
        if (hot_.load(std::memory_order_relaxed) && spins < (1 << 20)) { _mm_pause(); ++spins; continue; }
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen || hot_.load(std::memory_order_relaxed); });
        spins = 0;
      }
      seen = gen_.load(std::memory_order_acquire);
      if (stop_) return;
      work();
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }

  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<bool> stop_;
  std::atomic<uint64_t> gen_;
  std::atomic<int> pending_;
  std::atomic<bool> hot_{false};
  Job job_{};
  std::atomic<int64_t> next_{0};
  std::atomic<bool> ok_{true};
};

NarrowPool* narrow_pool_create(int n_threads, bool pin) { return new NarrowPool(n_threads > 0 ? n_threads : 1, pin); }
void narrow_pool_destroy(NarrowPool* p) { delete p; }
int narrow_pool_size(const NarrowPool* p) { return p ? p->size() + 1 : 0; }
bool narrow_f32_to_u16(NarrowPool* p, const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int64_t n, int32_t t,
                       bool stream_stores) {
  return p->run(src, ld_src, dst, ld_dst, n, t, stream_stores);
}
void narrow_pool_begin_call(NarrowPool* p) { p->begin_call(); }
void narrow_pool_end_call(NarrowPool* p) { p->end_call(); }

}  // namespace mmf
