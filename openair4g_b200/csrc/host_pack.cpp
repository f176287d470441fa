// Host side of the narrow input feed: range-check + pack int16 soft bits to int8 on the caller's CPU cores before the
// PCIe copy.  The host-buffer path of the decoder is bound by the link (36.9 KB of int16 LLRs per 6144 decoded bits);
// soft bits that fit 8 bits -- the common case: the reference's own 8-bit decoder and demappers work in that range --
// cross it at half the bytes.  The GPU's demultiplexing kernel widens them again, so the decoder sees the same values.
// A part whose values do not all fit is sent as int16 (the pack pass reports it); nothing is ever clipped.
//
// Measured on the B200 box's 24-vCPU host (tools/src/host_pack_probe.c, profiles/r2b_host_pack_probe.txt): 12 threads
// check + pack 93 GB/s of input, 16 threads 96 GB/s, against 52-55 GB/s for the int16 copy itself -- when nothing else
// uses the host's memory (see host_pack_threads below for what happens next to the copy engine).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <immintrin.h>

namespace oai {

// returns nonzero if some value of src[0..n) does not fit int8; dst receives the low bytes either way
__attribute__((target("avx2"))) static int pack_avx2(const int16_t* src, int8_t* dst, size_t n) {
  __m256i acc = _mm256_setzero_si256();
  size_t i = 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(dst) & 31) == 0);
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 16));
    const __m256i pk = _mm256_permute4x64_epi64(_mm256_packs_epi16(a, b), 0xD8);
    // exact iff sign-extending the packed bytes gives the inputs back
    const __m256i lo = _mm256_cvtepi8_epi16(_mm256_castsi256_si128(pk)), hi = _mm256_cvtepi8_epi16(_mm256_extracti128_si256(pk, 1));
    acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_xor_si256(a, lo), _mm256_xor_si256(b, hi)));
    if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), pk);
    else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), pk);
  }
  int bad = !_mm256_testz_si256(acc, acc);
  for (; i < n; ++i) { dst[i] = (int8_t)src[i]; bad |= (src[i] != (int16_t)(int8_t)src[i]); }
  _mm_sfence();
  return bad;
}
static int pack_scalar(const int16_t* src, int8_t* dst, size_t n) {
  int bad = 0;
  for (size_t i = 0; i < n; ++i) { dst[i] = (int8_t)src[i]; bad |= (src[i] != (int16_t)(int8_t)src[i]); }
  return bad;
}

class PackPool {
 public:
  explicit PackPool(int nthreads) : n_(nthreads) {
    for (int t = 0; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~PackPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int threads() const { return n_; }
  // one job at a time (callers serialise on run_mu_): src[0..n) -> dst, split into cache-line multiples
  int run(const int16_t* src, int8_t* dst, size_t n) {
    std::lock_guard<std::mutex> job(run_mu_);
    src_ = src; dst_ = dst; total_ = n; bad_.store(0);
    { std::lock_guard<std::mutex> lk(mu_); pending_ = n_; ++gen_; }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    return bad_.load();
  }

 private:
  void loop(int t) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      const size_t chunk = ((total_ + n_ - 1) / n_ + 63) & ~(size_t)63;
      const size_t lo = std::min(total_, chunk * t), hi = std::min(total_, lo + chunk);
      if (hi > lo) {
        const int bad = avx2 ? pack_avx2(src_ + lo, dst_ + lo, hi - lo) : pack_scalar(src_ + lo, dst_ + lo, hi - lo);
        if (bad) bad_.store(1);
      }
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
  const int16_t* src_ = nullptr; int8_t* dst_ = nullptr; size_t total_ = 0;
  std::atomic<int> bad_{0};
};

// Number of pack threads = OAI_TURBO_PACK_THREADS; unset or 0: the narrow feed is OFF (the default).  On the measured box
// the pass reaches 93-105 GB/s of input on 12-24 threads when it runs alone, but only 53-69 GB/s while the copy engine
// reads the same host memory, which makes the whole host-fed decode no faster than sending int16 (28-35 ms against
// 32.4 ms per 42 624 blocks, tools/e2e_pack_probe.py, profiles/r2c_narrow_feed_probe.txt): the host's memory system, not
// the link, is then the limit.  It remains an opt-in for hosts with more memory bandwidth per GPU.
int host_pack_threads() {
  const char* e = getenv("OAI_TURBO_PACK_THREADS");
  return (e && *e) ? std::max(0, std::min(64, atoi(e))) : 0;
}

// Packs src[0..n) into dst; returns 0 when every value fits int8 (dst is then an exact copy), 1 otherwise, -1 when the
// narrow feed is switched off.
int host_pack_i16_to_i8(const int16_t* src, int8_t* dst, size_t n) {
  const int nt = host_pack_threads();
  if (nt <= 0) return -1;
  static PackPool* pool = new PackPool(nt);     // sized at first use; lives for the process (workers are never joined)
  return pool->run(src, dst, n);
}

}  // namespace oai
