// Host side of the feature ingestion: gather the per-video [T_s, C_s] fp32 arrays of a batch into the
// pinned staging buffers the H2D copy engine reads. Replaces the collate + `.to(device)` of the reference's DataLoader
// path (libs/datasets/deepfake_video_audio.py:547-558, libs/modeling/av_fd_no_recon.py:431-477) on the host.
//
// The batch moves ~77 MB per 32 videos, i.e. ~40 GB/s of host memcpy at 16k videos/s: a job for several cores. The
// copies run on a persistent thread pool inside the library (no Python task overhead, no GIL), are handed out in
// 256 KB pieces from one atomic counter (ragged spans balance themselves) and use non-temporal stores: the
// destination is read next by the DMA engine, not by a core, so write-allocate traffic (a read of every destination
// line) and the eviction of the packer's other data are avoided.
#include <emmintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "common.cuh"

namespace avdf {
namespace hostpack {

constexpr size_t PIECE = 256 << 10;
constexpr int MAX_THREADS = 64;

struct Piece { unsigned char* dst; const unsigned char* src; size_t n; };

// dst/src arbitrary alignment; streams 64 B per iteration once dst is 64-byte aligned
static void copy_nt(unsigned char* dst, const unsigned char* src, size_t n) {
  if (n < 4096) { memcpy(dst, src, n); return; }
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
  if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
  const size_t body = n & ~size_t(63);
  for (size_t i = 0; i < body; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  if (n > body) memcpy(dst + body, src + body, n - body);
}

class Pool {
 public:
  // copies every piece with up to n_threads threads (the caller is one of them); returns when all are done
  void run(const std::vector<Piece>& pieces, int n_threads) {
    std::lock_guard<std::mutex> call(call_mu_);            // one gather at a time owns the workers
    if (n_threads > MAX_THREADS) n_threads = MAX_THREADS;
    const int helpers = (int)pieces.size() < n_threads ? (int)pieces.size() - 1 : n_threads - 1;
    if (helpers > 0) {
      std::unique_lock<std::mutex> lk(mu_);
      while ((int)workers_.size() < helpers) {
        const int id = (int)workers_.size();
        workers_.emplace_back([this, id] { worker(id); });
        workers_.back().detach();
      }
      pieces_ = &pieces;
      next_.store(0, std::memory_order_relaxed);
      want_ = helpers;
      pending_ = helpers;
      ++gen_;
      lk.unlock();
      cv_work_.notify_all();
    } else {
      pieces_ = &pieces;
      next_.store(0, std::memory_order_relaxed);
    }
    drain(pieces);
    if (helpers > 0) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_done_.wait(lk, [this] { return pending_ == 0; });
    }
    _mm_sfence();
  }

 private:
  void drain(const std::vector<Piece>& pieces) {
    for (;;) {
      const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
      if (i >= pieces.size()) break;
      copy_nt(pieces[i].dst, pieces[i].src, pieces[i].n);
    }
    _mm_sfence();                                          // non-temporal stores ordered before the hand-off
  }
  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::vector<Piece>* job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (id >= want_) continue;                         // this gather uses fewer threads
        job = pieces_;
      }
      drain(*job);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  const std::vector<Piece>* pieces_ = nullptr;
  std::atomic<size_t> next_{0};
  int want_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
};

// never destroyed: its detached workers may outlive static destruction at process exit. A forked child (DataLoader
// workers fork) inherits the object but none of its threads: it gets a fresh pool.
static Pool* g_pool = nullptr;
static std::once_flag g_once;
static Pool& pool() {
  std::call_once(g_once, [] {
    g_pool = new Pool();
    pthread_atfork(nullptr, nullptr, [] { g_pool = new Pool(); });
  });
  return *g_pool;
}

}  // namespace hostpack
}  // namespace avdf

extern "C" int avdf_host_pack(const void* const* src, void* const* dst, const size_t* nbytes, int32_t n, int32_t n_threads) {
  using namespace avdf;
  using namespace avdf::hostpack;
  AVDF_CHECK_ARG(n >= 0 && n_threads >= 1, "avdf_host_pack: n >= 0 and n_threads >= 1");
  AVDF_CHECK_ARG(n == 0 || (src && dst && nbytes), "avdf_host_pack: null span arrays");
  std::vector<Piece> pieces;
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] == 0) continue;
    AVDF_CHECK_ARG(src[i] && dst[i], "avdf_host_pack: null span pointer");
    unsigned char* d = static_cast<unsigned char*>(dst[i]);
    const unsigned char* s = static_cast<const unsigned char*>(src[i]);
    for (size_t o = 0; o < nbytes[i]; o += PIECE)
      pieces.push_back({d + o, s + o, nbytes[i] - o < PIECE ? nbytes[i] - o : PIECE});
  }
  if (pieces.empty()) return AVDF_OK;
  pool().run(pieces, n_threads);
  return AVDF_OK;
}

// 1 when every span lies in page-locked (pinned / registered) host memory, else 0 (also without a usable driver):
// such a batch needs no staging copy at all.
extern "C" int avdf_host_all_pinned(const void* const* src, const size_t* nbytes, int32_t n) {
  if (n <= 0 || !src || !nbytes) return 0;
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] == 0) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src[i]) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type != cudaMemoryTypeHost) return 0;
    // the last byte too: a span must not run past its registration
    if (cudaPointerGetAttributes(&at, static_cast<const unsigned char*>(src[i]) + nbytes[i] - 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (at.type != cudaMemoryTypeHost) return 0;
  }
  return 1;
}

// n asynchronous host -> device copies on `stream` (src: pinned host memory, dst: device): the zero-staging path of the
// feature ingestion - the copy engine reads the caller's arrays directly. The caller keeps the sources alive and
// unchanged until the stream has passed the copies.
extern "C" int avdf_h2d_gather(const void* const* src, void* const* dst, const size_t* nbytes, int32_t n, void* stream) {
  using namespace avdf;
  AVDF_CHECK_ARG(n >= 0, "n >= 0");
  AVDF_CHECK_ARG(n == 0 || (src && dst && nbytes), "null span arrays");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Spans that follow each other in BOTH address spaces are sent as one copy: a batch collated stream-major in one pinned
  // block (libs/datasets collate_pinned) leaves as one copy per stream instead of one per video and stream - measured on
  // the PCIe Gen5 link of the bench box: 46.9 GB/s with 96 copies of ~0.8 MB per batch, 55.3 GB/s with one (scripts/h2d_ceiling.py)
  int i = 0;
  while (i < n) {
    if (nbytes[i] == 0) { ++i; continue; }
    AVDF_CHECK_ARG(src[i] && dst[i], "null span pointer");
    const char* s0 = static_cast<const char*>(src[i]);
    char* d0 = static_cast<char*>(dst[i]);
    size_t len = nbytes[i];
    int j = i + 1;
    while (j < n && (nbytes[j] == 0 || (static_cast<const char*>(src[j]) == s0 + len && static_cast<char*>(dst[j]) == d0 + len))) {
      len += nbytes[j];
      ++j;
    }
    AVDF_CUDA(cudaMemcpyAsync(d0, s0, len, cudaMemcpyHostToDevice, st));
    i = j;
  }
  return AVDF_OK;
}
