// stage_pipeline.h -- pageable host memory -> page-locked ring -> wire, host side only (no CUDA in
// here: the wire is a policy, so tests/stage_pipeline_driver.cpp runs the same code against a
// software wire on any machine).
//
// A pageable block (what a Fortran ALLOCATE or an mmap of the WAVECAR gives the caller; the
// reference's reader fills such arrays, read_wavecar_mod.f90:537-583) left to cudaMemcpyAsync moves
// at 2-3 GB/s.  Here it is cut into chunks (one DMA each) that travel through a small ring of
// page-locked buffers, and every chunk into PIECES that a pool of host threads copies with
// streaming stores.  Pieces are handed out from one atomic counter over the whole block:
//   * no thread owns a fixed slice, so a thread the hypervisor parks for a millisecond delays one
//     256 KB piece, not a sixteenth of every chunk (the first version split every chunk evenly and
//     met at a barrier per chunk: 22-46 GB/s from run to run on the 16-vCPU boxes);
//   * threads run ahead across chunk boundaries as far as the ring has free buffers, so the copy of
//     chunk i+1 never waits for the stragglers of chunk i.
// The calling thread orchestrates -- puts a chunk on the wire as soon as its last piece has landed,
// releases the ring buffer of a finished DMA to the copiers -- and copies pieces itself whenever
// neither is due.  One job at a time (the ABI is single-caller per handle).
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

namespace bandup {

// The staging buffer is written once and next read by the DMA engine: streaming (non-temporal)
// stores skip the read-for-ownership of the destination lines, a third of the copy's DRAM traffic.
inline void stream_copy(char* dst, const char* src, size_t n) {
#if defined(__x86_64__) && defined(__SSE2__)
  size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
  if (head > n) head = n;
  std::memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  const size_t blocks = n / 64;
  for (size_t i = 0; i < blocks; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
    const __m128i c2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c2);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
    src += 64; dst += 64;
  }
  _mm_sfence();
  std::memcpy(dst, src, n - blocks * 64);
#else
  std::memcpy(dst, src, n);
#endif
}

inline void cpu_relax() {
#if defined(__x86_64__)
  _mm_pause();
#else
  std::this_thread::yield();
#endif
}

class StagePool {
 public:
  static constexpr int kMaxRing = 8;

  // n_threads copiers in total: n_threads - 1 pool threads + the calling thread
  explicit StagePool(int n_threads) : n_workers_(n_threads > 1 ? n_threads - 1 : 0) {
    for (int t = 0; t < n_workers_; ++t) threads_.emplace_back([this] { worker(); });
  }
  ~StagePool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& th : threads_) th.join();
  }
  StagePool(const StagePool&) = delete;
  StagePool& operator=(const StagePool&) = delete;

  int threads() const { return n_workers_ + 1; }

  // Moves `bytes` from src through ring[0..ring_size) (each at least `chunk` bytes), chunk i through
  // buffer (k0 + i) % ring_size.  Wire:
  //   bool ready(int k)                      the last send of ring buffer k has left it (never blocks)
  //   int  send(int k, size_t off, size_t n) put ring[k][0..n) on the wire for block offset off; 0 = ok
  // Chunks are sent in order.  Returns 0, or the first non-zero code of send() (the rest of the
  // block is then abandoned; the pool stays usable).
  // streaming: pieces are written with non-temporal stores (ring much larger than the last-level
  // cache) instead of memcpy (a ring small enough that the DMA engine finds the lines still in cache)
  template <class Wire>
  int run(const char* src, size_t bytes, char* const* ring, int ring_size, int k0, size_t chunk, size_t piece,
          Wire& wire, bool streaming = true) {
    if (bytes == 0) return 0;
    if (ring_size > kMaxRing) ring_size = kMaxRing;
    if (piece == 0 || piece > chunk) piece = chunk;
    const long n_chunks = (long)((bytes + chunk - 1) / chunk);
    const long ppc = (long)((chunk + piece - 1) / piece);  // pieces per (full) chunk
    {
      std::lock_guard<std::mutex> lk(m_);
      src_ = src;
      bytes_ = bytes;
      chunk_ = chunk;
      piece_ = piece;
      ppc_ = ppc;
      n_chunks_ = n_chunks;
      n_pieces_ = n_chunks * ppc;  // a short last chunk has empty trailing pieces: they count as done at once
      ring_size_ = ring_size;
      k0_ = k0;
      streaming_ = streaming;
      for (int k = 0; k < ring_size; ++k) ring_[k] = ring[k];
      if ((long)done_.size() < n_chunks) done_ = std::vector<std::atomic<int>>((size_t)n_chunks);
      for (long i = 0; i < n_chunks; ++i) done_[(size_t)i].store(0, std::memory_order_relaxed);
      next_.store(0, std::memory_order_relaxed);
      released_.store(0, std::memory_order_relaxed);
      abort_.store(false, std::memory_order_relaxed);
      active_.store(n_workers_, std::memory_order_release);
      ++job_;
    }
    cv_work_.notify_all();

    long rel = 0, iss = 0, pending = -1;
    bool exhausted = false;
    int rc = 0;
    while (iss < n_chunks && rc == 0) {
      bool progressed = false;
      // 1. every chunk whose pieces have all landed goes on the wire, in order
      while (iss < n_chunks && done_[(size_t)iss].load(std::memory_order_acquire) == (int)ppc) {
        const size_t off = (size_t)iss * chunk, n = bytes - off < chunk ? bytes - off : chunk;
        rc = wire.send((int)((k0 + iss) % ring_size), off, n);
        if (rc != 0) break;
        ++iss;
        progressed = true;
      }
      if (rc != 0) break;
      // 2. hand ring buffers whose DMA has finished to the copiers (chunk rel re-uses the buffer of
      //    chunk rel - ring_size, which must have been sent: rel < iss + ring_size)
      while (rel < n_chunks && rel < iss + ring_size && wire.ready((int)((k0 + rel) % ring_size))) {
        ++rel;
        {
          std::lock_guard<std::mutex> lk(m_rel_);
          released_.store(rel, std::memory_order_release);
        }
        cv_rel_.notify_all();
        progressed = true;
      }
      // 3. otherwise copy a piece ourselves (one that lies in a released chunk)
      if (pending < 0 && !exhausted) {
        const long g = next_.fetch_add(1, std::memory_order_relaxed);
        if (g < n_pieces_)
          pending = g;
        else
          exhausted = true;  // every piece has a taker
      }
      if (pending >= 0 && pending / ppc < rel) {
        copy_piece(pending);
        pending = -1;
        progressed = true;
      }
      if (!progressed) cpu_relax();
    }
    if (rc != 0) {
      abort_.store(true, std::memory_order_release);
      {
        std::lock_guard<std::mutex> lk(m_rel_);
      }
      cv_rel_.notify_all();
    }
    // the job's state lives in this object: nobody may still be reading it when the next job is set up
    while (active_.load(std::memory_order_acquire) != 0) cpu_relax();
    return rc;
  }

 private:
  void copy_piece(long g) {
    const long c = g / ppc_;
    const size_t in_chunk = (size_t)(g - c * ppc_) * piece_;
    const size_t off = (size_t)c * chunk_ + in_chunk;
    const size_t chunk_bytes = bytes_ - (size_t)c * chunk_ < chunk_ ? bytes_ - (size_t)c * chunk_ : chunk_;
    if (in_chunk < chunk_bytes) {
      const size_t n = chunk_bytes - in_chunk < piece_ ? chunk_bytes - in_chunk : piece_;
      char* to = ring_[(k0_ + c) % ring_size_] + in_chunk;
      if (streaming_)
        stream_copy(to, src_ + off, n);
      else
        std::memcpy(to, src_ + off, n);
    }
    done_[(size_t)c].fetch_add(1, std::memory_order_acq_rel);
  }

  void worker() {
    long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_work_.wait(lk, [&] { return stop_ || job_ != seen; });
        if (stop_) return;
        seen = job_;
      }
      for (;;) {
        if (abort_.load(std::memory_order_acquire)) break;
        const long g = next_.fetch_add(1, std::memory_order_relaxed);
        if (g >= n_pieces_) break;
        const long c = g / ppc_;
        if (released_.load(std::memory_order_acquire) <= c) {
          // the ring is full (the wire is the bottleneck): spin a little, then sleep
          int spins = 0;
          while (released_.load(std::memory_order_acquire) <= c && !abort_.load(std::memory_order_acquire) &&
                 ++spins < 2000)
            cpu_relax();
          if (released_.load(std::memory_order_acquire) <= c) {
            std::unique_lock<std::mutex> lk(m_rel_);
            cv_rel_.wait(lk, [&] {
              return abort_.load(std::memory_order_acquire) || released_.load(std::memory_order_acquire) > c;
            });
          }
        }
        if (abort_.load(std::memory_order_acquire)) break;
        copy_piece(g);
      }
      active_.fetch_sub(1, std::memory_order_acq_rel);
    }
  }

  const int n_workers_;
  std::vector<std::thread> threads_;
  std::mutex m_, m_rel_;
  std::condition_variable cv_work_, cv_rel_;
  bool stop_ = false;
  long job_ = 0;
  // the current job
  const char* src_ = nullptr;
  size_t bytes_ = 0, chunk_ = 0, piece_ = 0;
  long ppc_ = 1, n_chunks_ = 0, n_pieces_ = 0;
  char* ring_[kMaxRing] = {nullptr};
  int ring_size_ = 1, k0_ = 0;
  bool streaming_ = true;
  std::vector<std::atomic<int>> done_;
  std::atomic<long> next_{0}, released_{0};
  std::atomic<bool> abort_{false};
  std::atomic<int> active_{0};
};

}  // namespace bandup
