// Exercises bandup_b200/csrc/stage_pipeline.h (the pageable -> page-locked ring -> wire pipeline of
// bandup_gpu_submit_kpt) against a SOFTWARE wire: a thread that plays the DMA engine, copying ring
// buffers to the destination one after the other, optionally slower than the copiers.  Checks that
// every byte arrives, in order, for ragged sizes / misaligned sources / tiny rings, back-to-back
// jobs sharing one ring, and that a failing send aborts the block without hanging the pool.
//   g++ -O2 -std=c++17 -pthread [-fsanitize=thread] tests/stage_pipeline_driver.cpp -o driver && ./driver
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <random>

#include "../bandup_b200/csrc/stage_pipeline.h"

using bandup::StagePool;

struct SoftWire {
  struct Req { int k; size_t off, n; };
  char* dst = nullptr;
  char** ring = nullptr;
  int ring_size = 0;
  int fail_at = -1;          // send number that reports an error
  int delay_us = 0;          // per request: the wire is the bottleneck when > 0
  std::mutex m;
  std::condition_variable cv;
  std::deque<Req> q;
  std::vector<int> in_flight;  // per ring buffer
  bool stop = false;
  int sends = 0;
  size_t last_end = 0;
  bool in_order = true;
  std::thread th;

  SoftWire(char** r, int rs) : ring(r), ring_size(rs), in_flight((size_t)rs, 0) {
    th = std::thread([this] { engine(); });
  }
  ~SoftWire() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv.notify_all();
    th.join();
  }
  void engine() {
    for (;;) {
      Req r;
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return stop || !q.empty(); });
        if (q.empty()) return;
        r = q.front();
        q.pop_front();
      }
      if (delay_us) std::this_thread::sleep_for(std::chrono::microseconds(delay_us));
      std::memcpy(dst + r.off, ring[r.k], r.n);
      {
        std::lock_guard<std::mutex> lk(m);
        --in_flight[(size_t)r.k];
      }
    }
  }
  void drain() {
    for (;;) {
      {
        std::lock_guard<std::mutex> lk(m);
        bool idle = q.empty();
        for (int v : in_flight) idle = idle && v == 0;
        if (idle) return;
      }
      std::this_thread::yield();
    }
  }
  bool ready(int k) {
    std::lock_guard<std::mutex> lk(m);
    return in_flight[(size_t)k] == 0;
  }
  int send(int k, size_t off, size_t n) {
    std::lock_guard<std::mutex> lk(m);
    if (sends++ == fail_at) return 7;
    if (in_flight[(size_t)k] != 0) in_order = false;  // a buffer handed out while still on the wire
    if (off != last_end) in_order = false;
    last_end = off + n;
    ++in_flight[(size_t)k];
    q.push_back({k, off, n});
    cv.notify_all();
    return 0;
  }
};

static int failures = 0;
#define CHECK(x) do { if (!(x)) { std::printf("FAILED %s (line %d)\n", #x, __LINE__); ++failures; } } while (0)

int main(int argc, char** argv) {
  const int max_threads = argc > 1 ? std::atoi(argv[1]) : 6;
  std::mt19937_64 rng(20261020);
  const size_t cap = (size_t(9) << 20) + 4096;
  std::vector<char> src_store(cap + 64), dst(cap), ringmem;
  for (auto& b : src_store) b = (char)rng();
  int jobs = 0;
  for (int threads = 1; threads <= max_threads; ++threads) {
    for (int ring_size : {1, 2, 4}) {
      const size_t ring_chunk = size_t(1) << 20;
      ringmem.assign(ring_chunk * (size_t)ring_size + 64, 0);
      std::vector<char*> ring((size_t)ring_size);
      for (int k = 0; k < ring_size; ++k) ring[(size_t)k] = ringmem.data() + 32 + ring_chunk * (size_t)k;
      StagePool pool(threads);
      CHECK(pool.threads() == threads);
      SoftWire wire(ring.data(), ring_size);
      int k0 = 0;
      // sizes: empty tail, ragged tail, smaller than a piece, one byte, many chunks; sources off by 0/8/3 bytes
      const size_t sizes[] = {1, 100, 4096, 65536 + 17, ring_chunk, ring_chunk + 1, 3 * ring_chunk - 5, cap - 11};
      for (size_t bytes : sizes)
        for (size_t chunk : {ring_chunk, ring_chunk / 4, size_t(4096)})
          for (size_t piece : {size_t(0), size_t(65536), size_t(1000)}) {
            if (bytes / chunk > 4000) continue;  // keep the run short
            const size_t misalign = (size_t)(jobs % 3 == 0 ? 0 : jobs % 3 == 1 ? 8 : 3);
            const char* src = src_store.data() + misalign;
            std::fill(dst.begin(), dst.begin() + (long)bytes, 0);
            wire.dst = dst.data();
            wire.last_end = 0;
            wire.delay_us = (jobs % 5 == 0) ? 30 : 0;
            const int rc = pool.run(src, bytes, ring.data(), ring_size, k0, chunk, piece, wire);
            k0 = (int)((k0 + (long)((bytes + chunk - 1) / chunk)) % ring_size);
            wire.drain();
            CHECK(rc == 0);
            CHECK(wire.in_order && wire.last_end == bytes);
            CHECK(std::memcmp(dst.data(), src, bytes) == 0);
            ++jobs;
          }
      // a send that fails: the code comes back, nothing hangs, the pool takes the next block
      wire.fail_at = wire.sends + 2;
      wire.last_end = 0;
      const int rc = pool.run(src_store.data(), 5 * ring_chunk, ring.data(), ring_size, k0, ring_chunk / 2, 65536, wire);
      CHECK(rc == 7);
      wire.drain();
      wire.fail_at = -1;
      wire.last_end = 0;
      wire.in_order = true;
      std::fill(dst.begin(), dst.end(), 0);
      CHECK(pool.run(src_store.data(), cap, ring.data(), ring_size, 0, ring_chunk, 65536, wire) == 0);
      wire.drain();
      CHECK(wire.in_order && std::memcmp(dst.data(), src_store.data(), cap) == 0);
      ++jobs;
    }
  }
  std::printf("%d jobs, %d failures\n", jobs, failures);
  return failures ? 1 : 0;
}
