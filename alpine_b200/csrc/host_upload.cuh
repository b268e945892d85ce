// Host -> device upload of the expression matrix for ALPINE.fit on host buffers (the e2e path): the reference does
// `torch.tensor(X_array, device=...)` (main.py:445), i.e. a driver-staged copy from pageable memory (~11 GB/s).
// Here a pool of host threads copies row chunks of the caller's pageable array into a small ring of pinned staging
// buffers (non-temporal stores, one chunk per thread at a time) and queues one DMA per chunk on per-thread copy
// streams, so that staging and PCIe transfers overlap and neither Python nor the GIL is involved per chunk.
// The pinned ring belongs to the process (allocated on first use, reused by every later upload, never freed);
// slots are handed out under a mutex, so concurrent uploads from several host threads (one fit per GPU) share it.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace alpine {

constexpr size_t kUploadSlotBytes = 2u << 20;  // 2 MB chunks: >95 % PCIe efficiency, 1 ms to pin each
constexpr int kUploadSlotsPerThread = 2;
constexpr int kUploadMaxThreads = 32;

struct UploadSlot {
  float* buf = nullptr;
  cudaEvent_t ev = nullptr;  // completion of the last DMA that read this slot
  bool used = false;         // an event has been recorded
};

struct UploadPool {
  std::mutex mu;
  std::vector<UploadSlot*> free_slots;
  // take `n` slots (allocating pinned memory as needed); returns false on allocation failure
  bool acquire(int n, std::vector<UploadSlot*>* out) {
    std::lock_guard<std::mutex> lock(mu);
    while (static_cast<int>(out->size()) < n) {
      if (!free_slots.empty()) {
        out->push_back(free_slots.back());
        free_slots.pop_back();
        continue;
      }
      UploadSlot* s = new UploadSlot();
      if (cudaHostAlloc(reinterpret_cast<void**>(&s->buf), kUploadSlotBytes, cudaHostAllocPortable) != cudaSuccess ||
          cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (s->buf) cudaFreeHost(s->buf);
        delete s;
        return false;
      }
      out->push_back(s);
    }
    return true;
  }
  void release(const std::vector<UploadSlot*>& slots) {
    std::lock_guard<std::mutex> lock(mu);
    for (UploadSlot* s : slots) free_slots.push_back(s);
  }
};
// one pool per device (a slot's event belongs to the device that was current when it was created); leaked on
// purpose: no teardown-order issues with the CUDA runtime
constexpr int kUploadMaxDevices = 64;
inline UploadPool& upload_pool(int device) {
  static UploadPool* pools = new UploadPool[kUploadMaxDevices];
  return pools[device];
}

// dst (pinned, 32-byte aligned) <- src (pageable), n floats, with streaming stores: the staging buffer is read next
// by the DMA engine, not by this core, so it should not displace the source lines from the cache
#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void stream_copy_avx2(float* dst, const float* src, size_t n) {
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 16));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 24));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 8), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 16), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 24), d);
  }
  for (; i < n; ++i) dst[i] = src[i];
  _mm_sfence();
}
#endif
inline void stage_copy(float* dst, const float* src, size_t n) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) {
    stream_copy_avx2(dst, src, n);
    return;
  }
#endif
  memcpy(dst, src, n * sizeof(float));
}

// dst[r * ld_dst + c] (device) = src[r * ld_src + c] (pageable host), r < rows, c < cols.  On return every DMA has
// been queued and `stream` waits for all of them (stream-ordered completion; the host does not block on the GPU).
// `device` must be the calling thread's current device.  Returns cudaSuccess or the first error.
inline cudaError_t upload_rows_f32(int device, float* dst, int64_t ld_dst, const float* src, int64_t ld_src,
                                   int64_t rows, int64_t cols, int n_threads, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (device < 0 || device >= kUploadMaxDevices) return cudaErrorInvalidDevice;
  {
    // page-locked source (cudaHostAlloc / cudaHostRegister by the caller): the DMA engine reads it directly, no
    // staging and no host threads
    cudaPointerAttributes a0{}, a1{};
    const float* last = src + (rows - 1) * ld_src + (cols - 1);
    const bool pinned = cudaPointerGetAttributes(&a0, src) == cudaSuccess && a0.type == cudaMemoryTypeHost &&
                        cudaPointerGetAttributes(&a1, last) == cudaSuccess && a1.type == cudaMemoryTypeHost;
    cudaGetLastError();  // (older runtimes report unregistered memory as an error)
    if (pinned) {
      if (ld_src == cols && ld_dst == cols)
        return cudaMemcpyAsync(dst, src, static_cast<size_t>(rows) * cols * sizeof(float), cudaMemcpyHostToDevice, stream);
      return cudaMemcpy2DAsync(dst, ld_dst * sizeof(float), src, ld_src * sizeof(float), cols * sizeof(float), rows,
                               cudaMemcpyHostToDevice, stream);
    }
  }
  UploadPool& pool = upload_pool(device);
  if (n_threads < 1) n_threads = 1;
  if (n_threads > kUploadMaxThreads) n_threads = kUploadMaxThreads;
  const int64_t slot_floats = static_cast<int64_t>(kUploadSlotBytes / sizeof(float));
  // a chunk is a whole number of rows when a row fits a slot, otherwise a piece of one row
  const int64_t rows_per_chunk = cols <= slot_floats ? slot_floats / cols : 0;
  const int64_t pieces_per_row = rows_per_chunk > 0 ? 1 : (cols + slot_floats - 1) / slot_floats;
  const int64_t n_chunks = rows_per_chunk > 0 ? (rows + rows_per_chunk - 1) / rows_per_chunk : rows * pieces_per_row;
  if (n_chunks < n_threads) n_threads = static_cast<int>(n_chunks);
  std::vector<UploadSlot*> slots;
  if (!pool.acquire(n_threads * kUploadSlotsPerThread, &slots)) {
    pool.release(slots);
    return cudaErrorMemoryAllocation;
  }
  std::atomic<int64_t> next{0};
  std::atomic<int> first_error{static_cast<int>(cudaSuccess)};
  std::vector<cudaEvent_t> done(n_threads, nullptr);
  auto worker = [&](int t) {
    cudaError_t e = cudaSetDevice(device);
    cudaStream_t cs = nullptr;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    int turn = 0;
    while (e == cudaSuccess && first_error.load(std::memory_order_relaxed) == static_cast<int>(cudaSuccess)) {
      const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= n_chunks) break;
      UploadSlot* s = slots[t * kUploadSlotsPerThread + (turn++ % kUploadSlotsPerThread)];
      if (s->used) e = cudaEventSynchronize(s->ev);  // the DMA that last read this slot has finished
      if (e != cudaSuccess) break;
      if (rows_per_chunk > 0) {
        const int64_t r0 = i * rows_per_chunk;
        const int64_t nr = (r0 + rows_per_chunk <= rows) ? rows_per_chunk : rows - r0;
        if (ld_src == cols) {
          stage_copy(s->buf, src + r0 * ld_src, static_cast<size_t>(nr * cols));
        } else {
          for (int64_t r = 0; r < nr; ++r)  // (rows of the staging buffer are not 32-byte aligned: plain copies)
            memcpy(s->buf + r * cols, src + (r0 + r) * ld_src, static_cast<size_t>(cols) * sizeof(float));
        }
        if (ld_dst == cols)
          e = cudaMemcpyAsync(dst + r0 * ld_dst, s->buf, nr * cols * sizeof(float), cudaMemcpyHostToDevice, cs);
        else
          e = cudaMemcpy2DAsync(dst + r0 * ld_dst, ld_dst * sizeof(float), s->buf, cols * sizeof(float),
                                cols * sizeof(float), nr, cudaMemcpyHostToDevice, cs);
      } else {
        const int64_t r = i / pieces_per_row, c0 = (i % pieces_per_row) * slot_floats;
        const int64_t nc = (c0 + slot_floats <= cols) ? slot_floats : cols - c0;
        stage_copy(s->buf, src + r * ld_src + c0, static_cast<size_t>(nc));
        e = cudaMemcpyAsync(dst + r * ld_dst + c0, s->buf, nc * sizeof(float), cudaMemcpyHostToDevice, cs);
      }
      if (e == cudaSuccess) e = cudaEventRecord(s->ev, cs);
      s->used = true;
    }
    if (e == cudaSuccess && cs != nullptr) {
      // hand the completion of this thread's copy stream to the caller's stream
      e = cudaEventCreateWithFlags(&done[t], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventRecord(done[t], cs);
    }
    if (cs != nullptr) cudaStreamDestroy(cs);  // deferred by the runtime until its queued work has completed
    if (e != cudaSuccess) {
      int expected = static_cast<int>(cudaSuccess);
      first_error.compare_exchange_strong(expected, static_cast<int>(e));
    }
  };
  std::vector<std::thread> threads;
  threads.reserve(n_threads);
  for (int t = 1; t < n_threads; ++t) threads.emplace_back(worker, t);
  worker(0);
  for (auto& th : threads) th.join();
  cudaError_t result = static_cast<cudaError_t>(first_error.load());
  for (int t = 0; t < n_threads; ++t) {
    if (done[t] == nullptr) continue;
    if (result == cudaSuccess) result = cudaStreamWaitEvent(stream, done[t], 0);
    cudaEventDestroy(done[t]);  // released once the event has completed
  }
  pool.release(slots);  // their events guard the reuse by the next upload
  return result;
}

}  // namespace alpine
