// Cell sharding without a library collective on the critical path: the W update of main.py:592-612 fused with its
// exchange over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC).
//
// Every rank owns an "exchange block" [ reduce buffer (its partial X H^T, H H^T, rowsum H, B statistics) | W^T |
// flags ].  One iteration on rank r, all in stream order:
//   1. the contraction + reduce kernels leave the partial X_r H_r^T in the block              (alpine_mu_partials)
//   2. peer_signal_kernel: system-scope fence, then ready[r] = epoch in every peer's block
//   3. peer_wait_small_kernel: wait until ready[q] == epoch for all q, then sum the small statistics of all ranks
//      (rank order) into a local buffer
//   4. sym_long_kernel<EPI_W> on THIS RANK'S GENE SLICE only: numerator = sum over ranks of their partials, read
//      straight from peer memory; the updated columns of W^T are stored locally and into every peer's W^T
//      ("two-shot" all-reduce with the update between reduce-scatter and all-gather: 2 x 7/8 x 8 MB per GPU)
//   5. peer_signal_kernel: done[r] = epoch;  6. peer_wait_kernel: wait for done[q] == epoch for all q
// after which W^T is complete on every rank, bit-identical (same summation order everywhere).  A rank overwrites its
// partials only after step 6 of the same iteration, i.e. after every peer has finished reading them.
// All waits are bounded and report through the context's error flag instead of hanging the GPU.
#pragma once
#include "mu_small_kernels.cuh"

namespace alpine {

constexpr int kPeerFlagInts = 64;                   // [0, 8): ready[q], [8, 16): done[q]
enum { ERR_PEER_TIMEOUT = 20 };

struct PeerTable {
  int world, rank;
  int* flags[kMaxPeers];          // every rank's flag array (this rank's own included)
  const float* small[kMaxPeers];  // every rank's [S | hsum | Q] partials
};

__global__ void peer_signal_kernel(const PeerTable t, int which, int epoch) {
  __threadfence_system();
  if (threadIdx.x < t.world) {
    volatile int* f = t.flags[threadIdx.x] + which * kMaxPeers + t.rank;
    *f = epoch;
  }
  __threadfence_system();
}

__device__ __forceinline__ bool peer_wait_all(const PeerTable& t, int which, int epoch, int* err) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (threadIdx.x < t.world) {
    volatile int* f = t.flags[t.rank] + which * kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    while (*f < epoch) {
      if (clock64() - t0 > (4ll << 30)) {  // ~2 s
        if (atomicCAS(err, 0, ERR_PEER_TIMEOUT) == 0) {
          err[1] = t.rank, err[2] = threadIdx.x, err[3] = which, err[4] = epoch;
        }
        bad = 1;
        break;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  return bad == 0;
}

__global__ void __launch_bounds__(256) peer_wait_kernel(const PeerTable t, int which, int epoch, int* err) {
  peer_wait_all(t, which, epoch, err);
}

// wait for every rank's partials, then out[e] = sum over ranks (in rank order) of their small statistics
__global__ void __launch_bounds__(256) peer_wait_small_kernel(const PeerTable t, int epoch, int n_small,
                                                              float* __restrict__ out, int* err) {
  peer_wait_all(t, 0, epoch, err);  // every block polls this rank's own flag array (local memory)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_small; e += gridDim.x * blockDim.x) {
    float v[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) v[q] = (q < t.world) ? ld_sys_f32(t.small[q] + e) : 0.f;  // loads in flight together
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) acc += v[q];
    out[e] = acc;
  }
}

}  // namespace alpine
