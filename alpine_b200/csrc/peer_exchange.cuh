// Cell sharding without a library collective on the critical path: the W update of main.py:592-612 fused with its
// exchange over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC).
//
// Every rank owns an "exchange block" [ reduce buffer (its partial X H^T, H H^T, rowsum H, B statistics) | W^T |
// flags ].  One iteration on rank r, all in stream order:
//   1. the contraction + reduce kernels leave the partial X_r H_r^T in the block              (alpine_mu_partials)
//   2. peer_gather_reduce_kernel: system-scope fence, ready[r] = epoch in every peer's block; wait until
//      ready[q] == epoch for all q; then sum over the ranks (rank order), with loads straight from peer memory,
//      the small statistics and THIS RANK'S GENE SLICE of the numerator ("reduce-scatter")
//   3. the Z_W plan and w_update_kernel on the gene slice: the updated columns of W^T are stored locally and into
//      every peer's W^T ("all-gather" by peer stores; with step 2: 2 x 7/8 x 8 MB per GPU over NVLink)
//   4. peer_close_and_split_kernel: done[r] = epoch everywhere, wait for done[q] == epoch for all q, then the tf32
//      split of the gathered W^T
// after which W^T is complete on every rank, bit-identical (same summation order everywhere).  A rank overwrites its
// partials only after step 4 of the same iteration, i.e. after every peer has finished reading them.
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

__device__ __forceinline__ bool peer_wait_all(const PeerTable& t, int which, int epoch, int* err) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if (threadIdx.x < t.world) {
    volatile int* f = t.flags[t.rank] + which * kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    while (*f < epoch) {
      if (clock64() - t0 > (20ll << 30)) {  // ~10 s: a peer that late is lost, not slow
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

// The close of the exchange, fused with the first consumer of the gathered W^T: block 0 publishes done[r] = epoch
// to every rank (the W-update kernel, whose stores into the peers' W^T precede this one in the stream, is complete),
// every block waits until all ranks have said so, and then splits W^T [K][cols] into its tf32 / bf16 operand copies (the
// B operand of W^T W and W^T X).
__global__ void __launch_bounds__(256) peer_close_and_split_kernel(const PeerTable t, int which, int epoch, int* err,
                                                                   const float* __restrict__ src, long long ld_src,
                                                                   int rows, long long cols, float* __restrict__ hi,
                                                                   float* __restrict__ lo, long long ld_dst) {
  ptx::pdl_enter();
  if (blockIdx.x == 0) {
    __threadfence_system();
    if (threadIdx.x < t.world) {
      volatile int* f = t.flags[threadIdx.x] + which * kMaxPeers + t.rank;
      *f = epoch;
    }
    __threadfence_system();
  }
  peer_wait_all(t, which, epoch, err);  // every block polls this rank's own flag array (local memory)
  const long long c4n = (cols + 3) >> 2;
  const long long total = static_cast<long long>(rows) * c4n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c4n, c = (i - r * c4n) << 2;
    const float4 v = ld_sys_v4(src + r * ld_src + c);  // written by the peers a moment ago: not from a stale L1 line
    ptx::store_split4(v, hi, lo, r, c, ld_dst);
  }
}

// "Reduce-scatter" by peer loads: block 0 first publishes ready[r] = epoch (this rank's partials are complete: the
// kernels that wrote them precede this one in the stream); every block then waits for all ranks and sums, in rank
// order,  (a) the small statistics [S | hsum | Q] of all ranks -> small_out (and the operand copies of the summed
// S)  and  (b) the columns [g0, g1) of the
// partial numerators (K x ldG at the head of every exchange block) -> p_out (same pitch).  One thread per float4,
// all peers' loads in flight together, so the NVLink latency is paid once, not per peer.
__global__ void __launch_bounds__(256) peer_gather_reduce_kernel(const PeerTable t, int epoch, int n_small,
                                                                 float* __restrict__ small_out, int K, long long ldG,
                                                                 long long g0, long long g1,
                                                                 float* __restrict__ p_out, int* err,
                                                                 float* __restrict__ s_hi, float* __restrict__ s_lo,
                                                                 int ld_split) {
  ptx::pdl_enter();
  if (blockIdx.x == 0) {
    __threadfence_system();
    if (threadIdx.x < t.world) {
      volatile int* f = t.flags[threadIdx.x] + 0 * kMaxPeers + t.rank;
      *f = epoch;
    }
    __threadfence_system();
  }
  peer_wait_all(t, 0, epoch, err);  // every block polls this rank's own flag array (local memory)
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long e = tid; e < n_small; e += nth) {
    float v[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) v[q] = (q < t.world) ? ld_sys_f32(t.small[q] + e) : 0.f;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) acc += v[q];
    small_out[e] = acc;
    if (s_hi != nullptr && e < static_cast<long long>(K) * K) {
      // the first K * K entries are the summed H H^T: its split copies are the B operand of (H H^T) W^T
      const int r = static_cast<int>(e / K), c = static_cast<int>(e - static_cast<long long>(r) * K);
      ptx::store_split1(acc, s_hi, s_lo, r, c, ld_split);
    }
  }
  const long long w4 = (g1 - g0 + 3) >> 2;  // g0 is a multiple of 64 and ldG of 4: whole float4 groups stay in the pitch
  for (long long e = tid; e < static_cast<long long>(K) * w4; e += nth) {
    const long long k = e / w4, c = g0 + ((e - k * w4) << 2);
    const long long o = k * ldG + c;
    float4 v[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      v[q] = (q < t.world) ? ld_sys_v4(t.small[q] - static_cast<long long>(K) * ldG + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) acc.x += v[q].x, acc.y += v[q].y, acc.z += v[q].z, acc.w += v[q].w;
    *reinterpret_cast<float4*>(p_out + o) = acc;
  }
}

}  // namespace alpine
