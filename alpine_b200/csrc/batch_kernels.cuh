// Mini-batch epochs (main.py:509-521, 589-595, 662): the cells of one batch are gathered out of the full-data arrays into
// the batch context's contiguous buffers, one MU step runs on them, and the updated H columns go back.  One launch
// each way: rows of the cells-major X (contiguous, float4), columns of H and of every Y; the tail of the buffers
// (cells [cnt, n) of a context that is larger than the batch) is zero-filled -- an all-zero cell contributes nothing to
// any sum of the step and stays zero.
#pragma once
#include <cstdint>

#include "mu_small_kernels.cuh"

namespace alpine {

enum { ERR_BATCH_INDEX = 30 };

struct BatchGatherParams {
  const long long* idx;  // [cnt] cell numbers in the full-data arrays
  long long cnt, n, n_all;
  // dense X (null for a CSR context: the caller binds the batch's own CSR rows)
  const float* X_all;
  long long ldX_all;
  float* X;
  long long ldX, G;
  int vec4;  // both X arrays 16-byte aligned with pitches that are multiples of 4
  const float* H_all;
  long long ldH_all;
  float* H;
  long long ldH;
  int K;
  int n_cov;
  const float* Y_all[kMaxCov];  // [c][n_all]
  float* Y[kMaxCov];            // [c][n]
  int c[kMaxCov];
  int x_blocks;
  int* err;
};

__device__ __forceinline__ long long batch_index(const BatchGatherParams& p, long long j) {
  const long long i = p.idx[j];
  if (i < 0 || i >= p.n_all) {
    if (atomicCAS(p.err, 0, ERR_BATCH_INDEX) == 0) p.err[1] = static_cast<int>(j), p.err[2] = static_cast<int>(i);
    return -1;
  }
  return i;
}

// Blocks [0, x_blocks): rows of X, one row per warp at a time (512 B per warp instruction).  The remaining blocks:
// the (row, cell) elements of H and the Ys, cell fastest (coalesced stores; the loads are one 32-byte sector each).
__global__ void __launch_bounds__(256) batch_gather_kernel(const BatchGatherParams p) {
  if (static_cast<int>(blockIdx.x) < p.x_blocks) {
    const int lane = threadIdx.x & 31;
    const long long warp = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    const long long nwarps = static_cast<long long>(p.x_blocks) * 8;
    for (long long r = warp; r < p.n; r += nwarps) {
      const long long i = r < p.cnt ? batch_index(p, r) : -1;
      float* dst = p.X + r * p.ldX;
      if (i < 0) {
        if (p.vec4) {
          for (long long g = 4ll * lane; g < p.G; g += 128) *reinterpret_cast<float4*>(dst + g) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          for (long long g = lane; g < p.G; g += 32) dst[g] = 0.f;
        }
        continue;
      }
      const float* src = p.X_all + i * p.ldX_all;
      if (p.vec4) {
        // whole float4 groups: the pitch of both arrays is a multiple of 4, the columns [G, pitch) are padding
        long long g = 4ll * lane;
        for (; g + 384 < p.G; g += 512) {  // four independent 16-byte loads in flight per lane
          const float4 a = __ldcs(reinterpret_cast<const float4*>(src + g));
          const float4 b = __ldcs(reinterpret_cast<const float4*>(src + g + 128));
          const float4 c = __ldcs(reinterpret_cast<const float4*>(src + g + 256));
          const float4 d = __ldcs(reinterpret_cast<const float4*>(src + g + 384));
          *reinterpret_cast<float4*>(dst + g) = a;
          *reinterpret_cast<float4*>(dst + g + 128) = b;
          *reinterpret_cast<float4*>(dst + g + 256) = c;
          *reinterpret_cast<float4*>(dst + g + 384) = d;
        }
        for (; g < p.G; g += 128) *reinterpret_cast<float4*>(dst + g) = __ldcs(reinterpret_cast<const float4*>(src + g));
      } else {
        for (long long g = lane; g < p.G; g += 32) dst[g] = src[g];
      }
    }
    return;
  }
  int rows = p.K;
  for (int v = 0; v < p.n_cov; ++v) rows += p.c[v];
  const long long total = static_cast<long long>(rows) * p.n;
  const long long nth = static_cast<long long>(gridDim.x - p.x_blocks) * blockDim.x;
  for (long long e = static_cast<long long>(blockIdx.x - p.x_blocks) * blockDim.x + threadIdx.x; e < total; e += nth) {
    int row = static_cast<int>(e / p.n);
    const long long j = e - static_cast<long long>(row) * p.n;
    const long long i = j < p.cnt ? batch_index(p, j) : -1;
    if (row < p.K) {
      p.H[row * p.ldH + j] = i < 0 ? 0.f : p.H_all[row * p.ldH_all + i];
      continue;
    }
    row -= p.K;
    for (int v = 0; v < p.n_cov; ++v) {
      if (row < p.c[v]) {
        p.Y[v][row * p.n + j] = i < 0 ? 0.f : p.Y_all[v][row * p.n_all + i];
        break;
      }
      row -= p.c[v];
    }
  }
}

// H_all[:, idx[j]] = H[:, j] for j < cnt (main.py:662).  Duplicate cell numbers (sampling with replacement) carry
// identical columns -- the update of a column depends on that cell's data alone -- so the order of the stores is immaterial.
__global__ void __launch_bounds__(256) batch_scatter_kernel(const float* __restrict__ H, long long ldH, int K,
                                                            const long long* __restrict__ idx, long long cnt,
                                                            long long n_all, float* __restrict__ H_all,
                                                            long long ldH_all) {
  const long long total = static_cast<long long>(K) * cnt;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = e / cnt, j = e - row * cnt;
    const long long i = idx[j];
    if (i >= 0 && i < n_all) H_all[row * ldH_all + i] = H[row * ldH + j];
  }
}

}  // namespace alpine
