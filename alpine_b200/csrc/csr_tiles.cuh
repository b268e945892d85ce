// Sparse X: one-time conversion of a CSR matrix over cells (row j = cell, column ids = genes; the layout AnnData
// keeps for count matrices) into the two "tile list" copies the contraction kernel streams (mu_gemm_sm100.cuh,
// SRC_TILES): the nonzeros are grouped by (256-row super-tile, 32-deep k-block) of each contraction's work space,
//
//   ORIENT_XH  (X H^T,  main.py:596):  rows = genes, reduction = cells   -> block (g / 256, j / 32)
//   ORIENT_WX  (W^T X,  main.py:653):  rows = cells, reduction = genes   -> block (j / 256, g / 32)
//
// and every entry carries the byte offset of its element inside the shared-memory tile (row-major, 128-byte rows,
// TMA-style swizzle; the same form for both orientations) plus the fp32 value, so the producer warp only scatters.  8 bytes per nonzero and orientation,
// read front to back with coalesced 256-byte warp loads; the order of the entries inside a block is irrelevant
// (distinct positions), so the atomically assigned slots do not make results non-deterministic.  Column ids must
// be unique within a row (canonical CSR: scipy's sum_duplicates()).
#pragma once
#include "mu_gemm_sm100.cuh"

namespace alpine {

struct CsrView {
  const long long* indptr;  // [n_rows + 1]
  const int* indices;       // [nnz] gene ids
  const float* values;      // [nnz]
  long long n_rows;         // cells of this shard
  int n_cols;               // genes
};

// err[0] = 1: column id out of range; err[1] = 1: negative or non-finite value
__global__ void __launch_bounds__(256) csr_tile_count_kernel(const CsrView m, int kb_xh, int kb_wx,
                                                             unsigned int* __restrict__ cnt_xh,
                                                             unsigned int* __restrict__ cnt_wx, int* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long j = warp0; j < m.n_rows; j += n_warps) {
    const long long e0 = m.indptr[j], e1 = m.indptr[j + 1];
    for (long long e = e0 + lane; e < e1; e += 32) {
      const int g = m.indices[e];
      const float v = m.values[e];
      if (g < 0 || g >= m.n_cols) {
        err[0] = 1;
        continue;
      }
      if (!(v >= 0.f) || v > 3.0e38f) err[1] = 1;
      atomicAdd(cnt_xh + static_cast<long long>(g >> 8) * kb_xh + (j >> 5), 1u);
      atomicAdd(cnt_wx + (j >> 8) * kb_wx + (g >> 5), 1u);
    }
  }
}

// exclusive prefix sum of `n` counts into 64-bit offsets (ofs[n] = total); one block of 1024 threads.  The counts
// array is zeroed on the way out so that the fill kernel can reuse it as its per-block cursor.
__global__ void __launch_bounds__(1024) csr_tile_scan_kernel(unsigned int* __restrict__ cnt, long long n,
                                                             long long* __restrict__ ofs) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const long long chunk = (n + 1023) / 1024;
  const long long b = t * chunk, e = (b + chunk < n) ? b + chunk : n;
  long long s = 0;
  for (long long i = b; i < e; ++i) s += cnt[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    const long long v = (t >= o) ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = part[t] - s;
  for (long long i = b; i < e; ++i) {
    const unsigned int c = cnt[i];
    ofs[i] = run;
    run += c;
    cnt[i] = 0u;
  }
  if (t == 1023) ofs[n] = part[1023];
}

__global__ void __launch_bounds__(256) csr_tile_fill_kernel(const CsrView m, int kb_xh, int kb_wx,
                                                            const long long* __restrict__ ofs_xh,
                                                            const long long* __restrict__ ofs_wx,
                                                            unsigned int* __restrict__ cur_xh,
                                                            unsigned int* __restrict__ cur_wx,
                                                            uint2* __restrict__ ent_xh, uint2* __restrict__ ent_wx) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (gridDim.x * static_cast<long long>(blockDim.x)) >> 5;
  for (long long j = warp0; j < m.n_rows; j += n_warps) {
    const long long e0 = m.indptr[j], e1 = m.indptr[j + 1];
    for (long long e = e0 + lane; e < e1; e += 32) {
      const int g = m.indices[e];
      if (g < 0 || g >= m.n_cols) continue;
      const unsigned int bits = __float_as_uint(m.values[e]);
      {
        const long long blk = static_cast<long long>(g >> 8) * kb_xh + (j >> 5);
        const unsigned int slot = atomicAdd(cur_xh + blk, 1u);
        ent_xh[ofs_xh[blk] + slot] = make_uint2(x_tile_offset(ORIENT_WX, g & 255, static_cast<int>(j & 31)), bits);
      }
      {
        const long long blk = (j >> 8) * kb_wx + (g >> 5);
        const unsigned int slot = atomicAdd(cur_wx + blk, 1u);
        ent_wx[ofs_wx[blk] + slot] = make_uint2(x_tile_offset(ORIENT_WX, static_cast<int>(j & 255), g & 31), bits);
      }
    }
  }
}

// sum of squares (fp64) of a flat value array + the "not tf32-exact" flag of the count-matrix fast path
__global__ void __launch_bounds__(256) sumsq_flat_kernel(const float* __restrict__ v, long long n,
                                                         double* __restrict__ partial, int* __restrict__ inexact) {
  double acc = 0.0;
  unsigned int low = 0u;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = __ldg(v + i);
    acc += static_cast<double>(x) * static_cast<double>(x);
    low |= __float_as_uint(x) & 0x1FFFu;
  }
  __shared__ double red[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  if (__any_sync(0xffffffffu, low != 0u) && (threadIdx.x & 31) == 0) *inexact = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

}  // namespace alpine
