// The multiplicative updates of one MU iteration as two fused, persistent kernels (+ two small finish kernels), so
// that an iteration is six launches: X H^T contraction, W update, W finish, W^T X contraction, H update, H finish.
//
//   w_update_kernel  (main.py:596-612)  per 64-gene tile of W^T [K][G]:
//       numerator P = X H^T  (summed from the contraction's stream-K partial slots, or read from the reduce buffer)
//       Z = (H H^T) W^T      (K x K by K x 64, 3xTF32 mma.sync, fp32 accumulate)
//       W *= 2P / max(2Z + (1-l1) a W + orth (rowsum_K(W) - W) + l1 a, eps)
//       tf32 hi/lo copies of the new tile (B operand of the next contraction), peer stores (cell sharding),
//       and this CTA's running Gram partial  W_new^T W_new  (upper 16x16 blocks, 3xTF32 mma.sync)
//   w_finish_kernel  sums the Gram partials in a fixed order -> T = W^T W, and applies the B updates (main.py:615-628)
//   h_update_kernel  (main.py:631-663)  per 64-cell tile of H [K][n]:
//       numerator A = W^T X  (from the partial slots),  Z = T H,  guided terms from (old H, new B) per cell,
//       H *= (numG + 2A) / max(denG + 2Z, eps),  hi/lo copies,  t1 = sum A .* H_new (fp64),
//       statistics of (new H, new B) for the next iteration: rowsum(H), Q_i, prediction loss, Gram partial H H^T
//   h_finish_kernel  sums the partials -> S = H H^T, rowsum(H), Q_i into the reduce buffer, and the loss row
//       [t1, t2 = sum T .* S, pred_i] (main.py:666, 726-753 through the trace identity)
//
// The K x K products are < 1 % of the iteration's FLOPs; they run on the tensor cores through mma.sync.m16n8k8
// (tf32 operands split hi/lo, three products per term as in the big contractions) because the CUDA-core version of
// the same products was what bounded the update kernels (125 us of a 4.06 ms iteration at cfg 3, round 1).
// Everything is deterministic: fixed tile -> CTA assignment, fixed summation orders, no atomics on data.
#pragma once
#include "mu_small_kernels.cuh"

namespace alpine {

constexpr int kUpdCols = 64;       // columns (cells or genes) per tile
constexpr int kUpdPitch = 72;      // shared-memory row pitch of a tile (floats): conflict-free fragment loads
constexpr int kUpdThreads = 256;   // 8 warps
constexpr int kUpdWarps = 8;

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void mma_tf32_m16n8k8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32: small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bhi0,
                                     uint32_t bhi1, uint32_t blo0, uint32_t blo1) {
  mma_tf32_m16n8k8(d, alo, bhi0, bhi1);
  mma_tf32_m16n8k8(d, ahi, blo0, blo1);
  mma_tf32_m16n8k8(d, ahi, bhi0, bhi1);
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Where the numerator of a tile comes from: a finished [K][ld] array, or the stream-K partial-sum slots of the
// contraction that has just run (slot s holds [K][256] sums of one segment of a 256-row super-tile).
struct UpdNumSrc {
  const float* direct;
  long long ld;
  const float* partial;
  const int* slot_ofs;  // [super-tiles + 1]
  const int* slots;
  int K;                // row pitch of a slot is K * 256
};
// the two adjacent columns (col, col + 1) of row k; col is even
__device__ __forceinline__ float2 num_load2(const UpdNumSrc& s, int k, long long col, int s0, int s1) {
  if (s.direct != nullptr) return __ldg(reinterpret_cast<const float2*>(s.direct + static_cast<long long>(k) * s.ld + col));
  const int r = static_cast<int>(col & 255);
  float2 acc = make_float2(0.f, 0.f);
  for (int q = s0; q < s1; ++q) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(
        s.partial + (static_cast<size_t>(__ldg(s.slots + q)) * s.K + k) * 256 + r));
    acc.x += v.x, acc.y += v.y;
  }
  return acc;
}

// Upper-triangular 16x16 blocks (mt <= nt) of a (16 NC)^2 Gram matrix, enumerated row by row.
__device__ __forceinline__ void gram_block_decode(int b, int NC, int& mt, int& nt) {
  mt = 0;
  int rem = b;
  while (rem >= NC - mt) {
    rem -= NC - mt;
    ++mt;
  }
  nt = mt + rem;
}

template <int NC>
struct UpdGeom {
  static constexpr int Kp = 16 * NC;
  static constexpr int KS = 2 * NC;                  // k-steps of 8 over the padded K
  static constexpr int NB = NC * (NC + 1) / 2;       // Gram blocks
  static constexpr int NBW = (NB + kUpdWarps - 1) / kUpdWarps;  // Gram blocks per warp
  static constexpr int afrag_floats = NC * KS * 32 * 4;  // fp32 A fragments (split into hi / lo when loaded)
  static constexpr int tile_floats = Kp * kUpdPitch;
  static constexpr int n_tile_bufs = 4;                  // two cp.async targets + tf32 hi / lo of the current tile
  static constexpr int gram_floats = NBW * kUpdWarps * 2 * 32 * 4;  // per CTA, fragment order
};

// Stage Sym [K][K] (global) through `scratch` (>= K*K floats of shared memory) and write it in A-fragment order:
// afrag[((mt * KS + ks) * 32 + lane) * 4 + {a0..a3}],  a0 = (16 mt + g, 8 ks + t), a1 = (+8, .), a2 = (., +4),
// a3 = (+8, +4), g = lane / 4, t = lane % 4; entries outside K are zero.
template <int NC>
__device__ void build_afrag(float* afrag, float* scratch, const float* __restrict__ Sym, int ldS, int K) {
  constexpr int KS = 2 * NC;
  for (int e = threadIdx.x; e < K * K; e += kUpdThreads) {
    const int r = e / K, c = e - r * K;
    scratch[e] = __ldcg(Sym + static_cast<long long>(r) * ldS + c);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NC * KS * 32; e += kUpdThreads) {
    const int lane = e & 31, ks = (e >> 5) % KS, mt = (e >> 5) / KS;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = 16 * mt + g, r1 = r0 + 8, k0 = 8 * ks + t, k1 = k0 + 4;
    const float v0 = (r0 < K && k0 < K) ? scratch[r0 * K + k0] : 0.f;
    const float v1 = (r1 < K && k0 < K) ? scratch[r1 * K + k0] : 0.f;
    const float v2 = (r0 < K && k1 < K) ? scratch[r0 * K + k1] : 0.f;
    const float v3 = (r1 < K && k1 < K) ? scratch[r1 * K + k1] : 0.f;
    reinterpret_cast<float4*>(afrag)[e] = make_float4(v0, v1, v2, v3);
  }
  __syncthreads();
}

// tf32 hi / lo copies of the rows [0, K) of a tile (same pitch)
__device__ __forceinline__ void split_tile(const float* tile, float* thi, float* tlo, int K) {
  for (int e = threadIdx.x; e < K * 16; e += kUpdThreads) {
    const int o = (e >> 4) * kUpdPitch + 4 * (e & 15);
    const float4 v = *reinterpret_cast<const float4*>(tile + o);
    uint32_t h[4], l[4];
    ptx::split_tf32(v.x, h[0], l[0]);
    ptx::split_tf32(v.y, h[1], l[1]);
    ptx::split_tf32(v.z, h[2], l[2]);
    ptx::split_tf32(v.w, h[3], l[3]);
    *reinterpret_cast<float4*>(thi + o) =
        make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
    *reinterpret_cast<float4*>(tlo + o) =
        make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3]));
  }
}

// rows [0, K) x 64 columns [c0, c0 + 64) of Mat [K][ld] -> tile [Kp][72]; columns >= L arrive as zeros
__device__ __forceinline__ void load_tile_async(float* tile, const float* __restrict__ Mat, long long ld, int K,
                                                long long c0, long long L) {
  const uint32_t base = ptx::smem_u32(tile);
  for (int e = threadIdx.x; e < K * 16; e += kUpdThreads) {
    const int k = e >> 4, c4 = e & 15;
    const long long col = c0 + 4 * c4;
    const long long left = L - col;
    const int bytes = left >= 4 ? 16 : (left > 0 ? static_cast<int>(left) * 4 : 0);
    const float* src = bytes > 0 ? Mat + static_cast<long long>(k) * ld + col : Mat;
    cp_async16(base + (k * kUpdPitch + 4 * c4) * 4, src, bytes);
  }
  cp_async_commit();
}

// Z accumulators of warp `mt` (rows 16 mt ..): acc[nt][0..3] over the 8 column groups of the tile.  The three
// products of a k-step are issued product-major (eight independent accumulators between two MMAs on the same one),
// and even / odd k-steps go to two accumulator sets that are added at the end: twice the independent work for the
// tensor pipe and half the length of every round-toward-zero accumulation chain.
template <int NC>
__device__ __forceinline__ void z_product(float (&acc)[8][4], const float* afrag, const float* thi, const float* tlo,
                                          int mt, int ks_used, int lane) {
  constexpr int KS = 2 * NC;
  const int g = lane >> 2, t = lane & 3;
  float acc2[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] = acc2[nt][i] = 0.f;
  const float4* af = reinterpret_cast<const float4*>(afrag) + static_cast<size_t>(mt) * KS * 32 + lane;
  auto step = [&](int ks, float (&d)[8][4]) {
    const float4 a = af[ks * 32];
    uint32_t ahi[4], alo[4];
    ptx::split_tf32(a.x, ahi[0], alo[0]);
    ptx::split_tf32(a.y, ahi[1], alo[1]);
    ptx::split_tf32(a.z, ahi[2], alo[2]);
    ptx::split_tf32(a.w, ahi[3], alo[3]);
    const int o = (8 * ks + t) * kUpdPitch + g;
    uint32_t bh0[8], bh1[8], bl0[8], bl1[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      bh0[nt] = __float_as_uint(thi[o + 8 * nt]);
      bh1[nt] = __float_as_uint(thi[o + 4 * kUpdPitch + 8 * nt]);
      bl0[nt] = __float_as_uint(tlo[o + 8 * nt]);
      bl1[nt] = __float_as_uint(tlo[o + 4 * kUpdPitch + 8 * nt]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_tf32_m16n8k8(d[nt], alo, bh0[nt], bh1[nt]);  // small terms first
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_tf32_m16n8k8(d[nt], ahi, bl0[nt], bl1[nt]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_tf32_m16n8k8(d[nt], ahi, bh0[nt], bh1[nt]);
  };
  int ks = 0;
  for (; ks + 1 < ks_used; ks += 2) {
    step(ks, acc);
    step(ks + 1, acc2);
  }
  if (ks < ks_used) step(ks, acc);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[nt][i] += acc2[nt][i];
}

// Gram of one tile, added (round-to-nearest) to this CTA's running partial in global memory: block (mt_i, nt_i) of
// tile * tile^T over the 64 columns, from the tile's tf32 hi / lo copies.  Fresh accumulators per tile keep the
// tensor core's round-toward-zero chains short (12 MMAs); every thread owns its float4 slots of the partial, so the
// read-modify-write needs no synchronisation.  Fragment order: [(i * 8 + warp) * 2 + j][lane] float4.
template <int NBW>
__device__ __forceinline__ void gram_tile(float* gram_partial, int gram_floats, bool first, const float* thi,
                                          const float* tlo, const int (&bmt)[NBW], const int (&bnt)[NBW], int warp,
                                          int lane) {
  const int g = lane >> 2, t = lane & 3;
  float4* out = reinterpret_cast<float4*>(gram_partial + static_cast<size_t>(blockIdx.x) * gram_floats);
#pragma unroll
  for (int i = 0; i < NBW; ++i) {
    if (bmt[i] < 0) continue;
    float d[2][2][4];  // [j][k-step parity][c0..c3]
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int x = 0; x < 4; ++x) d[j][q][x] = 0.f;
    const int ao = (16 * bmt[i] + g) * kUpdPitch + t, bo = (16 * bnt[i] + g) * kUpdPitch + t;
#pragma unroll
    for (int ks = 0; ks < kUpdCols / 8; ++ks) {
      const int q = ks & 1;
      uint32_t ahi[4], alo[4], bh[2][2], bl[2][2];
      ahi[0] = __float_as_uint(thi[ao + 8 * ks]);
      ahi[1] = __float_as_uint(thi[ao + 8 * kUpdPitch + 8 * ks]);
      ahi[2] = __float_as_uint(thi[ao + 8 * ks + 4]);
      ahi[3] = __float_as_uint(thi[ao + 8 * kUpdPitch + 8 * ks + 4]);
      alo[0] = __float_as_uint(tlo[ao + 8 * ks]);
      alo[1] = __float_as_uint(tlo[ao + 8 * kUpdPitch + 8 * ks]);
      alo[2] = __float_as_uint(tlo[ao + 8 * ks + 4]);
      alo[3] = __float_as_uint(tlo[ao + 8 * kUpdPitch + 8 * ks + 4]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        bh[j][0] = __float_as_uint(thi[bo + 8 * j * kUpdPitch + 8 * ks]);
        bh[j][1] = __float_as_uint(thi[bo + 8 * j * kUpdPitch + 8 * ks + 4]);
        bl[j][0] = __float_as_uint(tlo[bo + 8 * j * kUpdPitch + 8 * ks]);
        bl[j][1] = __float_as_uint(tlo[bo + 8 * j * kUpdPitch + 8 * ks + 4]);
      }
      mma_tf32_m16n8k8(d[0][q], alo, bh[0][0], bh[0][1]);
      mma_tf32_m16n8k8(d[1][q], alo, bh[1][0], bh[1][1]);
      mma_tf32_m16n8k8(d[0][q], ahi, bl[0][0], bl[0][1]);
      mma_tf32_m16n8k8(d[1][q], ahi, bl[1][0], bl[1][1]);
      mma_tf32_m16n8k8(d[0][q], ahi, bh[0][0], bh[0][1]);
      mma_tf32_m16n8k8(d[1][q], ahi, bh[1][0], bh[1][1]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float4* slot = out + ((i * kUpdWarps + warp) * 2 + j) * 32 + lane;
      float4 v = make_float4(d[j][0][0] + d[j][1][0], d[j][0][1] + d[j][1][1], d[j][0][2] + d[j][1][2],
                             d[j][0][3] + d[j][1][3]);
      if (!first) {
        const float4 old = *slot;
        v.x += old.x, v.y += old.y, v.z += old.z, v.w += old.w;
      }
      *slot = v;
    }
  }
}

// tile rows [0, K) -> Mat, its tf32 hi / lo copies (from the tile's own hi / lo copies) and the peers' copies;
// columns >= L are skipped (Mat) / zero (split: pitches are multiples of 4, whole float4 groups are written)
__device__ __forceinline__ void store_tile(const float* tile, const float* thi, const float* tlo, float* __restrict__ Mat,
                                           long long ld, int K, long long c0, long long L, float* __restrict__ split_hi,
                                           float* __restrict__ split_lo, long long ld_split, int n_peers,
                                           float* const* mat_peer) {
  for (int e = threadIdx.x; e < K * 16; e += kUpdThreads) {
    const int k = e >> 4, c4 = e & 15;
    const long long col = c0 + 4 * c4;
    if (col >= L) continue;
    const int o = k * kUpdPitch + 4 * c4;
    const float4 v = *reinterpret_cast<const float4*>(tile + o);
    const bool full = col + 4 <= L;
    float* dst = Mat + static_cast<long long>(k) * ld + col;
    if (full) {
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      for (int x = 0; x < 4; ++x)
        if (col + x < L) dst[x] = vv[x];
    }
    for (int q = 0; q < n_peers; ++q) {
      if (mat_peer[q] == nullptr) continue;
      float* pd = mat_peer[q] + static_cast<long long>(k) * ld + col;
      if (full) {
        *reinterpret_cast<float4*>(pd) = v;
      } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
        for (int x = 0; x < 4; ++x)
          if (col + x < L) pd[x] = vv[x];
      }
    }
    if (split_hi != nullptr) {
      const long long so = static_cast<long long>(k) * ld_split + col;
      *reinterpret_cast<float4*>(split_hi + so) = *reinterpret_cast<const float4*>(thi + o);
      *reinterpret_cast<float4*>(split_lo + so) = *reinterpret_cast<const float4*>(tlo + o);
    }
  }
}

// new value of two adjacent tile entries: fp32 into the tile, tf32 hi / lo into its copies
__device__ __forceinline__ void put2(float* tile, float* thi, float* tlo, int o, float v0, float v1) {
  uint32_t h0, l0, h1, l1;
  ptx::split_tf32(v0, h0, l0);
  ptx::split_tf32(v1, h1, l1);
  *reinterpret_cast<float2*>(tile + o) = make_float2(v0, v1);
  *reinterpret_cast<float2*>(thi + o) = make_float2(__uint_as_float(h0), __uint_as_float(h1));
  *reinterpret_cast<float2*>(tlo + o) = make_float2(__uint_as_float(l0), __uint_as_float(l1));
}

// ---------------------------------------------------------------------------------------------------------------
// W update
struct WUpdParams {
  const float* S;       // [K][ldS]  H H^T (complete: all-reduced / summed over the peers)
  int ldS;
  float* WT;            // [K][ldG]  updated in place
  long long ldG;
  int K;
  long long col0, col1; // genes [col0, col1) are updated by this launch (col0 % 64 == 0)
  UpdNumSrc num;        // X H^T
  float c1, c2, orth, eps;
  float* split_hi;      // [K][ldG] tf32 copies of the new W^T, or nullptr
  float* split_lo;
  int n_peers;
  float* wt_peer[kMaxPeers];
  float* gram_partial;  // [gridDim.x][UpdGeom::gram_floats] or nullptr
};

template <int NC>
inline size_t w_update_smem_bytes() {
  using G = UpdGeom<NC>;
  return (static_cast<size_t>(G::afrag_floats) + G::n_tile_bufs * G::tile_floats) * sizeof(float) +
         kUpdCols * sizeof(double) + 16;
}

template <int NC>
__global__ void __launch_bounds__(kUpdThreads, 1) w_update_kernel(const WUpdParams p) {
  using G = UpdGeom<NC>;
  extern __shared__ __align__(16) uint8_t upd_smem[];
  float* afrag = reinterpret_cast<float*>(upd_smem);
  float* tiles = afrag + G::afrag_floats;   // [2][Kp][72] cp.async targets; the current one receives the new values
  float* thi = tiles + 2 * G::tile_floats;  // [Kp][72] tf32 hi of the current tile (old values, then new)
  float* tlo = thi + G::tile_floats;
  double* cs = reinterpret_cast<double*>(tlo + G::tile_floats);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int K = p.K;
  const int ks_used = (K + 7) >> 3;
  const long long n_tiles = (p.col1 - p.col0 + kUpdCols - 1) / kUpdCols;

  build_afrag<NC>(afrag, tiles, p.S, p.ldS, K);
  for (int e = tid; e < G::n_tile_bufs * G::tile_floats; e += kUpdThreads) tiles[e] = 0.f;  // pad rows stay zero
  __syncthreads();

  int bmt[G::NBW], bnt[G::NBW];
#pragma unroll
  for (int i = 0; i < G::NBW; ++i) {
    const int b = warp + kUpdWarps * i;
    bmt[i] = bnt[i] = -1;
    if (b < G::NB) gram_block_decode(b, NC, bmt[i], bnt[i]);
  }

  long long tile_i = blockIdx.x;
  int buf = 0;
  bool first = true;
  if (tile_i < n_tiles) load_tile_async(tiles, p.WT, p.ldG, K, p.col0 + tile_i * kUpdCols, p.col1);
  for (; tile_i < n_tiles; tile_i += gridDim.x, buf ^= 1, first = false) {
    float* tile = tiles + buf * G::tile_floats;
    const long long c0 = p.col0 + tile_i * kUpdCols;
    // numerator fragments of this warp's rows (issued before anything waits)
    float2 nu[8][2];
    {
      int s0 = 0, s1 = 0;
      if (p.num.direct == nullptr) {
        s0 = __ldg(p.num.slot_ofs + (c0 >> 8));
        s1 = __ldg(p.num.slot_ofs + (c0 >> 8) + 1);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 16 * warp + g + 8 * h;
          const long long col = c0 + 8 * nt + 2 * t;
          nu[nt][h] = (warp < NC && k < K && col < p.col1) ? num_load2(p.num, k, col, s0, s1) : make_float2(0.f, 0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();  // the tile has landed; everybody is done with the previous tile's buffers
    const long long next = tile_i + gridDim.x;
    if (next < n_tiles)
      load_tile_async(tiles + (buf ^ 1) * G::tile_floats, p.WT, p.ldG, K, p.col0 + next * kUpdCols, p.col1);
    split_tile(tile, thi, tlo, K);
    // rowsum_K W[g][:] per gene, fp64 so that (rowsum - w) does not cancel (the reference sums the other K-1 entries)
    if (tid < kUpdCols) {
      double sacc = 0.0;
      for (int k = 0; k < K; ++k) sacc += static_cast<double>(tile[k * kUpdPitch + tid]);
      cs[tid] = sacc;
    }
    __syncthreads();
    float acc[8][4];
    if (warp < NC) z_product<NC>(acc, afrag, thi, tlo, warp, ks_used, lane);
    __syncthreads();  // every warp has read the old hi / lo copies
    if (warp < NC) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 16 * warp + g + 8 * h;
          if (k >= K) continue;
          const int cl = 8 * nt + 2 * t;
          const int o = k * kUpdPitch + cl;
          const float2 old = *reinterpret_cast<const float2*>(tile + o);
          const float oldv[2] = {old.x, old.y}, numv[2] = {nu[nt][h].x, nu[nt][h].y};
          float outv[2];
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const float others = static_cast<float>(cs[cl + x] - static_cast<double>(oldv[x]));
            float den = (2.0f * acc[nt][2 * h + x] + p.c1 * oldv[x]) + p.orth * others;  // main.py:599-601
            den += p.c2;                                                                // main.py:603
            den = fmaxf(den, p.eps);                                                    // main.py:604
            outv[x] = (c0 + cl + x < p.col1) ? oldv[x] * ((2.0f * numv[x]) / den) : 0.f;  // main.py:596, 605
          }
          put2(tile, thi, tlo, o, outv[0], outv[1]);
        }
    }
    __syncthreads();  // the tile and its hi / lo copies hold the new values
    store_tile(tile, thi, tlo, p.WT, p.ldG, K, c0, p.col1, p.split_hi, p.split_lo, p.ldG, p.n_peers, p.wt_peer);
    if (p.gram_partial != nullptr) gram_tile<G::NBW>(p.gram_partial, G::gram_floats, first, thi, tlo, bmt, bnt, warp, lane);
    // (the next iteration's first barrier separates these reads from the next split / prefetch)
  }
  cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// H update
struct HUpdParams {
  const float* T;  // [K][ldT]  W^T W (transform: of the fixed W)
  int ldT;
  float* H;        // [K][ldH]  updated in place
  long long ldH;
  int K;
  long long n;
  UpdNumSrc num;   // W^T X
  float eps;
  CovTable cov;
  int loss_type;
  int Kg, c_total, q_total;
  float* split_hi;  // [K][ld_split] or nullptr
  float* split_lo;
  long long ld_split;
  float* gram_partial;    // [gridDim.x][gram_floats]
  float* hsum_partial;    // [gridDim.x][K]
  float* q_partial;       // [gridDim.x][q_total]
  double* pred_partial;   // [gridDim.x][n_cov]
  double* t1_partial;     // [gridDim.x]
};

template <int NC>
inline size_t h_update_smem_bytes(int K, int Kg, int c_total, int q_total) {
  using G = UpdGeom<NC>;
  const size_t q_pad = (static_cast<size_t>(q_total) + 3) & ~size_t(3);
  size_t f = static_cast<size_t>(G::afrag_floats) + G::n_tile_bufs * G::tile_floats;
  f += q_pad * 2 + 2 * static_cast<size_t>(c_total) * kUpdCols + Kg + K;  // Bs, qacc, rn, rd, dcol, hacc
  return f * sizeof(float) + (kUpdThreads / 32) * sizeof(double) + static_cast<size_t>(Kg + 1 + kMaxCov) * sizeof(int) + 32;
}

// FIT: the full update with statistics; otherwise the transform update H *= 2A / max(2 T H, eps) (main.py:705-709)
template <int NC, bool FIT>
__global__ void __launch_bounds__(kUpdThreads, 1) h_update_kernel(const HUpdParams p) {
  using G = UpdGeom<NC>;
  extern __shared__ __align__(16) uint8_t upd_smem[];
  float* afrag = reinterpret_cast<float*>(upd_smem);
  float* tiles = afrag + G::afrag_floats;
  const int q_pad = (p.q_total + 3) & ~3;           // keeps rn / rd 16-byte aligned
  float* thi = tiles + 2 * G::tile_floats;          // tf32 hi / lo of the current tile (old values, then new)
  float* tlo = thi + G::tile_floats;
  float* Bs = tlo + G::tile_floats;                 // [q_total]  all B_i, row-major [c][k] at q_off
  float* qacc = Bs + q_pad;                         // [q_total]  running Q partial of this CTA
  float* rn = qacc + q_pad;                         // [c_total][64]  per-cell numerator ratios (rho / Y)
  float* rd = rn + static_cast<size_t>(p.c_total) * kUpdCols;  // [c_total][64]  Frobenius: B H_i
  float* dcol = rd + static_cast<size_t>(p.c_total) * kUpdCols;  // [Kg]  KL: scale * colsum(B) per guided row
  float* hacc = dcol + p.Kg;                        // [K]  running row sums of the new H
  double* red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(hacc + p.K) + 7) & ~uintptr_t(7));  // [8]
  int* rowcov = reinterpret_cast<int*>(red + kUpdThreads / 32);  // [Kg]  covariate of guided row k
  int* rowc0 = rowcov + p.Kg + 1;                   // [n_cov]-indexed by covariate: first row of rn / rd (category offset)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int K = p.K;
  const int ks_used = (K + 7) >> 3;
  const long long n_tiles = (p.n + kUpdCols - 1) / kUpdCols;
  const float scale_kl = 1.0f, scale_fr = 2.0f;

  build_afrag<NC>(afrag, tiles, p.T, p.ldT, K);
  for (int e = tid; e < G::n_tile_bufs * G::tile_floats; e += kUpdThreads) tiles[e] = 0.f;
  if (FIT) {
    int coff = 0;
    for (int i = 0; i < p.cov.n_cov; ++i) {
      const CovDesc d = p.cov.d[i];
      for (int e = tid; e < d.c * d.k; e += kUpdThreads) Bs[d.q_off + e] = d.B[e];
      for (int k = tid; k < d.k; k += kUpdThreads) rowcov[d.row0 + k] = i;
      if (tid == 0) rowc0[i] = coff;
      coff += d.c;
    }
    for (int e = tid; e < p.q_total; e += kUpdThreads) qacc[e] = 0.f;
    for (int k = tid; k < K; k += kUpdThreads) hacc[k] = 0.f;
  }
  __syncthreads();
  if (FIT) {
    // KL: denG = (lam B^T) 1 = lam * colsum(B), the same for every cell (main.py:644)
    for (int k = tid; k < p.Kg; k += kUpdThreads) {
      const CovDesc d = p.cov.d[rowcov[k]];
      float s = 0.f;
      for (int c = 0; c < d.c; ++c) s += (scale_kl * d.lam * Bs[d.q_off + c * d.k + (k - d.row0)]) * 1.0f;
      dcol[k] = s;
    }
  }

  int bmt[G::NBW], bnt[G::NBW];
#pragma unroll
  for (int i = 0; i < G::NBW; ++i) {
    const int b = warp + kUpdWarps * i;
    bmt[i] = bnt[i] = -1;
    if (b < G::NB) gram_block_decode(b, NC, bmt[i], bnt[i]);
  }
  double t1 = 0.0, pl0 = 0.0, pl1 = 0.0;  // pred loss of covariates (tid / 64) and (tid / 64 + 4)

  long long tile_i = blockIdx.x;
  int buf = 0;
  bool first = true;
  if (tile_i < n_tiles) load_tile_async(tiles, p.H, p.ldH, K, tile_i * kUpdCols, p.n);
  for (; tile_i < n_tiles; tile_i += gridDim.x, buf ^= 1, first = false) {
    float* tile = tiles + buf * G::tile_floats;
    const long long c0 = tile_i * kUpdCols;
    float2 nu[8][2];
    {
      int s0 = 0, s1 = 0;
      if (p.num.direct == nullptr) {
        s0 = __ldg(p.num.slot_ofs + (c0 >> 8));
        s1 = __ldg(p.num.slot_ofs + (c0 >> 8) + 1);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 16 * warp + g + 8 * h;
          const long long col = c0 + 8 * nt + 2 * t;
          nu[nt][h] = (warp < NC && k < K && col < p.n) ? num_load2(p.num, k, col, s0, s1) : make_float2(0.f, 0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const long long next = tile_i + gridDim.x;
    if (next < n_tiles) load_tile_async(tiles + (buf ^ 1) * G::tile_floats, p.H, p.ldH, K, next * kUpdCols, p.n);
    split_tile(tile, thi, tlo, K);
    if (FIT) {
      // guided terms of the OLD H with the NEW B, per cell (main.py:637-650): thread (i, j) = (tid / 64 [+4], tid % 64)
      const int j = tid & 63;
      const bool live = c0 + j < p.n;
      for (int i = tid >> 6; i < p.cov.n_cov; i += 4) {
        const CovDesc d = p.cov.d[i];
        const int cbase = rowc0[i];
        for (int c = 0; c < d.c; ++c) {
          float yhat = 0.f;
          for (int k = 0; k < d.k; ++k) yhat += Bs[d.q_off + c * d.k + k] * tile[(d.row0 + k) * kUpdPitch + j];
          const float y = live ? __ldg(d.Y + static_cast<long long>(c) * p.n + c0 + j) : 0.f;
          rn[(cbase + c) * kUpdCols + j] = (p.loss_type == LOSS_KL) ? y / fmaxf(yhat, p.eps) : y;
          rd[(cbase + c) * kUpdCols + j] = yhat;
        }
      }
    }
    __syncthreads();  // hi / lo copies and rn / rd are complete
    float acc[8][4];
    if (warp < NC) z_product<NC>(acc, afrag, thi, tlo, warp, ks_used, lane);
    __syncthreads();  // every warp has read the old hi / lo copies
    if (warp < NC) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = 16 * warp + g + 8 * h;
          if (k >= K) continue;
          const int cl = 8 * nt + 2 * t;
          const int o = k * kUpdPitch + cl;
          const float2 old = *reinterpret_cast<const float2*>(tile + o);
          const float oldv[2] = {old.x, old.y}, numv[2] = {nu[nt][h].x, nu[nt][h].y};
          float gn[2] = {0.f, 0.f}, gd[2] = {0.f, 0.f};
          if (FIT && k < p.Kg) {
            const CovDesc d = p.cov.d[rowcov[k]];
            const int cbase = rowc0[rowcov[k]];
            const float scale = (p.loss_type == LOSS_KL) ? scale_kl * d.lam : scale_fr * d.lam;
            for (int c = 0; c < d.c; ++c) {
              const float lb = scale * Bs[d.q_off + c * d.k + (k - d.row0)];
              const float2 r = *reinterpret_cast<const float2*>(rn + (cbase + c) * kUpdCols + cl);
              gn[0] += lb * r.x, gn[1] += lb * r.y;
              if (p.loss_type != LOSS_KL) {
                const float2 q = *reinterpret_cast<const float2*>(rd + (cbase + c) * kUpdCols + cl);
                gd[0] += lb * q.x, gd[1] += lb * q.y;
              }
            }
            if (p.loss_type == LOSS_KL) gd[0] = gd[1] = dcol[k];
          }
          float outv[2];
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const float num = gn[x] + 2.0f * numv[x];                        // main.py:648, 653 / 706
            const float den = fmaxf(gd[x] + 2.0f * acc[nt][2 * h + x], p.eps);  // main.py:649, 654-655 / 707-708
            outv[x] = (c0 + cl + x < p.n) ? oldv[x] * (num / den) : 0.f;     // main.py:656 / 709
            if (FIT) t1 += static_cast<double>(numv[x]) * static_cast<double>(outv[x]);
          }
          put2(tile, thi, tlo, o, outv[0], outv[1]);
        }
    }
    __syncthreads();  // the tile and its hi / lo copies hold the new H
    store_tile(tile, thi, tlo, p.H, p.ldH, K, c0, p.n, p.split_hi, p.split_lo, p.ld_split, 0, nullptr);
    if (FIT) {
      // statistics of (new H, new B): rho' = Y / max(B H_i, eps) per cell, prediction loss (main.py:727-748)
      const int j = tid & 63;
      const bool live = c0 + j < p.n;
      int slot = 0;
      for (int i = tid >> 6; i < p.cov.n_cov; i += 4, ++slot) {
        const CovDesc d = p.cov.d[i];
        const int cbase = rowc0[i];
        double pl = 0.0;
        for (int c = 0; c < d.c; ++c) {
          float yhat = 0.f;
          for (int k = 0; k < d.k; ++k) yhat += Bs[d.q_off + c * d.k + k] * tile[(d.row0 + k) * kUpdPitch + j];
          const float y = live ? __ldg(d.Y + static_cast<long long>(c) * p.n + c0 + j) : 0.f;
          float r;
          if (p.loss_type == LOSS_KL) {
            const float yh = fmaxf(yhat, p.eps);
            r = y / yh;
            if (live) pl += static_cast<double>(y * logf(fmaxf(y / yh, p.eps)) - y + yh);
          } else {
            r = y;
            const float dlt = y - yhat;
            if (live) pl += static_cast<double>(dlt * dlt);
          }
          rn[(cbase + c) * kUpdCols + j] = live ? r : 0.f;
        }
        if (slot == 0) pl0 += pl; else pl1 += pl;
      }
      __syncthreads();  // rn now holds rho' of the new H
      // Q_i += rho' H_i^T over this tile's cells; row sums of the new H  (one owner thread per entry: fixed order)
      for (int e = tid; e < p.q_total + K; e += kUpdThreads) {
        if (e < p.q_total) {
          int i = 0;
          while (i + 1 < p.cov.n_cov && e >= p.cov.d[i + 1].q_off) ++i;
          const CovDesc d = p.cov.d[i];
          const int c = (e - d.q_off) / d.k, k = (e - d.q_off) - c * d.k;
          const float* rrow = rn + (rowc0[i] + c) * kUpdCols;
          const float* hrow = tile + (d.row0 + k) * kUpdPitch;
          float a = 0.f;
          for (int u = 0; u < kUpdCols; ++u) a += rrow[u] * hrow[u];
          qacc[e] += a;
        } else {
          const int k = e - p.q_total;
          const float* hrow = tile + k * kUpdPitch;
          float a = 0.f;
          for (int u = 0; u < kUpdCols; ++u) a += hrow[u];
          hacc[k] += a;
        }
      }
      gram_tile<G::NBW>(p.gram_partial, G::gram_floats, first, thi, tlo, bmt, bnt, warp, lane);
    }
  }
  cp_async_wait_all();
  if (!FIT) return;
  __syncthreads();
  for (int e = tid; e < p.q_total; e += kUpdThreads) p.q_partial[static_cast<size_t>(blockIdx.x) * p.q_total + e] = qacc[e];
  for (int k = tid; k < K; k += kUpdThreads) p.hsum_partial[static_cast<size_t>(blockIdx.x) * K + k] = hacc[k];
  // fixed-order block reductions of the fp64 scalars
  for (int o = 16; o > 0; o >>= 1) t1 += __shfl_down_sync(0xffffffffu, t1, o);
  if (lane == 0) red[warp] = t1;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < kUpdWarps; ++w) s += red[w];
    p.t1_partial[blockIdx.x] = s;
  }
  // prediction loss: warps (2i, 2i+1) hold covariate i in pl0 and covariate i + 4 in pl1
  for (int pass = 0; pass < 2; ++pass) {
    double v = pass == 0 ? pl0 : pl1;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid < 4) {
      const int i = tid + 4 * pass;
      if (i < p.cov.n_cov) p.pred_partial[static_cast<size_t>(blockIdx.x) * p.cov.n_cov + i] = red[2 * tid] + red[2 * tid + 1];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Finish kernels: fixed-order sums of the per-CTA partials
struct GramReduce {
  const float* partial;  // [n_parts][gram_floats]
  int n_parts, NC, K;
  float* out;            // [K][ld]
  int ld;
};
// one thread per float4 of the fragment layout; off-diagonal blocks are mirrored
__device__ __forceinline__ void gram_reduce_thread(const GramReduce& r, int idx) {
  const int NB = r.NC * (r.NC + 1) / 2;
  const int NBW = (NB + kUpdWarps - 1) / kUpdWarps;
  const int gram_floats = NBW * kUpdWarps * 2 * 32 * 4;
  if (idx >= NBW * kUpdWarps * 2 * 32) return;
  const int lane = idx & 31, j = (idx >> 5) & 1, w = (idx >> 6) % kUpdWarps, i = (idx >> 6) / kUpdWarps;
  const int b = w + kUpdWarps * i;
  if (b >= NB) return;
  int mt, nt;
  gram_block_decode(b, r.NC, mt, nt);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* src = reinterpret_cast<const float4*>(r.partial) + idx;
#pragma unroll 4
  for (int q = 0; q < r.n_parts; ++q) {
    const float4 v = __ldcg(src + static_cast<size_t>(q) * (gram_floats / 4));
    acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
  }
  const int g = lane >> 2, t = lane & 3;
  const int row = 16 * mt + g, col = 16 * nt + 8 * j + 2 * t;
  const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int rr = row + 8 * (x >> 1), cc = col + (x & 1);
    if (rr < r.K && cc < r.K) {
      r.out[rr * r.ld + cc] = v[x];
      if (mt != nt) r.out[cc * r.ld + rr] = v[x];
    }
  }
}
inline int gram_reduce_threads(int NC) {
  const int NB = NC * (NC + 1) / 2;
  return ((NB + kUpdWarps - 1) / kUpdWarps) * kUpdWarps * 2 * 32;
}

struct WFinishParams {
  GramReduce gram;   // -> T = W^T W  (n_parts == 0: T comes from elsewhere)
  int gram_blocks;
  CovTable cov;      // B updates (main.py:615-628)
  int loss_type;
  const float* stats_q;
  const float* hsum;
  const float* S;
  int ldS;
  float eps;
};
__global__ void __launch_bounds__(256) w_finish_kernel(const WFinishParams p) {
  if (static_cast<int>(blockIdx.x) < p.gram_blocks) {
    gram_reduce_thread(p.gram, blockIdx.x * 256 + threadIdx.x);
    return;
  }
  // the last block: every B_i, from the statistics of the old H / old B
  extern __shared__ float bs[];  // old B of one covariate [c][k]
  for (int i = 0; i < p.cov.n_cov; ++i) {
    const CovDesc d = p.cov.d[i];
    const float* Q = p.stats_q + d.q_off;
    for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) bs[e] = d.B[e];
    __syncthreads();
    for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) {
      const int c = e / d.k, k = e - c * d.k;
      float num, den;
      if (p.loss_type == LOSS_KL) {
        num = d.lam * Q[e];
        den = d.lam * p.hsum[d.row0 + k];
      } else {
        num = 2.0f * Q[e];
        float acc = 0.f;
        for (int k2 = 0; k2 < d.k; ++k2) acc += (2.0f * bs[c * d.k + k2]) * p.S[(d.row0 + k2) * p.ldS + d.row0 + k];
        den = acc;
      }
      den = fmaxf(den, p.eps);
      d.B[e] = bs[e] * (num / den);
    }
    __syncthreads();
  }
}

struct HFinishParams {
  GramReduce gram;   // -> S = H H^T of this shard
  int gram_blocks;
  int n_parts;
  const float* hsum_partial;  // [n_parts][K]
  float* hsum;                // [K]
  const float* q_partial;     // [n_parts][q_total]
  float* stats_q;             // [q_total]
  int q_total;
  const double* pred_partial;  // [n_parts][n_cov]
  int n_cov;
  const double* t1_partial;    // [n_parts]
  const float* T;              // [K][ldT]
  int ldT;
  double* loss_row;            // [2 + n_cov] or nullptr
  unsigned int* counter;       // zero before the first launch; left zero
};
// blocks [0, gram_blocks): Gram; block gram_blocks: hsum, Q, pred, t1.  The block that finishes last (all of S is
// then in memory) takes t2 = sum T .* S in a fixed order.
__global__ void __launch_bounds__(256) h_finish_kernel(const HFinishParams p) {
  __shared__ double red[256];
  __shared__ unsigned int last;
  const int K = p.gram.K;
  if (static_cast<int>(blockIdx.x) < p.gram_blocks) {
    gram_reduce_thread(p.gram, blockIdx.x * 256 + threadIdx.x);
  } else {
    for (int k = threadIdx.x; k < K; k += 256) {
      double a = 0.0;
      for (int q = 0; q < p.n_parts; ++q) a += static_cast<double>(p.hsum_partial[static_cast<size_t>(q) * K + k]);
      p.hsum[k] = static_cast<float>(a);
    }
    for (int e = threadIdx.x; e < p.q_total; e += 256) {
      double a = 0.0;
      for (int q = 0; q < p.n_parts; ++q) a += static_cast<double>(p.q_partial[static_cast<size_t>(q) * p.q_total + e]);
      p.stats_q[e] = static_cast<float>(a);
    }
    if (p.loss_row != nullptr) {
      for (int which = 0; which < 1 + p.n_cov; ++which) {  // t1, pred_0 ..
        double a = 0.0;
        for (int q = threadIdx.x; q < p.n_parts; q += 256)
          a += which == 0 ? p.t1_partial[q] : p.pred_partial[static_cast<size_t>(q) * p.n_cov + (which - 1)];
        const double s = block_sum_256(a, red);
        if (threadIdx.x == 0) p.loss_row[which == 0 ? 0 : 1 + which] = s;
      }
    }
  }
  if (p.loss_row == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(p.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last == 0u) return;
  __threadfence();
  double a = 0.0;
  for (int e = threadIdx.x; e < K * K; e += 256) {
    const int r = e / K, c = e - r * K;
    a += static_cast<double>(__ldcg(p.T + r * p.ldT + c)) * static_cast<double>(__ldcg(p.gram.out + r * p.gram.ld + c));
  }
  const double s = block_sum_256(a, red);
  if (threadIdx.x == 0) {
    p.loss_row[1] = s;
    *p.counter = 0u;
  }
}

}  // namespace alpine
