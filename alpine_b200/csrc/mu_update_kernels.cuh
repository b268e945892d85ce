// The multiplicative updates of one MU iteration as two fused epilogue kernels (+ two small finish kernels).
//
// All four K x K-sized products of the Gram reformulation run on the tcgen05 contraction kernel
// (csrc/mu_gemm_sm100.cuh) as small plans -- Z_W = (H H^T) W^T, T = W^T W, Z_H = T H, S = H H^T -- and leave their
// results in stream-K partial-sum slots; the kernels here consume those slots directly:
//
//   w_update_kernel  (main.py:596-612)  per 64-gene tile of W^T [K][G]:
//       P = X H^T (slots of the big contraction, or the all-reduced exchange buffer), Z_W (slots)
//       W *= 2P / max(2 Z_W + (1-l1) a W + orth (rowsum_K(W) - W) + l1 a, eps);  tf32 hi/lo copies of the new tile
//       (B operand of the next contraction);  peer stores under cell sharding
//   w_finish_kernel  T = W^T W from the Gram plan's slots (+ its hi/lo copies, B operand of Z_H) and the B updates
//       (main.py:615-628)
//   h_update_kernel  (main.py:631-663)  per 64-cell tile of H [K][n]:
//       A = W^T X (slots), Z_H (slots), guided terms from (old H, new B) per cell,
//       H *= (numG + 2A) / max(denG + 2 Z_H, eps);  hi/lo copies;  t1 = sum A .* H_new (fp64);
//       statistics of (new H, new B) for the next iteration: rowsum(H), Q_i, prediction loss
//   h_finish_kernel  S = H H^T from the Gram plan's slots (+ hi/lo copies, B operand of Z_W), rowsum(H), Q_i into
//       the exchange buffer, and the loss row [t1, t2 = sum T .* S, pred_i] (main.py:666, 726-753, trace identity)
//
// Measured dead end (round 2, kept in the history): the same products on mma.sync.m16n8k8.tf32 inside these kernels.
// Parity was green, but the legacy tensor path of sm_100 issues one such MMA per ~100 cycles per SM sub-partition
// (h_update 470 us at cfg 3, slower than the CUDA-core version it replaced), so the products moved to tcgen05.
// Everything is deterministic: fixed tile -> CTA assignment, fixed summation orders, no atomics on data.
#pragma once
#include "mu_small_kernels.cuh"

namespace alpine {

constexpr int kUpdCols = 64;       // columns (cells or genes) per tile
constexpr int kUpdPitch = 68;      // shared-memory row pitch of a tile (floats)
constexpr int kUpdThreads = 256;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// A [K][L] operand of an update: a finished array, or the stream-K partial-sum slots of the contraction that has
// just produced it (slot s holds [K][256] sums of one segment of a 256-row super-tile; the slots of a super-tile
// are added in the fixed order of its list).
struct SlotSrc {
  const float* direct;  // [K][ld] or nullptr
  long long ld;
  const float* partial;
  const int* slot_ofs;  // [super-tiles + 1]
  const int* slots;
  int K;                // components per slot
  long long origin;     // column of the caller's matrix that is row 0 of the plan's first super-tile
};
// Up to 256 components run as two contraction launches over component groups [0, split) and [split, K) (the tcgen05
// kernel accumulates at most 128 columns per launch): an operand is then two slot sets, selected by the row.
struct SlotSrc2 {
  SlotSrc g[2];
  int split;  // first row of the second group (>= K: one group)
};

// The slot ids of the tile's super-tile are staged in shared memory once per tile (ids[0 .. n)), so the loads of an
// element are independent of each other; they are issued four at a time and added in list order.
constexpr int kMaxTileSlots = 96;
// (a super-tile with more slots than the staging array -- very long reductions cut into many pieces -- is read
// through the global list instead)
__device__ __forceinline__ int stage_slot_ids(const SlotSrc& s, long long c0, int* ids, const int** idp) {
  *idp = ids;
  if (s.direct != nullptr) return 0;
  const long long t = (c0 - s.origin) >> 8;
  const int s0 = __ldg(s.slot_ofs + t), n = __ldg(s.slot_ofs + t + 1) - s0;
  if (n > kMaxTileSlots) {
    *idp = s.slots + s0;
    return n;
  }
  for (int q = threadIdx.x; q < n; q += kUpdThreads) ids[q] = __ldg(s.slots + s0 + q);
  return n;
}
// four adjacent columns (col .. col + 3) of row k; col % 4 == 0
__device__ __forceinline__ float4 slot_load4(const SlotSrc& s, int k, long long col, const int* ids, int n) {
  if (s.direct != nullptr) return __ldg(reinterpret_cast<const float4*>(s.direct + static_cast<long long>(k) * s.ld + col));
  const float* base = s.partial + static_cast<size_t>(k) * 256 + static_cast<int>((col - s.origin) & 255);
  const size_t stride = static_cast<size_t>(s.K) * 256;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int q = 0;
  for (; q + 4 <= n; q += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(base + ids[q + u] * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) acc.x += v[u].x, acc.y += v[u].y, acc.z += v[u].z, acc.w += v[u].w;
  }
  for (; q < n; ++q) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(base + ids[q] * stride));
    acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
  }
  return acc;
}

// the common case of at most two slots (or a finished array): both loads are issued without being consumed, so that
// a thread can have the loads of several elements in flight; sum = a + b
struct SlotPair {
  float4 a, b;
};
__device__ __forceinline__ SlotPair slot_load_pair(const SlotSrc& s, int k, long long col, const int* ids, int n) {
  SlotPair r;
  r.b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.direct != nullptr) {
    r.a = __ldg(reinterpret_cast<const float4*>(s.direct + static_cast<long long>(k) * s.ld + col));
    return r;
  }
  const float* base = s.partial + static_cast<size_t>(k) * 256 + static_cast<int>((col - s.origin) & 255);
  const size_t stride = static_cast<size_t>(s.K) * 256;
  r.a = n > 0 ? __ldcg(reinterpret_cast<const float4*>(base + ids[0] * stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (n > 1) r.b = __ldcg(reinterpret_cast<const float4*>(base + ids[1] * stride));
  return r;
}
__device__ __forceinline__ float4 pair_sum(const SlotPair& p) {
  return make_float4(p.a.x + p.b.x, p.a.y + p.b.y, p.a.z + p.b.z, p.a.w + p.b.w);
}

// The same for up to S slots (small shards: a super-tile of the W^T X plan is shared by three or four CTAs): all
// loads are issued before any is consumed; the sum adds them in list order, exactly like slot_load4.
template <int S>
struct SlotSet {
  float4 v[S];
};
template <int S>
__device__ __forceinline__ SlotSet<S> slot_load_set(const SlotSrc& s, int k, long long col, const int* ids, int n) {
  SlotSet<S> r;
#pragma unroll
  for (int u = 0; u < S; ++u) r.v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.direct != nullptr) {
    r.v[0] = __ldg(reinterpret_cast<const float4*>(s.direct + static_cast<long long>(k) * s.ld + col));
    return r;
  }
  const float* base = s.partial + static_cast<size_t>(k) * 256 + static_cast<int>((col - s.origin) & 255);
  const size_t stride = static_cast<size_t>(s.K) * 256;
#pragma unroll
  for (int u = 0; u < S; ++u)
    if (u < n) r.v[u] = __ldcg(reinterpret_cast<const float4*>(base + ids[u] * stride));
  return r;
}
template <int S>
__device__ __forceinline__ float4 set_sum(const SlotSet<S>& r) {
  float4 acc = r.v[0];
#pragma unroll
  for (int u = 1; u < S; ++u) acc.x += r.v[u].x, acc.y += r.v[u].y, acc.z += r.v[u].z, acc.w += r.v[u].w;
  return acc;
}

// rows [0, K) x 64 columns [c0, c0 + 64) of Mat [K][ld] -> tile [K][68]; columns >= L arrive as zeros
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

// one float4 of new values -> Mat (masked at L), its split copies for the next contraction (ptx::store_split4; whole groups of four: pitches are
// multiples of 4 and the values at columns >= L are zero) and the peers' copies of Mat
__device__ __forceinline__ void store4(float4 v, int k, long long col, long long L, float* __restrict__ Mat, long long ld,
                                       float* __restrict__ split_hi, float* __restrict__ split_lo, long long ld_split,
                                       int n_peers, float* const* mat_peer) {
  const bool full = col + 4 <= L;
  const float vv[4] = {v.x, v.y, v.z, v.w};
  float* dst = Mat + static_cast<long long>(k) * ld + col;
  if (full) {
    *reinterpret_cast<float4*>(dst) = v;
  } else {
    for (int x = 0; x < 4; ++x)
      if (col + x < L) dst[x] = vv[x];
  }
  for (int q = 0; q < n_peers; ++q) {
    if (mat_peer[q] == nullptr) continue;
    float* pd = mat_peer[q] + static_cast<long long>(k) * ld + col;
    if (full) {
      *reinterpret_cast<float4*>(pd) = v;
    } else {
      for (int x = 0; x < 4; ++x)
        if (col + x < L) pd[x] = vv[x];
    }
  }
  if (split_hi != nullptr) ptx::store_split4(v, split_hi, split_lo, k, col, ld_split);
}

// ---------------------------------------------------------------------------------------------------------------
// W update
struct WUpdParams {
  float* WT;            // [K][ldG]  updated in place
  long long ldG;
  int K;
  long long col0, col1; // genes [col0, col1) are updated by this launch (col0 % 64 == 0)
  SlotSrc2 num;         // X H^T
  SlotSrc2 z;           // (H H^T) W^T
  float c1, c2, orth, eps;
  float* split_hi;      // [K][ldG] tf32 copies of the new W^T, or nullptr
  float* split_lo;
  int n_peers;
  float* wt_peer[kMaxPeers];
};
inline size_t w_update_smem_bytes(int K) {
  return static_cast<size_t>(2) * K * kUpdPitch * sizeof(float) + kUpdCols * sizeof(double) + 16;
}

__global__ void __launch_bounds__(kUpdThreads, 3) w_update_kernel(const WUpdParams p) {
  extern __shared__ __align__(16) uint8_t upd_smem[];
  __shared__ int ids_num[2][kMaxTileSlots], ids_z[2][kMaxTileSlots];
  ptx::pdl_enter();
  const int K = p.K;
  float* tiles = reinterpret_cast<float*>(upd_smem);  // [2][K][68]
  double* cs = reinterpret_cast<double*>(tiles + 2 * K * kUpdPitch);
  const int tid = threadIdx.x;
  const long long n_tiles = (p.col1 - p.col0 + kUpdCols - 1) / kUpdCols;
  long long tile_i = blockIdx.x;
  int buf = 0;
  if (tile_i < n_tiles) load_tile_async(tiles, p.WT, p.ldG, K, p.col0 + tile_i * kUpdCols, p.col1);
  for (; tile_i < n_tiles; tile_i += gridDim.x, buf ^= 1) {
    const float* tile = tiles + buf * K * kUpdPitch;
    const long long c0 = p.col0 + tile_i * kUpdCols;
    cp_async_wait_all();
    __syncthreads();  // the tile has landed; everybody is done with the other buffer and the slot ids
    const int *idn[2], *idz[2];
    int nn[2], nz[2];
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      idn[gi] = idz[gi] = ids_num[0];
      nn[gi] = (gi == 0 || p.num.split < K) ? stage_slot_ids(p.num.g[gi], c0, ids_num[gi], &idn[gi]) : 0;
      nz[gi] = (gi == 0 || p.z.split < K) ? stage_slot_ids(p.z.g[gi], c0, ids_z[gi], &idz[gi]) : 0;
    }
    const bool few_slots = nn[0] <= 2 && nn[1] <= 2 && nz[0] <= 2 && nz[1] <= 2;
    const long long next = tile_i + gridDim.x;
    if (next < n_tiles)
      load_tile_async(tiles + (buf ^ 1) * K * kUpdPitch, p.WT, p.ldG, K, p.col0 + next * kUpdCols, p.col1);
    // rowsum_K W[g][:] per gene, fp64 so that (rowsum - w) does not cancel (the reference sums the other K-1 entries)
    if (tid < kUpdCols) {
      double sacc = 0.0;
      for (int k = 0; k < K; ++k) sacc += static_cast<double>(tile[k * kUpdPitch + tid]);
      cs[tid] = sacc;
    }
    __syncthreads();
    auto update = [&](int e, float4 num4, float4 z4) {
      const int k = e >> 4, c4 = e & 15;
      const long long col = c0 + 4 * c4;
      const float4 old4 = *reinterpret_cast<const float4*>(tile + k * kUpdPitch + 4 * c4);
      const float oldv[4] = {old4.x, old4.y, old4.z, old4.w}, numv[4] = {num4.x, num4.y, num4.z, num4.w};
      const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
      float outv[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float others = static_cast<float>(cs[4 * c4 + x] - static_cast<double>(oldv[x]));
        float den = (2.0f * zv[x] + p.c1 * oldv[x]) + p.orth * others;  // main.py:599-601
        den += p.c2;                                                   // main.py:603
        den = fmaxf(den, p.eps);                                       // main.py:604
        outv[x] = (col + x < p.col1) ? oldv[x] * ((2.0f * numv[x]) / den) : 0.f;  // main.py:596, 605
      }
      store4(make_float4(outv[0], outv[1], outv[2], outv[3]), k, col, p.col1, p.WT, p.ldG, p.split_hi, p.split_lo, p.ldG,
             p.n_peers, p.wt_peer);
    };
    const int n_items = K * 16;
    if (few_slots) {
      // two elements per thread and round: their (up to eight) loads are in flight together
      for (int e = tid; e < n_items; e += 2 * kUpdThreads) {
        const int e2 = e + kUpdThreads;
        const bool v1 = c0 + 4 * (e & 15) < p.col1, v2 = e2 < n_items && c0 + 4 * (e2 & 15) < p.col1;
        SlotPair n1{}, z1{}, n2{}, z2{};
        if (v1) {
          const int k = e >> 4, gn = k >= p.num.split, gz = k >= p.z.split;
          n1 = slot_load_pair(p.num.g[gn], k - gn * p.num.split, c0 + 4 * (e & 15), idn[gn], nn[gn]);
          z1 = slot_load_pair(p.z.g[gz], k - gz * p.z.split, c0 + 4 * (e & 15), idz[gz], nz[gz]);
        }
        if (v2) {
          const int k = e2 >> 4, gn = k >= p.num.split, gz = k >= p.z.split;
          n2 = slot_load_pair(p.num.g[gn], k - gn * p.num.split, c0 + 4 * (e2 & 15), idn[gn], nn[gn]);
          z2 = slot_load_pair(p.z.g[gz], k - gz * p.z.split, c0 + 4 * (e2 & 15), idz[gz], nz[gz]);
        }
        if (v1) update(e, pair_sum(n1), pair_sum(z1));
        if (v2) update(e2, pair_sum(n2), pair_sum(z2));
      }
    } else {
      for (int e = tid; e < n_items; e += kUpdThreads) {
        const long long col = c0 + 4 * (e & 15);
        if (col >= p.col1) continue;
        const int k = e >> 4, gn = k >= p.num.split, gz = k >= p.z.split;
        update(e, slot_load4(p.num.g[gn], k - gn * p.num.split, col, idn[gn], nn[gn]),
               slot_load4(p.z.g[gz], k - gz * p.z.split, col, idz[gz], nz[gz]));
      }
    }
    // (the next iteration's first barrier separates these reads from the prefetch into this buffer two tiles on)
  }
  cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------
// H update
struct HUpdParams {
  float* H;        // [K][ldH]  updated in place
  long long ldH;
  int K;
  long long n;
  SlotSrc2 num;    // W^T X
  SlotSrc2 z;      // (W^T W) H
  float eps;
  CovTable cov;
  int loss_type;
  int Kg, c_total, q_total;
  float* split_hi;  // [K][ld_split] or nullptr
  float* split_lo;
  long long ld_split;
  float* hsum_partial;    // [gridDim.x][K]
  float* q_partial;       // [gridDim.x][q_total]
  double* pred_partial;   // [gridDim.x][n_cov]
  double* t1_partial;     // [gridDim.x]
};
inline size_t h_update_smem_bytes(int K, int Kg, int c_total, int q_total) {
  const size_t q_pad = (static_cast<size_t>(q_total) + 3) & ~size_t(3);
  size_t f = static_cast<size_t>(2) * K * kUpdPitch;
  f += q_pad * 2 + 4 * static_cast<size_t>(c_total) * kUpdCols + ((Kg + 3) & ~3) + ((K + 3) & ~3);  // Bs, qacc, rn, rd, ys[2], dcol, hacc
  return f * sizeof(float) + (kUpdThreads / 32) * sizeof(double) + static_cast<size_t>(Kg + 1 + kMaxCov) * sizeof(int) + 32;
}

// FIT: the full update with statistics; otherwise the transform update H *= 2A / max(2 T H, eps) (main.py:705-709)
// SLOTS: how many partial-sum slots per operand the two-elements-in-flight path handles (2: three CTAs per SM; 4, for
// small shards whose super-tiles are cut into more pieces: two CTAs per SM, twice the registers)
template <bool FIT, int SLOTS = 2>
__global__ void __launch_bounds__(kUpdThreads, SLOTS == 2 ? 3 : 2) h_update_kernel(const HUpdParams p) {
  extern __shared__ __align__(16) uint8_t upd_smem[];
  __shared__ int ids_num[2][kMaxTileSlots], ids_z[2][kMaxTileSlots];
  ptx::pdl_enter();
  const int K = p.K;
  const int q_pad = (p.q_total + 3) & ~3;           // keeps the arrays behind 16-byte aligned
  float* tiles = reinterpret_cast<float*>(upd_smem);  // [2][K][68]: old H, overwritten with the new H
  float* Bs = tiles + 2 * K * kUpdPitch;            // [q_total]  all B_i, row-major [c][k] at q_off
  float* qacc = Bs + q_pad;                         // [q_total]  running Q partial of this CTA
  float* rn = qacc + q_pad;                         // [c_total][64]  per-cell numerator ratios (rho / Y)
  float* rd = rn + static_cast<size_t>(p.c_total) * kUpdCols;  // [c_total][64]  Frobenius: B H_i
  float* ys = rd + static_cast<size_t>(p.c_total) * kUpdCols;    // [2][c_total][64]  the tile's labels (with the H tile)
  float* dcol = ys + 2 * static_cast<size_t>(p.c_total) * kUpdCols;  // [Kg]  KL: lam * colsum(B) per guided row
  float* hacc = dcol + ((p.Kg + 3) & ~3);           // [K]  running row sums of the new H
  double* red = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(hacc + ((K + 3) & ~3)) + 7) & ~uintptr_t(7));  // [8]
  int* rowcov = reinterpret_cast<int*>(red + kUpdThreads / 32);  // [Kg]  covariate of guided row k
  int* rowc0 = rowcov + p.Kg + 1;                   // [n_cov]  first row of rn / rd of a covariate
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_tiles = (p.n + kUpdCols - 1) / kUpdCols;

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
    __syncthreads();
    // KL: denG = (lam B^T) 1 = lam * colsum(B), the same for every cell (main.py:644)
    for (int k = tid; k < p.Kg; k += kUpdThreads) {
      const CovDesc d = p.cov.d[rowcov[k]];
      float s = 0.f;
      for (int c = 0; c < d.c; ++c) s += (d.lam * Bs[d.q_off + c * d.k + (k - d.row0)]) * 1.0f;
      dcol[k] = s;
    }
  }
  double t1 = 0.0, pl0 = 0.0, pl1 = 0.0;  // pred loss of covariates (tid / 64) and (tid / 64 + 4)

  // the labels of a tile travel with its H tile (same cp.async group): no global-memory latency inside the two
  // per-cell phases; cells >= n arrive as zeros
  auto load_labels_async = [&](float* dst, long long col0) {
    if (!FIT) return;
    const uint32_t base = ptx::smem_u32(dst);
    int coff = 0;
    for (int i = 0; i < p.cov.n_cov; ++i) {
      const CovDesc d = p.cov.d[i];
      for (int e = tid; e < d.c * kUpdCols; e += kUpdThreads) {
        const int c = e >> 6, j = e & 63;
        const bool live = col0 + j < p.n;
        cp_async4(base + ((coff + c) * kUpdCols + j) * 4, live ? d.Y + static_cast<long long>(c) * p.n + col0 + j : d.Y,
                  live ? 4 : 0);
      }
      coff += d.c;
    }
  };
  long long tile_i = blockIdx.x;
  int buf = 0;
  if (tile_i < n_tiles) {
    load_labels_async(ys, tile_i * kUpdCols);
    load_tile_async(tiles, p.H, p.ldH, K, tile_i * kUpdCols, p.n);
  }
  for (; tile_i < n_tiles; tile_i += gridDim.x, buf ^= 1) {
    float* tile = tiles + buf * K * kUpdPitch;
    const float* ys_cur = ys + buf * static_cast<size_t>(p.c_total) * kUpdCols;
    const long long c0 = tile_i * kUpdCols;
    cp_async_wait_all();
    __syncthreads();  // the tile has landed (also orders the set-up above); everybody is done with the other buffer
    const int *idn[2], *idz[2];
    int nn[2], nz[2];
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
      idn[gi] = idz[gi] = ids_num[0];
      nn[gi] = (gi == 0 || p.num.split < K) ? stage_slot_ids(p.num.g[gi], c0, ids_num[gi], &idn[gi]) : 0;
      nz[gi] = (gi == 0 || p.z.split < K) ? stage_slot_ids(p.z.g[gi], c0, ids_z[gi], &idz[gi]) : 0;
    }
    const bool few_slots = nn[0] <= SLOTS && nn[1] <= SLOTS && nz[0] <= SLOTS && nz[1] <= SLOTS;
    if (!FIT) __syncthreads();  // (FIT: the barrier after the guided terms publishes the ids)
    const long long next = tile_i + gridDim.x;
    if (next < n_tiles) {
      load_labels_async(ys + (buf ^ 1) * static_cast<size_t>(p.c_total) * kUpdCols, next * kUpdCols);
      load_tile_async(tiles + (buf ^ 1) * K * kUpdPitch, p.H, p.ldH, K, next * kUpdCols, p.n);
    }
    if (FIT) {
      // guided terms of the OLD H with the NEW B, per cell (main.py:637-650): thread (i, j) = (tid / 64 [+4], tid % 64)
      const int j = tid & 63;
      for (int i = tid >> 6; i < p.cov.n_cov; i += 4) {
        const CovDesc d = p.cov.d[i];
        const int cbase = rowc0[i];
        for (int c = 0; c < d.c; ++c) {
          float yhat = 0.f;
          for (int k = 0; k < d.k; ++k) yhat += Bs[d.q_off + c * d.k + k] * tile[(d.row0 + k) * kUpdPitch + j];
          const float y = ys_cur[(cbase + c) * kUpdCols + j];
          rn[(cbase + c) * kUpdCols + j] = (p.loss_type == LOSS_KL) ? y / fmaxf(yhat, p.eps) : y;
          rd[(cbase + c) * kUpdCols + j] = yhat;
        }
      }
      __syncthreads();  // rn / rd are complete
    }
    auto update = [&](int e, float4 num4, float4 z4) {
      const int k = e >> 4, c4 = e & 15;
      const long long col = c0 + 4 * c4;
      float* cell = tile + k * kUpdPitch + 4 * c4;
      const float4 old4 = *reinterpret_cast<const float4*>(cell);
      const float oldv[4] = {old4.x, old4.y, old4.z, old4.w}, numv[4] = {num4.x, num4.y, num4.z, num4.w};
      const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
      float gn[4] = {0.f, 0.f, 0.f, 0.f}, gd[4] = {0.f, 0.f, 0.f, 0.f};
      if (FIT && k < p.Kg) {
        const CovDesc d = p.cov.d[rowcov[k]];
        const int cbase = rowc0[rowcov[k]];
        const float scale = (p.loss_type == LOSS_KL) ? d.lam : 2.0f * d.lam;
        for (int c = 0; c < d.c; ++c) {
          const float lb = scale * Bs[d.q_off + c * d.k + (k - d.row0)];
          const float4 r = *reinterpret_cast<const float4*>(rn + (cbase + c) * kUpdCols + 4 * c4);
          gn[0] += lb * r.x, gn[1] += lb * r.y, gn[2] += lb * r.z, gn[3] += lb * r.w;
          if (p.loss_type != LOSS_KL) {
            const float4 q = *reinterpret_cast<const float4*>(rd + (cbase + c) * kUpdCols + 4 * c4);
            gd[0] += lb * q.x, gd[1] += lb * q.y, gd[2] += lb * q.z, gd[3] += lb * q.w;
          }
        }
        if (p.loss_type == LOSS_KL) gd[0] = gd[1] = gd[2] = gd[3] = dcol[k];
      }
      float outv[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float num = gn[x] + 2.0f * numv[x];               // main.py:648, 653 / 706
        const float den = fmaxf(gd[x] + 2.0f * zv[x], p.eps);   // main.py:649, 654-655 / 707-708
        outv[x] = (col + x < p.n) ? oldv[x] * (num / den) : 0.f;  // main.py:656 / 709
        if (FIT) t1 += static_cast<double>(numv[x]) * static_cast<double>(outv[x]);
      }
      const float4 out4 = make_float4(outv[0], outv[1], outv[2], outv[3]);
      if (FIT) *reinterpret_cast<float4*>(cell) = out4;  // the statistics below read the new tile
      store4(out4, k, col, p.n, p.H, p.ldH, p.split_hi, p.split_lo, p.ld_split, 0, nullptr);
    };
    const int n_items = K * 16;
    if (few_slots) {
      // two elements per thread and round: their (up to eight) loads are in flight together
      for (int e = tid; e < n_items; e += 2 * kUpdThreads) {
        const int e2 = e + kUpdThreads;
        const bool v1 = c0 + 4 * (e & 15) < p.n, v2 = e2 < n_items && c0 + 4 * (e2 & 15) < p.n;
        SlotSet<SLOTS> n1{}, z1{}, n2{}, z2{};
        if (v1) {
          const int k = e >> 4, gn = k >= p.num.split, gz = k >= p.z.split;
          n1 = slot_load_set<SLOTS>(p.num.g[gn], k - gn * p.num.split, c0 + 4 * (e & 15), idn[gn], nn[gn]);
          z1 = slot_load_set<SLOTS>(p.z.g[gz], k - gz * p.z.split, c0 + 4 * (e & 15), idz[gz], nz[gz]);
        }
        if (v2) {
          const int k = e2 >> 4, gn = k >= p.num.split, gz = k >= p.z.split;
          n2 = slot_load_set<SLOTS>(p.num.g[gn], k - gn * p.num.split, c0 + 4 * (e2 & 15), idn[gn], nn[gn]);
          z2 = slot_load_set<SLOTS>(p.z.g[gz], k - gz * p.z.split, c0 + 4 * (e2 & 15), idz[gz], nz[gz]);
        }
        if (v1) update(e, set_sum<SLOTS>(n1), set_sum<SLOTS>(z1));
        if (v2) update(e2, set_sum<SLOTS>(n2), set_sum<SLOTS>(z2));
      }
    } else {
      for (int e = tid; e < n_items; e += kUpdThreads) {
        if (c0 + 4 * (e & 15) >= p.n) continue;
        const int k = e >> 4, gn = k >= p.num.split, gz = k >= p.z.split;
        update(e, slot_load4(p.num.g[gn], k - gn * p.num.split, c0 + 4 * (e & 15), idn[gn], nn[gn]),
               slot_load4(p.z.g[gz], k - gz * p.z.split, c0 + 4 * (e & 15), idz[gz], nz[gz]));
      }
    }
    if (FIT) {
      __syncthreads();  // the tile holds the new H
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
          const float y = ys_cur[(cbase + c) * kUpdCols + j];
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
      // Q_i += rho' H_i^T over this tile's cells; row sums of the new H.  One (entry, quarter) per thread, the four
      // quarters of an entry are combined in a fixed order by the owner lane.
      const int n_items = 4 * (p.q_total + K);
      for (int e0 = 0; e0 < n_items; e0 += kUpdThreads) {  // (warp-uniform trip count: the shuffles below are full-warp)
        const int e = e0 + tid;
        const int ent = e >> 2, quarter = e & 3;
        float a = 0.f;
        if (e < n_items) {
          if (ent < p.q_total) {
            int i = 0;
            while (i + 1 < p.cov.n_cov && ent >= p.cov.d[i + 1].q_off) ++i;
            const CovDesc d = p.cov.d[i];
            const int c = (ent - d.q_off) / d.k, k = (ent - d.q_off) - c * d.k;
            const float* rrow = rn + (rowc0[i] + c) * kUpdCols + 16 * quarter;
            const float* hrow = tile + (d.row0 + k) * kUpdPitch + 16 * quarter;
#pragma unroll
            for (int u = 0; u < 16; ++u) a += rrow[u] * hrow[u];
          } else {
            const float* hrow = tile + (ent - p.q_total) * kUpdPitch + 16 * quarter;
#pragma unroll
            for (int u = 0; u < 16; ++u) a += hrow[u];
          }
        }
        // the four quarters of an entry sit in adjacent lanes (e is a multiple of 4 at quarter 0)
        const float a1 = __shfl_down_sync(0xffffffffu, a, 1);
        const float a2 = __shfl_down_sync(0xffffffffu, a, 2);
        const float a3 = __shfl_down_sync(0xffffffffu, a, 3);
        if (e < n_items && quarter == 0) {
          const float s = (a + a1) + (a2 + a3);
          if (ent < p.q_total) qacc[ent] += s; else hacc[ent - p.q_total] += s;
        }
      }
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
    for (int w = 0; w < kUpdThreads / 32; ++w) s += red[w];
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
// Finish kernels
// K x K Gram matrix from the slots of its contraction (one super-tile: rows = components): out[k][r] = sum over the
// slots, in list order.  32 lanes share an output element's slots (strided) and are combined by a fixed shuffle
// tree, so that the many small loads overlap.  Also writes the split copies (B operand of the next Z plan).
struct GramFromSlots {
  SlotSrc2 src;      // slots [K][256] of the Gram plan (rows r < K are meaningful), per component group
  int tile;          // super-tile of the plan that holds the Gram (0 for a Gram plan; the additional tile of W^T X)
  int K;
  float* out;        // [K][ld]
  int ld;
  float* split_hi;   // [K][ld_split] or nullptr
  float* split_lo;
  int ld_split;
};
// one block per (component k, group of 32 rows r): lanes = 32 consecutive r (128-byte coalesced reads of a slot row),
// the 8 warps take the slots round-robin and are combined in a fixed order
inline int gram_blocks_for(int K) { return K * ((K + 31) / 32); }
__device__ __forceinline__ void gram_from_slots_block(const GramFromSlots& g, int b, float (*red)[32]) {
  const int RG = (g.K + 31) >> 5;
  const int k = b / RG, rg = b - k * RG;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = rg * 32 + lane;
  const int gi = k >= g.src.split;
  const SlotSrc& src = g.src.g[gi];
  const int kl = k - gi * g.src.split;
  const int s0 = __ldg(src.slot_ofs + g.tile), s1 = __ldg(src.slot_ofs + g.tile + 1);
  float a = 0.f;
  if (r < g.K)
    for (int q = s0 + w; q < s1; q += 8)
      a += __ldcg(src.partial + (static_cast<size_t>(__ldg(src.slots + q)) * src.K + kl) * 256 + r);
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && r < g.K) {
    float t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    g.out[k * g.ld + r] = t;
    if (g.split_hi != nullptr) {
      ptx::store_split1(t, g.split_hi, g.split_lo, k, r, g.ld_split);
    }
  }
}

struct WFinishParams {
  GramFromSlots gram;   // -> T = W^T W
  int gram_blocks;      // 0: T comes from elsewhere
  CovTable cov;         // B updates (main.py:615-628)
  int loss_type;
  const float* stats_q;
  const float* hsum;
  const float* S;
  int ldS;
  float eps;
};
__global__ void __launch_bounds__(256) w_finish_kernel(const WFinishParams p) {
  __shared__ float gred[8][32];
  ptx::pdl_enter();
  if (static_cast<int>(blockIdx.x) < p.gram_blocks) {
    gram_from_slots_block(p.gram, blockIdx.x, gred);
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
  GramFromSlots gram;   // -> S = H H^T of this shard
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
// fp64 column sums of 32 adjacent columns of a [n_parts][L] array of per-CTA partials: lanes = columns (coalesced), the
// 8 warps stride over the parts and are combined in a fixed order
__device__ __forceinline__ void colsum32_block(const float* __restrict__ part, int n_parts, int L, int c0,
                                               float* __restrict__ out, double (*red)[32]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = c0 + lane;
  double a = 0.0;
  if (col < L)
    for (int q = w; q < n_parts; q += 8) a += static_cast<double>(__ldcg(part + static_cast<size_t>(q) * L + col));
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && col < L) {
    double t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    out[col] = static_cast<float>(t);
  }
}
inline int h_finish_stat_blocks(int K, int q_total) { return (K + 31) / 32 + (q_total + 31) / 32 + 1; }
// blocks [0, gram_blocks): Gram; then ceil(K / 32) blocks of rowsum(H), ceil(q_total / 32) blocks of Q, one block of
// scalars (t1, pred_i).  The block that finishes last (all of S is then in memory) takes t2 = sum T .* S in a fixed
// order.
__global__ void __launch_bounds__(256) h_finish_kernel(const HFinishParams p) {
  __shared__ double red[256];
  __shared__ float gred[8][32];
  __shared__ unsigned int last;
  ptx::pdl_enter();
  const int K = p.gram.K;
  double (*red2)[32] = reinterpret_cast<double (*)[32]>(red);
  const int hb = (K + 31) / 32, qb = (p.q_total + 31) / 32;
  const int b2 = static_cast<int>(blockIdx.x) - p.gram_blocks;
  if (b2 < 0) {
    gram_from_slots_block(p.gram, blockIdx.x, gred);
  } else if (b2 < hb) {
    colsum32_block(p.hsum_partial, p.n_parts, K, 32 * b2, p.hsum, red2);
  } else if (b2 < hb + qb) {
    colsum32_block(p.q_partial, p.n_parts, p.q_total, 32 * (b2 - hb), p.stats_q, red2);
  } else if (p.loss_row != nullptr) {
    for (int which = 0; which < 1 + p.n_cov; ++which) {  // t1, pred_0 ..
      double a = 0.0;
      for (int q = threadIdx.x; q < p.n_parts; q += 256)
        a += which == 0 ? p.t1_partial[q] : p.pred_partial[static_cast<size_t>(q) * p.n_cov + (which - 1)];
      const double s = block_sum_256(a, red);
      if (threadIdx.x == 0) p.loss_row[which == 0 ? 0 : 1 + which] = s;
    }
  }
  if (p.loss_row == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(p.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last == 0u) return;
  __threadfence();
  // (T and S are both K x K with pitch K here: walk them linearly, four independent loads in flight per thread)
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const int KK = K * K;
  if (p.ldT == K && p.gram.ld == K) {
    int e = threadIdx.x;
    for (; e + 768 < KK; e += 1024) {
      const float t0 = __ldcg(p.T + e), t1 = __ldcg(p.T + e + 256), t2 = __ldcg(p.T + e + 512), t3 = __ldcg(p.T + e + 768);
      const float s0 = __ldcg(p.gram.out + e), s1 = __ldcg(p.gram.out + e + 256), s2 = __ldcg(p.gram.out + e + 512),
                  s3 = __ldcg(p.gram.out + e + 768);
      a0 += static_cast<double>(t0) * s0, a1 += static_cast<double>(t1) * s1;
      a2 += static_cast<double>(t2) * s2, a3 += static_cast<double>(t3) * s3;
    }
    for (; e < KK; e += 256) a0 += static_cast<double>(__ldcg(p.T + e)) * __ldcg(p.gram.out + e);
  } else {
    for (int e = threadIdx.x; e < KK; e += 256) {
      const int r = e / K, c = e - r * K;
      a0 += static_cast<double>(__ldcg(p.T + r * p.ldT + c)) * static_cast<double>(__ldcg(p.gram.out + r * p.gram.ld + c));
    }
  }
  const double s = block_sum_256((a0 + a1) + (a2 + a3), red);
  if (threadIdx.x == 0) {
    p.loss_row[1] = s;
    *p.counter = 0u;
  }
}

}  // namespace alpine
