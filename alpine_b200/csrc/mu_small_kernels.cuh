// Element-wise and reduction kernels around the tcgen05 contractions: operand splitting, the W / B
// multiplicative updates, guided (covariate) terms of the H update, per-iteration statistics and loss terms,
// post-fit scaling and the transform update.  All are HBM-bound streaming kernels: coalesced along the
// contiguous dimension, fp64 only in the loss accumulators, partial sums reduced in a fixed order.
#pragma once
#include "ptx_sm100.cuh"

namespace alpine {

enum { LOSS_KL = 0, LOSS_FROB = 1 };
constexpr int kMaxCov = 8;
constexpr int kMaxPeers = 8;

// loads / stores of memory that another GPU writes / reads inside the same fit: never served from a stale L1 line
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_sys_f32(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

struct CovDesc {
  int row0;         // first row of the block in H / first column in W
  int k;            // components of the block
  int c;            // categories
  const float* Y;   // [c][n] one-hot (zero column = missing label)
  float* B;         // [c][k]
  int q_off;        // offset (floats) of this block's Q / YH matrix inside the stats area
  float lam;
};
struct CovTable {
  int n_cov;
  CovDesc d[kMaxCov];
};

// ---------------------------------------------------------------------------------------------------------
// sum of squares of a pitched matrix in fp64 (||X||_F^2), two-stage deterministic reduction
// also raises *inexact when some value is not exactly representable in tf32 (low 13 mantissa bits set): only
// then does the contraction kernel need the lo half of the split of X
__global__ void sumsq_partial_kernel(const float* __restrict__ X, long long ld, long long rows, int cols,
                                     double* __restrict__ partial, int* __restrict__ inexact) {
  double acc = 0.0;
  unsigned int low = 0u;
  const int cols4 = cols >> 2;  // ld % 4 == 0 and base 16B aligned => float4 loads are aligned per row
  for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
    const float4* p4 = reinterpret_cast<const float4*>(X + r * ld);
    float s = 0.f;
    for (int c = threadIdx.x; c < cols4; c += blockDim.x) {
      const float4 v = __ldg(p4 + c);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      low |= (__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w)) & 0x1FFFu;
    }
    for (int c = (cols4 << 2) + threadIdx.x; c < cols; c += blockDim.x) {
      const float v = X[r * ld + c];
      s += v * v;
      low |= __float_as_uint(v) & 0x1FFFu;
    }
    acc += static_cast<double>(s);
  }
  if (__any_sync(0xffffffffu, low != 0u) && (threadIdx.x & 31) == 0) *inexact = 1;
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void sum_double_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += partial[i];
    *out = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// B update (reference main.py:615-628) from cell-reduced statistics; one block per covariate.
//   KL  : B *= (lam * Q) / max(lam * hsum, eps),  Q = (Y / max(B H_i, eps)) H_i^T
//   Frob: B *= (2 * YH) / max(2 * B (H_i H_i^T), eps)
__global__ void b_update_kernel(const CovTable tab, int loss_type, const float* __restrict__ stats_q,
                                const float* __restrict__ hsum, const float* __restrict__ S, int ldS, float eps) {
  const CovDesc d = tab.d[blockIdx.x];
  extern __shared__ float bs[];  // old B copy [c][k]
  for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) bs[e] = d.B[e];
  __syncthreads();
  const float* Q = stats_q + d.q_off;
  for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) {
    const int c = e / d.k, k = e - c * d.k;
    float num, den;
    if (loss_type == LOSS_KL) {
      num = d.lam * Q[e];
      den = d.lam * hsum[d.row0 + k];
    } else {
      num = 2.0f * Q[e];
      float acc = 0.f;
      for (int k2 = 0; k2 < d.k; ++k2) acc += (2.0f * bs[c * d.k + k2]) * S[(d.row0 + k2) * ldS + d.row0 + k];
      den = acc;
    }
    den = fmaxf(den, eps);
    d.B[e] = bs[e] * (num / den);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Guided terms of the H update for the covariate rows (reference main.py:637-650), per cell, old H, new B:
//   KL  : numG = (lam B^T) (Y / max(B H_i, eps)) ;  denG = (lam B^T) 1
//   Frob: numG = (2 lam B^T) Y                   ;  denG = (2 lam B^T) (B H_i)
// One thread per cell; B in shared memory; per-thread row accumulators in shared memory [k][tid].
__global__ void __launch_bounds__(128) guided_terms_kernel(const CovTable tab, int loss_type,
                                                           const float* __restrict__ H, long long ldH, int n,
                                                           float eps, float* __restrict__ numG,
                                                           float* __restrict__ denG, long long ldD) {
  extern __shared__ float sm[];
  const CovDesc d = tab.d[blockIdx.y];
  float* Bs = sm;                    // [c][k]
  float* acc_n = sm + d.c * d.k;     // [k][128]
  float* acc_d = acc_n + d.k * 128;  // [k][128]
  float* hcol = acc_d + d.k * 128;   // [k][128]
  for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) Bs[e] = d.B[e];
  __syncthreads();
  const long long j = blockIdx.x * 128ll + threadIdx.x;
  if (j >= n) return;
  const int t = threadIdx.x;
  for (int k = 0; k < d.k; ++k) {
    hcol[k * 128 + t] = H[(d.row0 + k) * ldH + j];
    acc_n[k * 128 + t] = 0.f;
    acc_d[k * 128 + t] = 0.f;
  }
  const float scale = (loss_type == LOSS_KL) ? d.lam : 2.0f * d.lam;
  for (int c = 0; c < d.c; ++c) {
    float yhat = 0.f;
    for (int k = 0; k < d.k; ++k) yhat += Bs[c * d.k + k] * hcol[k * 128 + t];
    const float y = __ldg(d.Y + static_cast<long long>(c) * n + j);
    const float rn = (loss_type == LOSS_KL) ? y / fmaxf(yhat, eps) : y;
    const float rd = (loss_type == LOSS_KL) ? 1.0f : yhat;
    for (int k = 0; k < d.k; ++k) {
      const float lb = scale * Bs[c * d.k + k];
      acc_n[k * 128 + t] += lb * rn;
      acc_d[k * 128 + t] += lb * rd;
    }
  }
  for (int k = 0; k < d.k; ++k) {
    numG[(d.row0 + k) * ldD + j] = acc_n[k * 128 + t];
    denG[(d.row0 + k) * ldD + j] = acc_d[k * 128 + t];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Per-covariate cell statistics from the NEW H and NEW B (inputs of the next B update and the prediction loss,
// reference main.py:618-626 and 727-748): per block of 256 cells
//   KL  : Qp[c][k] = sum_j rho[c][j] H_i[k][j], rho = Y / max(B H_i, eps);  pred = sum y log(max(y/yh,eps)) - y + yh
//   Frob: Qp[c][k] = sum_j Y[c][j] H_i[k][j];                               pred = sum (y - B H_i)^2
constexpr int kStatCells = 256;
__global__ void __launch_bounds__(kStatCells) cov_stats_kernel(const CovTable tab, int loss_type,
                                                               const float* __restrict__ H, long long ldH, int n,
                                                               float eps, int q_total,
                                                               float* __restrict__ q_partial,     // [blocks][q_total]
                                                               double* __restrict__ pred_partial  // [blocks][n_cov]
) {
  extern __shared__ float sm[];
  const CovDesc d = tab.d[blockIdx.y];
  constexpr int P = kStatCells + 1;
  float* Bs = sm;                // [c][k]
  float* hs = sm + d.c * d.k;    // [k][P]
  float* rs = hs + d.k * P;      // [c][P]
  __shared__ double red[kStatCells / 32];
  for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) Bs[e] = d.B[e];
  const int t = threadIdx.x;
  const long long j = blockIdx.x * static_cast<long long>(kStatCells) + t;
  const bool live = j < n;
  for (int k = 0; k < d.k; ++k) hs[k * P + t] = live ? H[(d.row0 + k) * ldH + j] : 0.f;
  __syncthreads();
  double pl = 0.0;
  for (int c = 0; c < d.c; ++c) {
    float yhat = 0.f;
    for (int k = 0; k < d.k; ++k) yhat += Bs[c * d.k + k] * hs[k * P + t];
    const float y = live ? __ldg(d.Y + static_cast<long long>(c) * n + j) : 0.f;
    float r;
    if (loss_type == LOSS_KL) {
      const float yh = fmaxf(yhat, eps);
      r = y / yh;
      if (live) pl += static_cast<double>(y * logf(fmaxf(y / yh, eps)) - y + yh);
    } else {
      r = y;
      const float dlt = y - yhat;
      if (live) pl += static_cast<double>(dlt * dlt);
    }
    rs[c * P + t] = live ? r : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) {
    const int c = e / d.k, k = e - c * d.k;
    float acc = 0.f;
    for (int u = 0; u < kStatCells; ++u) acc += rs[c * P + u] * hs[k * P + u];
    q_partial[static_cast<size_t>(blockIdx.x) * q_total + d.q_off + e] = acc;
  }
  for (int o = 16; o > 0; o >>= 1) pl += __shfl_down_sync(0xffffffffu, pl, o);
  if ((t & 31) == 0) red[t >> 5] = pl;
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
    for (int w = 0; w < kStatCells / 32; ++w) s += red[w];
    pred_partial[static_cast<size_t>(blockIdx.x) * tab.n_cov + blockIdx.y] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Post-fit scaling (reference main.py:772-781): column sums of W, then W /= s, H *= s, B /= s.
__global__ void scale_w_kernel(float* __restrict__ W, long long ldW, int G, int K, const float* __restrict__ s) {
  const long long total = static_cast<long long>(G) * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long g = i / K;
    const int k = static_cast<int>(i - g * K);
    W[g * ldW + k] = W[g * ldW + k] / s[k];
  }
}
__global__ void scale_h_kernel(float* __restrict__ H, long long ldH, int K, int n, const float* __restrict__ s) {
  const long long total = static_cast<long long>(K) * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / n);
    const long long j = i - static_cast<long long>(k) * n;
    H[k * ldH + j] = H[k * ldH + j] * s[k];
  }
}
__global__ void scale_b_kernel(const CovTable tab, const float* __restrict__ s) {
  const CovDesc d = tab.d[blockIdx.x];
  for (int e = threadIdx.x; e < d.c * d.k; e += blockDim.x) d.B[e] = d.B[e] / s[d.row0 + (e % d.k)];
}

// ---------------------------------------------------------------------------------------------------------
// Operand split of the small (K x R) operand of a contraction: plane `hi` = tf32 hi values, plane `lo` = the bf16
// images of hi and of the remainder (ptx::store_split4).  Pitches are multiples of 8 floats and bases 16-byte
// aligned, so whole groups of four are always in bounds.
__global__ void split_operand_kernel(const float* __restrict__ src, long long ld_src, int rows, long long cols,
                                     float* __restrict__ hi, float* __restrict__ lo, long long ld_dst) {
  ptx::pdl_enter();
  const long long c4n = (cols + 3) >> 2;
  const long long total = static_cast<long long>(rows) * c4n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c4n, c = (i - r * c4n) << 2;
    ptx::store_split4(*reinterpret_cast<const float4*>(src + r * ld_src + c), hi, lo, r, c, ld_dst);
  }
}

// the same for a small K x K matrix of any pitch / alignment (H H^T, W^T W): scalar accesses
__global__ void split_small_kernel(const float* __restrict__ src, int ld_src, int K, float* __restrict__ hi,
                                   float* __restrict__ lo, int ld_dst) {
  ptx::pdl_enter();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * K) return;
  const int r = e / K, c = e - r * K;
  ptx::store_split1(src[r * ld_src + c], hi, lo, r, c, ld_dst);
}

// ---------------------------------------------------------------------------------------------------------
// out[c][r] = in[r][c]  (W <-> W^T), 32x32 tiles through shared memory, both sides coalesced
__global__ void transpose_kernel(const float* __restrict__ in, long long ld_in, int rows, int cols,
                                 float* __restrict__ out, long long ld_out) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int rr = r0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (rr < rows && cc < cols) ? in[static_cast<long long>(rr) * ld_in + cc] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int cc = c0 + r, rr = r0 + threadIdx.x;
    if (cc < cols && rr < rows) out[static_cast<long long>(cc) * ld_out + rr] = tile[threadIdx.x][r];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Z = Sym (K x K, symmetric) * Mat (K x L, L long), fused with the multiplicative update that consumes Z.
// This is the Gram reformulation of the reference's G x n sized products:
//   EPI_W : Mat = W^T [K][G], Sym = S = H H^T :  Z^T = W (H H^T)  replaces ((2W) @ H) @ H^T   (main.py:599)
//           den = 2 Z + (1-l1) alpha W + orth (rowsum_k(W) - W) + l1 alpha ; W *= 2 P / max(den, eps)   (main.py:596-605)
//   EPI_H : Mat = H [K][n],   Sym = T = W^T W :  Z = (W^T W) H    replaces (2W^T) @ (W @ H)      (main.py:654)
//           H *= (numG + 2 A) / max(denG + 2 Z, eps)                                            (main.py:648-656)
//   EPI_TRANSFORM: H *= 2 A / max(2 Z, eps)                                                     (main.py:706-709)
// One block = all K rows x 64 columns; 16 x 16 threads, each KI rows (ty + 16 i) x 4 consecutive columns.
enum { EPI_W = 0, EPI_H = 1, EPI_TRANSFORM = 2 };
constexpr int kSLCols = 64;
struct SymLongParams {
  const float* Sym;
  int ldS;
  float* Mat;  // updated in place
  long long ldM;
  int K;
  long long L;
  const float* Num;  // EPI_W: P^T [K][ldNum];  EPI_H / EPI_TRANSFORM: A = W^T X [K][ldNum]
  long long ldNum;
  const float* numG;  // EPI_H: guided rows [Kg][ldD]
  const float* denG;
  long long ldD;
  int Kg;
  float c1, c2, orth, eps;
  double* t1_partial;   // EPI_H: [gridDim.x]  sum A .* Hnew
  float* rowsum_partial;  // EPI_H: [gridDim.x][K] row sums of the new H over this block's columns, or nullptr
  float* split_hi;      // EPI_W / EPI_H: split copies (tf32 hi plane, bf16 plane) of the updated matrix (B operand of the next
  float* split_lo;      //                contraction), pitch ld_split; nullptr to skip
  long long ld_split;
  // Peer mode (EPI_W under cell sharding, csrc/peer_exchange.cuh): the block handles columns col0 + 64 * blockIdx.x
  // (this rank's gene slice) and the updated columns are also stored into every peer's W^T over NVLink.
  long long col0;
  int n_peers;                    // 0: single-GPU / NCCL path
  float* mat_peer[kMaxPeers];     // peers' W^T (nullptr for this rank itself)
  // rows [r0, r1) of Mat are updated (all K rows take part in Z).  The whole matrix for the simultaneous update
  // (main.py:589-663); one component block for the block Gauss-Seidel ("ALS") sweep (main.py:523-588), where the
  // orthogonality term also only couples the columns of that block (main.py:541).
  int r0, r1;
};
constexpr int kSLKI = 8;  // rows per thread: K <= 128 (the contraction kernel's limit as well)
inline int sym_long_pitch(int K) { return (K + 3) / 4 * 4 + 4; }  // Ss row pitch (floats)
inline size_t sym_long_smem_bytes(int K) {
  return (static_cast<size_t>(K) * kSLCols + static_cast<size_t>(K) * sym_long_pitch(K)) * sizeof(float) +
         kSLCols * sizeof(double);
}
template <int KI, int EPI>
__global__ void __launch_bounds__(256) sym_long_kernel(const SymLongParams p) {
  extern __shared__ __align__(16) uint8_t sl_smem[];
  double* cs = reinterpret_cast<double*>(sl_smem);                // [64] column sums (EPI_W)
  float* Ms = reinterpret_cast<float*>(sl_smem + kSLCols * 8);    // [K][64]   tile of Mat
  float* Ss = Ms + static_cast<size_t>(p.K) * kSLCols;            // [K][SP]   Ss[k'][k] = Sym[k'][k] (symmetric)
  const int SP = (p.K + 3) / 4 * 4 + 4;
  __shared__ double red[8];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long c0 = p.col0 + static_cast<long long>(blockIdx.x) * kSLCols;
  const bool full = c0 + kSLCols <= p.L;

  // the whole K x K matrix and the K x 64 tile are staged once: one barrier, no per-chunk global latency
  for (int e = tid; e < p.K * p.K; e += 256) {
    const int kk = e / p.K, r = e - kk * p.K;
    Ss[kk * SP + r] = __ldg(p.Sym + static_cast<long long>(kk) * p.ldS + r);
  }
  for (int e = tid; e < p.K * 16; e += 256) {
    const int k = e >> 4, c4 = e & 15;
    const long long col = c0 + 4 * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = p.Mat + static_cast<long long>(k) * p.ldM + col;
    if (full) {
      v = *reinterpret_cast<const float4*>(src);
    } else {
      if (col + 0 < p.L) v.x = src[0];
      if (col + 1 < p.L) v.y = src[1];
      if (col + 2 < p.L) v.z = src[2];
      if (col + 3 < p.L) v.w = src[3];
    }
    reinterpret_cast<float4*>(Ms)[e] = v;
  }
  float4 acc[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // rows this thread owns; rows >= K are clamped to a valid address and their results are never used
  int rsel[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) rsel[i] = min(ty + 16 * i, p.K - 1);
#pragma unroll 2
  for (int kk = 0; kk < p.K; ++kk) {
    const float4 m = reinterpret_cast<const float4*>(Ms)[kk * 16 + tx];
    const float* srow = Ss + kk * SP;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const float sv = srow[rsel[i]];
      acc[i].x = fmaf(sv, m.x, acc[i].x);
      acc[i].y = fmaf(sv, m.y, acc[i].y);
      acc[i].z = fmaf(sv, m.z, acc[i].z);
      acc[i].w = fmaf(sv, m.w, acc[i].w);
    }
  }
  if (EPI == EPI_W) {
    // rowsum_k W[g][:] in fp64 so that (rowsum - w) does not cancel (the reference sums the other K-1 entries)
    if (tid < kSLCols) {
      double sacc = 0.0;
      for (int k = p.r0; k < p.r1; ++k) sacc += static_cast<double>(Ms[k * kSLCols + tid]);
      cs[tid] = sacc;
    }
    __syncthreads();
  }
  double t1 = 0.0;
  float rs[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) rs[i] = 0.f;
  const long long col = c0 + 4 * tx;
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const int k = ty + 16 * i;
    if (k < p.r0 || k >= p.r1 || col >= p.L) continue;
    const float4 old4 = reinterpret_cast<const float4*>(Ms)[k * 16 + tx];
    const float oldv[4] = {old4.x, old4.y, old4.z, old4.w};
    const float zv[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
    float nv[4] = {0.f, 0.f, 0.f, 0.f}, gn[4] = {0.f, 0.f, 0.f, 0.f}, gd[4] = {0.f, 0.f, 0.f, 0.f};
    const float* nsrc = p.Num + static_cast<long long>(k) * p.ldNum + col;
    const bool g_on = (EPI == EPI_H) && (k < p.Kg);
    if (full) {
      const float4 t = *reinterpret_cast<const float4*>(nsrc);
      nv[0] = t.x, nv[1] = t.y, nv[2] = t.z, nv[3] = t.w;
      if (g_on) {
        const float4 a = *reinterpret_cast<const float4*>(p.numG + static_cast<long long>(k) * p.ldD + col);
        const float4 b = *reinterpret_cast<const float4*>(p.denG + static_cast<long long>(k) * p.ldD + col);
        gn[0] = a.x, gn[1] = a.y, gn[2] = a.z, gn[3] = a.w;
        gd[0] = b.x, gd[1] = b.y, gd[2] = b.z, gd[3] = b.w;
      }
    } else {
      for (int x = 0; x < 4; ++x)
        if (col + x < p.L) {
          nv[x] = nsrc[x];
          if (g_on) {
            gn[x] = p.numG[static_cast<long long>(k) * p.ldD + col + x];
            gd[x] = p.denG[static_cast<long long>(k) * p.ldD + col + x];
          }
        }
    }
    float outv[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      float num, den;
      if (EPI == EPI_W) {
        const float others = static_cast<float>(cs[4 * tx + x] - static_cast<double>(oldv[x]));
        den = (2.0f * zv[x] + p.c1 * oldv[x]) + p.orth * others;  // main.py:599-601
        den += p.c2;                                              // main.py:603
        num = 2.0f * nv[x];                                       // main.py:596
      } else if (EPI == EPI_H) {
        num = gn[x] + 2.0f * nv[x];   // main.py:648, 653
        den = gd[x] + 2.0f * zv[x];   // main.py:649, 654
      } else {
        num = 2.0f * nv[x];  // main.py:706
        den = 2.0f * zv[x];  // main.py:707
      }
      den = fmaxf(den, p.eps);             // main.py:604, 655, 708
      outv[x] = oldv[x] * (num / den);     // main.py:605, 656, 709
      if (EPI == EPI_H && col + x < p.L) t1 += static_cast<double>(nv[x]) * static_cast<double>(outv[x]);
    }
    float* dst = p.Mat + static_cast<long long>(k) * p.ldM + col;
    if (full) {
      *reinterpret_cast<float4*>(dst) = make_float4(outv[0], outv[1], outv[2], outv[3]);
    } else {
      for (int x = 0; x < 4; ++x)
        if (col + x < p.L) dst[x] = outv[x];
    }
    if (EPI == EPI_W && p.n_peers > 0) {
      for (int q = 0; q < p.n_peers; ++q) {
        if (p.mat_peer[q] == nullptr) continue;
        float* pd = p.mat_peer[q] + static_cast<long long>(k) * p.ldM + col;
        if (full) {
          *reinterpret_cast<float4*>(pd) = make_float4(outv[0], outv[1], outv[2], outv[3]);
        } else {
          for (int x = 0; x < 4; ++x)
            if (col + x < p.L) pd[x] = outv[x];
        }
      }
    }
    if (EPI != EPI_TRANSFORM && p.split_hi != nullptr) {
      // pitches are multiples of 4 and the pad columns are never read unmasked, so whole float4 groups are written
      ptx::store_split4(make_float4((col + 0 < p.L) ? outv[0] : 0.f, (col + 1 < p.L) ? outv[1] : 0.f,
                                    (col + 2 < p.L) ? outv[2] : 0.f, (col + 3 < p.L) ? outv[3] : 0.f),
                        p.split_hi, p.split_lo, k, col, p.ld_split);
    }
    if (EPI == EPI_H && p.rowsum_partial != nullptr) {
      float v = 0.f;
#pragma unroll
      for (int x = 0; x < 4; ++x)
        if (col + x < p.L) v += outv[x];
      rs[i] = v;
    }
  }
  if (EPI == EPI_H && p.rowsum_partial != nullptr) {
    // fixed-order reduction over the 16 threads (tx) that share a row: xor-shuffles inside each half warp
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      float v = rs[i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const int k = ty + 16 * i;
      if (tx == 0 && k >= p.r0 && k < p.r1) p.rowsum_partial[static_cast<size_t>(blockIdx.x) * p.K + k] = v;
    }
  }
  if (EPI == EPI_H && p.t1_partial != nullptr) {
    for (int o = 16; o > 0; o >>= 1) t1 += __shfl_down_sync(0xffffffffu, t1, o);
    if ((tid & 31) == 0) red[tid >> 5] = t1;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += red[w];
      p.t1_partial[blockIdx.x] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Finish the per-iteration statistics on one block (fixed summation order):
//   stats_q  <- sum over cell blocks of q_partial                (input of the next B update)
//   loss_row <- [ t1 = sum A .* Hnew, t2 = sum T .* S, pred_0.. ] (fp64; recon = ||X||^2 - 2 t1 + t2)
struct StatsFinishParams {
  const float* q_partial;
  int q_blocks, q_total;
  float* stats_q;
  const double* pred_partial;
  int n_cov;
  const double* t1_partial;
  int t1_n;
  const float* T;
  int ldT;
  const float* S;
  int ldS;
  int K;
  const float* rowsum_partial;  // [rs_blocks][K] from the H update, or nullptr (row sums computed elsewhere)
  int rs_blocks;
  float* hsum;                  // [K]
  double* loss_row;  // [2 + n_cov] or nullptr
};
// fixed-order block reduction of a double (256 threads): thread-strided partial sums, then a shared-memory tree
__device__ __forceinline__ double block_sum_256(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}
// grid = K + q_total + 2 + n_cov blocks of 256 threads: K row sums of H, q_total Q entries, then t1, t2, pred_i
__global__ void __launch_bounds__(256) stats_finish_kernel(const StatsFinishParams p) {
  __shared__ double red[256];
  int b = blockIdx.x;
  if (b < p.K) {
    if (p.rowsum_partial == nullptr) return;
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.rs_blocks; i += 256) acc += static_cast<double>(p.rowsum_partial[static_cast<size_t>(i) * p.K + b]);
    const double r = block_sum_256(acc, red);
    if (threadIdx.x == 0) p.hsum[b] = static_cast<float>(r);
    return;
  }
  b -= p.K;
  if (b < p.q_total) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < p.q_blocks; i += 256) acc += static_cast<double>(p.q_partial[static_cast<size_t>(i) * p.q_total + b]);
    const double r = block_sum_256(acc, red);
    if (threadIdx.x == 0) p.stats_q[b] = static_cast<float>(r);
    return;
  }
  if (p.loss_row == nullptr) return;
  const int which = b - p.q_total;  // 0: t1, 1: t2, 2+i: pred_i
  double acc = 0.0;
  if (which == 0) {
    for (int i = threadIdx.x; i < p.t1_n; i += 256) acc += p.t1_partial[i];
  } else if (which == 1) {
    for (int e = threadIdx.x; e < p.K * p.K; e += 256) {
      const int r = e / p.K, c = e - r * p.K;
      acc += static_cast<double>(p.T[r * p.ldT + c]) * static_cast<double>(p.S[r * p.ldS + c]);
    }
  } else {
    const int i = which - 2;
    for (int q = threadIdx.x; q < p.q_blocks; q += 256) acc += p.pred_partial[static_cast<size_t>(q) * p.n_cov + i];
  }
  const double r = block_sum_256(acc, red);
  if (threadIdx.x == 0) p.loss_row[which] = r;
}

// t1 partials  sum A .* H  over 64-column blocks (same layout as the EPI_H partials), for the block-wise sweep where
// no single H-update launch sees all rows
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ A, long long ldA,
                                                          const float* __restrict__ H, long long ldH, int K, long long L,
                                                          double* __restrict__ partial) {
  __shared__ double red[256];
  const long long c0 = static_cast<long long>(blockIdx.x) * kSLCols;
  double acc = 0.0;
  for (int e = threadIdx.x; e < K * kSLCols; e += 256) {
    const int k = e / kSLCols;
    const long long col = c0 + (e - k * kSLCols);
    if (col < L) acc += static_cast<double>(A[k * ldA + col]) * static_cast<double>(H[k * ldH + col]);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// column sums of W (main.py:776) = row sums of W^T, one block per component, fp64 accumulation, fixed order
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ A, long long ld, long long L,
                                                     float* __restrict__ out) {
  __shared__ double red[256];
  const float* row = A + static_cast<long long>(blockIdx.x) * ld;
  double acc = 0.0;
  for (long long j = threadIdx.x; j < L; j += 256) acc += static_cast<double>(row[j]);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = static_cast<float>(red[0]);
}

}  // namespace alpine
