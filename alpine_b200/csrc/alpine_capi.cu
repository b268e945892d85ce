// C ABI of the B200-native MU-NMF path (see include/alpine_b200.h for the contract and the reference lines each
// entry point replaces).  Host side only: context bookkeeping, TMA descriptors, workspace and kernel launches.
#include "../../include/alpine_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "mu_gemm_sm100.cuh"
#include "mu_small_kernels.cuh"
#include "mu_update_kernels.cuh"
#include "csr_tiles.cuh"
#include "batch_kernels.cuh"
#include "peer_exchange.cuh"
#include "host_upload.cuh"

using namespace alpine;

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU_TRY(expr)                                                                                        \
  do {                                                                                                      \
    cudaError_t e__ = (expr);                                                                               \
    if (e__ != cudaSuccess)                                                                                 \
      return fail(ALPINE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define AL_TRY(expr)          \
  do {                        \
    int r__ = (expr);         \
    if (r__ != ALPINE_OK) return r__; \
  } while (0)
#define LAUNCH_CHECK()                                                                                      \
  do {                                                                                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                                     \
    cudaError_t e__ = cudaGetLastError();                                                                   \
    if (e__ != cudaSuccess)                                                                                 \
      return fail(ALPINE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Every entry point runs on its context's device and puts the caller's current device back on return (the host
// side keeps using torch on whatever device it had selected).
struct DeviceScope {
  int prev = -1;
  bool ok = false;
  explicit DeviceScope(int dev) {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return;
    if (cur == dev) {
      ok = true;
      return;
    }
    ok = cudaSetDevice(dev) == cudaSuccess;
    if (ok) prev = cur;
  }
  ~DeviceScope() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};
#define DEVICE_SCOPE(c)                                  \
  DeviceScope dev_scope__((c)->device);                  \
  if (!dev_scope__.ok) return fail(ALPINE_ERR_CUDA, "cannot select device %d", (c)->device)

// Launch with programmatic stream serialization (see ptx::pdl_enter): only for kernels that execute the wait in
// every CTA.  ALPINE_B200_PDL=0 launches them as ordinary stream-ordered kernels.
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ALPINE_B200_PDL");
    return e == nullptr || strcmp(e, "0") != 0;
  }();
  return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#define PDL_LAUNCH(...)                                                                                     \
  do {                                                                                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                                     \
    cudaError_t e__ = launch_pdl(__VA_ARGS__);                                                              \
    if (e__ != cudaSuccess)                                                                                 \
      return fail(ALPINE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: dims {inner, outer}, row pitch ld (floats), box {box_inner, box_outer}
int make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
             uint32_t box_outer, bool swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(ALPINE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 3) != 0)
    return fail(ALPINE_ERR_ARG, "TMA operand must be 16-byte aligned with a leading dimension divisible by 4");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ALPINE_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ALPINE_OK;
}

// 2-D bf16 tensor map over 16-bit data: dims {inner, outer}, row pitch ld16 (16-bit elements), box {32, box_outer}
// (64-byte rows in shared memory, SWIZZLE_64B)
int make_map_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld16, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(ALPINE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld16 & 7) != 0)
    return fail(ALPINE_ERR_ARG, "bf16 TMA operand must be 16-byte aligned with a leading dimension divisible by 8");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld16 * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ALPINE_ERR_CUDA, "cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r);
  return ALPINE_OK;
}

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }
inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// One contraction out[k][m] = sum_r A(m, r) * B[k][r] over a 2-D fp32 array Xmem[rows][cols] (pitch ldX):
//   ORIENT_XH: m = column of Xmem, r = row of Xmem;   ORIENT_WX: m = row of Xmem, r = column of Xmem.
// B is read from a pre-split workspace (hi at Bsplit, lo at Bsplit + K * ldS).
struct GemmOperands {
  int orient = ORIENT_XH;
  const float* Xmem = nullptr;
  long long ldX = 0, rows = 0, cols = 0;
  const float* Bsplit = nullptr;  // hi copy of the whole K-row operand; the lo copy follows K * ldS floats later
  long long ldS = 0;
  int k0 = 0, Kop = 0;            // component rows [k0, k0 + Kop) of the B operand take part (Kop = 0: all K)
  bool profiled = false;  // counted by alpine_profile (the two contractions over X)
  bool z_slots = false;   // partial sums go to the second slot buffer (they are consumed together with another plan's)
  int chunk_log2 = kChunkLog2;  // TMEM accumulation chain between two fp32 flushes, in k-blocks (log2)
  int group = 0;          // component group of this launch (K > 128 runs as two launches per contraction)
  // ORIENT_WX: one more super-tile taken from this [extra_rows][cols] matrix (pitch ldE): its Gram-type product with
  // the plan's B operand comes out of the same launch (W^T W out of W^T X)
  const float* extra = nullptr;
  long long ldE = 0, extra_rows = 0;
  // sparse X: tile lists instead of Xmem (csr_tiles.cuh)
  const long long* sp_ofs = nullptr;
  const uint2* sp_ent = nullptr;
};
// PLAN_WX_BLOCK + b: W_b^T X of component block b alone (block Gauss-Seidel sweep, main.py:567)
// PLAN_ZW / PLAN_ZH: the K x K-deep products of the Gram reformulation, Z_W = (H H^T) W^T and Z_H = (W^T W) H
// PLAN_WXG: W^T X with W^T W riding along as one more super-tile (the iteration's launch for dense fp32 X); PLAN_WX
// stays the plain contraction (alpine_wx_product, transform, block-wise sweep: same work split as the CSR plan, so
// dense and sparse products stay bit-identical)
enum { PLAN_XH = 0, PLAN_WX = 1, PLAN_GRAM_H = 2, PLAN_GRAM_W = 3, PLAN_ZW = 4, PLAN_ZH = 5, PLAN_WXG = 6,
       PLAN_WX_BLOCK = 7, PLAN_COUNT = 7 + kMaxCov + 1 };

struct GemmPlan {
  bool valid = false;
  GemmOperands op;
  int* d_slot_ofs = nullptr;  // device copies of the reduce kernel's slot lists
  int* d_slots = nullptr;
  CUtensorMap tmX, tmBhi, tmBh16, tmBl16, tmBlo, tmX2;
  int extra_tile = -1;
  GemmParams p{};
  ReduceParams r{};
  int grid = 0;
  size_t smem = 0;
};

}  // namespace

struct alpine_ctx {
  int device = 0;
  long long G = 0, n = 0;
  int K = 0, Kp = 0, n_blocks = 0, n_cov = 0, Kg = 0, q_total = 0;
  int kblk[kMaxCov + 1] = {0};
  int ccov[kMaxCov] = {0};
  int loss_type = LOSS_KL;
  int num_sms = 0;
  bool simt = false;

  const float* X = nullptr;
  long long ldX = 0;
  // sparse X (alpine_bind_csr): library-owned tile lists, one copy per contraction orientation
  bool sparse = false;
  long long nnz = 0;
  long long* sp_ofs[2] = {nullptr, nullptr};
  uint2* sp_ent[2] = {nullptr, nullptr};
  double* sp_xnorm2 = nullptr;
  int* flags = nullptr;  // device: [0] raised by the ||X||^2 pass when some X value is not tf32-exact
  bool x_exact = false;  // host copy of !flags[0] once that pass has run: X needs no lo half (count matrices)
  const float* Y[kMaxCov] = {nullptr};
  float* W = nullptr;
  long long ldW = 0;
  float* H = nullptr;
  long long ldH = 0;
  float* B[kMaxCov] = {nullptr};

  double lam[kMaxCov] = {0};
  double alpha = 0, l1 = 0, orth = 0, eps = 1e-6;
  bool hparams_set = false;

  // workspaces (device)
  long long ldG = 0, ldN = 0;
  float* WT = nullptr;     // [K][ldG]
  float* Hsplit = nullptr; // [3][K][ldN]  split copies of H (ptx::store_split4): tf32 hi plane, then [K][2 ldN]: bf16(hi), bf16(lo) | tf32(lo)  (B operand of X H^T)
  float* Wsplit = nullptr; // [3][K][ldG]  split copies of W^T, same layout (B operand of W^T X)
  float* A = nullptr;      // [K][ldN]   W^T X
  float* numG = nullptr;   // [Kg][ldN]
  float* denG = nullptr;
  float* T = nullptr;      // [K][K]     W^T W
  float* colsum = nullptr; // [K]
  float* q_partial = nullptr;
  double* pred_partial = nullptr;
  int stat_blocks = 0;
  double* t1_partial = nullptr;
  float* hsum_partial = nullptr;  // [sl_blocks_n][K]
  int sl_blocks_n = 0;
  double* sumsq_partial = nullptr;
  // fused update kernels (csrc/mu_update_kernels.cuh): per-CTA partials, summed by the finish kernels
  int upd_grid_h = 0;
  long long ldK = 0;
  float* Ssplit = nullptr;        // [3][K][ldK]  split copies of the complete H H^T (B operand of Z_W)
  float* Tsplit = nullptr;        // [3][K][ldK]  split copies of W^T W             (B operand of Z_H)
  float* hsum_part = nullptr;     // [upd_grid_h][K]
  float* q_part = nullptr;        // [upd_grid_h][q_total]
  double* pred_part = nullptr;    // [upd_grid_h][n_cov]
  double* t1_part = nullptr;      // [upd_grid_h]
  unsigned int* finish_counter = nullptr;
  bool xh_in_slots = false;       // the numerator of the pending W update is still in the contraction's slots
  bool w_stale = false;           // W^T (the master copy) is ahead of the caller's row-major W
  double* xnorm2 = nullptr;
  double* loss_hist = nullptr;
  double* eval_row = nullptr;  // alpine_eval_loss
  int loss_cap = 0;
  int* err = nullptr;
  // partial-sum slot buffers, by class = 2 * (Z plan) + component group: a Z plan's slots are consumed together
  // with the preceding contraction's, and the two groups of one contraction together
  float* pbuf[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t pbuf_floats[4] = {0, 0, 0, 0}, pbuf_hint[4] = {0, 0, 0, 0};
  // component groups: the tcgen05 kernel accumulates at most 128 columns per launch
  int n_groups = 1;
  int gk0[2] = {0, 0}, gK[2] = {0, 0};
  int split() const { return n_groups > 1 ? gk0[1] : 0x3fffffff; }
  // W^T W comes out of the W^T X launch as one more super-tile (dense X that needs its lo half: the count-matrix
  // variant drops the lo half of the streamed operand, which W^T needs; the sparse producer has no TMA path)
  bool gram_w_fused() const { return !sparse && !x_exact; }
  float* own_reduce = nullptr;
  float* reduce = nullptr;  // [Pt K*ldG | S K*K | hsum K | Q q_total]

  // peer exchange over NVLink (alpine_peer_export / _import): this rank's block and the peers' mapped blocks
  float* xchg = nullptr;            // [ reduce buffer | W^T | flags ]
  float* peer_base[kMaxPeers] = {nullptr};
  bool peer_opened[kMaxPeers] = {false};
  int peer_rank = -1, peer_world = 0, peer_epoch = 0;
  float* sum_small = nullptr;       // [S | hsum | Q] summed over the ranks (peer mode)
  float* sum_P = nullptr;           // [K][ldG]: this rank's gene slice of the summed numerator (peer mode)
  bool peer_on() const { return peer_world > 1; }
  // this rank's gene slice of the W update under peer exchange, in whole 64-column tiles
  void peer_slice(long long* g0, long long* g1) const {
    const long long tiles = (G + 63) / 64;
    *g0 = tiles * peer_rank / peer_world * 64;
    *g1 = tiles * (peer_rank + 1) / peer_world * 64;
    if (*g1 > G) *g1 = G;
  }
  long long xchg_wt_off() const { return round_up_ll(reduce_floats(), 64); }
  long long xchg_flag_off() const { return xchg_wt_off() + round_up_ll(static_cast<long long>(K) * ldG, 64); }
  long long xchg_floats() const { return xchg_flag_off() + kPeerFlagInts; }
  static long long round_up_ll(long long v, long long m) { return (v + m - 1) / m * m; }
  // the statistics the W / B updates consume: all-reduced in place by the caller, or summed over peers
  const float* use_S() const { return peer_on() ? sum_small : red_S(); }
  const float* use_hsum() const { return use_S() + static_cast<size_t>(K) * K; }
  const float* use_Q() const { return use_hsum() + K; }

  // Workspace arena lent by the caller (alpine_bind_workspace): a bump allocator over it replaces cudaMalloc /
  // cudaFree for the per-fit buffers, so that creating and destroying a context costs no driver allocation calls
  // (torch's caching allocator owns the memory).  Anything that does not fit falls back to cudaMalloc.
  char* arena = nullptr;
  size_t arena_size = 0, arena_used = 0;
  bool in_arena(const void* q) const {
    const char* b = static_cast<const char*>(q);
    return arena != nullptr && b >= arena && b < arena + arena_size;
  }

  GemmPlan plans[PLAN_COUNT][2];
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
  size_t prof_used = 0;
  bool ws_ready = false;
  bool fit_active = false;

  float* red_Pt() const { return reduce; }
  float* red_S() const { return reduce + static_cast<size_t>(K) * ldG; }
  float* red_hsum() const { return red_S() + static_cast<size_t>(K) * K; }
  float* red_Q() const { return red_hsum() + K; }
  long long reduce_floats() const { return static_cast<long long>(K) * ldG + static_cast<long long>(K) * K + K + q_total; }
  long long small_floats() const { return static_cast<long long>(K) * K + K + q_total; }
};

namespace {

CovTable make_cov_table(const alpine_ctx* c) {
  CovTable t;
  t.n_cov = c->n_cov;
  int row = 0, q = 0;
  for (int i = 0; i < c->n_cov; ++i) {
    t.d[i].row0 = row;
    t.d[i].k = c->kblk[i];
    t.d[i].c = c->ccov[i];
    t.d[i].Y = c->Y[i];
    t.d[i].B = c->B[i];
    t.d[i].q_off = q;
    t.d[i].lam = static_cast<float>(c->lam[i]);
    row += c->kblk[i];
    q += c->ccov[i] * c->kblk[i];
  }
  return t;
}

template <typename T>
int dev_alloc(T** p, size_t count) {
  if (*p != nullptr) return ALPINE_OK;
  CU_TRY(cudaMalloc(reinterpret_cast<void**>(p), (count > 0 ? count : 1) * sizeof(T)));
  return ALPINE_OK;
}

constexpr size_t kWsAlign = 256;
inline size_t ws_bytes(size_t count, size_t elem) { return ((count > 0 ? count : 1) * elem + kWsAlign - 1) / kWsAlign * kWsAlign; }

// workspace buffer of `count` elements: from the caller's arena when it fits, else cudaMalloc
template <typename T>
int ws_alloc(alpine_ctx* c, T** p, size_t count) {
  if (*p != nullptr) return ALPINE_OK;
  const size_t bytes = ws_bytes(count, sizeof(T));
  if (c->arena != nullptr && c->arena_used + bytes <= c->arena_size) {
    *p = reinterpret_cast<T*>(c->arena + c->arena_used);
    c->arena_used += bytes;
    return ALPINE_OK;
  }
  return dev_alloc(p, count);
}
inline void ws_free(const alpine_ctx* c, void* q) {
  if (q != nullptr && !c->in_arena(q)) cudaFree(q);
}

int set_kernel_attrs();

// (initialisations below are issued on the caller's stream, the one the consuming kernels are launched on)
int ensure_flags(alpine_ctx* c, cudaStream_t st) {
  if (c->flags != nullptr) return ALPINE_OK;
  AL_TRY(ws_alloc(c, &c->flags, 2));
  CU_TRY(cudaMemsetAsync(c->flags, 1, 2 * sizeof(int), st));  // any non-zero value = "not known to be tf32-exact"
  return ALPINE_OK;
}

int ensure_workspace(alpine_ctx* c, cudaStream_t st) {
  if (c->ws_ready) return ALPINE_OK;
  AL_TRY(set_kernel_attrs());
  const size_t K = c->K;
  AL_TRY(ws_alloc(c, &c->WT, K * c->ldG));
  AL_TRY(ws_alloc(c, &c->A, K * c->ldN));
  AL_TRY(ws_alloc(c, &c->Hsplit, 3 * K * c->ldN));
  AL_TRY(ws_alloc(c, &c->Wsplit, 3 * K * c->ldG));
  CU_TRY(cudaMemsetAsync(c->Hsplit, 0, 3 * K * c->ldN * sizeof(float), st));
  CU_TRY(cudaMemsetAsync(c->Wsplit, 0, 3 * K * c->ldG * sizeof(float), st));
  AL_TRY(ws_alloc(c, &c->numG, static_cast<size_t>(c->Kg) * c->ldN));
  AL_TRY(ws_alloc(c, &c->denG, static_cast<size_t>(c->Kg) * c->ldN));
  AL_TRY(ws_alloc(c, &c->T, K * K));
  AL_TRY(ws_alloc(c, &c->colsum, K));
  c->stat_blocks = ceil_div(c->n, kStatCells);
  AL_TRY(ws_alloc(c, &c->q_partial, static_cast<size_t>(c->stat_blocks) * (c->q_total > 0 ? c->q_total : 1)));
  AL_TRY(ws_alloc(c, &c->pred_partial, static_cast<size_t>(c->stat_blocks) * (c->n_cov > 0 ? c->n_cov : 1)));
  c->sl_blocks_n = ceil_div(c->n, kSLCols);
  AL_TRY(ws_alloc(c, &c->t1_partial, static_cast<size_t>(c->sl_blocks_n)));
  AL_TRY(ws_alloc(c, &c->hsum_partial, static_cast<size_t>(c->sl_blocks_n) * K));
  AL_TRY(ws_alloc(c, &c->sumsq_partial, 1024));
  {
    const int tiles_h = ceil_div(c->n, kUpdCols);
    c->upd_grid_h = tiles_h < 3 * c->num_sms ? tiles_h : 3 * c->num_sms;
    c->ldK = round_up(c->K, 8);
    AL_TRY(ws_alloc(c, &c->Ssplit, 3 * K * c->ldK));
    AL_TRY(ws_alloc(c, &c->Tsplit, 3 * K * c->ldK));
    CU_TRY(cudaMemsetAsync(c->Ssplit, 0, 3 * K * c->ldK * sizeof(float), st));
    CU_TRY(cudaMemsetAsync(c->Tsplit, 0, 3 * K * c->ldK * sizeof(float), st));
    AL_TRY(ws_alloc(c, &c->hsum_part, K * c->upd_grid_h));
    AL_TRY(ws_alloc(c, &c->q_part, static_cast<size_t>(c->q_total > 0 ? c->q_total : 1) * c->upd_grid_h));
    AL_TRY(ws_alloc(c, &c->pred_part, static_cast<size_t>(c->n_cov > 0 ? c->n_cov : 1) * c->upd_grid_h));
    AL_TRY(ws_alloc(c, &c->t1_part, static_cast<size_t>(c->upd_grid_h)));
    AL_TRY(ws_alloc(c, &c->finish_counter, 4));
    CU_TRY(cudaMemsetAsync(c->finish_counter, 0, 4 * sizeof(unsigned int), st));
  }
  AL_TRY(ws_alloc(c, &c->xnorm2, 1));
  AL_TRY(ws_alloc(c, &c->err, 8));
  CU_TRY(cudaMemsetAsync(c->err, 0, 8 * sizeof(int), st));
  AL_TRY(ensure_flags(c, st));
  CU_TRY(cudaMemsetAsync(c->t1_partial, 0, sizeof(double) * c->sl_blocks_n, st));
  if (c->reduce == nullptr) {
    AL_TRY(ws_alloc(c, &c->own_reduce, static_cast<size_t>(c->reduce_floats())));
    c->reduce = c->own_reduce;
  }
  c->ws_ready = true;
  return ALPINE_OK;
}

int set_kernel_attrs() {
  const int big = 227 * 1024;
#define ALPINE_GEMM_ATTR1(O, NC, EX) \
  CU_TRY(cudaFuncSetAttribute(mu_gemm_kernel<O, NC, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
#define ALPINE_GEMM_ATTR(NC)                                                     \
  ALPINE_GEMM_ATTR1(ORIENT_XH, NC, false) ALPINE_GEMM_ATTR1(ORIENT_WX, NC, false) \
  ALPINE_GEMM_ATTR1(ORIENT_XH, NC, true) ALPINE_GEMM_ATTR1(ORIENT_WX, NC, true)
  ALPINE_GEMM_ATTR(1) ALPINE_GEMM_ATTR(2) ALPINE_GEMM_ATTR(3) ALPINE_GEMM_ATTR(4)
  ALPINE_GEMM_ATTR(5) ALPINE_GEMM_ATTR(6) ALPINE_GEMM_ATTR(7) ALPINE_GEMM_ATTR(8)
#undef ALPINE_GEMM_ATTR
#undef ALPINE_GEMM_ATTR1
  const int sl = static_cast<int>(sym_long_smem_bytes(128));
  CU_TRY(cudaFuncSetAttribute(sym_long_kernel<kSLKI, EPI_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, sl));
  CU_TRY(cudaFuncSetAttribute(sym_long_kernel<kSLKI, EPI_H>, cudaFuncAttributeMaxDynamicSharedMemorySize, sl));
  CU_TRY(cudaFuncSetAttribute(sym_long_kernel<kSLKI, EPI_TRANSFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, sl));
  const int upd = big - 2048;  // (these kernels also hold a little static shared memory)
  CU_TRY(cudaFuncSetAttribute(w_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, upd));
  CU_TRY(cudaFuncSetAttribute(h_update_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd));
  CU_TRY(cudaFuncSetAttribute(h_update_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd));
  CU_TRY(cudaFuncSetAttribute(h_update_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd));
  CU_TRY(cudaFuncSetAttribute(h_update_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd));
  CU_TRY(cudaFuncSetAttribute(cov_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CU_TRY(cudaFuncSetAttribute(guided_terms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return ALPINE_OK;
}

int prof_mark(alpine_ctx* c, cudaStream_t st) {
  if (!c->prof) return ALPINE_OK;
  if (c->prof_used == c->prof_ev.size()) {
    cudaEvent_t e;
    CU_TRY(cudaEventCreate(&e));
    c->prof_ev.push_back(e);
  }
  CU_TRY(cudaEventRecord(c->prof_ev[c->prof_used++], st));
  return ALPINE_OK;
}

template <int ORIENT, bool EXACT>
int launch_gemm_t(const GemmPlan& pl, cudaStream_t st) {
  switch (pl.p.Kp / 16) {
#define ALPINE_GEMM_CASE(NC)                                                                     \
  case NC:                                                                                       \
    PDL_LAUNCH(mu_gemm_kernel<ORIENT, NC, EXACT>, dim3(pl.grid), dim3(kGemmThreads), pl.smem, st, pl.tmX, pl.tmBhi,  \
               pl.tmBh16, pl.tmBl16, pl.tmBlo, pl.tmX2, pl.p);                                   \
    break;
    ALPINE_GEMM_CASE(1) ALPINE_GEMM_CASE(2) ALPINE_GEMM_CASE(3) ALPINE_GEMM_CASE(4)
    ALPINE_GEMM_CASE(5) ALPINE_GEMM_CASE(6) ALPINE_GEMM_CASE(7) ALPINE_GEMM_CASE(8)
#undef ALPINE_GEMM_CASE
    default:
      return fail(ALPINE_ERR_ARG, "unsupported padded component count %d", pl.p.Kp);
  }
  return ALPINE_OK;
}

// Work space, grid and slot count of one contraction; returns the floats its partial-sum slots need.
size_t plan_geometry(const alpine_ctx* c, const GemmOperands& op, GemmParams& p, int* grid_out);
GemmOperands plan_operands(const alpine_ctx* c, int which, int group);
int build_plan_tail(alpine_ctx* c, GemmPlan* pl, const GemmOperands& op, long long R, int Kop, cudaStream_t st);

// Build the plan of one contraction (see GemmOperands).
int build_plan(alpine_ctx* c, GemmPlan* pl, const GemmOperands& op, cudaStream_t st) {
  const long long R = (op.orient == ORIENT_XH) ? op.rows : op.cols;
  pl->op = op;
  GemmParams& p = pl->p;
  const size_t need = plan_geometry(c, op, p, &pl->grid);
  const int Kop = p.K;
  // pipeline depths from the shared-memory budget
  const size_t budget = 227 * 1024 - 1024;
  int sb = 4, sx = 0;
  if (const char* e = getenv("ALPINE_B200_SB")) sb = atoi(e) >= 2 && atoi(e) <= kMaxBStages ? atoi(e) : sb;
  for (; sb >= 2; --sb) {
    sx = kMaxXStages;
    while (sx >= 2 && gemm_smem_layout(p.Kp, sx, sb).total > budget) --sx;
    if (sx >= 3 || sb == 2) break;
  }
  if (sx < 2) return fail(ALPINE_ERR_ARG, "K=%d does not fit the shared-memory pipeline", c->K);
  if (const char* e = getenv("ALPINE_B200_SX")) sx = atoi(e) < sx ? (atoi(e) < 1 ? 1 : atoi(e)) : sx;
  p.sx = sx;
  p.sb = sb;
  pl->smem = gemm_smem_layout(p.Kp, sx, sb).total + 1024;
  // partial-sum slots: per class, one buffer shared by the plans whose results are consumed before the next
  // contraction of that class runs
  const int cls = (op.z_slots ? 2 : 0) + op.group;
  if (need > c->pbuf_floats[cls]) {
    ws_free(c, c->pbuf[cls]);
    c->pbuf[cls] = nullptr;
    // with an arena, size the buffer for the largest of this context's standard plans at once (a bump allocator
    // cannot give memory back)
    const size_t want = (c->arena != nullptr && c->pbuf_hint[cls] > need) ? c->pbuf_hint[cls] : need;
    AL_TRY(ws_alloc(c, &c->pbuf[cls], want));
    c->pbuf_floats[cls] = want;
    for (auto& kind : c->plans)
      for (auto& other : kind)
        if ((other.op.z_slots ? 2 : 0) + other.op.group == cls) other.p.partial = c->pbuf[cls], other.r.partial = c->pbuf[cls];
  }
  return build_plan_tail(c, pl, op, R, Kop, st);
}

size_t plan_geometry(const alpine_ctx* c, const GemmOperands& op, GemmParams& p, int* grid_out) {
  const int rows = kRows;
  const long long M = (op.orient == ORIENT_XH) ? op.cols : op.rows;
  const long long R = (op.orient == ORIENT_XH) ? op.rows : op.cols;
  p.M = static_cast<int>(M);
  p.R = static_cast<int>(R);
  const int Kop = op.Kop > 0 ? op.Kop : c->K;
  p.K = Kop;
  p.Kp = static_cast<int>(round_up(Kop, 16));
  p.ws.num_tiles = ceil_div(M, rows);
  p.extra_tile = -1;
  p.extra_chunk_log2 = 1;  // (as the Gram plans: the result feeds a denominator through Z_H)
  if (op.extra != nullptr && op.orient == ORIENT_WX) p.extra_tile = p.ws.num_tiles++;
  p.ws.kb_per_tile = ceil_div(R, kBK);
  // pieces of the reduction axis: the live window of the B operand (one piece of its hi + lo copies, two around
  // a piece boundary) stays in L2 (evict-last) while X streams through with evict-first
  {
    const double b_bytes = 2.0 * Kop * static_cast<double>(R) * sizeof(float);
    long long piece_mb = 24;
    if (const char* e = getenv("ALPINE_B200_PIECE_MB")) piece_mb = atoll(e) > 0 ? atoll(e) : piece_mb;
    int pieces = static_cast<int>(b_bytes / (piece_mb * 1024.0 * 1024.0) + 0.999);
    if (pieces < 1 || p.ws.num_tiles == 1) pieces = 1;
    int plen = ceil_div(p.ws.kb_per_tile, pieces);
    plen = static_cast<int>(round_up(plen, 8));
    if (plen > p.ws.kb_per_tile) plen = p.ws.kb_per_tile;
    p.ws.piece_len = plen;
    p.ws.pieces = ceil_div(p.ws.kb_per_tile, plen);
  }
  const long long total = p.ws.total();
  const int grid = static_cast<int>(total < c->num_sms ? total : c->num_sms);
  p.max_segs = max_segments_per_cta(p.ws, grid);
  if (grid_out) *grid_out = grid;
  return static_cast<size_t>(grid) * p.max_segs * p.K * rows;
}

int build_plan_tail(alpine_ctx* c, GemmPlan* pl, const GemmOperands& op, long long R, int Kop, cudaStream_t st) {
  const int rows = kRows;
  GemmParams& p = pl->p;
  p.chunk_log2 = op.chunk_log2;
  p.partial = c->pbuf[(op.z_slots ? 2 : 0) + op.group];
  p.err = c->err;
  p.sp_ofs = op.sp_ofs;
  p.sp_ent = op.sp_ent;
  p.sp_total = c->nnz;
  // tensor maps: Xmem is [rows][cols] (inner = cols)
  if (op.sp_ofs != nullptr)
    memset(&pl->tmX, 0, sizeof(pl->tmX));  // the tile-list producer does not use TMA for X
  else if (op.orient == ORIENT_XH)
    AL_TRY(make_map(&pl->tmX, op.Xmem, op.cols, op.rows, op.ldX, rows, kBK, false));
  else
    AL_TRY(make_map(&pl->tmX, op.Xmem, op.cols, op.rows, op.ldX, kBK, rows, true));
  const float* b_hi = op.Bsplit + static_cast<size_t>(op.k0) * op.ldS;
  AL_TRY(make_map(&pl->tmBhi, b_hi, R, Kop, op.ldS, kBK, p.Kp, true));
  // plane 1 of the split copies, [K][2 * ldS] floats: row k = 16-bit bf16(hi)[ldS], bf16(lo)[ldS], then tf32(lo)[ldS]
  const float* b_p1 = op.Bsplit + static_cast<size_t>(c->K) * op.ldS + static_cast<size_t>(op.k0) * 2 * op.ldS;
  const uint16_t* b_16 = reinterpret_cast<const uint16_t*>(b_p1);
  AL_TRY(make_map_bf16(&pl->tmBh16, b_16, R, Kop, 4 * op.ldS, p.Kp));
  AL_TRY(make_map_bf16(&pl->tmBl16, b_16 + op.ldS, R, Kop, 4 * op.ldS, p.Kp));
  AL_TRY(make_map(&pl->tmBlo, b_p1 + op.ldS, R, Kop, 2 * op.ldS, kBK, p.Kp, true));
  pl->extra_tile = p.extra_tile;
  if (p.extra_tile >= 0)
    AL_TRY(make_map(&pl->tmX2, op.extra, op.cols, op.extra_rows, op.ldE, kBK, rows, true));
  else
    pl->tmX2 = pl->tmBhi;  // (unused; any valid descriptor)
  ReduceParams& r = pl->r;
  r.partial = p.partial;
  r.rows = rows;
  r.M = p.M;
  r.K = p.K;
  {
    std::vector<int> ofs(p.ws.num_tiles + 1, 0), slots;
    for (int t = 0; t < p.ws.num_tiles; ++t) {
      reduce_slots_of_tile(p.ws, pl->grid, p.max_segs, t, &slots);
      ofs[t + 1] = static_cast<int>(slots.size());
    }
    ws_free(c, pl->d_slot_ofs);
    ws_free(c, pl->d_slots);
    pl->d_slot_ofs = pl->d_slots = nullptr;
    AL_TRY(ws_alloc(c, &pl->d_slot_ofs, ofs.size()));
    AL_TRY(ws_alloc(c, &pl->d_slots, slots.size() + 1));
    // on the launch stream (and complete before the host vectors go away): ordered before the reduce kernel even when
    // the caller runs on a non-blocking side stream
    CU_TRY(cudaMemcpyAsync(pl->d_slot_ofs, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(pl->d_slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaStreamSynchronize(st));
    r.slot_ofs = pl->d_slot_ofs;
    r.slots = pl->d_slots;
  }
  pl->valid = true;
  return ALPINE_OK;
}

GemmOperands plan_operands(const alpine_ctx* c, int which, int group = 0) {
  GemmOperands op;
  switch (which) {
    case PLAN_XH:  // P^T[k][g] = sum_j X[j][g] H[k][j]                                   (main.py:596)
      op.orient = ORIENT_XH, op.Xmem = c->X, op.ldX = c->ldX, op.rows = c->n, op.cols = c->G;
      op.Bsplit = c->Hsplit, op.ldS = c->ldN, op.profiled = true;
      if (c->sparse) op.Xmem = nullptr, op.sp_ofs = c->sp_ofs[ORIENT_XH], op.sp_ent = c->sp_ent[ORIENT_XH];
      break;
    case PLAN_WX:  // A[k][j] = sum_g X[j][g] W^T[k][g]                                   (main.py:653)
      op.orient = ORIENT_WX, op.Xmem = c->X, op.ldX = c->ldX, op.rows = c->n, op.cols = c->G;
      op.Bsplit = c->Wsplit, op.ldS = c->ldG, op.profiled = true;
      if (c->sparse) op.Xmem = nullptr, op.sp_ofs = c->sp_ofs[ORIENT_WX], op.sp_ent = c->sp_ent[ORIENT_WX];
      break;
    case PLAN_WXG:
      op = plan_operands(c, PLAN_WX, group);
      if (c->gram_w_fused()) op.extra = c->WT, op.ldE = c->ldG, op.extra_rows = c->K;
      return op;
    case PLAN_GRAM_H:  // S[b][a] = sum_j H[a][j] H[b][j]            (H H^T of main.py:599 after the reformulation)
      op.orient = ORIENT_WX, op.Xmem = c->H, op.ldX = c->ldH, op.rows = c->K, op.cols = c->n;
      op.Bsplit = c->Hsplit, op.ldS = c->ldN, op.chunk_log2 = 1;
      break;
    case PLAN_GRAM_W:  // T[b][a] = sum_g W^T[a][g] W^T[b][g]        (W^T W of main.py:654 after the reformulation)
      op.orient = ORIENT_WX, op.Xmem = c->WT, op.ldX = c->ldG, op.rows = c->K, op.cols = c->G;
      op.Bsplit = c->Wsplit, op.ldS = c->ldG, op.chunk_log2 = 1;
      break;
    case PLAN_ZW:      // Z[k][g] = sum_k' S[k][k'] W^T[k'][g]      ((2W) @ H @ H^T of main.py:599 after the reformulation)
      op.orient = ORIENT_XH, op.Xmem = c->WT, op.ldX = c->ldG, op.rows = c->K, op.cols = c->G;
      op.Bsplit = c->Ssplit, op.ldS = c->ldK, op.z_slots = true, op.chunk_log2 = 0;
      if (c->peer_on()) {
        // only this rank's gene slice: the peers store their new slices into the other columns of W^T while this
        // product may still be running
        long long g0, g1;
        c->peer_slice(&g0, &g1);
        op.Xmem = c->WT + g0, op.cols = g1 - g0;
      }
      break;
    case PLAN_ZH:      // Z[k][j] = sum_k' T[k][k'] H[k'][j]        ((2W^T) @ (W @ H) of main.py:654 after the reformulation)
      op.orient = ORIENT_XH, op.Xmem = c->H, op.ldX = c->ldH, op.rows = c->K, op.cols = c->n;
      op.Bsplit = c->Tsplit, op.ldS = c->ldK, op.z_slots = true, op.chunk_log2 = 0;
      break;
    default: {         // A[k][j] = sum_g X[j][g] W^T[k][g] for the rows k of one component block     (main.py:567)
      op = plan_operands(c, PLAN_WX);
      op.k0 = 0;
      const int b = which - PLAN_WX_BLOCK;
      for (int i = 0; i < b; ++i) op.k0 += c->kblk[i];
      op.Kop = c->kblk[b];
      return op;
    }
  }
  if (c->n_groups > 1) {  // this launch produces the components [gk0, gk0 + gK) of the result
    op.group = group;
    op.k0 = c->gk0[group];
    op.Kop = c->gK[group];
  }
  return op;
}

// split copies (ptx::store_split4) of a small operand [K][R]
int run_split(alpine_ctx* c, const float* src, long long ld_src, long long R, float* dst, long long ldS,
              cudaStream_t st) {
  PDL_LAUNCH(split_operand_kernel, dim3(2 * c->num_sms), dim3(256), 0, st, src, ld_src, c->K, R, dst,
             dst + static_cast<size_t>(c->K) * ldS, ldS);
  return ALPINE_OK;
}

// out[k][m] (ld_out) = contraction `which`; its B operand must already be in the split workspace
// out[k][m] (ld_out) = sum of the partial-sum slots contraction `which` has just written (all component groups)
int reduce_slots(alpine_ctx* c, int which, float* out, long long ld_out, cudaStream_t st) {
  const int groups = which < PLAN_WX_BLOCK ? c->n_groups : 1;
  for (int g = 0; g < groups; ++g) {
    const GemmPlan* pl = &c->plans[which][g];
    ReduceParams r = pl->r;
    r.out = out + static_cast<long long>(groups > 1 ? c->gk0[g] : 0) * ld_out;
    r.ld = ld_out;
    if (pl->p.ws.num_tiles * 8 >= 2 * c->num_sms) {
      int gy = ceil_div(32 * c->num_sms, pl->p.ws.num_tiles * 8);
      const int gy_max = ceil_div(pl->p.K, 8);
      if (gy > gy_max) gy = gy_max;
      PDL_LAUNCH(reduce_partials_by_k_kernel, dim3(pl->p.ws.num_tiles * 8, gy), dim3(256), 0, st, r);
    } else {
      PDL_LAUNCH(reduce_partials_kernel, dim3(pl->p.ws.num_tiles * 8, pl->p.K), dim3(256), 0, st, r);
    }
  }
  return ALPINE_OK;
}

// Contraction `which` (one launch per component group).  out == nullptr: leave the result in the partial-sum slots;
// the fused update kernels add them up themselves.  Its B operand must already be in the split workspace.
int run_gemm(alpine_ctx* c, int which, float* out, long long ld_out, cudaStream_t st) {
  const int groups = which < PLAN_WX_BLOCK ? c->n_groups : 1;
  for (int g = 0; g < groups; ++g) {
    GemmPlan* pl = &c->plans[which][g];
    if (!pl->valid) AL_TRY(build_plan(c, pl, plan_operands(c, which, g), st));
    const GemmOperands& op = pl->op;
#ifdef ALPINE_B200_DEBUG_SIMT  // A/B checking builds only: the shipped library has no CUDA-core contraction
    if (c->simt && op.sp_ofs == nullptr && which < PLAN_ZW && out != nullptr && groups == 1) {
      const float* Bsrc = (which == PLAN_XH || which == PLAN_GRAM_H) ? c->H : c->WT;
      const long long ldB = (which == PLAN_XH || which == PLAN_GRAM_H) ? c->ldH : c->ldG;
      dim3 grid(ceil_div(pl->p.M, 128), c->K);
      if (op.orient == ORIENT_XH)
        simt_gemm_kernel<ORIENT_XH><<<grid, 128, 0, st>>>(op.Xmem, op.ldX, Bsrc, ldB, pl->p.M, pl->p.R, c->K, out, ld_out);
      else
        simt_gemm_kernel<ORIENT_WX><<<grid, 128, 0, st>>>(op.Xmem, op.ldX, Bsrc, ldB, pl->p.M, pl->p.R, c->K, out, ld_out);
      LAUNCH_CHECK();
      return ALPINE_OK;
    }
#endif
    if (op.profiled) AL_TRY(prof_mark(c, st));
    // op.profiled marks the contractions whose A operand is X; only those can use the count-matrix variant
    if (op.profiled && c->x_exact) {
      if (op.orient == ORIENT_XH)
        AL_TRY((launch_gemm_t<ORIENT_XH, true>(*pl, st)));
      else
        AL_TRY((launch_gemm_t<ORIENT_WX, true>(*pl, st)));
    } else if (op.orient == ORIENT_XH) {
      AL_TRY((launch_gemm_t<ORIENT_XH, false>(*pl, st)));
    } else {
      AL_TRY((launch_gemm_t<ORIENT_WX, false>(*pl, st)));
    }
    if (op.profiled) AL_TRY(prof_mark(c, st));
  }
  if (out == nullptr) return ALPINE_OK;
  return reduce_slots(c, which, out, ld_out, st);
}

template <int EPI>
int run_sym_long(alpine_ctx* c, const SymLongParams& p, cudaStream_t st) {
  const int blocks = ceil_div(p.L - p.col0, kSLCols);
  if (blocks <= 0) return ALPINE_OK;
  sym_long_kernel<kSLKI, EPI><<<blocks, 256, sym_long_smem_bytes(c->K), st>>>(p);
  LAUNCH_CHECK();
  return ALPINE_OK;
}

// statistics of the current (H, B): S = H H^T, hsum, Q_i (and the prediction-loss partials)
// `fresh_h_update`: the H update kernel just wrote the split copies of H and its per-block row sums
int run_stats(alpine_ctx* c, double* loss_row, bool fresh_h_update, cudaStream_t st) {
  const CovTable tab = make_cov_table(c);
  if (c->n_cov > 0) {
    int kmax = 0, cmax = 0;
    for (int i = 0; i < c->n_cov; ++i) {
      kmax = c->kblk[i] > kmax ? c->kblk[i] : kmax;
      cmax = c->ccov[i] > cmax ? c->ccov[i] : cmax;
    }
    const size_t smem = (static_cast<size_t>(cmax) * kmax + static_cast<size_t>(kmax + cmax) * (kStatCells + 1)) * 4;
    if (smem > 200 * 1024) return fail(ALPINE_ERR_ARG, "covariate block too large for the statistics kernel");
    cov_stats_kernel<<<dim3(c->stat_blocks, c->n_cov), kStatCells, smem, st>>>(
        tab, c->loss_type, c->H, c->ldH, (int)c->n, (float)c->eps, c->q_total, c->q_partial, c->pred_partial);
    LAUNCH_CHECK();
  }
  // S = H H^T as a tcgen05 contraction of H with itself; its split copies also feed the next X H^T
  if (!fresh_h_update) {
    AL_TRY(run_split(c, c->H, c->ldH, c->n, c->Hsplit, c->ldN, st));
    rowsum_kernel<<<c->K, 256, 0, st>>>(c->H, c->ldH, c->n, c->red_hsum());
    LAUNCH_CHECK();
  }
  AL_TRY(run_gemm(c, PLAN_GRAM_H, c->red_S(), c->K, st));
  StatsFinishParams f;
  f.q_partial = c->q_partial;
  f.q_blocks = c->stat_blocks;
  f.q_total = c->q_total;
  f.stats_q = c->red_Q();
  f.pred_partial = c->pred_partial;
  f.n_cov = c->n_cov;
  f.t1_partial = c->t1_partial;
  f.t1_n = c->sl_blocks_n;
  f.T = c->T;
  f.ldT = c->K;
  f.S = c->red_S();
  f.ldS = c->K;
  f.K = c->K;
  f.rowsum_partial = fresh_h_update ? c->hsum_partial : nullptr;
  f.rs_blocks = c->sl_blocks_n;
  f.hsum = c->red_hsum();
  f.loss_row = loss_row;
  stats_finish_kernel<<<c->K + c->q_total + 2 + c->n_cov, 256, 0, st>>>(f);
  LAUNCH_CHECK();
  return ALPINE_OK;
}

// ---- fused update kernels (csrc/mu_update_kernels.cuh)
SlotSrc2 src_slots(const alpine_ctx* c, int which) {
  SlotSrc2 s2{};
  s2.split = c->split();
  for (int g = 0; g < c->n_groups; ++g) {
    const GemmPlan& pl = c->plans[which][g];
    SlotSrc& s = s2.g[g];
    s.direct = nullptr;
    s.partial = pl.p.partial;
    s.slot_ofs = pl.r.slot_ofs;
    s.slots = pl.r.slots;
    s.K = pl.p.K;
  }
  return s2;
}
SlotSrc2 src_direct(const float* a, long long ld) {
  SlotSrc2 s2{};
  s2.split = 0x3fffffff;
  s2.g[0].direct = a;
  s2.g[0].ld = ld;
  return s2;
}

int launch_w_update(alpine_ctx* c, const WUpdParams& p, cudaStream_t st) {
  const long long tiles = ceil_div(p.col1 - p.col0, kUpdCols);
  if (tiles <= 0) return ALPINE_OK;
  const int grid = static_cast<int>(tiles < 3 * c->num_sms ? tiles : 3 * c->num_sms);
  PDL_LAUNCH(w_update_kernel, dim3(grid), dim3(kUpdThreads), w_update_smem_bytes(c->K), st, p);
  return ALPINE_OK;
}

template <bool FIT>
int launch_h_update(alpine_ctx* c, const HUpdParams& p, cudaStream_t st) {
  const int grid = c->upd_grid_h;
  if (grid <= 0) return ALPINE_OK;
  const size_t smem = h_update_smem_bytes(c->K, p.Kg, p.c_total, p.q_total);
  if (smem > 225 * 1024)
    return fail(ALPINE_ERR_ARG, "covariate blocks too large for the H update kernel (%zu bytes of shared memory)", smem);
  // small shards: the W^T X plan cuts every 256-cell super-tile into three or four pieces (fewer super-tiles than SMs),
  // and there are fewer H tiles than two waves of CTAs anyway: the variant that keeps four slots per operand in flight
  if (ceil_div(c->n, kUpdCols) <= 2 * c->num_sms)
    PDL_LAUNCH((h_update_kernel<FIT, 4>), dim3(grid), dim3(kUpdThreads), smem, st, p);
  else
    PDL_LAUNCH((h_update_kernel<FIT, 2>), dim3(grid), dim3(kUpdThreads), smem, st, p);
  return ALPINE_OK;
}

// split copies of a K x K matrix (pitch ld_src, any alignment) into a [3][K][ldK] operand buffer
int run_split_small(alpine_ctx* c, const float* src, int ld_src, float* dst, cudaStream_t st) {
  PDL_LAUNCH(split_small_kernel, dim3(ceil_div(static_cast<long long>(c->K) * c->K, 256)), dim3(256), 0, st, src, ld_src,
             c->K, dst, dst + static_cast<size_t>(c->K) * c->ldK, static_cast<int>(c->ldK));
  return ALPINE_OK;
}

// the caller's row-major W from the W^T master copy, when an update has run since the last export
int export_w(alpine_ctx* c, cudaStream_t st) {
  if (!c->w_stale) return ALPINE_OK;
  transpose_kernel<<<dim3(ceil_div(c->G, 32), ceil_div(c->K, 32)), dim3(32, 8), 0, st>>>(c->WT, c->ldG, c->K, (int)c->G,
                                                                                       c->W, c->ldW);
  LAUNCH_CHECK();
  c->w_stale = false;
  return ALPINE_OK;
}

int check_bound(const alpine_ctx* c, bool need_labels) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  if (c->X == nullptr && !c->sparse) return fail(ALPINE_ERR_STATE, "alpine_bind_dense / alpine_bind_csr has not been called");
  if (c->W == nullptr || c->H == nullptr) return fail(ALPINE_ERR_STATE, "alpine_bind_factors has not been called");
  if (need_labels) {
    for (int i = 0; i < c->n_cov; ++i)
      if (c->Y[i] == nullptr || c->B[i] == nullptr)
        return fail(ALPINE_ERR_STATE, "covariate %d has no labels / B bound", i);
    if (!c->hparams_set) return fail(ALPINE_ERR_STATE, "alpine_set_hparams has not been called");
  }
  return ALPINE_OK;
}

int check_kernel_error(alpine_ctx* c) {
  int h[8];
  CU_TRY(cudaMemcpy(h, c->err, sizeof(h), cudaMemcpyDeviceToHost));
  if (h[0] != 0) {
    cudaMemset(c->err, 0, sizeof(h));
    if (h[0] == ERR_BATCH_INDEX)
      return fail(ALPINE_ERR_ARG, "alpine_batch_gather: idx[%d] = %d is not a cell number of the full-data arrays", h[1], h[2]);
    if (h[0] == ERR_PEER_TIMEOUT)
      return fail(ALPINE_ERR_KERNEL, "peer exchange: rank %d waited too long for rank %d (flag set %d, epoch %d)", h[1], h[2], h[3], h[4]);
    return fail(ALPINE_ERR_KERNEL, "contraction kernel pipeline timeout: code %d block %d thread %d aux %d %d", h[0],
                h[1], h[2], h[3], h[4]);
  }
  return ALPINE_OK;
}

}  // namespace

extern "C" {

int alpine_abi_version(void) { return 11; }
const char* alpine_last_error(void) { return g_last_error.c_str(); }
long long alpine_launch_count(void) { return g_launches.load(); }

int alpine_create(alpine_ctx** out, int device, int64_t n_genes, int64_t n_cells, int n_blocks, const int* k_blocks,
                  int n_cov, const int* c_cov, int loss_type) {
  if (out == nullptr || k_blocks == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  if (n_genes <= 0 || n_cells <= 0) return fail(ALPINE_ERR_ARG, "empty matrix: %lld genes x %lld cells", (long long)n_genes, (long long)n_cells);
  if (n_genes > 0x7fffffffLL || n_cells > 0x7fffffffLL) return fail(ALPINE_ERR_ARG, "dimension exceeds int32");
  if (n_blocks != n_cov + 1 || n_cov < 0 || n_cov > kMaxCov)
    return fail(ALPINE_ERR_ARG, "n_blocks must be n_cov + 1 with n_cov <= %d", kMaxCov);
  if (loss_type != ALPINE_LOSS_KL && loss_type != ALPINE_LOSS_FROBENIUS) return fail(ALPINE_ERR_ARG, "unknown loss_type");
  int dev_count = 0;
  CU_TRY(cudaGetDeviceCount(&dev_count));
  if (device < 0 || device >= dev_count) return fail(ALPINE_ERR_ARG, "device %d not present (%d devices)", device, dev_count);
  // (attributes, not cudaGetDeviceProperties: that call takes ~0.1 s and a context is created per fit)
  int cc_major = 0, cc_minor = 0, sm_count = 0;
  CU_TRY(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  CU_TRY(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  CU_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  if (cc_major != 10) return fail(ALPINE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, cc_major, cc_minor);
  alpine_ctx* c = new alpine_ctx();
  c->device = device;
  c->G = n_genes;
  c->n = n_cells;
  c->n_blocks = n_blocks;
  c->n_cov = n_cov;
  c->loss_type = loss_type;
  c->num_sms = sm_count;
  int K = 0, Kg = 0, q = 0;
  for (int i = 0; i < n_blocks; ++i) {
    if (k_blocks[i] <= 0) {
      delete c;
      return fail(ALPINE_ERR_ARG, "block %d has %d components", i, k_blocks[i]);
    }
    c->kblk[i] = k_blocks[i];
    K += k_blocks[i];
    if (i < n_cov) {
      if (c_cov == nullptr || c_cov[i] <= 0) {
        delete c;
        return fail(ALPINE_ERR_ARG, "covariate %d has no categories", i);
      }
      c->ccov[i] = c_cov[i];
      Kg += k_blocks[i];
      q += c_cov[i] * k_blocks[i];
    }
  }
  if (K > 2 * kAccStride) {
    delete c;
    return fail(ALPINE_ERR_ARG, "total components %d > %d is not supported (two launches of at most %d accumulator "
                "columns per contraction)", K, 2 * kAccStride, kAccStride);
  }
  c->K = K;
  if (K > kAccStride) {  // two balanced component groups, the first a multiple of 16
    c->n_groups = 2;
    c->gK[0] = static_cast<int>(round_up((K + 1) / 2, 16));
    c->gK[1] = K - c->gK[0];
    c->gk0[1] = c->gK[0];
  } else {
    c->gK[0] = K;
  }
  c->Kp = static_cast<int>(round_up(K, 16));
  c->Kg = Kg;
  c->q_total = q;
#ifdef ALPINE_B200_DEBUG_SIMT
  if (const char* e = getenv("ALPINE_B200_GEMM")) c->simt = (strcmp(e, "simt") == 0);
#endif
  c->ldG = round_up(n_genes, 8);  // (8: the bf16 halves of the split copies start on 16-byte boundaries)
  c->ldN = round_up(n_cells, 8);
  *out = c;
  return ALPINE_OK;
}

int alpine_destroy(alpine_ctx* c) {
  if (c == nullptr) return ALPINE_OK;
  DeviceScope dev_scope__(c->device);
  for (int q = 0; q < kMaxPeers; ++q)
    if (c->peer_opened[q]) cudaIpcCloseMemHandle(c->peer_base[q]);
  if (c->xchg != nullptr) {
    c->WT = nullptr;  // lives inside the exchange block
    cudaFree(c->xchg);
  }
  ws_free(c, c->sum_small);
  ws_free(c, c->sum_P);
  void* ptrs[] = {c->WT, c->Hsplit, c->Wsplit, c->A, c->numG, c->denG, c->T, c->colsum, c->q_partial,
                  c->pred_partial, c->t1_partial, c->hsum_partial, c->sumsq_partial, c->xnorm2, c->loss_hist, c->eval_row, c->err, c->pbuf[0], c->pbuf[1], c->pbuf[2], c->pbuf[3],
                  c->own_reduce, c->sp_ofs[0], c->sp_ofs[1], c->sp_ent[0], c->sp_ent[1], c->sp_xnorm2, c->flags,
                  c->Ssplit, c->Tsplit, c->hsum_part, c->q_part, c->pred_part, c->t1_part, c->finish_counter};
  for (void* p : ptrs) ws_free(c, p);
  for (auto& kind : c->plans)
    for (auto& pl : kind) {
      ws_free(c, pl.d_slot_ofs);
      ws_free(c, pl.d_slots);
    }
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  delete c;
  return ALPINE_OK;
}

int64_t alpine_workspace_bytes(const alpine_ctx* c) {
  if (c == nullptr) return 0;
  const size_t K = c->K, f = sizeof(float);
  size_t b = 0;
  b += ws_bytes(K * c->ldG, f) + ws_bytes(K * c->ldN, f);                  // W^T, A
  b += ws_bytes(3 * K * c->ldN, f) + ws_bytes(3 * K * c->ldG, f);          // split copies
  b += 2 * ws_bytes(static_cast<size_t>(c->Kg) * c->ldN, f);              // guided terms
  b += ws_bytes(K * K, f) + ws_bytes(K, f);
  const size_t stat_blocks = ceil_div(c->n, kStatCells), sl_blocks = ceil_div(c->n, kSLCols);
  b += ws_bytes(stat_blocks * (c->q_total > 0 ? c->q_total : 1), f) + ws_bytes(stat_blocks * (c->n_cov > 0 ? c->n_cov : 1), 8);
  b += ws_bytes(sl_blocks, 8) + ws_bytes(sl_blocks * K, f) + ws_bytes(1024, 8) + 4 * kWsAlign;
  b += ws_bytes(static_cast<size_t>(c->reduce_floats()), f);
  {
    const size_t gh = ceil_div(c->n, kUpdCols) < 3 * c->num_sms ? ceil_div(c->n, kUpdCols) : 3 * c->num_sms;
    b += 2 * ws_bytes(3 * K * round_up(c->K, 8), f) + ws_bytes(K * gh, f);
    b += ws_bytes((c->q_total > 0 ? c->q_total : 1) * gh, f) + ws_bytes((c->n_cov > 0 ? c->n_cov : 1) * gh, 8);
    b += ws_bytes(gh, 8) + kWsAlign;
  }
  // partial-sum slots of the largest standard plan, and the slot lists of all of them
  size_t hint[4] = {0, 0, 0, 0};
  for (int which = 0; which < PLAN_WX_BLOCK; ++which)
    for (int g = 0; g < c->n_groups; ++g) {
      GemmParams p{};
      int grid = 0;
      const GemmOperands op = plan_operands(c, which, g);
      const size_t need = plan_geometry(c, op, p, &grid);
      const int cls = (op.z_slots ? 2 : 0) + op.group;
      hint[cls] = need > hint[cls] ? need : hint[cls];
      b += ws_bytes(p.ws.num_tiles + 1, 4) + ws_bytes(static_cast<size_t>(p.ws.num_tiles) * p.ws.pieces * 3 + grid + 1, 4);
    }
  for (size_t h : hint) b += ws_bytes(h, f);
  return static_cast<int64_t>(b + (2u << 20));  // + loss history, flags, alignment slack
}

int alpine_bind_workspace(alpine_ctx* c, void* base, int64_t bytes) {
  if (c == nullptr || base == nullptr || bytes <= 0) return fail(ALPINE_ERR_ARG, "null / empty workspace");
  if (c->ws_ready || c->arena != nullptr) return fail(ALPINE_ERR_STATE, "the workspace must be bound once, right after alpine_create");
  if ((reinterpret_cast<uintptr_t>(base) & (kWsAlign - 1)) != 0) return fail(ALPINE_ERR_ARG, "workspace must be 256-byte aligned");
  c->arena = static_cast<char*>(base);
  c->arena_size = static_cast<size_t>(bytes);
  c->arena_used = 0;
  for (int which = 0; which < PLAN_WX_BLOCK; ++which)
    for (int g = 0; g < c->n_groups; ++g) {
      GemmParams p{};
      const GemmOperands op = plan_operands(c, which, g);
      const size_t need = plan_geometry(c, op, p, nullptr);
      const int cls = (op.z_slots ? 2 : 0) + op.group;
      c->pbuf_hint[cls] = need > c->pbuf_hint[cls] ? need : c->pbuf_hint[cls];
    }
  return ALPINE_OK;
}

int alpine_bind_dense(alpine_ctx* c, const float* X, int64_t ldX) {
  if (c == nullptr || X == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  if (ldX < c->G || (ldX & 3) != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0)
    return fail(ALPINE_ERR_ARG, "X needs ldX >= n_genes, ldX %% 4 == 0 and a 16-byte aligned base");
  c->X = X;
  c->ldX = ldX;
  c->sparse = false;
  c->x_exact = false;  // until alpine_fit_begin has looked at the values
  for (auto& kind : c->plans)
    for (auto& pl : kind) pl.valid = false;
  return ALPINE_OK;
}

int alpine_bind_csr(alpine_ctx* c, const int64_t* indptr, const int32_t* indices, const float* values, int64_t nnz,
                    void* stream) {
  if (c == nullptr || indptr == nullptr || (nnz > 0 && (indices == nullptr || values == nullptr)))
    return fail(ALPINE_ERR_ARG, "null argument");
  if (nnz < 0) return fail(ALPINE_ERR_ARG, "negative nnz");
  DEVICE_SCOPE(c);
  AL_TRY(ensure_flags(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int o = 0; o < 2; ++o) {
    if (c->sp_ofs[o]) cudaFree(c->sp_ofs[o]);
    if (c->sp_ent[o]) cudaFree(c->sp_ent[o]);
    c->sp_ofs[o] = nullptr;
    c->sp_ent[o] = nullptr;
  }
  const int kb_xh = ceil_div(c->n, kBK), kb_wx = ceil_div(c->G, kBK);
  const long long nb_xh = static_cast<long long>(ceil_div(c->G, kRows)) * kb_xh;
  const long long nb_wx = static_cast<long long>(ceil_div(c->n, kRows)) * kb_wx;
  unsigned int *cnt_xh = nullptr, *cnt_wx = nullptr;
  int* err = nullptr;
  double* partial = nullptr;
  auto cleanup = [&]() {
    cudaFree(cnt_xh), cudaFree(cnt_wx), cudaFree(err), cudaFree(partial);
  };
#define CSR_TRY(expr)                                                                                        \
  do {                                                                                                       \
    cudaError_t e__ = (expr);                                                                                \
    if (e__ != cudaSuccess) {                                                                                \
      cleanup();                                                                                             \
      return fail(ALPINE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                                        \
  } while (0)
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&cnt_xh), nb_xh * sizeof(unsigned int)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&cnt_wx), nb_wx * sizeof(unsigned int)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&err), 2 * sizeof(int)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&partial), 1024 * sizeof(double)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&c->sp_ofs[ORIENT_XH]), (nb_xh + 1) * sizeof(long long)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&c->sp_ofs[ORIENT_WX]), (nb_wx + 1) * sizeof(long long)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&c->sp_ent[ORIENT_XH]), (nnz > 0 ? nnz : 1) * sizeof(uint2)));
  CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&c->sp_ent[ORIENT_WX]), (nnz > 0 ? nnz : 1) * sizeof(uint2)));
  if (c->sp_xnorm2 == nullptr) CSR_TRY(cudaMalloc(reinterpret_cast<void**>(&c->sp_xnorm2), sizeof(double)));
  CSR_TRY(cudaMemsetAsync(cnt_xh, 0, nb_xh * sizeof(unsigned int), st));
  CSR_TRY(cudaMemsetAsync(cnt_wx, 0, nb_wx * sizeof(unsigned int), st));
  CSR_TRY(cudaMemsetAsync(err, 0, 2 * sizeof(int), st));
  CSR_TRY(cudaMemsetAsync(c->flags, 0, sizeof(int), st));
  CsrView m{reinterpret_cast<const long long*>(indptr), indices, values, c->n, static_cast<int>(c->G)};
  const int grid = 8 * c->num_sms;
  csr_tile_count_kernel<<<grid, 256, 0, st>>>(m, kb_xh, kb_wx, cnt_xh, cnt_wx, err);
  csr_tile_scan_kernel<<<1, 1024, 0, st>>>(cnt_xh, nb_xh, c->sp_ofs[ORIENT_XH]);
  csr_tile_scan_kernel<<<1, 1024, 0, st>>>(cnt_wx, nb_wx, c->sp_ofs[ORIENT_WX]);
  csr_tile_fill_kernel<<<grid, 256, 0, st>>>(m, kb_xh, kb_wx, c->sp_ofs[ORIENT_XH], c->sp_ofs[ORIENT_WX], cnt_xh, cnt_wx,
                                            c->sp_ent[ORIENT_XH], c->sp_ent[ORIENT_WX]);
  // ||X||_F^2 (first term of the trace identity that replaces main.py:736) and the tf32-exactness flag
  sumsq_flat_kernel<<<1024, 256, 0, st>>>(values, nnz, partial, c->flags);
  sum_double_kernel<<<1, 32, 0, st>>>(partial, 1024, c->sp_xnorm2);
  g_launches.fetch_add(6, std::memory_order_relaxed);
  CSR_TRY(cudaGetLastError());
  int h_err[2] = {0, 0}, h_inexact = 1;
  long long total = 0;
  CSR_TRY(cudaMemcpyAsync(h_err, err, sizeof(h_err), cudaMemcpyDeviceToHost, st));
  CSR_TRY(cudaMemcpyAsync(&h_inexact, c->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  CSR_TRY(cudaMemcpyAsync(&total, c->sp_ofs[ORIENT_WX] + nb_wx, sizeof(long long), cudaMemcpyDeviceToHost, st));
  CSR_TRY(cudaStreamSynchronize(st));
#undef CSR_TRY
  cleanup();
  if (h_err[0]) return fail(ALPINE_ERR_ARG, "CSR column index outside [0, n_genes)");
  if (h_err[1]) return fail(ALPINE_ERR_ARG, "CSR values must be finite and non-negative");
  if (total != nnz) return fail(ALPINE_ERR_ARG, "indptr covers %lld nonzeros but nnz = %lld", total, (long long)nnz);
  c->X = nullptr;
  c->sparse = true;
  c->x_exact = (h_inexact == 0);
  c->nnz = nnz;
  for (auto& kind : c->plans)
    for (auto& pl : kind) pl.valid = false;
  return ALPINE_OK;
}

int alpine_bind_labels(alpine_ctx* c, int i, const float* Y) {
  if (c == nullptr || Y == nullptr || i < 0 || i >= c->n_cov) return fail(ALPINE_ERR_ARG, "bad covariate index %d", i);
  c->Y[i] = Y;
  return ALPINE_OK;
}

int alpine_bind_factors(alpine_ctx* c, float* W, int64_t ldW, float* H, int64_t ldH, float* const* Bs) {
  if (c == nullptr || W == nullptr || H == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  if (ldW < c->K) return fail(ALPINE_ERR_ARG, "ldW < K");
  if (ldH < c->n || (ldH & 3) != 0 || (reinterpret_cast<uintptr_t>(H) & 15) != 0)
    return fail(ALPINE_ERR_ARG, "H needs ldH >= n_cells, ldH %% 4 == 0 and a 16-byte aligned base");
  if (c->n_cov > 0 && Bs == nullptr) return fail(ALPINE_ERR_ARG, "Bs is null");
  c->W = W;
  c->ldW = ldW;
  c->H = H;
  c->ldH = ldH;
  for (int i = 0; i < c->n_cov; ++i) c->B[i] = Bs[i];
  for (auto& kind : c->plans)
    for (auto& pl : kind) pl.valid = false;
  c->w_stale = false;  // the bound W is the truth until an update runs
  return ALPINE_OK;
}

int alpine_set_hparams(alpine_ctx* c, const double* lam, double alpha_W, double l1_ratio_W, double orth_W, double eps) {
  if (c == nullptr || (c->n_cov > 0 && lam == nullptr)) return fail(ALPINE_ERR_ARG, "null argument");
  for (int i = 0; i < c->n_cov; ++i) c->lam[i] = lam[i];
  c->alpha = alpha_W;
  c->l1 = l1_ratio_W;
  c->orth = orth_W;
  c->eps = eps;
  c->hparams_set = true;
  return ALPINE_OK;
}

int64_t alpine_reduce_buffer_size(const alpine_ctx* c) { return c ? c->reduce_floats() : 0; }

int alpine_bind_reduce_buffer(alpine_ctx* c, float* buf) {
  if (c == nullptr || buf == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  if ((reinterpret_cast<uintptr_t>(buf) & 15) != 0) return fail(ALPINE_ERR_ARG, "reduce buffer must be 16-byte aligned");
  c->reduce = buf;
  return ALPINE_OK;
}

int alpine_fit_begin(alpine_ctx* c, int max_iter, void* stream) {
  AL_TRY(check_bound(c, true));
  if (max_iter <= 0) return fail(ALPINE_ERR_ARG, "max_iter must be positive");
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (max_iter > c->loss_cap) {
    ws_free(c, c->loss_hist);
    c->loss_hist = nullptr;
    AL_TRY(ws_alloc(c, &c->loss_hist, static_cast<size_t>(max_iter) * (2 + c->n_cov)));
    c->loss_cap = max_iter;
  }
  // ||X||_F^2 (first term of the trace identity that replaces main.py:736)
  if (c->sparse) {
    CU_TRY(cudaMemcpyAsync(c->xnorm2, c->sp_xnorm2, sizeof(double), cudaMemcpyDeviceToDevice, st));
  } else {
    const int sb = 1024;
    CU_TRY(cudaMemsetAsync(c->flags, 0, sizeof(int), st));
    sumsq_partial_kernel<<<sb, 256, 0, st>>>(c->X, c->ldX, c->n, (int)c->G, c->sumsq_partial, c->flags);
    LAUNCH_CHECK();
    sum_double_kernel<<<1, 32, 0, st>>>(c->sumsq_partial, sb, c->xnorm2);
    LAUNCH_CHECK();
    int h_inexact = 1;  // one 4-byte read-back per fit selects the 2-MMA kernel variant for count matrices
    CU_TRY(cudaMemcpyAsync(&h_inexact, c->flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (c->x_exact != (h_inexact == 0)) {  // the kernel variant (and what rides along with W^T X) depends on it
      for (auto& kind : c->plans)
        for (auto& pl : kind) pl.valid = false;
    }
    c->x_exact = (h_inexact == 0);
  }
  // W^T master copy for the gene-side kernels
  AL_TRY(export_w(c, st));
  transpose_kernel<<<dim3(ceil_div(c->K, 32), ceil_div(c->G, 32)), dim3(32, 8), 0, st>>>(c->W, c->ldW, (int)c->G, c->K,
                                                                                       c->WT, c->ldG);
  LAUNCH_CHECK();
  // split copies of the whole W^T: the simultaneous update rewrites all of them before their first use, the
  // block-wise sweep (alpine_als_block) only the rows of the block it has just updated
  AL_TRY(run_split(c, c->WT, c->ldG, c->G, c->Wsplit, c->ldG, st));
  AL_TRY(run_stats(c, nullptr, false, st));
  AL_TRY(run_split_small(c, c->red_S(), c->K, c->Ssplit, st));  // (re-done after the exchange under cell sharding)
  CU_TRY(cudaMemsetAsync(c->finish_counter, 0, 4 * sizeof(unsigned int), st));  // (a faulted launch may have left it set)
  c->fit_active = true;
  return ALPINE_OK;
}

int alpine_batch_begin(alpine_ctx* c, void* stream) {
  AL_TRY(check_bound(c, true));
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->loss_cap < 1) {
    AL_TRY(ws_alloc(c, &c->loss_hist, static_cast<size_t>(2 + c->n_cov)));
    c->loss_cap = 1;
  }
  if (!c->w_stale) {
    // the bound row-major W is the truth (first batch, or another context has updated it since): take W^T and its split
    // copies from it.  Otherwise this context's own update wrote both a moment ago and nobody has asked for W since
    // (no alpine_sync_w): consecutive batches of one context skip three G x K passes.
    transpose_kernel<<<dim3(ceil_div(c->K, 32), ceil_div(c->G, 32)), dim3(32, 8), 0, st>>>(c->W, c->ldW, (int)c->G,
                                                                                         c->K, c->WT, c->ldG);
    LAUNCH_CHECK();
    // split copies of the whole W^T: the simultaneous update rewrites all of them before their first use, the
    // block-wise sweep (alpine_als_block) only the rows of the block it has just updated
    AL_TRY(run_split(c, c->WT, c->ldG, c->G, c->Wsplit, c->ldG, st));
  }
  AL_TRY(run_stats(c, nullptr, false, st));
  AL_TRY(run_split_small(c, c->red_S(), c->K, c->Ssplit, st));
  CU_TRY(cudaMemsetAsync(c->finish_counter, 0, 4 * sizeof(unsigned int), st));
  c->fit_active = true;
  return ALPINE_OK;
}

int alpine_batch_gather(alpine_ctx* c, const float* X_all, int64_t ldX_all, float* X_batch, const float* H_all,
                        int64_t ldH_all, const float* const* Y_all, float* const* Y_batch, int64_t n_all,
                        const int64_t* idx, int64_t cnt, void* stream) {
  AL_TRY(check_bound(c, true));
  if (H_all == nullptr || idx == nullptr || n_all <= 0 || cnt < 0 || cnt > c->n || ldH_all < n_all)
    return fail(ALPINE_ERR_ARG, "bad batch: %lld of %lld cells into a context of %lld", (long long)cnt, (long long)n_all, (long long)c->n);
  if (!c->sparse && (X_all == nullptr || X_batch != c->X || ldX_all < c->G))
    return fail(ALPINE_ERR_ARG, "X_batch must be the array bound with alpine_bind_dense and X_all cells-major with ldX_all >= n_genes");
  if (c->n_cov > 0 && (Y_all == nullptr || Y_batch == nullptr)) return fail(ALPINE_ERR_ARG, "null label arrays");
  DEVICE_SCOPE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(ensure_workspace(c, st));
  BatchGatherParams p{};
  p.idx = reinterpret_cast<const long long*>(idx);
  p.cnt = cnt, p.n = c->n, p.n_all = n_all;
  p.X_all = c->sparse ? nullptr : X_all;
  p.ldX_all = ldX_all;
  p.X = c->sparse ? nullptr : X_batch;
  p.ldX = c->ldX, p.G = c->G;
  p.vec4 = !c->sparse && ((reinterpret_cast<uintptr_t>(X_all) | reinterpret_cast<uintptr_t>(X_batch)) & 15) == 0 &&
           ((ldX_all | c->ldX) & 3) == 0;
  p.H_all = H_all, p.ldH_all = ldH_all, p.H = c->H, p.ldH = c->ldH, p.K = c->K;
  p.n_cov = c->n_cov;
  long long rows = c->K;
  for (int i = 0; i < c->n_cov; ++i) {
    if (Y_all[i] == nullptr || Y_batch[i] != c->Y[i])
      return fail(ALPINE_ERR_ARG, "Y_batch[%d] must be the array bound with alpine_bind_labels", i);
    p.Y_all[i] = Y_all[i], p.Y[i] = Y_batch[i], p.c[i] = c->ccov[i];
    rows += c->ccov[i];
  }
  p.err = c->err;
  // one row of X per warp and pass: enough blocks to keep every SM's load queues full, no more than there are rows
  p.x_blocks = c->sparse ? 0 : static_cast<int>(std::min<long long>(ceil_div(c->n, 8), 8ll * c->num_sms));
  const int hy_blocks = static_cast<int>(std::min<long long>(ceil_div(rows * c->n, 256), 2ll * c->num_sms));
  batch_gather_kernel<<<p.x_blocks + hy_blocks, 256, 0, st>>>(p);
  LAUNCH_CHECK();
  return ALPINE_OK;
}

int alpine_batch_scatter(alpine_ctx* c, float* H_all, int64_t ldH_all, int64_t n_all, const int64_t* idx, int64_t cnt,
                         void* stream) {
  if (c == nullptr || c->H == nullptr) return fail(ALPINE_ERR_STATE, "alpine_bind_factors has not been called");
  if (H_all == nullptr || idx == nullptr || n_all <= 0 || cnt < 0 || cnt > c->n || ldH_all < n_all)
    return fail(ALPINE_ERR_ARG, "bad batch");
  if (cnt == 0) return ALPINE_OK;
  DEVICE_SCOPE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = static_cast<int>(std::min<long long>(ceil_div(static_cast<long long>(c->K) * cnt, 256), 4ll * c->num_sms));
  batch_scatter_kernel<<<blocks, 256, 0, st>>>(c->H, c->ldH, c->K, reinterpret_cast<const long long*>(idx), cnt, n_all,
                                               H_all, ldH_all);
  LAUNCH_CHECK();
  return ALPINE_OK;
}

int alpine_mu_partials(alpine_ctx* c, void* stream) {
  AL_TRY(check_bound(c, true));
  if (!c->fit_active) return fail(ALPINE_ERR_STATE, "alpine_fit_begin has not been called");
  DEVICE_SCOPE(c);
  // Single GPU (no caller-owned reduce buffer, no peers): nobody else needs the numerator as an array, the W update
  // sums it straight from the contraction's partial slots.  Otherwise it goes into the exchange buffer.
  c->xh_in_slots = !c->peer_on() && c->reduce == c->own_reduce;
  return run_gemm(c, PLAN_XH, c->xh_in_slots ? nullptr : c->red_Pt(), c->ldG, static_cast<cudaStream_t>(stream));  // Hsplit is current
}

}  // extern "C"

namespace {

PeerTable make_peer_table(const alpine_ctx* c) {
  PeerTable t{};
  t.world = c->peer_world;
  t.rank = c->peer_rank;
  for (int q = 0; q < c->peer_world; ++q) {
    t.flags[q] = reinterpret_cast<int*>(c->peer_base[q] + c->xchg_flag_off());
    t.small[q] = c->peer_base[q] + static_cast<size_t>(c->K) * c->ldG;
  }
  return t;
}

// One simultaneous-update iteration after alpine_mu_partials.  peer = false: the reduce buffer holds the complete
// numerator and statistics (single GPU, or all-reduced by the caller).  peer = true: csrc/peer_exchange.cuh.
int mu_apply_impl(alpine_ctx* c, int iter, void* stream, bool peer) {
  AL_TRY(check_bound(c, true));
  if (!c->fit_active) return fail(ALPINE_ERR_STATE, "alpine_fit_begin has not been called");
  if (iter < 0 || iter >= c->loss_cap) return fail(ALPINE_ERR_ARG, "iteration %d outside [0, %d)", iter, c->loss_cap);
  if (peer != c->peer_on())
    return fail(ALPINE_ERR_STATE, peer ? "alpine_peer_import has not been called"
                                       : "this context exchanges over peer memory: call alpine_mu_apply_peer");
  DEVICE_SCOPE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const CovTable tab = make_cov_table(c);
  int ckmax = 1, c_total = 0;
  for (int i = 0; i < c->n_cov; ++i) {
    ckmax = c->ccov[i] * c->kblk[i] > ckmax ? c->ccov[i] * c->kblk[i] : ckmax;
    c_total += c->ccov[i];
  }
  const bool single = !peer && c->reduce == c->own_reduce;  // no exchange: S is this GPU's own, already split
  // ---- W update (main.py:592-612) on W^T (the caller's row-major W is refreshed on demand: export_w)
  WUpdParams w{};
  w.WT = c->WT;
  w.ldG = c->ldG;
  w.K = c->K;
  w.col0 = 0, w.col1 = c->G;
  w.c1 = static_cast<float>((1.0 - c->l1) * c->alpha);
  w.c2 = static_cast<float>(c->l1 * c->alpha);
  w.orth = static_cast<float>(c->orth);
  w.eps = static_cast<float>(c->eps);
  w.split_hi = c->Wsplit;  // B operand of W^T W and W^T X below
  w.split_lo = c->Wsplit + static_cast<size_t>(c->K) * c->ldG;
  if (!peer) {
    w.num = c->xh_in_slots ? src_slots(c, PLAN_XH) : src_direct(c->red_Pt(), c->ldG);
    if (!single) AL_TRY(run_split_small(c, c->use_S(), c->K, c->Ssplit, st));  // the all-reduced H H^T
    AL_TRY(run_gemm(c, PLAN_ZW, nullptr, 0, st));   // Z_W = (H H^T) W^T, into the second slot buffer
    w.z = src_slots(c, PLAN_ZW);
    AL_TRY(launch_w_update(c, w, st));
  } else {
    const int epoch = ++c->peer_epoch;
    const PeerTable pt = make_peer_table(c);
    long long g0, g1;
    c->peer_slice(&g0, &g1);
    const int n_small = static_cast<int>(c->small_floats());
    PDL_LAUNCH(peer_gather_reduce_kernel, dim3(2 * c->num_sms), dim3(256), 0, st, pt, epoch, n_small, c->sum_small, c->K,
               c->ldG, g0, g1, c->sum_P, c->err, c->Ssplit, c->Ssplit + static_cast<size_t>(c->K) * c->ldK,
               static_cast<int>(c->ldK));
    AL_TRY(run_gemm(c, PLAN_ZW, nullptr, 0, st));  // (this rank's gene slice only, see plan_operands)
    w.z = src_slots(c, PLAN_ZW);
    w.z.g[0].origin = w.z.g[1].origin = g0;
    w.col0 = g0;
    w.col1 = g1;
    w.num = src_direct(c->sum_P, c->ldG);
    w.n_peers = c->peer_world;
    for (int q = 0; q < c->peer_world; ++q)
      w.wt_peer[q] = (q == c->peer_rank) ? nullptr : c->peer_base[q] + c->xchg_wt_off();
    w.split_hi = w.split_lo = nullptr;  // taken from the gathered W^T below
    AL_TRY(launch_w_update(c, w, st));
    // close of the exchange (every slice of the new W^T is in every block) fused with the split of the gathered W^T
    PDL_LAUNCH(peer_close_and_split_kernel, dim3(2 * c->num_sms), dim3(256), 0, st, pt, 1, epoch, c->err, c->WT, c->ldG,
               c->K, c->G, c->Wsplit, c->Wsplit + static_cast<size_t>(c->K) * c->ldG, c->ldG);
  }
  c->w_stale = true;
  // ---- A = W^T X (main.py:653), left in its slots; T = W^T W of the new W rides along as one more super-tile of the
  //      same launch (dense fp32 X), otherwise it is a Gram plan of its own.  The finish kernel sums T's slots, writes
  //      the split copies Z_H needs and applies the B updates (main.py:615-628).
  const bool fused_t = c->gram_w_fused();
  // (the Gram plan shares the slot buffer with W^T X: unfused, its slots are consumed before W^T X runs)
  AL_TRY(run_gemm(c, fused_t ? PLAN_WXG : PLAN_GRAM_W, nullptr, 0, st));
  {
    WFinishParams wf{};
    wf.gram.src = src_slots(c, fused_t ? PLAN_WXG : PLAN_GRAM_W);
    wf.gram.tile = fused_t ? c->plans[PLAN_WXG][0].extra_tile : 0;
    wf.gram.K = c->K;
    wf.gram.out = c->T;
    wf.gram.ld = c->K;
    wf.gram.split_hi = c->Tsplit;
    wf.gram.split_lo = c->Tsplit + static_cast<size_t>(c->K) * c->ldK;
    wf.gram.ld_split = static_cast<int>(c->ldK);
    wf.gram_blocks = gram_blocks_for(c->K);
    wf.cov = tab;
    wf.loss_type = c->loss_type;
    wf.stats_q = c->use_Q();
    wf.hsum = c->use_hsum();
    wf.S = c->use_S();
    wf.ldS = c->K;
    wf.eps = static_cast<float>(c->eps);
    PDL_LAUNCH(w_finish_kernel, dim3(wf.gram_blocks + (c->n_cov > 0 ? 1 : 0)), dim3(256), ckmax * sizeof(float), st, wf);
  }
  if (!fused_t) AL_TRY(run_gemm(c, PLAN_WX, nullptr, 0, st));
  // ---- Z_H = (W^T W) H, left in its slots
  AL_TRY(run_gemm(c, PLAN_ZH, nullptr, 0, st));
  // ---- H update (main.py:631-663) with the guided terms of (old H, new B), statistics of (new H, new B)
  HUpdParams h{};
  h.H = c->H;
  h.ldH = c->ldH;
  h.K = c->K;
  h.n = c->n;
  h.num = src_slots(c, fused_t ? PLAN_WXG : PLAN_WX);
  h.z = src_slots(c, PLAN_ZH);
  h.eps = static_cast<float>(c->eps);
  h.cov = tab;
  h.loss_type = c->loss_type;
  h.Kg = c->Kg;
  h.c_total = c_total;
  h.q_total = c->q_total;
  h.split_hi = c->Hsplit;  // B operand of H H^T below and of the next iteration's X H^T
  h.split_lo = c->Hsplit + static_cast<size_t>(c->K) * c->ldN;
  h.ld_split = c->ldN;
  h.hsum_partial = c->hsum_part;
  h.q_partial = c->q_part;
  h.pred_partial = c->pred_part;
  h.t1_partial = c->t1_part;
  AL_TRY(launch_h_update<true>(c, h, st));
  // ---- S = H H^T of the new H (Gram plan), statistics for the next iteration, loss terms of this one
  //      (main.py:666, 726-753)
  AL_TRY(run_gemm(c, PLAN_GRAM_H, nullptr, 0, st));
  HFinishParams hf{};
  hf.gram.src = src_slots(c, PLAN_GRAM_H);
  hf.gram.K = c->K;
  hf.gram.out = c->red_S();
  hf.gram.ld = c->K;
  if (single) {  // S stays this GPU's own: its split copies can be written right away
    hf.gram.split_hi = c->Ssplit;
    hf.gram.split_lo = c->Ssplit + static_cast<size_t>(c->K) * c->ldK;
    hf.gram.ld_split = static_cast<int>(c->ldK);
  }
  hf.gram_blocks = gram_blocks_for(c->K);
  hf.n_parts = c->upd_grid_h;
  hf.hsum_partial = c->hsum_part;
  hf.hsum = c->red_hsum();
  hf.q_partial = c->q_part;
  hf.stats_q = c->red_Q();
  hf.q_total = c->q_total;
  hf.pred_partial = c->pred_part;
  hf.n_cov = c->n_cov;
  hf.t1_partial = c->t1_part;
  hf.T = c->T;
  hf.ldT = c->K;
  hf.loss_row = c->loss_hist + static_cast<size_t>(iter) * (2 + c->n_cov);
  hf.counter = c->finish_counter;
  PDL_LAUNCH(h_finish_kernel, dim3(hf.gram_blocks + h_finish_stat_blocks(c->K, c->q_total)), dim3(256), 0, st, hf);
  return ALPINE_OK;
}

}  // namespace

extern "C" {

int alpine_mu_apply(alpine_ctx* c, int iter, void* stream) { return mu_apply_impl(c, iter, stream, false); }
int alpine_mu_apply_peer(alpine_ctx* c, int iter, void* stream) { return mu_apply_impl(c, iter, stream, true); }

int alpine_peer_export(alpine_ctx* c, void* handle_out) {
  if (c == nullptr || handle_out == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  if (c->ws_ready || c->reduce != nullptr)
    return fail(ALPINE_ERR_STATE, "alpine_peer_export must come before alpine_bind_reduce_buffer / alpine_fit_begin");
  DEVICE_SCOPE(c);
  CU_TRY(cudaMalloc(reinterpret_cast<void**>(&c->xchg), c->xchg_floats() * sizeof(float)));
  CU_TRY(cudaMemset(c->xchg, 0, c->xchg_floats() * sizeof(float)));
  CU_TRY(cudaDeviceSynchronize());  // the zero-fill is complete before any stream (or peer) touches the block
  AL_TRY(ws_alloc(c, &c->sum_small, static_cast<size_t>(c->small_floats())));
  AL_TRY(ws_alloc(c, &c->sum_P, static_cast<size_t>(c->K) * c->ldG));
  c->reduce = c->xchg;                          // [X H^T | H H^T | rowsum H | B statistics] partials of this rank
  c->WT = c->xchg + c->xchg_wt_off();           // peers store their gene slices of the new W^T here
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, c->xchg));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, sizeof(h));
  return ALPINE_OK;
}

int alpine_peer_import(alpine_ctx* c, int rank, int world, const void* handles) {
  if (c == nullptr || handles == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  if (c->xchg == nullptr) return fail(ALPINE_ERR_STATE, "alpine_peer_export has not been called");
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(ALPINE_ERR_ARG, "peer exchange needs 2..%d ranks (rank %d of %d)", kMaxPeers, rank, world);
  DEVICE_SCOPE(c);
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      c->peer_base[q] = c->xchg;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + 64 * q, sizeof(h));
    void* p = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[q] = static_cast<float*>(p);
    c->peer_opened[q] = true;
  }
  c->peer_rank = rank;
  c->peer_world = world;
  return ALPINE_OK;
}

int alpine_peer_disable(alpine_ctx* c) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  // back to the caller-side all-reduce (another rank could not map the peers): the exchange block, if any, keeps
  // serving as this rank's private reduce buffer / W^T storage
  c->peer_world = 0;
  c->peer_rank = -1;
  return ALPINE_OK;
}

int64_t alpine_reduce_stats_offset(const alpine_ctx* c) { return c ? static_cast<int64_t>(c->K) * c->ldG : 0; }

int alpine_als_block(alpine_ctx* c, int b, void* stream) {
  AL_TRY(check_bound(c, true));
  if (!c->fit_active) return fail(ALPINE_ERR_STATE, "alpine_fit_begin has not been called");
  if (b < 0 || b >= c->n_blocks) return fail(ALPINE_ERR_ARG, "block %d outside [0, %d)", b, c->n_blocks);
  if (c->peer_on()) return fail(ALPINE_ERR_STATE, "the block-wise sweep exchanges through the caller's all-reduce, not peer memory");
  if (c->n_groups > 1) return fail(ALPINE_ERR_ARG, "the block-wise sweep (use_als) supports at most %d components in total", kAccStride);
  DEVICE_SCOPE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int r0 = 0;
  for (int i = 0; i < b; ++i) r0 += c->kblk[i];
  const int r1 = r0 + c->kblk[b];
  if (c->xh_in_slots) {  // single GPU: alpine_mu_partials left the numerator in the contraction's slots
    AL_TRY(reduce_slots(c, PLAN_XH, c->red_Pt(), c->ldG, st));
    c->xh_in_slots = false;
  }
  // ---- W_b update (main.py:527-545): den = 2 W_cat (H_cat H_b^T) + (1-l1) alpha W_b + W_b orth(k_b) + l1 alpha
  SymLongParams w{};
  w.Sym = c->red_S();
  w.ldS = c->K;
  w.Mat = c->WT;
  w.ldM = c->ldG;
  w.K = c->K;
  w.r0 = r0, w.r1 = r1;
  w.L = c->G;
  w.Num = c->red_Pt();  // X H^T of the iteration's first sweep: H_b is still the H it was computed from
  w.ldNum = c->ldG;
  w.c1 = static_cast<float>((1.0 - c->l1) * c->alpha);
  w.c2 = static_cast<float>(c->l1 * c->alpha);
  w.orth = static_cast<float>(c->orth);
  w.eps = static_cast<float>(c->eps);
  w.split_hi = c->Wsplit;
  w.split_lo = c->Wsplit + static_cast<size_t>(c->K) * c->ldG;
  w.ld_split = c->ldG;
  AL_TRY(run_sym_long<EPI_W>(c, w, st));
  // ---- B_b update (main.py:548-562) from the statistics of (H_b, B_b), both unchanged since they were taken
  CovTable one;
  one.n_cov = 0;
  if (b < c->n_cov) {
    const CovTable tab = make_cov_table(c);
    one.n_cov = 1;
    one.d[0] = tab.d[b];
    b_update_kernel<<<1, 128, c->ccov[b] * c->kblk[b] * sizeof(float), st>>>(one, c->loss_type, c->red_Q(), c->red_hsum(),
                                                                             c->red_S(), c->K, (float)c->eps);
    LAUNCH_CHECK();
  }
  // ---- T = W^T W with the new W_b, A_b = W_b^T X (main.py:565-568): one sweep of X with a k_b-wide B operand
  AL_TRY(run_gemm(c, PLAN_GRAM_W, c->T, c->K, st));
  AL_TRY(run_gemm(c, PLAN_WX_BLOCK + b, c->A + static_cast<size_t>(r0) * c->ldN, c->ldN, st));
  if (b < c->n_cov) {
    const size_t smem = (static_cast<size_t>(c->ccov[b]) * c->kblk[b] + 3ull * c->kblk[b] * 128) * sizeof(float);
    if (smem > 200 * 1024) return fail(ALPINE_ERR_ARG, "covariate block too large for the guided-terms kernel");
    guided_terms_kernel<<<dim3(ceil_div(c->n, 128), 1), 128, smem, st>>>(one, c->loss_type, c->H, c->ldH, (int)c->n,
                                                                        (float)c->eps, c->numG, c->denG, c->ldN);
    LAUNCH_CHECK();
  }
  // ---- H_b update (main.py:570-588): den = 2 (W_b^T W_cat) H_cat + guided terms
  SymLongParams h{};
  h.Sym = c->T;
  h.ldS = c->K;
  h.Mat = c->H;
  h.ldM = c->ldH;
  h.K = c->K;
  h.r0 = r0, h.r1 = r1;
  h.L = c->n;
  h.Num = c->A;
  h.ldNum = c->ldN;
  h.numG = c->numG;
  h.denG = c->denG;
  h.ldD = c->ldN;
  h.Kg = c->Kg;
  h.eps = static_cast<float>(c->eps);
  h.t1_partial = nullptr;  // the loss terms are taken once all blocks are done (alpine_als_finish)
  h.rowsum_partial = c->hsum_partial;
  h.split_hi = c->Hsplit;
  h.split_lo = c->Hsplit + static_cast<size_t>(c->K) * c->ldN;
  h.ld_split = c->ldN;
  AL_TRY(run_sym_long<EPI_H>(c, h, st));
  // ---- the next block's W update needs H_cat H^T with this block's new rows (this shard's part; all-reduced by
  //      the caller under cell sharding).  The last block's is taken by alpine_als_finish.
  if (b + 1 < c->n_blocks) AL_TRY(run_gemm(c, PLAN_GRAM_H, c->red_S(), c->K, st));
  return ALPINE_OK;
}

int alpine_als_finish(alpine_ctx* c, int iter, void* stream) {
  AL_TRY(check_bound(c, true));
  if (!c->fit_active) return fail(ALPINE_ERR_STATE, "alpine_fit_begin has not been called");
  if (iter < 0 || iter >= c->loss_cap) return fail(ALPINE_ERR_ARG, "iteration %d outside [0, %d)", iter, c->loss_cap);
  DEVICE_SCOPE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  transpose_kernel<<<dim3(ceil_div(c->G, 32), ceil_div(c->K, 32)), dim3(32, 8), 0, st>>>(c->WT, c->ldG, c->K, (int)c->G,
                                                                                       c->W, c->ldW);
  LAUNCH_CHECK();
  c->w_stale = false;
  // every A_b was formed with its final W_b, every H_b is final: t1 = sum A .* H (main.py:666 via the trace identity)
  dot_partial_kernel<<<c->sl_blocks_n, 256, 0, st>>>(c->A, c->ldN, c->H, c->ldH, c->K, c->n, c->t1_partial);
  LAUNCH_CHECK();
  return run_stats(c, c->loss_hist + static_cast<size_t>(iter) * (2 + c->n_cov), true, st);
}

int alpine_eval_loss(alpine_ctx* c, double* terms, void* stream) {
  AL_TRY(check_bound(c, true));
  if (terms == nullptr) return fail(ALPINE_ERR_ARG, "null argument");
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(ws_alloc(c, &c->eval_row, static_cast<size_t>(2 + c->n_cov)));
  AL_TRY(export_w(c, st));
  transpose_kernel<<<dim3(ceil_div(c->K, 32), ceil_div(c->G, 32)), dim3(32, 8), 0, st>>>(c->W, c->ldW, (int)c->G, c->K,
                                                                                       c->WT, c->ldG);
  LAUNCH_CHECK();
  AL_TRY(run_split(c, c->WT, c->ldG, c->G, c->Wsplit, c->ldG, st));
  AL_TRY(run_gemm(c, PLAN_GRAM_W, c->T, c->K, st));  // T = W^T W
  AL_TRY(run_gemm(c, PLAN_WX, c->A, c->ldN, st));    // A = W^T X
  dot_partial_kernel<<<c->sl_blocks_n, 256, 0, st>>>(c->A, c->ldN, c->H, c->ldH, c->K, c->n, c->t1_partial);  // t1
  LAUNCH_CHECK();
  AL_TRY(run_stats(c, c->eval_row, false, st));      // S = H H^T, t2 = sum T .* S, prediction terms of (B, H)
  AL_TRY(run_split_small(c, c->red_S(), c->K, c->Ssplit, st));  // (keeps the context consistent for a following step)
  CU_TRY(cudaMemcpyAsync(terms, c->eval_row, sizeof(double) * (2 + c->n_cov), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  return check_kernel_error(c);
}

int alpine_fit_losses(alpine_ctx* c, int n_iter, double* xnorm2, double* rows, void* stream) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  if (!c->fit_active) return fail(ALPINE_ERR_STATE, "alpine_fit_begin has not been called");
  if (n_iter < 0 || n_iter > c->loss_cap) return fail(ALPINE_ERR_ARG, "n_iter outside the loss history");
  DEVICE_SCOPE(c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(export_w(c, st));  // the caller reads W next
  CU_TRY(cudaStreamSynchronize(st));
  AL_TRY(check_kernel_error(c));
  if (xnorm2) CU_TRY(cudaMemcpy(xnorm2, c->xnorm2, sizeof(double), cudaMemcpyDeviceToHost));
  if (rows && n_iter > 0)
    CU_TRY(cudaMemcpy(rows, c->loss_hist, sizeof(double) * n_iter * (2 + c->n_cov), cudaMemcpyDeviceToHost));
  return ALPINE_OK;
}

int alpine_scale(alpine_ctx* c, void* stream) {
  AL_TRY(check_bound(c, false));
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(export_w(c, st));
  transpose_kernel<<<dim3(ceil_div(c->K, 32), ceil_div(c->G, 32)), dim3(32, 8), 0, st>>>(c->W, c->ldW, (int)c->G, c->K,
                                                                                       c->WT, c->ldG);
  LAUNCH_CHECK();
  rowsum_kernel<<<c->K, 256, 0, st>>>(c->WT, c->ldG, c->G, c->colsum);  // s = W.sum(0), main.py:776
  LAUNCH_CHECK();
  scale_w_kernel<<<2 * c->num_sms, 256, 0, st>>>(c->W, c->ldW, (int)c->G, c->K, c->colsum);
  LAUNCH_CHECK();
  scale_h_kernel<<<4 * c->num_sms, 256, 0, st>>>(c->H, c->ldH, c->K, (int)c->n, c->colsum);
  LAUNCH_CHECK();
  if (c->n_cov > 0) {
    for (int i = 0; i < c->n_cov; ++i)
      if (c->B[i] == nullptr) return fail(ALPINE_ERR_STATE, "covariate %d has no B bound", i);
    scale_b_kernel<<<c->n_cov, 128, 0, st>>>(make_cov_table(c), c->colsum);
    LAUNCH_CHECK();
  }
  return ALPINE_OK;
}

int alpine_transform(alpine_ctx* c, int n_iter, void* stream) {
  AL_TRY(check_bound(c, false));
  if (n_iter < 0) return fail(ALPINE_ERR_ARG, "n_iter must be >= 0");
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(export_w(c, st));
  transpose_kernel<<<dim3(ceil_div(c->K, 32), ceil_div(c->G, 32)), dim3(32, 8), 0, st>>>(c->W, c->ldW, (int)c->G, c->K,
                                                                                       c->WT, c->ldG);
  LAUNCH_CHECK();
  AL_TRY(run_split(c, c->WT, c->ldG, c->G, c->Wsplit, c->ldG, st));
  AL_TRY(run_gemm(c, PLAN_GRAM_W, c->T, c->K, st));  // T = W^T W, loop-invariant
  AL_TRY(run_split_small(c, c->T, c->K, c->Tsplit, st));
  AL_TRY(run_gemm(c, PLAN_WX, c->A, c->ldN, st));    // A = W^T X, loop-invariant (main.py:706)
  HUpdParams h{};
  h.H = c->H;
  h.ldH = c->ldH;
  h.K = c->K;
  h.n = c->n;
  h.num = src_direct(c->A, c->ldN);
  h.eps = static_cast<float>(c->eps);
  for (int it = 0; it < n_iter; ++it) {  // main.py:705-709
    AL_TRY(run_gemm(c, PLAN_ZH, nullptr, 0, st));  // Z = T H
    h.z = src_slots(c, PLAN_ZH);
    AL_TRY(launch_h_update<false>(c, h, st));
  }
  CU_TRY(cudaStreamSynchronize(st));
  return check_kernel_error(c);
}

int alpine_sync_w(alpine_ctx* c, void* stream) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  if (c->W == nullptr) return fail(ALPINE_ERR_STATE, "alpine_bind_factors has not been called");
  DEVICE_SCOPE(c);
  return export_w(c, static_cast<cudaStream_t>(stream));
}

int alpine_xh_product(alpine_ctx* c, float* out, int64_t ld_out, void* stream) {
  AL_TRY(check_bound(c, false));
  if (out == nullptr || ld_out < c->G) return fail(ALPINE_ERR_ARG, "bad output");
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(run_split(c, c->H, c->ldH, c->n, c->Hsplit, c->ldN, st));
  AL_TRY(run_gemm(c, PLAN_XH, out, ld_out, st));
  CU_TRY(cudaStreamSynchronize(st));
  return check_kernel_error(c);
}

int alpine_wx_product(alpine_ctx* c, float* out, int64_t ld_out, void* stream) {
  AL_TRY(check_bound(c, false));
  if (out == nullptr || ld_out < c->n) return fail(ALPINE_ERR_ARG, "bad output");
  DEVICE_SCOPE(c);
  AL_TRY(ensure_workspace(c, static_cast<cudaStream_t>(stream)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AL_TRY(export_w(c, st));
  transpose_kernel<<<dim3(ceil_div(c->K, 32), ceil_div(c->G, 32)), dim3(32, 8), 0, st>>>(c->W, c->ldW, (int)c->G, c->K,
                                                                                       c->WT, c->ldG);
  LAUNCH_CHECK();
  AL_TRY(run_split(c, c->WT, c->ldG, c->G, c->Wsplit, c->ldG, st));
  AL_TRY(run_gemm(c, PLAN_WX, out, ld_out, st));
  CU_TRY(cudaStreamSynchronize(st));
  return check_kernel_error(c);
}

int alpine_upload_rows(int device, float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows,
                       int64_t cols, int threads, void* stream) {
  if (rows < 0 || cols < 0 || (rows > 0 && cols > 0 && (dst == nullptr || src == nullptr)))
    return fail(ALPINE_ERR_ARG, "null / negative argument");
  if (ld_dst < cols || ld_src < cols) return fail(ALPINE_ERR_ARG, "leading dimension smaller than the row length");
  DeviceScope scope(device);
  if (!scope.ok) return fail(ALPINE_ERR_CUDA, "cannot select device %d", device);
  const cudaError_t e = upload_rows_f32(device, dst, ld_dst, src, ld_src, rows, cols, threads, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(ALPINE_ERR_CUDA, "host -> device upload failed: %s", cudaGetErrorString(e));
  return ALPINE_OK;
}

int alpine_profile(alpine_ctx* c, int enable) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  c->prof = enable != 0;
  if (enable) c->prof_used = 0;
  return ALPINE_OK;
}

int alpine_profile_read(alpine_ctx* c, double* gemm_ms_total, long long* gemm_launches) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  DEVICE_SCOPE(c);
  double total = 0.0;
  const size_t pairs = c->prof_used / 2;
  for (size_t i = 0; i < pairs; ++i) {
    CU_TRY(cudaEventSynchronize(c->prof_ev[2 * i + 1]));
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
    total += ms;
  }
  if (gemm_ms_total) *gemm_ms_total = total;
  if (gemm_launches) *gemm_launches = static_cast<long long>(pairs);
  return ALPINE_OK;
}

int alpine_query(const alpine_ctx* c, int* num_sms, int* gemm_grid, int* smem_stages, int* k_padded) {
  if (c == nullptr) return fail(ALPINE_ERR_ARG, "null context");
  if (num_sms) *num_sms = c->num_sms;
  if (gemm_grid) *gemm_grid = c->plans[PLAN_XH][0].valid ? c->plans[PLAN_XH][0].grid : 0;
  if (smem_stages) *smem_stages = c->plans[PLAN_XH][0].valid ? c->plans[PLAN_XH][0].p.sx : 0;
  if (k_padded) *k_padded = c->Kp;
  return ALPINE_OK;
}

}  // extern "C"
