// The two dense contractions of ALPINE's multiplicative-update step over the gene x cell matrix X, as one
// warp-specialised, persistent, stream-K tcgen05 kernel for sm_100a (split-precision products, fp32 accumulate):
//
//   ORIENT_XH:  D[g][k] = sum_j X[g][j] * H[k][j]     (reference main.py:596,  (2*X_batch) @ H^T)
//   ORIENT_WX:  D[j][k] = sum_g X[g][j] * W[g][k]     (reference main.py:653,  (2*W^T) @ X_batch)
//
// X lives in HBM exactly as the reference holds it: cells-major, X_phys[j][g] (main.py:104, 445).  Both
// contractions stream that single copy once through TMA; the 128-row (M) side of the MMA is genes (XH) or
// cells (WX), the N side is the K components (padded to a multiple of 16), the reduction runs over the other
// axis of X in blocks of 32.
//
//  * Precision: every fp32 value is split into hi = tf32(x) and the exact remainder lo = x - hi (|lo| <= 2^-11 |x|);
//    x * y = hi * hi' + hi * lo' + lo * hi' + O(2^-22).  The main term runs as tf32 MMAs, the two correction terms
//    as bf16 MMAs (kind::f16) on the round-to-nearest bf16 images of hi and lo, at twice the tf32 rate: 2 instead of 3
//    tf32-MMA equivalents per product, which moves the dense sweep from the tensor roofline to the HBM roofline.
//    The bf16 rounding of a correction operand is unbiased and costs ~2^-9 of a term that is itself 2^-11 of the
//    product; over the long reductions of this workload it averages out like the 2^-22 of plain 3xTF32 (the golden
//    trajectories deviate from the reference by the same amount with either scheme, tests/test_split_precision.py).
//  * A operand (X super-tile, 256 rows x 32 reduction elements, fp32): TMA -> shared memory ring -> 8
//    converter warps split every value and write tf32 hi (32 columns per MMA tile), bf16(hi) and bf16(lo) (two
//    reduction elements per 32-bit column, 16 columns each) to TENSOR MEMORY (tcgen05.st), so the tensor core reads
//    A from TMEM and the memory orientation of X does not matter (XH reads the tile transposed out of shared
//    memory, WX reads 128B-swizzled rows).
//  * B operand (H or W^T, [K][R] K-major, pre-split by the kernels that produce it -- ptx::store_split4: a tf32 hi
//    plane and a 16-bit plane with the bf16 images of hi and lo; 90 MB at most, L2 resident): three TMA loads
//    (SWIZZLE_128B for the fp32 tile, SWIZZLE_64B for the two bf16 tiles, rows >= K zero-filled) -> shared memory
//    ring -> UMMA shared-memory descriptors.  One B stage serves 256 rows of X, which keeps L2->SM traffic of the
//    small operand below that of X itself.
//  * D accumulates in TMEM (fp32, 2 accumulators of Kp columns): per 32-deep k-block two 16-deep bf16 steps of
//    lo*hi' and hi*lo' (small terms first), then four 8-deep tf32 steps of hi*hi'.  The tensor core accumulates with
//    round-toward-zero, so after every `chunk` k-blocks the accumulator is drained into an fp32 master sum held in
//    the registers of the epilogue threads (round-to-nearest adds); the drain of one tile overlaps the MMAs of the
//    other.
//  * Work split (stream-K inside pieces): the reduction axis is cut into pieces whose B-operand window fits L2;
//    the (super-tile, k-block) units of every piece are cut into gridDim.x equal contiguous ranges, so all CTAs
//    move through the pieces together and the small operand is read from HBM about once.  Every contiguous
//    run inside one tile ("segment") is stored to a partial-sum slot; the consumers add the slots of a
//    tile in a fixed order, so results are deterministic and no CTA ever waits for another.
//  * One MMA-issuing warp per 128-row tile: a 128x112x8 tf32 MMA lasts only ~56 cycles, so a single issuing
//    thread (and any per-instruction register shuffling) would leave the tensor pipe idle.
//  * Every mbarrier wait is bounded (clock64): a protocol bug reports an error code instead of hanging the GPU.
//  * SRC_TILES (sparse X): instead of TMA, the X-producer warp scatters per-(super-tile, k-block) nonzero lists
//    (csr_tiles.cuh, built once from the CSR matrix) into a zeroed shared-memory tile, so HBM only sees 8 bytes
//    per nonzero while everything downstream (split, TMEM staging, MMAs, drains) is unchanged.  The tile is
//    row-major + swizzled for both orientations (the lists carry ready-made offsets), and every converter
//    thread clears the 128 bytes it has just read, so a stage returns to the producer already zeroed: no
//    zero-fill traffic through L2 and no single-warp memset.
//  * Count matrices (EXACT): when every X value is exactly representable in tf32 (integers < 2048, detected by
//    the pass that computes ||X||^2), the lo half of A is zero: the converters write hi only and the one remaining
//    correction term hi*lo' stays a tf32 MMA on a tf32 copy of lo' (two tf32 MMAs per k-step; the sparse path is
//    bound by the converters, so the bf16 packing would cost it more than the cheaper MMA saves -- measured).
//  * Measured (round 2, probe on one B200): the sweep is no longer bound by one resource.  Dropping the bf16 MMAs
//    gives -13 %, dropping all converter arithmetic -8 %, halving / removing the B-operand loads -2 % / -6.5 % (so a
//    cta_group::2 variant, whose point is to halve them, would gain ~2 %), deeper X rings nothing.
#pragma once
#include <vector>

#include "ptx_sm100.cuh"

namespace alpine {

constexpr int kBM = 128;         // rows per MMA tile == TMEM lanes
constexpr int kBK = 32;          // reduction elements per pipeline stage (128 bytes of fp32)
constexpr int kUmmaK = 8;        // tf32: 32 bytes per MMA k-step
constexpr int kUmmaK16 = 16;     // bf16: 32 bytes per MMA k-step
constexpr int kTmemCols = 512;
constexpr int kTmemAOff = 256;   // A staging starts here; accumulators occupy [0, 256)
constexpr int kMaxXStages = 8;
constexpr int kMaxBStages = 6;
constexpr int kMaxAStages = 4;
constexpr long long kTimeoutCycles = 1ll << 30;  // ~0.5 s: a wait that long is a bug, never a hang

enum { ORIENT_XH = 0, ORIENT_WX = 1 };
enum { SRC_DENSE = 0, SRC_TILES = 1 };

// Byte offset of element (row of the 256-row super-tile, reduction column of the 32-deep k-block) inside the
// shared-memory X tile, i.e. where the TMA box of the dense path puts it (and where the converters read it).
// The tile lists of the sparse path use the ORIENT_WX form for both contractions.
__host__ __device__ inline uint32_t x_tile_offset(int orient, int row, int col) {
  if (orient == ORIENT_XH) return static_cast<uint32_t>(col * 256 + row) * 4u;  // [32][256], no swizzle
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>((col >> 2) ^ (row & 7)) << 4) +
         (static_cast<uint32_t>(col & 3) << 2);                                   // [256][32], 128B swizzle
}

// error codes written to GemmParams::err[0]; err[1..4] = blockIdx, threadIdx, k-block counter, aux
enum { ERR_NONE = 0, ERR_XPROD_EMPTY = 1, ERR_BPROD_EMPTY = 2, ERR_CONV_XFULL = 3, ERR_CONV_BFULL = 4,
       ERR_CONV_AEMPTY = 5, ERR_MMA_CFULL = 6, ERR_MMA_ACCEMPTY = 7, ERR_EPI_ACCFULL = 8, ERR_MMA_BFULL = 9 };

__host__ __device__ inline long long gemm_range_begin(long long total, int grid, int cta) {
  return total * cta / grid;
}

// Work space of one contraction.  The reduction axis (kb_per_tile k-blocks) is cut into `pieces` of piece_len
// k-blocks (the last may be shorter); a "run" is one (piece, tile) pair and the linear work order is piece-major:
//   pos -> piece = pos / (num_tiles * piece_len), then tile, then k-block inside the piece.
// Stream-K hands each CTA a contiguous range of that order, so at any moment all CTAs read the same one or two
// pieces of the small (B) operand: its live window stays a few MB and is served from L2 instead of HBM.
struct WorkSpace {
  int num_tiles, kb_per_tile, piece_len, pieces;
  __host__ __device__ long long total() const { return static_cast<long long>(num_tiles) * kb_per_tile; }
  __host__ __device__ int len_of_piece(int pc) const {
    const int rest = kb_per_tile - pc * piece_len;
    return rest < piece_len ? rest : piece_len;
  }
  // run index (piece * num_tiles + tile), first k-block and k-blocks left in the run at linear position pos
  __host__ __device__ void decode(long long pos, int& run, int& tile, int& kb, int& left_in_run) const {
    const long long per_piece = static_cast<long long>(num_tiles) * piece_len;
    const int pc = static_cast<int>(pos / per_piece);
    const long long rem = pos - pc * per_piece;
    const int lp = len_of_piece(pc);
    tile = static_cast<int>(rem / lp);
    const int off = static_cast<int>(rem - static_cast<long long>(tile) * lp);
    kb = pc * piece_len + off;
    left_in_run = lp - off;
    run = pc * num_tiles + tile;
  }
  __host__ __device__ long long run_begin(int pc, int tile) const {
    return static_cast<long long>(pc) * num_tiles * piece_len + static_cast<long long>(tile) * len_of_piece(pc);
  }
  // Share of CTA `cta` in piece pc: the piece's (tile, k-block) units are cut into `grid` equal contiguous ranges,
  // so all CTAs walk through the pieces together and only one piece of the B operand is live at a time.
  __host__ __device__ void cta_range(int pc, int grid, int cta, long long& b, long long& e) const {
    const long long units = static_cast<long long>(num_tiles) * len_of_piece(pc);
    const long long base = static_cast<long long>(pc) * num_tiles * piece_len;
#ifdef ALPINE_B200_GLOBAL_SPLIT  // A/B build only: one contiguous range per CTA over the whole piece-major order
    const long long gb = total() * cta / grid, ge = total() * (cta + 1) / grid;
    b = gb > base ? gb : base;
    e = ge < base + units ? ge : base + units;
    if (e < b) e = b;
#else
    b = base + units * cta / grid;
    e = base + units * (cta + 1) / grid;
#endif
  }
  // number of segments (contiguous runs inside one tile) CTA `cta` executes in piece pc
  __host__ __device__ int cta_segments(int pc, int grid, int cta) const {
    long long b, e;
    cta_range(pc, grid, cta, b, e);
    if (e <= b) return 0;
    const long long base = static_cast<long long>(pc) * num_tiles * piece_len;
    const int lp = len_of_piece(pc);
    return static_cast<int>((e - 1 - base) / lp - (b - base) / lp) + 1;
  }
};

struct GemmParams {
  int M;            // rows of D (genes for XH, cells for WX)
  int R;            // reduction length
  int K;            // real component count
  int Kp;           // MMA N: K padded to a multiple of 16 (<= 128)
  WorkSpace ws;     // num_tiles = ceil(M / 256), kb_per_tile = ceil(R / 32), piece decomposition
  int sx;           // X ring depth
  int sb;           // B ring depth
  int max_segs;     // partial slots per CTA
  int extra_tile;   // ORIENT_WX only: index of one additional super-tile whose rows come from a second matrix with the
                    // same reduction axis (tmX2; its first 256 rows), or -1.  Used to get W^T W out of the W^T X launch:
                    // W^T [K][G] has exactly the layout of a 256-row block of X (rows x genes).
  int extra_chunk_log2;  // flush interval (log2 k-blocks) of the additional super-tile
  int chunk_log2;   // k-blocks (2^chunk_log2) accumulated in TMEM between two round-to-nearest flushes: every MMA
                    // adds into the accumulator with round-toward-zero, so shorter chains mean less bias (the small
                    // K x K-deep plans use 1 or 2 k-blocks: their flushes cost nothing next to their launch)
  float* partial;   // [gridDim.x * max_segs][K][256]
  int* err;         // [8]
  // SRC_TILES: nonzeros of X grouped by (super-tile, k-block); entry = {x_tile_offset, fp32 bits}
  const long long* sp_ofs;  // [num_tiles * kb_per_tile + 1]; nullptr => dense X through TMA
  const uint2* sp_ent;
  long long sp_total;       // entries in sp_ent
};

// Walks the k-blocks of a CTA's stream-K range in execution order.
struct KbIter {
  WorkSpace ws;
  long long pos, end;
  int tile, kb, left, pc, grid, cta;
  __device__ KbIter(const WorkSpace& w, int g, int c) : ws(w), pos(0), end(0), tile(0), kb(0), left(0), pc(-1), grid(g), cta(c) {}
  __device__ bool next(long long& blk) {
    if (left == 0) {
      while (pos >= end) {
        if (++pc >= ws.pieces) return false;
        ws.cta_range(pc, grid, cta, pos, end);
      }
      int run;
      ws.decode(pos, run, tile, kb, left);
      if (left > end - pos) left = static_cast<int>(end - pos);
    }
    blk = static_cast<long long>(tile) * ws.kb_per_tile + kb;
    ++kb, --left, ++pos;
    return true;
  }
};


struct AbortCtx {
  volatile int* flag;  // shared
  int* err;            // global
};

__device__ __noinline__ void report_timeout(const AbortCtx& a, int code, int aux0, int aux1) {
  if (atomicCAS(a.err, 0, code) == 0) {
    a.err[1] = blockIdx.x;
    a.err[2] = threadIdx.x;
    a.err[3] = aux0;
    a.err[4] = aux1;
    __threadfence();
  }
  *a.flag = 1;
}

// Bounded mbarrier wait.  try_wait suspends in hardware; the clock and the abort flag are looked at only every
// 256 polls.  The slow path is out of line so that the hot loops stay small.  Safe to call from one lane or from
// a whole warp.
__device__ __forceinline__ bool wait_bar_slow(uint64_t* bar, uint32_t parity, volatile int* flag, int* err, int code,
                                           int aux0, int aux1) {
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (true) {
    if (ptx::mbar_try_wait(bar, parity)) return true;
    if ((++polls & 255u) == 0) {
      if (*flag) return false;
      if (clock64() - t0 > kTimeoutCycles) {
        report_timeout(AbortCtx{flag, err}, code, aux0, aux1);
        return false;
      }
    }
  }
}
__device__ __forceinline__ bool wait_bar(uint64_t* bar, uint32_t parity, const AbortCtx& a, int code, int aux0,
                                         int aux1) {
  if (ptx::mbar_try_wait(bar, parity)) return true;
  return wait_bar_slow(bar, parity, a.flag, a.err, code, aux0, aux1);
}
// warp-uniform variant: every lane polls (the instruction is warp-wide anyway); the verdict is made uniform
__device__ __forceinline__ bool warp_wait_bar(uint64_t* bar, uint32_t parity, const AbortCtx& a, int code, int aux0,
                                              int aux1) {
  const bool ok = wait_bar(bar, parity, a, code, aux0, aux1);
  return __all_sync(0xffffffffu, ok);
}

// Position in a ring of n slots without integer division: slot index and parity of the completed wraps.
struct RingPos {
  int s = 0;
  uint32_t ph = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++s == n) {
      s = 0;
      ph ^= 1u;
    }
  }
};
constexpr int kChunkLog2 = 3;  // default: 8 k-blocks accumulated in TMEM between two round-to-nearest flushes

// Fixed geometry: a super-tile is 2 MMA tiles (256 rows of X); 8 converter/epilogue warps, one thread per row.
constexpr int kMT = 2;
constexpr int kRows = kMT * kBM;
constexpr int kConvWarps = 4 * kMT;
constexpr int kGemmThreads = (kConvWarps + 4) * 32;  // + X producer, B producer, one MMA issuer per tile
constexpr int kXTileBytes = kRows * kBK * 4;
constexpr int kAStageCols = kMT * 64;   // per MMA tile: 32 tf32 hi columns + 16 bf16(hi) + 16 bf16(lo) columns
constexpr int kAStages = (kTmemCols - kTmemAOff) / kAStageCols;
constexpr int kAccStride = kTmemAOff / kMT;  // column distance between the two accumulators (=> Kp <= 128)

struct GemmSmemLayout {
  size_t x_off, b_off, bar_off, total;
};
__host__ __device__ inline GemmSmemLayout gemm_smem_layout(int Kp, int sx, int sb) {
  GemmSmemLayout l;
  l.x_off = 0;
  l.b_off = static_cast<size_t>(sx) * kXTileBytes;
  l.bar_off = l.b_off + static_cast<size_t>(sb) * 2 * Kp * kBK * 4;
  l.total = l.bar_off + (2 * kMaxXStages + 2 * kMaxBStages + 2 * kMaxAStages + 4) * 8 + 16;
  return l;
}

// NC = Kp / 16: number of 16-column groups of the accumulator (compile time: the fp32 master sum of a thread's
// accumulator row lives in 16*NC registers).
template <int ORIENT, int NC, bool EXACT>
__global__ void __launch_bounds__(kGemmThreads, 1)
mu_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBh16, const __grid_constant__ CUtensorMap tmBl16,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmX2,
               const GemmParams p) {
  constexpr int Kp = 16 * NC;
  constexpr int kWarpXProd = kConvWarps, kWarpBProd = kConvWarps + 1, kWarpMma = kConvWarps + 2;  // + kMT MMA warps
  // one B stage: tf32 hi tile [Kp][32] (128-byte rows, SWIZZLE_128B), then the bf16 images of hi and of lo
  // [Kp][32] each (64-byte rows, SWIZZLE_64B) -- or, for the count-matrix variant (EXACT), the tf32 lo tile
  constexpr int b_hi_bytes = Kp * kBK * 4;
  constexpr int b_16_bytes = Kp * kBK * 2;
  constexpr int b_stage_bytes = b_hi_bytes + 2 * b_16_bytes;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  ptx::pdl_launch_dependents();  // the set-up below (barriers, TMEM) overlaps the predecessor's tail

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int SX = p.sx, SB = p.sb;
  const GemmSmemLayout lay = gemm_smem_layout(Kp, SX, SB);
  uint8_t* smem_x = smem + lay.x_off;
  uint8_t* smem_b = smem + lay.b_off;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* xfull_bar = bars;                           // [SX]  TMA -> converters
  uint64_t* xempty_bar = xfull_bar + kMaxXStages;       // [SX]  converters -> X producer
  uint64_t* bfull_bar = xempty_bar + kMaxXStages;       // [SB]  TMA -> converters
  uint64_t* bempty_bar = bfull_bar + kMaxBStages;       // [SB]  MMA commit -> B producer
  uint64_t* cfull_bar = bempty_bar + kMaxBStages;       // [kAStages]  converters -> MMA (A in TMEM + B split)
  uint64_t* aempty_bar = cfull_bar + kMaxAStages;       // [kAStages]  MMA commit -> converters
  uint64_t* accfull_bar = aempty_bar + kMaxAStages;     // [kMT]  MMA commit (end of a chunk) -> flush
  uint64_t* accempty_bar = accfull_bar + kMT;           // [kMT]  flush -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty_bar + kMT);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  AbortCtx actx{abort_flag, p.err};

  if (threadIdx.x == 0) {
    for (int i = 0; i < SX; ++i) {
      ptx::mbar_init(&xfull_bar[i], 1);
      ptx::mbar_init(&xempty_bar[i], kConvWarps);
    }
    for (int i = 0; i < SB; ++i) {
      ptx::mbar_init(&bfull_bar[i], 1);
      ptx::mbar_init(&bempty_bar[i], kMT);
    }
    for (int i = 0; i < kAStages; ++i) {
      ptx::mbar_init(&cfull_bar[i], kConvWarps);
      ptx::mbar_init(&aempty_bar[i], kMT);
    }
    for (int i = 0; i < kMT; ++i) {
      ptx::mbar_init(&accfull_bar[i], 1);
      ptx::mbar_init(&accempty_bar[i], 4);
    }
    *abort_flag = 0;
    ptx::fence_barrier_init();
  }
  const bool tiles = p.sp_ofs != nullptr;  // SRC_TILES
  if (!tiles && warp == kWarpXProd && lane == 0) {
    ptx::prefetch_tensormap(&tmX);
    if (ORIENT == ORIENT_WX && p.extra_tile >= 0) ptx::prefetch_tensormap(&tmX2);
  }
  if (tiles) {  // the sparse producer only scatters nonzeros: stages start zeroed and are re-zeroed by their readers
    const uint32_t base = ptx::smem_u32(smem_x);
    for (int i = threadIdx.x; i < SX * (kXTileBytes / 16); i += kGemmThreads) ptx::sts_zero_v4(base + i * 16);
  }
  if (warp == kWarpBProd && lane == 0) {
    ptx::prefetch_tensormap(&tmBhi);
    if (EXACT) {
      ptx::prefetch_tensormap(&tmBlo);
    } else {
      ptx::prefetch_tensormap(&tmBh16);
      ptx::prefetch_tensormap(&tmBl16);
    }
  }
  if (warp == kWarpMma) {  // the first MMA warp owns the TMEM allocation
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::pdl_wait();  // from here on global memory written by the predecessor is read
  const uint32_t tmem_base = *tmem_slot;
  const WorkSpace ws = p.ws;
  const int cta = blockIdx.x;

  if (warp == kWarpXProd && tiles) {
    // ===================================================== sparse producer of the X ring (whole warp)
    // Per k-block: scatter the block's nonzeros into a stage its readers left zeroed.  Entry loads run one k-block
    // ahead (registers), their offsets two k-blocks ahead, so no global-memory latency sits on the critical path.
    constexpr int NE = 16;  // entries per lane kept in registers: 512 per k-block (mean at 5 % density: 410)
    const uint32_t smem_x_u32 = ptx::smem_u32(smem_x);
    KbIter ahead(ws, gridDim.x, cta);
    long long my_units = 0;
    for (int pc = 0; pc < ws.pieces; ++pc) {
      long long b, e;
      ws.cta_range(pc, gridDim.x, cta, b, e);
      my_units += e - b;
    }
    long long blk = 0, o_beg = 0, o_end = 0, cur_beg = 0;
    int cur_cnt = 0;
    uint2 cur[NE], nxt[NE];
    auto load_entries = [&](long long beg, int cnt, uint2(&e)[NE]) {
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int idx = i * 32 + lane;
        e[i] = (idx < cnt) ? __ldcs(p.sp_ent + beg + idx) : make_uint2(0u, 0u);
      }
    };
    bool have_next = ahead.next(blk);
    if (have_next) {
      cur_beg = __ldg(p.sp_ofs + blk);
      cur_cnt = static_cast<int>(__ldg(p.sp_ofs + blk + 1) - cur_beg);
      load_entries(cur_beg, cur_cnt, cur);
      have_next = ahead.next(blk);
      if (have_next) {
        o_beg = __ldg(p.sp_ofs + blk);
        o_end = __ldg(p.sp_ofs + blk + 1);
      }
    }
    bool ok = true;
    RingPos rx;
    for (uint32_t it = 0; it < static_cast<uint32_t>(my_units) && ok; ++it, rx.advance(SX)) {
      const int s = rx.s;
      // entries of the next k-block, offsets of the one after
      const long long nxt_beg = o_beg;
      const int nxt_cnt = have_next ? static_cast<int>(o_end - o_beg) : 0;
      load_entries(nxt_beg, nxt_cnt, nxt);
      {
        // The lists of consecutive k-blocks of a tile are adjacent in memory: pull the lines about four k-blocks
        // ahead into L2, so that the register prefetch above sees L2 latency instead of HBM latency.
        const long long pf = nxt_beg + 4ll * nxt_cnt + static_cast<long long>(lane) * ((nxt_cnt >> 5) + 1);
        if (nxt_cnt > 0 && pf < p.sp_total) ptx::prefetch_l2(p.sp_ent + pf);
      }
      if (have_next) {
        have_next = ahead.next(blk);
        if (have_next) {
          o_beg = __ldg(p.sp_ofs + blk);
          o_end = __ldg(p.sp_ofs + blk + 1);
        }
      }
      if (!warp_wait_bar(&xempty_bar[s], rx.ph ^ 1u, actx, ERR_XPROD_EMPTY, it, s)) {
        ok = false;
        break;
      }
      const uint32_t sX = smem_x_u32 + static_cast<uint32_t>(s) * kXTileBytes;
#pragma unroll
      for (int i = 0; i < NE; ++i)
        if (i * 32 + lane < cur_cnt) ptx::sts_f32(sX + cur[i].x, cur[i].y);
      for (int idx = NE * 32 + lane; idx < cur_cnt; idx += 32) {  // rare: a k-block denser than 512 nonzeros
        const uint2 e = __ldcs(p.sp_ent + cur_beg + idx);
        ptx::sts_f32(sX + e.x, e.y);
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&xfull_bar[s]);
#pragma unroll
      for (int i = 0; i < NE; ++i) cur[i] = nxt[i];
      cur_beg = nxt_beg;
      cur_cnt = nxt_cnt;
    }
    __syncwarp();
  } else if (warp == kWarpXProd) {
    // ===================================================== TMA producer of the X ring
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      RingPos rx;
      for (int pc = 0; pc < ws.pieces && ok; ++pc) {
      long long range_begin, range_end;
      ws.cta_range(pc, gridDim.x, cta, range_begin, range_end);
      for (long long pos = range_begin; pos < range_end && ok;) {
        int run, tile, kb0, len;
        ws.decode(pos, run, tile, kb0, len);
        if (len > range_end - pos) len = static_cast<int>(range_end - pos);
        for (int kb = kb0; kb < kb0 + len; ++kb, ++it, rx.advance(SX)) {
          const int s = rx.s;
          if (!wait_bar(&xempty_bar[s], rx.ph ^ 1u, actx, ERR_XPROD_EMPTY, it, s)) {
            ok = false;
            break;
          }
          uint8_t* dst = smem_x + static_cast<size_t>(s) * kXTileBytes;
          ptx::mbar_arrive_expect_tx(&xfull_bar[s], kXTileBytes);
          if (ORIENT == ORIENT_XH)  // box {256 genes, 32 cells} at (gene0, cell0): smem [32 cells][256 genes]
            ptx::tma_load_2d(dst, &tmX, &xfull_bar[s], tile * kRows, kb * kBK, ptx::kEvictFirst);
          else if (tile == p.extra_tile)  // the additional super-tile: rows [0, 256) of the second matrix
            ptx::tma_load_2d(dst, &tmX2, &xfull_bar[s], kb * kBK, 0, ptx::kEvictLast);
          else  // box {32 genes, 256 cells} at (gene0, cell0): smem [256 cells][32 genes], 128B swizzle
            ptx::tma_load_2d(dst, &tmX, &xfull_bar[s], kb * kBK, tile * kRows, ptx::kEvictFirst);
        }
        pos += len;
      }
      }
    }
    __syncwarp();
  } else if (warp == kWarpBProd) {
    // ===================================================== TMA producer of the B ring
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      RingPos rb;
      for (int pc = 0; pc < ws.pieces && ok; ++pc) {
      long long range_begin, range_end;
      ws.cta_range(pc, gridDim.x, cta, range_begin, range_end);
      for (long long pos = range_begin; pos < range_end && ok;) {
        int run, tile, kb0, len;
        ws.decode(pos, run, tile, kb0, len);
        if (len > range_end - pos) len = static_cast<int>(range_end - pos);
        for (int kb = kb0; kb < kb0 + len; ++kb, ++it, rb.advance(SB)) {
          const int s = rb.s;
          if (!wait_bar(&bempty_bar[s], rb.ph ^ 1u, actx, ERR_BPROD_EMPTY, it, s)) {
            ok = false;
            break;
          }
          uint8_t* dst = smem_b + static_cast<size_t>(s) * b_stage_bytes;
          ptx::mbar_arrive_expect_tx(&bfull_bar[s], b_stage_bytes);
          ptx::tma_load_2d(dst, &tmBhi, &bfull_bar[s], kb * kBK, 0, ptx::kEvictLast);
          if (EXACT) {
            ptx::tma_load_2d(dst + b_hi_bytes, &tmBlo, &bfull_bar[s], kb * kBK, 0, ptx::kEvictLast);
          } else {
            ptx::tma_load_2d(dst + b_hi_bytes, &tmBh16, &bfull_bar[s], kb * kBK, 0, ptx::kEvictLast);
            ptx::tma_load_2d(dst + b_hi_bytes + b_16_bytes, &tmBl16, &bfull_bar[s], kb * kBK, 0, ptx::kEvictLast);
          }
        }
        pos += len;
      }
      }
    }
    __syncwarp();
  } else if (warp >= kWarpMma) {
    // ===================================================== MMA issuers: one warp per 128-row tile.
    // The whole warp runs the loop with warp-uniform values; one elected lane issues (see ptx::mma_tf32_ts_if).
    const int mt = __shfl_sync(0xffffffffu, warp - kWarpMma, 0);
    const uint32_t leader = ptx::elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_b_u32 = __shfl_sync(0xffffffffu, ptx::smem_u32(smem_b), 0);
    const uint32_t idesc = ptx::make_idesc_tf32(kBM, Kp);
    const uint32_t idesc16 = ptx::make_idesc_bf16(kBM, Kp);
    const uint32_t d_acc = tb + mt * kAccStride;
    uint32_t it = 0, mc = 0;  // k-block counter, chunk counter
    bool ok = true;
    RingPos rb;
    for (int pc = 0; pc < ws.pieces && ok; ++pc) {
    long long range_begin, range_end;
    ws.cta_range(pc, gridDim.x, cta, range_begin, range_end);
    for (long long pos = range_begin; pos < range_end && ok;) {
      int run, tile, kb0, len;
      ws.decode(pos, run, tile, kb0, len);
      if (len > range_end - pos) len = static_cast<int>(range_end - pos);
      const int CM = (1 << ((ORIENT == ORIENT_WX && tile == p.extra_tile) ? p.extra_chunk_log2 : p.chunk_log2)) - 1;
      for (int li = 0; li < len; ++li, ++it, rb.advance(SB)) {
        const int t = it % kAStages;
        const int sbi = rb.s;
        const bool c_first = (li & CM) == 0;
        const bool c_last = (li & CM) == CM || li == len - 1;
        if (!warp_wait_bar(&bfull_bar[sbi], rb.ph, actx, ERR_MMA_BFULL, it, sbi) ||
            !warp_wait_bar(&cfull_bar[t], (it / kAStages) & 1, actx, ERR_MMA_CFULL, it, t)) {
          ok = false;
          break;
        }
        if (c_first && mc > 0) {  // the previous chunk of this accumulator must have been flushed
          if (!warp_wait_bar(&accempty_bar[mt], (mc - 1) & 1, actx, ERR_MMA_ACCEMPTY, mc, mt)) {
            ok = false;
            break;
          }
        }
        ptx::tc_fence_after();
        const uint32_t sb_addr = smem_b_u32 + static_cast<uint32_t>(sbi) * b_stage_bytes;
        const uint64_t dhi = ptx::make_kmajor_sw128_desc(sb_addr);
        const uint32_t a_hi = tb + kTmemAOff + t * kAStageCols + mt * 64;
        if (EXACT) {
          // count matrix: A has no lo half; hi * lo' stays a tf32 MMA on the tf32 lo tile (small term first)
          const uint64_t dlo = ptx::make_kmajor_sw128_desc(sb_addr + b_hi_bytes);
#pragma unroll
          for (int ks = 0; ks < kBK / kUmmaK; ++ks) {
            // advance 32 bytes along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            const uint32_t fresh = (c_first && ks == 0) ? 0u : 1u;
            ptx::mma_tf32_ts_if(leader, d_acc, a_hi + ks * kUmmaK, dlo + static_cast<uint64_t>(2 * ks), idesc, fresh);
            ptx::mma_tf32_ts_if(leader, d_acc, a_hi + ks * kUmmaK, dhi + static_cast<uint64_t>(2 * ks), idesc, 1u);
          }
        } else {
          const uint64_t dh16 = ptx::make_kmajor_sw64_desc(sb_addr + b_hi_bytes);
          const uint64_t dl16 = ptx::make_kmajor_sw64_desc(sb_addr + b_hi_bytes + b_16_bytes);
          const uint32_t a_h16 = a_hi + 32, a_l16 = a_hi + 48;
          // the small terms first: lo * hi' and hi * lo' as bf16 MMAs (16 reduction elements = 32 bytes = 8 TMEM
          // columns per step), then hi * hi' as tf32 MMAs (8 elements per step)
#pragma unroll
          for (int ks = 0; ks < kBK / kUmmaK16; ++ks) {
            // advance 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
            const uint32_t fresh = (c_first && ks == 0) ? 0u : 1u;
            ptx::mma_bf16_ts_if(leader, d_acc, a_l16 + 8 * ks, dh16 + static_cast<uint64_t>(2 * ks), idesc16, fresh);
            ptx::mma_bf16_ts_if(leader, d_acc, a_h16 + 8 * ks, dl16 + static_cast<uint64_t>(2 * ks), idesc16, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < kBK / kUmmaK; ++ks)
            ptx::mma_tf32_ts_if(leader, d_acc, a_hi + ks * kUmmaK, dhi + static_cast<uint64_t>(2 * ks), idesc, 1u);
        }
        if (c_last) {
          ptx::tc_commit_if(leader, &accfull_bar[mt]);
          ++mc;
        }
        ptx::tc_commit_if(leader, &aempty_bar[t]);
        ptx::tc_commit_if(leader, &bempty_bar[sbi]);
      }
      pos += len;
    }
    }
    __syncwarp();
  } else {
    // ===================================================== converter + flush/epilogue warps (one thread per X row)
    const int row = warp * 32 + lane;          // row inside the super-tile
    const int mt = warp >> 2;                  // which 128-row MMA tile
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
    const uint32_t acc_addr = tmem_base + lane_sel + mt * kAccStride;
    const uint32_t smem_x_u32 = ptx::smem_u32(smem_x);
    uint32_t it = 0, fc = 0, seg = 0;  // k-block counter, flushed-chunk counter, segment counter
    bool ok = true;
    RingPos rx;
    // fp32 master sum of this thread's accumulator row.  The tensor core adds into its accumulator with
    // round-toward-zero, which biases long sums of positive terms low (~2.5e-8 per MMA); chunks of C k-blocks
    // are therefore summed here with round-to-nearest.
    float master[Kp];
#pragma unroll
    for (int i = 0; i < Kp; ++i) master[i] = 0.f;

    auto flush = [&]() -> bool {
      if (!warp_wait_bar(&accfull_bar[mt], fc & 1, actx, ERR_EPI_ACCFULL, fc, mt)) return false;
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        uint32_t v[16];
        ptx::tmem_ld_x16(acc_addr + 16 * c, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) master[16 * c + i] += __uint_as_float(v[i]);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&accempty_bar[mt]);
      ++fc;
      return true;
    };

    for (int pc = 0; pc < ws.pieces && ok; ++pc) {
    long long range_begin, range_end;
    ws.cta_range(pc, gridDim.x, cta, range_begin, range_end);
    for (long long pos = range_begin; pos < range_end && ok; ++seg) {
      int run, tile, kb0, len;
      ws.decode(pos, run, tile, kb0, len);
      if (len > range_end - pos) len = static_cast<int>(range_end - pos);
      const int CL = (ORIENT == ORIENT_WX && tile == p.extra_tile) ? p.extra_chunk_log2 : p.chunk_log2;
      const int C = 1 << CL, CM = C - 1;
      for (int li = 0; li < len; ++li, ++it, rx.advance(SX)) {
        const int s = rx.s;
        const int t = it % kAStages;
        // ---- lagged flush: the chunk that ended kAStages k-blocks ago (its MMAs are the ones the aempty wait
        //      below waits for anyway), overlapped with the other tile's MMAs
        if (li >= kAStages && ((li - kAStages) & CM) == CM) {
          if (!flush()) {
            ok = false;
            break;
          }
        }
        // ---- X tile: shared memory fp32 -> (hi, lo) -> tensor memory stage t
        if (!warp_wait_bar(&xfull_bar[s], rx.ph, actx, ERR_CONV_XFULL, it, s) ||
            !warp_wait_bar(&aempty_bar[t], ((it / kAStages) & 1) ^ 1, actx, ERR_CONV_AEMPTY, it, t)) {
          ok = false;
          break;
        }
        ptx::tc_fence_after();
        const uint32_t sX = smem_x_u32 + static_cast<uint32_t>(s) * kXTileBytes;
        const uint32_t a_addr = tmem_base + lane_sel + kTmemAOff + t * kAStageCols + mt * 64;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          uint32_t hi[8], lo[8];
          if (ORIENT == ORIENT_XH && !tiles) {
            // tile is [32 cells][256 genes]; this thread owns gene `row`
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              ptx::split_tf32_fast(ptx::lds_f32(sX + ((8 * h + kk) * kRows + row) * 4), hi[kk], lo[kk]);
          } else {
            // tile is [256 rows][32] with the TMA 128B swizzle: 16-byte chunk c of row r sits at c ^ (r & 7)
            // (W^T X, and both contractions of the sparse path)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint32_t qa = sX + row * (kBK * 4) + (((2 * h + c) ^ (row & 7)) << 4);
              const float4 q = ptx::lds_v4(qa);
              if (tiles) ptx::sts_zero_v4(qa);  // hand the stage back zeroed
              ptx::split_tf32_fast(q.x, hi[4 * c + 0], lo[4 * c + 0]);
              ptx::split_tf32_fast(q.y, hi[4 * c + 1], lo[4 * c + 1]);
              ptx::split_tf32_fast(q.z, hi[4 * c + 2], lo[4 * c + 2]);
              ptx::split_tf32_fast(q.w, hi[4 * c + 3], lo[4 * c + 3]);
            }
          }
          ptx::tmem_st_x8(a_addr + 8 * h, hi);
          if (!EXACT) {
            // bf16 images for the correction terms, two reduction elements per column (even element in the low half)
            uint32_t h16[4], l16[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              h16[q] = ptx::pack_bf16x2(__uint_as_float(hi[2 * q]), __uint_as_float(hi[2 * q + 1]));
              l16[q] = ptx::pack_bf16x2(__uint_as_float(lo[2 * q]), __uint_as_float(lo[2 * q + 1]));
            }
            ptx::tmem_st_x4(a_addr + 32 + 4 * h, h16);
            ptx::tmem_st_x4(a_addr + 48 + 4 * h, l16);
          }
        }
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&xempty_bar[s]);  // the X slot can be refilled
          ptx::mbar_arrive(&cfull_bar[t]);   // A (TMEM) and B (shared memory) of this k-block are ready
        }
      }
      if (!ok) break;
      // ---- remaining chunks of this segment, then the segment's sum -> partial-sum slot [K][256]
      const int n_chunks = (len + C - 1) >> CL;
      const int flushed = (len >= kAStages) ? (len - kAStages) >> CL : 0;
      for (int r = flushed; r < n_chunks && ok; ++r) ok = flush();
      if (!ok) break;
      float* dst = p.partial + (static_cast<size_t>(cta) * p.max_segs + seg) * p.K * kRows + row;
#pragma unroll
      for (int i = 0; i < Kp; ++i) {
        if (i < p.K) dst[static_cast<size_t>(i) * kRows] = master[i];
        master[i] = 0.f;
      }
      pos += len;
    }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == kWarpMma) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------
// out[k][m] = sum over the segments of m's tile, piece by piece in stream-K order (fixed => deterministic).
// grid = (num_tiles, k-slices); threads stride over the 256 rows of the tile.
struct ReduceParams {
  const float* partial;
  int rows;  // 256
  int M, K;
  const int* slot_ofs;  // [num_tiles + 1]  prefix offsets into `slots` (built on the host from the WorkSpace)
  const int* slots;     // partial-sum slots of every tile, piece by piece in stream-K order
  float* out;
  long long ld;
};
// Host side: the slots that hold tile `tile` (same enumeration the kernel uses for its segment numbering).
inline void reduce_slots_of_tile(const WorkSpace& ws, int grid, int max_segs, int tile, std::vector<int>* out) {
  for (int pc = 0; pc < ws.pieces; ++pc) {
    const long long r_begin = ws.run_begin(pc, tile), r_end = r_begin + ws.len_of_piece(pc);
    const long long base = static_cast<long long>(pc) * ws.num_tiles * ws.piece_len;
    const int lp = ws.len_of_piece(pc);
    for (int q = 0; q < grid; ++q) {
      long long b, e;
      ws.cta_range(pc, grid, q, b, e);
      if (e <= b || e <= r_begin || b >= r_end) continue;
      int before = 0;  // segments CTA q has stored in earlier pieces, then earlier tiles of this piece
      for (int p2 = 0; p2 < pc; ++p2) before += ws.cta_segments(p2, grid, q);
      const int first_tile = static_cast<int>((b - base) / lp);
      out->push_back(q * max_segs + before + (tile - first_tile));
    }
  }
}
// Largest number of segments any CTA stores (slots per CTA).
inline int max_segments_per_cta(const WorkSpace& ws, int grid) {
  int best = 1;
  for (int q = 0; q < grid; ++q) {
    int n = 0;
    for (int pc = 0; pc < ws.pieces; ++pc) n += ws.cta_segments(pc, grid, q);
    best = n > best ? n : best;
  }
  return best;
}
// Few slots per tile (the X contractions): grid = (num_tiles * 8, gy); block = (tile, 32-row group); lane = row
// (coalesced 128-byte reads), each warp owns a strided set of components and adds the tile's slots in order.
__global__ void __launch_bounds__(256) reduce_partials_by_k_kernel(const ReduceParams p) {
  ptx::pdl_enter();
  const int tile = blockIdx.x >> 3, rg = blockIdx.x & 7;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = rg * 32 + lane;
  const long long m = static_cast<long long>(tile) * p.rows + r;
  if (m >= p.M) return;
  const int s0 = p.slot_ofs[tile], s1 = p.slot_ofs[tile + 1];
  for (int k = blockIdx.y * 8 + w; k < p.K; k += gridDim.y * 8) {
    float acc = 0.f;
    for (int q = s0; q < s1; ++q)
      acc += __ldcg(p.partial + (static_cast<size_t>(__ldg(p.slots + q)) * p.K + k) * p.rows + r);
    p.out[static_cast<long long>(k) * p.ld + m] = acc;
  }
}
// Many slots per tile (the Gram contractions: one tile, one slot per CTA): grid = (num_tiles * 8, K); one block
// per (tile, 32-row group, component); the 8 warps take the slots round-robin and are combined in a fixed order.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceParams p) {
  ptx::pdl_enter();
  __shared__ float red[8][32];
  const int tile = blockIdx.x >> 3, rg = blockIdx.x & 7, k = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = rg * 32 + lane;
  const long long m = static_cast<long long>(tile) * p.rows + r;
  if (static_cast<long long>(tile) * p.rows + rg * 32 >= p.M) return;  // whole row group out of range
  const int s0 = p.slot_ofs[tile], s1 = p.slot_ofs[tile + 1];
  float acc = 0.f;
  for (int q = s0 + w; q < s1; q += 8)
    acc += __ldcg(p.partial + (static_cast<size_t>(__ldg(p.slots + q)) * p.K + k) * p.rows + r);
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && m < p.M) {
    float t = red[0][lane];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += red[i][lane];
    p.out[static_cast<long long>(k) * p.ld + m] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
#ifdef ALPINE_B200_DEBUG_SIMT
// Plain fp32 CUDA-core version of the same contractions: compiled only into A/B-checking builds
// (-DALPINE_B200_DEBUG_SIMT, then selected with ALPINE_B200_GEMM=simt); the shipped library does not contain it.
// out[k][m] = sum_r A(m, r) * B[k][r];  XH: A(m, r) = X[r * ldX + m];  WX: A(m, r) = X[m * ldX + r].
template <int ORIENT>
__global__ void simt_gemm_kernel(const float* __restrict__ X, long long ldX, const float* __restrict__ B,
                                 long long ldB, int M, int R, int K, float* __restrict__ out, long long ld_out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (m >= M || k >= K) return;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) {
    const float a = (ORIENT == ORIENT_XH) ? X[static_cast<long long>(r) * ldX + m] : X[static_cast<long long>(m) * ldX + r];
    acc = fmaf(a, __ldg(B + static_cast<long long>(k) * ldB + r), acc);
  }
  out[static_cast<long long>(k) * ld_out + m] = acc;
}
#endif  // ALPINE_B200_DEBUG_SIMT

}  // namespace alpine
