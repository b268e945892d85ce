// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma kind::tf32 / st / ld / commit / fences).  No library dependencies.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace alpine {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- programmatic dependent launch
// Kernels of the iteration are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel lets its
// successor in the stream be scheduled as soon as all of its own CTAs have started (launch_dependents), and waits
// for the complete, flushed predecessor grid before it touches global memory (wait).  Both are no-ops in a kernel
// that was launched without the attribute.  EVERY CTA of such a kernel must execute the wait (grid completion is
// what the successor waits for, so a CTA that skipped it could let the successor overtake the predecessor).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// shared-memory accesses through 32-bit shared-window addresses (generic pointers cost 64-bit address math)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ void sts_f32(uint32_t addr, uint32_t bits) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(bits) : "memory");
}
__device__ __forceinline__ void sts_zero_v4(uint32_t addr) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<uint64_t>(p)) : "memory");
}

// ---------------------------------------------------------------- TMA
// L2 cache-policy descriptors (the encodings createpolicy.fractional produces).
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on an mbarrier when all previously issued tcgen05 async ops of this thread are done
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// one lane of a converged warp
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred;
}
// The `_if` variants are executed by a whole converged warp with warp-uniform operands; only lanes with
// leader != 0 (one elected lane) issue.  Keeping the operands uniform lets ptxas feed the instruction from
// uniform registers instead of a per-operand R2UR "waterfall" loop.
__device__ __forceinline__ void mma_tf32_ts_if(uint32_t leader, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// same for kind::f16 with bf16 operands (A: two elements per 32-bit TMEM column, B: shared-memory descriptor), fp32 D
__device__ __forceinline__ void mma_bf16_ts_if(uint32_t leader, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tc_commit_if(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}"
      ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}

// 32 lanes x N consecutive columns, one 32-bit register per column
__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Operand split of a FINITE fp32 value: hi = x rounded to nearest (ties away) to tf32's 10 explicit
// mantissa bits, lo = (x - hi) (exact in fp32, |lo| <= 2^-11 |x|) rounded the same way, both with the low 13
// bits zero so the result does not depend on how the tensor core treats them.  |x - hi - lo| <= 2^-22 |x|.
// (cvt.rna.tf32.f32 compiles to the same add/mask plus an Inf/NaN guard; X, W, H are finite by contract.)
__device__ __forceinline__ uint32_t round_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// Same split with lo left as the exact fp32 remainder: the tensor core reads only the tf32 bits of an operand
// (the low 13 mantissa bits are ignored), i.e. it truncates lo; lo has a random sign, so this adds no bias and
// |x - hi - trunc(lo)| <= 2^-21 |x|.  Three integer/float ops per element, used on the streamed X tiles.
__device__ __forceinline__ void split_tf32_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = round_tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// The two correction products of the split (hi * lo and lo * hi, each ~2^-11 of the main product) run as bf16 MMAs at
// twice the tf32 rate: their operands are the round-to-nearest-even bf16 values of hi and of the exact remainder
// lo = x - hi.  The rounding is unbiased and contributes ~2^-9 * 2^-11 = 2^-20 relative per product term with a random
// sign, which averages out over the reduction exactly like the 2^-22 of the plain 3xTF32 split (measured on the golden
// trajectories: same deviation from the reference).
// {even element -> low half, odd element -> high half}: the order of two adjacent K elements in a 32-bit word
__device__ __forceinline__ uint32_t pack_bf16x2(float even, float odd) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
  return r;
}
__device__ __forceinline__ uint16_t bf16_bits(float x) {
  const uint32_t b = __float_as_uint(x);
  return static_cast<uint16_t>((b + 0x7FFFu + ((b >> 16) & 1u)) >> 16);
}
// The split copies of a [K][ld] operand: plane 0 = [K][ld] tf32 hi values; plane 1 = [K][2 * ld] floats, row k =
//   [ ld floats read as 16-bit: bf16(hi)[ld] followed by bf16(lo)[ld] | ld floats: tf32(lo) ]
// (ld % 8 == 0).  The bf16 images feed the correction MMAs of the general kernel, the tf32 lo copy the count-matrix
// variant, whose streamed operand has no lo half and whose correction term stays a tf32 MMA.
__device__ __forceinline__ void store_split1(float x, float* plane0, float* plane1, long long k, long long col, long long ld) {
  const uint32_t hi = round_tf32(x);
  const float lo = x - __uint_as_float(hi);
  plane0[k * ld + col] = __uint_as_float(hi);
  float* row1 = plane1 + 2 * k * ld;
  uint16_t* p16 = reinterpret_cast<uint16_t*>(row1) + col;
  p16[0] = bf16_bits(__uint_as_float(hi));
  p16[ld] = bf16_bits(lo);
  row1[ld + col] = __uint_as_float(round_tf32(lo));
}
// four adjacent columns (col % 4 == 0)
__device__ __forceinline__ void store_split4(float4 v, float* plane0, float* plane1, long long k, long long col, long long ld) {
  const float x[4] = {v.x, v.y, v.z, v.w};
  float h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __uint_as_float(round_tf32(x[i]));
    l[i] = x[i] - h[i];
  }
  *reinterpret_cast<float4*>(plane0 + k * ld + col) = make_float4(h[0], h[1], h[2], h[3]);
  float* row1 = plane1 + 2 * k * ld;
  uint16_t* p16 = reinterpret_cast<uint16_t*>(row1) + col;
  *reinterpret_cast<uint2*>(p16) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
  *reinterpret_cast<uint2*>(p16 + ld) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
  *reinterpret_cast<float4*>(row1 + ld + col) =
      make_float4(__uint_as_float(round_tf32(l[0])), __uint_as_float(round_tf32(l[1])),
                  __uint_as_float(round_tf32(l[2])), __uint_as_float(round_tf32(l[3])));
}

// ---------------------------------------------------------------- descriptors
// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 B apart.
// (layout of cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type [61,64) with SWIZZLE_128B = 2)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;             // LBO (ignored for swizzled K-major), canonical value 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO = 1024 B between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
  return d;
}
// Same for 64-byte rows (32 bf16 reduction elements per row): SWIZZLE_64B (layout_type 4), 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// UMMA instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major.
// (cute::UMMA::InstrDescriptor: c_format [4,6)=1 F32, a_format [7,10)=2 TF32, b_format [10,13)=2 TF32,
//  a_major bit 15, b_major bit 16, n_dim [17,23) = N>>3, m_dim [24,29) = M>>4)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// kind::f16 with bf16 A and B (format 1), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace ptx
}  // namespace alpine
