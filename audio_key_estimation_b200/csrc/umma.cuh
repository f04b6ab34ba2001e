// umma.cuh -- thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path used by libake_b200:
// tcgen05.mma with shared-memory operand descriptors, TMEM allocation / loads, mbarriers, bulk async copies.
//
// Operand layout used everywhere in this library: SWIZZLE_NONE, K-major "planar chunk" form.  A chunk is 16 bytes
// (8 fp16 values) of K for one row; chunk c of row r lives at   start + c * LBO + r * 16   (SBO = 128 B: eight rows of
// 16 B form one contiguous core matrix).  Consequences the kernels exploit (all verified on hardware by
// tools/umma_probe.cu):
//   * the start address may be ANY multiple of 16 B, so "row r of the operand" can be "position q0 + r" of a
//     flattened activation tile -- a convolution tap is just a different start address (no im2col copy);
//   * LBO is free: the second K chunk of an MMA can be the next channel group (planar layout), the next time frame
//     (LBO = 16) or the next pitch row (LBO = row pitch), so one K = 16 MMA covers two taps of 8 channels;
//   * an MMA costs about (A bytes + B bytes) / 128 cycles of shared-memory bandwidth (M = 128, K = 16 fp16:
//     4 KB of A), so wide N amortises the A read.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace ake {

// fp32x2 packed arithmetic (sm_100: FMUL2 / FADD2) halves the scale / residual instructions of the hi/lo split
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// o[e] += x[e] (four fp32 pairs) where v < 0: one predicate, four predicated FADD2 (a C++ conditional around f2_add costs
// two extra moves per pair)
__device__ __forceinline__ void f2_add4_if_neg(uint64_t (&o)[4], uint64_t x0, uint64_t x1, uint64_t x2, uint64_t x3, int v) {
  asm("{\n\t.reg .pred q;\n\tsetp.lt.s32 q, %8, 0;\n\t@q add.rn.f32x2 %0, %0, %4;\n\t@q add.rn.f32x2 %1, %1, %5;\n\t"
      "@q add.rn.f32x2 %2, %2, %6;\n\t@q add.rn.f32x2 %3, %3, %7;\n\t}\n"
      : "+l"(o[0]), "+l"(o[1]), "+l"(o[2]), "+l"(o[3])
      : "l"(x0), "l"(x1), "l"(x2), "l"(x3), "r"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor (SWIZZLE_NONE, K-major): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1 [46,48).
__host__ __device__ constexpr uint64_t desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes = 128) {
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t make_desc(uint64_t hi, uint32_t smem_addr) { return hi | (uint64_t)((smem_addr >> 4) & 0x3FFF); }

// Instruction descriptor, kind::f16 with fp16 inputs (format 0) and fp32 accumulation, K-major A and B, M = 128.
__host__ __device__ constexpr uint32_t idesc_f16(int N, int M = 128) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- single-lane issue ---------------------------------------------------------------------------
// tcgen05.mma / commit take their operands from the uniform datapath.  Under `if (lane == 0)` the compiler cannot prove
// uniformity and wraps every such instruction in an ELECT / BRA.U.ANY "waterfall" with R2UR moves (~60-100 cycles per MMA
// from one thread: slower than the MMA itself).  Keep the issuing warp converged instead: a warp index the compiler knows
// to be uniform (shuffle broadcast) for the role branch, and elect.sync for the one lane that issues.
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// ---- MMA issue / completion ---------------------------------------------------------------------
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (count 1) on an mbarrier once every MMA previously issued by this thread has completed.
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tensor core / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
// Call from ONE full warp.  `cols` is a power of two in [32, 512]; the base address is written to *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// 32 lanes x 16 consecutive columns: thread i of warp w receives lane 32 * (w % 4) + i.  `addr` = base + (lane0 << 16) + column.
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Asynchronous form: issue the load, then tmem_ld_wait16() on the same registers before reading them (the wait names the
// registers as in/out operands so that neither nvcc nor ptxas can move a consumer above it).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t addr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// asynchronous 8-column form (see tmem_ld16_issue)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t addr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]) : : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
// Same with a back-off between polls: for the roles of a warp-specialised kernel that are not on the critical path (their
// spinning would otherwise take issue slots and instruction-cache bandwidth from the warps that have work).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(40);
  }
}
// Same with a caller-chosen back-off: many waiting warps polling every 40 ns take a visible share of the issue slots.
template <int NS>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(NS);
  }
}
// 1-D bulk async copy global -> shared, completion signalled on `bar` (bytes: multiple of 16; both addresses 16 B aligned).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Same with an L2 eviction-priority hint (createpolicy): `evict_first` for data that is read exactly once (the input audio of
// the first cascade pass), so that the stream does not push the intermediate levels the later passes re-read out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// Pull `bytes` (multiple of 16) of global memory into L2 ahead of the bulk copy that will land it in shared memory: a
// bulk copy cannot start before its shared-memory buffer is free, a prefetch can.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// ---- fp16 two-term split ------------------------------------------------------------------------
// x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits while |x| stays in fp16's normal range.
// Products are formed as hi*Whi + lo*Whi + hi*Wlo (fp32 accumulation in TMEM); the dropped lo*lo term is ~2^-22.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
// Two values at once with the packed conversions (one F2FP per pair instead of two F2F + PRMT): hi/lo as half2 bit patterns.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h), lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

}  // namespace umma
}  // namespace ake
