// sm_100a building blocks for the fused MLP kernels: mbarrier, bulk async copies (TMA 1-D),
// tcgen05 MMA / TMEM, UMMA shared-memory descriptors and the 128B-swizzled "slab" layout.
//
// Slab layout (used for activations, weights and the HBM stash alike): a slab holds R rows of
// 64 bf16 (128 B). Row r lives at byte r*128; inside a row the eight 16-byte chunks are
// XOR-swizzled with (r & 7): chunk c is stored at chunk position c ^ (r & 7). This is the
// canonical UMMA SWIZZLE_128B layout, so the same bytes serve as
//   - a K-major operand  (rows = M or N index, the 64 columns = K)          [fwd, dgrad]
//   - an MN-major operand (rows = K index, the 64 columns = M or N index)   [wgrad]
// Slabs must start on a 1024-byte boundary.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace nerfb200 {
namespace tc {

constexpr int kSlabCols = 64;          // bf16 per slab row
constexpr int kSlabRowBytes = 128;
constexpr uint32_t kWaitTimeoutCycles = 2000000000u;  // ~1 s: never hang the GPU box

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// byte offset of element (row, col) inside a slab
__device__ __host__ __forceinline__ uint32_t slab_offset(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 3) ^ (row & 7u)) << 4) | ((col & 7u) << 1));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// The suspend-time hint lets the hardware park the thread until the phase completes (or the
// hint, in nanoseconds, expires) instead of returning after the short default limit: pollers
// then cost no issue slots of the scheduler they share with the epilogue warps.
#ifndef NB_SUSPEND_NS
#define NB_SUSPEND_NS 20000u
#endif
constexpr uint32_t kTryWaitSuspendNs = NB_SUSPEND_NS;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kTryWaitSuspendNs)
      : "memory");
  return ok != 0;
}
// Waits for the phase with the given parity; on timeout records the error and traps, so a
// protocol bug surfaces as a launch failure instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if ((unsigned long long)(clock64() - t0) > kWaitTimeoutCycles) {
      printf("nerfb200: mbarrier wait timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---- proxies / fences ----------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Where the generic-proxy -> async-proxy fence of a slab publication is executed. Writer side (default
// off): every row warp fences before its mbarrier arrival — the fence is a MEMBAR.ALL.CTA, which also
// waits for the warp's global stores in flight. Consumer side (NB_CONSUMER_FENCE): the row warps only
// release-arrive; the ONE thread that issues the async-proxy reads (MMA issuer / stash copier) fences
// after it has acquired the barrier phase (the writes are then ordered before the fence in causality
// order, the async operations after it in program order).
#ifdef NB_CONSUMER_FENCE
__device__ __forceinline__ void writer_proxy_fence() {}
__device__ __forceinline__ void consumer_proxy_fence() { fence_proxy_async(); }
#else
__device__ __forceinline__ void writer_proxy_fence() { fence_proxy_async(); }
__device__ __forceinline__ void consumer_proxy_fence() {}
#endif
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- bulk async copies (TMA engine, 1-D) -----------------------------------------------------
// global -> shared, completion on an mbarrier (bytes: multiple of 16, 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global, tracked by the issuing thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until all but the newest N committed groups have finished READING their smem source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMEM ----------------------------------------------------------------------------------
// one full warp; ncols: power of two in [32,512]; the base address lands in *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets lane (base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// eight columns per thread (32 lanes x 8 x 32 bit), and the matching store: registers parked in TMEM
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
// As tmem_ld_wait, and ties the destination registers of an earlier tcgen05.ld to the wait, so
// that the compiler cannot schedule their uses above it while another load is still in flight.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                 "+r"(v[31])
               :
               : "memory");
}

// ---- UMMA ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a SWIZZLE_128B slab (cute::UMMA::SmemDescriptor bits:
// [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=2).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major operand: rows = M/N index (8-row atoms 1024 B apart), one 64-wide K slab.
// k16 = which 16-element K step inside the slab (advances the start address by 32 B).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t slab_addr, uint32_t row0,
                                                     uint32_t k16) {
  return umma_desc(slab_addr + row0 * 128u + k16 * 32u, 0u, 1024u);
}
// MN-major operand: rows = K index; MN atoms of 64 elements are whole slabs `slab_stride`
// bytes apart. k16 = which group of 16 rows (advances by 2048 B).
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t slab_addr, uint32_t slab_stride,
                                                      uint32_t k16) {
  return umma_desc(slab_addr + k16 * 2048u, slab_stride, 1024u);
}

// K-major operand in the SWIZZLE_32B layout: rows of 16 bf16 (32 B), 8-row groups 256 B apart, the
// two 16-byte halves of a row swapped when row bit 2 is set; one [rows][16] image = one K step.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw32(uint32_t image_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((image_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major layouts)
  d |= (uint64_t)((256u >> 4) & 0x3FFFu) << 32; // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;                       // SWIZZLE_32B
  return d;
}

// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor bits).
__device__ __host__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)  // a_major
         | ((b_mn_major ? 1u : 0u) << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)      // n_dim
         | ((uint32_t)(M >> 4) << 24);     // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                     uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace tc
}  // namespace nerfb200
