// Thin inline-PTX wrappers for the sm_100a features the tensor-core ALS kernel uses: mbarrier, cp.async (SASS LDGSTS),
// tensor memory (tcgen05.alloc/ld/st, SASS LDTM/STTM) and tcgen05.mma kind::f16 (SASS UTCHMMA) with shared-memory matrix
// descriptors.  (The 1-D bulk-copy / TMA wrappers of the first version live on in benchmarks/gather_probe.cu.)  No CUTLASS: every string below is plain PTX ISA 8.8.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bfl {
namespace sm100 {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(s32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// for waits that are expected to be long (an idle role): back off between polls so that the spinning warp does not take
// issue slots from the working warps of its scheduler (a failed try_wait returns after a few cycles: the tight loop of the
// four idle epilogue warps was 13 % of all executed instructions on the item side)
__device__ __forceinline__ void mbar_wait_idle(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}

// generic-proxy writes to shared memory -> visible to the async proxy (tensor-core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 16-byte cp.async (SASS LDGSTS), completion by commit / wait groups ---------------------------------
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// waits until at most N of the executing thread's most recent commit groups are still pending
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- tensor memory ----------------------------------------------------------------------------
// whole warp; ncols: power of two in [32, 512]; the base address is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets lane (lane_base + t), columns [col, col+32)
// taddr = tmem_base + (lane_base << 16) + col; a warp may only touch lanes 32*(warp_id % 4) .. +31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// ---- tcgen05.mma, operands from shared memory, fp32 accumulator in tensor memory ------------------------
// Shared-memory matrix descriptor (PTX ISA "matrix descriptor", tcgen05 flavour: bits 46-47 = 0b01):
//   [0,14) start address >> 4 | [16,30) leading-dimension byte offset >> 4 | [32,46) stride-dimension byte offset >> 4
//   [61,64) swizzle mode (0 = none)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar))
                 : "memory");
}

// ---- tcgen05.mma kind::f16 (fp16 operands, fp32 accumulator) -------------------------------------------
// K-major un-swizzled ("interleaved") operand: core matrices of 8 rows (M/N) x 16 bytes (8 fp16 along K), stored as 128
// contiguous bytes; SBO = distance between core matrices that are neighbours along M/N, LBO = distance between
// neighbours along K (probe: benchmarks/mma_probe.cu).  One instruction covers K = 16 (two core matrices deep).
__host__ __device__ constexpr uint32_t idesc_f16_k(int M, int N) {   // D fp32, A/B fp16, both K-major, dense
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t IDESC_NEGATE_A = 1u << 13;
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// packed fp32 pairs (Blackwell FMUL2 / FFMA2)
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
// Two-term fp16 split of a pair of fp32 values (k even -> low half, k odd -> high half of each 32-bit word):
// head = the value with its low 13 mantissa bits cleared (exactly an fp16 number while it is in the normal fp16 range),
// tail = value - head rounded to fp16; head + tail carries >= 21 significant bits.
__device__ __forceinline__ void split_f16x2(float2 s, uint32_t& head, uint32_t& tail) {
    float2 h;
    h.x = __uint_as_float(__float_as_uint(s.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(s.y) & 0xffffe000u);
    const float2 t = f2fma(h, make_float2(-1.f, -1.f), s);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(head) : "f"(h.y), "f"(h.x));
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(tail) : "f"(t.y), "f"(t.x));
}

}  // namespace sm100
}  // namespace bfl
