// Layout probe for the tensor-core ALS kernel: tcgen05.mma kind::f16 (fp16 operands, fp32 accumulator in tensor memory)
// with K-major, un-swizzled ("interleaved") shared-memory operands -- 8-row x 16-byte core matrices, SBO = distance of
// core matrices along M/N, LBO = distance along K.  Checks D = A B^T (M = N = 128, K = 32 as two K = 16 instructions)
// against a host reference, and measures how the fp32 accumulator rounds (sum of many equal small terms).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../buffalo_b200/csrc -o mma_probe mma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "sm100_ptx.cuh"
using namespace bfl::sm100;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// operand slab for KT k-values: element (m, k) at byte  (k/8)*LBO + (m/8)*128 + (m%8)*16 + (k%8)*2,  LBO = 2048
constexpr int KT = 32, LBO = 2048, SBO = 128;

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int reps, int swap) {
    __shared__ __align__(1024) unsigned char sa[KT / 8 * LBO];
    __shared__ __align__(1024) unsigned char sb[KT / 8 * LBO];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = 0; k < KT; ++k) {
        const int off = (k / 8) * LBO + (tid / 8) * 128 + (tid % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__half*>(sa + off) = A[tid * KT + k];
        *reinterpret_cast<__half*>(sb + off) = B[tid * KT + k];
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tmem_alloc(&tbase, 128);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tbase;
    if (tid == 0) {
        const uint32_t id = idesc_f16_k(128, 128);
        for (int r = 0; r < reps; ++r)
            for (int ks = 0; ks < KT / 16; ++ks) {
                const uint32_t lbo = swap ? SBO : LBO, sbo = swap ? LBO : SBO;
                const uint64_t da = smem_desc(s32(sa) + ks * 2 * LBO, lbo, sbo);
                const uint64_t db = smem_desc(s32(sb) + ks * 2 * LBO, lbo, sbo);
                mma_f16(tmem, da, db, id, (r > 0 || ks > 0) ? 1u : 0u);
            }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) D[tid * 128 + c * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    std::vector<__half> hA(128 * KT), hB(128 * KT);
    std::vector<float> fA(128 * KT), fB(128 * KT);
    srand(1);
    for (int i = 0; i < 128 * KT; ++i) {
        hA[i] = __float2half((rand() % 2001 - 1000) / 500.0f);
        hB[i] = __float2half((rand() % 2001 - 1000) / 500.0f);
        fA[i] = __half2float(hA[i]);
        fB[i] = __half2float(hB[i]);
    }
    __half *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 128 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    std::vector<float> hD(128 * 128);
    for (int swap = 0; swap < 1; ++swap) {   // the swapped (LBO, SBO) assignment faults (reads beyond the slab): checked once
        probe<<<1, 128>>>(dA, dB, dD, 1, swap);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                double ref = 0;
                for (int k = 0; k < KT; ++k) ref += (double)fA[m * KT + k] * fB[n * KT + k];
                maxerr = fmax(maxerr, fabs(ref - hD[m * 128 + n]));
                maxref = fmax(maxref, fabs(ref));
            }
        printf("kind::f16 K-major interleaved, (LBO,SBO) = (%d,%d): max |err| %.3e (max |ref| %.3e) -> %s\n", swap ? SBO : LBO,
               swap ? LBO : SBO, maxerr, maxref, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
    }
    // accumulator rounding: every operand entry 1 + 2^-10 (exact in fp16): each K=16 instruction adds 16 (1+2^-10)^2 to
    // every accumulator entry; after `reps` x 2 instructions compare with the exact sum
    for (int i = 0; i < 128 * KT; ++i) hA[i] = hB[i] = __float2half(1.0f + 1.0f / 1024.0f);
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    for (int reps : {1, 8, 64, 512}) {
        probe<<<1, 128>>>(dA, dB, dD, reps, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        const double x = 1.0 + 1.0 / 1024.0, exact = (double)reps * KT * x * x;
        printf("accumulate %4d x %d entries: got %.9g exact %.9g rel err %.3e\n", reps, KT, (double)hD[5 * 128 + 7], exact,
               (hD[5 * 128 + 7] - exact) / exact);
    }
    return 0;
}
