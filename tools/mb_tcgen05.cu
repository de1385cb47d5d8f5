// Microbenchmark: raw tcgen05.mma issue/execute throughput on B200 for the shapes the Ozaki kernel
// can use.  One CTA per SM, one thread issues `iters` back-to-back MMAs on fixed (zeroed) shared-memory
// operands into TMEM accumulators, commits, waits.  Reports clk / MMA and dense TOP/s for
//   kind::i8  (int8 x int8 -> int32)  and  kind::f8f6f4 (e4m3 x e4m3 -> fp32),  M = 128, N in {64,128,256},
// with the accumulator either re-used (same D every time) or rotated over several D regions.
// This is the measured denominator for the int8 roofline (MEASURED_PEAKS.json only has bf16).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>  // 0: i8, 1: f8f6f4
__device__ __forceinline__ void mma(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
                     " tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n}" ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %5, 0;\n mov.b64 da, {%1, %3};\n mov.b64 db, {%2, %3};\n"
                     " tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n}" ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int N, int ROT>
__global__ void __launch_bounds__(128, 1) mb_kernel(int iters, long long* clk_out) {
    extern __shared__ unsigned char raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    // zero operands: A 128 x 64 B, B 256 x 64 B (SWIZZLE_64B K-major)
    for (int i = threadIdx.x; i < (128 + 256) * 64 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t hi = 32u | (1u << 14) | (4u << 29);           // SBO 512 B, version 1, SWIZZLE_64B
        const uint32_t alo = (base >> 4) | (1u << 16), blo = ((base + 128 * 64) >> 4) | (1u << 16);
        uint32_t idesc;
        if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);   // S32, INT8, INT8
        else idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);             // F32, E4M3, E4M3
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) mma<KIND>(tmem + (uint32_t)((j % ROT) * N), alo, blo, hi, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) *clk_out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int KIND, int N, int ROT>
int run(const char* name, int sms, int iters) {
    long long* d;
    CK(cudaMalloc(&d, 8));
    const size_t smem = (128 + 256) * 64 + 2048;
    CK(cudaFuncSetAttribute(mb_kernel<KIND, N, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    long long clk = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        mb_kernel<KIND, N, ROT><<<sms, 128, smem>>>(iters, d);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; CK(cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost)); }
    }
    const double ops = 2.0 * 128 * N * 32 * (double)iters * sms;
    printf("%-28s N=%3d rot=%d : %7.2f clk/MMA  %8.1f TOP/s dense (%d SMs, %.3f ms)\n", name, N, ROT, (double)clk / iters,
           ops / best * 1e-9, sms, best);
    cudaFree(d);
    return 0;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 40000;
    if (run<0, 64, 1>("i8 (same accumulator)", sms, iters)) return 1;
    if (run<0, 64, 4>("i8 (4 accumulators)", sms, iters)) return 1;
    if (run<0, 128, 1>("i8 (same accumulator)", sms, iters)) return 1;
    if (run<0, 128, 4>("i8 (4 accumulators)", sms, iters)) return 1;
    if (run<0, 256, 1>("i8 (same accumulator)", sms, iters)) return 1;
    if (run<0, 256, 2>("i8 (2 accumulators)", sms, iters)) return 1;
    if (run<1, 64, 4>("e4m3 (4 accumulators)", sms, iters)) return 1;
    if (run<1, 128, 4>("e4m3 (4 accumulators)", sms, iters)) return 1;
    if (run<1, 256, 2>("e4m3 (2 accumulators)", sms, iters)) return 1;
    return 0;
}
