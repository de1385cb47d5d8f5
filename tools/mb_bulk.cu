// Microbenchmark: throughput of 1-D TMA bulk copies (cp.async.bulk global->shared, UBLKCP) as a
// function of copy size, issued by one thread per CTA into a 4-slot ring, 2 CTAs per SM, source
// L2-resident.  Question: is the per-copy cost of many 1 KB copies what starves gemm_nt's ring?
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory"); } while (!ok);
}
// each "slab" = 24576 bytes moved with `ncopies` copies of 24576/ncopies bytes
__global__ void __launch_bounds__(32) k_bulk(const char* src, size_t src_bytes, int ncopies, int slabs, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int SLAB = 24576, ST = 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + ST * SLAB);
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int csz = SLAB / ncopies;
    size_t off = ((size_t)blockIdx.x * 7919 * SLAB) % (src_bytes - SLAB);
    off &= ~(size_t)127;
    auto issue = [&](int q) {
        int slot = q % ST;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[slot])), "r"(SLAB) : "memory");
        for (int c = 0; c < ncopies; c++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(smem + slot * SLAB + c * csz)), "l"(src + off + (size_t)c * csz), "r"(csz), "r"(s32(&full[slot])) : "memory");
        off += SLAB * 37; if (off + SLAB > src_bytes) off -= (src_bytes - SLAB) & ~(size_t)127;
    };
    for (int q = 0; q < ST - 1 && q < slabs; q++) issue(q);
    for (int q = 0; q < slabs; q++) {
        if (q + ST - 1 < slabs) issue(q + ST - 1);
        mbar_wait(&full[q % ST], (q / ST) & 1);
    }
    if (smem[5] == 77) sink[0] = 1;
}
// gemm_nt-like pattern: one slab = 16 segments of 1024 B + 16 segments of 512 B, consecutive
// segments `ld` bytes apart (column-major panel with leading dimension ld), or contiguous (ld=0).
__global__ void __launch_bounds__(32) k_strided(const char* src, size_t src_bytes, size_t ld, int slabs, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int SLAB = 24576, ST = 4;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + ST * SLAB);
    if (threadIdx.x == 0) {
        for (int s = 0; s < ST; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // CTA b streams "tiles": row offset (b*1024) bytes within the columns, 16 slabs (256 columns) per tile
    size_t rows_bytes = ld ? ld : 1024;
    size_t row_off = ((size_t)blockIdx.x * 1024) % (rows_bytes - 1024 + 1);
    int col = 0;
    auto issue = [&](int q) {
        int slot = q % ST;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[slot])), "r"(SLAB) : "memory");
        for (int c = 0; c < 16; c++) {
            size_t a = ld ? ((size_t)(col + c) * ld + row_off) : ((size_t)blockIdx.x * 393216 + (size_t)(col + c) * 1536) % (src_bytes - 4096);
            a &= ~(size_t)15;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(smem + slot * SLAB + c * 1536)), "l"(src + a), "r"(1024), "r"(s32(&full[slot])) : "memory");
            size_t b2 = ld ? ((size_t)(col + c) * ld + (row_off + 77 * 1024) % (rows_bytes - 1024 + 1)) : a + 1024;
            b2 &= ~(size_t)15;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s32(smem + slot * SLAB + c * 1536 + 1024)), "l"(src + b2), "r"(512), "r"(s32(&full[slot])) : "memory");
        }
        col += 16; if (col >= 256) { col = 0; row_off = (row_off + 296 * 1024) % (rows_bytes - 1024 + 1); }
    };
    for (int q = 0; q < ST - 1 && q < slabs; q++) issue(q);
    for (int q = 0; q < slabs; q++) {
        if (q + ST - 1 < slabs) issue(q + ST - 1);
        mbar_wait(&full[q % ST], (q / ST) & 1);
    }
    if (smem[5] == 77) sink[0] = 1;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    size_t src_bytes = 96u << 20;
    char* src; CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 1, src_bytes));
    double* sink; CK(cudaMalloc(&sink, 8));
    size_t smem = 4 * 24576 + 64;
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int slabs = 4000;
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ctas_per_sm++)
        for (int nc : {1, 2, 6, 12, 24, 48, 96}) {
            int blocks = sms * ctas_per_sm;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            k_bulk<<<blocks, 32, smem>>>(src, src_bytes, nc, 200, sink); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); k_bulk<<<blocks, 32, smem>>>(src, src_bytes, nc, slabs, sink); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double bytes = (double)blocks * slabs * 24576.0;
            printf("{\"ctas_per_sm\":%d,\"copy_bytes\":%d,\"copies_per_slab\":%d,\"GBps_total\":%.0f,\"GBps_per_sm\":%.1f,\"ns_per_copy_per_sm\":%.1f}\n",
                   ctas_per_sm, 24576 / nc, nc, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / sms,
                   ms * 1e6 / ((double)slabs * nc * ctas_per_sm));
        }
    {
        // panel-like source: 256 columns x ld bytes
        for (size_t ld : {(size_t)0, (size_t)65536, (size_t)262144, (size_t)524288}) {
            size_t need = ld ? ld * 256 : (size_t)128 << 20;
            char* psrc; CK(cudaMalloc(&psrc, need)); CK(cudaMemset(psrc, 1, need));
            CK(cudaFuncSetAttribute(k_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int blocks = sms * 2;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            k_strided<<<blocks, 32, smem>>>(psrc, need, ld, 200, sink); CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); k_strided<<<blocks, 32, smem>>>(psrc, need, ld, slabs, sink); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            double bytes = (double)blocks * slabs * 24576.0;
            printf("{\"pattern\":\"gemm-like 16x1024B+16x512B per slab\",\"ld_bytes\":%zu,\"panel_MB\":%.0f,\"GBps_total\":%.0f,\"GBps_per_sm\":%.1f,\"us_per_slab_per_cta\":%.2f}\n",
                   ld, need / 1048576.0, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / sms, ms * 1e3 / slabs);
            CK(cudaFree(psrc));
        }
    }
    return 0;
}
