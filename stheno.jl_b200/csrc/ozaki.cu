// K3': fp64 trailing update on the 5th-generation tensor cores (tcgen05 + TMEM) by integer slicing.
//
// tcgen05.mma has no fp64 kind; the legacy DMMA path tops out at 37 TFLOP/s on B200.  The only way
// past that ceiling is the Ozaki scheme on the int8 kind (4.5 POP/s dense on sm_100a): every row
// of an operand panel is scaled by a power of two and cut into S = 7 signed 8-bit digits
//      x_ik  ~=  2^(e_i - 55) * sum_{p=1..7} d_p[i,k] * 2^(8(7-p)),     d_p in [-128, 127]
// (the bytes of the 56-bit two's-complement fixed-point value, re-centred so every digit is
// balanced), and   sum_k x_ik x_jk = s_i s_j * sum_{t=2..8} 2^(8(8-t)) G_t,   s = 2^(e-31),
//      G_t[i,j] = sum_{p+q=t} sum_k d_p[i,k] d_q[j,k]
// is EXACT int32 arithmetic on the tensor cores (|G_t| <= 7 * 512 * 2^14 < 2^31).  Digit pairs with
// p+q > 8 are dropped: a zero-mean truncation of ~1e-15 relative to |row_i|_max |row_j|_max, i.e.
// fp64-level.  28 int8 MMAs replace one fp64 MMA and still run ~3x faster than DMMA.
//
// Kernel structure (one persistent CTA per SM, 10 warps, warp-specialised):
//   warp 8  TMA producer: cp.async.bulk.tensor (tensor-map TMA, SASS UTMALDG) of the 7 digit
//           planes of a 64-byte k-chunk of A (128 rows) and B (64 rows) into a 2-stage smem ring,
//           128B-free SWIZZLE_64B layout, completion on mbarriers.
//   warp 9  MMA issuer: one elected lane issues tcgen05.mma.kind::i8 (M=128, N=64, K=32), 56 per
//           k-chunk, accumulating the seven G_t in seven 64-column TMEM accumulators (448 of the
//           512 columns); tcgen05.commit releases smem stages and publishes finished accumulators.
//   warps 0-7 epilogue: tcgen05.ld the int32 accumulators, convert + weight into fp64 registers,
//           release each TMEM accumulator as soon as it is read (so the next tile's MMAs overlap the
//           rest of the epilogue), then C -= s_i s_j * acc on the packed fp64 matrix.
// Replaces the dsyrk/dgemm inside LAPACK dpotrf (AbstractGPs `cholesky(Symmetric(cov(fx)))`).
#include <cuda.h>

#include "sb_common.cuh"

namespace sb {
namespace {

constexpr int OZ_S = 7;              // digit planes
constexpr int OZ_BM = 128, OZ_BN = 64;
constexpr int OZ_THREADS = 320;
constexpr int OZ_EPI_WARPS = 8;
constexpr int OZ_KMAX = 512;         // bytes of K per row in the digit planes (row pitch)

// Pipeline shape: KC = bytes of K per smem stage, STAGES = ring depth.  The MMA floor is 7.5 us per
// 128x64x512 tile and a tile needs 688 KB of operands, i.e. ~92 GB/s per SM: with ~1-2 us of TMA
// latency the ring must keep >= 100-180 KB in flight.  (KC=64, 2 stages) keeps only 84 KB in flight
// and measured 19.6 us/tile (ncu: tensor pipe 39 %, profiles/ncu_ozaki_r2.txt); (KC=32, 5 stages)
// keeps 168 KB in flight inside the same shared memory.
template <int KC, int STAGES>
struct OzCfg {
    static constexpr int A_PLANE = OZ_BM * KC;
    static constexpr int B_PLANE = OZ_BN * KC;
    static constexpr int A_STAGE = OZ_S * A_PLANE;
    static constexpr int B_STAGE = OZ_S * B_PLANE;
    static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
    static constexpr int BAR_BYTES = 16 * STAGES + 16 * OZ_S + 16;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct OzArgs {
    int mode;         // 1: packed SYRK (A == B == the panel), 0: plain  C[M x Ncols] -= A B^T, dense C
    Packed Pk;
    int64_t J0, w;    // packed: first owned trailing block column, column stride (world)
    int64_t total_tiles;
    int kchunks;      // K / KC
    const double* scaleA;  // row scales s_i = 2^(e_i - 31), indexed by plane row of A / B
    const double* scaleB;
    // plain mode: tile (rt, ct) = (t % mtiles, t / mtiles); plane rows rowA0 + 128 rt, rowB0 + 64 ct
    double* C;
    int64_t ldc, mtiles;
    int64_t rowA0, rowB0;
    // shared-memory matrix descriptor fields (runtime so the test harness can probe encodings)
    uint32_t a_kk_adv, b_kk_adv;  // start-address advance (16-byte units) per K=32 step
    uint32_t a_lbo, b_lbo, sbo;   // 16-byte units
    uint32_t layout;              // 3-bit layout_type (4 = SWIZZLE_64B, 0 = none)
    int tma_mode;                 // 0: 3-D SWIZZLE_64B box (KC 64), 1: 4-D un-swizzled interleave (KC 64), 2: 3-D SWIZZLE_32B (KC 32)
    int* dbg;                     // optional: raw int32 accumulators of tile 0  [7][128][64]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin with a watchdog: a protocol bug must trap (error returned to the caller), never hang the box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; it++) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((it & 0xfff) == 0xfff) {
            long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000LL) __trap();  // ~10 s
        }
    }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M=128 N=64 K=32
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// thread <-> its own TMEM lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// same enumeration as gemm_nt.cu's packed-SYRK cursor: owned block columns J0, J0+w, ...; column J
// holds nblk-J row blocks, two 64-wide half tiles per block
struct OzTile {
    int rowA, rowB;    // first plane row of the A (128 rows) / B (64 rows) operand
    double* C;         // top-left element of the 128 x 64 output tile
    int64_t ldc;
};

struct OzCursor {
    int64_t t, J, s0;
    __device__ __forceinline__ void init(const OzArgs& g, int64_t t0) {
        t = t0; J = g.J0; s0 = 0;
        seek(g);
    }
    __device__ __forceinline__ void seek(const OzArgs& g) {
        if (g.mode != 1) return;
        const int64_t nblk = g.Pk.nblk();
        while (t < g.total_tiles && t - s0 >= 2 * (nblk - J)) {
            s0 += 2 * (nblk - J);
            J += g.w;
        }
    }
    __device__ __forceinline__ void advance(const OzArgs& g, int64_t step) {
        t += step;
        seek(g);
    }
    __device__ __forceinline__ OzTile tile(const OzArgs& g) const {
        OzTile o;
        if (g.mode == 1) {
            const int64_t loc = t - s0;
            const int64_t I = J + (loc >> 1);
            const int h = (int)(loc & 1);
            o.rowA = (int)(I * NB);
            o.rowB = (int)(J * NB + h * OZ_BN);
            o.ldc = g.Pk.ld(J);
            o.C = g.Pk.blk(I, J) + (int64_t)(h * OZ_BN) * o.ldc;
        } else {
            const int64_t rt = t % g.mtiles, ct = t / g.mtiles;
            o.rowA = (int)(g.rowA0 + rt * OZ_BM);
            o.rowB = (int)(g.rowB0 + ct * OZ_BN);
            o.ldc = g.ldc;
            o.C = g.C + ct * OZ_BN * g.ldc + rt * OZ_BM;
        }
        return o;
    }
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    // tcgen05 shared-memory matrix descriptor: start address [0,14) (>>4), leading byte offset
    // [16,30) (>>4), stride byte offset [32,46) (>>4), version = 1 at [46,48), layout type [61,64)
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)(lbo & 0x3fff) << 16;
    d |= (uint64_t)(sbo & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

template <int KC, int STAGES>
__global__ void __launch_bounds__(OZ_THREADS, 1)
ozaki_syrk_kernel(const __grid_constant__ OzArgs g, const __grid_constant__ CUtensorMap tmA,
                  const __grid_constant__ CUtensorMap tmB) {
    using Cfg = OzCfg<KC, STAGES>;
    constexpr int OZ_STAGES = STAGES, OZ_KC = KC;
    constexpr int OZ_A_PLANE = Cfg::A_PLANE, OZ_B_PLANE = Cfg::B_PLANE, OZ_A_STAGE = Cfg::A_STAGE;
    constexpr int OZ_STAGE_BYTES = Cfg::STAGE_BYTES;
    extern __shared__ unsigned char oz_smem_raw[];
    const uint32_t raw = smem_u32(oz_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;          // stage buffers: 1024-byte aligned
    const uint32_t bars = base + OZ_STAGES * OZ_STAGE_BYTES;  // 8-byte mbarriers
    // layout of the barrier block: full[STAGES], empty[STAGES], tfull[7], tempty[7], tmem ptr
    const uint32_t full0 = bars, empty0 = bars + 8 * OZ_STAGES, tfull0 = bars + 16 * OZ_STAGES,
                   tempty0 = tfull0 + 8 * OZ_S;
    const uint32_t tmem_slot = tempty0 + 8 * OZ_S;
    unsigned char* gen_base = oz_smem_raw + (base - raw);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(gen_base + OZ_STAGES * OZ_STAGE_BYTES + 16 * OZ_STAGES + 16 * OZ_S);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int t = 0; t < OZ_S; t++) {
            mbar_init(tfull0 + 8 * t, 1);
            mbar_init(tempty0 + 8 * t, OZ_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 9) {  // TMEM: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_ptr;

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            OzCursor cur;
            cur.init(g, blockIdx.x);
            uint32_t n = 0;
            for (; cur.t < g.total_tiles; cur.advance(g, gridDim.x)) {
                const OzTile tl = cur.tile(g);
                const int rowA = tl.rowA, rowB = tl.rowB;
                for (int kc = 0; kc < g.kchunks; kc++, n++) {
                    const uint32_t st = n % OZ_STAGES, ph = (n / OZ_STAGES) & 1;
                    mbar_wait(empty0 + 8 * st, ph ^ 1);
                    const uint32_t fb = full0 + 8 * st;
                    mbar_expect_tx(fb, OZ_STAGE_BYTES);
                    const uint32_t dA = base + st * OZ_STAGE_BYTES, dB = dA + OZ_A_STAGE;
                    if (g.tma_mode != 1) {
                        tma_load_3d(dA, &tmA, kc * OZ_KC, rowA, 0, fb);
                        tma_load_3d(dB, &tmB, kc * OZ_KC, rowB, 0, fb);
                    } else {
                        tma_load_4d(dA, &tmA, 0, rowA, kc * (OZ_KC / 16), 0, fb);
                        tma_load_4d(dB, &tmB, 0, rowB, kc * (OZ_KC / 16), 0, fb);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor, kind::i8: D = S32 (2 @ [4,6)), A = B = INT8 (1 @ [7,10), [10,13)),
            // K-major both, N >> 3 @ [17,23), M >> 4 @ [24,29)
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BN >> 3) << 17) |
                                   ((uint32_t)(OZ_BM >> 4) << 24);
            uint32_t n = 0, it = 0;
            for (int64_t t = blockIdx.x; t < g.total_tiles; t += gridDim.x, it++) {
                for (int kc = 0; kc < g.kchunks; kc++, n++) {
                    const uint32_t st = n % OZ_STAGES, ph = (n / OZ_STAGES) & 1;
                    mbar_wait(full0 + 8 * st, ph);
                    tc_fence_after();
                    const uint32_t sA = base + st * OZ_STAGE_BYTES, sB = sA + OZ_A_STAGE;
                    const uint64_t dA0 = make_desc(sA, g.a_lbo, g.sbo, g.layout);
                    const uint64_t dB0 = make_desc(sB, g.b_lbo, g.sbo, g.layout);
#pragma unroll
                    for (int grp = 0; grp < OZ_S; grp++) {         // grp = p + q - 2
                        if (kc == 0) {                              // accumulator must have been drained
                            mbar_wait(tempty0 + 8 * grp, (it & 1) ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t d = tmem + (uint32_t)grp * OZ_BN;
#pragma unroll
                        for (int p = 0; p <= grp; p++) {            // digit planes p (of A) and q = grp - p (of B)
                            const int q = grp - p;
#pragma unroll
                            for (int kk = 0; kk < OZ_KC / 32; kk++) {
                                const uint64_t da = dA0 + (uint64_t)(p * (OZ_A_PLANE >> 4) + kk * g.a_kk_adv);
                                const uint64_t db = dB0 + (uint64_t)(q * (OZ_B_PLANE >> 4) + kk * g.b_kk_adv);
                                mma_i8(d, da, db, idesc, (kc > 0 || p > 0 || kk > 0) ? 1u : 0u);
                            }
                        }
                        if (kc == g.kchunks - 1) tc_commit(tfull0 + 8 * grp);   // G_{grp+2} of this tile is final
                    }
                    tc_commit(empty0 + 8 * st);  // smem stage free once these MMAs have read it
                }
            }
        }
    } else {
        // ===================== epilogue warps 0..7 =====================
        const int lq = warp & 3, ch = warp >> 2;  // TMEM lane quarter (hardware: warp % 4), column half
        OzCursor cur;
        cur.init(g, blockIdx.x);
        uint32_t it = 0;
        for (; cur.t < g.total_tiles; cur.advance(g, gridDim.x), it++) {
            const OzTile tl = cur.tile(g);
            double acc[32];
#pragma unroll
            for (int c = 0; c < 32; c++) acc[c] = 0.0;
#pragma unroll
            for (int grp = 0; grp < OZ_S; grp++) {
                mbar_wait(tfull0 + 8 * grp, it & 1);
                tc_fence_after();
                int v[32];
                tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(grp * OZ_BN + ch * 32), v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * grp);   // accumulator may be overwritten
                const double wt = (double)(1ull << (8 * (OZ_S - 1 - grp)));   // 2^(8(8 - t)), t = grp + 2
                if (g.dbg != nullptr && cur.t == 0) {
#pragma unroll
                    for (int c = 0; c < 32; c++) g.dbg[(grp * OZ_BM + lq * 32 + lane) * OZ_BN + ch * 32 + c] = v[c];
                }
#pragma unroll
                for (int c = 0; c < 32; c++) acc[c] = fma((double)v[c], wt, acc[c]);
            }
            // C[rows lq*32 + lane, cols ch*32 .. +32 of the tile] -= s_i s_j acc
            const int64_t ldc = tl.ldc;
            const int row = lq * 32 + lane;
            double* cp = tl.C + (int64_t)(ch * 32) * ldc + row;
            const double si = g.scaleA[tl.rowA + row];
            const double* sj = g.scaleB + tl.rowB + ch * 32;
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8) {
                double old[8];
#pragma unroll
                for (int c = 0; c < 8; c++) old[c] = __ldcs(cp + (int64_t)(c0 + c) * ldc);
#pragma unroll
                for (int c = 0; c < 8; c++) cp[(int64_t)(c0 + c) * ldc] = fma(-(si * sj[c0 + c]), acc[c0 + c], old[c]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// ---- digit planes of an operand panel ------------------------------------------------------------
// Source: up to 4 segments of 128 columns each (K = 128 * nseg), element (row block rb, column k, row r)
// of segment q at  base[q] + rb*rbs[q] + k*ld[q] + r  -- covers both the TILED Cholesky panels of
// gemm_nt.cu (rbs = 128*132, ld = 132) and plain column-major panels (rbs = 128, ld = ld).
// Output, for plane row i = out_row_base + rb*128 + r and byte kb = 128*q + k:
//   plane[p][i*512 + kb]  (K-major, 512-byte pitch: the 3-D tensor the TMA box walks),
//   scale[i] = 2^(e_i - 31),  expo[i] = e_i.
__global__ void __launch_bounds__(512)
oz_rowscale_kernel(const OzSrc src, int64_t rb_lo, int64_t out_row_base, double* __restrict__ scale,
                   int* __restrict__ expo) {
    // one CTA per 128-row block; 512 threads = 128 rows x 4 column groups
    __shared__ double red[4][NB];
    const int64_t rb = rb_lo + blockIdx.x;
    const int r = threadIdx.x & 127, cg = threadIdx.x >> 7;
    double m = 0.0;
    for (int q = 0; q < src.nseg; q++) {
        const double* P = src.base[q] + rb * src.rbs[q] + r;
        const int64_t ld = src.ld[q];
        for (int k = cg; k < NB; k += 4) m = fmax(m, fabs(P[(int64_t)k * ld]));
    }
    red[cg][r] = m;
    __syncthreads();
    if (cg == 0) {
        m = fmax(fmax(red[0][r], red[1][r]), fmax(red[2][r], red[3][r]));
        const int64_t i = out_row_base + rb * NB + r;
        const bool ok = m > 1e-280 && m < 1e280;
        const int e = ok ? ilogb(m) + 2 : 0;   // |x| * 2^-e < 0.5
        scale[i] = ok ? scalbn(1.0, e - 31) : 0.0;
        expo[i] = ok ? e : 0x7fffffff;
    }
}

__global__ void __launch_bounds__(256)
oz_slice_kernel(const OzSrc src, int64_t rb_lo, int64_t out_row_base, int64_t plane_rows,
                const int* __restrict__ expo, signed char* __restrict__ planes) {
    // grid: (row blocks, nseg * 4 column chunks of 32); 256 threads = 128 rows x 2 halves of 16 cols
    __shared__ __align__(16) signed char sd[OZ_S][NB][32];
    const int64_t rb = rb_lo + blockIdx.x;
    const int q = blockIdx.y >> 2, kc = blockIdx.y & 3;
    const int r = threadIdx.x & 127, kh = threadIdx.x >> 7;
    const int64_t row0 = out_row_base + rb * NB;
    const int e = expo[row0 + r];
    const int64_t ld = src.ld[q];
    const double* P = src.base[q] + rb * src.rbs[q] + (int64_t)(kc * 32 + kh * 16) * ld + r;
#pragma unroll 4
    for (int kk = 0; kk < 16; kk++) {
        const double x = P[(int64_t)kk * ld];
        long long Z = 0x0000808080808080LL;
        if (e != 0x7fffffff) Z += __double2ll_rn(scalbn(x, 55 - e));
        const int kb = kh * 16 + kk;
        sd[0][r][kb] = (signed char)(Z >> 48);
#pragma unroll
        for (int p = 1; p < OZ_S; p++) sd[p][r][kb] = (signed char)(((Z >> (8 * (OZ_S - 1 - p))) & 0xff) ^ 0x80);
    }
    __syncthreads();
    // 7 planes x 128 rows x 32 bytes: 16-byte stores, two per row
    for (int idx = threadIdx.x; idx < OZ_S * NB * 2; idx += 256) {
        const int p = idx / (NB * 2), rr = (idx >> 1) & 127, hf = idx & 1;
        const int4 v = *reinterpret_cast<const int4*>(&sd[p][rr][hf * 16]);
        *reinterpret_cast<int4*>(planes + ((int64_t)p * plane_rows + row0 + rr) * OZ_KMAX + q * NB + kc * 32 + hf * 16) = v;
    }
}

typedef CUresult (*EncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiled_t g_encode = nullptr;
bool g_oz_attr = false;
int g_oz_sms = 0;

// tma_mode -> (KC, STAGES) instance
#define OZ_DISPATCH(mode, CALL)                                  \
    do {                                                         \
        if ((mode) == 2) { CALL(32, 5); } else { CALL(64, 2); }  \
    } while (0)

int oz_init() {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
        g_encode = (EncodeTiled_t)fn;
    }
    if (!g_oz_attr) {
        if (cudaFuncSetAttribute(ozaki_syrk_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)OzCfg<64, 2>::SMEM) != cudaSuccess) return -2;
        if (cudaFuncSetAttribute(ozaki_syrk_kernel<32, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)OzCfg<32, 5>::SMEM) != cudaSuccess) return -2;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_oz_sms, cudaDevAttrMultiProcessorCount, dev);
        g_oz_attr = true;
    }
    return 0;
}

}  // namespace

size_t oz_planes_bytes(int64_t Np) { return (size_t)OZ_S * Np * OZ_KMAX; }

// Build the two tensor maps (A box: 128 rows, B box: 64 rows) over the digit planes.
int oz_make_maps(signed char* planes, int64_t Np, int tma_mode, OzMaps* out) {
    if (oz_init() != 0) return -1;
    static_assert(sizeof(out->a) >= sizeof(CUtensorMap), "OzMaps too small");
    CUtensorMap* ma = reinterpret_cast<CUtensorMap*>(out->a);
    CUtensorMap* mb = reinterpret_cast<CUtensorMap*>(out->b);
    CUresult r1, r2;
    if (tma_mode == 0 || tma_mode == 2) {
        const cuuint32_t kc = tma_mode == 0 ? 64 : 32;
        const CUtensorMapSwizzle sw = tma_mode == 0 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        cuuint64_t dims[3] = {(cuuint64_t)OZ_KMAX, (cuuint64_t)Np, (cuuint64_t)OZ_S};
        cuuint64_t strides[2] = {(cuuint64_t)OZ_KMAX, (cuuint64_t)Np * OZ_KMAX};
        cuuint32_t estr[3] = {1, 1, 1};
        cuuint32_t boxA[3] = {kc, OZ_BM, OZ_S}, boxB[3] = {kc, OZ_BN, OZ_S};
        r1 = g_encode(ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, planes, dims, strides, boxA, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        r2 = g_encode(mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, planes, dims, strides, boxB, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // un-swizzled "interleave" operand layout: smem [plane][16-byte k column][row][16 B]
        cuuint64_t dims[4] = {16, (cuuint64_t)Np, (cuuint64_t)(OZ_KMAX / 16), (cuuint64_t)OZ_S};
        cuuint64_t strides[3] = {(cuuint64_t)OZ_KMAX, 16, (cuuint64_t)Np * OZ_KMAX};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        cuuint32_t boxA[4] = {16, OZ_BM, 64 / 16, OZ_S}, boxB[4] = {16, OZ_BN, 64 / 16, OZ_S};
        r1 = g_encode(ma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, planes, dims, strides, boxA, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        r2 = g_encode(mb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, planes, dims, strides, boxB, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) ? 0 : -3;
}

// digit planes + row scales of row blocks [rb_lo, rb_lo + nrb) of the source panel
void launch_oz_slice(const OzSrc& src, int64_t rb_lo, int64_t nrb, int64_t out_row_base, int64_t plane_rows,
                     double* scale, int* expo, signed char* planes, cudaStream_t s) {
    if (nrb <= 0 || src.nseg <= 0) return;
    oz_rowscale_kernel<<<(unsigned)nrb, 512, 0, s>>>(src, rb_lo, out_row_base, scale, expo);
    oz_slice_kernel<<<dim3((unsigned)nrb, (unsigned)(src.nseg * 4)), 256, 0, s>>>(src, rb_lo, out_row_base, plane_rows,
                                                                                  expo, planes);
    g_launch_count += 2;
}

// the nseg TILED panels of the Cholesky outer step at block column k0 (row block 0 <-> block row k0+1)
OzSrc oz_src_tiled(const double* const* Pt, int nseg) {
    OzSrc src{};
    src.nseg = nseg;
    for (int q = 0; q < nseg; q++) { src.base[q] = Pt[q]; src.ld[q] = NB + 4; src.rbs[q] = (int64_t)NB * (NB + 4); }
    return src;
}

void oz_default_desc(OzDesc* d, int tma_mode) {
    if (tma_mode == 0) {
        d->a_kk_adv = d->b_kk_adv = 2;  // +32 bytes inside the 64-byte swizzled row
        d->a_lbo = d->b_lbo = 1;        // unused for swizzled K-major
        d->sbo = 32;                    // 8 rows x 64 B
        d->layout = 4;                  // SWIZZLE_64B
    } else if (tma_mode == 2) {
        d->a_kk_adv = d->b_kk_adv = 0;  // one K=32 step per stage
        d->a_lbo = d->b_lbo = 1;
        d->sbo = 16;                    // 8 rows x 32 B
        d->layout = 6;                  // SWIZZLE_32B
    } else {
        d->a_kk_adv = 2 * (OZ_BM * 16 >> 4);  // two 16-byte k columns of 128 rows
        d->b_kk_adv = 2 * (OZ_BN * 16 >> 4);
        d->a_lbo = OZ_BM * 16 >> 4;           // next 16-byte k column
        d->b_lbo = OZ_BN * 16 >> 4;
        d->sbo = 8;                           // next 8-row core matrix: 128 B
        d->layout = 0;
    }
}

static void oz_fill_desc(OzArgs& g, const OzDesc* desc, int tma_mode, int* dbg) {
    g.a_kk_adv = desc->a_kk_adv; g.b_kk_adv = desc->b_kk_adv;
    g.a_lbo = desc->a_lbo; g.b_lbo = desc->b_lbo; g.sbo = desc->sbo; g.layout = desc->layout;
    g.tma_mode = tma_mode;
    g.dbg = dbg;
}

static int oz_launch(OzArgs& g, const OzMaps* mapsA, const OzMaps* mapsB, cudaStream_t s, int reserve_sms) {
    int64_t cap = g_oz_sms - reserve_sms;
    if (cap < 1) cap = 1;
    const int64_t grid = g.total_tiles < cap ? g.total_tiles : cap;
    const CUtensorMap* ma = reinterpret_cast<const CUtensorMap*>(mapsA->a);
    const CUtensorMap* mb = reinterpret_cast<const CUtensorMap*>(mapsB->b);
#define OZ_LAUNCH(KC, ST) ozaki_syrk_kernel<KC, ST><<<(unsigned)grid, OZ_THREADS, OzCfg<KC, ST>::SMEM, s>>>(g, *ma, *mb)
    OZ_DISPATCH(g.tma_mode, OZ_LAUNCH);
#undef OZ_LAUNCH
    g_launch_count++;
    return 0;
}

// A[I, J] -= P_I P_J^T on the packed lower matrix for the owned block columns J in [jlo, jhi), with
// the digit planes / scales produced by launch_oz_slice.  K = 128 * nseg.
int launch_syrk_ozaki(Packed Apk, int64_t k, int nseg, int64_t jlo, int64_t jhi, int rank, int world,
                      const OzMaps* maps, const double* scale, const OzDesc* desc, int tma_mode, cudaStream_t s,
                      int reserve_sms, int* dbg) {
    if (oz_init() != 0) return -1;
    const int64_t nblk = Apk.nblk();
    if (jlo < k + 1) jlo = k + 1;
    if (jhi > nblk) jhi = nblk;
    const int64_t J0 = jlo + ((rank - jlo % world) % world + world) % world;
    const int64_t tiles = syrk_packed_tiles(nblk, k, jlo, jhi, rank, world);
    if (tiles <= 0 || nseg <= 0) return 0;
    OzArgs g{};
    g.mode = 1;
    g.Pk = Apk; g.J0 = J0; g.w = world;
    g.total_tiles = tiles * 2;
    g.kchunks = nseg * NB / (tma_mode == 2 ? 32 : 64);
    g.scaleA = g.scaleB = scale;
    oz_fill_desc(g, desc, tma_mode, dbg);
    return oz_launch(g, maps, maps, s, reserve_sms);
}

// plain product  C[M x Ncols] -= A B^T  (C dense column-major, ldc; M % 128 == 0, Ncols % 64 == 0):
// A = plane rows [rowA0, rowA0 + M) of the (mapsA, scaleA) set, B = plane rows [rowB0, rowB0 + Ncols) of
// (mapsB, scaleB); K = 128 * nseg.  Used by the posterior / VFE matrix-TRSM sweeps.
int launch_gemm_ozaki(double* C, int64_t ldc, int64_t M, int64_t Ncols, int nseg, const OzMaps* mapsA,
                      const double* scaleA, int64_t rowA0, const OzMaps* mapsB, const double* scaleB, int64_t rowB0,
                      const OzDesc* desc, int tma_mode, cudaStream_t s) {
    if (oz_init() != 0) return -1;
    if (M <= 0 || Ncols <= 0 || nseg <= 0) return 0;
    OzArgs g{};
    g.mode = 0;
    g.C = C; g.ldc = ldc; g.mtiles = M / OZ_BM;
    g.total_tiles = (M / OZ_BM) * (Ncols / OZ_BN);
    g.kchunks = nseg * NB / (tma_mode == 2 ? 32 : 64);
    g.scaleA = scaleA; g.scaleB = scaleB;
    g.rowA0 = rowA0; g.rowB0 = rowB0;
    oz_fill_desc(g, desc, tma_mode, nullptr);
    return oz_launch(g, mapsA, mapsB, s, 0);
}

}  // namespace sb
