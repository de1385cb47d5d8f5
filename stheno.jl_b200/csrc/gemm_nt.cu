// K3: fp64 tensor-core NT contraction   C = beta*C + alpha * A * B^T   (all column-major).
//
// This one kernel is the whole O(N^3) of the path: the Cholesky trailing update (SYRK on the
// packed lower matrix), the panel TRSM (as a product with the explicit inverse of the
// diagonal block), the posterior matrix TRSM sweep and the posterior covariance downdate.
// It replaces LAPACK dpotrf's dsyrk/dgemm/dtrsm inside `cholesky(Symmetric(K + Sigma_y))` and
// the `C.U' \ K_fx` solves of AbstractGPs.
//
// sm_100a notes.  tcgen05.mma has no fp64 kind, so fp64 tensor math is the DMMA path
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  Operand k-slabs are staged global->shared by the
// TMA engine with 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx -> SASS UBLKCP) into a
// 4-stage ring; one mbarrier per stage.  Shared tiles are [k][row] with a 4-double pad so the
// m8n8k4 fragment loads (8 rows x 4 k per operand) are bank-conflict free.  The MMA is issued
// "transposed" (m <-> C columns, n <-> C rows) so each thread owns 2 consecutive rows of C and
// the epilogue uses 16-byte loads/stores.  CTA tile 128x64, 8 warps of 32x32, 2 CTAs per SM so
// one CTA's C read-modify-write epilogue overlaps the other's MMA main loop.
#include "sb_common.cuh"

namespace sb {
namespace {

constexpr int BM = 128;  // C rows per CTA
constexpr int BN = 64;   // C cols per CTA
constexpr int KC = 16;   // k-slab per pipeline stage
constexpr int STAGES = 4;
constexpr int LDA_S = BM + 4;
constexpr int LDB_S = BN + 4;
constexpr int THREADS = 256;
constexpr size_t SMEM_BYTES = (size_t)STAGES * KC * (LDA_S + LDB_S) * 8 + 2 * STAGES * 8 + 8 * 128;

struct GemmArgs {
    int mode;  // 0 plain, 1 packed SYRK
    const double* A;
    int64_t lda;
    const double* B;
    int64_t ldb;
    double* C;
    int64_t ldc;
    int64_t mtiles;  // plain: number of row tiles
    int64_t total_tiles;
    int64_t K;
    double alpha, beta;
    // packed SYRK
    Packed Pk;
    int64_t k;      // panel index
    int64_t J0;     // first owned block column >= jlo
    int64_t w;      // column stride (world)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

struct TilePtrs {
    const double* A;
    const double* B;
    double* C;
    int64_t lda, ldb, ldc;
};

// Walks the tiles blockIdx.x, +gridDim.x, ... of one launch.  Packed SYRK tiles are enumerated
// block column by block column (columns J0, J0+w, ... ; column J holds nblk-J row blocks; two
// 64-wide half tiles per block), so advancing is a couple of integer compares -- no division,
// no square root.
struct TileCursor {
    int64_t t;     // linear tile index
    int64_t J;     // packed: current block column
    int64_t s0;    // packed: linear index of the first (half-)tile of column J
    __device__ __forceinline__ void init(const GemmArgs& g, int64_t t0) {
        t = t0;
        J = g.J0;
        s0 = 0;
        if (g.mode == 1) seek(g);
    }
    __device__ __forceinline__ void seek(const GemmArgs& g) {
        const int64_t nblk = g.Pk.nblk();
        while (t - s0 >= 2 * (nblk - J)) {
            s0 += 2 * (nblk - J);
            J += g.w;
        }
    }
    __device__ __forceinline__ void advance(const GemmArgs& g, int64_t step) {
        t += step;
        if (g.mode == 1 && t < g.total_tiles) seek(g);
    }
    __device__ __forceinline__ TilePtrs ptrs(const GemmArgs& g) const {
        TilePtrs p;
        if (g.mode == 0) {
            int64_t rt = t % g.mtiles, ct = t / g.mtiles;
            p.lda = g.lda; p.ldb = g.ldb; p.ldc = g.ldc;
            p.A = g.A + rt * BM;
            p.B = g.B + ct * BN;
            p.C = g.C + ct * BN * g.ldc + rt * BM;
        } else {
            const int64_t loc = t - s0;
            const int64_t I = J + (loc >> 1);
            const int h = (int)(loc & 1);
            const int64_t m = g.Pk.Np - (g.k + 1) * NB;  // panel rows below diagonal block k
            p.lda = p.ldb = m;
            p.A = g.A + (I - g.k - 1) * NB;
            p.B = g.A + (J - g.k - 1) * NB + h * BN;
            p.ldc = g.Pk.ld(J);
            p.C = g.Pk.blk(I, J) + (int64_t)h * BN * p.ldc;
        }
        return p;
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent CTAs (2 per SM): each walks tiles blockIdx.x, +gridDim.x, ...  The operand ring is
// addressed by a chunk counter that runs ACROSS tiles, so the TMA engine is already filling the
// next tile's first k-slabs while the warps are still in the current tile's epilogue.  No
// CTA-wide barrier in the steady state: full[] (TMA -> warps, tx-count) and empty[] (8 warps ->
// producer lane) mbarriers only.
__global__ void __launch_bounds__(THREADS, 2) gemm_nt_kernel(const __grid_constant__ GemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + STAGES * KC * LDA_S;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * KC * LDB_S);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int nchunks = (int)(g.K / KC);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], THREADS / 32);  // one arrive.expect_tx per issuing warp
            mbar_init(&empty[s], THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // ---- distributed producer: lane 0 of EVERY warp issues its 4 of the 32 bulk copies of a
    // k-slab (a bulk copy costs its issuing thread ~30 ns -- tools/mb_bulk.cu -- so one lane
    // issuing all 32 sat on warp 0's critical path).  Lane 0 of warp 0 additionally arms the
    // slab's mbarrier with the byte count.  Each warp tracks the producer cursor redundantly
    // (identical values in all warps), in registers of lane 0 only via a small shared struct.
    struct ProducerState {
        TileCursor cur;
        TilePtrs pp;
        int pc, pslot;
        uint32_t pphase;  // parity to wait for on empty[pslot] before refilling it
        int pfirst;       // first pass over the ring: slots are fresh, no wait
        int pdone;
    };
    ProducerState* ps = reinterpret_cast<ProducerState*>(empty + STAGES) + warp;
    auto issue_next = [&]() {  // lane 0 of each warp
        ProducerState st = *ps;
        if (!st.pfirst) mbar_wait(&empty[st.pslot], st.pphase);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[st.pslot], (BM + BN) * 2 * 8);  // this warp's 2 + 2 columns
        const double* srcA = st.pp.A + ((int64_t)st.pc * KC + warp * 2) * st.pp.lda;
        const double* srcB = st.pp.B + ((int64_t)st.pc * KC + warp * 2) * st.pp.ldb;
        double* dA = sA + (st.pslot * KC + warp * 2) * LDA_S;
        double* dB = sB + (st.pslot * KC + warp * 2) * LDB_S;
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            bulk_g2s(dA + kk * LDA_S, srcA + kk * st.pp.lda, BM * 8, &full[st.pslot]);
            bulk_g2s(dB + kk * LDB_S, srcB + kk * st.pp.ldb, BN * 8, &full[st.pslot]);
        }
        if (++st.pslot == STAGES) {
            st.pslot = 0;
            if (st.pfirst) st.pfirst = 0; else st.pphase ^= 1;
        }
        if (++st.pc == nchunks) {
            st.pc = 0;
            st.cur.advance(g, gridDim.x);
            st.pdone = st.cur.t >= g.total_tiles;
            if (!st.pdone) st.pp = st.cur.ptrs(g);
        }
        *ps = st;
    };
    const bool producer = (lane == 0);
    if (producer) {
        ProducerState st;
        st.cur.init(g, blockIdx.x);
        st.pc = 0; st.pslot = 0; st.pphase = 0; st.pfirst = 1;
        st.pdone = st.cur.t >= g.total_tiles;
        if (!st.pdone) st.pp = st.cur.ptrs(g);
        *ps = st;
        for (int c = 0; c < STAGES - 1 && !ps->pdone; c++) issue_next();
    }

    const int wr = warp & 3, wc = warp >> 2;  // 4 warps along rows, 2 along cols; 32x32 each
    const double alpha = g.alpha, beta = g.beta;
    const double* a_base = sA + wr * 32 + gid + tig * LDA_S;
    const double* b_base = sB + wc * 32 + gid + tig * LDB_S;
    int slot = 0;
    uint32_t phase = 0;
    TileCursor cur;
    cur.init(g, blockIdx.x);
    for (; cur.t < g.total_tiles; cur.advance(g, gridDim.x)) {
        double acc[4][4][2];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[j][i][0] = acc[j][i][1] = 0.0;

        for (int c = 0; c < nchunks; c++) {
            // keep the ring STAGES-1 slabs ahead: refill the slot released by the previous chunk
            if (producer && !ps->pdone) issue_next();
            __syncwarp();
            mbar_wait(&full[slot], phase);
            const double* a = a_base + slot * KC * LDA_S;
            const double* b = b_base + slot * KC * LDB_S;
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; k4++) {
                double rf[4], cf[4];
#pragma unroll
                for (int i = 0; i < 4; i++) rf[i] = a[k4 * 4 * LDA_S + i * 8];
#pragma unroll
                for (int j = 0; j < 4; j++) cf[j] = b[k4 * 4 * LDB_S + j * 8];
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int i = 0; i < 4; i++) dmma(acc[j][i], cf[j], rf[i]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            if (++slot == STAGES) { slot = 0; phase ^= 1; }
        }

        // epilogue: thread owns rows (2*tig, 2*tig+1) of column gid in each 8x8 fragment
        TilePtrs tp = cur.ptrs(g);
        double* cbase = tp.C + (int64_t)(wc * 32 + gid) * tp.ldc + wr * 32 + 2 * tig;
        const int64_t ldc = tp.ldc;
        if (beta != 0.0) {
#pragma unroll
            for (int jh = 0; jh < 2; jh++) {  // 8 x 16-byte loads in flight per thread
                double2 old[2][4];
#pragma unroll
                for (int jj = 0; jj < 2; jj++)
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        old[jj][i] = __ldcs(reinterpret_cast<const double2*>(
                            cbase + (int64_t)(jh * 2 + jj) * 8 * ldc + i * 8));
#pragma unroll
                for (int jj = 0; jj < 2; jj++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int j = jh * 2 + jj;
                        double2 o;
                        o.x = fma(alpha, acc[j][i][0], beta * old[jj][i].x);
                        o.y = fma(alpha, acc[j][i][1], beta * old[jj][i].y);
                        *reinterpret_cast<double2*>(cbase + (int64_t)j * 8 * ldc + i * 8) = o;
                    }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    double2 o = make_double2(alpha * acc[j][i][0], alpha * acc[j][i][1]);
                    *reinterpret_cast<double2*>(cbase + (int64_t)j * 8 * ldc + i * 8) = o;
                }
        }
    }
}

int g_num_sms = 0;
bool g_attr_set = false;
void ensure_attr() {
    if (!g_attr_set) {
        cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)SMEM_BYTES);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        g_attr_set = true;
    }
}

}  // namespace

void launch_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                    int64_t ldc, int64_t M, int64_t Ncols, int64_t K, double alpha, double beta,
                    cudaStream_t s) {
    if (M <= 0 || Ncols <= 0) return;
    ensure_attr();
    GemmArgs g{};
    g.mode = 0;
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
    g.mtiles = M / BM;
    g.K = K; g.alpha = alpha; g.beta = beta;
    int64_t tiles = (M / BM) * (Ncols / BN);
    g.total_tiles = tiles;
    int64_t grid = tiles < 2 * g_num_sms ? tiles : 2 * g_num_sms;
    gemm_nt_kernel<<<(unsigned)grid, THREADS, SMEM_BYTES, s>>>(g);
    g_launch_count++;
}

int64_t syrk_packed_tiles(int64_t nblk, int64_t k, int64_t jlo, int64_t jhi, int rank, int world) {
    if (jlo < k + 1) jlo = k + 1;
    if (jhi > nblk) jhi = nblk;
    int64_t J0 = jlo + ((rank - jlo % world) % world + world) % world;
    int64_t tiles = 0;
    for (int64_t J = J0; J < jhi; J += world) tiles += nblk - J;
    return tiles;
}

void launch_syrk_packed(Packed Apk, int64_t k, const double* P, int64_t K, int64_t jlo, int64_t jhi,
                        int rank, int world, cudaStream_t s) {
    int64_t nblk = Apk.nblk();
    if (jlo < k + 1) jlo = k + 1;
    if (jhi > nblk) jhi = nblk;
    int64_t J0 = jlo + ((rank - jlo % world) % world + world) % world;
    int64_t tiles = syrk_packed_tiles(nblk, k, jlo, jhi, rank, world);
    if (tiles <= 0) return;
    ensure_attr();
    GemmArgs g{};
    g.mode = 1;
    g.A = P;
    g.K = K; g.alpha = -1.0; g.beta = 1.0;
    g.Pk = Apk; g.k = k; g.J0 = J0; g.w = world;
    g.total_tiles = tiles * 2;
    int64_t grid = g.total_tiles < 2 * g_num_sms ? g.total_tiles : 2 * g_num_sms;
    gemm_nt_kernel<<<(unsigned)grid, THREADS, SMEM_BYTES, s>>>(g);
    g_launch_count++;
}

}  // namespace sb
