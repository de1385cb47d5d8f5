// K2: Cholesky of one NB x NB diagonal block, entirely in shared memory, plus the explicit
// inverse of the resulting triangle (used to turn the panel TRSM and every later block
// triangular solve into tensor-core products) and the block's share of logdet.
// Replaces the unblocked dpotf2 step inside LAPACK dpotrf (AbstractGPs
// `cholesky(Symmetric(cov(fx)))`).  Latency-bound, one CTA; it is kept off
// the critical path by look-ahead in the driver loop (api.cu).
#include "sb_common.cuh"

namespace sb {
namespace {

constexpr int LDS = NB + 4;
constexpr int PT = 512;
constexpr size_t POTRF_SMEM = (size_t)NB * LDS * 8 + NB * 8 + 16;

__global__ void __launch_bounds__(PT, 1)
potrf_inv_kernel(Packed A, int64_t k, int64_t N, double* __restrict__ invL,
                 double* __restrict__ logdet_blk, long long* __restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* s = reinterpret_cast<double*>(smem_raw);   // s[c*LDS + r]
    double* col = s + NB * LDS;                        // NB
    int* bad = reinterpret_cast<int*>(col + NB);

    const int tid = threadIdx.x;
    double* Akk = A.blk(k, k);
    const int64_t ld = A.ld(k);

    // load lower triangle (upper = 0)
    for (int idx = tid; idx < NB * NB; idx += PT) {
        int r = idx % NB, c = idx / NB;
        s[c * LDS + r] = (r >= c) ? Akk[(int64_t)c * ld + r] : 0.0;
    }
    if (tid == 0) *bad = 0;
    __syncthreads();

    // ---- right-looking unblocked Cholesky --------------------------------------------------
    const int ri = tid % NB, lg = tid / NB;  // row, column group (PT/NB = 4 groups)
    for (int j = 0; j < NB; j++) {
        double d = s[j * LDS + j];
        __syncthreads();
        if (!(d > 0.0)) {  // also catches NaN
            if (tid == 0 && *bad == 0) *bad = j + 1;
            d = 1.0;  // keep going with finite garbage; info reports the failure
        }
        double piv = sqrt(d);
        if (tid == j) s[j * LDS + j] = piv;
        if (tid > j && tid < NB) s[j * LDS + tid] /= piv;
        __syncthreads();
        if (ri > j) {
            const double lij = s[j * LDS + ri];
            int l = j + 1 + lg;
            for (; l + 12 <= ri; l += 16) {  // 4 independent updates in flight
                double a0 = s[j * LDS + l], a1 = s[j * LDS + l + 4], a2 = s[j * LDS + l + 8],
                       a3 = s[j * LDS + l + 12];
                double c0 = s[l * LDS + ri], c1 = s[(l + 4) * LDS + ri], c2 = s[(l + 8) * LDS + ri],
                       c3 = s[(l + 12) * LDS + ri];
                s[l * LDS + ri] = c0 - lij * a0;
                s[(l + 4) * LDS + ri] = c1 - lij * a1;
                s[(l + 8) * LDS + ri] = c2 - lij * a2;
                s[(l + 12) * LDS + ri] = c3 - lij * a3;
            }
            for (; l <= ri; l += 4) s[l * LDS + ri] -= lij * s[j * LDS + l];
        }
        __syncthreads();
    }

    // write L_kk back (clean upper triangle), per-block logdet share and info
    for (int idx = tid; idx < NB * NB; idx += PT) {
        int r = idx % NB, c = idx / NB;
        Akk[(int64_t)c * ld + r] = s[c * LDS + r];
    }
    if (tid < 32) {
        double acc = 0.0;
        for (int j = tid; j < NB; j += 32) acc += log(s[j * LDS + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (tid == 0) {
            logdet_blk[k] = 2.0 * acc;
            if (*bad != 0) {
                long long v = (long long)(k * NB + *bad);
                // keep the FIRST failing pivot (smallest index); 0 means "ok so far"
                long long old = *info;
                if (old == 0 || v < old) *info = v;
            }
        }
    }
    __syncthreads();

    // ---- in-place inverse of the lower triangle (unblocked trtri, columns right to left) ----
    // X[j][j] = 1/L[j][j];  X[i][j] = -X[j][j] * sum_{p=j+1..i} X[i][p] * L[p][j]
    const int row = tid >> 2, part = tid & 3;
    for (int j = NB - 1; j >= 0; j--) {
        if (tid < NB) col[tid] = s[j * LDS + tid];  // original column j of L
        __syncthreads();
        double xjj = 1.0 / col[j];
        double acc = 0.0;
        if (row > j) {
            double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
            int p = j + 1 + part;
            for (; p + 12 <= row; p += 16) {
                b0 = fma(s[p * LDS + row], col[p], b0);
                b1 = fma(s[(p + 4) * LDS + row], col[p + 4], b1);
                b2 = fma(s[(p + 8) * LDS + row], col[p + 8], b2);
                b3 = fma(s[(p + 12) * LDS + row], col[p + 12], b3);
            }
            for (; p <= row; p += 4) b0 = fma(s[p * LDS + row], col[p], b0);
            acc = (b0 + b1) + (b2 + b3);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (part == 0) {
            if (row > j) s[j * LDS + row] = -xjj * acc;
            if (row == j) s[j * LDS + j] = xjj;
        }
        __syncthreads();
    }
    double* out = invL + k * (int64_t)NB * NB;
    for (int idx = tid; idx < NB * NB; idx += PT) {
        int r = idx % NB, c = idx / NB;
        out[idx] = (r >= c) ? s[c * LDS + r] : 0.0;
    }
}

bool g_attr = false;

}  // namespace

void launch_potrf_inv(Packed A, int64_t k, int64_t N, double* invL, double* logdet_blk,
                      long long* info, cudaStream_t st) {
    if (!g_attr) {
        cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)POTRF_SMEM);
        g_attr = true;
    }
    potrf_inv_kernel<<<1, PT, POTRF_SMEM, st>>>(A, k, N, invL, logdet_blk, info);
    g_launch_count++;
}

}  // namespace sb
