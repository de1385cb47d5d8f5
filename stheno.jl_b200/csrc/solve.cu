// K4/K5/K6: vector triangular solves on the packed factor, reductions, posterior mean/var
// helpers.  All HBM-bound streaming kernels: every element of L is read exactly once per
// sweep with coalesced, row-contiguous accesses; reductions use warp shuffles.
// Replaces `C.U' \ delta`, `C \ delta`, `logdet(C)`, `sum(abs2, .)`, `diag_At_A`,
// `K_{*f} * alpha` of AbstractGPs.
#include "sb_common.cuh"

namespace sb {
namespace {

constexpr int MAXS = 8;  // right-hand sides processed per pass

// b_k <- invL_kk * b_k   (or invL_kk^T * b_k); one CTA per right-hand side, 4 lanes per row
__global__ void __launch_bounds__(4 * NB)
trsv_diag_kernel(const double* __restrict__ invLk, double* __restrict__ bk, int64_t ldb,
                 int transpose) {
    __shared__ double x[NB];
    double* b = bk + (int64_t)blockIdx.x * ldb;
    const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
    if (threadIdx.x < NB) x[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    double acc = 0.0;
    if (!transpose) {
#pragma unroll 8
        for (int p = part; p <= i; p += 4) acc = fma(invLk[p * NB + i], x[p], acc);
    } else {
#pragma unroll 8
        for (int p = i + part; p < NB; p += 4) acc = fma(invLk[i * NB + p], x[p], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) b[i] = acc;
}

// rows below block k:  b[r] -= sum_c L[r, k*NB + c] * x_k[c]
// CTA = 64 rows x 4 column groups of 32 columns: 1024 CTAs at m = 64K rows, 16 loads in flight
// per thread, so the sweep streams L at HBM speed instead of being latency-bound.
__global__ void __launch_bounds__(256)
gemv_below_kernel(Packed L, int64_t k, double* __restrict__ b, int S) {
    __shared__ double xs[MAXS][NB];
    __shared__ double red[3][MAXS][64];
    const int64_t Np = L.Np;
    for (int idx = threadIdx.x; idx < S * NB; idx += 256) {
        int s = idx / NB, c = idx % NB;
        xs[s][c] = b[(int64_t)s * Np + k * NB + c];
    }
    __syncthreads();
    const int64_t ld = L.ld(k);
    const int rl = threadIdx.x & 63, cg = threadIdx.x >> 6;
    const int64_t lr = (int64_t)blockIdx.x * 64 + rl;  // m is a multiple of 128
    const double* p = L.blk(k + 1, k) + lr + (int64_t)(cg * 32) * ld;
    double acc[MAXS];
#pragma unroll
    for (int s = 0; s < MAXS; s++) acc[s] = 0.0;
#pragma unroll 16
    for (int c = 0; c < 32; c++) {
        double l = p[(int64_t)c * ld];
#pragma unroll
        for (int s = 0; s < MAXS; s++)
            if (s < S) acc[s] = fma(l, xs[s][cg * 32 + c], acc[s]);
    }
    if (cg > 0) {
#pragma unroll
        for (int s = 0; s < MAXS; s++)
            if (s < S) red[cg - 1][s][rl] = acc[s];
    }
    __syncthreads();
    if (cg == 0) {
        int64_t r = (k + 1) * NB + lr;
#pragma unroll
        for (int s = 0; s < MAXS; s++)
            if (s < S) b[(int64_t)s * Np + r] -= acc[s] + red[0][s][rl] + red[1][s][rl] + red[2][s][rl];
    }
}

// b_k[c] -= sum_{r below} L[r, k*NB + c] * x[r]   (transposed product, atomics across CTAs)
// CTA = 256 rows; each lane keeps its 8 x-values in registers, each warp owns 16 columns.
constexpr int GT_ROWS = 256;
__global__ void __launch_bounds__(256)
gemvT_below_kernel(Packed L, int64_t k, double* __restrict__ b, int S) {
    const int64_t Np = L.Np;
    const int64_t ld = L.ld(k);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r0 = (int64_t)blockIdx.x * GT_ROWS;  // m is a multiple of 128; GT_ROWS | 256
    const int64_t m = ld - NB;
    const double* base = L.blk(k + 1, k) + r0 + lane;
    for (int s = 0; s < S; s++) {
        const double* x = b + (int64_t)s * Np + (k + 1) * NB + r0 + lane;
        double xr[GT_ROWS / 32];
#pragma unroll
        for (int i = 0; i < GT_ROWS / 32; i++) xr[i] = (r0 + lane + 32 * i < m) ? x[32 * i] : 0.0;
#pragma unroll 4
        for (int cc = 0; cc < 16; cc++) {
            int c = warp * 16 + cc;
            const double* colp = base + (int64_t)c * ld;
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < GT_ROWS / 32; i++)
                if (r0 + lane + 32 * i < m) acc = fma(colp[32 * i], xr[i], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) atomicAdd(&b[(int64_t)s * Np + k * NB + c], -acc);
        }
    }
}

__global__ void __launch_bounds__(256)
colsumsq_kernel(const double* __restrict__ v, int64_t n, int64_t ld, double* __restrict__ out) {
    const double* p = v + (int64_t)blockIdx.y * ld;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        acc = fma(p[i], p[i], acc);
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[w];
        atomicAdd(&out[blockIdx.y], t);
    }
}

// y[M] += W[M x n] * a[n], columns split across blockIdx.y
constexpr int GN_COLS = 512;
__global__ void __launch_bounds__(256)
gemv_n_kernel(const double* __restrict__ W, int64_t ld, int64_t M, int64_t n,
              const double* __restrict__ a, double* __restrict__ y) {
    __shared__ double as[GN_COLS];
    int64_t c0 = (int64_t)blockIdx.y * GN_COLS;
    int64_t nc = n - c0 < GN_COLS ? n - c0 : GN_COLS;
    for (int i = threadIdx.x; i < nc; i += 256) as[i] = a[c0 + i];
    __syncthreads();
    int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= M) return;
    const double* p = W + c0 * ld + r;
    double acc = 0.0;
#pragma unroll 8
    for (int c = 0; c < nc; c++) acc = fma(p[(int64_t)c * ld], as[c], acc);
    atomicAdd(&y[r], acc);
}

__global__ void __launch_bounds__(256)
rowsumsq_acc_kernel(const double* __restrict__ X, int64_t ld, int64_t M, int64_t ncols,
                    double* __restrict__ acc_out) {
    int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= M) return;
    int64_t c0 = (int64_t)blockIdx.y * 32;
    double acc = 0.0;
#pragma unroll 8
    for (int64_t c = c0; c < c0 + 32 && c < ncols; c++) {
        double v = X[c * ld + r];
        acc = fma(v, v, acc);
    }
    atomicAdd(&acc_out[r], acc);
}

__global__ void axpy1_kernel(double* y, const double* x, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

__global__ void sub_kernel(double* out, const double* a, const double* b, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}

__global__ void unpack_lower_kernel(Packed L, int64_t N, double* __restrict__ out) {
    int64_t c = blockIdx.x;
    int64_t r = (int64_t)blockIdx.y * 256 + threadIdx.x;
    if (r >= N || c >= N) return;
    out[c * N + r] = (r >= c) ? *L.at(r, c) : 0.0;
}

// out[r] += sum_c L[r, c] z[c] over one NB x NB block (I, J), J <= I
__global__ void __launch_bounds__(NB)
trmv_lower_kernel(Packed L, int64_t N, const double* __restrict__ z, double* __restrict__ out,
                  int S) {
    int64_t I = blockIdx.x, J = blockIdx.y;
    if (J > I) return;
    __shared__ double zs[NB];
    const int i = threadIdx.x;
    const double* blk = L.blk(I, J);
    const int64_t ld = L.ld(J);
    for (int s = 0; s < S; s++) {
        __syncthreads();
        zs[i] = z[(int64_t)s * L.Np + J * NB + i];
        __syncthreads();
        int cmax = (I == J) ? i + 1 : NB;
        double acc = 0.0;
        for (int c = 0; c < cmax; c++) acc = fma(blk[(int64_t)c * ld + i], zs[c], acc);
        atomicAdd(&out[(int64_t)s * L.Np + I * NB + i], acc);
    }
}

// W[r, c] *= s[r]
__global__ void __launch_bounds__(256)
rowscale_kernel(double* __restrict__ W, int64_t ld, int64_t rows, const double* __restrict__ s) {
    int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    W[(int64_t)blockIdx.y * ld + r] *= s[r];
}

// y[c] += sum_r W[r, c] * x[r]   (columns are contiguous: one warp per column chunk)
__global__ void __launch_bounds__(256)
gemv_t_kernel(const double* __restrict__ W, int64_t ld, int64_t rows, const double* __restrict__ x,
              double* __restrict__ y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.y * 8 + warp;
    const int64_t r0 = (int64_t)blockIdx.x * 4096;
    const int64_t r1 = r0 + 4096 < rows ? r0 + 4096 : rows;
    const double* col = W + c * ld;
    double acc = 0.0;
#pragma unroll 8
    for (int64_t r = r0 + lane; r < r1; r += 32) acc = fma(col[r], x[r], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(&y[c], acc);
}

// out[c, r] = in[r, c]  (in: rows x cols, ld_in; out: cols x rows, ld_out), 32x32 smem tiles
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ in, int64_t ld_in, double* __restrict__ out, int64_t ld_out) {
    __shared__ double t[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int j = ty; j < 32; j += 8) t[j][tx] = in[(c0 + j) * ld_in + r0 + tx];
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) out[(r0 + j) * ld_out + c0 + tx] = t[tx][j];
}

// packed lower <- dense (n x n, ld), plus `shift` added on the diagonal
__global__ void __launch_bounds__(256)
pack_lower_kernel(Packed L, const double* __restrict__ D, int64_t ld, double shift) {
    int64_t c = blockIdx.x;
    int64_t r = (int64_t)blockIdx.y * 256 + threadIdx.x;
    if (r >= L.Np || r / NB < c / NB) return;
    double v = D[c * ld + r];
    if (r == c) v += shift;
    *L.at(r, c) = v;
}

// dst[(c) * ld + rb*128 + r] = Pt[((row_blk0 + rb) * 128 + c) * 132 + r]
__global__ void __launch_bounds__(128)
untile_panel_kernel(const double* __restrict__ Pt, int64_t row_blk0, double* __restrict__ dst, int64_t ld) {
    const int64_t rb = blockIdx.x;
    const int c = blockIdx.y, r = threadIdx.x;
    dst[(int64_t)c * ld + rb * NB + r] = Pt[((row_blk0 + rb) * NB + c) * (int64_t)(NB + 4) + r];
}

// packed lower += dense symmetric matrix (n x n, ld), lower triangle only
__global__ void __launch_bounds__(256)
add_dense_lower_kernel(Packed L, const double* __restrict__ D, int64_t ld, int64_t n) {
    int64_t c = blockIdx.x;
    int64_t r = (int64_t)blockIdx.y * 256 + threadIdx.x;
    if (r >= n || c >= n || r < c) return;
    *L.at(r, c) += D[c * ld + r];
}

__global__ void add_diag_kernel(Packed L, const double* __restrict__ d, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) *L.at(i, i) += d[i];
}

}  // namespace

void launch_add_dense_lower(Packed L, const double* D, int64_t ld, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    add_dense_lower_kernel<<<dim3((unsigned)n, (unsigned)((n + 255) / 256)), 256, 0, st>>>(L, D, ld, n);
    g_launch_count++;
}

void launch_add_diag(Packed L, const double* d, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    add_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(L, d, n);
    g_launch_count++;
}

void launch_untile_panel(const double* Pt, int64_t row_blk0, int64_t nrow_blks, double* dst, int64_t ld,
                         cudaStream_t st) {
    if (nrow_blks <= 0) return;
    untile_panel_kernel<<<dim3((unsigned)nrow_blks, NB), NB, 0, st>>>(Pt, row_blk0, dst, ld);
    g_launch_count++;
}

void launch_rowscale(double* W, int64_t ld, int64_t rows, int64_t cols, const double* s, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;
    rowscale_kernel<<<dim3((unsigned)((rows + 255) / 256), (unsigned)cols), 256, 0, st>>>(W, ld, rows, s);
    g_launch_count++;
}

void launch_gemv_t(const double* W, int64_t ld, int64_t rows, int64_t cols, const double* x, double* y,
                   cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;  // cols must be a multiple of 8
    gemv_t_kernel<<<dim3((unsigned)((rows + 4095) / 4096), (unsigned)(cols / 8)), 256, 0, st>>>(W, ld, rows, x, y);
    g_launch_count++;
}

void launch_transpose(const double* in, int64_t ld_in, int64_t rows, int64_t cols, double* out, int64_t ld_out,
                      cudaStream_t st) {
    // rows, cols multiples of 32
    transpose_kernel<<<dim3((unsigned)(rows / 32), (unsigned)(cols / 32)), 256, 0, st>>>(in, ld_in, out, ld_out);
    g_launch_count++;
}

void launch_pack_lower(Packed L, const double* D, int64_t ld, double shift, cudaStream_t st) {
    pack_lower_kernel<<<dim3((unsigned)L.Np, (unsigned)((L.Np + 255) / 256)), 256, 0, st>>>(L, D, ld, shift);
    g_launch_count++;
}

void launch_trsv_diag(const double* invLk, double* bk, int64_t ldb, int S, bool transpose,
                      cudaStream_t s) {
    trsv_diag_kernel<<<S, 4 * NB, 0, s>>>(invLk, bk, ldb, transpose ? 1 : 0);
    g_launch_count++;
}

void launch_gemv_below(Packed L, int64_t k, double* b, int S, cudaStream_t s) {
    int64_t m = L.ld(k) - NB;
    if (m <= 0) return;
    gemv_below_kernel<<<(unsigned)(m / 64), 256, 0, s>>>(L, k, b, S);
    g_launch_count++;
}

void launch_gemvT_below(Packed L, int64_t k, double* b, int S, cudaStream_t s) {
    int64_t m = L.ld(k) - NB;
    if (m <= 0) return;
    gemvT_below_kernel<<<(unsigned)((m + GT_ROWS - 1) / GT_ROWS), 256, 0, s>>>(L, k, b, S);
    g_launch_count++;
}

void launch_colsumsq(const double* v, int64_t n, int64_t ld, int S, double* out, cudaStream_t s) {
    cudaMemsetAsync(out, 0, sizeof(double) * S, s);
    int64_t gx = (n + 255) / 256;
    if (gx > 1024) gx = 1024;
    colsumsq_kernel<<<dim3((unsigned)gx, S), 256, 0, s>>>(v, n, ld, out);
    g_launch_count++;
}

void launch_gemv_n(const double* W, int64_t ld, int64_t M, int64_t n, const double* a, double* y,
                   cudaStream_t s) {
    cudaMemsetAsync(y, 0, sizeof(double) * M, s);
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)((n + GN_COLS - 1) / GN_COLS));
    gemv_n_kernel<<<grid, 256, 0, s>>>(W, ld, M, n, a, y);
    g_launch_count++;
}

void launch_rowsumsq_acc(const double* X, int64_t ld, int64_t M, int64_t ncols, double* acc,
                         cudaStream_t s) {
    dim3 grid((unsigned)((M + 255) / 256), (unsigned)((ncols + 31) / 32));
    rowsumsq_acc_kernel<<<grid, 256, 0, s>>>(X, ld, M, ncols, acc);
    g_launch_count++;
}

void launch_axpy1(double* y, const double* x, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    axpy1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, x, n);
    g_launch_count++;
}

void launch_sub(double* out, const double* a, const double* b, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    sub_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, a, b, n);
    g_launch_count++;
}

void launch_unpack_lower(Packed L, int64_t N, double* out, cudaStream_t s) {
    dim3 grid((unsigned)N, (unsigned)((N + 255) / 256));
    unpack_lower_kernel<<<grid, 256, 0, s>>>(L, N, out);
    g_launch_count++;
}

void launch_trmv_lower(Packed L, int64_t N, const double* z, double* out, int S, cudaStream_t s) {
    cudaMemsetAsync(out, 0, sizeof(double) * L.Np * S, s);
    dim3 grid((unsigned)L.nblk(), (unsigned)L.nblk());
    trmv_lower_kernel<<<grid, NB, 0, s>>>(L, N, z, out, S);
    g_launch_count++;
}

}  // namespace sb
