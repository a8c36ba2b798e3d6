// limbo_b200/csrc/potrf.cu — blocked right-looking Cholesky, trailing update on
// fp64 tensor cores (DMMA).
//
// Replaces `_matrixL = Eigen::LLT<Eigen::MatrixXd>(_kernel).matrixL()`
// (model/gp.hpp:565).  Per 128-column panel k:
//   1. potf2_inv_kernel : factor the 128x128 diagonal block in one CTA
//                         (32x32 sub-blocks in registers via warp shuffles) and
//                         form its inverse (kept in invD for every later
//                         triangular solve);
//   2. trsm_panel_kernel: L[i,k] = A[i,k] * inv(L[k,k])^T   (DMMA GEMM)
//   3. syrk_kernel      : A[i,j] -= L[i,k] * L[j,k]^T, k < j <= i (DMMA GEMM)
// A non-positive pivot is reported LAPACK-style through info (the reference
// never checks Eigen's info()).
#include "gemm.cuh"

namespace {

constexpr int PS = LB_TILE + 1;   // pitch of the diagonal block in smem
constexpr int XB = 32 * 33;       // one 32x32 inverse block, pitch 33
constexpr size_t POTF2_SMEM = (size_t)(LB_TILE * PS + 10 * XB + LB_TILE) * sizeof(double);

__device__ __forceinline__ int blk(int ib, int jb) { return ib * (ib + 1) / 2 + jb; }

__global__ void __launch_bounds__(256, 1)
potf2_inv_kernel(double* __restrict__ L, int64_t ld, int k, double* __restrict__ invD, int* __restrict__ info, int do_factor)
{
    extern __shared__ __align__(16) double smem[];
    double* S = smem;                         // [128][PS]
    double* Xb = smem + LB_TILE * PS;         // 10 blocks [32][33]
    double* sInv = Xb + 10 * XB;              // [128] reciprocal diagonal
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t k0 = (int64_t)k * LB_TILE;
    double* Lkk = L + k0 + k0 * ld;

    for (int idx = tid; idx < LB_TILE * LB_TILE; idx += 256) {
        int r = idx & 127, c = idx >> 7;
        S[r * PS + c] = (c <= r) ? Lkk[r + (int64_t)c * ld] : 0.0;
    }
    __syncthreads();
    if (!do_factor) { // block already factored (incremental update): only (re)build its inverse
        if (tid < LB_TILE) sInv[tid] = 1.0 / S[tid * PS + tid];
        __syncthreads();
    }

    for (int jb = 0; jb < (do_factor ? 4 : 0); ++jb) {
        const int c0 = 32 * jb;
        // (1) 32x32 diagonal sub-block, one row per lane, right-looking
        if (warp == 0) {
            double a[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) a[c] = S[(c0 + lane) * PS + c0 + c];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                double ajj = __shfl_sync(0xffffffffu, a[j], j);
                if (!(ajj > 0.0) && lane == 0) atomicCAS(info, 0, (int)(k0 + c0 + j + 1));
                double d = sqrt(ajj);
                double inv = 1.0 / d;
                double lij = (lane > j) ? a[j] * inv : ((lane == j) ? d : 0.0);
                a[j] = lij;
                if (lane == j) sInv[c0 + j] = inv;
#pragma unroll
                for (int kk = j + 1; kk < 32; ++kk) {
                    double lkj = __shfl_sync(0xffffffffu, lij, kk);
                    a[kk] = fma(-lij, lkj, a[kk]);
                }
            }
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (c <= lane) S[(c0 + lane) * PS + c0 + c] = a[c];
        }
        __syncthreads();
        const int nrem = LB_TILE - c0 - 32;
        // (2) rows below: x = a * Ld^-T (forward substitution along the row)
        if (tid < nrem) {
            const int r = c0 + 32 + tid;
            double x[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = S[r * PS + c0 + c];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                double s = x[c];
#pragma unroll
                for (int kk = 0; kk < c; ++kk) s = fma(-x[kk], S[(c0 + c) * PS + c0 + kk], s);
                x[c] = s * sInv[c0 + c];
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) S[r * PS + c0 + c] = x[c];
        }
        __syncthreads();
        // (3) trailing update inside the block (lower part)
        for (int idx = tid; idx < nrem * nrem; idx += 256) {
            const int r = c0 + 32 + idx % nrem, c = c0 + 32 + idx / nrem;
            if (c <= r) {
                double s = S[r * PS + c];
#pragma unroll 8
                for (int kk = 0; kk < 32; ++kk) s = fma(-S[r * PS + c0 + kk], S[c * PS + c0 + kk], s);
                S[r * PS + c] = s;
            }
        }
        __syncthreads();
    }

    // (4) inverses of the four 32x32 diagonal sub-blocks, one column per lane
    if (warp < 4) {
        const int c0 = 32 * warp;
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int kk = 0; kk < i; ++kk) s = fma(-S[(c0 + i) * PS + c0 + kk], x[kk], s);
            x[i] = s * sInv[c0 + i];
        }
        double* X = Xb + blk(warp, warp) * XB;
#pragma unroll
        for (int i = 0; i < 32; ++i) X[i * 33 + lane] = (i >= lane) ? x[i] : 0.0;
    }
    __syncthreads();
    // (5) off-diagonal inverse blocks, block row by block row:
    //     X[ib,jb] = -X[ib,ib] * sum_{kb=jb}^{ib-1} L[ib,kb] X[kb,jb]
    for (int ib = 1; ib < 4; ++ib) {
        for (int o = tid; o < 1024 * ib; o += 256) {
            const int r = o & 31, c = (o >> 5) & 31, jb = o >> 10;
            double s = 0.0;
            for (int kb = jb; kb < ib; ++kb) {
                const double* Xk = Xb + blk(kb, jb) * XB;
#pragma unroll 8
                for (int kk = 0; kk < 32; ++kk) s = fma(S[(32 * ib + r) * PS + 32 * kb + kk], Xk[kk * 33 + c], s);
            }
            S[(32 * jb + r) * PS + 32 * ib + c] = s; // temp T in the free upper block (jb, ib)
        }
        __syncthreads();
        for (int o = tid; o < 1024 * ib; o += 256) {
            const int r = o & 31, c = (o >> 5) & 31, jb = o >> 10;
            const double* Xd = Xb + blk(ib, ib) * XB;
            double s = 0.0;
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) s = fma(Xd[r * 33 + kk], S[(32 * jb + kk) * PS + 32 * ib + c], s);
            Xb[blk(ib, jb) * XB + r * 33 + c] = -s;
        }
        __syncthreads();
    }
    // (6) write back L[k,k] (clean lower) and inv(L[k,k])
    double* inv_out = invD + (int64_t)k * LB_TILE * LB_TILE;
    for (int idx = tid; idx < LB_TILE * LB_TILE; idx += 256) {
        int r = idx & 127, c = idx >> 7;
        Lkk[r + (int64_t)c * ld] = (c <= r) ? S[r * PS + c] : 0.0;
        inv_out[r + c * LB_TILE] = (c <= r) ? Xb[blk(r >> 5, c >> 5) * XB + (r & 31) * 33 + (c & 31)] : 0.0;
    }
}

// L[i,k] <- A[i,k] * inv(L[k,k])^T for i = k+1 .. T-1
__global__ void __launch_bounds__(lbg::THREADS, 1)
trsm_panel_kernel(double* __restrict__ L, int64_t ld, int k, const double* __restrict__ invD)
{
    extern __shared__ __align__(16) double smem[];
    const int i = k + 1 + blockIdx.x;
    double* A = L + (int64_t)i * LB_TILE + (int64_t)k * LB_TILE * ld;
    const double* B = invD + (int64_t)k * LB_TILE * LB_TILE; // B(kk,n) = inv[n + kk*128]
    lbg::Acc<128> acc;
    acc.zero();
    lbg::mainloop<128, false, false>(acc, A, ld, B, LB_TILE, LB_TILE, smem);
    lbg::for_each_acc<128>(acc, [&](int r, int c, double v) { A[r + (int64_t)c * ld] = v; });
}

// Trailing update with the panel block columns [kb, kb + kd):
//   A[i,j] -= L[i, kb:kb+kd] L[j, kb:kb+kd]^T   for j in [j0, j0 + nc), j <= i < T
// (kd = 1: one 128-column panel, K = 128; kd = 2: two panels at once, K = 256 —
// half the C traffic and half the tile prologues per flop).  C is preloaded into
// the accumulators so its latency overlaps the operand pipeline's prologue.
__global__ void __launch_bounds__(lbg::THREADS, 1)
syrk_kernel(double* __restrict__ L, int64_t ld, int kb, int kd, int j0, int nc, int T)
{
    extern __shared__ __align__(16) double smem[];
    int idx = blockIdx.x, c = 0;
    while (c < nc && idx >= T - j0 - c) { idx -= T - j0 - c; ++c; }
    const int j = j0 + c, i = j + idx;
    const double* A = L + (int64_t)i * LB_TILE + (int64_t)kb * LB_TILE * ld;
    const double* B = L + (int64_t)j * LB_TILE + (int64_t)kb * LB_TILE * ld;
    double* C = L + (int64_t)i * LB_TILE + (int64_t)j * LB_TILE * ld;
    lbg::Acc<128> acc;
    lbg::load_acc<128>(acc, C, ld);
    lbg::mainloop<128, false, false, true>(acc, A, ld, B, ld, kd * LB_TILE, smem);
    lbg::store_acc<128>(acc, C, ld);
}

bool g_attr_done = false;
int set_attrs()
{
    if (g_attr_done) return LB_OK;
    LB_CUDA(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTF2_SMEM));
    LB_CUDA(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lbg::PIPE_BYTES));
    LB_CUDA(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lbg::PIPE_BYTES));
    g_attr_done = true;
    return LB_OK;
}

} // namespace

int lb_launch_potf2_block(lb_gp* h, int k, int do_factor)
{
    int rc = set_attrs();
    if (rc) return rc;
    potf2_inv_kernel<<<1, 256, POTF2_SMEM, h->stream>>>(h->dL, h->Np, k, h->dInvD, h->dInfo, do_factor);
    h->launches++;
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}

static inline int syrk_tiles(int T, int j0, int nc)
{
    int n = 0;
    for (int c = 0; c < nc; ++c) n += T - j0 - c;
    return n;
}

// Right-looking factorisation in pairs of 128-column panels with look-ahead:
//   side stream : panel(p)  = potf2(k) trsm(k) syrk[k -> col k+1] potf2(k+1) trsm(k+1)
//   main stream : a(p)      = K=256 update of the next pair's two block columns (k+2, k+3)
//                 b(p)      = K=256 update of everything right of them
// panel(p+1) only needs a(p), so the latency-bound panel work runs under b(p).
int lb_launch_potrf(lb_gp* h)
{
    int rc = set_attrs();
    if (rc) return rc;
    const int T = (int)(h->Np / LB_TILE);
    const int64_t ld = h->Np;
    cudaStream_t main = h->stream, side = h->side ? h->side : h->stream;
    LB_CUDA(cudaMemsetAsync(h->dInfo, 0, 2 * sizeof(int), main));
    if (side != main) {
        LB_CUDA(cudaEventRecord(h->ev[0], main)); // inputs (K) ready
        LB_CUDA(cudaStreamWaitEvent(side, h->ev[0], 0));
    }
    int p = 0;
    for (int k = 0; k < T; k += 2, ++p) {
        const bool pair = (k + 1 < T);
        // ---- panel(p) on the side stream ----
        {
            LbProfScope ps(h, side, LB_PC_POTF2);
            potf2_inv_kernel<<<1, 256, POTF2_SMEM, side>>>(h->dL, ld, k, h->dInvD, h->dInfo, 1);
        }
        h->launches++;
        if (pair) {
            {
                LbProfScope ps(h, side, LB_PC_TRSM_PANEL);
                trsm_panel_kernel<<<T - k - 1, lbg::THREADS, lbg::PIPE_BYTES, side>>>(h->dL, ld, k, h->dInvD);
            }
            {
                LbProfScope ps(h, side, LB_PC_SYRK_COL);
                syrk_kernel<<<T - k - 1, lbg::THREADS, lbg::PIPE_BYTES, side>>>(h->dL, ld, k, 1, k + 1, 1, T);
            }
            {
                LbProfScope ps(h, side, LB_PC_POTF2);
                potf2_inv_kernel<<<1, 256, POTF2_SMEM, side>>>(h->dL, ld, k + 1, h->dInvD, h->dInfo, 1);
            }
            h->launches += 3;
            if (k + 2 < T) {
                LbProfScope ps(h, side, LB_PC_TRSM_PANEL);
                trsm_panel_kernel<<<T - k - 2, lbg::THREADS, lbg::PIPE_BYTES, side>>>(h->dL, ld, k + 1, h->dInvD);
                h->launches++;
            }
        }
        if (k + 2 >= T) break;
        if (side != main) {
            LB_CUDA(cudaEventRecord(h->ev[1 + (p & 1)], side));
            LB_CUDA(cudaStreamWaitEvent(main, h->ev[1 + (p & 1)], 0));
        }
        // ---- a(p): the next pair's block columns ----
        const int j0 = k + 2;
        const int nca = (T - j0 < 2) ? (T - j0) : 2;
        {
            LbProfScope ps(h, main, LB_PC_SYRK);
            syrk_kernel<<<syrk_tiles(T, j0, nca), lbg::THREADS, lbg::PIPE_BYTES, main>>>(h->dL, ld, k, 2, j0, nca, T);
        }
        h->launches++;
        if (side != main) {
            LB_CUDA(cudaEventRecord(h->ev[3 + (p & 1)], main));
            LB_CUDA(cudaStreamWaitEvent(side, h->ev[3 + (p & 1)], 0));
        }
        // ---- b(p): the rest of the trailing matrix ----
        const int ncb = T - j0 - nca;
        if (ncb > 0) {
            LbProfScope ps(h, main, LB_PC_SYRK);
            syrk_kernel<<<syrk_tiles(T, j0 + nca, ncb), lbg::THREADS, lbg::PIPE_BYTES, main>>>(h->dL, ld, k, 2, j0 + nca, ncb, T);
            h->launches++;
        }
    }
    if (side != main) { // join
        LB_CUDA(cudaEventRecord(h->ev[5], side));
        LB_CUDA(cudaStreamWaitEvent(main, h->ev[5], 0));
    }
    LB_CUDA(cudaGetLastError());
    return LB_OK;
}
