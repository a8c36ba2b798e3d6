// limbo_b200/csrc/gemm.cuh — fp64 tensor-core (DMMA) tile GEMM building block.
//
// One CTA (512 threads = 16 warps, 4 along M x 4 along N; a single warp can only
// drive the DMMA pipe of its SM sub-partition at half rate — measured,
// profiles/r01_microbench.json — so every sub-partition gets 4 warps) accumulates a
// 128 x BN tile  acc += A(128 x K) * B(K x BN)  with a 3-stage cp.async
// pipeline (BK = 32: one CTA barrier per 32 k).  Both operands can be "outer-contiguous" (the m / n index
// is the unit-stride one, i.e. a column-major 128 x K block) or "K-contiguous"
// (the k index is unit-stride).  Shared-memory tiles are padded (+4 doubles)
// so every 64-bit fragment load is bank-conflict free.
//
// This replaces the arithmetic the reference delegates to Eigen:
//   LLT trailing update / panel solve            model/gp.hpp:565
//   triangular solves for alpha, sigma^2, K^-1   model/gp.hpp:260-261,608-610,620
#pragma once
#include "common.cuh"

namespace lbg {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int STAGES = 3;
constexpr int THREADS = 512;
constexpr int PITCH_OC = BM + 4; // outer-contiguous tile: [BK][128+4]
constexpr int PITCH_KC = BK + 4; // k-contiguous tile:     [128][16+4]
constexpr int STAGE_DOUBLES = BM * PITCH_KC; // 2560 >= BK*PITCH_OC (2112)
constexpr size_t PIPE_BYTES = (size_t)2 * STAGES * STAGE_DOUBLES * sizeof(double); // A + B stages

// Load one BK-slab of an operand tile (128 "outer" x 16 k) into shared memory.
template <bool KC>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ g, int64_t ld, int nouter)
{
    const int tid = threadIdx.x;
    if (KC) {
        // element (o, k) at g[k + o*ld]; smem [o][k], BK/2 chunks of 16 B per row
        constexpr int CPR = BK / 2;
        for (int c = tid; c < nouter * CPR; c += THREADS) {
            int o = c / CPR, kc = c - o * CPR;
            lb_cp_async16(s + o * PITCH_KC + 2 * kc, g + (int64_t)o * ld + 2 * kc);
        }
    }
    else {
        // element (o, k) at g[o + k*ld]; smem [k][o]
        const int cpr = nouter >> 1; // 16 B chunks per k-row
        for (int c = tid; c < BK * cpr; c += THREADS) {
            int k = c / cpr, oc = c - k * cpr;
            lb_cp_async16(s + k * PITCH_OC + 2 * oc, g + (int64_t)k * ld + 2 * oc);
        }
    }
}

// Accumulators of one warp: MT m16-tiles x NT n8-tiles.
// Warp grid is 4 (M) x 4 (N): warp tile = 32 x (BN/4) -> MT = 2, NT = BN/32.
template <int BN>
struct Acc {
    static constexpr int NT = BN / 32;
    double v[2][NT][4];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int c = 0; c < 4; ++c) v[a][b][c] = 0.0;
    }
};

template <int BN, bool A_KC, bool B_KC, bool NEG_A = false>
__device__ __forceinline__ void compute_stage(Acc<BN>& acc, const double* sA, const double* sB)
{
    constexpr int NT = BN / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int m_base = wm * 32, n_base = wn * (BN / 4);
#pragma unroll
    for (int k0 = 0; k0 < BK; k0 += 8) {
        double a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int m = m_base + mt * 16 + g + 8 * (i & 1);
                int k = k0 + t + 4 * (i >> 1);
                a[mt][i] = A_KC ? sA[m * PITCH_KC + k] : sA[k * PITCH_OC + m];
                if (NEG_A) a[mt][i] = -a[mt][i];
            }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            double b[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int n = n_base + nt * 8 + g;
                int k = k0 + t + 4 * i;
                b[i] = B_KC ? sB[n * PITCH_KC + k] : sB[k * PITCH_OC + n];
            }
            lb_dmma_16x8x8(acc.v[0][nt], a[0], b);
            lb_dmma_16x8x8(acc.v[1][nt], a[1], b);
        }
    }
}

// acc += A * B (acc -= A * B with NEG_A, so a tile update C - A*B can preload C
// into the accumulators and overlap those loads with the pipeline prologue
// instead of paying a dependent load-subtract-store epilogue) over K (multiple of 16).  gA/gB point at the (0,0) element of
// the operand tile for k = 0; stepping k by 16 advances an outer-contiguous
// operand by 16*ld and a K-contiguous one by 16.  All threads must call.
// smem: PIPE_BYTES.  On return all cp.async groups are drained and the CTA is
// synchronised (smem may be reused).
template <int BN, bool A_KC, bool B_KC, bool NEG_A = false>
__device__ __forceinline__ void mainloop(Acc<BN>& acc, const double* __restrict__ gA, int64_t lda,
    const double* __restrict__ gB, int64_t ldb, int K, double* smem)
{
    double* sA = smem;
    double* sB = smem + STAGES * STAGE_DOUBLES;
    const int nk = K / BK;
    const int64_t stepA = A_KC ? BK : (int64_t)BK * lda;
    const int64_t stepB = B_KC ? BK : (int64_t)BK * ldb;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_tile<A_KC>(sA + s * STAGE_DOUBLES, gA + s * stepA, lda, BM);
            load_tile<B_KC>(sB + s * STAGE_DOUBLES, gB + s * stepB, ldb, BN);
        }
        lb_cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        lb_cp_async_wait<STAGES - 2>();
        __syncthreads();
        int nx = kt + STAGES - 1;
        if (nx < nk) {
            int s = nx % STAGES;
            load_tile<A_KC>(sA + s * STAGE_DOUBLES, gA + nx * stepA, lda, BM);
            load_tile<B_KC>(sB + s * STAGE_DOUBLES, gB + nx * stepB, ldb, BN);
        }
        lb_cp_async_commit();
        int s = kt % STAGES;
        compute_stage<BN, A_KC, B_KC, NEG_A>(acc, sA + s * STAGE_DOUBLES, sB + s * STAGE_DOUBLES);
    }
    lb_cp_async_wait<0>();
    __syncthreads();
}

// Apply f(row, col, value) to every accumulator element of this thread
// (row in [0,128), col in [0,BN)).
template <int BN, typename F>
__device__ __forceinline__ void for_each_acc(Acc<BN>& acc, F&& f)
{
    constexpr int NT = BN / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                int row = wm * 32 + mt * 16 + g + 8 * (i >> 1);
                int col = wn * (BN / 4) + nt * 8 + 2 * t + (i & 1);
                f(row, col, acc.v[mt][nt][i]);
            }
}

// acc <- C tile (column-major, ld) : plain loads, no arithmetic, so they stay in
// flight while the cp.async prologue is issued.
template <int BN>
__device__ __forceinline__ void load_acc(Acc<BN>& acc, const double* __restrict__ C, int64_t ld)
{
    for_each_acc<BN>(acc, [&](int r, int c, double& v) { v = __ldcs(C + r + (int64_t)c * ld); });
}
template <int BN>
__device__ __forceinline__ void store_acc(Acc<BN>& acc, double* __restrict__ C, int64_t ld)
{
    for_each_acc<BN>(acc, [&](int r, int c, double& v) { C[r + (int64_t)c * ld] = v; });
}

// Second-phase product with a resident B operand: acc2 += A(128 x 128) * Bres
// where Bres is in shared memory as [n][k] with pitch BM+4 (k-contiguous) and A
// (outer-contiguous, ld = lda) streams through the A pipeline stages.
// smem_pipe: the A stage area (STAGES*STAGE_DOUBLES doubles).
template <int BN>
__device__ __forceinline__ void mainloop_resB(Acc<BN>& acc, const double* __restrict__ gA, int64_t lda,
    const double* sBres, double* smem_pipe)
{
    constexpr int NT = BN / 32;
    constexpr int PB = BM + 4;
    const int nk = BM / BK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int m_base = wm * 32, n_base = wn * (BN / 4);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        load_tile<false>(smem_pipe + s * STAGE_DOUBLES, gA + (int64_t)s * BK * lda, lda, BM);
        lb_cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        lb_cp_async_wait<STAGES - 2>();
        __syncthreads();
        int nx = kt + STAGES - 1;
        if (nx < nk) load_tile<false>(smem_pipe + (nx % STAGES) * STAGE_DOUBLES, gA + (int64_t)nx * BK * lda, lda, BM);
        lb_cp_async_commit();
        const double* sA = smem_pipe + (kt % STAGES) * STAGE_DOUBLES;
#pragma unroll
        for (int k0 = 0; k0 < BK; k0 += 8) {
            double a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int m = m_base + mt * 16 + g + 8 * (i & 1);
                    int k = k0 + t + 4 * (i >> 1);
                    a[mt][i] = sA[k * PITCH_OC + m];
                }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                double b[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    int n = n_base + nt * 8 + g;
                    int k = kt * BK + k0 + t + 4 * i;
                    b[i] = sBres[n * PB + k];
                }
                lb_dmma_16x8x8(acc.v[0][nt], a[0], b);
                lb_dmma_16x8x8(acc.v[1][nt], a[1], b);
            }
        }
    }
    lb_cp_async_wait<0>();
    __syncthreads();
}

} // namespace lbg
