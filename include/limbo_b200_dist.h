/* include/limbo_b200_dist.h — C ABI of the multi-GPU Cholesky building blocks (BASELINE.json config 5).
 *
 * The reference factors its kernel matrix in one address space (`Eigen::LLT<MatrixXd>(_kernel).matrixL()`,
 * src/limbo/model/gp.hpp:565); nothing in it corresponds to a distributed factor, so there is no reference interface to
 * mirror here: these are the device-side steps of a 1-D block-cyclic right-looking factorisation whose only exchange step
 * (the panel broadcast) is left to the host runtime (NCCL through torch.distributed in limbo_b200/dist_chol.py; any
 * MPI/NCCL binding can drive the same calls).
 *
 * Layout: panels of 256 columns ("pairs" of 128-blocks).  Pair p (global 128-block columns 2p, 2p+1) lives on rank
 * p mod G; a rank stores its pairs side by side, full height, column-major with leading dimension Nd (the order padded to
 * a multiple of 256, identity in the padding):
 *      local 128-block column l  <->  global 128-block column  2 * (G * (l / 2) + rank) + (l & 1)
 * All pointers are DEVICE pointers unless stated; every call is asynchronous on the handle's stream (lb_set_stream) and
 * returns 0 or a negative LB_ERR_* code (limbo_b200.h).
 */
#ifndef LIMBO_B200_DIST_H
#define LIMBO_B200_DIST_H

#include "limbo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Samples only: stages the N x D row-major HOST array X on the device without allocating the N x N factor storage of the
 * handle (the data part of GP::compute, gp.hpp:88-116).  Follow with lb_set_kernel. */
int lb_dchol_set_points(lb_gp* h, int64_t N, int D, const double* X_rowmajor_host);

/* This rank's columns of K = k(X, X) + (noise + 1e-8) I (gp.hpp:552-562, kernel/kernel.hpp:81-84), generated in place:
 * dLoc is Nd x ncols_local, ncols_local = 256 * (number of pairs of this rank). */
int lb_dchol_build(lb_gp* h, int64_t Nd, int rank, int G, int64_t ncols_local, double* dLoc);

/* Owner step for the pair whose first global 128-block column is kpair (even): dCols = the pair's 256 local columns
 * (Nd x 256, ld = Nd).  Factors them in place (diagonal blocks + panel below) and packs rows [(kpair+2)*128, Nd) into
 * dPanel (column-major, ld = Nd - (kpair+2)*128) for the broadcast.  dInvD: 2*128*128 doubles of scratch; dInfo: 2 ints,
 * dInfo[0] receives the 1-based index of the first non-positive pivot (LAPACK style) if there is one. */
int lb_dchol_panel(lb_gp* h, double* dCols, int64_t Nd, int kpair, double* dInvD, int* dInfo, double* dPanel);

/* Trailing update with the (broadcast) panel of pair kpair: C[i, j] -= P[i, :] P[j, :]^T for the local 128-block columns
 * l in [l0, l1) that lie right of the pair, all i >= j (K = 256 on the fp64 tensor cores). */
int lb_dchol_update(lb_gp* h, double* dLoc, int64_t Nd, const double* dPanel, int kpair, int l0, int l1, int rank, int G);

/* Zero the strictly upper part of the local columns (matrixL() has a zero upper triangle, gp.hpp:565) and write
 * sum_j log L_jj over this rank's columns with global index < N to *dLogdetPart (the log-det term of gp.hpp:272-274). */
int lb_dchol_finish(lb_gp* h, double* dLoc, int64_t Nd, int64_t N, int rank, int G, int64_t ncols_local, double* dLogdetPart);

#ifdef __cplusplus
}
#endif
#endif
