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

/* ---- distributed FIT of one GP (limbo_b200/dist_fit.py): the factorisation above, with every rank assembling the complete
 * factor in a regular handle from the panel messages, so that lb_query / lb_acq_argmax can then shard the candidates over the
 * ranks with no further exchange (GP::compute, gp.hpp:88-116, spread over the GPUs; the result is bit-identical to lb_fit).
 * Message of pair kpair:  [ 256 x 256 diagonal block (column-major, ld 256) | inv(L_kk), inv(L_k+1,k+1) (2 x 128 x 128) |
 * rows below the pair (ld = Nd - (kpair + 2) * 128) ]:  LB_DCHOL_HEAD doubles, then what lb_dchol_panel packs. */
#define LB_DCHOL_HEAD (256 * 256 + 2 * 128 * 128)
/* owner: head of the message (call after lb_dchol_panel, same handle / stream; dInvD = lb_dchol_panel's scratch) */
int lb_dchol_pack_head(lb_gp* h, const double* dCols, int64_t Nd, int kpair, const double* dInvD, double* dMsg);
/* target handle (lb_set_data + lb_set_kernel done; its padded order must equal Nd, else LB_ERR_UNSUPPORTED -> use lb_fit) */
int lb_dchol_adopt_begin(lb_gp* h, int64_t Nd);
/* copy one panel message into the handle's factor and diagonal-block inverses, on cuda_stream (NULL: the handle's stream) */
int lb_dchol_unpack(lb_gp* h, const double* dMsg, int64_t Nd, int kpair, void* cuda_stream);
/* all messages unpacked: the handle is fitted, alpha = K^-1 obs_mean is solved locally (gp.hpp:605-611); info = the
 * factorisation's LAPACK-style info (> 0 is returned as is) */
int lb_dchol_adopt_end(lb_gp* h, int info);

/* ---- inversion of the factor spread over G GPUs, for the reduced-precision candidate path (LB_PREC_TF32 / FP16 / FP16X3,
 * BASELINE.json config 4; limbo_b200/dist_inv.py).  sigma^2(v) = k(v,v) - |L^-1 k(v)|^2 (gp.hpp:618-624) is scored against a
 * reduced-precision copy of ALL of L^-1 on every rank (the candidates are sharded, lb_acq_argmax), but the inverse itself is
 * independent per column: rank r computes the 128-column tiles c = r, r + G, r + 2G, ... (N^3 / (3 G) flops, blocked forward
 * solve of identity columns on the fp64 tensor cores), casts them, and one all_gather of the chunks gives every rank the whole
 * copy.  Chunk of a rank = [ hi plane: Np x W row-major, W = 128 * ceil(Np / 128 / G), fp16 or fp32(tf32) | lo plane (FP16X3) |
 * W doubles |L^-1 e_k|^2 (weights of the rounding-bias term) ]; local column tile t is global tile r + t * G.
 * The handle must be fitted (lb_fit or lb_dchol_adopt_end); no other call on it between lb_dinv_columns and lb_dinv_pack. */
/* bytes of one chunk (> 0), or a negative LB_ERR_* (LB_ERR_UNSUPPORTED for an fp64 handle) */
long long lb_dinv_chunk_bytes(const lb_gp* h, int G);
/* compute this rank's columns of L^-1 (kept in the handle's query workspace); *absmax_host = max |.| over them (the fp16 scale
 * needs the maximum over ALL ranks: reduce it before lb_dinv_pack).  Synchronises the handle's stream. */
int lb_dinv_columns(lb_gp* h, int rank, int G, double* absmax_host);
/* cast this rank's columns into dChunk (device, lb_dinv_chunk_bytes) with the scale derived from absmax_all */
int lb_dinv_pack(lb_gp* h, int rank, int G, double absmax_all, void* dChunk);
/* dAll = the G chunks in rank order (device): assemble the handle's reduced-precision L^-1; after this lb_query / lb_acq_argmax
 * score without inverting anything */
int lb_dinv_adopt(lb_gp* h, int G, const void* dAll, double absmax_all);

#ifdef __cplusplus
}
#endif
#endif
