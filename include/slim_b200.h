/*
 * slim_b200.h -- extension entry points of libslim.so (slim-b200) that have no counterpart in the
 * reference header.  They expose the two halves of SLIM_Learn separately so that a caller can
 *   (1) keep the training matrix resident in HBM across many learn calls (regularisation sweeps,
 *       model selection: reference src/programs/slim_mselect.c:99-108 re-stages R every time), and
 *   (2) solve an arbitrary subset of target item columns on one GPU and leave the result on the
 *       device, which is what column-sharded multi-GPU learning needs: the reference's
 *       `#pragma omp for` over columns (src/libslim/estimate.c:402-403) becomes one process per
 *       GPU, each calling SLIMB200_LearnColumns on its shard, followed by one all-gather of W.
 * Plain C ABI: pointers and sizes only.
 */
#ifndef SLIM_B200_EXT_H
#define SLIM_B200_EXT_H

#include "slim.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slimb200_matrix slimb200_matrix_t; /* R staged in HBM (CSR + padded CSC + norms) */
typedef struct slimb200_result slimb200_result_t; /* solved columns, device resident          */

/* Number of usable CUDA devices (0 when there is none). */
int32_t SLIMB200_DeviceCount(void);
/* Message of the last failure on the calling thread. */
const char *SLIMB200_LastError(void);

/* Stage a CSR training matrix on `device`: the work of CreateTrainingMatrix
 * (reference src/libslim/setup.c:109-135).  rowval may be NULL (all-ones ratings).
 * Host pointers; for best H2D rates pass page-locked memory. */
slimb200_matrix_t *SLIMB200_Stage(int32_t device, int32_t nrows, const ssize_t *rowptr,
                                  const int32_t *rowind, const float *rowval, int32_t *r_status);
/* Same, from CSR arrays that already live on `device` (nnz = rowptr[nrows]). */
slimb200_matrix_t *SLIMB200_StageDevice(int32_t device, int32_t nrows, int64_t nnz,
                                        const int64_t *d_rowptr, const int32_t *d_rowind,
                                        const float *d_rowval, int32_t *r_status);
void SLIMB200_FreeMatrix(slimb200_matrix_t **matrix);
int32_t SLIMB200_MatrixInfo(const slimb200_matrix_t *matrix, int32_t *nrows, int32_t *ncols,
                            int64_t *nnz, int32_t *device, double *stage_ms, int32_t *stage_launches);
/* Copy the staged column view back to the host (colptr int64[ncols+1], unpadded). */
int32_t SLIMB200_MatrixCSC(const slimb200_matrix_t *matrix, int64_t *colptr, int32_t *colind,
                           float *colval, float *cnorms);

/* The engine relabels items by popularity internally (internal id = position when items are sorted by
 * descending nnz, ties by ascending id); rank[original id] = internal id.  Results always use original ids. */
int32_t SLIMB200_MatrixItemOrder(const slimb200_matrix_t *matrix, int32_t *rank);
/* Copy the staged window Gram blocks back (tests): double[ceil(ncols/32)][32][32], block w holds
 * <a_k, a_m> for the items with INTERNAL ids 32w+k and 32w+m (zero diagonal). */
int32_t SLIMB200_MatrixWindowGram(const slimb200_matrix_t *matrix, double *out);

/* Gram matrix G = R^T R (internal item order) that staging builds for the Gram-space solver when it fits in
 * HBM: every <a_i, a_k> the coordinate-descent sweeps of cd.c:117-133 need.  elem_bytes = 0 (not staged:
 * too large, or SLIMB200_GRAM=0), 4 (exact integer sums: non-negative integer ratings and sums below 2^24; stored
 * PACKED, 32 / 16 / 8 bits per element depending on the column's bound, see SLIMB200_MatrixGramLayout) or 8 (double);
 * build_ms = CUDA-event time of the build.  SLIMB200_MatrixGram copies it back densely, ncols x ncols elements
 * without row padding -- float for elem_bytes 4, double for 8 (tests). */
int32_t SLIMB200_MatrixGramInfo(const slimb200_matrix_t *matrix, int32_t *elem_bytes, double *build_ms);
int32_t SLIMB200_MatrixGram(const slimb200_matrix_t *matrix, void *out);
/* Footprint in HBM and column ranges of the packed layout: columns (internal ids) [0, h32) hold 32-bit elements,
 * [h32, h16) 16-bit, the rest 8-bit; a column's width follows from the bound G[k][i] <= rmax * sum_u r_ui^2.
 * h32 = h16 = 0 for the fp64 layout.  Any pointer may be NULL. */
int32_t SLIMB200_MatrixGramLayout(const slimb200_matrix_t *matrix, int64_t *bytes, int32_t *h32, int32_t *h16);
/* STAIR layout: when the full packed matrix does not fit in HBM (BASELINE configs[4]: 500 K items = 308 GB) staging
 * keeps, for every panel p of 64 columns, only the rows [0, max(64 (p + 1), hd)) -- the upper triangle by panels plus
 * a full square of the hd most popular items (133 GB for configs[4]); G is symmetric, so an element that is not
 * stored is read as its mirror image.  stair = 1 when this layout is in use, hd = side of the square.
 * SLIMB200_GRAM_LAYOUT=stair|full forces / forbids it, SLIMB200_GRAM_HD sets hd (default 32768). */
int32_t SLIMB200_MatrixGramStair(const slimb200_matrix_t *matrix, int32_t *stair, int32_t *hd);

/* Solve target columns cols[0..ncols_sel) (cols == NULL: every column): the per-column body of
 * EstimateModelCD (reference src/libslim/estimate.c:405-505) + CoordinateDescent (cd.c:101-142).
 * Options as for SLIM_Learn.  imodel: optional warm-start model handle.  Calls on DIFFERENT staged matrices run
 * concurrently; concurrent calls on the same matrix are serialised inside the library (its solve scratch is per matrix). */
slimb200_result_t *SLIMB200_LearnColumns(slimb200_matrix_t *matrix, const int32_t *ioptions,
                                         const double *doptions, const int32_t *cols,
                                         int32_t ncols_sel, const slim_t *imodel, int32_t *r_status);
void SLIMB200_FreeResult(slimb200_result_t **result);
/* nsel, total nnz, CUDA-event time of the solve kernel(s) and of the gather kernel (ms),
 * kernels launched. Any pointer may be NULL. */
int32_t SLIMB200_ResultInfo(const slimb200_result_t *result, int32_t *nsel, int64_t *nnz,
                            double *solve_ms, double *gather_ms, int32_t *launches);
/* Per-column counters in the order of `cols` (any pointer may be NULL): sweeps executed
 * (cd.c:140), |A_j|, sum of nnz over active columns, sum of row lengths over the users of column j,
 * 1/2||y - yhat||^2 and the objective (estimate.c:477-489). */
int32_t SLIMB200_ResultStats(const slimb200_result_t *result, int32_t *niters, int32_t *nactive,
                             int64_t *active_nnz, int64_t *expand_nnz, double *rnorm, double *objval);
/* Profiling counters of the cluster kernel per column: phase_us float[nsel][4] = microseconds spent in
 * candidate expansion, active-set build, sweeps, epilogue; rounds int32[nsel] = barrier rounds per sweep.
 * Zeros when the single-CTA kernel was used. */
int32_t SLIMB200_ResultPhases(const slimb200_result_t *result, float *phase_us, int32_t *rounds);
/* Solved columns as CSC: colptr int64[nsel+1], colind int32[nnz] (ascending), colval float[nnz]. */
int32_t SLIMB200_ResultToHost(const slimb200_result_t *result, int64_t *colptr, int32_t *colind,
                              float *colval);
/* Same into DEVICE buffers on the matrix's device: counts int32[nsel], colind, colval. */
int32_t SLIMB200_ResultToDevice(const slimb200_result_t *result, int32_t *d_counts,
                                int32_t *d_colind, float *d_colval);

/* ---- multi-GPU --------------------------------------------------------------------------------
 * The target columns are independent problems (reference src/libslim/estimate.c:402-403 is a plain parallel-for),
 * so learning shards by column: R replicated on every GPU, every rank solves its own columns with
 * SLIMB200_LearnColumns and ONE NCCL all-gather at the end assembles W on every rank -- the counterpart of
 * SaveModel's concatenation of the per-column lists (estimate.c:570-588).
 *
 * Inside one process SLIM_Learn / Py_SLIM_Learn do all of this themselves: they use the devices named by the
 * environment (SLIMB200_DEVICES="0,1,..", or SLIMB200_GPUS=n; default: every visible device for matrices with
 * >= 20 M nonzeros, else one), one host thread per device, communicators created once per process.
 * With one process per GPU (torchrun, MPI) the caller exchanges the 128-byte id of rank 0 by its own means and uses
 * the three calls below.  NCCL is bound at run time (libnccl.so.2); without it these calls fail with SLIM_ERROR. */
typedef struct slimb200_comm slimb200_comm_t;
/* ncclGetUniqueId: fills 128 bytes on the calling rank (rank 0), to be sent to every other rank. */
int32_t SLIMB200_CommUniqueId(void *id128);
/* ncclCommInitRank on `device`; collective over the nranks callers. */
slimb200_comm_t *SLIMB200_CommInitRank(int32_t device, int32_t nranks, int32_t rank, const void *id128,
                                       int32_t *r_status);
void SLIMB200_CommFree(slimb200_comm_t **comm);
/* Collective.  `local` holds the columns this rank solved; positions[k] (host, int32[nsel of local]) is the index of
 * local column k in the global list of ncols_total columns; every index must be owned by exactly one rank.  Returns
 * the CSC of ALL ncols_total columns, resident on this rank's device and identical on every rank (read it with
 * SLIMB200_ResultToHost / ResultToDevice / ResultToModel).  One grouped NCCL collective (variable-length all-gather)
 * plus two placement kernels; nothing is padded and no shard passes through the host. */
slimb200_result_t *SLIMB200_AllGatherColumns(slimb200_comm_t *comm, const slimb200_result_t *local,
                                             const int32_t *positions, int32_t ncols_total, int32_t *r_status);
/* Model handle (both views of SaveModel, estimate.c:570-593) from a result that holds ALL item columns 0..n-1 in
 * order (SLIMB200_LearnColumns with cols == NULL, or SLIMB200_AllGatherColumns with positions = column ids); the
 * CSR index (gk_csr_CreateIndex, lib/GKlib/csr.c:1546-1584) is built on the GPU. */
slim_t *SLIMB200_ResultToModel(const slimb200_result_t *result, int32_t *r_status);

/* Build a model handle (both views, reference SaveModel src/libslim/estimate.c:570-593) from the
 * CSC of all nitems columns, e.g. after gathering the shards of every GPU. */
slim_t *SLIMB200_AssembleModel(int32_t nitems, const int64_t *colptr, const int32_t *colind,
                               const float *colval, int32_t *r_status);

#ifdef __cplusplus
}
#endif
#endif /* SLIM_B200_EXT_H */
