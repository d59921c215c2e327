// engine.h -- internal interface between the C-ABI shims (api.cpp) and the CUDA engine
// (engine.cu).  Not installed; the public headers are include/slim.h and include/slim_b200.h.
#pragma once
#include <stdint.h>
#include <sys/types.h>

namespace slimb200 {

// Return codes (include/slim.h slim_rstatus_et; reference include/slim.h:177-182)
enum { kOk = 1, kErrInput = -2, kErrMemory = -3, kErr = -4 };

constexpr double kEps = 1e-7;  // EPSILON, reference src/libslim/def.h:14

struct LearnParams {
  double l1r, l2r, opttol;
  int32_t maxniters;
  int32_t dbglvl;
  // fSLIM (reference src/libslim/neighbors.c, estimate.c:424-431): nnbrs > 0 restricts the active set of every
  // target to its nnbrs most similar columns; simtype as include/slim.h (0 cos, 1 jac, 2 dotp)
  int32_t nnbrs = 0;
  int32_t simtype = 0;
};

// Host view of a model used for warm starts: CSC of a previous W (reference
// src/libslim/estimate.c:453-464 reads imat->colptr/colind/colval).
struct WarmStart {
  int32_t ncols;
  const ssize_t *colptr;
  const int32_t *colind;
  const float *colval;
};

struct Matrix;  // R staged in HBM: CSR + 16-byte-padded CSC + norms
struct Result;  // solved columns, device resident

// Kernel timing / work counters of one learn call (CUDA events on the engine stream).
struct Timings {
  double solve_ms;     // the CD kernel(s): candidate expansion + sweeps + compaction
  double gather_ms;    // ordered gather of the solved columns
  int32_t launches;    // kernels launched by the call
  int32_t solve_launches;
};

Matrix *stage(int device, int32_t nrows, const ssize_t *rowptr, const int32_t *rowind,
              const float *rowval, bool inputs_on_device, int64_t nnz_if_device, int32_t *status);
void free_matrix(Matrix *m);
void matrix_info(const Matrix *m, int32_t *nrows, int32_t *ncols, int64_t *nnz, int32_t *device,
                 double *stage_ms, int32_t *stage_launches);
// Copies the staged column view back (tests): colptr is in UNPADDED units.
int matrix_csc_to_host(const Matrix *m, int64_t *colptr, int32_t *colind, float *colval,
                       float *cnorms);

// Internal item order (tests): rank[original id] = internal id (items sorted by descending nnz).
int matrix_item_order_to_host(const Matrix *m, int32_t *rank);
// Window Gram blocks in INTERNAL item order (tests): double[ceil(ncols/32)][32][32].
int matrix_window_gram_to_host(const Matrix *m, double *out);

// Gram matrix G = R^T R staged for the Gram-space solver (INTERNAL item order).  elem_bytes: 0 when G
// was not staged (too large / disabled), 4 = float (exact integer sums), 8 = double.
void matrix_gram_info(const Matrix *m, int32_t *elem_bytes, double *build_ms);
// Dense ncols x ncols copy (no row padding): float when elem_bytes == 4 (the packed integer layout, exact), double
// when elem_bytes == 8 (tests).
int matrix_gram_to_host(const Matrix *m, void *out);
// Footprint of the staged Gram matrix in HBM and the column ranges of the packed layout: 32-bit elements for
// columns [0, h32), 16-bit for [h32, h16), 8-bit from h16 on (zeros for the fp64 layout).
void matrix_gram_layout(const Matrix *m, int64_t *bytes, int32_t *h32, int32_t *h16);
// stair = 1 when the packed matrix is held in the STAIR layout (panel p stores rows [0, max(64(p+1), hd)) only: the
// layout for item counts whose full matrix does not fit), hd = side of its full head square.
void matrix_gram_stair(const Matrix *m, int32_t *stair, int32_t *hd);

Result *learn(Matrix *m, const LearnParams &p, const int32_t *cols, int32_t nsel,
              const WarmStart *warm, int32_t *status);
void free_result(Result *r);
void result_info(const Result *r, int32_t *nsel, int64_t *nnz, Timings *t);
// Per-column statistics in the caller's column order; any pointer may be null.
int result_stats(const Result *r, int32_t *niters, int32_t *nactive, int64_t *active_nnz,
                 int64_t *expand_nnz, double *rnorm, double *objval);
// Per-column phase times of the cluster kernel in microseconds [nsel][4] (candidates, active set,
// sweeps, epilogue) and the number of barrier rounds per sweep; zeros for the team kernel.
int result_phases(const Result *r, float *phase_us, int32_t *ngroups);
int result_to_host(const Result *r, int64_t *colptr, int32_t *colind, float *colval);
// Same, into caller-owned DEVICE buffers on the matrix's device (counts int32[nsel]).
int result_to_device(const Result *r, int32_t *d_counts, int32_t *d_colind, float *d_colval);

// Batched top-N for every row of a user-history matrix on the GPU: the loop of Py_SLIM_Predict (reference
// src/libslim/pyapi.c:530-563) around GetRecommendations (src/libslim/predict.c:15-71).  Host arrays in and
// out; out_ids / out_scores are [nusers][nrcmds] and only the first out_counts[u] entries of a row are
// written (the arrays are read first, so the rest keeps the caller's content).  Lists and scores are
// bit-identical to the host restatement in api.cpp.
int predict_topn(int device, int32_t wrows, int32_t wcols, const ssize_t *wrowptr, const int32_t *wrowind,
                 const float *wrowval, int32_t nusers, const ssize_t *urowptr, const int32_t *urowind,
                 const float *urowval, int32_t nrcmds, int32_t *out_ids, float *out_scores, int32_t *out_counts,
                 double *kernel_ms);

// Column nnz by ORIGINAL item id (host copy kept by stage()).
int matrix_colcounts(const Matrix *m, int32_t *cnt);

// ---- multi-GPU (gather.cuh): NCCL communicators and the final all-gather of W ------------------
struct Comm;
int comm_unique_id(void *id128);  // 128 bytes (ncclUniqueId), to be broadcast to the other ranks by the caller
Comm *comm_init(int device, int nranks, int rank, const void *id128, int32_t *status);
int comm_init_all(int ndev, const int *devices, Comm **out);  // one communicator per device, all in this process
void comm_free(Comm *c);
void comm_info(const Comm *c, int32_t *nranks, int32_t *rank, int32_t *device);
// All-gather of the column shards: positions[k] = index of local column k in the global list of ncols_total columns.
// Returns the CSC of all ncols_total columns on the communicator's device (same content on every rank).
Result *allgather_columns(Comm *c, const Result *local, const int32_t *positions, int32_t ncols_total, int32_t *status);
// Both views of a model (SaveModel, estimate.c:570-593) from a Result holding ALL columns 0..n-1 in order; the CSR
// index is built on the GPU.  Arrays are caller-allocated ([n+1] / [nnz]); index_ms = CUDA-event time of the index build.
int model_to_host(const Result *r, ssize_t *colptr, int32_t *colind, float *colval, ssize_t *rowptr, int32_t *rowind,
                  float *rowval, double *index_ms);

int device_count();
const char *last_error();

}  // namespace slimb200
