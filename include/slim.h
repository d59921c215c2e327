/*
 * slim.h -- public C ABI of libslim.so (slim-b200).
 *
 * Drop-in for the header of KarypisLab/SLIM: every declaration below keeps the name, argument
 * order, argument meaning and return convention of the reference interface it replaces (cited
 * per entry as reference file:line), so existing callers -- the reference CLI programs, the
 * reference python-package (ctypes) -- link and run unchanged.  What changes is what is behind
 * SLIM_Learn / Py_SLIM_Learn: the OpenMP coordinate-descent learner is replaced by a CUDA engine
 * for NVIDIA B200 (sm_100a).  There is no CPU fallback for the learner: when no CUDA device is
 * usable SLIM_Learn returns NULL and sets *r_status to SLIM_ERROR.
 *
 * Device selection / tuning is by environment variable so the ABI stays frozen:
 *   SLIMB200_DEVICE   CUDA device ordinal used by SLIM_Learn / Py_SLIM_Learn (default 0)
 *   SLIMB200_NT       threads cooperating on one target item column (32, 128 or 512)
 * The nthreads option is accepted and ignored by the learner.
 */
#ifndef SLIM_B200_SLIM_H
#define SLIM_B200_SLIM_H

#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque model / matrix handle.  As in the reference (include/slim.h:47, src/libslim/api.c:95) the
 * object behind it has the memory layout of GKlib's gk_csr_t (lib/GKlib/gk_struct.h:75-88), every
 * array malloc()-allocated, because reference callers dereference it directly
 * (src/programs/slim_learn.c:83). */
typedef void slim_t;

#define SLIM_VERSION "2.0"
#define SLIM_NOPTIONS 40 /* length of the ioptions[] / doptions[] arrays (include/slim.h:56) */

/* Return codes (reference include/slim.h:177-182). */
typedef enum {
  SLIM_OK = 1,
  SLIM_ERROR_INPUT = -2,
  SLIM_ERROR_MEMORY = -3,
  SLIM_ERROR = -4
} slim_rstatus_et;

/* Model types, similarity types, algorithms (reference include/slim.h:185-210). */
typedef enum { SLIM_MTYPE_SLIM = 0, SLIM_MTYPE_FSLIM = 1, SLIM_MTYPE_OSLIM = 2, SLIM_MTYPE_OFSLIM = 3 } slim_mtype_et;
typedef enum { SLIM_SIMTYPE_COS = 0, SLIM_SIMTYPE_JAC = 1, SLIM_SIMTYPE_DOTP = 2 } slim_simtype_et;
typedef enum { SLIM_ALGO_ADMM = 0, SLIM_ALGO_CD = 1 } slim_algo_et;

/* Text labels indexed by the enums above; the reference CLIs print them
 * (src/programs/slim_learn.c:39-43), so the header keeps them (reference include/slim.h:193,203,212). */
static const char slim_mtypenames[][10] = {"SLIM", "FSLIM", "OSLIM", "OFSLIM", ""};
static const char slim_simtypenames[][10] = {"cos", "jac", "dotp", ""};
static const char slim_algonames[][10] = {"admm", "cd", ""};

/* Slots of the option arrays (reference include/slim.h:215-230).  A NULL array or a slot holding
 * -1 selects the default (src/libslim/macros.h:14-15; defaults src/libslim/api.c:42-52). */
typedef enum {
  SLIM_OPTION_DBGLVL = 0,
  SLIM_OPTION_NNBRS = 1,
  SLIM_OPTION_SIMTYPE = 2,
  SLIM_OPTION_NTHREADS = 3,
  SLIM_OPTION_MAXNITERS = 4,
  SLIM_OPTION_ALGO = 5,
  SLIM_OPTION_ORDERED = 6,
  SLIM_OPTION_L1R = 7,
  SLIM_OPTION_L2R = 8,
  SLIM_OPTION_OPTTOL = 9,
  SLIM_OPTION_NRCMDS = 10
} slim_options_et;

/* Debug bits (reference include/slim.h:233-239). */
typedef enum {
  SLIM_DBG_INFO = 1,
  SLIM_DBG_TIME = 2,
  SLIM_DBG_PROGRESS = 4,
  SLIM_DBG_PROGRESS2 = 16,
  SLIM_DBG_MEMORY = 2048
} slim_dbglvl_et;

/* ---- reference include/slim.h:79-167 ---------------------------------------------------------- */

/* replaces SLIM_iSetDefaults, src/libslim/api.c:149-153: fills the 40 slots with -1, returns 1 */
int32_t SLIM_iSetDefaults(int32_t *options);
/* replaces SLIM_dSetDefaults, src/libslim/api.c:161-165 */
int32_t SLIM_dSetDefaults(double *options);

/* replaces SLIM_Learn, include/slim.h:107-110 / src/libslim/api.c:33-96.
 * Inputs are borrowed (copied to the GPU); the returned model is owned by the caller and released
 * with SLIM_FreeModel.  imodel (optional) warm-starts the solver (src/libslim/estimate.c:453-464).
 * Only SLIM_ALGO_CD with nnbrs == 0 is implemented; anything else returns NULL with
 * *r_status = SLIM_ERROR_INPUT (the reference exit()s or runs ADMM/fSLIM). */
slim_t *SLIM_Learn(int32_t nrows, ssize_t *rowptr, int32_t *rowind, float *rowval,
                   int32_t *ioptions, double *doptions, slim_t *imodel, int32_t *r_status);

/* replaces SLIM_GetTopN, include/slim.h:125-127 / src/libslim/api.c:111-141 */
int32_t SLIM_GetTopN(slim_t *model, int32_t nratings, int32_t *itemids, float *ratings,
                     int32_t *ioptions, int32_t nrcmds, int32_t *rids, float *rscores);

/* replace SLIM_WriteModel / SLIM_ReadModel / SLIM_FreeModel, src/libslim/api.c:174-204
 * (binary row format of lib/GKlib/csr.c:825-838 / :436-460) */
int32_t SLIM_WriteModel(slim_t *model, char *filename);
slim_t *SLIM_ReadModel(char *filename);
void SLIM_FreeModel(slim_t **model);

/* replaces SLIM_DetermineHeadAndTail, src/libslim/api.c:215-245 (malloc'd int32[ncols]: 0 head, 1 tail) */
int32_t *SLIM_DetermineHeadAndTail(int32_t nrows, int32_t ncols, ssize_t *rowptr, int32_t *rowind);

/* ---- ctypes entry points of the reference python-package (src/libslim/pyapi.c; the reference
 *      does not declare them in a header, python-package/SLIM/core.py:366-385,692-804 binds them) -- */

int32_t Py_csr_wrapper(int32_t nrows, ssize_t *rowptr, int32_t *rowind, float *rowval,
                       slim_t **matrix_out);                        /* pyapi.c:22-38  */
int32_t Py_csr_save(slim_t *mathandle, char *fname);                /* pyapi.c:47-51  */
int32_t Py_csr_load(slim_t **mathandle, char *fname);               /* pyapi.c:60-64  */
int32_t Py_csr_free(slim_t *mathandle);                             /* pyapi.c:72-76  */
int32_t Py_csr_stat(slim_t *mathandle, int32_t *nnz);               /* pyapi.c:85-89  */
int32_t Py_csr_export(slim_t *mathandle, int32_t *indptr, int32_t *indices,
                      float *data);                                 /* pyapi.c:101-123 */
/* replaces Py_SLIM_Learn, pyapi.c:134-199 (same engine as SLIM_Learn, no warm start) */
int32_t Py_SLIM_Learn(slim_t *trnhandle, int32_t *ioptions, double *doptions, slim_t **model_out);
/* replaces Py_SLIM_Mselect, pyapi.c:214-412 */
int32_t Py_SLIM_Mselect(slim_t *trnhandle, slim_t *tsthandle, int32_t *ioptions, double *doptions,
                        double *arrayl1, double *arrayl2, int32_t nl1, int32_t nl2,
                        double *bestl1HR, double *bestl2HR, double *bestHRHR, double *bestARHR,
                        double *bestl1AR, double *bestl2AR, double *bestHRAR, double *bestARAR);
int32_t Py_SLIM_GetTopN(slim_t *model, int32_t nratings, int32_t *itemids, float *ratings,
                        int32_t nrcmds, int32_t *rids, float *rscores,
                        int32_t dbglvl);                            /* pyapi.c:414-440 */
int32_t Py_SLIM_GetTopN_1vsk(slim_t *model, int32_t nratings, int32_t *itemids, float *ratings,
                             int32_t nrcmds, int32_t *rids, float *rscores, int32_t nnegs,
                             int32_t *negitems, int32_t dbglvl);    /* pyapi.c:442-469 */
int32_t Py_SLIM_Predict_1vsk(int32_t nrcmds, int32_t nnegs, slim_t *slimhandle, slim_t *trnhandle,
                             int32_t *negitems, int32_t *output, float *scores); /* pyapi.c:483-517 */
int32_t Py_SLIM_Predict(int32_t nrcmds, slim_t *slimhandle, slim_t *trnhandle, int32_t *output,
                        float *scores);                             /* pyapi.c:530-563 */

#ifdef __cplusplus
}
#endif
#endif /* SLIM_B200_SLIM_H */
