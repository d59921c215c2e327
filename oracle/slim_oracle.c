/*
 * slim_oracle.c -- plain-C CPU restatement of the reference SLIM coordinate-descent learner.
 *
 * TEST INFRASTRUCTURE ONLY (see slim_oracle.h).  Not linked into libslim.so, never imported
 * by slim_b200/.  Every routine cites the reference file:line (relative to the KarypisLab/SLIM
 * tree) whose behaviour it restates.  Parity status: PINNED (tests/test_oracle.py checks it
 * bit-for-bit against oracle/_ref/libslim_ref.so and against tests/golden/).
 */
#include "slim_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_EPS 1e-7 /* src/libslim/def.h:14 */

/* ------------------------------------------------------------------------------------------
 * Training-matrix setup.  Restates src/libslim/setup.c:109-135:
 *   ncols = max(rowind)+1                         (setup.c:117)
 *   CSC by serial counting sort, so every column lists users in ascending order
 *                                                 (lib/GKlib/csr.c:1549-1584)
 *   cnorms[i] = (float)sqrt( float-accumulated sum of v*v ), sequential order, no FMA
 *               (csr.c:1929-1931 + gk_mkblas.h:161-170); sqrt(count) when there are no values
 *               (csr.c:1934-1936).
 * ---------------------------------------------------------------------------------------- */
oracle_csc_t *oracle_setup(int32_t nrows, const int64_t *rowptr, const int32_t *rowind,
                           const float *rowval) {
  oracle_csc_t *m = (oracle_csc_t *)calloc(1, sizeof(*m));
  int64_t nnz = rowptr[nrows], k;
  int32_t ncols = 0, i;

  for (k = 0; k < nnz; k++)
    if (rowind[k] + 1 > ncols) ncols = rowind[k] + 1;
  if (nnz == 0) ncols = 0;
  m->nrows = nrows;
  m->ncols = ncols;
  m->colptr = (int64_t *)calloc((size_t)ncols + 1, sizeof(int64_t));
  m->colind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  m->colval = rowval ? (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1)) : NULL;
  m->cnorms = (float *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(float));

  m->rowptr = (int64_t *)malloc(sizeof(int64_t) * ((size_t)nrows + 1));
  m->rowind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  m->rowval = rowval ? (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1)) : NULL;
  memcpy(m->rowptr, rowptr, sizeof(int64_t) * ((size_t)nrows + 1));
  memcpy(m->rowind, rowind, sizeof(int32_t) * (size_t)nnz);
  if (rowval) memcpy(m->rowval, rowval, sizeof(float) * (size_t)nnz);

  for (k = 0; k < nnz; k++) m->colptr[rowind[k] + 1]++;
  for (i = 0; i < ncols; i++) m->colptr[i + 1] += m->colptr[i];
  {
    int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncols + 1));
    memcpy(pos, m->colptr, sizeof(int64_t) * (size_t)(ncols + 1));
    for (i = 0; i < nrows; i++) {
      for (k = rowptr[i]; k < rowptr[i + 1]; k++) {
        int64_t p = pos[rowind[k]]++;
        m->colind[p] = i;
        if (rowval) m->colval[p] = rowval[k];
      }
    }
    free(pos);
  }
  for (i = 0; i < ncols; i++) {
    if (m->colval) {
      volatile float partial = 0.0f; /* volatile: forbid vector reassociation */
      for (k = m->colptr[i]; k < m->colptr[i + 1]; k++) {
        float prod = m->colval[k] * m->colval[k];
        partial = partial + prod;
      }
      m->cnorms[i] = (float)sqrt((double)partial);
    } else {
      m->cnorms[i] = (float)sqrt((double)(m->colptr[i + 1] - m->colptr[i]));
    }
  }
  return m;
}

void oracle_free_csc(oracle_csc_t *m) {
  if (!m) return;
  free(m->colptr);
  free(m->colind);
  free(m->colval);
  free(m->cnorms);
  free(m->rowptr);
  free(m->rowind);
  free(m->rowval);
  free(m);
}

void oracle_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Sparse helpers restating src/libslim/cd.c:24-38 (AddSpVec, with its |x|<=EPS skip) and
 * cd.c:51-65 (SpVecInnerProduct).
 * ---------------------------------------------------------------------------------------- */
static void add_spvec(const oracle_csc_t *m, int32_t c, double xi, double *yhat) {
  int64_t k;
  if (xi > ORACLE_EPS || xi < -ORACLE_EPS) {
    if (m->colval) {
      for (k = m->colptr[c]; k < m->colptr[c + 1]; k++) yhat[m->colind[k]] += xi * m->colval[k];
    } else {
      for (k = m->colptr[c]; k < m->colptr[c + 1]; k++) yhat[m->colind[k]] += xi;
    }
  }
}

static double spvec_dot(const oracle_csc_t *m, int32_t c, const double *yhat) {
  int64_t k;
  double res = 0.0;
  if (m->colval) {
    for (k = m->colptr[c]; k < m->colptr[c + 1]; k++) res += m->colval[k] * yhat[m->colind[k]];
  } else {
    for (k = m->colptr[c]; k < m->colptr[c + 1]; k++) res += yhat[m->colind[k]];
  }
  return res;
}

typedef struct {
  float key;   /* aTy rounded to float: gk_fkv_t.key, lib/GKlib/gk_struct.h:32 */
  int32_t val; /* item id */
} act_t;

/* cd.c:76-86 -- the biased swap shuffle driven by the process-global glibc rand(). */
static void shuffle_ref(act_t *list, int32_t n) {
  int32_t i;
  for (i = 0; i < n; i++) {
    act_t t = list[i];
    int32_t idx = rand() % n;
    list[i] = list[idx];
    list[idx] = t;
  }
}

/* ------------------------------------------------------------------------------------------
 * One target column.  Restates src/libslim/estimate.c:405-505 (target scatter, ATy full sweep,
 * strict >l1r active set with float-rounded key, iteration cap, warm start, compaction) and
 * src/libslim/cd.c:101-142 (the sweeps).
 *
 *   order == ORACLE_ORDER_REF_RAND : literal three-pass update (cd.c:122-129) in the shuffled
 *                                    order -- bit-identical to the reference at nthreads=1.
 *   order == ORACLE_ORDER_ASCENDING: fixed ascending order and the algebraically identical
 *       one-gather form ip = <a_i,yhat> - x_i*sum(v^2), yhat += (x_i' - x_i) a_i that the CUDA
 *       engine uses (same EPS skip semantics).
 * Work arrays x[ncols], aty[ncols], y[nrows], yhat[nrows] must be zero on entry and are zero
 * again on return (estimate.c:520-530).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double *x, *aty, *y, *yhat, *csq;
  act_t *act;
  const int32_t *rank; /* position of every item when sorted by (descending nnz, ascending id) */
  act_t *tmp;
  int32_t *marker; /* fSLIM: candidate slot of an item, -1 when it is none (neighbors.c:36-37) */
  act_t *cand;
} work_t;

/* ------------------------------------------------------------------------------------------
 * fSLIM neighbour search, restating src/libslim/neighbors.c:16-125 (FindColumnNeighbors).
 * Candidates are the items that share a user with column jc, collected in first-encounter order
 * (users of jc ascending, items of a user's row in row order); the similarity accumulates in
 * FLOAT (gk_fkv_t.key), cosine divides by the CANDIDATE's norm only (neighbors.c:82-83) and
 * "jaccard" uses norms, not squared norms (neighbors.c:108-110) -- both kept as they are.
 * Returns min(nnbrs, ncand); the winners are cand[0 .. ret).
 * ---------------------------------------------------------------------------------------- */
static int pop_before(const work_t *w, const act_t *a, const act_t *b) {
  if (a->key != b->key) return a->key > b->key;
  return w->rank[a->val] < w->rank[b->val];
}

static int32_t find_neighbors(const oracle_csc_t *m, const oracle_params_t *p, int32_t jc, work_t *w) {
  int64_t kk, j;
  int32_t n = 0, i, want;
  act_t *c = w->cand;
  if (m->colptr[jc] == m->colptr[jc + 1]) return 0; /* neighbors.c:33-34 */
  for (kk = m->colptr[jc]; kk < m->colptr[jc + 1]; kk++) {
    const int32_t u = m->colind[kk];
    const float cval = m->colval ? m->colval[kk] : 1.0f;
    for (j = m->rowptr[u]; j < m->rowptr[u + 1]; j++) {
      const int32_t it = m->rowind[j];
      if (it == jc) continue;
      if (w->marker[it] == -1) {
        c[n].val = it;
        c[n].key = 0.0f;
        w->marker[it] = n++;
      }
      if (m->rowval) {
        volatile float prod = m->rowval[j] * cval; /* separate multiply and add, as the -std=c99 build */
        c[w->marker[it]].key = c[w->marker[it]].key + prod;
      } else {
        c[w->marker[it]].key = c[w->marker[it]].key + cval;
      }
    }
  }
  if (p->simtype == ORACLE_SIM_COS) {
    for (i = 0; i < n; i++) c[i].key = c[i].key / m->cnorms[c[i].val];
  } else if (p->simtype == ORACLE_SIM_JAC) {
    for (i = 0; i < n; i++) c[i].key = c[i].key / (m->cnorms[c[i].val] + m->cnorms[jc] - c[i].key);
  }
  for (i = 0; i < n; i++) w->marker[c[i].val] = -1;
  want = p->nnbrs < n ? p->nnbrs : n;
  if (n <= want) return n;

  if (p->nbr_ties == ORACLE_TIES_POPULARITY) {
    /* deterministic rule: full order by (similarity desc, popularity rank asc); selection by repeated minimum
       removal would be O(n*k): a heap-free merge sort on the scratch list keeps it O(n log n) */
    int32_t width, a;
    act_t *src = c, *dst = w->tmp;
    for (width = 1; width < n; width *= 2) {
      for (a = 0; a < n; a += 2 * width) {
        int32_t l = a, r = a + width, le = r < n ? r : n, re = a + 2 * width < n ? a + 2 * width : n, o = a;
        while (l < le && r < re) dst[o++] = pop_before(w, &src[r], &src[l]) ? src[r++] : src[l++];
        while (l < le) dst[o++] = src[l++];
        while (r < re) dst[o++] = src[r++];
      }
      { act_t *t = src; src = dst; dst = t; }
    }
    if (src != c) memcpy(c, src, sizeof(act_t) * (size_t)n);
    return want;
  }

  /* ORACLE_TIES_REFERENCE: the selection of lib/GKlib/fkvkselect.c:22-60, step by step -- three-way pivot choice
     between the ends and the middle, pivot parked at the right end, one left-to-right pass that moves every key
     >= pivot to the front, then the side that contains position `want` is searched again. */
  {
    int32_t lo = 0, hi = n - 1;
    while (lo < hi) {
      int32_t mid = lo + ((hi - lo) >> 1), store, scan;
      act_t t;
      float pivot;
      if (c[lo].key < c[mid].key) mid = lo;
      if (c[hi].key > c[mid].key) {
        mid = hi;
        if (c[lo].key < c[mid].key) mid = lo;
      }
      t = c[mid];
      c[mid] = c[hi];
      c[hi] = t;
      pivot = c[hi].key;
      store = lo - 1;
      for (scan = lo; scan < hi; scan++) {
        if (c[scan].key >= pivot) {
          store++;
          t = c[store];
          c[store] = c[scan];
          c[scan] = t;
        }
      }
      store++;
      t = c[store];
      c[store] = c[hi];
      c[hi] = t;
      if (store > want) hi = store - 1;
      else if (store < want) lo = store + 1;
      else break;
    }
  }
  return want;
}

static int32_t solve_column(const oracle_csc_t *m, const oracle_params_t *p, int32_t jc,
                            int32_t incols, const int64_t *icolptr, const int32_t *icolind,
                            const float *icolval, work_t *w, int32_t *out_ind, float *out_val,
                            int32_t *r_niters, int32_t *r_nact, int64_t *r_actnnz, double *r_rnorm,
                            double *r_obj) {
  const int32_t ncols = m->ncols, nrows = m->nrows;
  int64_t k, maxit64;
  int32_t i, t, na = 0, maxit, nnz = 0, niters;
  double l1r = p->l1r, l2r = p->l2r;
  int64_t actnnz = 0;

  /* estimate.c:406-408 */
  for (k = m->colptr[jc]; k < m->colptr[jc + 1]; k++)
    w->y[m->colind[k]] = m->colval ? (double)m->colval[k] : 1.0;

  /* estimate.c:412-421 full sweep; :433-444 active set (ascending i, key = (float)ATy) */
  for (i = 0; i < ncols; i++) {
    double ip = 0.0;
    for (k = m->colptr[i]; k < m->colptr[i + 1]; k++)
      ip += m->colval ? m->colval[k] * w->y[m->colind[k]] : w->y[m->colind[k]];
    w->aty[i] = ip;
  }
  if (p->nnbrs > 0) {
    /* fSLIM, estimate.c:424-431: the active set is the neighbour list, WITHOUT the > l1r filter, keys = (float)ATy.
       No x = -0.1 flags are set on this branch, so a warm-start model is ignored (estimate.c:453-464 writes 0.0). */
    na = find_neighbors(m, p, jc, w);
    for (i = 0; i < na; i++) {
      w->act[i].val = w->cand[i].val;
      w->act[i].key = (float)w->aty[w->cand[i].val];
      actnnz += m->colptr[w->cand[i].val + 1] - m->colptr[w->cand[i].val];
    }
    if (p->order == ORACLE_ORDER_ASCENDING && na > 1) { /* the reference order is by similarity; ours is free */
      int32_t a, b;
      for (a = 1; a < na; a++) {
        act_t key = w->act[a];
        b = a - 1;
        while (b >= 0 && w->act[b].val > key.val) {
          w->act[b + 1] = w->act[b];
          b--;
        }
        w->act[b + 1] = key;
      }
    }
  } else {
    for (i = 0; i < ncols; i++) {
      if (w->aty[i] > l1r && i != jc) {
        w->act[na].val = i;
        w->act[na].key = (float)w->aty[i];
        na++;
        w->x[i] = -0.1; /* estimate.c:440 flag */
        actnnz += m->colptr[i + 1] - m->colptr[i];
      }
    }
  }

  /* ORACLE_ORDER_POPULARITY: reorder the active list once (the order is free, cd.c shuffles it) */
  if (p->order == ORACLE_ORDER_POPULARITY && na > 1) {
    /* counting placement by rank: ranks are distinct, so a sort by rank via a dense scatter */
    int32_t q = 0;
    for (i = 0; i < na; i++) w->tmp[i] = w->act[i];
    /* simple insertion into rank order: na is small compared with ncols; use qsort-free merge */
    {
      int32_t a, b;
      for (a = 1; a < na; a++) {
        act_t key = w->tmp[a];
        b = a - 1;
        while (b >= 0 && w->rank[w->tmp[b].val] > w->rank[key.val]) {
          w->tmp[b + 1] = w->tmp[b];
          b--;
        }
        w->tmp[b + 1] = key;
      }
    }
    for (i = 0; i < na; i++) w->act[i] = w->tmp[i];
    (void)q;
  }

  /* estimate.c:448-449 */
  maxit64 = 50 * (m->colptr[jc + 1] - m->colptr[jc]);
  if (maxit64 > p->maxniters) maxit64 = p->maxniters;
  maxit = (int32_t)maxit64;

  /* estimate.c:453-471 */
  if (icolptr != NULL && jc < incols) {
    for (k = icolptr[jc]; k < icolptr[jc + 1]; k++) {
      int32_t r = icolind[k];
      if (r < 0 || r >= ncols) continue;
      w->x[r] = w->x[r] < 0. ? (double)icolval[k] : 0.0;
    }
  }
  for (i = 0; i < na; i++) {
    int32_t c = w->act[i].val;
    w->x[c] = w->x[c] < 0. ? 0.0 : w->x[c];
  }

  /* cd.c:108-110 */
  for (i = 0; i < na; i++) add_spvec(m, w->act[i].val, w->x[w->act[i].val], w->yhat);

  /* cd.c:112-140 */
  for (t = 0; t < maxit; t++) {
    double dltx = 0.0;
    if (p->order == ORACLE_ORDER_REF_RAND) shuffle_ref(w->act, na);
    for (i = 0; i < na; i++) {
      int32_t c = w->act[i].val;
      double aTy = w->act[i].key;
      double aTa = m->cnorms[c];
      double xi = w->x[c], newxi, ip, num;
      if (p->order == ORACLE_ORDER_REF_RAND) {
        add_spvec(m, c, -xi, w->yhat);
        ip = spvec_dot(m, c, w->yhat);
        num = aTy - ip;
        newxi = num > l1r ? (num - l1r) / ((aTa * aTa) + l2r) : 0.0;
        add_spvec(m, c, newxi, w->yhat);
      } else {
        double in_old = (xi > ORACLE_EPS || xi < -ORACLE_EPS) ? xi : 0.0, in_new, d;
        ip = spvec_dot(m, c, w->yhat) - in_old * w->csq[c];
        num = aTy - ip;
        newxi = num > l1r ? (num - l1r) / ((aTa * aTa) + l2r) : 0.0;
        in_new = (newxi > ORACLE_EPS || newxi < -ORACLE_EPS) ? newxi : 0.0;
        d = in_new - in_old;
        if (d != 0.0) {
          if (m->colval) {
            for (k = m->colptr[c]; k < m->colptr[c + 1]; k++)
              w->yhat[m->colind[k]] += d * m->colval[k];
          } else {
            for (k = m->colptr[c]; k < m->colptr[c + 1]; k++) w->yhat[m->colind[k]] += d;
          }
        }
      }
      w->x[c] = newxi;
      dltx += (newxi - xi) * (newxi - xi);
    }
    if (dltx < p->optTol) break;
  }
  niters = t + 1; /* cd.c:140 (also when the loop ran to the cap or maxit == 0) */

  /* estimate.c:477-489 */
  {
    double rn = 0.0, obj;
    for (i = 0; i < nrows; i++) rn += (w->y[i] - w->yhat[i]) * (w->y[i] - w->yhat[i]);
    rn *= 0.5;
    obj = rn;
    for (i = 0; i < ncols; i++) obj += 0.5 * l2r * (w->x[i] * w->x[i]) + l1r * fabs(w->x[i]);
    if (r_rnorm) *r_rnorm = rn;
    if (r_obj) *r_obj = obj;
  }

  /* estimate.c:492-505 : |x| > EPS, ascending item id, value cast to float */
  for (i = 0; i < ncols; i++) {
    if (fabs(w->x[i]) > ORACLE_EPS) {
      out_ind[nnz] = i;
      out_val[nnz] = (float)w->x[i];
      nnz++;
    }
  }

  /* estimate.c:520-530 */
  for (k = m->colptr[jc]; k < m->colptr[jc + 1]; k++) w->y[m->colind[k]] = 0.0;
  memset(w->x, 0, sizeof(double) * (size_t)ncols);
  memset(w->yhat, 0, sizeof(double) * (size_t)nrows);

  if (r_niters) *r_niters = niters;
  if (r_nact) *r_nact = na;
  if (r_actnnz) *r_actnnz = actnnz;
  return nnz;
}

/* Driver restating the column loop of estimate.c:371-403 and the CSC assembly of SaveModel,
 * estimate.c:570-588 (restricted to the requested columns). */
int oracle_learn(const oracle_csc_t *m, const oracle_params_t *p, const int32_t *cols,
                 int32_t nsel, int32_t incols, const int64_t *icolptr, const int32_t *icolind,
                 const float *icolval, int64_t **wptr, int32_t **wind, float **wval,
                 oracle_stats_t *stats) {
  const int32_t ncols = m->ncols, nrows = m->nrows;
  int32_t s, nthreads = p->nthreads > 0 ? p->nthreads : 1;
  int32_t **linds;
  float **lvals;
  int32_t *lnnz;
  double *csq;
  int64_t *rowlen = NULL;
  int64_t tot;
  int32_t *rank = NULL;

  if (cols == NULL) nsel = ncols;
  if (p->order == ORACLE_ORDER_REF_RAND) nthreads = 1;

  linds = (int32_t **)calloc((size_t)(nsel > 0 ? nsel : 1), sizeof(*linds));
  lvals = (float **)calloc((size_t)(nsel > 0 ? nsel : 1), sizeof(*lvals));
  lnnz = (int32_t *)calloc((size_t)(nsel > 0 ? nsel : 1), sizeof(*lnnz));

  /* exact sum of squares per column (double), used by the one-gather form */
  csq = (double *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(double));
  for (s = 0; s < ncols; s++) {
    int64_t k;
    double a = 0.0;
    if (m->colval)
      for (k = m->colptr[s]; k < m->colptr[s + 1]; k++) a += (double)m->colval[k] * m->colval[k];
    else
      a = (double)(m->colptr[s + 1] - m->colptr[s]);
    csq[s] = a;
  }
  if (p->order == ORACLE_ORDER_POPULARITY || (p->nnbrs > 0 && p->nbr_ties == ORACLE_TIES_POPULARITY)) {
    /* rank[i] = position of item i when items are sorted by (descending nnz, ascending id):
       counting sort over the column lengths, stable in the id */
    int64_t maxc = 0, *start;
    int32_t i;
    for (i = 0; i < ncols; i++)
      if (m->colptr[i + 1] - m->colptr[i] > maxc) maxc = m->colptr[i + 1] - m->colptr[i];
    start = (int64_t *)calloc((size_t)maxc + 2, sizeof(int64_t));
    for (i = 0; i < ncols; i++) start[maxc - (m->colptr[i + 1] - m->colptr[i]) + 1]++;
    for (i = 0; i <= maxc; i++) start[i + 1] += start[i];
    rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncols > 0 ? ncols : 1));
    for (i = 0; i < ncols; i++) rank[i] = (int32_t)start[maxc - (m->colptr[i + 1] - m->colptr[i])]++;
    free(start);
  }
  if (stats && stats->expand_nnz) {
    int64_t k;
    rowlen = (int64_t *)calloc((size_t)(nrows > 0 ? nrows : 1), sizeof(int64_t));
    for (k = 0; k < m->colptr[ncols]; k++) rowlen[m->colind[k]]++;
  }

#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    work_t w;
    int32_t *tind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncols > 0 ? ncols : 1));
    float *tval = (float *)malloc(sizeof(float) * (size_t)(ncols > 0 ? ncols : 1));
    int32_t q;
    w.x = (double *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(double));
    w.aty = (double *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(double));
    w.y = (double *)calloc((size_t)(nrows > 0 ? nrows : 1), sizeof(double));
    w.yhat = (double *)calloc((size_t)(nrows > 0 ? nrows : 1), sizeof(double));
    w.act = (act_t *)malloc(sizeof(act_t) * (size_t)(ncols > 0 ? ncols : 1));
    w.tmp = (act_t *)malloc(sizeof(act_t) * (size_t)(ncols > 0 ? ncols : 1));
    w.csq = csq;
    w.rank = rank;
    w.marker = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncols > 0 ? ncols : 1));
    w.cand = (act_t *)malloc(sizeof(act_t) * (size_t)(ncols > 0 ? ncols : 1));
    memset(w.marker, 0xff, sizeof(int32_t) * (size_t)(ncols > 0 ? ncols : 1));

#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (q = 0; q < nsel; q++) {
      int32_t jc = cols ? cols[q] : q, nnz, nit = 0, na = 0;
      int64_t an = 0;
      double rn = 0.0, ob = 0.0;
      nnz = solve_column(m, p, jc, incols, icolptr, icolind, icolval, &w, tind, tval, &nit, &na,
                         &an, &rn, &ob);
      lnnz[q] = nnz;
      linds[q] = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
      lvals[q] = (float *)malloc(sizeof(float) * (size_t)(nnz > 0 ? nnz : 1));
      memcpy(linds[q], tind, sizeof(int32_t) * (size_t)nnz);
      memcpy(lvals[q], tval, sizeof(float) * (size_t)nnz);
      if (stats) {
        if (stats->niters) stats->niters[q] = nit;
        if (stats->nactive) stats->nactive[q] = na;
        if (stats->active_nnz) stats->active_nnz[q] = an;
        if (stats->rnorm) stats->rnorm[q] = rn;
        if (stats->objval) stats->objval[q] = ob;
        if (stats->expand_nnz) {
          int64_t k, e = 0;
          for (k = m->colptr[jc]; k < m->colptr[jc + 1]; k++) e += rowlen[m->colind[k]];
          stats->expand_nnz[q] = e;
        }
      }
    }
    free(w.x);
    free(w.aty);
    free(w.y);
    free(w.yhat);
    free(w.act);
    free(w.tmp);
    free(w.marker);
    free(w.cand);
    free(tind);
    free(tval);
  }

  tot = 0;
  for (s = 0; s < nsel; s++) tot += lnnz[s];
  *wptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nsel + 1));
  *wind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(tot > 0 ? tot : 1));
  *wval = (float *)malloc(sizeof(float) * (size_t)(tot > 0 ? tot : 1));
  (*wptr)[0] = 0;
  for (s = 0; s < nsel; s++) {
    memcpy(*wind + (*wptr)[s], linds[s], sizeof(int32_t) * (size_t)lnnz[s]);
    memcpy(*wval + (*wptr)[s], lvals[s], sizeof(float) * (size_t)lnnz[s]);
    (*wptr)[s + 1] = (*wptr)[s] + lnnz[s];
    free(linds[s]);
    free(lvals[s]);
  }
  free(linds);
  free(lvals);
  free(lnnz);
  free(csq);
  free(rowlen);
  free(rank);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * CSC -> CSR of the model: the counting-sort transpose of lib/GKlib/csr.c:1546-1584 as called
 * by SaveModel (estimate.c:590).  Row entries come out in ascending column order.
 * ---------------------------------------------------------------------------------------- */
void oracle_transpose(int32_t n, const int64_t *ptr, const int32_t *ind, const float *val,
                      int64_t *tptr, int32_t *tind, float *tval) {
  int32_t i;
  int64_t k;
  int64_t *pos;
  memset(tptr, 0, sizeof(int64_t) * (size_t)(n + 1));
  for (k = 0; k < ptr[n]; k++) tptr[ind[k] + 1]++;
  for (i = 0; i < n; i++) tptr[i + 1] += tptr[i];
  pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  memcpy(pos, tptr, sizeof(int64_t) * (size_t)(n + 1));
  for (i = 0; i < n; i++) {
    for (k = ptr[i]; k < ptr[i + 1]; k++) {
      int64_t q = pos[ind[k]]++;
      tind[q] = i;
      if (val) tval[q] = val[k];
    }
  }
  free(pos);
}

/* ------------------------------------------------------------------------------------------
 * Top-N for one user.  Restates src/libslim/predict.c:15-71: history items are excluded
 * (:34-37,48-49), scores accumulate in FLOAT in history order then model-row order (:40-57),
 * sort descending (:59).  Exact score ties: ascending item id (reference order is
 * implementation-defined there, lib/GKlib/gk_mksort.h:118-269).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float key;
  int32_t val;
} cand_t;

static int cand_cmp(const void *a, const void *b) {
  const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
  if (x->key > y->key) return -1;
  if (x->key < y->key) return 1;
  return (x->val > y->val) - (x->val < y->val);
}

int32_t oracle_topn(int32_t nitems, const int64_t *wrowptr, const int32_t *wrowind,
                    const float *wrowval, int32_t nratings, const int32_t *itemids,
                    const float *ratings, int32_t nrcmds, int32_t *rids, float *rscores) {
  int32_t *marker = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nitems > 0 ? nitems : 1));
  cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)(nitems > 0 ? nitems : 1));
  int32_t r, ncand = 0, n;
  int64_t k;
  for (r = 0; r < nitems; r++) marker[r] = -1;
  for (r = 0; r < nratings; r++)
    if (itemids[r] < nitems && itemids[r] >= 0) marker[itemids[r]] = -2;
  for (r = 0; r < nratings; r++) {
    int32_t i = itemids[r];
    float rating = ratings ? ratings[r] : 1.0f;
    if (i >= nitems || i < 0) continue; /* the reference's test (predict.c:42) can never fire;
                                           ids outside the model are skipped here instead of
                                           reading out of bounds */
    for (k = wrowptr[i]; k < wrowptr[i + 1]; k++) {
      int32_t c = wrowind[k];
      if (marker[c] == -2) continue;
      if (marker[c] == -1) {
        cand[ncand].val = c;
        cand[ncand].key = 0.0f;
        marker[c] = ncand++;
      }
      {
        volatile float prod = rating * wrowval[k]; /* separate multiply and add, as -std=c99 */
        cand[marker[c]].key = cand[marker[c]].key + prod;
      }
    }
  }
  qsort(cand, (size_t)ncand, sizeof(cand_t), cand_cmp);
  n = ncand < nrcmds ? ncand : nrcmds;
  for (r = 0; r < n; r++) {
    rids[r] = cand[r].val;
    rscores[r] = cand[r].key;
  }
  free(marker);
  free(cand);
  return n;
}
