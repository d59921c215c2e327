// api.cpp -- the C ABI of libslim.so: the reference's 20 exported symbols (include/slim.h) plus the
// SLIMB200_* extension (include/slim_b200.h).  Host code only; the learner itself is engine.cu.
//
// What the reference keeps in src/libslim/api.c + pyapi.c (option decoding, matrix wrapping,
// prediction, model I/O) is re-stated here in plain C++ with the same argument meaning and error
// behaviour; the learner entry points route to the CUDA engine and fail (NULL / SLIM_ERROR) when no
// GPU is usable -- there is no CPU learner in this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

// the library is built with -fvisibility=hidden: only the C ABI is exported
#pragma GCC visibility push(default)
#include "../../include/slim_b200.h"
#pragma GCC visibility pop
#include "engine.h"

using namespace slimb200;

namespace {

// Layout of the handle: GKlib's gk_csr_t (reference lib/GKlib/gk_struct.h:75-88), 184 bytes on LP64.
struct CsrHandle {
  int32_t nrows, ncols;
  ssize_t *rowptr, *colptr;
  int32_t *rowind, *colind;
  int32_t *rowids, *colids;
  int32_t *rlabels, *clabels;
  int32_t *rmap, *cmap;
  float *rowval, *colval;
  float *rnorms, *cnorms;
  float *rsums, *csums;
  float *rsizes, *csizes;
  float *rvols, *cvols;
  float *rwgts, *cwgts;
};
static_assert(sizeof(CsrHandle) == 184, "handle must stay layout-compatible with gk_csr_t");

template <class T>
T *xmalloc(size_t n) {
  return static_cast<T *>(malloc(std::max<size_t>(n, 1) * sizeof(T)));
}

CsrHandle *handle_create() { return static_cast<CsrHandle *>(calloc(1, sizeof(CsrHandle))); }

// gk_csr_Free (lib/GKlib/csr.c:47-72): free every array, then the struct.
void handle_free(CsrHandle *h) {
  if (!h) return;
  void *ptrs[] = {h->rowptr, h->colptr, h->rowind, h->colind, h->rowids, h->colids, h->rlabels, h->clabels,
                  h->rmap, h->cmap, h->rowval, h->colval, h->rnorms, h->cnorms, h->rsums, h->csums,
                  h->rsizes, h->csizes, h->rvols, h->cvols, h->rwgts, h->cwgts};
  for (void *p : ptrs) free(p);
  free(h);
}

// Counting-sort transpose; entries of every output segment come out in ascending source order
// (the order gk_csr_CreateIndex produces, lib/GKlib/csr.c:1546-1584).
void transpose(int32_t nsrc, int32_t ndst, const ssize_t *sptr, const int32_t *sind, const float *sval,
               ssize_t *dptr, int32_t *dind, float *dval) {
  std::fill(dptr, dptr + ndst + 1, (ssize_t)0);
  for (ssize_t k = 0; k < sptr[nsrc]; k++) dptr[sind[k] + 1]++;
  for (int32_t i = 0; i < ndst; i++) dptr[i + 1] += dptr[i];
  std::vector<ssize_t> pos(dptr, dptr + ndst);
  for (int32_t i = 0; i < nsrc; i++)
    for (ssize_t k = sptr[i]; k < sptr[i + 1]; k++) {
      const ssize_t q = pos[sind[k]]++;
      dind[q] = i;
      if (sval) dval[q] = sval[k];
    }
}

int32_t max_index_plus1(ssize_t nnz, const int32_t *ind) {
  int32_t mx = -1;
  for (ssize_t k = 0; k < nnz; k++) mx = std::max(mx, ind[k]);
  return mx + 1;  // gk_i32max(...) + 1, setup.c:117 / pyapi.c:26
}

struct Options {
  int32_t nthreads, nnbrs, simtype, dbglvl, algo, ordered, maxniters, nrcmds;
  double l1r, l2r, opttol;
};

// GETOPTION (src/libslim/macros.h:14-15) with the defaults of src/libslim/api.c:42-52.
template <class T>
T getopt(const T *o, int idx, T dflt) {
  return (o == nullptr || o[idx] == (T)-1) ? dflt : o[idx];
}

Options decode(const int32_t *io, const double *dopt) {
  Options p;
  p.nthreads = getopt<int32_t>(io, SLIM_OPTION_NTHREADS, 1);
  p.nnbrs = getopt<int32_t>(io, SLIM_OPTION_NNBRS, 0);
  p.simtype = getopt<int32_t>(io, SLIM_OPTION_SIMTYPE, SLIM_SIMTYPE_COS);
  p.dbglvl = getopt<int32_t>(io, SLIM_OPTION_DBGLVL, 0);
  p.algo = getopt<int32_t>(io, SLIM_OPTION_ALGO, SLIM_ALGO_CD);
  p.ordered = getopt<int32_t>(io, SLIM_OPTION_ORDERED, 0);
  p.maxniters = getopt<int32_t>(io, SLIM_OPTION_MAXNITERS, 10000);
  p.nrcmds = getopt<int32_t>(io, SLIM_OPTION_NRCMDS, 10);
  p.l1r = getopt<double>(dopt, SLIM_OPTION_L1R, 1.0);
  p.l2r = getopt<double>(dopt, SLIM_OPTION_L2R, 1.0);
  p.opttol = getopt<double>(dopt, SLIM_OPTION_OPTTOL, 1e-7);
  return p;
}

// PrintParams, src/libslim/api.c:251-281
void print_params(const Options &p) {
  printf(" Runtime parameters:\n");
  printf("   Model type: %s\n", (p.nnbrs > 0 && p.ordered == 0) ? "fSLIM" : "SLIM");
  printf("   Optimization: l1r: %.2le, l2r: %.2le\n                 optTol: %.2le, maxniters: %d\n", p.l1r,
         p.l2r, p.opttol, p.maxniters);
  printf("   nthreads: %d\n\n", p.nthreads);
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int env_device() {
  const char *v = getenv("SLIMB200_DEVICE");
  return (v && *v) ? atoi(v) : 0;
}

// Model types of the reference (api.c:54-60): nnbrs > 0 with ordered == 0 is fSLIM; `ordered` alone only changes a
// label, and nnbrs > 0 with ordered == 1 ("OFSLIM") runs the plain SLIM path in EstimateModelCD, which tests for
// SLIM_MTYPE_FSLIM only (estimate.c:396, 424).  Same here.
bool supported(const Options &p, int32_t *status) {
  if (p.algo != SLIM_ALGO_CD) {
    fprintf(stderr, "libslim (slim-b200): only algo=cd is implemented on the GPU engine (ADMM needs MKL in the reference)\n");
    if (status) *status = SLIM_ERROR_INPUT;
    return false;
  }
  if (p.nnbrs > 0 && (p.simtype < SLIM_SIMTYPE_COS || p.simtype > SLIM_SIMTYPE_DOTP)) {
    fprintf(stderr, "libslim (slim-b200): unknown similarity measure %d\n", p.simtype);
    if (status) *status = SLIM_ERROR_INPUT;
    return false;
  }
  return true;
}

LearnParams learn_params(const Options &p, double l1r, double l2r) {
  LearnParams lp{l1r, l2r, p.opttol, p.maxniters, p.dbglvl};
  if (p.nnbrs > 0 && p.ordered == 0) {
    lp.nnbrs = p.nnbrs;
    lp.simtype = p.simtype;
  }
  return lp;
}

CsrHandle *assemble(int32_t nitems, const int64_t *colptr, const int32_t *colind, const float *colval) {
  CsrHandle *h = handle_create();
  if (!h) return nullptr;
  const int64_t nnz = colptr[nitems];
  h->nrows = h->ncols = nitems;  // SaveModel(tnnz, ncols, ncols, ...), estimate.c:544
  h->colptr = xmalloc<ssize_t>((size_t)nitems + 1);
  h->colind = xmalloc<int32_t>(nnz);
  h->colval = xmalloc<float>(nnz);
  h->rowptr = xmalloc<ssize_t>((size_t)nitems + 1);
  h->rowind = xmalloc<int32_t>(nnz);
  h->rowval = xmalloc<float>(nnz);
  if (!h->colptr || !h->colind || !h->colval || !h->rowptr || !h->rowind || !h->rowval) {
    handle_free(h);
    return nullptr;
  }
  for (int32_t i = 0; i <= nitems; i++) h->colptr[i] = (ssize_t)colptr[i];
  memcpy(h->colind, colind, sizeof(int32_t) * nnz);
  memcpy(h->colval, colval, sizeof(float) * nnz);
  transpose(nitems, nitems, h->colptr, h->colind, h->colval, h->rowptr, h->rowind, h->rowval);
  return h;
}

// SaveModel (estimate.c:570-593) from a device-resident Result that holds ALL item columns in order: the CSC is
// copied as is, the CSR index (gk_csr_CreateIndex(ROW), csr.c:1546-1584) is built on the GPU (gather.cuh) and
// lands directly in the malloc'd arrays of the handle.
CsrHandle *model_from_result(const Result *res, int32_t *st) {
  int32_t nitems = 0;
  int64_t nnz = 0;
  result_info(res, &nitems, &nnz, nullptr);
  CsrHandle *h = handle_create();
  if (h) {
    h->nrows = h->ncols = nitems;  // SaveModel(tnnz, ncols, ncols, ...), estimate.c:544
    h->colptr = xmalloc<ssize_t>((size_t)nitems + 1);
    h->colind = xmalloc<int32_t>(nnz);
    h->colval = xmalloc<float>(nnz);
    h->rowptr = xmalloc<ssize_t>((size_t)nitems + 1);
    h->rowind = xmalloc<int32_t>(nnz);
    h->rowval = xmalloc<float>(nnz);
  }
  if (!h || !h->colptr || !h->colind || !h->colval || !h->rowptr || !h->rowind || !h->rowval) {
    handle_free(h);
    if (st) *st = SLIM_ERROR_MEMORY;
    return nullptr;
  }
  const int rc = model_to_host(res, h->colptr, h->colind, h->colval, h->rowptr, h->rowind, h->rowval, nullptr);
  if (rc != kOk) {
    handle_free(h);
    if (st) *st = rc;
    return nullptr;
  }
  if (st) *st = SLIM_OK;
  return h;
}

// ---- in-library multi-GPU -----------------------------------------------------------
// Devices used by SLIM_Learn / Py_SLIM_Learn / Py_SLIM_Mselect.  SLIMB200_DEVICES="0,2,3" names them;
// SLIMB200_GPUS=n takes n devices starting at SLIMB200_DEVICE (default 0).  Without either: every visible device
// when the matrix is large enough for the replicated staging to pay (>= 20 M nonzeros), else one.
std::vector<int> learn_devices(int64_t nnz) {
  const int have = device_count();
  std::vector<int> devs;
  if (const char *list = getenv("SLIMB200_DEVICES")) {
    for (const char *p = list; *p;) {
      char *e;
      const long d = strtol(p, &e, 10);
      if (e == p) break;
      if (d >= 0 && d < have && std::find(devs.begin(), devs.end(), (int)d) == devs.end()) devs.push_back((int)d);
      p = (*e == ',') ? e + 1 : e;
    }
    if (!devs.empty()) return devs;
  }
  const int base = env_device();
  int n = 1;
  if (const char *g = getenv("SLIMB200_GPUS")) n = atoi(g);
  else if (nnz >= 20000000) n = have;
  n = std::max(1, std::min(n, std::max(have - base, 1)));
  for (int k = 0; k < n; k++) devs.push_back(base + k);
  return devs;
}

// one NCCL communicator per device, created once per device set and kept for the life of the process
std::mutex g_comm_mutex;
std::map<std::vector<int>, std::vector<Comm *>> g_comms;
std::vector<Comm *> comms_for(const std::vector<int> &devs) {
  std::lock_guard<std::mutex> lock(g_comm_mutex);
  auto it = g_comms.find(devs);
  if (it != g_comms.end()) return it->second;
  std::vector<Comm *> cs(devs.size(), nullptr);
  if (comm_init_all((int)devs.size(), devs.data(), cs.data()) != kOk) return {};
  g_comms[devs] = cs;
  return cs;
}

// columns of rank r: dealt round-robin in descending-nnz order (every rank gets the same head / tail mix)
std::vector<int32_t> shard_of(const std::vector<int32_t> &order, int r, int W) {
  std::vector<int32_t> cols;
  for (size_t k = (size_t)r; k < order.size(); k += (size_t)W) cols.push_back(order[k]);
  std::sort(cols.begin(), cols.end());
  return cols;
}

struct MultiGpu {  // R replicated on every device of `devs`
  std::vector<int> devs;
  std::vector<Matrix *> mats;
  std::vector<Comm *> comms;
  std::vector<std::vector<int32_t>> shards;
  int32_t ncols = 0;
  std::string error;
  int32_t status = SLIM_OK;
  ~MultiGpu() {
    for (Matrix *m : mats) free_matrix(m);
  }
};

bool stage_multi(MultiGpu &g, int32_t nrows, const ssize_t *rowptr, const int32_t *rowind, const float *rowval) {
  const size_t W = g.devs.size();
  g.mats.assign(W, nullptr);
  std::vector<int32_t> st(W, SLIM_OK);
  std::vector<std::string> err(W);
  std::vector<std::thread> th;
  for (size_t r = 0; r < W; r++)
    th.emplace_back([&, r] {
      g.mats[r] = stage(g.devs[r], nrows, rowptr, rowind, rowval, false, 0, &st[r]);
      if (!g.mats[r]) err[r] = last_error();
    });
  for (auto &t : th) t.join();
  for (size_t r = 0; r < W; r++)
    if (!g.mats[r]) {
      g.error = err[r];
      g.status = st[r];
      return false;
    }
  matrix_info(g.mats[0], nullptr, &g.ncols, nullptr, nullptr, nullptr, nullptr);
  std::vector<int32_t> cnt((size_t)std::max(g.ncols, 1));
  matrix_colcounts(g.mats[0], cnt.data());
  std::vector<int32_t> order(g.ncols);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cnt[a] > cnt[b]; });
  g.shards.clear();
  for (size_t r = 0; r < W; r++) g.shards.push_back(shard_of(order, (int)r, (int)W));
  g.comms = comms_for(g.devs);
  if (g.comms.empty()) {
    g.error = last_error();
    g.status = SLIM_ERROR;
    return false;
  }
  return true;
}

// every device solves its shard, then ONE NCCL all-gather assembles W on every device; returns the gathered
// Result of rank 0 (all columns, device resident) and sums the per-column residual / objective statistics
Result *learn_multi(MultiGpu &g, const LearnParams &lp, const WarmStart *wsp, double *error, double *objval) {
  const size_t W = g.devs.size();
  std::vector<Result *> full(W, nullptr);
  std::vector<int32_t> st(W, SLIM_OK);
  std::vector<std::string> err(W);
  std::vector<double> e_r(W, 0.0), o_r(W, 0.0);
  // phase 1: the shards, side by side (no communication: the columns are independent problems)
  std::vector<Result *> loc(W, nullptr);
  {
    std::vector<std::thread> th;
    for (size_t r = 0; r < W; r++)
      th.emplace_back([&, r] {
        const std::vector<int32_t> &cols = g.shards[r];
        loc[r] = learn(g.mats[r], lp, cols.data(), (int32_t)cols.size(), wsp, &st[r]);
        if (!loc[r]) {
          err[r] = last_error();
          return;
        }
        std::vector<double> rn(cols.size()), ob(cols.size());
        result_stats(loc[r], nullptr, nullptr, nullptr, nullptr, rn.data(), ob.data());
        for (size_t k = 0; k < cols.size(); k++) {
          e_r[r] += rn[k];
          o_r[r] += ob[k];
        }
      });
    for (auto &t : th) t.join();
  }
  bool solved = true;
  for (size_t r = 0; r < W; r++) solved = solved && loc[r] != nullptr;
  // phase 2: the all-gather (every rank must take part, so it only starts when every shard was solved)
  if (solved) {
    std::vector<std::thread> th;
    for (size_t r = 0; r < W; r++)
      th.emplace_back([&, r] {
        full[r] = allgather_columns(g.comms[r], loc[r], g.shards[r].data(), g.ncols, &st[r]);
        if (!full[r]) err[r] = last_error();
      });
    for (auto &t : th) t.join();
  }
  for (size_t r = 0; r < W; r++) free_result(loc[r]);
  bool ok = true;
  for (size_t r = 0; r < W; r++)
    if (!full[r]) {
      ok = false;
      g.error = err[r];
      g.status = st[r];
    }
  for (size_t r = 1; r < W; r++) free_result(full[r]);
  if (!ok) {
    free_result(full[0]);
    return nullptr;
  }
  if (error) *error = std::accumulate(e_r.begin(), e_r.end(), 0.0);
  if (objval) *objval = std::accumulate(o_r.begin(), o_r.end(), 0.0);
  return full[0];
}

// Body shared by SLIM_Learn (api.c:33-96) and Py_SLIM_Learn (pyapi.c:134-199).
CsrHandle *learn_entry(int32_t nrows, const ssize_t *rowptr, const int32_t *rowind, const float *rowval,
                       const int32_t *io, const double *dopt, const CsrHandle *imodel, int32_t *r_status) {
  const Options p = decode(io, dopt);
  int32_t st = SLIM_OK;
  if (r_status) *r_status = SLIM_ERROR;
  if (!supported(p, r_status)) return nullptr;
  if (nrows < 0 || !rowptr) {
    if (r_status) *r_status = SLIM_ERROR_INPUT;
    return nullptr;
  }
  if (p.dbglvl & SLIM_DBG_INFO) print_params(p);

  const LearnParams lp = learn_params(p, p.l1r, p.l2r);
  WarmStart ws{};
  const WarmStart *wsp = nullptr;
  if (imodel && imodel->colptr) {
    ws.ncols = imodel->ncols;
    ws.colptr = imodel->colptr;
    ws.colind = imodel->colind;
    ws.colval = imodel->colval;
    wsp = &ws;
  }

  const double t0 = now_s();
  const std::vector<int> devs = learn_devices((int64_t)rowptr[nrows]);
  CsrHandle *model = nullptr;
  double t1 = t0, error = 0.0, objval = 0.0;
  Result *res = nullptr;  // all columns, device resident
  MultiGpu mg;
  Matrix *m = nullptr;
  const char *force_multi = getenv("SLIMB200_FORCE_MULTI");  // tests: the sharded path on a single device
  if (devs.size() > 1 || (force_multi && atoi(force_multi))) {
    // column-sharded over the devices of this process: R replicated, one NCCL all-gather of W at the end
    mg.devs = devs;
    const bool staged = stage_multi(mg, nrows, rowptr, rowind, rowval);
    t1 = now_s();
    if (!staged) {
      fprintf(stderr, "libslim (slim-b200): multi-GPU staging failed: %s\n", mg.error.c_str());
      if (r_status) *r_status = mg.status;
      return nullptr;
    }
    printf("Using Coordinate Descent! \n");  // estimate.c:329
    res = learn_multi(mg, lp, wsp, &error, &objval);
    if (!res) {
      st = mg.status;
      fprintf(stderr, "libslim (slim-b200): multi-GPU learn failed: %s\n", mg.error.c_str());
    }
  } else {
    m = stage(devs.empty() ? env_device() : devs[0], nrows, rowptr, rowind, rowval, false, 0, &st);
    t1 = now_s();
    if (!m) {
      fprintf(stderr, "libslim (slim-b200): staging failed: %s\n", last_error());
      if (r_status) *r_status = st;
      return nullptr;
    }
    printf("Using Coordinate Descent! \n");  // estimate.c:329
    res = learn(m, lp, nullptr, 0, wsp, &st);
    if (res) {
      int32_t nsel = 0;
      result_info(res, &nsel, nullptr, nullptr);
      std::vector<double> rn(nsel), ob(nsel);
      result_stats(res, nullptr, nullptr, nullptr, nullptr, rn.data(), ob.data());
      for (int32_t j = 0; j < nsel; j++) {
        error += rn[j];
        objval += ob[j];
      }
    } else {
      fprintf(stderr, "libslim (slim-b200): learn failed: %s\n", last_error());
    }
  }
  if (res) {
    model = model_from_result(res, &st);
    if (!model) fprintf(stderr, "libslim (slim-b200): model assembly failed: %s\n", last_error());
    if (model && (p.dbglvl & SLIM_DBG_INFO))  // estimate.c:552-555
      printf("Done estimation: loss: %.5le, fit: %.5le, ffrac: %.3lf,  #nzs: %zd\n", objval, error, error / objval,
             model->colptr[model->ncols]);
    free_result(res);
  }
  const double t2 = now_s();
  free_matrix(m);
  if (p.dbglvl & SLIM_DBG_TIME) {  // PrintTimers, timing.c:28-45
    printf("\nTiming Information -------------------------------------------------");
    printf("\n Total: \t %7.3lf", t2 - t0);
    printf("\n   Setup: \t\t %7.3lf", t1 - t0);
    printf("\n   Learn: \t\t %7.3lf", t2 - t1);
    printf("\n********************************************************************\n");
  }
  if (r_status) *r_status = model ? SLIM_OK : st;
  return model;
}

// GetRecommendations, src/libslim/predict.c:15-71.  Scores accumulate in float, in history order then
// model-row order; history items are excluded; descending score, ties by ascending item id (the
// reference's quicksort leaves exact ties unspecified).
struct Cand {
  float key;
  int32_t val;
};

int32_t recommend(const CsrHandle *w, int32_t nratings, const int32_t *itemids, const float *ratings,
                  int32_t nrcmds, int32_t *rids, float *rscores) {
  if (!w || !w->rowptr || nrcmds < 0) return -1;
  const int32_t ncols = w->ncols;
  std::vector<int32_t> marker((size_t)std::max(ncols, 1), -1);
  std::vector<Cand> cand;
  cand.reserve(256);
  for (int32_t r = 0; r < nratings; r++)
    if (itemids[r] < ncols && itemids[r] >= 0) marker[itemids[r]] = -2;
  for (int32_t r = 0; r < nratings; r++) {
    const int32_t i = itemids[r];
    if (i >= w->nrows || i < 0) continue;  // ids outside the model contribute nothing
    const float rating = ratings ? ratings[r] : 1.0f;
    for (ssize_t k = w->rowptr[i]; k < w->rowptr[i + 1]; k++) {
      const int32_t c = w->rowind[k];
      if (marker[c] == -2) continue;
      if (marker[c] == -1) {
        marker[c] = (int32_t)cand.size();
        cand.push_back(Cand{0.0f, c});
      }
      volatile float prod = rating * w->rowval[k];  // separate multiply and add, like the -std=c99 build
      cand[marker[c]].key = cand[marker[c]].key + prod;
    }
  }
  std::sort(cand.begin(), cand.end(), [](const Cand &a, const Cand &b) {
    return a.key > b.key || (a.key == b.key && a.val < b.val);
  });
  const int32_t n = std::min<int32_t>((int32_t)cand.size(), nrcmds);
  for (int32_t r = 0; r < n; r++) {
    rids[r] = cand[r].val;
    rscores[r] = cand[r].key;
  }
  return n;
}

// GetRec_1vsk, src/libslim/predict.c:77-133: scores only the supplied negative/candidate items.
int32_t recommend_1vsk(const CsrHandle *w, int32_t nratings, const int32_t *itemids, const float *ratings,
                       int32_t nrcmds, int32_t *rids, float *rscores, int32_t nnegs, const int32_t *negitems) {
  if (!w || !w->rowptr || nrcmds < 0) return -1;
  const int32_t ncols = w->ncols;
  std::vector<int32_t> marker((size_t)std::max(ncols, 1), -2);
  std::vector<Cand> cand((size_t)std::max(nnegs, 0));
  for (int32_t r = 0; r < nnegs; r++) {
    cand[r] = Cand{0.0f, negitems[r]};
    if (negitems[r] >= 0 && negitems[r] < ncols) marker[negitems[r]] = r;
  }
  for (int32_t r = 0; r < nratings; r++) {
    const int32_t i = itemids[r];
    if (i >= w->nrows || i < 0) continue;
    const float rating = ratings ? ratings[r] : 1.0f;
    for (ssize_t k = w->rowptr[i]; k < w->rowptr[i + 1]; k++) {
      const int32_t c = w->rowind[k];
      if (marker[c] == -2) continue;
      volatile float prod = rating * w->rowval[k];
      cand[marker[c]].key = cand[marker[c]].key + prod;
    }
  }
  std::sort(cand.begin(), cand.end(), [](const Cand &a, const Cand &b) {
    return a.key > b.key || (a.key == b.key && a.val < b.val);
  });
  const int32_t n = std::min<int32_t>((int32_t)cand.size(), nrcmds);
  for (int32_t r = 0; r < n; r++) {
    rids[r] = cand[r].val;
    rscores[r] = cand[r].key;
  }
  return n;
}

// Descending sort by frequency with the tie order of the reference.  SLIM_DetermineHeadAndTail sorts its
// (count, item) pairs with gk_ikvsortd (lib/GKlib/sort.c:256-261), GKlib's instance of the classic libc quicksort
// (lib/GKlib/gk_mksort.h:118-275): median-of-three partitioning with an explicit stack, partitions of at most 8
// elements left to one closing insertion sort.  It is not stable, and items with equal counts around the 50 % mark
// decide which items are "head" -- so the same sequence of comparisons and swaps is followed here, on indices.
struct FreqItem {
  int64_t key;
  int32_t val;
};

void sort_by_frequency_like_reference(std::vector<FreqItem> &a) {
  const ptrdiff_t n = (ptrdiff_t)a.size();
  if (n < 1) return;
  auto before = [&](ptrdiff_t x, ptrdiff_t y) { return a[x].key > a[y].key; };  // "less than" of a descending sort
  constexpr ptrdiff_t kSmall = 8;
  if (n > kSmall) {
    ptrdiff_t lo = 0, hi = n - 1;
    std::vector<std::pair<ptrdiff_t, ptrdiff_t>> todo;
    for (bool more = true; more;) {
      ptrdiff_t mid = lo + ((hi - lo) >> 1);
      if (before(mid, lo)) std::swap(a[mid], a[lo]);
      if (before(hi, mid)) {
        std::swap(a[mid], a[hi]);
        if (before(mid, lo)) std::swap(a[mid], a[lo]);
      }
      ptrdiff_t left = lo + 1, right = hi - 1;
      do {
        while (before(left, mid)) ++left;
        while (before(mid, right)) --right;
        if (left < right) {
          std::swap(a[left], a[right]);
          if (mid == left) mid = right;
          else if (mid == right) mid = left;
          ++left;
          --right;
        } else if (left == right) {
          ++left;
          --right;
          break;
        }
      } while (left <= right);
      // continue with the smaller side, remember the larger one; sides of at most kSmall elements are skipped
      const bool small_l = right - lo <= kSmall, small_r = hi - left <= kSmall;
      if (small_l && small_r) {
        if (todo.empty()) {
          more = false;
        } else {
          lo = todo.back().first;
          hi = todo.back().second;
          todo.pop_back();
        }
      } else if (small_l) {
        lo = left;
      } else if (small_r) {
        hi = right;
      } else if (right - lo > hi - left) {
        todo.emplace_back(lo, right);
        lo = left;
      } else {
        todo.emplace_back(left, hi);
        hi = right;
      }
    }
  }
  // closing insertion sort; the first of the leading kSmall + 1 elements in sort order becomes the sentinel at [0]
  const ptrdiff_t end = n - 1, thresh = std::min<ptrdiff_t>(kSmall, end);
  ptrdiff_t first = 0;
  for (ptrdiff_t run = 1; run <= thresh; ++run)
    if (before(run, first)) first = run;
  if (first != 0) std::swap(a[first], a[0]);
  for (ptrdiff_t run = 2; run <= end; ++run) {
    ptrdiff_t pos = run - 1;
    while (before(run, pos)) --pos;
    ++pos;
    if (pos != run) {
      const FreqItem hold = a[run];
      for (ptrdiff_t k = run; k > pos; --k) a[k] = a[k - 1];
      a[pos] = hold;
    }
  }
}

// SLIM_DetermineHeadAndTail, src/libslim/api.c:215-245
int32_t *head_and_tail(int32_t nrows, int32_t ncols, const ssize_t *rowptr, const int32_t *rowind) {
  int32_t *fm = xmalloc<int32_t>(ncols);
  if (!fm) return nullptr;
  std::vector<FreqItem> cand((size_t)std::max(ncols, 0));
  for (int32_t c = 0; c < ncols; c++) {
    fm[c] = 1;
    cand[c] = FreqItem{0, c};
  }
  for (ssize_t k = 0; k < rowptr[nrows]; k++)
    if (rowind[k] >= 0 && rowind[k] < ncols) cand[rowind[k]].key++;
  sort_by_frequency_like_reference(cand);
  ssize_t left = rowptr[nrows] / 2;
  for (int32_t c = 0; c < ncols && left > 0; c++) {
    fm[cand[c].val] = 0;
    left -= cand[c].key;
  }
  return fm;
}

CsrHandle *wrap_csr(int32_t nrows, const ssize_t *rowptr, const int32_t *rowind, const float *rowval) {
  CsrHandle *h = handle_create();
  if (!h) return nullptr;
  const ssize_t nnz = rowptr[nrows];
  h->nrows = nrows;
  h->ncols = max_index_plus1(nnz, rowind);
  h->rowptr = xmalloc<ssize_t>((size_t)nrows + 1);
  h->rowind = xmalloc<int32_t>(nnz);
  if (rowval) h->rowval = xmalloc<float>(nnz);
  if (!h->rowptr || !h->rowind || (rowval && !h->rowval)) {
    handle_free(h);
    return nullptr;
  }
  memcpy(h->rowptr, rowptr, sizeof(ssize_t) * ((size_t)nrows + 1));
  memcpy(h->rowind, rowind, sizeof(int32_t) * nnz);
  if (rowval) memcpy(h->rowval, rowval, sizeof(float) * nnz);
  return h;
}

}  // namespace

extern "C" {

int32_t SLIM_iSetDefaults(int32_t *options) {
  for (int i = 0; i < SLIM_NOPTIONS; i++) options[i] = -1;
  return SLIM_OK;
}

int32_t SLIM_dSetDefaults(double *options) {
  for (int i = 0; i < SLIM_NOPTIONS; i++) options[i] = -1;
  return SLIM_OK;
}

slim_t *SLIM_Learn(int32_t nrows, ssize_t *rowptr, int32_t *rowind, float *rowval, int32_t *ioptions,
                   double *doptions, slim_t *imodel, int32_t *r_status) {
  return learn_entry(nrows, rowptr, rowind, rowval, ioptions, doptions,
                     static_cast<const CsrHandle *>(imodel), r_status);
}

int32_t SLIM_GetTopN(slim_t *model, int32_t nratings, int32_t *itemids, float *ratings, int32_t *ioptions,
                     int32_t nrcmds, int32_t *rids, float *rscores) {
  (void)ioptions;
  const int32_t n = recommend(static_cast<const CsrHandle *>(model), nratings, itemids, ratings, nrcmds,
                              rids, rscores);
  return n < 0 ? SLIM_ERROR : n;
}

// binary row format, lib/GKlib/csr.c:825-838
int32_t SLIM_WriteModel(slim_t *model, char *filename) {
  const CsrHandle *h = static_cast<const CsrHandle *>(model);
  FILE *f = fopen(filename, "wb");
  if (!f || !h || !h->rowptr) {
    if (f) fclose(f);
    return SLIM_ERROR;
  }
  const ssize_t nnz = h->rowptr[h->nrows];
  fwrite(&h->nrows, sizeof(int32_t), 1, f);
  fwrite(&h->ncols, sizeof(int32_t), 1, f);
  fwrite(h->rowptr, sizeof(ssize_t), (size_t)h->nrows + 1, f);
  fwrite(h->rowind, sizeof(int32_t), nnz, f);
  if (h->rowval) fwrite(h->rowval, sizeof(float), nnz, f);
  fclose(f);
  return SLIM_OK;
}

// lib/GKlib/csr.c:436-460 + CreateIndex(COL) (api.c:188-193)
slim_t *SLIM_ReadModel(char *filename) {
  FILE *f = fopen(filename, "rb");
  if (!f) return nullptr;
  CsrHandle *h = handle_create();
  if (!h) {
    fclose(f);
    return nullptr;
  }
  bool ok = fread(&h->nrows, sizeof(int32_t), 1, f) == 1 && fread(&h->ncols, sizeof(int32_t), 1, f) == 1 &&
            h->nrows >= 0 && h->ncols >= 0;
  if (ok) {
    h->rowptr = xmalloc<ssize_t>((size_t)h->nrows + 1);
    ok = h->rowptr && fread(h->rowptr, sizeof(ssize_t), (size_t)h->nrows + 1, f) == (size_t)h->nrows + 1;
  }
  if (ok) {  // the row pointer must be a monotone prefix sum starting at 0
    ok = h->rowptr[0] == 0;
    for (int32_t i = 0; ok && i < h->nrows; i++) ok = h->rowptr[i + 1] >= h->rowptr[i];
  }
  if (ok) {
    const ssize_t nnz = h->rowptr[h->nrows];
    h->rowind = xmalloc<int32_t>(nnz);
    h->rowval = xmalloc<float>(nnz);
    ok = h->rowind && h->rowval && fread(h->rowind, sizeof(int32_t), nnz, f) == (size_t)nnz &&
         fread(h->rowval, sizeof(float), nnz, f) == (size_t)nnz;
    for (ssize_t k = 0; ok && k < nnz; k++) ok = h->rowind[k] >= 0 && h->rowind[k] < h->ncols;
    if (ok) {
      h->colptr = xmalloc<ssize_t>((size_t)h->ncols + 1);
      h->colind = xmalloc<int32_t>(nnz);
      h->colval = xmalloc<float>(nnz);
      ok = h->colptr && h->colind && h->colval;
      if (ok) transpose(h->nrows, h->ncols, h->rowptr, h->rowind, h->rowval, h->colptr, h->colind, h->colval);
    }
  }
  fclose(f);
  if (!ok) {
    handle_free(h);
    return nullptr;
  }
  return h;
}

void SLIM_FreeModel(slim_t **model) {
  if (!model) return;
  handle_free(static_cast<CsrHandle *>(*model));
  *model = nullptr;
}

int32_t *SLIM_DetermineHeadAndTail(int32_t nrows, int32_t ncols, ssize_t *rowptr, int32_t *rowind) {
  return head_and_tail(nrows, ncols, rowptr, rowind);
}

int32_t Py_csr_wrapper(int32_t nrows, ssize_t *rowptr, int32_t *rowind, float *rowval, slim_t **matrix_out) {
  CsrHandle *h = wrap_csr(nrows, rowptr, rowind, rowval);
  *matrix_out = h;
  return h ? SLIM_OK : SLIM_ERROR_MEMORY;
}

// text CSR, lib/GKlib/csr.c:904-923: " %d %f" pairs, one row per line
int32_t Py_csr_save(slim_t *mathandle, char *fname) {
  const CsrHandle *h = static_cast<const CsrHandle *>(mathandle);
  FILE *f = fopen(fname, "w");
  if (!f) return SLIM_ERROR;
  for (int32_t i = 0; i < h->nrows; i++) {
    for (ssize_t k = h->rowptr[i]; k < h->rowptr[i + 1]; k++) {
      fprintf(f, " %d", h->rowind[k]);
      fprintf(f, " %f", h->rowval ? h->rowval[k] : 1.0f);
    }
    fprintf(f, "\n");
  }
  fclose(f);
  return SLIM_OK;
}

// text CSR with values, 0-based (lib/GKlib/csr.c:655-769 as called from pyapi.c:61)
int32_t Py_csr_load(slim_t **mathandle, char *fname) {
  FILE *f = fopen(fname, "r");
  if (!f) {
    *mathandle = nullptr;
    return SLIM_ERROR;
  }
  std::vector<ssize_t> rp(1, 0);
  std::vector<int32_t> ri;
  std::vector<float> rv;
  char *line = nullptr;
  size_t cap = 0;
  while (getline(&line, &cap, f) >= 0) {
    char *s = line;
    for (;;) {
      char *e;
      const long c = strtol(s, &e, 10);
      if (e == s) break;
      s = e;
      const float v = strtof(s, &e);
      if (e == s) break;
      s = e;
      ri.push_back((int32_t)c);
      rv.push_back(v);
    }
    rp.push_back((ssize_t)ri.size());
  }
  free(line);
  fclose(f);
  CsrHandle *h = wrap_csr((int32_t)rp.size() - 1, rp.data(), ri.data(), rv.data());
  *mathandle = h;
  return h ? SLIM_OK : SLIM_ERROR_MEMORY;
}

int32_t Py_csr_free(slim_t *mathandle) {
  handle_free(static_cast<CsrHandle *>(mathandle));
  return SLIM_OK;
}

int32_t Py_csr_stat(slim_t *mathandle, int32_t *nnz) {
  const CsrHandle *h = static_cast<const CsrHandle *>(mathandle);
  *nnz = (int32_t)h->rowptr[h->nrows];
  return SLIM_OK;
}

int32_t Py_csr_export(slim_t *mathandle, int32_t *indptr, int32_t *indices, float *data) {
  const CsrHandle *h = static_cast<const CsrHandle *>(mathandle);
  const ssize_t nnz = h->rowptr[h->nrows];
  for (int32_t i = 0; i <= h->nrows; i++) indptr[i] = (int32_t)h->rowptr[i];
  for (ssize_t k = 0; k < nnz; k++) indices[k] = h->rowind[k];
  if (h->rowval)
    for (ssize_t k = 0; k < nnz; k++) data[k] = h->rowval[k];
  return SLIM_OK;
}

int32_t Py_SLIM_Learn(slim_t *trnhandle, int32_t *ioptions, double *doptions, slim_t **model_out) {
  const CsrHandle *t = static_cast<const CsrHandle *>(trnhandle);
  int32_t st = SLIM_ERROR;
  CsrHandle *m = learn_entry(t->nrows, t->rowptr, t->rowind, t->rowval, ioptions, doptions, nullptr, &st);
  *model_out = m;
  return m ? SLIM_OK : st;
}

int32_t Py_SLIM_GetTopN(slim_t *model, int32_t nratings, int32_t *itemids, float *ratings, int32_t nrcmds,
                        int32_t *rids, float *rscores, int32_t dbglvl) {
  (void)dbglvl;
  const int32_t n = recommend(static_cast<const CsrHandle *>(model), nratings, itemids, ratings, nrcmds,
                              rids, rscores);
  return n < 0 ? SLIM_ERROR : n;
}

int32_t Py_SLIM_GetTopN_1vsk(slim_t *model, int32_t nratings, int32_t *itemids, float *ratings, int32_t nrcmds,
                             int32_t *rids, float *rscores, int32_t nnegs, int32_t *negitems, int32_t dbglvl) {
  (void)dbglvl;
  const int32_t n = recommend_1vsk(static_cast<const CsrHandle *>(model), nratings, itemids, ratings, nrcmds,
                                   rids, rscores, nnegs, negitems);
  return n < 0 ? SLIM_ERROR : n;
}

int32_t Py_SLIM_Predict_1vsk(int32_t nrcmds, int32_t nnegs, slim_t *slimhandle, slim_t *trnhandle,
                             int32_t *negitems, int32_t *output, float *scores) {
  const CsrHandle *w = static_cast<const CsrHandle *>(slimhandle);
  const CsrHandle *t = static_cast<const CsrHandle *>(trnhandle);
  std::vector<int32_t> rids((size_t)std::max(nrcmds, 1));
  std::vector<float> rsc((size_t)std::max(nrcmds, 1));
  int32_t nvalid = 0;
  for (int32_t u = 0; u < t->nrows; u++) {
    const ssize_t a = t->rowptr[u], b = t->rowptr[u + 1];
    const int32_t n = recommend_1vsk(w, (int32_t)(b - a), t->rowind + a, t->rowval ? t->rowval + a : nullptr,
                                     nrcmds, rids.data(), rsc.data(), nnegs, negitems + (size_t)u * nnegs);
    if (n >= 0) {
      for (int32_t r = 0; r < n; r++) {
        output[(size_t)u * nrcmds + r] = rids[r];
        scores[(size_t)u * nrcmds + r] = rsc[r];
      }
      nvalid++;
    }
  }
  return nvalid < 1 ? SLIM_ERROR : SLIM_OK;
}

int32_t Py_SLIM_Predict(int32_t nrcmds, slim_t *slimhandle, slim_t *trnhandle, int32_t *output, float *scores) {
  const CsrHandle *w = static_cast<const CsrHandle *>(slimhandle);
  const CsrHandle *t = static_cast<const CsrHandle *>(trnhandle);
  if (!w || !t || !w->rowptr || !t->rowptr) return SLIM_ERROR;
  // batched on the GPU (predict.cuh): one CTA per user, lists and scores bit-identical to recommend() below; a
  // failing GPU call is an error.  The per-user host loop of the reference (pyapi.c:530-563) runs only when it is
  // asked for (SLIMB200_PREDICT_HOST=1) or when the process has no CUDA device at all (prediction is not the
  // learner: a model can be served on a host without a GPU, exactly as with the reference library).
  const char *force_host = getenv("SLIMB200_PREDICT_HOST");
  if (nrcmds > 0 && t->nrows > 0 && device_count() > 0 && !(force_host && atoi(force_host))) {
    const int rc = predict_topn(env_device(), w->nrows, w->ncols, w->rowptr, w->rowind, w->rowval, t->nrows, t->rowptr,
                                t->rowind, t->rowval, nrcmds, output, scores, nullptr, nullptr);
    if (rc != kOk) fprintf(stderr, "libslim (slim-b200): GPU top-N failed: %s\n", last_error());
    return rc == kOk ? SLIM_OK : SLIM_ERROR;  // no silent host fallback once a device was chosen
  }
  std::vector<int32_t> rids((size_t)std::max(nrcmds, 1));
  std::vector<float> rsc((size_t)std::max(nrcmds, 1));
  int32_t nvalid = 0;
  for (int32_t u = 0; u < t->nrows; u++) {
    const ssize_t a = t->rowptr[u], b = t->rowptr[u + 1];
    const int32_t n = recommend(w, (int32_t)(b - a), t->rowind + a, t->rowval ? t->rowval + a : nullptr, nrcmds,
                                rids.data(), rsc.data());
    if (n >= 0) {
      for (int32_t r = 0; r < n; r++) {
        output[(size_t)u * nrcmds + r] = rids[r];
        scores[(size_t)u * nrcmds + r] = rsc[r];
      }
      nvalid++;
    }
  }
  return nvalid < 1 ? SLIM_ERROR : SLIM_OK;
}

// Py_SLIM_Mselect, pyapi.c:214-412: (l1, l2) grid with warm starts; R stays staged in HBM for the
// whole grid (the reference re-runs CreateTrainingMatrix inside every SLIM_Learn call).
int32_t Py_SLIM_Mselect(slim_t *trnhandle, slim_t *tsthandle, int32_t *ioptions, double *doptions,
                        double *arrayl1, double *arrayl2, int32_t nl1, int32_t nl2, double *bestl1HR,
                        double *bestl2HR, double *bestHRHR, double *bestARHR, double *bestl1AR,
                        double *bestl2AR, double *bestHRAR, double *bestARAR) {
  const CsrHandle *trn = static_cast<const CsrHandle *>(trnhandle);
  const CsrHandle *tst = static_cast<const CsrHandle *>(tsthandle);
  Options p = decode(ioptions, doptions);
  int32_t st = SLIM_OK;
  if (!supported(p, &st)) return st;
  const int32_t ncols = std::max(max_index_plus1(trn->rowptr[trn->nrows], trn->rowind),
                                 max_index_plus1(tst->rowptr[tst->nrows], tst->rowind));
  printf("------------------------------------------------------------------\n");
  printf("SLIM, version %s\n", SLIM_VERSION);
  printf("------------------------------------------------------------------\n");
  printf("  trn matrix, nrows: %d, ncols: %d, nnz: %zd\n", trn->nrows, ncols, trn->rowptr[trn->nrows]);
  printf("  tst matrix, nrows: %d, ncols: %d, nnz: %zd\n", tst->nrows, tst->ncols, tst->rowptr[tst->nrows]);
  printf("  optTol: %.2le, niters: %d\n", p.opttol, p.maxniters);
  printf("\nEstimating & evaluating models...\n\n");

  Matrix *m = stage(env_device(), trn->nrows, trn->rowptr, trn->rowind, trn->rowval, false, 0, &st);
  if (!m) {
    fprintf(stderr, "libslim (slim-b200): staging failed: %s\n", last_error());
    return st;
  }
  int32_t *fm = head_and_tail(trn->nrows, ncols, trn->rowptr, trn->rowind);
  std::vector<int32_t> rids((size_t)std::max(p.nrcmds, 1)), rmarker((size_t)std::max(ncols, 1), -1);
  std::vector<float> rsc((size_t)std::max(p.nrcmds, 1));
  CsrHandle *model = nullptr;
  *bestHRHR = *bestARHR = *bestHRAR = *bestARAR = 0.0;
  int32_t rc = SLIM_OK;
  std::vector<int32_t> all_ids, all_cnt;
  std::vector<float> all_sc;
  for (int32_t i1 = 0; i1 < nl1 && rc == SLIM_OK; i1++) {
    for (int32_t i2 = 0; i2 < nl2; i2++) {
      doptions[SLIM_OPTION_L1R] = arrayl1[i1];
      doptions[SLIM_OPTION_L2R] = arrayl2[i2];
      const double t0 = now_s();
      const LearnParams lp = learn_params(p, arrayl1[i1], arrayl2[i2]);
      WarmStart ws{};
      if (model) ws = WarmStart{model->ncols, model->colptr, model->colind, model->colval};
      printf("Using Coordinate Descent! \n");
      Result *res = learn(m, lp, nullptr, 0, model ? &ws : nullptr, &st);
      CsrHandle *next = nullptr;
      if (res) {
        next = model_from_result(res, &st);
        free_result(res);
      }
      handle_free(model);
      model = next;
      const double t1 = now_s();
      if (!model) {
        printf("ERROR: Something went wrong with model estimation [%.3le %.3le]: rstatus: %d\n", arrayl1[i1],
               arrayl2[i2], st);
        continue;
      }
      // evaluation, pyapi.c:306-366
      std::fill(rmarker.begin(), rmarker.end(), -1);
      float hr[3] = {0, 0, 0}, arhr = 0.0f;  // float accumulators, as pyapi.c:231 declares them
      int32_t nvalid = 0, nvalid_head = 0, nvalid_tail = 0;
      // top-N lists of all users in one batched GPU call (predict.cuh); SLIMB200_PREDICT_HOST=1 asks for the
      // per-user loop of the reference instead.  A failing GPU call ends the grid with an error (no silent fallback).
      bool gpu_lists = false;
      {
        const char *force_host = getenv("SLIMB200_PREDICT_HOST");
        if (p.nrcmds > 0 && trn->nrows > 0 && !(force_host && atoi(force_host))) {
          all_ids.assign((size_t)trn->nrows * p.nrcmds, -1);
          all_sc.assign((size_t)trn->nrows * p.nrcmds, 0.f);
          all_cnt.assign((size_t)trn->nrows, 0);
          gpu_lists = predict_topn(env_device(), model->nrows, model->ncols, model->rowptr, model->rowind,
                                   model->rowval, trn->nrows, trn->rowptr, trn->rowind, trn->rowval, p.nrcmds,
                                   all_ids.data(), all_sc.data(), all_cnt.data(), nullptr) == kOk;
          if (!gpu_lists) {
            fprintf(stderr, "libslim (slim-b200): GPU top-N failed: %s\n", last_error());
            *bestl1HR = arrayl1[i1];
            *bestl2HR = arrayl2[i2];
            rc = SLIM_ERROR;
            break;
          }
        }
      }
      for (int32_t u = 0; u < trn->nrows; u++) {
        if (u >= tst->nrows || tst->rowptr[u + 1] - tst->rowptr[u] < 1) continue;
        const ssize_t a = trn->rowptr[u], b = trn->rowptr[u + 1];
        int32_t n;
        const int32_t *rids_u = rids.data();
        if (gpu_lists) {
          n = all_cnt[u];
          rids_u = all_ids.data() + (size_t)u * p.nrcmds;
        } else {
          n = recommend(model, (int32_t)(b - a), trn->rowind + a, trn->rowval ? trn->rowval + a : nullptr, p.nrcmds,
                        rids.data(), rsc.data());
        }
        nvalid += n >= 0 ? 1 : 0;
        int is_tail = 0, is_head = 0, ntrue[2] = {0, 0}, nhits[3] = {0, 0, 0};
        float larhr = 0.0f, baseline = 0.0f;
        for (ssize_t z = tst->rowptr[u]; z < tst->rowptr[u + 1]; z++) {
          rmarker[tst->rowind[z]] = u;
          ntrue[fm[tst->rowind[z]]]++;
          if (fm[tst->rowind[z]]) is_tail = 1; else is_head = 1;
          baseline += 1.0 / (1.0 + z - tst->rowptr[u]);
        }
        nvalid_tail += is_tail;
        nvalid_head += is_head;
        for (int32_t r = 0; r < n; r++)
          if (rmarker[rids_u[r]] == u) {
            nhits[fm[rids_u[r]]]++;
            nhits[2]++;
            larhr += 1.0 / (1.0 + r);
          }
        hr[0] += nhits[0] > 0 ? 1.0 * nhits[0] / ntrue[0] : 0.0;
        hr[1] += nhits[1] > 0 ? 1.0 * nhits[1] / ntrue[1] : 0.0;
        hr[2] += 1.0 * nhits[2] / (tst->rowptr[u + 1] - tst->rowptr[u]);
        arhr += larhr / baseline;
      }
      const float all_hr = nvalid > 0 ? hr[2] / nvalid : 0;
      const float head_hr = nvalid_head > 0 ? hr[0] / nvalid_head : 0;
      const float tail_hr = nvalid_tail > 0 ? hr[1] / nvalid_tail : 0;
      const float ar = nvalid > 0 ? arhr / nvalid : 0;
      printf("\nnvalid: %d nvalid_head: %d nvalid_tail: %d", nvalid, nvalid_head, nvalid_tail);
      printf("\nl1r: %.2le l2r: %.2le nnz: %7zd hr: %.4f hr_head: %.4f hr_tail: %.4f arhr: %.4f time: %.2lf\n",
             arrayl1[i1], arrayl2[i2], model->rowptr[model->nrows], all_hr, head_hr, tail_hr, ar, t1 - t0);
      if (nvalid < 1) {
        *bestl1HR = arrayl1[i1];
        *bestl2HR = arrayl2[i2];
        rc = SLIM_ERROR;
        break;
      }
      if (all_hr > *bestHRHR) {
        *bestHRHR = all_hr;
        *bestARHR = ar;
        *bestl1HR = arrayl1[i1];
        *bestl2HR = arrayl2[i2];
      }
      if (ar > *bestARAR) {
        *bestHRAR = all_hr;
        *bestARAR = ar;
        *bestl1AR = arrayl1[i1];
        *bestl2AR = arrayl2[i2];
      }
    }
  }
  if (rc == SLIM_OK) {
    printf("\nDone.\n");
    printf("------------------------------------------------------------------\n");
  }
  handle_free(model);
  free(fm);
  free_matrix(m);
  return rc;
}

// ---- include/slim_b200.h ------------------------------------------------------------------------

int32_t SLIMB200_DeviceCount(void) { return device_count(); }
const char *SLIMB200_LastError(void) { return last_error(); }

slimb200_matrix_t *SLIMB200_Stage(int32_t device, int32_t nrows, const ssize_t *rowptr, const int32_t *rowind,
                                  const float *rowval, int32_t *r_status) {
  return reinterpret_cast<slimb200_matrix_t *>(stage(device, nrows, rowptr, rowind, rowval, false, 0, r_status));
}

slimb200_matrix_t *SLIMB200_StageDevice(int32_t device, int32_t nrows, int64_t nnz, const int64_t *d_rowptr,
                                        const int32_t *d_rowind, const float *d_rowval, int32_t *r_status) {
  return reinterpret_cast<slimb200_matrix_t *>(
      stage(device, nrows, reinterpret_cast<const ssize_t *>(d_rowptr), d_rowind, d_rowval, true, nnz, r_status));
}

void SLIMB200_FreeMatrix(slimb200_matrix_t **matrix) {
  if (!matrix) return;
  free_matrix(reinterpret_cast<Matrix *>(*matrix));
  *matrix = nullptr;
}

int32_t SLIMB200_MatrixInfo(const slimb200_matrix_t *matrix, int32_t *nrows, int32_t *ncols, int64_t *nnz,
                            int32_t *device, double *stage_ms, int32_t *stage_launches) {
  if (!matrix) return SLIM_ERROR_INPUT;
  matrix_info(reinterpret_cast<const Matrix *>(matrix), nrows, ncols, nnz, device, stage_ms, stage_launches);
  return SLIM_OK;
}

int32_t SLIMB200_MatrixCSC(const slimb200_matrix_t *matrix, int64_t *colptr, int32_t *colind, float *colval,
                           float *cnorms) {
  if (!matrix) return SLIM_ERROR_INPUT;
  return matrix_csc_to_host(reinterpret_cast<const Matrix *>(matrix), colptr, colind, colval, cnorms);
}

int32_t SLIMB200_MatrixItemOrder(const slimb200_matrix_t *matrix, int32_t *rank) {
  if (!matrix) return SLIM_ERROR_INPUT;
  return matrix_item_order_to_host(reinterpret_cast<const Matrix *>(matrix), rank);
}

int32_t SLIMB200_MatrixWindowGram(const slimb200_matrix_t *matrix, double *out) {
  if (!matrix) return SLIM_ERROR_INPUT;
  return matrix_window_gram_to_host(reinterpret_cast<const Matrix *>(matrix), out);
}

int32_t SLIMB200_MatrixGramInfo(const slimb200_matrix_t *matrix, int32_t *elem_bytes, double *build_ms) {
  if (!matrix) return SLIM_ERROR_INPUT;
  matrix_gram_info(reinterpret_cast<const Matrix *>(matrix), elem_bytes, build_ms);
  return SLIM_OK;
}

int32_t SLIMB200_MatrixGramLayout(const slimb200_matrix_t *matrix, int64_t *bytes, int32_t *h32, int32_t *h16) {
  if (!matrix) return SLIM_ERROR_INPUT;
  matrix_gram_layout(reinterpret_cast<const Matrix *>(matrix), bytes, h32, h16);
  return SLIM_OK;
}

int32_t SLIMB200_MatrixGramStair(const slimb200_matrix_t *matrix, int32_t *stair, int32_t *hd) {
  if (!matrix) return SLIM_ERROR_INPUT;
  matrix_gram_stair(reinterpret_cast<const Matrix *>(matrix), stair, hd);
  return SLIM_OK;
}

int32_t SLIMB200_MatrixGram(const slimb200_matrix_t *matrix, void *out) {
  if (!matrix || !out) return SLIM_ERROR_INPUT;
  return matrix_gram_to_host(reinterpret_cast<const Matrix *>(matrix), out);
}

slimb200_result_t *SLIMB200_LearnColumns(slimb200_matrix_t *matrix, const int32_t *ioptions, const double *doptions,
                                         const int32_t *cols, int32_t ncols_sel, const slim_t *imodel,
                                         int32_t *r_status) {
  if (r_status) *r_status = SLIM_ERROR_INPUT;
  if (!matrix) return nullptr;
  const Options p = decode(ioptions, doptions);
  if (!supported(p, r_status)) return nullptr;
  const LearnParams lp = learn_params(p, p.l1r, p.l2r);
  const CsrHandle *im = static_cast<const CsrHandle *>(imodel);
  WarmStart ws{};
  if (im && im->colptr) ws = WarmStart{im->ncols, im->colptr, im->colind, im->colval};
  return reinterpret_cast<slimb200_result_t *>(
      learn(reinterpret_cast<Matrix *>(matrix), lp, cols, ncols_sel, (im && im->colptr) ? &ws : nullptr, r_status));
}

void SLIMB200_FreeResult(slimb200_result_t **result) {
  if (!result) return;
  free_result(reinterpret_cast<Result *>(*result));
  *result = nullptr;
}

int32_t SLIMB200_ResultInfo(const slimb200_result_t *result, int32_t *nsel, int64_t *nnz, double *solve_ms,
                            double *gather_ms, int32_t *launches) {
  if (!result) return SLIM_ERROR_INPUT;
  Timings t{};
  result_info(reinterpret_cast<const Result *>(result), nsel, nnz, &t);
  if (solve_ms) *solve_ms = t.solve_ms;
  if (gather_ms) *gather_ms = t.gather_ms;
  if (launches) *launches = t.launches;
  return SLIM_OK;
}

int32_t SLIMB200_ResultStats(const slimb200_result_t *result, int32_t *niters, int32_t *nactive,
                             int64_t *active_nnz, int64_t *expand_nnz, double *rnorm, double *objval) {
  if (!result) return SLIM_ERROR_INPUT;
  return result_stats(reinterpret_cast<const Result *>(result), niters, nactive, active_nnz, expand_nnz, rnorm,
                      objval);
}

int32_t SLIMB200_ResultPhases(const slimb200_result_t *result, float *phase_us, int32_t *rounds) {
  if (!result) return SLIM_ERROR_INPUT;
  return result_phases(reinterpret_cast<const Result *>(result), phase_us, rounds);
}

int32_t SLIMB200_ResultToHost(const slimb200_result_t *result, int64_t *colptr, int32_t *colind, float *colval) {
  if (!result) return SLIM_ERROR_INPUT;
  return result_to_host(reinterpret_cast<const Result *>(result), colptr, colind, colval);
}

int32_t SLIMB200_ResultToDevice(const slimb200_result_t *result, int32_t *d_counts, int32_t *d_colind,
                                float *d_colval) {
  if (!result) return SLIM_ERROR_INPUT;
  return result_to_device(reinterpret_cast<const Result *>(result), d_counts, d_colind, d_colval);
}

// ---- multi-GPU ------------------------------------------------------------------------------------
int32_t SLIMB200_CommUniqueId(void *id128) { return id128 ? comm_unique_id(id128) : SLIM_ERROR_INPUT; }

slimb200_comm_t *SLIMB200_CommInitRank(int32_t device, int32_t nranks, int32_t rank, const void *id128,
                                       int32_t *r_status) {
  return reinterpret_cast<slimb200_comm_t *>(comm_init(device, nranks, rank, id128, r_status));
}

void SLIMB200_CommFree(slimb200_comm_t **comm) {
  if (!comm) return;
  comm_free(reinterpret_cast<Comm *>(*comm));
  *comm = nullptr;
}

slimb200_result_t *SLIMB200_AllGatherColumns(slimb200_comm_t *comm, const slimb200_result_t *local,
                                             const int32_t *positions, int32_t ncols_total, int32_t *r_status) {
  if (r_status) *r_status = SLIM_ERROR_INPUT;
  if (!comm || !local) return nullptr;
  return reinterpret_cast<slimb200_result_t *>(allgather_columns(
      reinterpret_cast<Comm *>(comm), reinterpret_cast<const Result *>(local), positions, ncols_total, r_status));
}

slim_t *SLIMB200_ResultToModel(const slimb200_result_t *result, int32_t *r_status) {
  if (r_status) *r_status = SLIM_ERROR_INPUT;
  if (!result) return nullptr;
  return model_from_result(reinterpret_cast<const Result *>(result), r_status);
}

slim_t *SLIMB200_AssembleModel(int32_t nitems, const int64_t *colptr, const int32_t *colind, const float *colval,
                               int32_t *r_status) {
  CsrHandle *h = (nitems >= 0 && colptr) ? assemble(nitems, colptr, colind, colval) : nullptr;
  if (r_status) *r_status = h ? SLIM_OK : SLIM_ERROR_MEMORY;
  return h;
}

}  // extern "C"
