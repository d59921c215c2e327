// gather.cuh -- assembling the model W on the device; included by engine.cu after `struct Result`.
//
//  * allgather_columns(): the multi-GPU exchange of SURVEY.md 8e.  Every rank solved a disjoint set of target
//    columns (reference src/libslim/estimate.c:402-403 is a plain parallel-for over columns); ONE grouped NCCL
//    collective (an all-gather with per-rank lengths: ncclBroadcast from every root inside one ncclGroup) moves
//    the (position, count) headers and the (row id, weight) payloads of all ranks over NVLink, and two small
//    kernels put every column segment at its final place of the CSC of SaveModel (estimate.c:570-588).  No
//    padding to the largest shard, no host copy of the shards, no host loop over columns.
//  * model_to_host(): the CSR view of W (gk_csr_CreateIndex(ROW), estimate.c:590 -> lib/GKlib/csr.c:1546-1584)
//    built on the GPU: histogram + scan of the row ids, stable LSD radix sort of (row id, entry position).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- in a torch process that is the copy torch already
// loaded), so libslim.so has no link-time dependency on it and single-GPU callers never touch it.
#pragma once

// (engine.cu includes <dlfcn.h> and <cub/device/device_scan.cuh> at file scope: this header sits inside namespace slimb200)

namespace nccl_shim {
// the handful of NCCL declarations this file needs (nccl.h 2.27 / 2.28: values are ABI-stable)
struct UniqueId {
  char internal[128];
};
typedef void *Comm;
enum { kSuccess = 0 };
enum { kInt8 = 0, kInt32 = 2, kInt64 = 4, kFloat32 = 7 };
struct Api {
  void *handle = nullptr;
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
  int (*CommInitAll)(Comm *, int, const int *) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
};

static Api *load() {
  static Api api;
  static bool tried = false;
  static std::string why;
  if (!tried) {
    tried = true;
    // 1. a copy that is already in the process (a torch process has its bundled NCCL loaded: use THAT one, two
    //    NCCL builds with the same soname in one process do not mix); 2. SLIMB200_NCCL_LIBRARY; 3. the system's
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char *names[] = {getenv("SLIMB200_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (api.handle) break;
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (!api.handle) why = dlerror();
    }
    if (api.handle) {
#define SLIM_NCCL_SYM(field, sym)                                                \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym));     \
  if (!api.field) {                                                              \
    why = std::string("missing symbol ") + sym;                                  \
    api.handle = nullptr;                                                        \
  }
      SLIM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
      SLIM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
      SLIM_NCCL_SYM(CommInitAll, "ncclCommInitAll")
      SLIM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
      SLIM_NCCL_SYM(GroupStart, "ncclGroupStart")
      SLIM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
      SLIM_NCCL_SYM(AllGather, "ncclAllGather")
      SLIM_NCCL_SYM(Broadcast, "ncclBroadcast")
      SLIM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
      SLIM_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef SLIM_NCCL_SYM
    }
  }
  if (!api.handle) throw EngineError(kErr, "NCCL is not available (dlopen libnccl.so.2): " + why);
  return &api;
}
}  // namespace nccl_shim

static inline void nck(int rc, const char *what) {
  if (rc != nccl_shim::kSuccess) {
    nccl_shim::Api *n = nccl_shim::load();
    throw EngineError(kErr, std::string(what) + ": " + n->GetErrorString(rc));
  }
}
#define NCK(x) nck((x), #x)

struct Comm {
  int device = 0;
  int nranks = 1, rank = 0;
  nccl_shim::Comm comm = nullptr;
  cudaStream_t stream = nullptr;
};

int comm_unique_id(void *id128) {
  try {
    nccl_shim::Api *n = nccl_shim::load();
    nccl_shim::UniqueId id;
    NCK(n->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

Comm *comm_init(int device, int nranks, int rank, const void *id128, int32_t *status) {
  Comm *c = nullptr;
  try {
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) throw EngineError(kErrInput, "comm_init: bad rank / size / id");
    if (device < 0 || device >= device_count()) throw EngineError(kErr, "comm_init: no usable CUDA device");
    nccl_shim::Api *n = nccl_shim::load();
    DeviceGuard guard(device);
    c = new Comm();
    c->device = device;
    c->nranks = nranks;
    c->rank = rank;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    nccl_shim::UniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCK(n->CommInitRank(&c->comm, nranks, id, rank));
    if (status) *status = kOk;
    return c;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    if (status) *status = e.status;
  }
  if (c) {
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
  }
  return nullptr;
}

// One communicator per device of `devices`, all in THIS process (the in-library multi-GPU path of SLIM_Learn).
int comm_init_all(int ndev, const int *devices, Comm **out) {
  std::vector<nccl_shim::Comm> cs((size_t)std::max(ndev, 1), nullptr);
  try {
    if (ndev < 1) throw EngineError(kErrInput, "comm_init_all: no devices");
    nccl_shim::Api *n = nccl_shim::load();
    NCK(n->CommInitAll(cs.data(), ndev, devices));
    for (int r = 0; r < ndev; r++) {
      DeviceGuard guard(devices[r]);
      Comm *c = new Comm();
      c->device = devices[r];
      c->nranks = ndev;
      c->rank = r;
      c->comm = cs[r];
      CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
      out[r] = c;
    }
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  }
}

void comm_free(Comm *c) {
  if (!c) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  if (c->comm) nccl_shim::load()->CommDestroy(c->comm);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
}

void comm_info(const Comm *c, int32_t *nranks, int32_t *rank, int32_t *device) {
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  if (device) *device = c->device;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// header entry h = (position of the column in the global list, nnz): counts_full[position] = nnz
__global__ void place_counts_kernel(int32_t n, const int32_t *__restrict__ head, int32_t ntotal, int32_t *counts_full,
                                    int32_t *bad) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int pos = head[2 * k], c = head[2 * k + 1];
    if (pos < 0 || pos >= ntotal || c < 0) atomicOr(bad, 1);
    else if (atomicExch(counts_full + pos, c) != -1) atomicOr(bad, 2);  // a position owned by two ranks
  }
}

// one warp per gathered column: staging (rank-major, segments in header order) -> final CSC position
__global__ void place_columns_kernel(int32_t n, const int32_t *__restrict__ head, const int64_t *__restrict__ src_off,
                                     const int64_t *__restrict__ colptr, const int32_t *__restrict__ st_ind,
                                     const float *__restrict__ st_val, int32_t *out_ind, float *out_val) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
    const int pos = head[2 * k], c = head[2 * k + 1];
    const int64_t s = src_off[k], d = colptr[pos];
    for (int e = lane; e < c; e += 32) {
      out_ind[d + e] = st_ind[s + e];
      out_val[d + e] = st_val[s + e];
    }
  }
}

__global__ void header_pack_kernel(int32_t n, const int32_t *__restrict__ positions, const int32_t *__restrict__ counts,
                                   int32_t *head) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    head[2 * k] = positions[k];
    head[2 * k + 1] = counts[k];
  }
}

__global__ void header_counts_kernel(int32_t n, const int32_t *__restrict__ head, int32_t *cnt) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) cnt[k] = head[2 * k + 1];
}

__global__ void fill_i32_kernel(int32_t n, int32_t v, int32_t *p) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) p[k] = v;
}

// entry position -> its column (expands colptr), and the row histogram of the CSC
__global__ void expand_columns_kernel(int32_t ncols, const int64_t *__restrict__ colptr, const int32_t *__restrict__ colind,
                                      int32_t *colof, int32_t *rowcnt) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < ncols; j += warps) {
    const int64_t a = colptr[j], b = colptr[j + 1];
    for (int64_t e = a + lane; e < b; e += 32) {
      colof[e] = j;
      atomicAdd(rowcnt + colind[e], 1);
    }
  }
}

__global__ void fill_rows_kernel(int64_t nnz, const uint32_t *__restrict__ spos, const int32_t *__restrict__ colof,
                                 const float *__restrict__ colval, int32_t *rowind, float *rowval) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = spos[k];
    rowind[k] = colof[p];
    rowval[k] = colval[p];
  }
}

__global__ void iota_u32_kernel(int64_t n, uint32_t *p) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    p[k] = (uint32_t)k;
}

// exclusive prefix sum with caller-provided scratch (scan_scratch_bytes): nothing is allocated or freed here
template <class In, class Out>
static size_t scan_scratch_bytes(int64_t n) {
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp, (const In *)nullptr, (Out *)nullptr, cub::Sum(), (Out)0, n));
  return tmp;
}
template <class In, class Out>
static void exclusive_sum(const In *d_in, Out *d_out, int64_t n, void *scratch, size_t scratch_bytes, cudaStream_t s) {
  CK(cub::DeviceScan::ExclusiveScan(scratch, scratch_bytes, d_in, d_out, cub::Sum(), (Out)0, n, s));
}

// ------------------------------------------------------------------------------------------------
// allgather_columns
// ------------------------------------------------------------------------------------------------
// `local`: this rank's solved columns; positions[k] = index of local column k in the global column list of
// `ncols_total` entries (every position owned by exactly one rank).  Returns a Result holding the CSC of ALL
// ncols_total columns on this rank's device (identical on every rank); per-column statistics are not exchanged.
Result *allgather_columns(Comm *c, const Result *local, const int32_t *positions, int32_t ncols_total, int32_t *status) {
  Result *res = nullptr;
  try {
    if (!c || !local || ncols_total < 0 || (local->nsel > 0 && !positions))
      throw EngineError(kErrInput, "allgather_columns: bad arguments");
    if (local->device != c->device) throw EngineError(kErrInput, "allgather_columns: result and communicator live on different devices");
    nccl_shim::Api *n = nccl_shim::load();
    DeviceGuard guard(c->device);
    cudaStream_t s = c->stream;
    const int W = c->nranks;
    const int32_t nloc = local->nsel;
    const int64_t nnz_loc = local->nnz;
    EventPair ev;
    CK(cudaEventRecord(ev.a, s));

    // (1) sizes of every rank: 2 x int64 per rank
    DevBuf<int64_t> d_meta, d_metas;
    d_meta.alloc(2);
    d_metas.alloc((size_t)2 * W);
    const int64_t meta[2] = {nloc, nnz_loc};
    CK(cudaMemcpyAsync(d_meta.p, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    NCK(n->AllGather(d_meta.p, d_metas.p, 2, nccl_shim::kInt64, c->comm, s));
    std::vector<int64_t> metas((size_t)2 * W);
    CK(cudaMemcpyAsync(metas.data(), d_metas.p, sizeof(int64_t) * 2 * W, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int64_t> hoff(W + 1, 0), poff(W + 1, 0);
    for (int r = 0; r < W; r++) {
      hoff[r + 1] = hoff[r] + metas[2 * r];
      poff[r + 1] = poff[r] + metas[2 * r + 1];
    }
    const int64_t ngath = hoff[W], nnz = poff[W];
    if (ngath != ncols_total) throw EngineError(kErrInput, "allgather_columns: the ranks' column counts do not add up to ncols_total");

    // (2) local header (position, count) pairs
    DevBuf<int32_t> d_pos, d_head, d_heads;
    d_pos.alloc(nloc);
    d_head.alloc((size_t)2 * nloc);
    d_heads.alloc((size_t)2 * ngath);
    if (nloc > 0) {
      CK(cudaMemcpyAsync(d_pos.p, positions, sizeof(int32_t) * nloc, cudaMemcpyHostToDevice, s));
      header_pack_kernel<<<grid_for(nloc, 256, 148), 256, 0, s>>>(nloc, d_pos.p, local->d_counts, d_head.p);
      CK(cudaGetLastError());
    }
    // (3) ONE grouped collective: headers and payloads of every rank, each at its offset (all-gather-v)
    DevBuf<int32_t> st_ind;
    DevBuf<float> st_val;
    st_ind.alloc(nnz);
    st_val.alloc(nnz);
    // every allocation happens BEFORE the collective is enqueued: with several ranks in one process a
    // cudaMalloc / cudaFree issued while a peer still has to launch its half can deadlock against NCCL
    res = new Result();
    res->device = c->device;
    res->nsel = ncols_total;
    res->nnz = nnz;
    res->h_colptr.assign((size_t)ncols_total + 1, 0);
    res->niters.assign(ncols_total, 0);
    res->nactive.assign(ncols_total, 0);
    res->actnnz.assign(ncols_total, 0);
    res->expand.assign(ncols_total, 0);
    res->rnorm.assign(ncols_total, 0.0);
    res->obj.assign(ncols_total, 0.0);
    res->phase.assign((size_t)ncols_total * 4, 0.f);
    res->ngroups.assign(ncols_total, 0);
    CK(cudaMalloc(&res->d_colptr, sizeof(int64_t) * ((size_t)ncols_total + 1)));
    CK(cudaMalloc(&res->d_counts, sizeof(int32_t) * ((size_t)ncols_total + 1)));
    CK(cudaMalloc(&res->d_colind, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    CK(cudaMalloc(&res->d_colval, sizeof(float) * std::max<int64_t>(nnz, 1)));
    DevBuf<int32_t> d_bad, d_scnt;
    DevBuf<int64_t> d_soff;
    d_bad.alloc_zero(1, s);
    d_scnt.alloc((size_t)ngath + 1);
    d_soff.alloc((size_t)ngath + 1);
    const size_t scan_bytes = scan_scratch_bytes<int32_t, int64_t>(std::max<int64_t>(ncols_total, ngath) + 1);
    DevBuf<unsigned char> d_scan;
    d_scan.alloc(scan_bytes);
    NCK(n->GroupStart());
    for (int r = 0; r < W; r++) {
      const size_t hc = (size_t)2 * metas[2 * r], pc = (size_t)metas[2 * r + 1];
      if (hc) NCK(n->Broadcast(d_head.p, d_heads.p + 2 * hoff[r], hc, nccl_shim::kInt32, r, c->comm, s));
      if (pc) {
        NCK(n->Broadcast(local->d_colind, st_ind.p + poff[r], pc, nccl_shim::kInt32, r, c->comm, s));
        NCK(n->Broadcast(local->d_colval, st_val.p + poff[r], pc, nccl_shim::kFloat32, r, c->comm, s));
      }
    }
    NCK(n->GroupEnd());

    // (4) final CSC: counts by position -> colptr; source offsets = scan of the counts in staging order
    fill_i32_kernel<<<grid_for(ncols_total + 1, 256, 148), 256, 0, s>>>(ncols_total + 1, -1, res->d_counts);
    if (ngath > 0) {
      place_counts_kernel<<<grid_for(ngath, 256, 148), 256, 0, s>>>((int32_t)ngath, d_heads.p, ncols_total, res->d_counts,
                                                                     d_bad.p);
      header_counts_kernel<<<grid_for(ngath, 256, 148), 256, 0, s>>>((int32_t)ngath, d_heads.p, d_scnt.p);
    }
    CK(cudaGetLastError());
    int32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemsetAsync(res->d_counts + ncols_total, 0, sizeof(int32_t), s));
    CK(cudaMemsetAsync(d_scnt.p + ngath, 0, sizeof(int32_t), s));
    exclusive_sum<int32_t, int64_t>(res->d_counts, res->d_colptr, (int64_t)ncols_total + 1, d_scan.p, scan_bytes, s);
    exclusive_sum<int32_t, int64_t>(d_scnt.p, d_soff.p, ngath + 1, d_scan.p, scan_bytes, s);
    CK(cudaStreamSynchronize(s));
    if (bad) throw EngineError(kErrInput, "allgather_columns: column positions out of range or owned by two ranks");
    if (ngath > 0) {
      place_columns_kernel<<<grid_for(ngath * 32, 256, 148), 256, 0, s>>>((int32_t)ngath, d_heads.p, d_soff.p, res->d_colptr,
                                                                          st_ind.p, st_val.p, res->d_colind, res->d_colval);
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(res->h_colptr.data(), res->d_colptr, sizeof(int64_t) * ((size_t)ncols_total + 1),
                       cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ev.b, s));
    CK(cudaStreamSynchronize(s));
    if (res->h_colptr[ncols_total] != nnz) throw EngineError(kErr, "allgather_columns: a column position is missing");
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
    res->tm.gather_ms = ms;
    res->tm.launches = ngath > 0 ? 6 : 3;
    if (status) *status = kOk;
    return res;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    if (status) *status = e.status;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    if (status) *status = kErrMemory;
  }
  free_result(res);
  return nullptr;
}

// ------------------------------------------------------------------------------------------------
// model_to_host: both views of SaveModel (estimate.c:570-593) from a Result that holds ALL columns
// ------------------------------------------------------------------------------------------------
// colptr/rowptr: ssize_t[n+1]; colind/colval/rowind/rowval: [nnz] (caller allocated, e.g. malloc for a model handle)
int model_to_host(const Result *r, ssize_t *colptr, int32_t *colind, float *colval, ssize_t *rowptr, int32_t *rowind,
                  float *rowval, double *index_ms) {
  try {
    DeviceGuard guard(r->device);
    const int32_t n = r->nsel;
    const int64_t nnz = r->nnz;
    if (nnz >= (int64_t)0xffffffffLL) throw EngineError(kErrInput, "model_to_host: nnz(W) must be below 2^32");
    cudaStream_t s = nullptr;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct StreamOwner {
      cudaStream_t s;
      ~StreamOwner() { cudaStreamDestroy(s); }
    } owner{s};
    for (int32_t j = 0; j <= n; j++) colptr[j] = (ssize_t)r->h_colptr[j];
    EventPair ev;
    CK(cudaEventRecord(ev.a, s));
    DevBuf<int32_t> d_rowcnt, d_colof, d_srow, d_rowind;
    DevBuf<int64_t> d_rowptr;
    DevBuf<uint32_t> d_pos, d_spos;
    DevBuf<float> d_rowval;
    d_rowcnt.alloc_zero((size_t)n + 1, s);
    d_rowptr.alloc((size_t)n + 1);
    d_colof.alloc(nnz);
    d_srow.alloc(nnz);
    d_pos.alloc(nnz);
    d_spos.alloc(nnz);
    d_rowind.alloc(nnz);
    d_rowval.alloc(nnz);
    if (n > 0 && nnz > 0) {
      expand_columns_kernel<<<grid_for((int64_t)n * 32, 256, 148), 256, 0, s>>>(n, r->d_colptr, r->d_colind, d_colof.p,
                                                                                d_rowcnt.p);
      iota_u32_kernel<<<grid_for(nnz, 256, 148), 256, 0, s>>>(nnz, d_pos.p);
      CK(cudaGetLastError());
    }
    const size_t scan_bytes = scan_scratch_bytes<int32_t, int64_t>((int64_t)n + 1);
    DevBuf<unsigned char> d_scan;
    d_scan.alloc(scan_bytes);
    exclusive_sum<int32_t, int64_t>(d_rowcnt.p, d_rowptr.p, (int64_t)n + 1, d_scan.p, scan_bytes, s);
    if (nnz > 0) {
      int bits = 1;
      while ((1LL << bits) < (int64_t)n) bits++;
      size_t tmp_bytes = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, r->d_colind, d_srow.p, d_pos.p, d_spos.p, nnz, 0, bits, s));
      DevBuf<unsigned char> d_tmp;
      d_tmp.alloc(tmp_bytes);
      CK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, r->d_colind, d_srow.p, d_pos.p, d_spos.p, nnz, 0, bits, s));
      fill_rows_kernel<<<grid_for(nnz, 256, 148), 256, 0, s>>>(nnz, d_spos.p, d_colof.p, r->d_colval, d_rowind.p,
                                                               d_rowval.p);
      CK(cudaGetLastError());
      CK(cudaEventRecord(ev.b, s));
      CK(cudaMemcpyAsync(colind, r->d_colind, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(colval, r->d_colval, sizeof(float) * nnz, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(rowind, d_rowind.p, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, s));
      CK(cudaMemcpyAsync(rowval, d_rowval.p, sizeof(float) * nnz, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
    } else {
      CK(cudaEventRecord(ev.b, s));
    }
    std::vector<int64_t> rp((size_t)n + 1);
    CK(cudaMemcpyAsync(rp.data(), d_rowptr.p, sizeof(int64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int32_t i = 0; i <= n; i++) rowptr[i] = (ssize_t)rp[i];
    if (index_ms) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
      *index_ms = ms;
    }
    return kOk;
  } catch (const EngineError &e) {
    g_last_error = e.what();
    return e.status;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return kErrMemory;
  }
}
