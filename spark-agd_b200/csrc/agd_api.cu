// agd_api.cu -- the C-ABI of include/agd_b200.h: handle, shard loading, the collective, and the
// AcceleratedGradientDescent.run driver loop (AGD.scala:177-338) executed natively around the
// K1 / all-reduce / K3 kernels.  No CPU fallback: every compute entry point needs an sm_100 GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; the library is dlopen'ed (torch ships its own libnccl.so.2)
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "agd_common.cuh"

using namespace agd;

// ---------------------------------------------------------------- NCCL through dlopen
namespace {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
  std::string why;
};

NcclApi &nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define AGD_NCCL_SYM(name)                                                   \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(lib, "nccl" #name)); \
  if (!api.name) { api.why = "missing symbol nccl" #name; return; }
    AGD_NCCL_SYM(GetUniqueId)
    AGD_NCCL_SYM(CommInitRank)
    AGD_NCCL_SYM(AllReduce)
    AGD_NCCL_SYM(AllGather)
    AGD_NCCL_SYM(GroupStart)
    AGD_NCCL_SYM(GroupEnd)
    AGD_NCCL_SYM(CommDestroy)
    AGD_NCCL_SYM(GetErrorString)
#undef AGD_NCCL_SYM
    api.ok = true;
  });
  return api;
}

std::string g_create_error;
std::mutex g_create_mu;

struct Shard {
  void *X = nullptr;          // dense, row-major, ld == d
  double *labels = nullptr;   // cap + pad
  int64_t rows = 0, cap = 0;
  int elem_bytes = 0;         // 4 (fp32) or 8 (fp64) storage
  bool csr = false;
  int64_t *rowptr = nullptr;
  int32_t *idx = nullptr;
  void *val = nullptr;
  int64_t nnz = 0, nnz_cap = 0;
};

struct Dev {
  int ordinal = -1;
  int sm_count = 0;
  cudaStream_t st = nullptr;
  Shard sh;
  // d-vectors of the driver loop (AGD.scala:224-230,241,249) + packed pass result
  double *x = nullptr, *z = nullptr, *x_old = nullptr, *z_old = nullptr, *y = nullptr, *g_y = nullptr,
         *g_x = nullptr, *wtmp = nullptr, *acc = nullptr, *y_spec = nullptr;
  int32_t vec_d = 0;
  double *slabs = nullptr;
  size_t slabs_doubles = 0;
  double *partials = nullptr;
  unsigned int *ticket = nullptr;
  double *scalars_dev = nullptr;   // device alias of scalars_host: K3 writes its scalars straight to the host
  double *scalars_host = nullptr;  // pinned + mapped, 2*K3_NS
  void *stage_dev = nullptr;
  size_t stage_bytes = 0;
  ncclComm_t comm = nullptr;
  std::vector<cudaEvent_t> ev;     // K1 start/stop pairs (device 0 only)
  size_t ev_used = 0;
  std::vector<cudaEvent_t> ev_ar;  // all-reduce start/stop pairs
  size_t ev_ar_used = 0;
  std::mutex *mu = nullptr;
  long long row_base = 0;          // global index of this shard's first row (for the sampling mask)
  double *hist_host = nullptr;            // pinned + mapped: [2k] = loss sum, [2k+1] = count of the history pass of iteration k
  double *hist_dev = nullptr;             // device alias of hist_host (k3_step stores the pair that rode along with a fused sweep)
  size_t hist_cap = 0;
  // K2' peer-memory exchange (xchg.cu)
  double *xbuf = nullptr;                 // [2][W][d+4], written by every rank over NVLink
  unsigned long long *xflags = nullptr;   // [2][W] epochs
  unsigned int *xticket = nullptr;
  XchgPeers xpeers;                       // every rank's xbuf / xflags as mapped into this device
  std::vector<void *> xopened;            // cudaIpcOpenMemHandle mappings to close
};

}  // namespace

struct agd_handle {
  std::vector<Dev> devs;
  int32_t d = 0;       // INTERNAL dimension: d_user padded with zero columns so that rows are whole 16-byte vectors
  int32_t d_user = 0;  // the caller's feature count (what agd_dim reports)
  int world = 1, first_rank = 0;
  bool comm_ready = false;
  bool comm_auto = false;   // the world is this process's own GPUs (agd_create): NCCL is only built if it is ever needed
  bool ipc_only = false;    // agd_comm_init_ipc: no NCCL at all, the host ships the CUDA IPC handles (agd_xchg_export/import)
  int k1_variant = 0;  // 0 auto, 1 ring, 2 generic, 4 tcgen05 (bf16)
  int ring_stages = 0;
  int tune_rows = 0, tune_ctas = 0, tune_full = 0;
  int k1_diag = 0;
  int tc_margins_f64 = 0;    // tcgen05 kernel: fp64-exact margins instead of the fp32 phase 1
  unsigned long long sample_seed = 0, sample_thresh = 0;  // mini-batch row mask of the current pass (0 = every row)
  int collective = 0;        // 0 = auto (peer memory if every pair of ranks can map each other, else NCCL), 1 = nccl, 2 = p2p
  int32_t x_d = 0;           // dimension the exchange buffers were built for (0 = not built)
  bool x_p2p = false;        // exchange buffers are live
  unsigned long long x_epoch = 0;
  // a sweep whose gather was left to the K3 kernel that consumes it (smooth_device(..., defer_gather))
  bool xg_pending = false;
  bool xg_rs = false;        // the pending exchange is of the reduce-scatter form
  unsigned long long xg_epoch = 0;
  int xg_n = 0;
  std::string err;
  std::mutex mu;
  unsigned long long seq_base = 0;   // last round sequence number handed out (wait_scalars)
  // AGD_TRACE=1 (diagnostics): an event after every launch on device 0; agd_run prints per-kernel totals (gap + run time) to stderr
  int trace = -1;
  std::vector<std::pair<const char *, cudaEvent_t>> tr;
  std::vector<cudaEvent_t> tr_pool;
  int64_t launches = 0;  // per device, current call
  int64_t collectives = 0;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
};

namespace {

int fail(agd_handle *h, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else {
    std::lock_guard<std::mutex> g(g_create_mu);
    g_create_error = buf;
  }
  return 1;
}

#define CK(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKN(call)                                                                                        \
  do {                                                                                                   \
    ncclResult_t r_ = (call);                                                                            \
    if (r_ != ncclSuccess) return fail(h, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

// Java Math.max / Math.min (NaN-propagating), used at AGD.scala:274,275,286,292,322
double jmax(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a == 0.0 && b == 0.0) return std::signbit(a) ? b : a;
  return a > b ? a : b;
}
double jmin(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
  return a < b ? a : b;
}

int dtype_bytes(int dt) { return dt == AGD_F64 ? 8 : (dt == AGD_F32 ? 4 : (dt == AGD_BF16 ? 2 : 0)); }
int bytes_dtype(int eb) { return eb == 8 ? AGD_F64 : (eb == 4 ? AGD_F32 : AGD_BF16); }

int free_shard(agd_handle *h, Dev &D) {
  CK(cudaSetDevice(D.ordinal));
  Shard &s = D.sh;
  if (s.X) cudaFree(s.X);
  if (s.labels) cudaFree(s.labels);
  if (s.rowptr) cudaFree(s.rowptr);
  if (s.idx) cudaFree(s.idx);
  if (s.val) cudaFree(s.val);
  s = Shard();
  return 0;
}

int ensure_vectors(agd_handle *h, Dev &D, int32_t d) {
  if (D.vec_d == d) return 0;
  CK(cudaSetDevice(D.ordinal));
  double **v[] = {&D.x, &D.z, &D.x_old, &D.z_old, &D.y, &D.g_y, &D.g_x, &D.wtmp, &D.y_spec};
  for (double **p : v) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    CK(cudaMalloc(p, ((size_t)d + 4) * sizeof(double)));
    CK(cudaMemsetAsync(*p, 0, ((size_t)d + 4) * sizeof(double), D.st));
  }
  if (D.acc) cudaFree(D.acc);
  // [grad(d) | loss | count | loss at w2 | count at w2], and after a two-gradient sweep a second block [grad at w2 | loss | count | 0 | 0]
  CK(cudaMalloc(&D.acc, 2 * ((size_t)d + 4) * sizeof(double)));
  if (D.partials) cudaFree(D.partials);
  CK(cudaMalloc(&D.partials, (size_t)k3_blocks(d) * K3_NS * sizeof(double)));
  D.vec_d = d;
  return 0;
}

int ensure_slabs(agd_handle *h, Dev &D, int blocks, int32_t n) {
  const size_t need = (size_t)blocks * (size_t)n;
  if (need <= D.slabs_doubles) return 0;
  CK(cudaSetDevice(D.ordinal));
  if (D.slabs) cudaFree(D.slabs);
  D.slabs = nullptr;
  CK(cudaMalloc(&D.slabs, need * sizeof(double)));
  D.slabs_doubles = need;
  return 0;
}

// elem_bytes > 0: dense storage -- rows are padded with zero columns to whole 16-byte vectors when that puts the
// shard on the TMA-ring / tcgen05 kernels.  The solver then runs in the padded dimension, which is bit-identical:
// the extra gradient entries are exactly 0, so the extra weights stay exactly 0 under every updater.
int set_dim(agd_handle *h, int32_t d, int elem_bytes = 0) {
  std::lock_guard<std::mutex> g(h->mu);
  if (d <= 0) return fail(h, "feature dimension must be positive (got %d)", d);
  if (h->d_user == 0) {
    h->d_user = d;
    h->d = d;
    if (elem_bytes > 0) {
      const int epv = 16 / elem_bytes;
      const int padded = (d + epv - 1) / epv * epv;
      if (padded != d && k1_ring_supported(padded, elem_bytes)) h->d = padded;
    }
  } else if (h->d_user != d) {
    return fail(h, "feature dimension mismatch: shard has d=%d, call passed d=%d", h->d_user, d);
  }
  return 0;
}

int reserve_locked(agd_handle *h, Dev &D, int64_t cap, int32_t d, int store_dtype) {
  const int eb = dtype_bytes(store_dtype);
  if (!eb) return fail(h, "store_dtype must be AGD_F64, AGD_F32 or AGD_BF16");
  if (cap < 0) return fail(h, "negative capacity");
  CK(cudaSetDevice(D.ordinal));
  Shard &s = D.sh;
  if (s.csr) return fail(h, "device already holds a CSR shard");
  if (s.cap > 0 && s.elem_bytes != eb) return fail(h, "storage dtype mismatch with the resident shard");
  if (cap <= s.cap) {
    if (!s.elem_bytes) s.elem_bytes = eb;  // an empty shard still records its storage type
    return 0;
  }
  void *nx = nullptr;
  double *nl = nullptr;
  const size_t xbytes = (size_t)cap * d * eb + 64;
  CK(cudaMalloc(&nx, xbytes));
  CK(cudaMalloc(&nl, ((size_t)cap + 64) * sizeof(double)));
  CK(cudaMemsetAsync(nl, 0, ((size_t)cap + 64) * sizeof(double), D.st));
  if (s.rows > 0) {
    CK(cudaMemcpyAsync(nx, s.X, (size_t)s.rows * d * eb, cudaMemcpyDeviceToDevice, D.st));
    CK(cudaMemcpyAsync(nl, s.labels, (size_t)s.rows * sizeof(double), cudaMemcpyDeviceToDevice, D.st));
  }
  CK(cudaStreamSynchronize(D.st));
  if (s.X) cudaFree(s.X);
  if (s.labels) cudaFree(s.labels);
  s.X = nx;
  s.labels = nl;
  s.cap = cap;
  s.elem_bytes = eb;
  return 0;
}

int ensure_stage(agd_handle *h, Dev &D, size_t bytes) {
  if (D.stage_bytes >= bytes) return 0;
  if (D.stage_dev) cudaFree(D.stage_dev);
  D.stage_dev = nullptr;
  CK(cudaMalloc(&D.stage_dev, bytes));
  D.stage_bytes = bytes;
  return 0;
}

cudaEvent_t next_event(std::vector<cudaEvent_t> &pool, size_t &used) {
  if (used == pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    pool.push_back(e);
  }
  return pool[used++];
}


void free_xchg(agd_handle *h) {
  for (Dev &D : h->devs) {
    cudaSetDevice(D.ordinal);
    for (void *p : D.xopened) cudaIpcCloseMemHandle(p);
    D.xopened.clear();
    if (D.xbuf) cudaFree(D.xbuf);
    if (D.xflags) cudaFree(D.xflags);
    if (D.xticket) cudaFree(D.xticket);
    D.xbuf = nullptr; D.xflags = nullptr; D.xticket = nullptr;
  }
  h->x_d = 0;
  h->x_p2p = false;
}

// What one rank tells the others about its exchange buffers (AGD_XCHG_HANDLE_BYTES on the wire).
struct XHandles {
  cudaIpcMemHandle_t buf, flags;
  int32_t can_peer, d, rank, pad;
  unsigned char reserve[AGD_XCHG_HANDLE_BYTES - 2 * sizeof(cudaIpcMemHandle_t) - 16];
};
static_assert(sizeof(XHandles) == AGD_XCHG_HANDLE_BYTES, "exchange handle blob size is part of the ABI");

void destroy_comms(agd_handle *h) {
  for (Dev &D : h->devs) {
    if (D.comm && nccl_api().ok) { cudaSetDevice(D.ordinal); nccl_api().CommDestroy(D.comm); }
    D.comm = nullptr;
  }
}

// NCCL communicator of a single-process world, built on first need (fallback / collective=nccl)
int ensure_nccl(agd_handle *h) {
  if (h->world <= 1) return 0;
  bool have = true;
  for (Dev &D : h->devs) have = have && D.comm != nullptr;
  if (have) return 0;
  if (h->ipc_only) return fail(h, "this world was set up with agd_comm_init_ipc: there is no NCCL communicator to fall back to");
  if (!h->comm_auto) return fail(h, "world_ranks=%d but agd_comm_init was not called", h->world);
  NcclApi &N = nccl_api();
  if (!N.ok) return fail(h, "NCCL unavailable: %s", N.why.c_str());
  ncclUniqueId id;
  CKN(N.GetUniqueId(&id));
  const int nd = (int)h->devs.size();
  CKN(N.GroupStart());
  for (int i = 0; i < nd; ++i) {
    CK(cudaSetDevice(h->devs[i].ordinal));
    CKN(N.CommInitRank(&h->devs[i].comm, nd, id, i));
  }
  CKN(N.GroupEnd());
  return 0;
}

// step 1 of the exchange setup: allocate this process's buffers for the current dimension and describe them
int xchg_alloc(agd_handle *h, std::vector<XHandles> &mine) {
  const int W = h->world, nd = (int)h->devs.size();
  const int S = 2 * (h->d + 4);              // slot stride: room for a two-gradient sweep
  const size_t total = xchg_total_doubles(S, W), nflags = 6 * (size_t)W;   // one-shot + reduce-scatter areas (agd_common.cuh)
  mine.assign((size_t)nd, XHandles());
  for (int i = 0; i < nd; ++i) {
    Dev &D = h->devs[i];
    CK(cudaSetDevice(D.ordinal));
    CK(cudaMalloc(&D.xbuf, total * sizeof(double)));
    CK(cudaMalloc(&D.xflags, nflags * sizeof(unsigned long long)));
    CK(cudaMalloc(&D.xticket, sizeof(unsigned int)));
    CK(cudaMemset(D.xbuf, 0, total * sizeof(double)));
    CK(cudaMemset(D.xflags, 0, nflags * sizeof(unsigned long long)));
    CK(cudaMemset(D.xticket, 0, sizeof(unsigned int)));
    memset(&mine[i], 0, sizeof(XHandles));
    mine[i].can_peer = 1;
    mine[i].d = h->d;
    mine[i].rank = h->first_rank + i;
    for (int j = 0; j < nd; ++j) {
      if (j == i) continue;
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, D.ordinal, h->devs[j].ordinal));
      if (!can) mine[i].can_peer = 0;
      else { cudaError_t e = cudaDeviceEnablePeerAccess(h->devs[j].ordinal, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) mine[i].can_peer = 0; cudaGetLastError(); }
    }
    if (nd < W) {  // buffers of other processes are reached through CUDA IPC
      if (cudaIpcGetMemHandle(&mine[i].buf, D.xbuf) != cudaSuccess || cudaIpcGetMemHandle(&mine[i].flags, D.xflags) != cudaSuccess) {
        mine[i].can_peer = 0;
        cudaGetLastError();
      }
    }
  }
  return 0;
}

// step 3: map every rank's buffers into every local device; *ok_out = false when some pair cannot be mapped
int xchg_map(agd_handle *h, const std::vector<XHandles> &all, bool *ok_out) {
  const int W = h->world, nd = (int)h->devs.size();
  bool ok = true;
  for (int r = 0; r < W; ++r) ok = ok && all[(size_t)r].can_peer && all[(size_t)r].d == h->d && all[(size_t)r].rank == r;
  for (int i = 0; i < nd && ok; ++i) {
    Dev &D = h->devs[i];
    CK(cudaSetDevice(D.ordinal));
    for (int r = 0; r < W; ++r) {
      const int lj = r - h->first_rank;
      if (lj >= 0 && lj < nd) {  // same process: direct pointers (peer access enabled above)
        D.xpeers.slot[r] = h->devs[lj].xbuf;
        D.xpeers.flag[r] = h->devs[lj].xflags;
        continue;
      }
      void *pb = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&pb, all[(size_t)r].buf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&pf, all[(size_t)r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      D.xopened.push_back(pb);
      D.xopened.push_back(pf);
      D.xpeers.slot[r] = (double *)pb;
      D.xpeers.flag[r] = (unsigned long long *)pf;
    }
  }
  *ok_out = ok;
  return 0;
}

// Builds the peer-memory exchange for the current dimension.  Three transports for the setup (never for the data):
//  * every rank is a GPU of this process: nothing travels (direct peer pointers), NCCL is not touched;
//  * agd_comm_init worlds: the CUDA IPC handles travel once through an NCCL all-gather, the yes/no decision through a
//    one-int all-reduce; if some pair of ranks cannot map each other NCCL carries the all-reduce instead;
//  * agd_comm_init_ipc worlds: the host ships the handles (agd_xchg_export / agd_xchg_import) -- no NCCL anywhere.
// Collective: every rank calls it from the same entry point.
int ensure_xchg(agd_handle *h) {
  if (h->world <= 1) return 0;
  if (h->collective == 1) return ensure_nccl(h);
  if (h->x_d == h->d) return h->x_p2p ? 0 : ensure_nccl(h);
  if (h->ipc_only)
    return fail(h, "peer-memory exchange not established for d=%d: call agd_xchg_export / agd_xchg_import after loading the shards", h->d);
  free_xchg(h);
  h->x_d = h->d;
  const int W = h->world, nd = (int)h->devs.size();
  if (W > kMaxRanks) { if (h->collective == 2) return fail(h, "peer exchange supports at most %d ranks", kMaxRanks); return ensure_nccl(h); }
  std::vector<XHandles> mine, all((size_t)W);
  if (xchg_alloc(h, mine)) return 1;
  bool ok = true;
  if (nd == W) {
    all = mine;
    if (xchg_map(h, all, &ok)) return 1;
  } else {
    if (ensure_nccl(h)) return 1;
    NcclApi &N = nccl_api();
    // all-gather the handles (device staging through NCCL; setup only)
    std::vector<void *> stage(nd);
    for (int i = 0; i < nd; ++i) {
      Dev &D = h->devs[i];
      CK(cudaSetDevice(D.ordinal));
      CK(cudaMalloc(&stage[i], (size_t)W * sizeof(XHandles)));
      CK(cudaMemcpyAsync((char *)stage[i] + (size_t)(h->first_rank + i) * sizeof(XHandles), &mine[i], sizeof(XHandles), cudaMemcpyHostToDevice, D.st));
    }
    CKN(N.GroupStart());
    for (int i = 0; i < nd; ++i) {
      Dev &D = h->devs[i];
      CKN(N.AllGather((char *)stage[i] + (size_t)(h->first_rank + i) * sizeof(XHandles), stage[i], sizeof(XHandles), ncclChar, D.comm, D.st));
    }
    CKN(N.GroupEnd());
    CK(cudaSetDevice(h->devs[0].ordinal));
    CK(cudaMemcpyAsync(all.data(), stage[0], (size_t)W * sizeof(XHandles), cudaMemcpyDeviceToHost, h->devs[0].st));
    for (int i = 0; i < nd; ++i) { CK(cudaSetDevice(h->devs[i].ordinal)); CK(cudaStreamSynchronize(h->devs[i].st)); }
    for (int i = 0; i < nd; ++i) { cudaSetDevice(h->devs[i].ordinal); cudaFree(stage[i]); }
    if (xchg_map(h, all, &ok)) return 1;
    // every rank must take the same decision: agree through one more (tiny) all-reduce of the ok flag
    int v = ok ? 1 : 0;
    std::vector<int *> fl(nd);
    for (int i = 0; i < nd; ++i) {
      CK(cudaSetDevice(h->devs[i].ordinal));
      CK(cudaMalloc(&fl[i], sizeof(int)));
      CK(cudaMemcpyAsync(fl[i], &v, sizeof(int), cudaMemcpyHostToDevice, h->devs[i].st));
    }
    CKN(N.GroupStart());
    for (int i = 0; i < nd; ++i) CKN(N.AllReduce(fl[i], fl[i], 1, ncclInt, ncclMin, h->devs[i].comm, h->devs[i].st));
    CKN(N.GroupEnd());
    CK(cudaSetDevice(h->devs[0].ordinal));
    CK(cudaMemcpyAsync(&v, fl[0], sizeof(int), cudaMemcpyDeviceToHost, h->devs[0].st));
    for (int i = 0; i < nd; ++i) { CK(cudaSetDevice(h->devs[i].ordinal)); CK(cudaStreamSynchronize(h->devs[i].st)); cudaFree(fl[i]); }
    ok = v == 1;
  }
  if (!ok) {
    if (h->collective == 2) return fail(h, "peer-memory exchange unavailable (no P2P / IPC mapping between every pair of ranks)");
    const int32_t keep = h->x_d;
    free_xchg(h);
    h->x_d = keep;  // do not retry every pass; NCCL carries the all-reduce
    return ensure_nccl(h);
  }
  h->x_p2p = true;
  h->x_epoch = 0;
  return 0;
}

void trace_mark(agd_handle *h, const char *tag) {
  if (h->trace < 0) { const char *e = getenv("AGD_TRACE"); h->trace = (e && *e && *e != '0') ? 1 : 0; }
  if (!h->trace) return;
  Dev &D = h->devs[0];
  cudaSetDevice(D.ordinal);
  cudaEvent_t ev;
  if (h->tr.size() < h->tr_pool.size()) ev = h->tr_pool[h->tr.size()];
  else { cudaEventCreate(&ev); h->tr_pool.push_back(ev); }
  cudaEventRecord(ev, D.st);
  h->tr.push_back({tag, ev});
}

void trace_report(agd_handle *h, const char *what) {
  if (h->trace != 1 || h->tr.size() < 2) { h->tr.clear(); return; }
  struct Acc { double ms = 0; int n = 0; };
  std::vector<std::pair<std::string, Acc>> sums;
  double total = 0;
  for (size_t i = 1; i < h->tr.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->tr[i - 1].second, h->tr[i].second);
    total += ms;
    bool found = false;
    for (auto &kv : sums) if (kv.first == h->tr[i].first) { kv.second.ms += ms; kv.second.n++; found = true; break; }
    if (!found) { Acc a; a.ms = ms; a.n = 1; sums.push_back({h->tr[i].first, a}); }
  }
  fprintf(stderr, "[AGD_TRACE %s rank %d] total %.3f ms:", what, h->first_rank, total);
  for (auto &kv : sums) fprintf(stderr, " %s %.3f ms / %d = %.1f us;", kv.first.c_str(), kv.second.ms, kv.second.n, kv.second.ms / kv.second.n * 1e3);
  fprintf(stderr, "\n");
  h->tr.clear();
}

typedef const double *(*WSel)(Dev &);

// which K1 kernel a dense shard of this handle runs on (0 generic, 1 ring, 3 tcgen05)
int dense_kernel_of(const agd_handle *h, int eb) {
  bool ring = k1_ring_supported(h->d, eb) != 0;
  if (h->k1_variant == 2) ring = false;
  if (k1_tc_supported(h->d, eb) && (h->k1_variant == 0 || h->k1_variant == 4)) return 3;
  return ring ? 1 : 0;
}

// can one sweep evaluate the loss at a second point as well (pass fusion)?
bool dual_supported(const agd_handle *h) {
  for (const Dev &D : h->devs) {
    const Shard &s = D.sh;
    if (s.csr) continue;
    const int eb = s.elem_bytes ? s.elem_bytes : 4;
    const int k = dense_kernel_of(h, eb);
    if (k == 3 && (h->tune_rows != 0 || h->tc_margins_f64)) return false;   // tcgen05: the default (fp32-margin) mapping has one
    if (k == 1 && !k1_ring_dual_supported(h->d, eb)) return false;
  }
  return h->k1_diag == 0;
}

// ... and the gradient there too (the speculative sweep of the memoised pass structure)?
bool dual_full_supported(const agd_handle *h) {
  for (const Dev &D : h->devs) {
    const Shard &s = D.sh;
    if (s.csr) return false;
    const int eb = s.elem_bytes ? s.elem_bytes : 4;
    const int k = dense_kernel_of(h, eb);
    if (k == 3) { if (h->tune_rows != 0 || h->tc_margins_f64) return false; continue; }   // tcgen05: default mapping has one
    if (k != 1 || !k1_ring_dual_full_supported(h->d, eb)) return false;
  }
  return h->k1_diag == 0;
}

// One applySmooth (AGD.scala:192-208) at the device-resident point `w_of(dev)`: K1 over every local
// shard, slab reduction, one all-reduce of [grad | loss | count | loss2 | count2].  Result: Dev::acc on every device.
// w2_of != nullptr: the same sweep also evaluates the loss (not the gradient) at `w2_of(dev)` -> acc[d+2], acc[d+3];
// with dual_full also the gradient there -> a second block acc[d+4 .. 2d+7] = [grad | loss | count | 0 | 0].
// defer_gather: on the peer-memory path the gather is left to the next K3 kernel (h->xg_pending; see XchgGather) -- one launch
// fewer per sweep; the caller must hand the pending exchange to a k3_step / k3_gx launch before anything else reads acc.
int smooth_device(agd_handle *h, int kind, WSel w_of, bool timed, WSel w2_of = nullptr, bool dual_full = false,
                  bool defer_gather = false) {
  if (h->xg_pending) return fail(h, "internal: an exchange is still waiting for its consumer");
  const int32_t d = h->d;
  const int32_t n = (dual_full ? 2 : 1) * (d + 4);   // doubles this sweep produces and exchanges
  const bool p2p = h->world > 1 && h->x_p2p;
  const bool rs = p2p && n >= kXchgRsMin;    // large payloads: reduce-scatter + all-gather instead of the one-shot exchange
  const unsigned long long epoch = p2p ? ++h->x_epoch : 0ull;
  auto make_rs = [&](Dev &D, size_t i) {
    XchgRs x;
    x.peers = D.xpeers; x.world = h->world; x.my_rank = h->first_rank + (int)i; x.buf = (int)(epoch & 1ull);
    x.n = n; x.slot_stride = 2 * (d + 4); x.epoch = epoch; x.ticket = D.xticket;
    return x;
  };
  auto make_pub = [&](Dev &D, size_t i) {
    XchgPub pub;
    pub.peers = D.xpeers; pub.world = h->world; pub.my_rank = h->first_rank + (int)i; pub.buf = (int)(epoch & 1ull);
    pub.n = n; pub.slot_stride = 2 * (d + 4); pub.epoch = epoch; pub.ticket = D.xticket;
    return pub;
  };
  for (size_t i = 0; i < h->devs.size(); ++i) {
    Dev &D = h->devs[i];
    CK(cudaSetDevice(D.ordinal));
    const Shard &s = D.sh;
    const bool t0 = timed && i == 0;
    if (s.csr) {
      if (dual_full) return fail(h, "internal: two-gradient sweep requested on a CSR shard");
      K1CsrArgs a;
      a.rowptr = s.rowptr; a.idx = s.idx; a.val = s.val; a.labels = s.labels; a.w = w_of(D);
      a.w2 = w2_of ? w2_of(D) : nullptr;
      a.gacc = D.acc; a.rows = s.rows; a.d = d; a.kind = kind;
      a.sample_seed = h->sample_seed; a.sample_thresh = h->sample_thresh; a.row_base = D.row_base; a.tune = h->tune_rows;
      if (t0) CK(cudaEventRecord(next_event(D.ev, D.ev_used), D.st));
      CK(k1_csr_launch(a, s.elem_bytes, D.sm_count, D.st));
      if (t0) CK(cudaEventRecord(next_event(D.ev, D.ev_used), D.st));
      if (rs) {
        const XchgRs x = make_rs(D, i);
        CK(xchg_rs_publish_launch(D.acc, x, D.st));
        CK(xchg_rs_reduce_bcast_launch(D.xbuf, D.xflags, x, D.st));
      } else if (p2p) { const XchgPub pub = make_pub(D, i); CK(xchg_publish_launch(D.acc, pub, D.st)); }
      h->launches += (i == 0) ? (rs ? 4 : (p2p ? 3 : 2)) : 0;
      continue;
    }
    K1Args a;
    a.X = s.X; a.labels = s.labels; a.w = w_of(D); a.w2 = w2_of ? w2_of(D) : nullptr; a.dual_full = dual_full ? 1 : 0;
    a.rows = s.rows; a.d = d; a.kind = h->k1_diag ? h->k1_diag : kind;
    a.stages = h->ring_stages; a.slab_stride = n;
    a.sample_seed = h->sample_seed; a.sample_thresh = h->sample_thresh; a.row_base = D.row_base; a.tune_rows = h->tune_rows; a.tune_ctas = h->tune_ctas; a.tune_full = h->tune_full;
    a.tc_margins_f64 = h->tc_margins_f64;
    const int eb = s.elem_bytes ? s.elem_bytes : 4;
    bool ring = k1_ring_supported(d, eb) != 0;
    if (h->k1_variant == 2) ring = false;
    const bool tc = k1_tc_supported(d, eb) && (h->k1_variant == 0 || h->k1_variant == 4);
    if (h->k1_variant == 4 && !tc) return fail(h, "tcgen05 kernel needs bf16 storage with d %% 128 == 0 and d <= 4096 (d=%d)", d);
    if (h->k1_variant == 1 && !ring) return fail(h, "ring kernel does not support d=%d with %d-byte elements", d, eb);
    if (a.w2 && ((tc && (h->tune_rows != 0 || h->tc_margins_f64)) || (!tc && ring && !k1_ring_dual_supported(d, eb))))
      return fail(h, "internal: two-point sweep requested on a kernel without one");
    if (dual_full && ((tc && (h->tune_rows != 0 || h->tc_margins_f64)) || (!tc && (!ring || !k1_ring_dual_full_supported(d, eb)))))
      return fail(h, "internal: two-gradient sweep requested on a kernel without one");
    int max_blocks = k1_max_blocks(D.sm_count);
    if (!ring) {  // generic: bound the slab memory for very wide rows
      const long long lim = (32LL << 20) / ((long long)d + 4);
      if (lim < max_blocks) max_blocks = lim < 1 ? 1 : (int)lim;
    }
    if (ensure_slabs(h, D, max_blocks, n)) return 1;
    a.slabs = D.slabs;
    int blocks = 0;
    if (t0) CK(cudaEventRecord(next_event(D.ev, D.ev_used), D.st));
    if (tc) CK(k1_tc_launch(a, D.sm_count, &blocks, D.st));
    else if (ring) CK(k1_ring_launch(a, eb, D.sm_count, &blocks, D.st));
    else CK(k1_generic_launch(a, eb, D.sm_count, max_blocks, &blocks, D.st));
    if (t0) CK(cudaEventRecord(next_event(D.ev, D.ev_used), D.st));
    if (i == 0) trace_mark(h, dual_full ? "K1x2" : (w2_of ? "K1+loss" : "K1"));
    if (p2p && !rs) { const XchgPub pub = make_pub(D, i); CK(k1_reduce_launch(D.slabs, blocks, n, D.acc, &pub, D.st)); }
    else CK(k1_reduce_launch(D.slabs, blocks, n, D.acc, nullptr, D.st));
    if (rs) {
      const XchgRs x = make_rs(D, i);
      CK(xchg_rs_publish_launch(D.acc, x, D.st));
      CK(xchg_rs_reduce_bcast_launch(D.xbuf, D.xflags, x, D.st));
      if (i == 0) h->launches += 2;
    }
    if (i == 0) trace_mark(h, p2p ? "reduce+publish" : "reduce");
    if (i == 0) h->launches += (s.rows > 0 ? 2 : 1);
  }
  if (p2p && defer_gather) {
    h->xg_pending = true;
    h->xg_rs = rs;
    h->xg_epoch = epoch;
    h->xg_n = n;
    h->collectives += 1;
  } else if (p2p) {  // K2': every rank already holds every rank's partial sums; add them in rank order
    Dev &D0 = h->devs[0];
    if (timed) { CK(cudaSetDevice(D0.ordinal)); CK(cudaEventRecord(next_event(D0.ev_ar, D0.ev_ar_used), D0.st)); }
    for (Dev &D : h->devs) {
      CK(cudaSetDevice(D.ordinal));
      if (rs) CK(xchg_rs_gather_launch(D.xbuf, D.xflags, h->world, (int)(epoch & 1ull), n, 2 * (d + 4), epoch, D.acc, D.st));
      else CK(xchg_gather_launch(D.xbuf, D.xflags, h->world, (int)(epoch & 1ull), n, 2 * (d + 4), epoch, D.acc, D.st));
    }
    if (timed) { CK(cudaSetDevice(D0.ordinal)); CK(cudaEventRecord(next_event(D0.ev_ar, D0.ev_ar_used), D0.st)); }
    trace_mark(h, "gather");
    h->launches += 1;
    h->collectives += 1;
  } else if (h->world > 1) {
    if (!h->comm_ready || !h->devs[0].comm) return fail(h, "world_ranks=%d but there is no communicator (agd_comm_init)", h->world);
    NcclApi &N = nccl_api();
    Dev &D0 = h->devs[0];
    if (timed) { CK(cudaSetDevice(D0.ordinal)); CK(cudaEventRecord(next_event(D0.ev_ar, D0.ev_ar_used), D0.st)); }
    CKN(N.GroupStart());
    for (Dev &D : h->devs) CKN(N.AllReduce(D.acc, D.acc, (size_t)n, ncclDouble, ncclSum, D.comm, D.st));
    CKN(N.GroupEnd());
    if (timed) { CK(cudaSetDevice(D0.ordinal)); CK(cudaEventRecord(next_event(D0.ev_ar, D0.ev_ar_used), D0.st)); }
    h->collectives += 1;
  }
  return 0;
}

int read_scalars(agd_handle *h, double *out) {
  Dev &D = h->devs[0];
  CK(cudaSetDevice(D.ordinal));
  CK(cudaStreamSynchronize(D.st));  // the kernel's zero-copy stores are visible once the stream has drained
  memcpy(out, D.scalars_host, K3_NS * sizeof(double));
  return 0;
}

// The same without draining the stream: the last K3 kernel of a round stores `seq` behind its scalars in mapped pinned memory
// (after a system-scope fence); the host polls that word.  No driver call sits between the kernel's last store and the host
// loop continuing, and the stream may already hold later work.  A stream query every few thousand polls catches errors.
int wait_scalars(agd_handle *h, unsigned long long seq, double *out) {
  Dev &D = h->devs[0];
  volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(D.scalars_host + 2 * K3_NS);
  unsigned int spins = 0;
  while (*flag != seq) {
    if ((++spins & 0x3fffu) == 0) {
      CK(cudaSetDevice(D.ordinal));
      const cudaError_t q = cudaStreamQuery(D.st);
      if (q == cudaSuccess) { if (*flag != seq) return fail(h, "internal: the round's scalars never arrived"); break; }
      if (q != cudaErrorNotReady) return fail(h, "stream failed while waiting for a round: %s", cudaGetErrorString(q));
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  memcpy(out, D.scalars_host, K3_NS * sizeof(double));
  return 0;
}

int sum_events(agd_handle *h, std::vector<cudaEvent_t> &pool, size_t used, double *ms_out) {
  double ms = 0.0;
  for (size_t i = 0; i + 1 < used; i += 2) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, pool[i], pool[i + 1]));
    ms += t;
  }
  *ms_out = ms;
  return 0;
}

int check_ready(agd_handle *h) {
  if (!h) return 1;
  if (h->d <= 0) return fail(h, "no shard loaded (call agd_load_dense / agd_load_csr / agd_generate first)");
  for (Dev &D : h->devs)
    if (ensure_vectors(h, D, h->d)) return 1;
  if (h->world > 1 && !h->comm_ready) return fail(h, "world_ranks=%d but agd_comm_init was not called", h->world);
  return ensure_xchg(h);
}

int call_begin(agd_handle *h) {
  h->xg_pending = false;   // a call that failed half-way may have left one behind
  Dev &D = h->devs[0];
  CK(cudaSetDevice(D.ordinal));
  if (!h->ev_begin) { CK(cudaEventCreate(&h->ev_begin)); CK(cudaEventCreate(&h->ev_end)); }
  h->launches = 0;
  h->collectives = 0;
  D.ev_used = D.ev_ar_used = 0;
  CK(cudaEventRecord(h->ev_begin, D.st));
  return 0;
}

// all devices drained; fills the timing part of the stats
int call_end(agd_handle *h, agd_stats &s, std::chrono::steady_clock::time_point t_begin) {
  Dev &D0 = h->devs[0];
  CK(cudaSetDevice(D0.ordinal));
  CK(cudaEventRecord(h->ev_end, D0.st));
  for (Dev &D : h->devs) { CK(cudaSetDevice(D.ordinal)); CK(cudaStreamSynchronize(D.st)); }
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end));
  s.device_ms_total = ms;
  if (sum_events(h, D0.ev, D0.ev_used, &s.k1_ms_total)) return 1;
  if (sum_events(h, D0.ev_ar, D0.ev_ar_used, &s.allreduce_ms_total)) return 1;
  s.k1_launches = (int64_t)(D0.ev_used / 2);
  s.gpu_launches = h->launches;
  s.collective_calls = h->collectives;
  s.collective_kind = h->x_p2p ? 1 : 0;
  s.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  return 0;
}

double reg_value(int updater, double reg, double sum_sq, double sum_abs) {
  if (updater == AGD_UPD_SQUARED_L2) {
    const double nrm = std::sqrt(sum_sq);  // brzNorm(w, 2.0)
    return 0.5 * reg * nrm * nrm;
  }
  if (updater == AGD_UPD_L1) return sum_abs * reg;
  return 0.0;
}

}  // namespace

namespace agd {
void set_last_error(agd_handle *h, const char *msg) { fail(h, "%s", msg); }
}  // namespace agd

// ================================================================ C-ABI
extern "C" {

int agd_abi_version(void) { return AGD_B200_ABI_VERSION; }
int agd_sizeof_params(void) { return (int)sizeof(agd_params); }
int agd_sizeof_stats(void) { return (int)sizeof(agd_stats); }

void agd_default_params(agd_params *p) {  // AGD.scala:44-51
  memset(p, 0, sizeof *p);
  p->convergence_tol = 1e-4;
  p->num_iterations = 100;
  p->reg_param = 0.0;
  p->L0 = 1.0;
  p->Lexact = std::numeric_limits<double>::infinity();
  p->beta = 0.5;
  p->alpha = 0.9;
  p->may_restart = 1;
  p->gradient = AGD_GRAD_LOGISTIC;
  p->updater = AGD_UPD_SIMPLE;
  p->flags = 0;
}

const char *agd_last_error(const agd_handle *h) {
  if (h) return h->err.c_str();
  return g_create_error.c_str();
}

int agd_create(const int32_t *device_ids, int32_t n_dev, agd_handle **out) {
  agd_handle *h = nullptr;  // errors before the handle exists go to the global slot
  if (!out) return fail(h, "out is NULL");
  *out = nullptr;
  if (n_dev < 1 || !device_ids) return fail(h, "need at least one device");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(h, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  agd_handle *nh = new agd_handle();
  nh->devs.resize(n_dev);
  for (int i = 0; i < n_dev; ++i) {
    Dev &D = nh->devs[i];
    D.ordinal = device_ids[i];
    D.mu = new std::mutex();
    if (D.ordinal < 0 || D.ordinal >= count) { fail(h, "device ordinal %d out of range (0..%d)", D.ordinal, count - 1); agd_destroy(nh); return 1; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, D.ordinal) != cudaSuccess || prop.major != 10) {
      fail(h, "device %d is not an sm_100 (Blackwell B200) GPU; kernels are built for sm_100a only", D.ordinal);
      agd_destroy(nh);
      return 1;
    }
    D.sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(D.ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&D.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&D.ticket, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(D.ticket, 0, sizeof(unsigned int)) != cudaSuccess ||
        cudaHostAlloc(&D.scalars_host, (2 * K3_NS + 2) * sizeof(double), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&D.scalars_dev, D.scalars_host, 0) != cudaSuccess) {
      fail(h, "device %d setup failed: %s", D.ordinal, cudaGetErrorString(cudaGetLastError()));
      agd_destroy(nh);
      return 1;
    }
    memset(D.scalars_host, 0, (2 * K3_NS + 2) * sizeof(double));   // the sequence word behind the scalars starts at 0
  }
  nh->world = n_dev;
  nh->first_rank = 0;
  for (int i = 0; i < n_dev; ++i) nh->devs[i].row_base = (long long)i << 40;  // loaded shards: one mask stream per rank
  *out = nh;
  // A single-process owner of several GPUs is a complete world of its own: ranks 0..n_dev-1 exchange through direct peer
  // pointers, and NCCL is built only if it is ever needed (ensure_nccl).  agd_comm_init / agd_comm_init_ipc replace this
  // default when the process is part of a larger world.
  nh->comm_auto = true;
  nh->comm_ready = true;
  return 0;
}

int agd_destroy(agd_handle *h) {
  if (!h) return 0;
  free_xchg(h);
  for (Dev &D : h->devs) {
    cudaSetDevice(D.ordinal);
    cudaStreamSynchronize(D.st);
    if (D.comm && nccl_api().ok) nccl_api().CommDestroy(D.comm);
    D.comm = nullptr;
    free_shard(h, D);
    double *v[] = {D.x, D.z, D.x_old, D.z_old, D.y, D.g_y, D.g_x, D.wtmp, D.y_spec, D.acc, D.slabs, D.partials};
    for (double *p : v)
      if (p) cudaFree(p);
    if (D.ticket) cudaFree(D.ticket);
    if (D.scalars_host) cudaFreeHost(D.scalars_host);
    if (D.hist_host) cudaFreeHost(D.hist_host);
    if (D.stage_dev) cudaFree(D.stage_dev);
    for (cudaEvent_t e : D.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : D.ev_ar) cudaEventDestroy(e);
    if (&D == &h->devs[0] && h->ev_begin) { cudaEventDestroy(h->ev_begin); cudaEventDestroy(h->ev_end); }
    if (D.st) cudaStreamDestroy(D.st);
    delete D.mu;
  }
  delete h;
  return 0;
}

int agd_comm_unique_id(void *out128) {
  agd_handle *h = nullptr;
  NcclApi &N = nccl_api();
  if (!N.ok) return fail(h, "NCCL unavailable: %s", N.why.c_str());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  ncclResult_t r = N.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(h, "ncclGetUniqueId failed: %s", N.GetErrorString(r));
  memcpy(out128, &id, 128);
  return 0;
}

int agd_comm_init(agd_handle *h, const void *id128, int32_t world_ranks, int32_t first_rank) {
  if (!h) return 1;
  NcclApi &N = nccl_api();
  if (!N.ok) return fail(h, "NCCL unavailable: %s", N.why.c_str());
  const int nd = (int)h->devs.size();
  if (world_ranks < nd || first_rank < 0 || first_rank + nd > world_ranks)
    return fail(h, "bad rank layout: world=%d first=%d local=%d", world_ranks, first_rank, nd);
  if (h->comm_ready && !h->comm_auto) return fail(h, "communicator already initialised");
  // replaces the default single-process world of agd_create (its NCCL communicator, if one was ever built, and its exchange)
  free_xchg(h);
  destroy_comms(h);
  h->comm_auto = false;
  h->ipc_only = false;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  if (world_ranks > 1) {
    CKN(N.GroupStart());
    for (int i = 0; i < nd; ++i) {
      CK(cudaSetDevice(h->devs[i].ordinal));
      CKN(N.CommInitRank(&h->devs[i].comm, world_ranks, id, first_rank + i));
    }
    CKN(N.GroupEnd());
  }
  h->world = world_ranks;
  h->first_rank = first_rank;
  h->comm_ready = true;
  for (int i = 0; i < nd; ++i) h->devs[i].row_base = (long long)(first_rank + i) << 40;
  return 0;
}

// ---- the same world without NCCL: the host language ships the CUDA IPC handles of the exchange buffers
int agd_comm_init_ipc(agd_handle *h, int32_t world_ranks, int32_t first_rank) {
  if (!h) return 1;
  const int nd = (int)h->devs.size();
  if (world_ranks < nd || first_rank < 0 || first_rank + nd > world_ranks)
    return fail(h, "bad rank layout: world=%d first=%d local=%d", world_ranks, first_rank, nd);
  if (world_ranks > kMaxRanks) return fail(h, "the peer-memory exchange supports at most %d ranks", kMaxRanks);
  if (h->comm_ready && !h->comm_auto) return fail(h, "communicator already initialised");
  free_xchg(h);
  destroy_comms(h);
  h->comm_auto = false;
  h->ipc_only = world_ranks > nd;   // a world of local GPUs only needs no handles at all
  if (!h->ipc_only) h->comm_auto = true;
  h->world = world_ranks;
  h->first_rank = first_rank;
  h->comm_ready = true;
  for (int i = 0; i < nd; ++i) h->devs[i].row_base = (long long)(first_rank + i) << 40;
  return 0;
}

int agd_xchg_export(agd_handle *h, void *out, int64_t capacity_bytes, int64_t *bytes_written) {
  if (!h || !out || !bytes_written) return 1;
  if (!h->ipc_only) return fail(h, "agd_xchg_export needs a world set up with agd_comm_init_ipc");
  if (h->d <= 0) return fail(h, "load the shards first: the exchange buffers are sized by the feature dimension");
  const int nd = (int)h->devs.size();
  if (capacity_bytes < (int64_t)nd * AGD_XCHG_HANDLE_BYTES) return fail(h, "capacity too small: need %d bytes", nd * AGD_XCHG_HANDLE_BYTES);
  free_xchg(h);
  std::vector<XHandles> mine;
  if (xchg_alloc(h, mine)) return 1;
  for (Dev &D : h->devs) { CK(cudaSetDevice(D.ordinal)); CK(cudaDeviceSynchronize()); }  // the zeroed buffers are visible before any peer maps them
  memcpy(out, mine.data(), (size_t)nd * sizeof(XHandles));
  *bytes_written = (int64_t)nd * AGD_XCHG_HANDLE_BYTES;
  return 0;
}

int agd_xchg_import(agd_handle *h, const void *all_ranks, int64_t bytes) {
  if (!h || !all_ranks) return 1;
  if (!h->ipc_only) return fail(h, "agd_xchg_import needs a world set up with agd_comm_init_ipc");
  if (!h->devs[0].xbuf) return fail(h, "call agd_xchg_export first");
  if (bytes != (int64_t)h->world * AGD_XCHG_HANDLE_BYTES) return fail(h, "expected %d handle blobs (%d bytes)", h->world, h->world * AGD_XCHG_HANDLE_BYTES);
  std::vector<XHandles> all((size_t)h->world);
  memcpy(all.data(), all_ranks, (size_t)bytes);
  bool ok = true;
  if (xchg_map(h, all, &ok)) return 1;
  if (!ok) {
    free_xchg(h);
    return fail(h, "peer-memory exchange unavailable: some pair of ranks cannot map each other's buffers (or the blobs are not in rank order / of another dimension)");
  }
  h->x_d = h->d;
  h->x_p2p = true;
  h->x_epoch = 0;
  return 0;
}

int agd_reserve(agd_handle *h, int32_t dev, int64_t rows_capacity, int32_t d, int32_t store_dtype) {
  if (!h) return 1;
  if (dev < 0 || dev >= (int)h->devs.size()) return fail(h, "bad local device index %d", dev);
  if (set_dim(h, d, dtype_bytes(store_dtype))) return 1;
  Dev &D = h->devs[dev];
  std::lock_guard<std::mutex> g(*D.mu);
  return reserve_locked(h, D, rows_capacity, h->d, store_dtype);
}

int agd_load_dense(agd_handle *h, int32_t dev, const void *X, int32_t src_dtype, const double *labels,
                   int64_t rows, int32_t d, int64_t ld, int32_t store_dtype) {
  if (!h) return 1;
  if (dev < 0 || dev >= (int)h->devs.size()) return fail(h, "bad local device index %d", dev);
  const int sb = dtype_bytes(src_dtype);
  if (sb != 4 && sb != 8) return fail(h, "src_dtype must be AGD_F32 or AGD_F64");
  if (rows < 0 || ld < d) return fail(h, "bad geometry rows=%lld d=%d ld=%lld", (long long)rows, d, (long long)ld);
  if (rows > 0 && (!X || !labels)) return fail(h, "NULL data pointer");
  if (set_dim(h, d, dtype_bytes(store_dtype))) return 1;
  const int32_t di = h->d;  // stored row length (>= d, zero-padded)
  Dev &D = h->devs[dev];
  std::lock_guard<std::mutex> g(*D.mu);
  Shard &s = D.sh;
  const int64_t need = s.rows + rows;
  if (s.cap == 0 || need > s.cap) {
    const int64_t cap = s.cap == 0 ? need : (need > 2 * s.cap ? need : 2 * s.cap);
    if (reserve_locked(h, D, cap, di, s.cap ? bytes_dtype(s.elem_bytes) : store_dtype)) return 1;
  }
  const int eb = s.elem_bytes;
  if (dtype_bytes(store_dtype) != eb) return fail(h, "storage dtype mismatch with the resident shard");
  CK(cudaSetDevice(D.ordinal));
  unsigned char *dst = (unsigned char *)s.X + (size_t)s.rows * di * eb;
  const unsigned char *src = (const unsigned char *)X;
  if (rows > 0) {
    if (sb == eb && ld == d && di == d) {
      CK(cudaMemcpyAsync(dst, src, (size_t)rows * d * eb, cudaMemcpyHostToDevice, D.st));
    } else {
      int64_t chunk = (int64_t)((64u << 20) / ((size_t)ld * sb));
      if (chunk < 1) chunk = 1;
      if (chunk > rows) chunk = rows;
      if (ensure_stage(h, D, (size_t)chunk * ld * sb)) return 1;
      for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        const int64_t rc = rows - r0 < chunk ? rows - r0 : chunk;
        CK(cudaMemcpyAsync(D.stage_dev, src + (size_t)r0 * ld * sb, (size_t)rc * ld * sb, cudaMemcpyHostToDevice, D.st));
        CK(convert_rows_launch(dst + (size_t)r0 * di * eb, eb, D.stage_dev, sb, rc, d, ld, di, D.st));
        CK(cudaStreamSynchronize(D.st));  // the staging buffer is reused
      }
    }
    CK(cudaMemcpyAsync(s.labels + s.rows, labels, (size_t)rows * sizeof(double), cudaMemcpyHostToDevice, D.st));
  }
  CK(cudaStreamSynchronize(D.st));
  s.rows += rows;
  return 0;
}

// grows the CSR arrays of a device to hold `rows_cap` rows and `nnz_cap` entries (contents preserved)
static int csr_reserve_locked(agd_handle *h, Dev &D, int64_t rows_cap, int64_t nnz_cap, int eb) {
  Shard &s = D.sh;
  CK(cudaSetDevice(D.ordinal));
  if (!s.csr && (s.cap > 0 || s.rows > 0)) return fail(h, "device already holds a dense shard");
  if (s.csr && s.elem_bytes != eb) return fail(h, "storage dtype mismatch with the resident shard");
  if (rows_cap > s.cap || !s.rowptr) {
    int64_t *nr = nullptr;
    double *nl = nullptr;
    CK(cudaMalloc(&nr, ((size_t)rows_cap + 1) * sizeof(int64_t)));
    CK(cudaMalloc(&nl, ((size_t)rows_cap + 64) * sizeof(double)));
    if (s.rowptr) {
      CK(cudaMemcpyAsync(nr, s.rowptr, ((size_t)s.rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, D.st));
      CK(cudaMemcpyAsync(nl, s.labels, (size_t)s.rows * sizeof(double), cudaMemcpyDeviceToDevice, D.st));
    } else {
      CK(cudaMemsetAsync(nr, 0, sizeof(int64_t), D.st));
    }
    CK(cudaStreamSynchronize(D.st));
    if (s.rowptr) cudaFree(s.rowptr);
    if (s.labels) cudaFree(s.labels);
    s.rowptr = nr; s.labels = nl; s.cap = rows_cap;
  }
  if (nnz_cap > s.nnz_cap || !s.idx) {
    int32_t *ni = nullptr;
    void *nv = nullptr;
    CK(cudaMalloc(&ni, ((size_t)nnz_cap + 4) * sizeof(int32_t)));
    CK(cudaMalloc(&nv, ((size_t)nnz_cap + 4) * eb));
    if (s.idx && s.nnz > 0) {
      CK(cudaMemcpyAsync(ni, s.idx, (size_t)s.nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, D.st));
      CK(cudaMemcpyAsync(nv, s.val, (size_t)s.nnz * eb, cudaMemcpyDeviceToDevice, D.st));
    }
    CK(cudaStreamSynchronize(D.st));
    if (s.idx) cudaFree(s.idx);
    if (s.val) cudaFree(s.val);
    s.idx = ni; s.val = nv; s.nnz_cap = nnz_cap;
  }
  s.csr = true;
  s.elem_bytes = eb;
  return 0;
}

int agd_load_csr(agd_handle *h, int32_t dev, const int64_t *rowptr, const int32_t *idx, const void *val,
                 int32_t src_dtype, const double *labels, int64_t rows, int32_t d, int32_t store_dtype) {
  if (!h) return 1;
  if (dev < 0 || dev >= (int)h->devs.size()) return fail(h, "bad local device index %d", dev);
  const int sb = dtype_bytes(src_dtype), eb = dtype_bytes(store_dtype);
  if ((sb != 4 && sb != 8) || (eb != 4 && eb != 8)) return fail(h, "CSR dtypes must be AGD_F32 or AGD_F64");
  if (rows < 0 || (rows > 0 && (!rowptr || !labels))) return fail(h, "bad CSR arguments");
  if (set_dim(h, d)) return 1;
  Dev &D = h->devs[dev];
  std::lock_guard<std::mutex> g(*D.mu);
  Shard &s = D.sh;
  if (rows > 0 && rowptr[0] != 0) return fail(h, "rowptr[0] must be 0");
  const int64_t nnz = rows > 0 ? rowptr[rows] : 0;
  if (nnz < 0) return fail(h, "rowptr[rows] = %lld is negative", (long long)nnz);
  if (nnz > 0 && (!idx || !val)) return fail(h, "NULL index / value pointer");
  // APPENDS rows (Spark hands partitions over one at a time); arrays grow geometrically
  const int64_t need_rows = s.rows + rows, need_nnz = s.nnz + nnz;
  const int64_t rows_cap = need_rows > s.cap ? (need_rows > 2 * s.cap ? need_rows : 2 * s.cap) : s.cap;
  const int64_t nnz_cap = need_nnz > s.nnz_cap ? (need_nnz > 2 * s.nnz_cap ? need_nnz : 2 * s.nnz_cap) : s.nnz_cap;
  if (csr_reserve_locked(h, D, rows_cap, nnz_cap, eb)) return 1;
  if (rows > 0) {
    // device rowptr entries for the new rows = host rowptr[1..rows] + resident nnz
    if (ensure_stage(h, D, ((size_t)rows + 1) * sizeof(int64_t) + (size_t)nnz * sb + 64)) return 1;
    CK(cudaMemcpyAsync(D.stage_dev, rowptr, ((size_t)rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, D.st));
    CK(csr_shift_rowptr_launch(s.rowptr + s.rows, (const int64_t *)D.stage_dev, rows + 1, s.nnz, D.st));
    CK(cudaMemcpyAsync(s.labels + s.rows, labels, (size_t)rows * sizeof(double), cudaMemcpyHostToDevice, D.st));
  }
  if (nnz > 0) {
    CK(cudaMemcpyAsync(s.idx + s.nnz, idx, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, D.st));
    if (sb == eb) {
      CK(cudaMemcpyAsync((unsigned char *)s.val + (size_t)s.nnz * eb, val, (size_t)nnz * eb, cudaMemcpyHostToDevice, D.st));
    } else {
      unsigned char *stage_vals = (unsigned char *)D.stage_dev + (((size_t)rows + 1) * sizeof(int64_t) + 63) / 64 * 64;
      CK(cudaMemcpyAsync(stage_vals, val, (size_t)nnz * sb, cudaMemcpyHostToDevice, D.st));
      CK(convert_rows_launch((unsigned char *)s.val + (size_t)s.nnz * eb, eb, stage_vals, sb, nnz, 1, 1, 1, D.st));
    }
  }
  // Validate on the device before the partition becomes part of the shard: the gradient kernel gathers w[idx] and
  // scatters into g[idx] with these raw indices, so one bad column id would be an out-of-bounds device write.
  if (rows > 0) {
    int *flag = nullptr, host_flag = 0;   // D.stage_dev still holds the caller's rowptr
    CK(cudaMalloc(&flag, sizeof(int)));
    CK(cudaMemsetAsync(flag, 0, sizeof(int), D.st));
    CK(csr_validate_launch((const int64_t *)D.stage_dev, rows, s.idx + s.nnz, nnz, d, flag, D.st));
    CK(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, D.st));
    CK(cudaStreamSynchronize(D.st));
    cudaFree(flag);
    if (host_flag == 1) return fail(h, "bad CSR partition: rowptr must be non-decreasing from 0 to nnz=%lld", (long long)nnz);
    if (host_flag == 2) return fail(h, "bad CSR partition: a column index lies outside [0, %d) (SparseVector size differs from the weights?)", d);
  }
  CK(cudaStreamSynchronize(D.st));
  s.rows = need_rows;
  s.nnz = need_nnz;
  return 0;
}

int agd_generate_csr(agd_handle *h, int64_t total_rows, int32_t d, int32_t nnz_per_row, int32_t store_dtype,
                     uint64_t seed, int32_t gradient) {
  if (!h) return 1;
  const int eb = dtype_bytes(store_dtype);
  if (eb != 4 && eb != 8) return fail(h, "CSR store_dtype must be AGD_F32 or AGD_F64");
  if (total_rows < 0 || nnz_per_row < 1 || nnz_per_row > d) return fail(h, "bad CSR geometry");
  if (agd_clear(h)) return 1;
  if (set_dim(h, d)) return 1;
  for (size_t i = 0; i < h->devs.size(); ++i) {
    Dev &D = h->devs[i];
    std::lock_guard<std::mutex> g(*D.mu);
    const long long rank = h->first_rank + (long long)i, W = h->world;
    const int64_t lo = (int64_t)(((__int128)rank * total_rows) / W), hi = (int64_t)(((__int128)(rank + 1) * total_rows) / W);
    if (csr_reserve_locked(h, D, hi - lo, (hi - lo) * nnz_per_row, eb)) return 1;
    if (ensure_vectors(h, D, d)) return 1;
    CK(cudaSetDevice(D.ordinal));
    D.row_base = lo;
    CK(synth_wtrue_launch(D.wtmp, seed, d, D.st));
    CK(synth_csr_launch(D.sh.rowptr, D.sh.idx, D.sh.val, eb, D.wtmp, D.sh.labels, seed, gradient, lo, hi - lo, d,
                        nnz_per_row, D.st));
    D.sh.rows = hi - lo;
    D.sh.nnz = (hi - lo) * nnz_per_row;
  }
  for (Dev &D : h->devs) { CK(cudaSetDevice(D.ordinal)); CK(cudaStreamSynchronize(D.st)); }
  return 0;
}

int agd_get_csr_rows(agd_handle *h, int32_t dev, int64_t row0, int64_t rows, int64_t *rowptr_out, int32_t *idx_out,
                     void *val_out, int64_t nnz_capacity, double *labels_out) {
  if (!h) return 1;
  if (dev < 0 || dev >= (int)h->devs.size()) return fail(h, "bad local device index %d", dev);
  Dev &D = h->devs[dev];
  const Shard &s = D.sh;
  if (!s.csr) return fail(h, "agd_get_csr_rows serves CSR shards only");
  if (row0 < 0 || rows < 0 || row0 + rows > s.rows) return fail(h, "row range out of bounds");
  CK(cudaSetDevice(D.ordinal));
  CK(cudaMemcpyAsync(rowptr_out, s.rowptr + row0, ((size_t)rows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, D.st));
  CK(cudaStreamSynchronize(D.st));
  const int64_t a = rowptr_out[0], b = rowptr_out[rows];
  if (b - a > nnz_capacity) return fail(h, "nnz_capacity too small: need %lld", (long long)(b - a));
  if (b > a) {
    CK(cudaMemcpyAsync(idx_out, s.idx + a, (size_t)(b - a) * sizeof(int32_t), cudaMemcpyDeviceToHost, D.st));
    CK(cudaMemcpyAsync(val_out, (const unsigned char *)s.val + (size_t)a * s.elem_bytes, (size_t)(b - a) * s.elem_bytes,
                       cudaMemcpyDeviceToHost, D.st));
  }
  if (labels_out && rows) CK(cudaMemcpyAsync(labels_out, s.labels + row0, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, D.st));
  CK(cudaStreamSynchronize(D.st));
  for (int64_t i = rows; i >= 0; --i) rowptr_out[i] -= a;
  return 0;
}

int agd_clear(agd_handle *h) {
  if (!h) return 1;
  for (Dev &D : h->devs) {
    std::lock_guard<std::mutex> g(*D.mu);
    CK(cudaSetDevice(D.ordinal));
    CK(cudaStreamSynchronize(D.st));
    if (free_shard(h, D)) return 1;
  }
  h->d = 0;
  h->d_user = 0;
  return 0;
}

int64_t agd_rows(const agd_handle *h, int32_t dev) {
  if (!h || dev < 0 || dev >= (int)h->devs.size()) return -1;
  return h->devs[dev].sh.rows;
}
int32_t agd_dim(const agd_handle *h) { return h ? h->d_user : 0; }

int agd_generate(agd_handle *h, int64_t total_rows, int32_t d, int32_t store_dtype, uint64_t seed, int32_t gradient) {
  if (!h) return 1;
  const int eb = dtype_bytes(store_dtype);
  if (!eb) return fail(h, "store_dtype must be AGD_F64, AGD_F32 or AGD_BF16");
  if (total_rows < 0) return fail(h, "negative row count");
  if (agd_clear(h)) return 1;
  if (set_dim(h, d, eb)) return 1;
  const int32_t di = h->d;
  for (size_t i = 0; i < h->devs.size(); ++i) {
    Dev &D = h->devs[i];
    std::lock_guard<std::mutex> g(*D.mu);
    const long long rank = h->first_rank + (long long)i, W = h->world;
    const int64_t lo = (int64_t)(((__int128)rank * total_rows) / W), hi = (int64_t)(((__int128)(rank + 1) * total_rows) / W);
    if (reserve_locked(h, D, hi - lo, di, store_dtype)) return 1;
    if (ensure_vectors(h, D, di)) return 1;
    CK(cudaSetDevice(D.ordinal));
    D.row_base = lo;
    CK(synth_dense_launch(D.sh.X, eb, seed, lo, hi - lo, d, di, D.st));
    CK(synth_wtrue_launch(D.wtmp, seed, d, D.st));
    CK(synth_labels_launch(D.sh.X, eb, D.wtmp, D.sh.labels, seed, gradient, lo, hi - lo, d, di, D.st));
    CK(cudaMemsetAsync(D.wtmp, 0, ((size_t)di + 2) * sizeof(double), D.st));
    D.sh.rows = hi - lo;
  }
  for (Dev &D : h->devs) { CK(cudaSetDevice(D.ordinal)); CK(cudaStreamSynchronize(D.st)); }
  return 0;
}

int agd_get_rows(agd_handle *h, int32_t dev, int64_t row0, int64_t rows, void *X_out, double *labels_out) {
  if (!h) return 1;
  if (dev < 0 || dev >= (int)h->devs.size()) return fail(h, "bad local device index %d", dev);
  Dev &D = h->devs[dev];
  const Shard &s = D.sh;
  if (s.csr) return fail(h, "agd_get_rows serves dense shards only");
  if (row0 < 0 || rows < 0 || row0 + rows > s.rows) return fail(h, "row range out of bounds");
  CK(cudaSetDevice(D.ordinal));
  const size_t rb = (size_t)h->d * s.elem_bytes, ub = (size_t)h->d_user * s.elem_bytes;  // stored / user row bytes
  if (X_out && rows) CK(cudaMemcpy2DAsync(X_out, ub, (const unsigned char *)s.X + (size_t)row0 * rb, rb, ub, (size_t)rows, cudaMemcpyDeviceToHost, D.st));
  if (labels_out && rows) CK(cudaMemcpyAsync(labels_out, s.labels + row0, (size_t)rows * sizeof(double), cudaMemcpyDeviceToHost, D.st));
  CK(cudaStreamSynchronize(D.st));
  return 0;
}

int agd_synth_wtrue(agd_handle *h, uint64_t seed, int32_t d, double *w_out) {
  if (!h) return 1;
  Dev &D = h->devs[0];
  CK(cudaSetDevice(D.ordinal));
  double *tmp = nullptr;
  CK(cudaMalloc(&tmp, (size_t)d * sizeof(double)));
  CK(synth_wtrue_launch(tmp, seed, d, D.st));
  CK(cudaMemcpyAsync(w_out, tmp, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, D.st));
  CK(cudaStreamSynchronize(D.st));
  cudaFree(tmp);
  return 0;
}

const char *agd_kernel_name(const agd_handle *h, int32_t dev) {
  if (!h || dev < 0 || dev >= (int)h->devs.size() || h->d <= 0) return "";
  const Shard &s = h->devs[dev].sh;
  if (s.csr) return s.elem_bytes == 8 ? "k1_csr_pipelined_kernel<double>" : "k1_csr_pipelined_kernel<float>";
  const int eb = s.elem_bytes ? s.elem_bytes : 4;
  const char *t = eb == 8 ? "double" : (eb == 4 ? "float" : "__nv_bfloat16");
  static thread_local char buf[96];
  switch (dense_kernel_of(h, eb)) {
    case 3: return "k1_tc_kernel (tcgen05, bf16 storage)";
    case 1: snprintf(buf, sizeof buf, "k1_ring_kernel<%s,...>", t); return buf;
    default: snprintf(buf, sizeof buf, "k1_generic_kernel<%s>", t); return buf;
  }
}

int agd_set_option(agd_handle *h, const char *key, const char *value) {
  if (!h || !key || !value) return 1;
  if (!strcmp(key, "k1_variant")) {
    if (!strcmp(value, "auto")) h->k1_variant = 0;
    else if (!strcmp(value, "ring")) h->k1_variant = 1;
    else if (!strcmp(value, "generic")) h->k1_variant = 2;
    else if (!strcmp(value, "tc")) h->k1_variant = 4;
    else return fail(h, "k1_variant must be auto|ring|generic|tc");
    return 0;
  }
  if (!strcmp(key, "ring_stages")) { h->ring_stages = atoi(value); return 0; }
  if (!strcmp(key, "collective")) {
    if (!strcmp(value, "auto")) h->collective = 0;
    else if (!strcmp(value, "nccl")) h->collective = 1;
    else if (!strcmp(value, "p2p")) h->collective = 2;
    else return fail(h, "collective must be auto|nccl|p2p");
    free_xchg(h);
    return 0;
  }
  if (!strcmp(key, "k1_diag")) { h->k1_diag = atoi(value); return 0; }
  if (!strcmp(key, "tc_margins")) {
    if (!strcmp(value, "f32")) h->tc_margins_f64 = 0;
    else if (!strcmp(value, "f64")) h->tc_margins_f64 = 1;
    else return fail(h, "tc_margins must be f32|f64");
    return 0;
  }
  if (!strcmp(key, "ring_rows")) { h->tune_rows = atoi(value); return 0; }
  if (!strcmp(key, "ring_ctas")) { h->tune_ctas = atoi(value); return 0; }
  if (!strcmp(key, "ring_predicated")) { h->tune_full = atoi(value); return 0; }
  return fail(h, "unknown option %s", key);
}

// ---------------------------------------------------------------- applySmooth with host buffers
static int smooth_host(agd_handle *h, int32_t gradient, const double *w, const double *w2, double *loss, double *grad,
                       int64_t *count, double *loss2, double *grad2 = nullptr) {
  if (check_ready(h)) return 1;
  if (gradient < 0 || gradient > AGD_GRAD_LEAST_SQUARES_HALF) return fail(h, "unknown gradient %d", gradient);
  if (!w || !loss || !grad) return fail(h, "NULL argument");
  if (w2 && !dual_supported(h)) return fail(h, "this shard's gradient kernel has no two-point form (use two agd_smooth calls)");
  if (grad2 && !dual_full_supported(h)) return fail(h, "this shard's gradient kernel has no two-gradient form (use two agd_smooth calls)");
  const int32_t d = h->d;
  for (Dev &D : h->devs) {
    CK(cudaSetDevice(D.ordinal));
    CK(cudaMemsetAsync(D.wtmp, 0, ((size_t)d + 2) * sizeof(double), D.st));                            // zero weights on padded columns
    CK(cudaMemcpyAsync(D.wtmp, w, (size_t)h->d_user * sizeof(double), cudaMemcpyHostToDevice, D.st));  // = broadcast, AGD.scala:193
    if (w2) {
      CK(cudaMemsetAsync(D.g_x, 0, ((size_t)d + 2) * sizeof(double), D.st));   // g_x doubles as the staging vector of w2
      CK(cudaMemcpyAsync(D.g_x, w2, (size_t)h->d_user * sizeof(double), cudaMemcpyHostToDevice, D.st));
    }
  }
  h->devs[0].ev_used = h->devs[0].ev_ar_used = 0;
  h->launches = h->collectives = 0;
  h->xg_pending = false;
  if (smooth_device(h, gradient, [](Dev &D) { return (const double *)D.wtmp; }, false,
                    w2 ? (WSel)[](Dev &D) { return (const double *)D.g_x; } : (WSel) nullptr, grad2 != nullptr))
    return 1;
  Dev &D0 = h->devs[0];
  CK(cudaSetDevice(D0.ordinal));
  std::vector<double> host(2 * ((size_t)d + 4));
  CK(cudaMemcpyAsync(host.data(), D0.acc, (grad2 ? 2 : 1) * ((size_t)d + 4) * sizeof(double), cudaMemcpyDeviceToHost, D0.st));
  for (Dev &D : h->devs) { CK(cudaSetDevice(D.ordinal)); CK(cudaStreamSynchronize(D.st)); }
  const double cnt = host[(size_t)d + 1];
  *loss = host[d] / cnt;                                    // AGD.scala:207
  for (int32_t j = 0; j < h->d_user; ++j) grad[j] = host[j] / cnt;
  if (count) *count = (int64_t)cnt;
  if (w2 && loss2) *loss2 = host[(size_t)d + 2] / host[(size_t)d + 3];
  if (grad2) {  // second block: [grad at w2 | loss | count | 0 | 0]
    const double *b2 = host.data() + (size_t)d + 4;
    for (int32_t j = 0; j < h->d_user; ++j) grad2[j] = b2[j] / b2[(size_t)d + 1];
  }
  return 0;
}

int agd_smooth(agd_handle *h, int32_t gradient, const double *w, double *loss, double *grad, int64_t *count) {
  return smooth_host(h, gradient, w, nullptr, loss, grad, count, nullptr);
}

int agd_smooth_pair(agd_handle *h, int32_t gradient, const double *w, const double *w2, double *loss, double *grad,
                    int64_t *count, double *loss2) {
  if (h && (!w2 || !loss2)) return fail(h, "NULL argument");
  return smooth_host(h, gradient, w, w2, loss, grad, count, loss2);
}

int agd_smooth_two(agd_handle *h, int32_t gradient, const double *w, const double *w2, double *loss, double *grad,
                   int64_t *count, double *loss2, double *grad2) {
  if (h && (!w2 || !loss2 || !grad2)) return fail(h, "NULL argument");
  return smooth_host(h, gradient, w, w2, loss, grad, count, loss2, grad2);
}

// ---------------------------------------------------------------- applyProjector with host buffers
int agd_prox(agd_handle *h, int32_t updater, const double *w, const double *g, double step, double reg,
             int32_t d, double *w_out, double *reg_val) {
  if (!h) return 1;
  if (updater < 0 || updater > AGD_UPD_L1) return fail(h, "unknown updater %d", updater);
  if (d <= 0) return fail(h, "bad dimension");
  Dev &D = h->devs[0];
  CK(cudaSetDevice(D.ordinal));
  double *buf = nullptr, *partials = nullptr;
  CK(cudaMalloc(&buf, 3 * (size_t)d * sizeof(double)));
  CK(cudaMalloc(&partials, (size_t)k3_blocks(d) * K3_NS * sizeof(double)));
  CK(cudaMemcpyAsync(buf, w, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, D.st));
  CK(cudaMemcpyAsync(buf + d, g, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, D.st));
  K3ProxArgs a;
  a.w = buf; a.g = buf + d; a.w_out = buf + 2 * (size_t)d; a.partials = partials; a.ticket = D.ticket;
  a.scalars = D.scalars_dev; a.step = step; a.reg = reg; a.acc_tail = nullptr; a.d = d; a.updater = updater;
  CK(k3_prox_launch(a, D.st));
  CK(cudaMemcpyAsync(w_out, buf + 2 * (size_t)d, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, D.st));
  double sc[K3_NS];
  if (read_scalars(h, sc)) return 1;
  if (reg_val) *reg_val = reg_value(updater, reg, sc[2], sc[5]);
  cudaFree(buf);
  cudaFree(partials);
  return 0;
}

// ---------------------------------------------------------------- AcceleratedGradientDescent.run
int agd_run(agd_handle *h, const agd_params *p, const double *w0, double *w_out, double *loss_hist,
            int32_t *n_hist, agd_stats *stats) {
  if (check_ready(h)) return 1;
  if (!p || !w0 || !w_out || !loss_hist || !n_hist) return fail(h, "NULL argument");
  if (p->gradient < 0 || p->gradient > AGD_GRAD_LEAST_SQUARES_HALF) return fail(h, "unknown gradient %d", p->gradient);
  if (p->updater < 0 || p->updater > AGD_UPD_L1) return fail(h, "unknown updater %d", p->updater);
  const auto t_begin = std::chrono::steady_clock::now();
  const int32_t d = h->d;
  const size_t vb = (size_t)d * sizeof(double);
  const double INF = std::numeric_limits<double>::infinity();
  agd_stats s;
  memset(&s, 0, sizeof s);
  if (call_begin(h)) return 1;
  const bool memoize = (p->flags & AGD_FLAG_MEMOIZE_FX) != 0;
  // Pass fusion: the history evaluation applySmooth(x) (:304) of iteration k and applySmooth(y) (:250) of iteration k+1
  // (first backtracking round) do not depend on each other, so ONE sweep over X evaluates both.
  const bool fuse = (p->flags & AGD_FLAG_NO_FUSE) == 0 && dual_supported(h);
  bool y_ready = false;   // acc already holds applySmooth(y) for the first round of the coming iteration
  // Speculative sweep of the memoised pass structure: applySmooth(x) of the backtracking test (:269) and applySmooth(y) of the
  // NEXT iteration (:250) -- y_{k+1} = x_k (1 - theta') + z_k theta' with theta' from L alpha, i.e. assuming the test accepts and
  // the gradient test does not restart -- are evaluated by ONE two-gradient sweep over X.  An accepted iteration then reads X
  // once.  Every evaluation is the arithmetic a sweep of its own would do (same kernel code path), so weights and history
  // stay bit-identical; a rejected guess only wastes the extra FMAs.  On a restart y_{k+1} = x_k exactly, and the memoised
  // (f_x, g_x) are reused without any evaluation at all.
  const bool reuse_fx = memoize && (p->flags & AGD_FLAG_NO_FUSE) == 0 && p->beta < 1.0;   // restart: applySmooth(y_{k+1}) = (f_x, g_x)
  const bool speculate_y = reuse_fx && dual_full_supported(h);
  size_t acc_off = 0;     // where applySmooth(y) of the current round lives inside Dev::acc (0, or d + 4 after a good guess)

  for (Dev &D : h->devs) {                                                 // :224-225  x = w0 ; z = x
    CK(cudaSetDevice(D.ordinal));
    CK(cudaMemsetAsync(D.x, 0, vb, D.st));  // padded columns carry zero weights throughout
    CK(cudaMemcpyAsync(D.x, w0, (size_t)h->d_user * sizeof(double), cudaMemcpyHostToDevice, D.st));
    CK(k3_copy2_launch(D.z, D.x, nullptr, nullptr, d, D.st));
  }
  h->launches += 1;
  double theta = INF;                                                      // :226
  int nh = 0;                                                              // :227
  double f_y = 0.0;                                                        // :229
  double L = p->L0;                                                        // :232
  bool backtrack_simple = true;                                            // :234
  const double backtrack_tol = 1e-10;                                      // :235
  const double Lexact = p->Lexact, beta = p->beta;
  double sc[K3_NS] = {0}, sg[K3_NS] = {0};
  unsigned long long &round_seq = h->seq_base;  // every round of every call gets a fresh sequence number

  auto launch_all = [&](auto fn) -> int {
    for (Dev &D : h->devs) {
      CK(cudaSetDevice(D.ordinal));
      CK(fn(D));
    }
    trace_mark(h, "k3");
    h->launches += 1;
    return 0;
  };
  trace_mark(h, "start");

  // Host round trips: the host needs device scalars once per backtracking round.  Pass 2 (applySmooth(x), :269) is
  // enqueued speculatively right behind pass 1 -- it is wasted only when ||x - y||^2 == 0 (:265) -- and the f_x of the
  // history pass (:304) is copied to pinned memory asynchronously and read after the loop, so the GPU runs
  // pass 3(k) -> pass 1(k+1) -> pass 2(k+1) back to back with a single synchronisation per round.
  Dev &H0 = h->devs[0];
  if (H0.hist_cap < 2 * (size_t)(p->num_iterations > 0 ? p->num_iterations : 1)) {
    CK(cudaSetDevice(H0.ordinal));
    if (H0.hist_host) cudaFreeHost(H0.hist_host);
    H0.hist_cap = 2 * (size_t)(p->num_iterations > 0 ? p->num_iterations : 1);
    CK(cudaHostAlloc(&H0.hist_host, H0.hist_cap * sizeof(double), cudaHostAllocMapped | cudaHostAllocPortable));
    CK(cudaHostGetDevicePointer((void **)&H0.hist_dev, H0.hist_host, 0));
  }
  // the gather of a sweep is done by the K3 kernel that consumes its sums (one launch fewer per sweep)
  auto take_gather = [&](Dev &D) {
    XchgGather g;
    if (h->xg_pending) {
      const int S = 2 * (d + 4), W = h->world;
      g.world = W; g.buf = (int)(h->xg_epoch & 1ull); g.n = h->xg_n; g.slot_stride = S; g.epoch = h->xg_epoch;
      g.rs = h->xg_rs ? 1 : 0;
      g.xbuf = h->xg_rs ? D.xbuf + xchg_off_res(S, W) : D.xbuf;       // rs: the area of finished sums
      g.flags = h->xg_rs ? D.xflags + 4 * W : D.xflags;               // rs: the "finished slice arrived" flags
    }
    return g;
  };
  long long pending_hist = -1;        // slot of H0.hist_host the next k3_step fills from the fused sweep it consumes
  std::vector<double> cx_of;          // c_x per iteration (:305)
  std::vector<char> fx_deferred;      // f_x of iteration k still sits in H0.hist_host[2k..2k+1]
  for (int nIter = 1; nIter <= p->num_iterations; ++nIter) {              // :237
    const double L_old = L;                                                // :242
    L = L * p->alpha;                                                      // :243
    const double theta_old = theta;                                        // :244
    bool nonterminating = false, have_fx = false, first_round = true;
    bool guess_live = false;     // acc[d+4 ..] holds applySmooth at y_spec, sharing its sweep with this round's applySmooth(x)
    double guess_theta = 0.0;
    double f_x = 0.0;
    for (;;) {                                                             // :246
      theta = 2.0 / (1.0 + std::sqrt(1.0 + 4.0 * (L / L_old) / (theta_old * theta_old)));  // :248
      const double omt = 1.0 - theta;
      if (first_round) {  // (x_old, z_old) = (x, z) :241 fused with y = x_old*(1-theta) + z_old*theta :249
        if (!y_ready && launch_all([&](Dev &D) { return k3_begin_launch(D.x_old, D.z_old, D.y, D.x, D.z, omt, theta, d, D.st); })) return 1;
        first_round = false;
      } else if (launch_all([&](Dev &D) { return k3_combine_launch(D.y, D.x_old, omt, D.z_old, theta, d, D.st); })) return 1;  // :249
      if (!y_ready) {
        if (smooth_device(h, p->gradient, [](Dev &D) { return (const double *)D.y; }, true, nullptr, false, true)) return 1;  // :250
        acc_off = 0;
      }
      y_ready = false;
      s.passes++;
      const double step = 1.0 / (theta * L);                               // :253
      const bool speculate = beta < 1.0;                                   // :257 is known up front
      // this round's applySmooth(x) sweep also carries a guess of y_{k+1} (formed inside k3_step from the new x and z)
      const bool guessed = speculate && speculate_y && nIter < p->num_iterations;
      double theta_guess = 0.0;
      if (guessed) {
        const double L_n = L * p->alpha;                                   // :243-248 of iteration nIter + 1 if this round accepts
        theta_guess = 2.0 / (1.0 + std::sqrt(1.0 + 4.0 * (L_n / L) / (theta * theta)));
      }
      if (launch_all([&](Dev &D) {                                         // :254-255,263-264 fused
            K3StepArgs a;
            a.y_spec = guessed ? D.y_spec : nullptr; a.spec_ca = 1.0 - theta_guess; a.spec_cb = theta_guess;
            a.acc = D.acc + acc_off; a.x_old = D.x_old; a.z_old = D.z_old; a.y = D.y; a.g_y = D.g_y; a.z = D.z; a.x = D.x;
            a.partials = D.partials; a.ticket = D.ticket; a.scalars = D.scalars_dev;
            a.theta = theta; a.one_minus_theta = omt; a.step = step; a.reg = p->reg_param; a.d = d; a.updater = p->updater;
            if (!speculate) { a.seq_out = reinterpret_cast<unsigned long long *>(D.scalars_dev + 2 * K3_NS); a.seq = round_seq + 1; }
            a.xg = take_gather(D); a.acc_w = D.acc;
            a.hist_out = (pending_hist >= 0 && &D == &h->devs[0]) ? D.hist_dev + pending_hist : nullptr;
            return k3_step_launch(a, D.st);
          })) return 1;
      h->xg_pending = false;
      pending_hist = -1;
      if (speculate) {                                                     // :269, enqueued before :265 is known
        if (guessed) {
          if (smooth_device(h, p->gradient, [](Dev &D) { return (const double *)D.x; }, true,
                            [](Dev &D) { return (const double *)D.y_spec; }, true, true))
            return 1;
        } else if (smooth_device(h, p->gradient, [](Dev &D) { return (const double *)D.x; }, true, nullptr, false, true)) return 1;
        if (launch_all([&](Dev &D) {
              K3GxArgs a;
              a.acc = D.acc; a.x = D.x; a.y = D.y; a.g_y = D.g_y; a.g_x = D.g_x; a.partials = D.partials;
              a.ticket = D.ticket; a.scalars = D.scalars_dev + K3_NS; a.d = d;
              a.seq_out = reinterpret_cast<unsigned long long *>(D.scalars_dev + 2 * K3_NS); a.seq = round_seq + 1;
              a.xg = take_gather(D); a.acc_w = D.acc;
              return k3_gx_launch(a, D.st);
            })) return 1;
        h->xg_pending = false;
      }
      if (wait_scalars(h, ++round_seq, sc)) return 1;                      // the one host wait of this round (no stream drain)
      trace_mark(h, "host-gap");
      memcpy(sg, H0.scalars_host + K3_NS, K3_NS * sizeof(double));
      f_y = sc[6] / sc[7];                                                 // :207
      have_fx = false;
      if (beta >= 1.0) break;                                              // :257
      const double nxy = std::sqrt(sc[0]);
      const double xy_sq = nxy * nxy;                                      // :264  math.pow(norm(xy), 2)
      if (xy_sq == 0) { s.wasted_passes++; guess_live = false; break; }    // :265  (the speculative pass is discarded)
      s.passes++;
      f_x = sg[6] / sg[7];
      have_fx = true;
      guess_live = guessed;                                                // valid only if this round is the accepted one
      guess_theta = theta_guess;
      double localL;
      if (backtrack_simple) {                                              // :272
        const double q_x = f_y + sc[1] + 0.5 * L * xy_sq;                  // :273
        localL = L + 2.0 * jmax(f_x - q_x, 0.0) / xy_sq;                   // :274
        backtrack_simple = (std::fabs(f_y - f_x) >= backtrack_tol * jmax(std::fabs(f_x), std::fabs(f_y)));  // :275
      } else {
        localL = 2.0 * sg[0] / xy_sq;                                      // :278
      }
      if (localL <= L || L >= Lexact) break;                               // :281
      if (!std::isinf(localL)) L = jmin(Lexact, localL);                   // :285-287
      else localL = L;                                                     // :288-290
      L = jmin(Lexact, jmax(localL, L / beta));                            // :292
      s.backtracks++;
      if (guess_live) { s.wasted_passes++; guess_live = false; }   // rejected: y_{k+1} was guessed from an L that did not survive
      if (L != L) { nonterminating = true; break; }  // the reference never leaves :246-293 once L is NaN
    }
    const double c_x = reg_value(p->updater, p->reg_param, sc[2], sc[5]);  // :305  applyProjector(x, g_x, 0.0)._1
    cx_of.push_back(c_x);
    s.iterations = nIter;
    // The exits of :309-324 and the restart of :327-331 depend on nothing the history evaluation (:304) produces, so they
    // are decided first: when another iteration follows, that evaluation shares its sweep with the next applySmooth(y).
    bool stop = false;
    if (nonterminating) { s.stopped_nan = 1; s.nonterminating = 1; stop = true; }
    else if (std::isnan(f_y) || std::isinf(f_y)) { s.stopped_nan = 1; stop = true; }  // :309-312
    else {
      const double norm_x = std::sqrt(sc[2]);                              // :315
      const double norm_dx = std::sqrt(sc[3]);                             // :316
      if (norm_dx == 0.0 && nIter > 1) { s.converged = 1; stop = true; }   // :317-321
      else if (norm_dx < p->convergence_tol * jmax(norm_x, 1)) { s.converged = 1; stop = true; }  // :322-324
    }
    const bool restart = !stop && p->may_restart && sc[4] > 0.0;           // :327
    const bool need_hist = !(memoize && have_fx);
    const bool fuse_now = need_hist && fuse && !stop && nIter < p->num_iterations;
    if (need_hist && !fuse_now) {                                          // :304  (f_x, g_x) = applySmooth(x)
      if (smooth_device(h, p->gradient, [](Dev &D) { return (const double *)D.x; }, true)) return 1;
      s.passes++;
      CK(cudaSetDevice(H0.ordinal));
      CK(cudaMemcpyAsync(H0.hist_host + 2 * (size_t)nh, H0.acc + d, 2 * sizeof(double), cudaMemcpyDeviceToHost, H0.st));
      fx_deferred.push_back(1);                                            // read after the loop
      loss_hist[nh++] = 0.0;
    } else if (!need_hist) {
      fx_deferred.push_back(0);
      loss_hist[nh++] = f_x + c_x;                                         // :306
    }
    if (restart) {
      if (launch_all([&](Dev &D) { return k3_copy2_launch(D.z, D.x, nullptr, nullptr, d, D.st); })) return 1;  // :328
      theta = INF;                                                         // :329
      backtrack_simple = true;                                             // :330
      s.restarts++;
    }
    if (reuse_fx && have_fx && !stop && nIter < p->num_iterations && (guess_live || restart)) {
      // The coming iteration's first round is already evaluated.  (x_old, z_old) = (x, z) and y (:241,:249) are formed as the
      // loop head would form them -- the same kernel, the same doubles -- and pass 1 is skipped (y_ready):
      //   * no restart: y_{k+1} is the guess; its sums are the second block of acc;
      //   * restart: theta' = 1, so y_{k+1} = x_k * 0 + z * 1 with z = x_k (:328) = x_k exactly; applySmooth(x_k) is the first block.
      const double L_n = L * p->alpha;
      const double theta_n = 2.0 / (1.0 + std::sqrt(1.0 + 4.0 * (L_n / L) / (theta * theta)));
      if (restart || theta_n == guess_theta) {
        const double omt_n = 1.0 - theta_n;
        if (launch_all([&](Dev &D) { return k3_begin_launch(D.x_old, D.z_old, D.y, D.x, D.z, omt_n, theta_n, d, D.st); })) return 1;
        acc_off = restart ? 0 : (size_t)d + 4;
        if (restart && guess_live) s.wasted_passes++;
        if (!restart) s.fused_passes++;      // pass 1 of iteration nIter + 1 shared its sweep with pass 2 of this one
        y_ready = true;
      } else if (guess_live) s.wasted_passes++;
    } else if (guess_live) s.wasted_passes++;
    if (fuse_now) {
      // :243-249 of iteration nIter + 1, first round (the loop head recomputes the same doubles), then one sweep:
      // acc[0..d+1] = applySmooth(y) sums, acc[d+2..d+3] = loss sum / count at x
      const double L_n = L * p->alpha;
      const double theta_n = 2.0 / (1.0 + std::sqrt(1.0 + 4.0 * (L_n / L) / (theta * theta)));
      const double omt_n = 1.0 - theta_n;
      if (launch_all([&](Dev &D) { return k3_begin_launch(D.x_old, D.z_old, D.y, D.x, D.z, omt_n, theta_n, d, D.st); })) return 1;
      if (smooth_device(h, p->gradient, [](Dev &D) { return (const double *)D.y; }, true,
                        [](Dev &D) { return (const double *)D.x; }, false, true))
        return 1;
      s.passes++;          // the history evaluation; the applySmooth(y) half is counted by the next iteration
      s.fused_passes++;
      y_ready = true;
      pending_hist = 2 * (long long)nh;   // the k3_step that consumes this sweep stores {loss sum, count} at x into the history slot
      fx_deferred.push_back(1);
      loss_hist[nh++] = 0.0;
    }
    if (stop) break;
  }
  {
    Dev &D = h->devs[0];
    CK(cudaSetDevice(D.ordinal));
    CK(cudaMemcpyAsync(w_out, D.x, (size_t)h->d_user * sizeof(double), cudaMemcpyDeviceToHost, D.st));    // :337
  }
  if (h->xg_pending || pending_hist >= 0) return fail(h, "internal: a sweep was left without its consumer");
  if (call_end(h, s, t_begin)) return 1;
  trace_report(h, memoize ? "agd_run memoised" : (fuse ? "agd_run" : "agd_run unfused"));
  for (int k = 0; k < nh; ++k)                                             // :306 for the deferred f_x values
    if (fx_deferred[(size_t)k]) loss_hist[k] = H0.hist_host[2 * (size_t)k] / H0.hist_host[2 * (size_t)k + 1] + cx_of[(size_t)k];
  *n_hist = nh;
  s.final_L = L;
  s.final_theta = theta;
  if (stats) *stats = s;
  return 0;
}

// ---------------------------------------------------------------- GradientDescent.runMiniBatchSGD (fraction 1.0)
int agd_gd_run(agd_handle *h, int32_t gradient, int32_t updater, double step_size, int32_t num_iterations,
               double reg_param, const double *w0, double *w_out, double *loss_hist, int32_t *n_hist,
               agd_stats *stats) {
  return agd_gd_run_minibatch(h, gradient, updater, step_size, num_iterations, reg_param, 1.0, w0, w_out, loss_hist,
                              n_hist, stats);
}

int agd_gd_run_minibatch(agd_handle *h, int32_t gradient, int32_t updater, double step_size, int32_t num_iterations,
                         double reg_param, double mini_batch_fraction, const double *w0, double *w_out,
                         double *loss_hist, int32_t *n_hist, agd_stats *stats) {
  if (check_ready(h)) return 1;
  if (!(mini_batch_fraction > 0.0)) return fail(h, "miniBatchFraction must be positive");
  if (gradient < 0 || gradient > AGD_GRAD_LEAST_SQUARES_HALF) return fail(h, "unknown gradient %d", gradient);
  if (updater < 0 || updater > AGD_UPD_L1) return fail(h, "unknown updater %d", updater);
  const auto t_begin = std::chrono::steady_clock::now();
  const int32_t d = h->d;
  const size_t vb = (size_t)d * sizeof(double);
  agd_stats s;
  memset(&s, 0, sizeof s);
  if (call_begin(h)) return 1;
  double sc[K3_NS];
  int64_t total_rows_local = 0;
  for (Dev &D : h->devs) total_rows_local += D.sh.rows;
  auto prox_all = [&](const double *acc_tail_sel, double step, bool from_acc) -> int {
    for (Dev &D : h->devs) {
      CK(cudaSetDevice(D.ordinal));
      K3ProxArgs a;
      a.w = D.x; a.g = from_acc ? D.acc : D.g_x; a.w_out = D.z; a.partials = D.partials; a.ticket = D.ticket;
      a.scalars = D.scalars_dev; a.step = step; a.reg = reg_param; a.acc_tail = from_acc ? D.acc + d : nullptr;
      a.d = d; a.updater = updater;
      CK(k3_prox_launch(a, D.st));
      CK(k3_copy2_launch(D.x, D.z, nullptr, nullptr, d, D.st));
    }
    (void)acc_tail_sel;
    h->launches += 2;
    return 0;
  };
  for (Dev &D : h->devs) {
    CK(cudaSetDevice(D.ordinal));
    CK(cudaMemsetAsync(D.x, 0, vb, D.st));  // padded columns carry zero weights throughout
    CK(cudaMemcpyAsync(D.x, w0, (size_t)h->d_user * sizeof(double), cudaMemcpyHostToDevice, D.st));
    CK(cudaMemsetAsync(D.g_x, 0, vb, D.st));
  }
  // regVal = updater.compute(weights, zeros, 0, 1, regParam)._2
  if (prox_all(nullptr, 0.0, false)) return 1;
  if (read_scalars(h, sc)) return 1;
  double reg_val = reg_value(updater, reg_param, sc[2], sc[5]);
  int nh = 0;
  for (int i = 1; i <= num_iterations; ++i) {
    // data.sample(false, miniBatchFraction, 42 + i): Bernoulli row mask keyed by the iteration
    h->sample_seed = 42ull + (unsigned long long)i;
    h->sample_thresh = mini_batch_fraction >= 1.0 ? 0ull : (unsigned long long)std::ldexp(mini_batch_fraction, 64);
    const int rc_smooth = smooth_device(h, gradient, [](Dev &D) { return (const double *)D.x; }, true);
    h->sample_thresh = 0ull;
    if (rc_smooth) return 1;
    s.passes++;
    const double this_step = step_size / std::sqrt((double)i);
    if (prox_all(nullptr, this_step, true)) return 1;
    if (read_scalars(h, sc)) return 1;
    if (!(sc[7] > 0)) continue;  // miniBatchSize == 0: the reference logs a warning and skips the update
    loss_hist[nh++] = sc[6] / sc[7] + reg_val;
    reg_val = reg_value(updater, reg_param, sc[2], sc[5]);
    s.iterations = i;
  }
  {
    Dev &D = h->devs[0];
    CK(cudaSetDevice(D.ordinal));
    CK(cudaMemcpyAsync(w_out, D.x, (size_t)h->d_user * sizeof(double), cudaMemcpyDeviceToHost, D.st));
  }
  if (call_end(h, s, t_begin)) return 1;
  *n_hist = nh;
  if (stats) *stats = s;
  return 0;
}

}  // extern "C"
