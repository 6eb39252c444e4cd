// lq_capi.cu -- context, kernel launchers and the extern "C" surface declared in include/lqcd_b200.h.
//
// Default build: CUDA for sm_100a (nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU path in that
// library: without a device lq_ctx_create returns LQ_E_NODEVICE.
// -DLQ_HOST_EMU (g++ -x c++): the same kernel bodies in host loops -- test infrastructure for the CPU CI only
// (tests/emu.py); the package never loads it.
#include "../../include/lqcd_b200.h"
#include "lq_kernels.cuh"
#include "lq_geom_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#if !defined(LQ_HOST_EMU) && defined(LQ_HAVE_TUNED)
#include "lq_tuned.cuh"
#define LQ_TUNED 1
#endif

#define LQ_BLOCK 128
#define LQ_RBLOCK 256
#define LQ_P2P_NBUF 9      /* U U2 E E2 G G2 T T2 flags */
#define LQ_P2P_MAXNB 8     /* 3^2 - 1 neighbours of a 2-D process grid */
#define LQ_P2P_HANDLE 64   /* sizeof(cudaIpcMemHandle_t) */
#define LQ_PROF_CAP 8192     /* event pairs in flight; a full ring is drained into the per-class totals */
#define LQ_PROF_NCLASS 16

static thread_local char g_cuda_err[512] = "";

// ------------------------------------------------------------------------------------------------ runtime shim
#ifdef LQ_HOST_EMU
typedef void* lq_stream_t;
#define LQ_CHECK(x) \
  do {              \
  } while (0)
static int rt_malloc(void** p, size_t n) {
  *p = calloc(n ? n : 1, 1);
  return *p ? LQ_OK : LQ_E_CUDA;
}
static void rt_free(void* p) { free(p); }
static int rt_malloc_host(void** p, size_t n) { return rt_malloc(p, n); }
static void rt_free_host(void* p) { free(p); }
static int rt_copy(void* d, const void* s, size_t n, int /*kind*/, lq_stream_t) {
  memcpy(d, s, n);
  return LQ_OK;
}
static int rt_memset(void* d, int v, size_t n, lq_stream_t) {
  memset(d, v, n);
  return LQ_OK;
}
static int rt_sync(lq_stream_t) { return LQ_OK; }
enum { H2D = 0, D2H = 1, D2D = 2 };
#else
typedef cudaStream_t lq_stream_t;
static int cuda_fail(cudaError_t e, const char* what, int line) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s failed at lq_capi.cu:%d: %s", what, line, cudaGetErrorString(e));
  return LQ_E_CUDA;
}
#define LQ_CHECK(x)                                              \
  do {                                                           \
    cudaError_t e_ = (x);                                        \
    if (e_ != cudaSuccess) return cuda_fail(e_, #x, __LINE__);   \
  } while (0)
static int rt_malloc(void** p, size_t n) {
  LQ_CHECK(cudaMalloc(p, n ? n : 1));
  return LQ_OK;
}
static void rt_free(void* p) {
  if (p) cudaFree(p);
}
static int rt_malloc_host(void** p, size_t n) {
  LQ_CHECK(cudaMallocHost(p, n));
  return LQ_OK;
}
static void rt_free_host(void* p) {
  if (p) cudaFreeHost(p);
}
enum { H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost, D2D = cudaMemcpyDeviceToDevice };
static int rt_copy(void* d, const void* s, size_t n, int kind, lq_stream_t st) {
  LQ_CHECK(cudaMemcpyAsync(d, s, n, (cudaMemcpyKind)kind, st));
  return LQ_OK;
}
static int rt_memset(void* d, int v, size_t n, lq_stream_t st) {
  LQ_CHECK(cudaMemsetAsync(d, v, n, st));
  return LQ_OK;
}
static int rt_sync(lq_stream_t st) {
  LQ_CHECK(cudaStreamSynchronize(st));
  return LQ_OK;
}
#endif

// ------------------------------------------------------------------------------------------------ context
struct lq_ctx {
  int device;
  lq_stream_t stream;
  bool own_stream;
  LqGeom g;
  double a, beta, CA;
  int flags;
  bool g_pp_safe;  // the last Gauss-field evaluation was the two-pass peer-memory one (see gauss_field)
  int integ_kind, integ_exp;  // lq_set_integrator: what lq_md_n / lq_hmc_trajectory run (0, 0 = the reference's)
  double integ_lambda;
  int64_t t;
  int64_t launches;
  bool decomposed;
  int nproc[LQ_MAXD];
  bool even_extents;
  int odd_mask;  // bit d: ext[d] is odd (sweeps then use the colour classes of lq_site_class); 0 when all are even
  cx *U, *U2, *E, *E2, *G, *G2;
  cx *T, *T2;  // transported field U^+ E U of the projection loop (lq_gausst4_kernel), D = 4 tuned path only
  cx *snapU, *snapE;
  cx* hmcU;  // reject-path copy of lq_hmc_trajectory (its own buffer: lq_snapshot / lq_restore keep theirs)
  int64_t snap_t;
  bool has_snap;
  double* d_aos;
  size_t d_aos_bytes;
  double* d_partial;
  size_t partial_cap;  // doubles
  double* d_result;
  double* h_result;
  bool halo_ok[3];
  bool g_valid;  // the Gauss field in G matches the current (U, E)
  lq_comm comm;
  bool has_comm;
  bool reduced_globally;  // the last reduce() result in h_result is already the global sum
  // peer-to-peer halo transport (CUDA IPC mappings of the neighbours' buffers, written over NVLink by our kernels)
  bool p2p_on;
  cx* own[LQ_P2P_NBUF - 1];                       // the field allocations in export order: U U2 E E2 G G2 T T2
  unsigned long long* p2p_flags;                  // [LQ_P2P_MAXNB] incoming flags + [LQ_P2P_MAXNB] error word
  unsigned long long p2p_epoch;
  int p2p_nnb, p2p_npeers;
  int p2p_off[LQ_P2P_MAXNB][LQ_MAXD];             // neighbour offsets (-1, 0, +1 per direction)
  int p2p_peer[LQ_P2P_MAXNB];                     // neighbour -> opened peer
  int p2p_rev[LQ_P2P_MAXNB];                      // my slot in that neighbour's flag array
  void* p2p_base[LQ_P2P_MAXNB][LQ_P2P_NBUF];      // opened peer buffers (per unique peer)
  int64_t p2p_exchanges;
  void* d_push;                                   // LqPush[8] on the device: peer table per field-buffer allocation
  // pipelined marshalling (lq_links_upload_begin / _commit, lq_links_download_begin, lq_copies_wait)
  double *stage_in, *stage_out;  // AoS staging buffers on the device
  bool up_pending;
#ifndef LQ_HOST_EMU
  cudaStream_t s_h2d, s_d2h;
  cudaEvent_t ev_in_ready, ev_in_consumed, ev_out_ready, ev_out_done;
  bool copy_streams;
#endif
  unsigned int* d_fold_counter;                   // boundary blocks finished (halo synchronisation folded into the kernels)
  bool p2p_pending;                               // the last folded launch released an epoch nobody has acquired yet
  // optional per-kernel-class CUDA-event timing (lq_profile_*)
  bool prof_on;
  int prof_n;                 // event pairs recorded since the last reset
#ifndef LQ_HOST_EMU
  cudaEvent_t* prof_ev;       // 2 * LQ_PROF_CAP events
#endif
  int prof_class[LQ_PROF_CAP];
  double prof_ms[LQ_PROF_NCLASS];     // per-class totals of the pairs already drained
  int64_t prof_cnt[LQ_PROF_NCLASS];
  size_t u_bytes() const { return (size_t)g.nchunk * 32 * 9 * g.D * sizeof(cx); }
  size_t e_bytes() const { return (size_t)g.nchunk * 32 * 4 * g.D * sizeof(cx); }
  size_t g_bytes() const { return (size_t)g.nchunk * 32 * 9 * sizeof(cx); }
  size_t t_bytes() const { return (size_t)g.nchunk * 32 * 5 * g.D * sizeof(cx); }
};

#ifndef LQ_HOST_EMU
struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) {
    ok = cudaGetDevice(&prev) == cudaSuccess;
    if (ok && prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (ok) cudaSetDevice(prev);
  }
};
#define LQ_GUARD(c) DeviceGuard guard_((c)->device)
#else
#define LQ_GUARD(c) \
  do {              \
  } while (0)
#endif

// Folds the recorded event pairs into the per-class totals (synchronises the stream) and empties the ring.
static void prof_drain(lq_ctx* c) {
#ifndef LQ_HOST_EMU
  if (!c->prof_ev || c->prof_n == 0) return;
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < c->prof_n; ++i) {
    float ms = 0.f;
    const int cls = c->prof_class[i];
    if (cls >= 0 && cls < LQ_PROF_NCLASS && cudaEventElapsedTime(&ms, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]) == cudaSuccess) {
      c->prof_ms[cls] += ms;
      c->prof_cnt[cls] += 1;
    }
  }
  c->prof_n = 0;
#else
  (void)c;
#endif
}
// Times the launches issued while it is alive (one CUDA-event pair on the context stream) when profiling is on.
struct ProfScope {
  lq_ctx* c;
  int slot;
  ProfScope(lq_ctx* c_, int cls) : c(c_), slot(-1) {
#ifndef LQ_HOST_EMU
    if (c->prof_on && c->prof_ev) {
      if (c->prof_n >= LQ_PROF_CAP) prof_drain(c);  // one stream synchronisation per LQ_PROF_CAP timed launches
      slot = c->prof_n++;
      c->prof_class[slot] = cls;
      cudaEventRecord(c->prof_ev[2 * slot], c->stream);
    }
#else
    (void)cls;
#endif
  }
  ~ProfScope() {
#ifndef LQ_HOST_EMU
    if (slot >= 0) cudaEventRecord(c->prof_ev[2 * slot + 1], c->stream);
#endif
  }
};

// ------------------------------------------------------------------------------------------------ launchers
#ifdef LQ_HOST_EMU
template <class F>
static int launch(lq_ctx* c, lq_i64 n, const F& f) {
  c->launches++;
#pragma omp parallel for schedule(static)
  for (lq_i64 i = 0; i < n; ++i) f(i);
  return LQ_OK;
}
// result lands in c->h_result[0..K)
template <class F>
static int reduce(lq_ctx* c, lq_i64 n, const F& f) {
  c->launches += 2;
  double acc[F::K];
  for (int k = 0; k < F::K; ++k) acc[k] = 0.0;
  for (lq_i64 b = 0; b < n; b += LQ_RBLOCK) {
    double v[F::K];
    for (int k = 0; k < F::K; ++k) v[k] = 0.0;
    for (lq_i64 i = b; i < n && i < b + LQ_RBLOCK; ++i) f(i, v);
    for (int k = 0; k < F::K; ++k) acc[k] += v[k];
  }
  for (int k = 0; k < F::K; ++k) c->h_result[k] = acc[k];
  c->reduced_globally = false;
  if (c->decomposed && c->has_comm && c->comm.allreduce_sum_device) {
    if (c->comm.allreduce_sum_device(c->comm.user, c->h_result, F::K)) return LQ_E_COMM;
    c->reduced_globally = true;
  }
  return LQ_OK;
}
#else
template <class F>
__global__ void __launch_bounds__(LQ_BLOCK) lq_k(F f, lq_i64 n) {
  lq_i64 i = (lq_i64)blockIdx.x * LQ_BLOCK + threadIdx.x;
  if (i < n) f(i);
}
template <class F>
__global__ void __launch_bounds__(LQ_RBLOCK) lq_reduce_k(F f, lq_i64 n, double* partial) {
  constexpr int K = F::K;
  double v[K];
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = 0.0;
  lq_i64 i = (lq_i64)blockIdx.x * LQ_RBLOCK + threadIdx.x;
  if (i < n) f(i, v);
  __shared__ double sm[K][LQ_RBLOCK / 32];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[k][w] = x;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    double x = 0.0;
#pragma unroll
    for (int j = 0; j < LQ_RBLOCK / 32; ++j) x += sm[threadIdx.x][j];
    partial[(lq_i64)blockIdx.x * K + threadIdx.x] = x;
  }
}
// fixed-order final sum of the per-block partials (deterministic: no atomics)
template <int K>
__global__ void __launch_bounds__(LQ_RBLOCK) lq_final_k(const double* partial, lq_i64 nblk, double* out) {
  __shared__ double sm[K][LQ_RBLOCK];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double x = 0.0;
    for (lq_i64 b = threadIdx.x; b < nblk; b += LQ_RBLOCK) x += partial[b * K + k];
    sm[k][threadIdx.x] = x;
  }
  __syncthreads();
  for (int o = LQ_RBLOCK / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
#pragma unroll
      for (int k = 0; k < K; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < K) out[threadIdx.x] = sm[threadIdx.x][0];
}
template <class F>
static int launch(lq_ctx* c, lq_i64 n, const F& f) {
  if (n <= 0) return LQ_OK;
  lq_i64 nb = (n + LQ_BLOCK - 1) / LQ_BLOCK;
  lq_k<F><<<(unsigned)nb, LQ_BLOCK, 0, c->stream>>>(f, n);
  c->launches++;
  LQ_CHECK(cudaGetLastError());
  return LQ_OK;
}
static int reduce_reserve(lq_ctx* c, lq_i64 nb, int K) {
  if ((size_t)(nb * K) > c->partial_cap) {
    rt_free(c->d_partial);
    c->d_partial = nullptr;
    int rc = rt_malloc((void**)&c->d_partial, (size_t)nb * K * sizeof(double));
    if (rc) return rc;
    c->partial_cap = (size_t)nb * K;
  }
  return LQ_OK;
}
// second stage of every reduction: fixed-order sum of the nb per-block partials, optional device-side all-reduce,
// one host synchronisation; result in c->h_result[0..K)
template <int K>
static int reduce_finish(lq_ctx* c, lq_i64 nb) {
  lq_final_k<K><<<1, LQ_RBLOCK, 0, c->stream>>>(c->d_partial, nb, c->d_result);
  c->launches += 2;
  LQ_CHECK(cudaGetLastError());
  // decomposed contexts: the global sum is taken on the DEVICE buffer, stream-ordered (one NCCL all-reduce enqueued by
  // the caller's plumbing), so a reduction costs one host synchronisation in total
  c->reduced_globally = false;
  if (c->decomposed && c->has_comm && c->comm.allreduce_sum_device) {
    if (c->comm.allreduce_sum_device(c->comm.user, c->d_result, K)) return LQ_E_COMM;
    c->reduced_globally = true;
  }
  int rc = rt_copy(c->h_result, c->d_result, K * sizeof(double), D2H, c->stream);
  if (rc) return rc;
  if (c->p2p_on) {
    LQ_CHECK(cudaMemcpyAsync(&c->h_result[15], c->p2p_flags + LQ_P2P_MAXNB, sizeof(double), cudaMemcpyDeviceToHost,
                             c->stream));
  }
  rc = rt_sync(c->stream);
  if (rc) return rc;
  if (c->p2p_on) {  // a flag wait that timed out (a neighbour died or left the SPMD sequence) latched an error word
    unsigned long long err;
    memcpy(&err, &c->h_result[15], sizeof(err));
    if (err) return LQ_E_COMM;
  }
  return LQ_OK;
}
template <class F>
static int reduce(lq_ctx* c, lq_i64 n, const F& f) {
  constexpr int K = F::K;
  lq_i64 nb = (n + LQ_RBLOCK - 1) / LQ_RBLOCK;
  int rc = reduce_reserve(c, nb, K);
  if (rc) return rc;
  lq_reduce_k<F><<<(unsigned)nb, LQ_RBLOCK, 0, c->stream>>>(f, n, c->d_partial);
  return reduce_finish<K>(c, nb);
}
#endif

#define LQ_DISPATCH(c, CALL)          \
  switch ((c)->g.D) {                 \
    case 2: { constexpr int DD = 2; CALL; } break; \
    case 3: { constexpr int DD = 3; CALL; } break; \
    case 4: { constexpr int DD = 4; CALL; } break; \
    default: return LQ_E_BADARG;      \
  }

#define LQ_TRY(x)          \
  do {                     \
    int rc_ = (x);         \
    if (rc_) return rc_;   \
  } while (0)

// ------------------------------------------------------------------------------------------------ helpers
static int p2p_exchange(lq_ctx* c, int which);
#ifndef LQ_HOST_EMU
static int p2p_barrier(lq_ctx* c);
#endif
static int ensure_halo(lq_ctx* c, int which) {
#ifndef LQ_HOST_EMU
  // a folded launch (LqFold) left its epoch for the next kernel of its chain to acquire; anybody else waits for it here
  if (c->decomposed && c->p2p_pending) LQ_TRY(p2p_barrier(c));
#endif
  if (!c->decomposed || c->halo_ok[which]) return LQ_OK;
  if (c->p2p_on) {
    LQ_TRY(p2p_exchange(c, which));
    c->halo_ok[which] = true;
    return LQ_OK;
  }
  if (!c->has_comm || !c->comm.halo_exchange) return LQ_E_COMM;
  int rc = c->comm.halo_exchange(c->comm.user, c, which);
  if (rc) return LQ_E_COMM;
  c->halo_ok[which] = true;
  return LQ_OK;
}
static int global_sum(lq_ctx* c, double* v, int n) {
  if (!c->decomposed) return LQ_OK;
  if (c->reduced_globally) {  // reduce() already summed over the ranks on the device
    c->reduced_globally = false;
    return LQ_OK;
  }
  if (!c->has_comm || !c->comm.allreduce_sum) return LQ_OK;  // rank-local partial sums
  return c->comm.allreduce_sum(c->comm.user, v, n) ? LQ_E_COMM : LQ_OK;
}
static lq_i64 global_sites(const lq_ctx* c) {
  lq_i64 n = 1;
  for (int d = 0; d < c->g.D; ++d) n *= c->g.gext[d];
  return n;
}
static int ensure_aos(lq_ctx* c, size_t bytes) {
  if (c->d_aos_bytes >= bytes) return LQ_OK;
  rt_free(c->d_aos);
  c->d_aos = nullptr;
  c->d_aos_bytes = 0;
  LQ_TRY(rt_malloc((void**)&c->d_aos, bytes));
  c->d_aos_bytes = bytes;
  return LQ_OK;
}
static int ensure_buf(cx** p, size_t bytes, lq_ctx* c) {
  if (*p) return LQ_OK;
  LQ_TRY(rt_malloc((void**)p, bytes));
  return rt_memset(*p, 0, bytes, c->stream);
}

static int ctx_create_common(lq_ctx** out, int device, int D, const int64_t* gext, const int* nproc, const int* coord,
                             double a, double beta, double CA) {
  if (!out || !gext) return LQ_E_BADARG;
  *out = nullptr;
  LqGeom g;
  LQ_TRY(init_geom(g, D, gext, nproc, coord));
#ifndef LQ_HOST_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "no CUDA device visible; this library has no CPU fallback");
    return LQ_E_NODEVICE;
  }
  if (device < 0 || device >= ndev) return LQ_E_BADARG;
#endif
  lq_ctx* c = new (std::nothrow) lq_ctx();
  if (!c) return LQ_E_CUDA;
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->g = g;
  c->a = a;
  c->beta = beta;
  c->CA = CA;
  c->decomposed = false;
  c->even_extents = true;
  for (int d = 0; d < LQ_MAXD; ++d) {
    c->nproc[d] = (nproc && d < D) ? nproc[d] : 1;
    if (c->nproc[d] > 1) c->decomposed = true;
    if (d < D && (g.gext[d] & 1 || g.ext[d] & 1)) c->even_extents = false;
    if (d < D && (g.ext[d] & 1)) c->odd_mask |= 1 << d;
  }
  int rc = LQ_OK;
#ifndef LQ_HOST_EMU
  DeviceGuard guard(device);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) rc = LQ_E_CUDA;
  c->own_stream = true;
#endif
  if (!rc) rc = rt_malloc((void**)&c->U, c->u_bytes());
  if (!rc) rc = rt_malloc((void**)&c->E, c->e_bytes());
  if (!rc) rc = rt_malloc((void**)&c->d_result, 16 * sizeof(double));
  if (!rc) rc = rt_malloc_host((void**)&c->h_result, 16 * sizeof(double));
  if (!rc) rc = rt_memset(c->U, 0, c->u_bytes(), c->stream);
  if (!rc) rc = rt_memset(c->E, 0, c->e_bytes(), c->stream);
  if (rc) {
    lq_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return LQ_OK;
}

extern "C" {

const char* lq_strerror(int code) {
  switch (code) {
    case LQ_OK: return "ok";
    case LQ_E_BADARG: return "bad argument";
    case LQ_E_SIZE: return "incompatible size (StateInitializationError::IncompatibleSize)";
    case LQ_E_CUDA: return "CUDA error";
    case LQ_E_COMM: return "halo / all-reduce transport error (is lq_set_comm registered?)";
    case LQ_E_ODD_EXTENT: return "sweeps on a decomposed context need even extents";
    case LQ_E_GAUSS_DIVERGED: return "Gauss projection did not converge (GaussProjectionError)";
    case LQ_E_ZERO_STEPS: return "zero integration steps (MultiIntegrationError::ZeroIntegration)";
    case LQ_E_NOSNAPSHOT: return "no snapshot to restore";
    case LQ_E_NODEVICE: return "no CUDA device (there is no CPU fallback)";
    default: return "unknown error";
  }
}
const char* lq_last_cuda_error(void) { return g_cuda_err; }
int lq_version(void) { return 100; }
int lq_device_count(int* n) {
  if (!n) return LQ_E_BADARG;
#ifdef LQ_HOST_EMU
  *n = 1;
#else
  int k = 0;
  if (cudaGetDeviceCount(&k) != cudaSuccess) k = 0;
  *n = k;
#endif
  return LQ_OK;
}

int lq_ctx_create(lq_ctx** out, int device, int D, const int64_t* extent, double a, double beta, double CA) {
  return ctx_create_common(out, device, D, extent, nullptr, nullptr, a, beta, CA);
}
int lq_ctx_create_dist(lq_ctx** out, int device, int D, const int64_t* gext, const int* proc_grid, const int* coord,
                       double a, double beta, double CA) {
  if (!proc_grid || !coord) return LQ_E_BADARG;
  return ctx_create_common(out, device, D, gext, proc_grid, coord, a, beta, CA);
}
int lq_ctx_clone(const lq_ctx* src, lq_ctx** out) {
  if (!src || !out) return LQ_E_BADARG;
  *out = nullptr;
  int64_t gext[LQ_MAXD];
  int coord[LQ_MAXD];
  for (int d = 0; d < LQ_MAXD; ++d) {
    gext[d] = src->g.gext[d];
    coord[d] = src->g.ext[d] > 0 ? src->g.goff[d] / src->g.ext[d] : 0;
  }
  lq_ctx* c = nullptr;
  LQ_TRY(ctx_create_common(&c, src->device, src->g.D, gext, src->nproc, coord, src->a, src->beta, src->CA));
  LQ_GUARD(c);
  c->flags = src->flags;
  c->integ_kind = src->integ_kind;
  c->integ_exp = src->integ_exp;
  c->integ_lambda = src->integ_lambda;
  c->t = src->t;
  c->comm = src->comm;
  c->has_comm = src->has_comm;
  // order the copy after everything already queued on the source stream
  int rc = rt_sync(src->stream);
  if (!rc) rc = rt_copy(c->U, src->U, c->u_bytes(), D2D, c->stream);
  if (!rc) rc = rt_copy(c->E, src->E, c->e_bytes(), D2D, c->stream);
  if (!rc) rc = rt_sync(c->stream);
  if (rc) {
    lq_ctx_destroy(c);
    return rc;
  }
  c->halo_ok[0] = src->halo_ok[0];
  c->halo_ok[1] = src->halo_ok[1];
  *out = c;
  return LQ_OK;
}
int lq_ctx_destroy(lq_ctx* c) {
  if (!c) return LQ_OK;
  LQ_GUARD(c);
  rt_sync(c->stream);
  rt_free(c->U);
  rt_free(c->U2);
  rt_free(c->E);
  rt_free(c->E2);
  rt_free(c->G);
  rt_free(c->G2);
  rt_free(c->T);
  rt_free(c->T2);
#ifndef LQ_HOST_EMU
  for (int q = 0; q < c->p2p_npeers; ++q)
    for (int b = 0; b < LQ_P2P_NBUF; ++b)
      if (c->p2p_base[q][b]) cudaIpcCloseMemHandle(c->p2p_base[q][b]);
  rt_free(c->p2p_flags);
  rt_free(c->d_push);
  rt_free(c->d_fold_counter);
  rt_free(c->stage_in);
  rt_free(c->stage_out);
#ifndef LQ_HOST_EMU
  if (c->copy_streams) {
    cudaStreamDestroy(c->s_h2d);
    cudaStreamDestroy(c->s_d2h);
    cudaEventDestroy(c->ev_in_ready);
    cudaEventDestroy(c->ev_in_consumed);
    cudaEventDestroy(c->ev_out_ready);
    cudaEventDestroy(c->ev_out_done);
  }
#endif
#endif
  rt_free(c->snapU);
  rt_free(c->snapE);
  rt_free(c->hmcU);
  rt_free(c->d_aos);
  rt_free(c->d_partial);
  rt_free(c->d_result);
  rt_free_host(c->h_result);
#ifndef LQ_HOST_EMU
  if (c->prof_ev) {
    for (int i = 0; i < 2 * LQ_PROF_CAP; ++i)
      if (c->prof_ev[i]) cudaEventDestroy(c->prof_ev[i]);
    free(c->prof_ev);
  }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
#endif
  delete c;
  return LQ_OK;
}
int lq_set_flags(lq_ctx* c, int flags) {
  if (!c) return LQ_E_BADARG;
  c->flags = flags;
  return LQ_OK;
}
int lq_get_flags(lq_ctx* c, int* flags) {
  if (!c || !flags) return LQ_E_BADARG;
  *flags = c->flags;
  return LQ_OK;
}
int lq_set_beta(lq_ctx* c, double beta) {
  if (!c) return LQ_E_BADARG;
  c->beta = beta;
  return LQ_OK;
}
int lq_sync(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  return rt_sync(c->stream);
}
int lq_stream(lq_ctx* c, void** s) {
  if (!c || !s) return LQ_E_BADARG;
  *s = (void*)c->stream;
  return LQ_OK;
}
int lq_set_stream(lq_ctx* c, void* s) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  LQ_TRY(rt_sync(c->stream));
#ifndef LQ_HOST_EMU
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
#else
  c->stream = s;
#endif
  c->own_stream = false;
  return LQ_OK;
}
int64_t lq_num_sites(const lq_ctx* c) { return c ? c->g.vol : 0; }
int64_t lq_num_links(const lq_ctx* c) { return c ? c->g.vol * c->g.D : 0; }
int64_t lq_t(const lq_ctx* c) { return c ? c->t : 0; }
int lq_set_t(lq_ctx* c, int64_t t) {
  if (!c) return LQ_E_BADARG;
  c->t = t;
  return LQ_OK;
}
int64_t lq_kernel_launches(const lq_ctx* c) { return c ? c->launches : 0; }
int lq_set_comm(lq_ctx* c, const lq_comm* comm) {
  if (!c) return LQ_E_BADARG;
  if (comm) {
    c->comm = *comm;
    c->has_comm = true;
  } else {
    c->has_comm = false;
  }
  return LQ_OK;
}
int lq_is_decomposed(const lq_ctx* c, int dir) { return (c && dir >= 0 && dir < c->g.D) ? c->g.ghost[dir] : 0; }

// ---------------------------------------------------------------------------------------------- marshalling
static int links_from_device_aos(lq_ctx* c, const double* d_aos) {
#if defined(LQ_HAVE_TUNED) && !defined(LQ_HOST_EMU)
  // the bulk-copy (TMA) row kernel needs a 16-byte aligned AoS pointer; anything else takes the per-thread functor
  if (lq_tuned_aos_ok(c->g) && ((uintptr_t)d_aos & 15) == 0 && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
    LQ_CHECK(lq_tuned_links_aos(c->stream, c->g, c->U, const_cast<double*>(d_aos), false));
    c->launches++;
  } else
#endif
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KLinksFromAos<DD>{c->g, d_aos, c->U}))));
  c->halo_ok[0] = false;
  c->g_valid = false;
  return LQ_OK;
}
static int efield_from_device_aos(lq_ctx* c, const double* d_aos) {
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KEFromAos<DD>{c->g, d_aos, c->E}))));
  c->halo_ok[1] = false;
  c->g_valid = false;
  return LQ_OK;
}
static int links_to_device_aos(lq_ctx* c, double* d_aos) {
#if defined(LQ_HAVE_TUNED) && !defined(LQ_HOST_EMU)
  if (lq_tuned_aos_ok(c->g) && ((uintptr_t)d_aos & 15) == 0 && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
    LQ_CHECK(lq_tuned_links_aos(c->stream, c->g, c->U, d_aos, true));
    c->launches++;
    return LQ_OK;
  }
#endif
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KLinksToAos<DD>{c->g, c->U, d_aos}))));
  return LQ_OK;
}
static int efield_to_device_aos(lq_ctx* c, double* d_aos) {
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KEToAos<DD>{c->g, c->E, d_aos}))));
  return LQ_OK;
}
int lq_links_upload(lq_ctx* c, const double* aos, int64_t n_links) {
  if (!c || !aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_links * 18 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(rt_copy(c->d_aos, aos, bytes, H2D, c->stream));
  LQ_TRY(links_from_device_aos(c, c->d_aos));
  return rt_sync(c->stream);
}
int lq_links_download(lq_ctx* c, double* aos, int64_t n_links) {
  if (!c || !aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_links * 18 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(links_to_device_aos(c, c->d_aos));
  LQ_TRY(rt_copy(aos, c->d_aos, bytes, D2H, c->stream));
  return rt_sync(c->stream);
}
// ---- pipelined marshalling: the host <-> device copies of consecutive, independent states overlap the kernels.
// A host that streams a batch of configurations through one context (measurement runs over a stored ensemble; the e2e leg
// of bench.py) begins the upload of the NEXT configuration and the download of the PREVIOUS result while the current
// trajectory computes: PCIe is idle during compute and full duplex.  Two staging buffers and two copy streams; events order
// the transpositions (context stream) against the copies.
#ifndef LQ_HOST_EMU
static int copy_streams(lq_ctx* c) {
  if (c->copy_streams) return LQ_OK;
  LQ_CHECK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  LQ_CHECK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  LQ_CHECK(cudaEventCreateWithFlags(&c->ev_in_ready, cudaEventDisableTiming));
  LQ_CHECK(cudaEventCreateWithFlags(&c->ev_in_consumed, cudaEventDisableTiming));
  LQ_CHECK(cudaEventCreateWithFlags(&c->ev_out_ready, cudaEventDisableTiming));
  LQ_CHECK(cudaEventCreateWithFlags(&c->ev_out_done, cudaEventDisableTiming));
  c->copy_streams = true;
  return LQ_OK;
}
#endif
int lq_links_upload_begin(lq_ctx* c, const double* aos, int64_t n_links) {
  if (!c || !aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  if (c->up_pending) return LQ_E_BADARG;  // one upload in flight at a time
  LQ_GUARD(c);
  const size_t bytes = (size_t)n_links * 18 * sizeof(double);
  if (!c->stage_in) LQ_TRY(rt_malloc((void**)&c->stage_in, bytes));
#ifndef LQ_HOST_EMU
  LQ_TRY(copy_streams(c));
  LQ_CHECK(cudaStreamWaitEvent(c->s_h2d, c->ev_in_consumed, 0));  // the previous commit is done reading the staging buffer
  LQ_CHECK(cudaMemcpyAsync(c->stage_in, aos, bytes, cudaMemcpyHostToDevice, c->s_h2d));
  LQ_CHECK(cudaEventRecord(c->ev_in_ready, c->s_h2d));
#else
  memcpy(c->stage_in, aos, bytes);
#endif
  c->up_pending = true;
  return LQ_OK;
}
int lq_links_upload_commit(lq_ctx* c) {
  if (!c || !c->up_pending) return LQ_E_BADARG;
  LQ_GUARD(c);
#ifndef LQ_HOST_EMU
  LQ_CHECK(cudaStreamWaitEvent(c->stream, c->ev_in_ready, 0));
#endif
  LQ_TRY(links_from_device_aos(c, c->stage_in));
#ifndef LQ_HOST_EMU
  LQ_CHECK(cudaEventRecord(c->ev_in_consumed, c->stream));
#endif
  c->up_pending = false;
  return LQ_OK;
}
int lq_links_download_begin(lq_ctx* c, double* aos, int64_t n_links) {
  if (!c || !aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  const size_t bytes = (size_t)n_links * 18 * sizeof(double);
  if (!c->stage_out) LQ_TRY(rt_malloc((void**)&c->stage_out, bytes));
#ifndef LQ_HOST_EMU
  LQ_TRY(copy_streams(c));
  LQ_CHECK(cudaStreamWaitEvent(c->stream, c->ev_out_done, 0));  // the previous download has left the staging buffer
  LQ_TRY(links_to_device_aos(c, c->stage_out));
  LQ_CHECK(cudaEventRecord(c->ev_out_ready, c->stream));
  LQ_CHECK(cudaStreamWaitEvent(c->s_d2h, c->ev_out_ready, 0));
  LQ_CHECK(cudaMemcpyAsync(aos, c->stage_out, bytes, cudaMemcpyDeviceToHost, c->s_d2h));
  LQ_CHECK(cudaEventRecord(c->ev_out_done, c->s_d2h));
#else
  LQ_TRY(links_to_device_aos(c, c->stage_out));
  memcpy(aos, c->stage_out, bytes);
#endif
  return LQ_OK;
}
int lq_copies_wait(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
#ifndef LQ_HOST_EMU
  if (c->copy_streams) {
    LQ_CHECK(cudaStreamSynchronize(c->s_h2d));
    LQ_CHECK(cudaStreamSynchronize(c->s_d2h));
  }
#endif
  return LQ_OK;
}
int lq_efield_upload(lq_ctx* c, const double* aos, int64_t n_links) {
  if (!c || !aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_links * 8 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(rt_copy(c->d_aos, aos, bytes, H2D, c->stream));
  LQ_TRY(efield_from_device_aos(c, c->d_aos));
  return rt_sync(c->stream);
}
int lq_efield_download(lq_ctx* c, double* aos, int64_t n_links) {
  if (!c || !aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_links * 8 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(efield_to_device_aos(c, c->d_aos));
  LQ_TRY(rt_copy(aos, c->d_aos, bytes, D2H, c->stream));
  return rt_sync(c->stream);
}
int lq_links_upload_device(lq_ctx* c, const double* d_aos, int64_t n_links) {
  if (!c || !d_aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  return links_from_device_aos(c, d_aos);
}
int lq_links_download_device(lq_ctx* c, double* d_aos, int64_t n_links) {
  if (!c || !d_aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  return links_to_device_aos(c, d_aos);
}
int lq_efield_upload_device(lq_ctx* c, const double* d_aos, int64_t n_links) {
  if (!c || !d_aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  return efield_from_device_aos(c, d_aos);
}
int lq_efield_download_device(lq_ctx* c, double* d_aos, int64_t n_links) {
  if (!c || !d_aos) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  return efield_to_device_aos(c, d_aos);
}
int lq_links_set_cold(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KLinksCold<DD>{c->g, c->U}))));
  c->halo_ok[0] = false;
  c->g_valid = false;
  return LQ_OK;
}
int lq_efield_set_zero(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  c->g_valid = false;
  c->halo_ok[1] = true;  // zero everywhere, ghosts included
  return rt_memset(c->E, 0, c->e_bytes(), c->stream);
}
int lq_links_set_random(lq_ctx* c, uint64_t seed, uint64_t counter) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KLinksRandom<DD>{c->g, c->U, seed, counter}))));
  c->halo_ok[0] = false;
  c->g_valid = false;
  return LQ_OK;
}

// ---------------------------------------------------------------------------------------------- observables
static int plaquette_all(lq_ctx* c, double v[3]) {
  LQ_TRY(ensure_halo(c, 0));
  ProfScope ps(c, LQ_PROF_PLAQUETTE);
#if defined(LQ_HAVE_TUNED) && !defined(LQ_HOST_EMU)
  if (lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
    const lq_i64 nb = lq_tuned_plaquette_blocks(c->g);
    LQ_TRY(reduce_reserve(c, nb, 3));
    LQ_CHECK(lq_tuned_plaquette(c->stream, c->g, c->U, c->CA, c->d_partial));
    LQ_TRY(reduce_finish<3>(c, nb));
  } else
#endif
  LQ_DISPATCH(c, LQ_TRY((reduce(c, KPlaquette<DD>::items(c->g), KPlaquette<DD>{c->g, c->U, c->CA}))));
  for (int k = 0; k < 3; ++k) v[k] = c->h_result[k];
  return global_sum(c, v, 3);
}
int lq_plaquette_sum(lq_ctx* c, double out[2]) {
  if (!c || !out) return LQ_E_BADARG;
  LQ_GUARD(c);
  double v[3];
  LQ_TRY(plaquette_all(c, v));
  out[0] = v[0];
  out[1] = v[1];
  return LQ_OK;
}
int lq_average_trace_plaquette(lq_ctx* c, double out[2]) {
  if (!c || !out) return LQ_E_BADARG;
  LQ_GUARD(c);
  double v[3];
  LQ_TRY(plaquette_all(c, v));
  double npl = (double)(global_sites(c) * (c->g.D * (c->g.D - 1)) / 2);  // field.rs:801-803
  out[0] = v[0] / npl;
  out[1] = v[1] / npl;
  return LQ_OK;
}
int lq_hamiltonian_links(lq_ctx* c, double* h) {
  if (!c || !h) return LQ_E_BADARG;
  LQ_GUARD(c);
  double v[3];
  LQ_TRY(plaquette_all(c, v));
  *h = v[2] * c->beta;
  return LQ_OK;
}
int lq_hamiltonian_efield(lq_ctx* c, double* h) {
  if (!c || !h) return LQ_E_BADARG;
  LQ_GUARD(c);
  {
    ProfScope ps(c, LQ_PROF_EFIELD_ENERGY);
    LQ_DISPATCH(c, LQ_TRY((reduce(c, c->g.vol, KEfieldEnergy<DD>{c->g, c->E}))));
  }
  double v = c->h_result[0];
  LQ_TRY(global_sum(c, &v, 1));
  *h = v * c->beta;
  return LQ_OK;
}
int lq_hamiltonian_total(lq_ctx* c, double* h) {
  double a = 0, b = 0;
  LQ_TRY(lq_hamiltonian_links(c, &a));
  LQ_TRY(lq_hamiltonian_efield(c, &b));
  *h = a + b;
  return LQ_OK;
}

// field-strength observables (SURVEY section 8f "next": same stencil machinery, golden vectors field.rs:1580-1711)
static int field_strength(lq_ctx* c, int mode, int p0, int p1, double* aos_out, int64_t n_sites) {
  if (!c || !aos_out) return LQ_E_BADARG;
  if (n_sites != c->g.vol) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_sites * 18 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(ensure_halo(c, 0));
  LQ_DISPATCH(c, LQ_TRY((launch(c, c->g.vol, KFieldStrength<DD>{c->g, c->U, c->d_aos, mode, p0, p1, c->a}))));
  LQ_TRY(rt_copy(aos_out, c->d_aos, bytes, D2H, c->stream));
  return rt_sync(c->stream);
}
static bool signed_dir_ok(const lq_ctx* c, int sd) { return c && sd != 0 && abs(sd) <= c->g.D; }
int lq_clover(lq_ctx* c, int sdir_i, int sdir_j, double* aos_out, int64_t n_sites) {
  if (!signed_dir_ok(c, sdir_i) || !signed_dir_ok(c, sdir_j)) return LQ_E_BADARG;
  return field_strength(c, 0, sdir_i, sdir_j, aos_out, n_sites);
}
int lq_f_mu_nu(lq_ctx* c, int dir_i, int dir_j, double* aos_out, int64_t n_sites) {
  if (!c || dir_i < 0 || dir_j < 0 || dir_i >= c->g.D || dir_j >= c->g.D) return LQ_E_BADARG;
  return field_strength(c, 1, dir_i, dir_j, aos_out, n_sites);
}
int lq_magnetic_field(lq_ctx* c, int dir, double* aos_out, int64_t n_sites) {
  if (!c || dir < 0 || dir >= c->g.D) return LQ_E_BADARG;
  return field_strength(c, 2, dir, 0, aos_out, n_sites);
}

// ---------------------------------------------------------------------------------------------- molecular dynamics
static double force_coef(const lq_ctx* c) { return -sqrt(2.0 / c->CA) / c->a; }  // state.rs:1428
static double link_coef(const lq_ctx* c) { return sqrt(2.0 * c->CA) / c->a; }    // state.rs:1412-1416

int lq_staples(lq_ctx* c, double* aos_out, int64_t n_links) {
  if (!c || !aos_out) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_links * 18 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(ensure_halo(c, 0));
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KStaplesToAos<DD>{c->g, c->U, c->d_aos}))));
  LQ_TRY(rt_copy(aos_out, c->d_aos, bytes, D2H, c->stream));
  return rt_sync(c->stream);
}
int lq_force(lq_ctx* c, double* aos_out, int64_t n_links) {
  if (!c || !aos_out) return LQ_E_BADARG;
  if (n_links != lq_num_links(c)) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_links * 8 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(ensure_halo(c, 0));
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KForceToAos<DD>{c->g, c->U, c->d_aos, force_coef(c)}))));
  LQ_TRY(rt_copy(aos_out, c->d_aos, bytes, D2H, c->stream));
  return rt_sync(c->stream);
}
static int efield_step(lq_ctx* c, double dt, int nkick) {
  LQ_TRY(ensure_halo(c, 0));
  ProfScope ps(c, LQ_PROF_EFIELD_STEP);
#ifdef LQ_TUNED
  if (lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
    LQ_CHECK(lq_tuned_efield_step(c->stream, c->g, c->U, c->E, force_coef(c), dt, nkick));
    c->launches++;
    c->halo_ok[1] = false;
    c->g_valid = false;
    return LQ_OK;
  }
#endif
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KEfieldStep<DD>{c->g, c->U, c->E, force_coef(c), dt, nkick}))));
  c->halo_ok[1] = false;
  c->g_valid = false;
  return LQ_OK;
}
static int link_step(lq_ctx* c, const cx* Uin, cx* Uout, double dt, int use_exp) {
  ProfScope ps(c, LQ_PROF_LINK_STEP);
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KLinkStep<DD>{c->g, Uin, Uout, c->E, dt, link_coef(c), use_exp}))));
  c->halo_ok[0] = false;
  c->g_valid = false;
  return LQ_OK;
}
#ifdef LQ_TUNED
// Schedule and epochs of a launch with the halo synchronisation folded in (lq_tuned.cuh: LqFold).  false: not applicable
// (geometry, transport, or LQ_FLAG_FOLD_HALO_SYNC not set: the default) -- the caller keeps the barrier kernels.  The launch acquires the
// epoch a preceding folded launch of the same chain released (p2p_pending) and releases the next one.
static bool fold_fill(lq_ctx* c, int block_sites, int zfirst, LqFold& f) {
  if (!c->p2p_on || !c->d_push || !c->d_fold_counter || !(c->flags & LQ_FLAG_FOLD_HALO_SYNC)) return false;
  memset(&f, 0, sizeof(f));
  if (!lq_fold_geom(c->g, block_sites, zfirst, f)) return false;
  f.n = c->p2p_nnb;
  for (int k = 0; k < c->p2p_nnb; ++k)
    f.remote[k] = (unsigned long long*)c->p2p_base[c->p2p_peer[k]][LQ_P2P_NBUF - 1] + c->p2p_rev[k];
  f.mine = c->p2p_flags;
  f.counter = c->d_fold_counter;
  f.wait_value = c->p2p_pending ? c->p2p_epoch : 0;
  c->p2p_epoch += 1;
  f.signal_value = c->p2p_epoch;
  return true;
}
#endif
// E += nkick * (dt_e F[U]);  U <- step(U, E_new, dt_u)  in one kernel (second link buffer, then swap)
// chain: the previous operation on this context was the same fused step (inside one lq_symplectic_n / lq_md_n loop)
static int efield_link_step(lq_ctx* c, double dt_e, int nkick, double dt_u, int use_exp = 0, bool chain = false) {
  // inside a chain of folded launches the boundary blocks of this kernel acquire the previous step's epoch themselves
  if (!(chain && c->p2p_pending)) LQ_TRY(ensure_halo(c, 0));
  LQ_TRY(ensure_buf(&c->U2, c->u_bytes(), c));
#ifdef LQ_TUNED
  const bool tuned = lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS);
  if (tuned && c->p2p_on && c->d_push) {
    // compute + halo push in one kernel: boundary links go straight into the neighbours' ghost layers (NVLink)
    int bi = -1;
    for (int b = 0; b < LQ_P2P_NBUF - 1; ++b)
      if (c->own[b] == c->U2) bi = b;
    if (bi < 0 || !c->d_push) return LQ_E_COMM;
    // ready: the neighbours are done reading the ghost layers this kernel overwrites.  Inside a chain of fused steps
    // the two link buffers alternate, so the ghosts written now were last read by the neighbours' step BEFORE the
    // previous one -- and a neighbour only arrived at the previous step's data barrier (which this rank has passed)
    // after that kernel had finished: one barrier per step is enough.
    if (!chain) LQ_TRY(p2p_barrier(c));
    LqFold fold;
    if ((c->flags & 2048) && fold_fill(c, 32, (c->flags & 1024) ? 1 : 0, fold)) {  // A/B bits (both measured slower): 2048 folds the MD chain too, 1024 schedules its z faces first
      // the data barrier is folded in as well: the last boundary block releases this step's epoch, the boundary blocks of
      // the next step (or the barrier in ensure_halo, when something else follows) acquire it
      {
        ProfScope ps2(c, LQ_PROF_EFIELD_LINK_STEP);
        LQ_CHECK(lq_tuned_efield_link_step_fold(c->stream, c->g, c->U, c->U2, c->E, force_coef(c), dt_e, dt_u,
                                                link_coef(c), nkick, (const LqPush*)c->d_push + bi, use_exp, fold));
        c->launches++;
      }
      c->p2p_pending = true;
      c->p2p_exchanges++;
      cx* t = c->U;
      c->U = c->U2;
      c->U2 = t;
      c->halo_ok[0] = true;
      c->halo_ok[1] = false;
      c->g_valid = false;
      return LQ_OK;
    }
    {
      ProfScope ps2(c, LQ_PROF_EFIELD_LINK_STEP);
      LQ_CHECK(lq_tuned_efield_link_step_push(c->stream, c->g, c->U, c->U2, c->E, force_coef(c), dt_e, dt_u,
                                              link_coef(c), nkick, (const LqPush*)c->d_push + bi, use_exp));
      c->launches++;
    }
    LQ_TRY(p2p_barrier(c));  // data: their pushes into my ghost layers have landed
    c->p2p_exchanges++;
    cx* t = c->U;
    c->U = c->U2;
    c->U2 = t;
    c->halo_ok[0] = true;
    c->halo_ok[1] = false;
    c->g_valid = false;
    return LQ_OK;
  }
  if (tuned) {
    ProfScope ps(c, LQ_PROF_EFIELD_LINK_STEP);
    LQ_CHECK(lq_tuned_efield_link_step(c->stream, c->g, c->U, c->U2, c->E, force_coef(c), dt_e, dt_u, link_coef(c),
                                       nkick, use_exp));
    c->launches++;
  } else
#endif
  {
    ProfScope ps(c, LQ_PROF_EFIELD_LINK_STEP);
    LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g),
                                   KEfieldLinkStep<DD>{c->g, c->U, c->U2, c->E, force_coef(c), dt_e, dt_u, link_coef(c),
                                                       nkick, use_exp}))));
  }
  cx* t = c->U;
  c->U = c->U2;
  c->U2 = t;
  c->halo_ok[0] = false;
  c->halo_ok[1] = false;
  c->g_valid = false;
  return LQ_OK;
}
int lq_efield_step(lq_ctx* c, double dt) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  return efield_step(c, dt, 1);
}
int lq_link_step(lq_ctx* c, double dt, int use_exp) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  return link_step(c, c->U, c->U, dt, use_exp);
}
int lq_integrate(lq_ctx* c, int kind, double dt) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  switch (kind) {
    case LQ_SYNC_SYNC: {  // both from the old state, symplectic_euler_rayon.rs:127-143
      LQ_TRY(ensure_buf(&c->U2, c->u_bytes(), c));
      LQ_TRY(link_step(c, c->U, c->U2, dt, 0));
      LQ_TRY(efield_step(c, dt, 1));  // still reads the old links in c->U
      cx* t = c->U;
      c->U = c->U2;
      c->U2 = t;
      c->halo_ok[0] = false;
      c->g_valid = false;
      c->t += 1;
      return LQ_OK;
    }
    case LQ_LEAP_LEAP:  // :145-168
      LQ_TRY(link_step(c, c->U, c->U, dt, 0));
      LQ_TRY(efield_step(c, dt, 1));
      c->t += 1;
      return LQ_OK;
    case LQ_SYNC_LEAP:  // :170-191 (t unchanged)
      return efield_step(c, dt / 2.0, 1);
    case LQ_LEAP_SYNC:  // :193-218
      LQ_TRY(link_step(c, c->U, c->U, dt, 0));
      LQ_TRY(efield_step(c, dt / 2.0, 1));
      c->t += 1;
      return LQ_OK;
    case LQ_SYMPLECTIC:  // :220-252
      LQ_TRY(efield_step(c, dt / 2.0, 1));
      LQ_TRY(link_step(c, c->U, c->U, dt, 0));
      LQ_TRY(efield_step(c, dt / 2.0, 1));
      c->t += 1;
      return LQ_OK;
    default:
      return LQ_E_BADARG;
  }
}
int lq_symplectic_n(lq_ctx* c, double dt, int64_t n) {
  if (!c) return LQ_E_BADARG;
  if (n <= 0) return LQ_E_ZERO_STEPS;
  LQ_GUARD(c);
  if (c->flags & LQ_FLAG_NO_KICK_MERGE) {
    for (int64_t k = 0; k < n; ++k) LQ_TRY(lq_integrate(c, LQ_SYMPLECTIC, dt));
    return LQ_OK;
  }
  // n x [E(dt/2) U(dt) E(dt/2)]: the trailing kick of step k and the leading kick of step k+1 see the same links,
  // so the force is evaluated once and applied twice (same rounding sequence); each kick is fused with the link
  // step that follows it.
  for (int64_t k = 0; k < n; ++k) LQ_TRY(efield_link_step(c, dt / 2.0, k == 0 ? 1 : 2, dt, 0, k > 0));
  LQ_TRY(efield_step(c, dt / 2.0, 1));
  c->t += n;
  return LQ_OK;
}
// Integrator options beyond the reference's (not parity paths of the crate, checked against the
// oracle's composition of the same E / U steps).  All are compositions of the two reference updates, integrate_efield
// (integrator/mod.rs:240-254) and integrate_link (:216-233, Euler) or its exponential form (use_exp, su3.rs:832-855):
//   symplectic Euler / leap-frog   E(dt/2) U(dt) E(dt/2)                                   (symplectic_euler_rayon.rs:220-252)
//   Omelyan 2nd-order minimum norm E(l dt) U(dt/2) E((1-2l) dt) U(dt/2) E(l dt),  l = 0.1931833...
// with the adjacent E kicks of consecutive steps merged.  (0, any, 0) is lq_symplectic_n itself (fused kernels).
int lq_set_integrator(lq_ctx* c, int kind, double lambda, int use_exp) {
  if (!c || (kind != LQ_INTEGRATOR_SYMPLECTIC_EULER && kind != LQ_INTEGRATOR_OMELYAN)) return LQ_E_BADARG;
  if (kind == LQ_INTEGRATOR_OMELYAN && !(lambda > 0.0 && lambda < 0.5)) return LQ_E_BADARG;
  c->integ_kind = kind;
  c->integ_lambda = lambda;
  c->integ_exp = use_exp != 0;
  return LQ_OK;
}
int lq_md_n(lq_ctx* c, double dt, int64_t n) {
  if (!c) return LQ_E_BADARG;
  if (n <= 0) return LQ_E_ZERO_STEPS;
  LQ_GUARD(c);
  if (c->integ_kind == LQ_INTEGRATOR_SYMPLECTIC_EULER && !c->integ_exp) return lq_symplectic_n(c, dt, n);
  // every E kick is fused with the link step that follows it (one kernel: force + kick + link update, the new boundary
  // links pushed to the neighbour ranks); only the closing kick runs alone
  const int ex = c->integ_exp;
  if (c->integ_kind == LQ_INTEGRATOR_SYMPLECTIC_EULER) {
    for (int64_t k = 0; k < n; ++k) LQ_TRY(efield_link_step(c, k == 0 ? dt / 2.0 : dt, 1, dt, ex, k > 0));
    LQ_TRY(efield_step(c, dt / 2.0, 1));
  } else {
    const double l = c->integ_lambda;
    for (int64_t k = 0; k < n; ++k) {
      LQ_TRY(efield_link_step(c, k == 0 ? l * dt : 2.0 * l * dt, 1, dt / 2.0, ex, k > 0));
      LQ_TRY(efield_link_step(c, (1.0 - 2.0 * l) * dt, 1, dt / 2.0, ex, true));
    }
    LQ_TRY(efield_step(c, l * dt, 1));
  }
  c->t += n;
  return LQ_OK;
}
int lq_leapfrog_n(lq_ctx* c, double dt, int64_t n) {
  if (!c) return LQ_E_BADARG;
  if (n <= 0) return LQ_E_ZERO_STEPS;
  LQ_GUARD(c);
  LQ_TRY(lq_integrate(c, LQ_SYNC_LEAP, dt));
  for (int64_t k = 0; k + 1 < n; ++k) LQ_TRY(lq_integrate(c, LQ_LEAP_LEAP, dt));
  return lq_integrate(c, LQ_LEAP_SYNC, dt);
}
int lq_reunitarize(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  ProfScope ps(c, LQ_PROF_REUNITARIZE);
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KReunitarize<DD>{c->g, c->U}))));
  c->halo_ok[0] = false;
  c->g_valid = false;
  return LQ_OK;
}

// ---------------------------------------------------------------------------------------------- momenta + Gauss
int lq_momenta_refresh(lq_ctx* c, uint64_t seed, uint64_t counter, double sigma) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  ProfScope ps(c, LQ_PROF_MOMENTA);
  LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KMomentaRefresh<DD>{c->g, c->E, seed, counter, sigma}))));
  c->halo_ok[1] = false;
  c->g_valid = false;
  return LQ_OK;
}
static int gauss_field(lq_ctx* c) {
  if (c->g_valid && c->G) return LQ_OK;  // E and U unchanged since the last evaluation (e.g. residual check -> next step)
  LQ_TRY(ensure_buf(&c->G, c->g_bytes(), c));
  LQ_TRY(ensure_halo(c, 0));
  LQ_TRY(ensure_halo(c, 1));
#ifndef LQ_HOST_EMU
  if (c->p2p_on && c->d_push) {
    // compute + halo push in one kernel: boundary sites of G go straight into the neighbours' ghost layers
    // Successive evaluations alternate between the two Gauss-field buffers.  The ghosts of the buffer written now were
    // last read (by the projection step) two evaluations ago, and the data barrier of the evaluation in between -- which
    // this rank has passed -- was reached by every neighbour only after that reader had finished: no "ready" barrier is
    // needed as long as the previous evaluation also went through this path (g_pp_safe).
    const bool skip_ready = c->g_pp_safe && c->G2;
    if (skip_ready) {
      cx* t = c->G;
      c->G = c->G2;
      c->G2 = t;
    }
    int bi = -1;
    for (int b = 0; b < LQ_P2P_NBUF - 1; ++b)
      if (c->own[b] == c->G) bi = b;
    if (bi < 0) return LQ_E_COMM;
    if (!skip_ready) LQ_TRY(p2p_barrier(c));
    {
      ProfScope ps(c, LQ_PROF_GAUSS_FIELD);
#ifdef LQ_HAVE_TUNED
      if (lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
        LQ_CHECK(lq_tuned_gauss_field(c->stream, c->g, c->U, c->E, c->G, (const LqPush*)c->d_push + bi));
        c->launches++;
      } else
#endif
      LQ_DISPATCH(c, LQ_TRY((launch(c, c->g.vol, KGaussField<DD>{c->g, c->U, c->E, c->G, (const LqPush*)c->d_push + bi}))));
    }
    LQ_TRY(p2p_barrier(c));
    c->p2p_exchanges++;
    c->halo_ok[2] = true;
    c->g_valid = true;
    c->g_pp_safe = true;
    return LQ_OK;
  }
#endif
  ProfScope ps(c, LQ_PROF_GAUSS_FIELD);
#if defined(LQ_HAVE_TUNED) && !defined(LQ_HOST_EMU)
  if (lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
    LQ_CHECK(lq_tuned_gauss_field(c->stream, c->g, c->U, c->E, c->G, nullptr));
    c->launches++;
  } else
#endif
  LQ_DISPATCH(c, LQ_TRY((launch(c, c->g.vol, KGaussField<DD>{c->g, c->U, c->E, c->G, nullptr}))));
  c->halo_ok[2] = false;
  c->g_valid = true;
  return LQ_OK;
}
int lq_gauss_field(lq_ctx* c, double* aos_out, int64_t n_sites) {
  if (!c || !aos_out) return LQ_E_BADARG;
  if (n_sites != c->g.vol) return LQ_E_SIZE;
  LQ_GUARD(c);
  size_t bytes = (size_t)n_sites * 18 * sizeof(double);
  LQ_TRY(ensure_aos(c, bytes));
  LQ_TRY(gauss_field(c));
  LQ_DISPATCH(c, LQ_TRY((launch(c, c->g.vol, KGaussToAos<DD>{c->g, c->G, c->d_aos}))));
  LQ_TRY(rt_copy(aos_out, c->d_aos, bytes, D2H, c->stream));
  return rt_sync(c->stream);
}
int lq_gauss_sum_div(lq_ctx* c, double* out) {
  if (!c || !out) return LQ_E_BADARG;
  LQ_GUARD(c);
  LQ_TRY(gauss_field(c));
  {
    ProfScope ps(c, LQ_PROF_GAUSS_DIV);
    LQ_DISPATCH(c, LQ_TRY((reduce(c, c->g.vol, KGaussDiv<DD>{c->g, c->G}))));
  }
  double v = c->h_result[0];
  LQ_TRY(global_sum(c, &v, 1));
  *out = v;
  return LQ_OK;
}
int lq_gauss_project_step(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  LQ_TRY(gauss_field(c));     // G of the current (U, E); a no-op when the previous iteration left it behind
  LQ_TRY(ensure_halo(c, 2));  // G(x +- i)
  LQ_TRY(ensure_halo(c, 1));  // E_i(x - i)
  LQ_TRY(ensure_halo(c, 0));
  LQ_TRY(ensure_buf(&c->E2, c->e_bytes(), c));
  if (!(c->flags & LQ_FLAG_GAUSS_FUSED)) {
    ProfScope ps(c, LQ_PROF_GAUSS_STEP);
#if defined(LQ_HAVE_TUNED) && !defined(LQ_HOST_EMU)
    if (lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
      LQ_CHECK(lq_tuned_gauss_step(c->stream, c->g, c->U, c->G, c->E, c->E2));
      c->launches++;
    } else
#endif
    LQ_DISPATCH(c, LQ_TRY((launch(c, lq_link_items(c->g), KGaussProjectStep<DD>{c->g, c->U, c->G, c->E, c->E2}))));
    // low-ghost links of the split directions: recomputed, not exchanged.  (Folding them into the launch above was
    // measured slower on 4 GPUs, 49.1 vs 42.9 ms per trajectory: the combined functor costs the main path occupancy.)
    for (int hd = 1; hd < c->g.D; ++hd)
      if (c->g.ghost[hd])
        LQ_DISPATCH(c, LQ_TRY((launch(c, c->g.vol / c->g.ext[hd],
                                       KGaussProjectGhost<DD>{c->g, c->U, c->G, c->E, c->E2, hd}))));
    cx* t = c->E;
    c->E = c->E2;
    c->E2 = t;
    c->halo_ok[1] = true;  // ensure_halo(c, 1) above made the inputs valid; the entries consumers read were updated
    c->g_valid = false;
    return LQ_OK;
  }
  LQ_TRY(ensure_buf(&c->G2, c->g_bytes(), c));
  c->g_pp_safe = false;  // this path swaps the two G buffers without a barrier of its own
  {
    // projection step + Gauss field of the projected E in one pass (KGaussIter)
    ProfScope ps(c, LQ_PROF_GAUSS_STEP);
    LQ_DISPATCH(c, LQ_TRY((launch(c, c->g.vol, KGaussIter<DD>{c->g, c->U, c->G, c->E, c->E2, c->G2}))));
  }
  cx* t = c->E;
  c->E = c->E2;
  c->E2 = t;
  t = c->G;
  c->G = c->G2;
  c->G2 = t;
  c->halo_ok[1] = true;   // the low-ghost entries consumers read were recomputed by the kernel
  c->halo_ok[2] = false;
  c->g_valid = true;
  return LQ_OK;
}
#ifdef LQ_TUNED
static int push_table_index(const lq_ctx* c, const cx* buf) {
  for (int b = 0; b < LQ_P2P_NBUF - 1; ++b)
    if (c->own[b] == buf) return b;
  return -1;
}
// project_to_gauss (field.rs:1265-1294) on the transported field: T = U^+ E U once, then ONE kernel per iteration
// (lq_gausst4_kernel: links read once, no Gauss-field array).  The residual of the state after steps 1, 5, 9, ... comes
// out of the NEXT iteration's kernel (it forms G of its input anyway), which therefore runs speculatively: when the
// check passes its output is simply not swapped in.
static int gauss_project_transported(lq_ctx* c, int64_t max_steps, int64_t* steps_out) {
  LQ_TRY(ensure_halo(c, 0));
  LQ_TRY(ensure_halo(c, 1));
  LQ_TRY(ensure_buf(&c->E2, c->e_bytes(), c));
  LQ_TRY(ensure_buf(&c->T, c->t_bytes(), c));
  LQ_TRY(ensure_buf(&c->T2, c->t_bytes(), c));
  const bool push = c->p2p_on && c->d_push;
  const int variant = (c->flags >> 6) & 3;  // A/B switch of the iteration kernel (bits 64, 128 of the flags)
  const lq_i64 nb = lq_tuned_gausst_blocks(c->g, variant);
  LQ_TRY(reduce_reserve(c, nb, 1));
  const double thr = LQ_EPS * (double)(global_sites(c) * 4 * 8 * 10);  // field.rs:1285
  auto table = [&](const cx* buf) -> const LqPush* {
    if (!push) return nullptr;
    const int bi = push_table_index(c, buf);
    return bi < 0 ? nullptr : (const LqPush*)c->d_push + bi;
  };
  {
    if (push) {
      if (!table(c->T)) return LQ_E_COMM;
      LQ_TRY(p2p_barrier(c));  // ready: nobody still reads the ghost layers of T
    }
    ProfScope ps(c, LQ_PROF_GAUSS_FIELD);
    LQ_CHECK(lq_tuned_gauss_tinit(c->stream, c->g, c->U, c->E, c->T, table(c->T)));
    c->launches++;
  }
  if (push) {
    LQ_TRY(p2p_barrier(c));
    c->p2p_exchanges++;
  }
  int64_t steps = 0;
  int rc = LQ_OK;
  for (;;) {
    // the state after `steps` projection steps is in (E, T); check it when steps = 1, 5, 9, ...
    const bool want_res = steps >= 1 && ((steps - 1) & 3) == 0;
    if (push && (!table(c->E2) || !table(c->T2))) return LQ_E_COMM;
    LqFold fold;
    if (push && variant == 0 && fold_fill(c, 128, 1, fold)) {
      // halo synchronisation inside the iteration kernel: its boundary blocks acquire the previous iteration's epoch and
      // the last of them releases this one's
      ProfScope ps(c, LQ_PROF_GAUSS_STEP);
      LQ_CHECK(lq_tuned_gauss_titer_fold(c->stream, c->g, c->U, c->E, c->T, c->E2, c->T2, c->d_partial, want_res,
                                         table(c->E2), table(c->T2), fold));
      c->launches++;
      c->p2p_pending = true;
      c->p2p_exchanges++;
    } else {
      {
        ProfScope ps(c, LQ_PROF_GAUSS_STEP);
        LQ_CHECK(lq_tuned_gauss_titer(c->stream, c->g, c->U, c->E, c->T, c->E2, c->T2, c->d_partial, want_res,
                                      table(c->E2), table(c->T2), variant));
        c->launches++;
      }
      if (push) {
        LQ_TRY(p2p_barrier(c));  // data (the ping-pong of the two buffer pairs makes a separate "ready" barrier unnecessary)
        c->p2p_exchanges++;
      }
    }
    if (want_res) {
      ProfScope ps(c, LQ_PROF_GAUSS_DIV);
      LQ_TRY(reduce_finish<1>(c, nb));
      double v = c->h_result[0];
      LQ_TRY(global_sum(c, &v, 1));
      if (v != v) {
        rc = LQ_E_GAUSS_DIVERGED;
        break;
      }
      if (v <= thr) break;  // (E, T) is the projected state; the speculative step in (E2, T2) is dropped
      if (steps >= max_steps) {
        rc = LQ_E_GAUSS_DIVERGED;
        break;
      }
    }
    cx* t = c->E;
    c->E = c->E2;
    c->E2 = t;
    t = c->T;
    c->T = c->T2;
    c->T2 = t;
    ++steps;
  }
  c->halo_ok[1] = push;  // the pushes kept every ghost entry of E current
  c->g_valid = false;
  c->g_pp_safe = false;
  if (steps_out) *steps_out = steps;
  return rc;
}
#endif
int lq_gauss_project(lq_ctx* c, int64_t max_steps, int64_t* steps_out) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  if (max_steps <= 0) max_steps = 1 << 20;
#ifdef LQ_TUNED
  if (lq_tuned_ok(c->g) && !(c->flags & (LQ_FLAG_GENERIC_KERNELS | LQ_FLAG_GAUSS_FUSED | LQ_FLAG_GAUSS_TWO_PASS)) &&
      (!c->decomposed || (c->p2p_on && c->d_push)))
    return gauss_project_transported(c, max_steps, steps_out);
#endif
  LQ_TRY(lq_gauss_project_step(c));
  int64_t steps = 1;
  const double thr = LQ_EPS * (double)(global_sites(c) * 4 * 8 * 10);  // field.rs:1285
  for (;;) {
    double v;
    LQ_TRY(lq_gauss_sum_div(c, &v));
    if (v != v) {
      if (steps_out) *steps_out = steps;
      return LQ_E_GAUSS_DIVERGED;
    }
    if (v <= thr) break;
    if (steps >= max_steps) {
      if (steps_out) *steps_out = steps;
      return LQ_E_GAUSS_DIVERGED;
    }
    for (int k = 0; k < 4; ++k) {
      LQ_TRY(lq_gauss_project_step(c));
      ++steps;
    }
  }
  if (steps_out) *steps_out = steps;
  return LQ_OK;
}

// ---------------------------------------------------------------------------------------------- sweeps
// Sub-steps of a sweep: for dir, [for colour class of an odd lattice,] for parity.  Even extents: two colours, vol/2
// sites per launch.  Odd extents (single-rank contexts only): classes (boundary mask, parity), every launch scans the
// whole volume (lq_site_class).
static bool sweep_supported(const lq_ctx* c) { return c->even_extents || !c->decomposed; }
#define LQ_SWEEP_LOOP(c, BODY)                                             \
  for (int d = 0; d < (c)->g.D; ++d)                                      \
    for (int cm = 0; cm <= (c)->odd_mask; ++cm) {                         \
      if (cm & ~(c)->odd_mask) continue;                                  \
      for (int p = 0; p < 2; ++p) {                                       \
        const int om = (c)->odd_mask;                                     \
        const lq_i64 nitems = om ? (c)->g.vol : (c)->g.vol / 2;           \
        BODY                                                              \
      }                                                                   \
    }
int lq_sweep_heatbath(lq_ctx* c, uint64_t seed, uint64_t counter, double coupling_scale) {
  if (!c) return LQ_E_BADARG;
  if (!sweep_supported(c)) return LQ_E_ODD_EXTENT;
  LQ_GUARD(c);
  LQ_SWEEP_LOOP(c, {
    LQ_TRY(ensure_halo(c, 0));
    ProfScope ps(c, LQ_PROF_HEATBATH);
#ifdef LQ_TUNED
    if (!om && lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
      LQ_CHECK(lq_tuned_sweep(c->stream, c->g, c->U, 0, d, p, c->flags, 0, c->beta * coupling_scale, seed, counter));
      c->launches++;
    } else
#endif
    LQ_DISPATCH(c, LQ_TRY((launch(c, nitems, KHeatBath<DD>{c->g, c->U, d, p, c->flags, c->beta * coupling_scale, seed,
                                                          counter, om, cm}))));
    c->halo_ok[0] = false;
    c->g_valid = false;
  })
  return LQ_OK;
}
int lq_sweep_overrelax(lq_ctx* c, int kind) {
  if (!c || (kind != LQ_OR_ROTATION && kind != LQ_OR_REVERSE && kind != LQ_OR_SU2_SUBGROUPS)) return LQ_E_BADARG;
  if (!sweep_supported(c)) return LQ_E_ODD_EXTENT;
  LQ_GUARD(c);
  LQ_SWEEP_LOOP(c, {
    LQ_TRY(ensure_halo(c, 0));
    ProfScope ps(c, LQ_PROF_OVERRELAX);
#ifdef LQ_TUNED
    if (!om && lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
      LQ_CHECK(lq_tuned_sweep(c->stream, c->g, c->U, 1, d, p, c->flags, kind, 0.0, 0, 0));
      c->launches++;
    } else
#endif
    LQ_DISPATCH(c, LQ_TRY((launch(c, nitems, KOverrelax<DD>{c->g, c->U, d, p, kind, om, cm}))));
    c->halo_ok[0] = false;
    c->g_valid = false;
  })
  return LQ_OK;
}
int lq_sweep_metropolis(lq_ctx* c, uint64_t seed, uint64_t counter, double spread, int n_update, int64_t* n_accept,
                        double* sum_prob) {
  if (!c || n_update < 1 || !(spread > 0.0 && spread < 1.0)) return LQ_E_BADARG;  // metropolis_hastings_sweep.rs:73-80
  if (!sweep_supported(c)) return LQ_E_ODD_EXTENT;
  LQ_GUARD(c);
  double acc[2] = {0.0, 0.0};
#if defined(LQ_HAVE_TUNED) && !defined(LQ_HOST_EMU)
  if (!c->odd_mask && lq_tuned_ok(c->g) && !(c->flags & LQ_FLAG_GENERIC_KERNELS)) {
    // tuned D = 4 sub-steps; their per-block statistics are reduced ONCE per sweep (one host synchronisation, one
    // all-reduce on decomposed contexts) instead of once per sub-step
    const lq_i64 nb = lq_tuned_metropolis_blocks(c->g);
    LQ_TRY(reduce_reserve(c, nb * 8, 2));
    int sub = 0;
    for (int d = 0; d < 4; ++d)
      for (int p = 0; p < 2; ++p, ++sub) {
        LQ_TRY(ensure_halo(c, 0));
        ProfScope ps(c, LQ_PROF_METROPOLIS);
        LQ_CHECK(lq_tuned_metropolis(c->stream, c->g, c->U, d, p, c->flags, n_update, c->beta, c->CA, spread, seed, counter,
                                     c->d_partial + (size_t)sub * nb * 2));
        c->launches++;
        c->halo_ok[0] = false;
        c->g_valid = false;
      }
    LQ_TRY(reduce_finish<2>(c, nb * 8));
    acc[0] = c->h_result[0];
    acc[1] = c->h_result[1];
    LQ_TRY(global_sum(c, acc, 2));
    if (n_accept) *n_accept = (int64_t)(acc[0] + 0.5);
    if (sum_prob) *sum_prob = acc[1];
    return LQ_OK;
  }
#endif
  LQ_SWEEP_LOOP(c, {
    LQ_TRY(ensure_halo(c, 0));
    ProfScope ps(c, LQ_PROF_METROPOLIS);
    LQ_DISPATCH(c, LQ_TRY((reduce(c, nitems, KMetropolis<DD>{c->g, c->U, d, p, c->flags, n_update, c->beta, c->CA, spread,
                                                            seed, counter, om, cm}))));
    acc[0] += c->h_result[0];
    acc[1] += c->h_result[1];
    c->halo_ok[0] = false;
    c->g_valid = false;
  })
  LQ_TRY(global_sum(c, acc, 2));
  if (n_accept) *n_accept = (int64_t)(acc[0] + 0.5);
  if (sum_prob) *sum_prob = acc[1];
  return LQ_OK;
}

// MetropolisHastingsDeltaDiagnostic::next_element (metropolis_hastings.rs:374-417), n_hits independent single-link hits
// per call (KMetropolisHits).  force_accept = 1: apply every proposal (MetropolisHastings::potential_next_element,
// metropolis_hastings.rs:96-118; the caller then accepts or rejects the whole state on the Hamiltonians).
int lq_metropolis_hits(lq_ctx* c, uint64_t seed, uint64_t counter, double spread, int64_t n_hits, int force_accept,
                       int64_t* n_performed, int64_t* n_accept, double* sum_prob) {
  if (!c || n_hits < 1 || n_hits > 0x7fffffff || !(spread > 0.0 && spread < 1.0)) return LQ_E_BADARG;
  if (c->decomposed) return LQ_E_BADARG;  // single-rank contexts only
  LQ_GUARD(c);
  double acc[3] = {0.0, 0.0, 0.0};
  ProfScope ps(c, LQ_PROF_METROPOLIS);
  if (c->odd_mask) {
    // lattices with an odd extent have no two-colour classes: the hits run one after the other, each on a uniformly
    // random link (the reference's own sequence of calls)
    for (int64_t h = 0; h < n_hits; ++h) {
      LQ_DISPATCH(c, LQ_TRY((reduce(c, 1, KMetropolisHits<DD>{c->g, c->U, nullptr, 1, 0, 0, c->flags, force_accept, c->beta,
                                                              c->CA, spread, seed, counter, (lq_i64)h}))));
      for (int k = 0; k < 3; ++k) acc[k] += c->h_result[k];
    }
  } else {
    const lq_i64 half = c->g.vol / 2;
    if (half > 0x7fffffff) return LQ_E_BADARG;
    LqStream pick(seed, counter, 0xFFFFFFFFFDull);
    int dir = (int)(pick.uniform01() * c->g.D);
    if (dir >= c->g.D) dir = c->g.D - 1;
    const int parity = pick.uniform01() < 0.5 ? 0 : 1;
    LQ_TRY(ensure_aos(c, (size_t)half * sizeof(int)));
    int* claim = (int*)c->d_aos;
    LQ_TRY(launch(c, half, KFillInt{claim, 0x7fffffff}));
    LQ_DISPATCH(c, LQ_TRY((launch(c, n_hits, KNoReduce<KMetropolisHits<DD>>{{c->g, c->U, claim, 0, dir, parity, c->flags,
                                                                            force_accept, c->beta, c->CA, spread, seed,
                                                                            counter, 0}}))));
    LQ_DISPATCH(c, LQ_TRY((reduce(c, n_hits, KMetropolisHits<DD>{c->g, c->U, claim, 1, dir, parity, c->flags, force_accept,
                                                                 c->beta, c->CA, spread, seed, counter, 0}))));
    for (int k = 0; k < 3; ++k) acc[k] = c->h_result[k];
  }
  c->halo_ok[0] = false;
  c->g_valid = false;
  if (n_accept) *n_accept = (int64_t)(acc[0] + 0.5);
  if (sum_prob) *sum_prob = acc[1];
  if (n_performed) *n_performed = (int64_t)(acc[2] + 0.5);
  return LQ_OK;
}

// ---------------------------------------------------------------------------------------------- HMC
int lq_snapshot(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
  if (!c->snapU) LQ_TRY(rt_malloc((void**)&c->snapU, c->u_bytes()));
  if (!c->snapE) LQ_TRY(rt_malloc((void**)&c->snapE, c->e_bytes()));
  LQ_TRY(rt_copy(c->snapU, c->U, c->u_bytes(), D2D, c->stream));
  LQ_TRY(rt_copy(c->snapE, c->E, c->e_bytes(), D2D, c->stream));
  c->snap_t = c->t;
  c->has_snap = true;
  return LQ_OK;
}
int lq_restore(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  if (!c->has_snap) return LQ_E_NOSNAPSHOT;
  LQ_GUARD(c);
  LQ_TRY(rt_copy(c->U, c->snapU, c->u_bytes(), D2D, c->stream));
  LQ_TRY(rt_copy(c->E, c->snapE, c->e_bytes(), D2D, c->stream));
  c->t = c->snap_t;
  c->halo_ok[0] = c->halo_ok[1] = false;
  c->g_valid = false;
  return LQ_OK;
}
int lq_hmc_trajectory(lq_ctx* c, double dt, int64_t n_steps, uint64_t seed, uint64_t counter, double sigma,
                      int use_current_e, int do_project, double* h_old, double* h_new, double* prob, int* accepted,
                      int64_t* gauss_steps) {
  if (!c) return LQ_E_BADARG;
  if (n_steps <= 0) return LQ_E_ZERO_STEPS;
  LQ_GUARD(c);
  if (!use_current_e) LQ_TRY(lq_momenta_refresh(c, seed, counter, sigma));  // state.rs:1093-1099
  int64_t gs = 0;
  if (do_project) LQ_TRY(lq_gauss_project(c, 0, &gs));                      // state.rs:1100
  if (gauss_steps) *gauss_steps = gs;
  // keep the old links for the reject path (hybrid_monte_carlo.rs:603-611)
  if (!c->hmcU) LQ_TRY(rt_malloc((void**)&c->hmcU, c->u_bytes()));
  {
    ProfScope ps(c, LQ_PROF_COPY);
    LQ_TRY(rt_copy(c->hmcU, c->U, c->u_bytes(), D2D, c->stream));
  }
  int64_t t0 = c->t;
  double h0, h1;
  LQ_TRY(lq_hamiltonian_total(c, &h0));
  LQ_TRY(lq_md_n(c, dt, n_steps));  // hybrid_monte_carlo.rs:573-582 (lq_symplectic_n unless lq_set_integrator chose otherwise)
  LQ_TRY(lq_hamiltonian_total(c, &h1));
  double p = fmax(fmin(exp(h0 - h1), 1.0), 0.0);  // :584-589
  LqStream acc(seed, counter, 0xFFFFFFFFFEull);
  bool ok = acc.bernoulli(p);
  if (!ok) {
    LQ_TRY(rt_copy(c->U, c->hmcU, c->u_bytes(), D2D, c->stream));
    c->halo_ok[0] = false;
    c->g_valid = false;
    c->t = t0;
  }
  if (h_old) *h_old = h0;
  if (h_new) *h_new = h1;
  if (prob) *prob = p;
  if (accepted) *accepted = ok ? 1 : 0;
  return LQ_OK;
}

// ---------------------------------------------------------------------------------------------- device peaks
// The two ceilings the rooflines are quoted against, measured on THIS device in the caller's run: f64 FMA issue rate
// (eight independent chains per thread, every SM full) and a streaming copy of the link buffer into the second one.
#ifndef LQ_HOST_EMU
}  // extern "C"
__global__ void lq_dfma_peak_k(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, cc = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
    a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 1.2345e-300) out[0] = a0;
}
__global__ void lq_copy_peak_k(const double2* __restrict__ a, double2* __restrict__ b, lq_i64 n) {
  lq_i64 i = (lq_i64)blockIdx.x * blockDim.x + threadIdx.x;
  const lq_i64 stride = (lq_i64)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}
extern "C" {
#endif
int lq_measure_peaks(lq_ctx* c, double* fp64_tflops, double* copy_gbs) {
  if (!c) return LQ_E_BADARG;
#ifdef LQ_HOST_EMU
  (void)fp64_tflops;
  (void)copy_gbs;
  return LQ_E_NODEVICE;
#else
  LQ_GUARD(c);
  int sms = 0;
  LQ_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  cudaEvent_t e0, e1;
  LQ_CHECK(cudaEventCreate(&e0));
  LQ_CHECK(cudaEventCreate(&e1));
  float ms = 0.f;
  if (fp64_tflops) {
    const int iters = 40000, blocks = sms * 8, threads = 256;
    lq_dfma_peak_k<<<blocks, threads, 0, c->stream>>>(c->d_result, 1000);
    LQ_CHECK(cudaEventRecord(e0, c->stream));
    lq_dfma_peak_k<<<blocks, threads, 0, c->stream>>>(c->d_result, iters);
    LQ_CHECK(cudaEventRecord(e1, c->stream));
    LQ_CHECK(cudaEventSynchronize(e1));
    LQ_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *fp64_tflops = 2.0 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    c->launches += 2;
  }
  if (copy_gbs) {
    LQ_TRY(ensure_buf(&c->U2, c->u_bytes(), c));
    const lq_i64 n = (lq_i64)(c->u_bytes() / sizeof(cx));
    const int reps = 5;
    lq_copy_peak_k<<<sms * 16, 256, 0, c->stream>>>(c->U, c->U2, n);
    LQ_CHECK(cudaEventRecord(e0, c->stream));
    for (int r = 0; r < reps; ++r) lq_copy_peak_k<<<sms * 16, 256, 0, c->stream>>>(c->U, c->U2, n);
    LQ_CHECK(cudaEventRecord(e1, c->stream));
    LQ_CHECK(cudaEventSynchronize(e1));
    LQ_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *copy_gbs = (double)reps * 2.0 * (double)c->u_bytes() / (ms * 1e-3) / 1e9;
    c->launches += reps + 1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  LQ_CHECK(cudaGetLastError());
  return LQ_OK;
#endif
}

// ---------------------------------------------------------------------------------------------- profiling
int lq_profile_enable(lq_ctx* c, int on) {
  if (!c) return LQ_E_BADARG;
  LQ_GUARD(c);
#ifndef LQ_HOST_EMU
  if (on && !c->prof_ev) {
    c->prof_ev = (cudaEvent_t*)calloc(2 * LQ_PROF_CAP, sizeof(cudaEvent_t));
    if (!c->prof_ev) return LQ_E_CUDA;
    for (int i = 0; i < 2 * LQ_PROF_CAP; ++i) LQ_CHECK(cudaEventCreate(&c->prof_ev[i]));
  }
#endif
  c->prof_on = on != 0;
  return lq_profile_reset(c);
}
int lq_profile_reset(lq_ctx* c) {
  if (!c) return LQ_E_BADARG;
  c->prof_n = 0;
  for (int k = 0; k < LQ_PROF_NCLASS; ++k) {
    c->prof_ms[k] = 0.0;
    c->prof_cnt[k] = 0;
  }
  return LQ_OK;
}
int lq_profile_get(lq_ctx* c, int kernel_class, int64_t* launches, double* total_ms) {
  if (!c || !launches || !total_ms) return LQ_E_BADARG;
  LQ_GUARD(c);
  *launches = 0;
  *total_ms = 0.0;
  if (kernel_class < 0 || kernel_class >= LQ_PROF_NCLASS) return LQ_E_BADARG;
  prof_drain(c);
  *launches = c->prof_cnt[kernel_class];
  *total_ms = c->prof_ms[kernel_class];
  return LQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- peer-to-peer halos
// One process per GPU; every rank maps its neighbours' field buffers (CUDA IPC) and WRITES its boundary slices
// straight into their ghost layers over NVLink -- no pack buffers, no NCCL, no host round trip.  An exchange is
//   barrier(ready) -> push kernels (one per neighbour) -> barrier(data)
// where a barrier is one tiny kernel: lane t releases a monotonically increasing epoch into its slot of neighbour
// t's flag array (st.release.sys), then spins (ld.acquire.sys) until neighbour t's epoch arrives in mine.
// "ready" orders the push after the neighbours' earlier kernels that still read their ghosts (WAR), "data" orders
// the consumers after the neighbours' pushes (RAW).  All ranks issue the same sequence of exchanges (SPMD), so
// the epochs match; a spin gives up after ~20 s and latches an error word that lq_sync / reductions report.
#ifndef LQ_HOST_EMU
struct P2pBarrierArgs {
  unsigned long long* remote[LQ_P2P_MAXNB];
  int n;
};
__global__ void lq_p2p_barrier_k(P2pBarrierArgs a, unsigned long long* mine, unsigned long long value) {
  const int t = threadIdx.x;
  if (t >= a.n) return;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.remote[t]), "l"(value) : "memory");
  const long long t0 = clock64();
  unsigned long long v;
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + t) : "memory");
    if (v >= value) break;
    if (clock64() - t0 > 40000000000ll) {
      mine[LQ_P2P_MAXNB] = value;  // error latch
      break;
    }
    __nanosleep(200);
  }
}
// boundary slice -> the neighbour's ghost slice.  region: lo[d], len[d] in storage coordinates; delta = slot shift.
template <int D>
struct KHaloPush {
  LqGeom g;
  const cx* F;
  cx* dst;
  int planes;
  int lo[LQ_MAXD], len[LQ_MAXD];
  lq_i64 nsite, delta;
  LQ_HD void operator()(lq_i64 i) const {
    lq_i64 pl = i / nsite;
    lq_i64 n = i - pl * nsite;
    if (pl >= planes) return;
    Site<D> st;
    lq_i64 q = n / len[0];
    int lane = (int)(n - q * len[0]);  // direction 0 is never split: the whole row, even x0 first
    int x0 = lane < g.ne0 ? 2 * lane : 2 * (lane - g.ne0) + 1;
    st.x[0] = x0;
    st.s = x0;
#pragma unroll
    for (int d = 1; d < D; ++d) {
      lq_i64 r = q / len[d];
      int xd = lo[d] + (int)(q - r * len[d]);
      q = r;
      st.x[d] = xd;
      st.s += (lq_i64)xd * g.sstride[d];
    }
    lq_i64 p = lq_slot<D>(g, st);
    dst[lq_addr(p + delta, planes, (int)pl)] = F[lq_addr(p, planes, (int)pl)];
  }
};
static int p2p_barrier(lq_ctx* c) {
  P2pBarrierArgs a;
  a.n = c->p2p_nnb;
  for (int k = 0; k < c->p2p_nnb; ++k)
    a.remote[k] = (unsigned long long*)c->p2p_base[c->p2p_peer[k]][LQ_P2P_NBUF - 1] + c->p2p_rev[k];
  c->p2p_epoch += 1;
  c->p2p_pending = false;  // a full barrier is behind every epoch released before it
  lq_p2p_barrier_k<<<1, 32, 0, c->stream>>>(a, c->p2p_flags, c->p2p_epoch);
  c->launches++;
  LQ_CHECK(cudaGetLastError());
  return LQ_OK;
}
static int p2p_exchange(lq_ctx* c, int which) {
  cx* f;
  int planes;
  switch (which) {
    case 0: f = c->U; planes = 9 * c->g.D; break;
    case 1: f = c->E; planes = 4 * c->g.D; break;
    case 2: f = c->G; planes = 9; break;
    default: return LQ_E_BADARG;
  }
  if (!f) return LQ_E_BADARG;
  int bi = -1;  // which of my allocations is it: the neighbours' current buffer is their allocation of the same index
  for (int b = 0; b < LQ_P2P_NBUF - 1; ++b)
    if (c->own[b] == f) bi = b;
  if (bi < 0) return LQ_E_COMM;
  LQ_TRY(p2p_barrier(c));  // ready: the neighbours' earlier kernels are done with their ghosts
  const LqGeom& g = c->g;
  for (int k = 0; k < c->p2p_nnb; ++k) {
    int lo[LQ_MAXD], len[LQ_MAXD];
    lq_i64 ns = 1, delta = 0;
    for (int d = 0; d < LQ_MAXD; ++d) {
      int o = d < g.D ? c->p2p_off[k][d] : 0;
      if (d >= g.D) {
        lo[d] = 0;
        len[d] = 1;
      } else if (o < 0) {  // my first interior slice -> their high ghost
        lo[d] = 1;
        len[d] = 1;
        delta += (lq_i64)g.ext[d] * g.sstride[d];
      } else if (o > 0) {  // my last interior slice -> their low ghost
        lo[d] = g.ext[d];
        len[d] = 1;
        delta -= (lq_i64)g.ext[d] * g.sstride[d];
      } else {
        lo[d] = g.ghost[d];
        len[d] = g.ext[d];
      }
      ns *= len[d];
    }
    cx* dst = (cx*)c->p2p_base[c->p2p_peer[k]][bi];
#define LQ_PUSH(DD_)                                                                                   \
  {                                                                                                    \
    KHaloPush<DD_> kp;                                                                                 \
    kp.g = g; kp.F = f; kp.dst = dst; kp.planes = planes; kp.nsite = ns; kp.delta = delta;            \
    for (int d = 0; d < LQ_MAXD; ++d) { kp.lo[d] = lo[d]; kp.len[d] = len[d]; }                       \
    LQ_TRY(launch(c, ns * planes, kp));                                                                \
  }
    switch (g.D) {
      case 2: LQ_PUSH(2) break;
      case 3: LQ_PUSH(3) break;
      case 4: LQ_PUSH(4) break;
      default: return LQ_E_BADARG;
    }
#undef LQ_PUSH
  }
  LQ_TRY(p2p_barrier(c));  // data: the neighbours' pushes into my ghosts have landed
  c->p2p_exchanges++;
  return LQ_OK;
}
#else
static int p2p_exchange(lq_ctx*, int) { return LQ_E_COMM; }
#endif

extern "C" {
int lq_p2p_export(lq_ctx* c, void* handles_out, int64_t bytes) {
  if (!c || !handles_out || bytes != LQ_P2P_NBUF * LQ_P2P_HANDLE) return LQ_E_BADARG;
#ifdef LQ_HOST_EMU
  return LQ_E_NODEVICE;
#else
  if (!c->decomposed) return LQ_E_BADARG;
  LQ_GUARD(c);
  // every buffer a neighbour may have to write must exist before the handles travel
  LQ_TRY(ensure_buf(&c->U2, c->u_bytes(), c));
  LQ_TRY(ensure_buf(&c->E2, c->e_bytes(), c));
  LQ_TRY(ensure_buf(&c->G, c->g_bytes(), c));
  LQ_TRY(ensure_buf(&c->G2, c->g_bytes(), c));
  LQ_TRY(ensure_buf(&c->T, c->t_bytes(), c));
  LQ_TRY(ensure_buf(&c->T2, c->t_bytes(), c));
  if (!c->p2p_flags) {
    LQ_TRY(rt_malloc((void**)&c->p2p_flags, 2 * LQ_P2P_MAXNB * sizeof(unsigned long long)));
    LQ_TRY(rt_memset(c->p2p_flags, 0, 2 * LQ_P2P_MAXNB * sizeof(unsigned long long), c->stream));
  }
  LQ_TRY(rt_sync(c->stream));
  cx* bufs[LQ_P2P_NBUF - 1] = {c->U, c->U2, c->E, c->E2, c->G, c->G2, c->T, c->T2};
  for (int b = 0; b < LQ_P2P_NBUF - 1; ++b) c->own[b] = bufs[b];
  static_assert(sizeof(cudaIpcMemHandle_t) == LQ_P2P_HANDLE, "IPC handle size");
  cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles_out;
  for (int b = 0; b < LQ_P2P_NBUF - 1; ++b) LQ_CHECK(cudaIpcGetMemHandle(&h[b], bufs[b]));
  LQ_CHECK(cudaIpcGetMemHandle(&h[LQ_P2P_NBUF - 1], c->p2p_flags));
  return LQ_OK;
#endif
}
int lq_p2p_attach(lq_ctx* c, int n_peers, const void* peer_handles, int n_neighbors, const int* offsets,
                  const int* peer_index) {
  if (!c || !peer_handles || !offsets || !peer_index) return LQ_E_BADARG;
#ifdef LQ_HOST_EMU
  return LQ_E_NODEVICE;
#else
  if (!c->decomposed || !c->p2p_flags || n_peers < 1 || n_peers > LQ_P2P_MAXNB || n_neighbors < 1 ||
      n_neighbors > LQ_P2P_MAXNB)
    return LQ_E_BADARG;
  // one attach per context: a second one would reuse flag slots that still hold the epochs of the first (every barrier
  // would pass at once) and leak the mappings already opened
  if (c->p2p_on || c->p2p_npeers) return LQ_E_BADARG;
  LQ_GUARD(c);
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)peer_handles;
  for (int q = 0; q < n_peers; ++q)
    for (int b = 0; b < LQ_P2P_NBUF; ++b)
      LQ_CHECK(cudaIpcOpenMemHandle(&c->p2p_base[q][b], h[q * LQ_P2P_NBUF + b], cudaIpcMemLazyEnablePeerAccess));
  c->p2p_npeers = n_peers;
  // neighbour slots are numbered by their offset vector, identically on every rank: slot(o) = sum (o_d + 1) 3^d
  // over the directions; a neighbour at offset o finds me at offset -o.
  for (int k = 0; k < n_neighbors; ++k) {
    if (peer_index[k] < 0 || peer_index[k] >= n_peers) return LQ_E_BADARG;
    c->p2p_peer[k] = peer_index[k];
    for (int d = 0; d < LQ_MAXD; ++d) c->p2p_off[k][d] = d < c->g.D ? offsets[k * c->g.D + d] : 0;
  }
  for (int k = 0; k < n_neighbors; ++k) {
    int rev = -1;
    for (int j = 0; j < n_neighbors; ++j) {
      bool opp = true;
      for (int d = 0; d < c->g.D; ++d) opp = opp && c->p2p_off[j][d] == -c->p2p_off[k][d];
      if (opp) rev = j;
    }
    if (rev < 0) return LQ_E_BADARG;  // the neighbour list must be closed under negation, in the same order everywhere
    c->p2p_rev[k] = rev;
  }
  c->p2p_nnb = n_neighbors;
  c->p2p_epoch = 0;
  if (c->g.svol * 36 < ((lq_i64)1 << 31)) {  // peer tables of the fused "compute + halo push" kernels, one per buffer
    const int D = c->g.D;
    LqPush tab[LQ_P2P_NBUF - 1];
    memset(tab, 0, sizeof(tab));
    for (int b = 0; b < LQ_P2P_NBUF - 1; ++b) {
      for (int i = 0; i < 9; ++i) (&tab[b].nbmap[0][0])[i] = -1;
      for (int k = 0; k < n_neighbors; ++k) {
        lq_i64 delta = 0;
        for (int d = 0; d < D; ++d) delta -= (lq_i64)c->p2p_off[k][d] * c->g.ext[d] * c->g.sstride[d];
        tab[b].peer[k] = (cx*)c->p2p_base[c->p2p_peer[k]][b];
        tab[b].delta[k] = (int)delta;
        tab[b].nbmap[(D >= 3 ? c->p2p_off[k][D - 2] : 0) + 1][c->p2p_off[k][D - 1] + 1] = k;
      }
    }
    if (!c->d_fold_counter) {
      LQ_TRY(rt_malloc((void**)&c->d_fold_counter, sizeof(unsigned int)));
      LQ_TRY(rt_memset(c->d_fold_counter, 0, sizeof(unsigned int), c->stream));
    }
    if (!c->d_push) LQ_TRY(rt_malloc(&c->d_push, sizeof(tab)));
    LQ_TRY(rt_copy(c->d_push, tab, sizeof(tab), H2D, c->stream));
    LQ_TRY(rt_sync(c->stream));
  }
  c->p2p_on = true;
  return LQ_OK;
#endif
}
int lq_p2p_enabled(const lq_ctx* c) { return c && c->p2p_on ? 1 : 0; }
int64_t lq_p2p_exchanges(const lq_ctx* c) { return c ? c->p2p_exchanges : 0; }

// ---------------------------------------------------------------------------------------------- halos
static int halo_field(lq_ctx* c, int which, cx** f, int* planes) {
  switch (which) {
    case 0: *f = c->U; *planes = 9 * c->g.D; return LQ_OK;
    case 1: *f = c->E; *planes = 4 * c->g.D; return LQ_OK;
    case 2: *f = c->G; *planes = 9; return c->G ? LQ_OK : LQ_E_BADARG;
    default: return LQ_E_BADARG;
  }
}
static lq_i64 face_sites(const lq_ctx* c, int dir) { return c->g.svol / c->g.sext[dir]; }
int lq_halo_bytes(lq_ctx* c, int which, int dir, int64_t* bytes) {
  if (!c || !bytes || dir < 0 || dir >= c->g.D || !c->g.ghost[dir]) return LQ_E_BADARG;
  int planes = which == 0 ? 9 * c->g.D : which == 1 ? 4 * c->g.D : which == 2 ? 9 : -1;
  if (planes < 0) return LQ_E_BADARG;
  *bytes = (int64_t)(face_sites(c, dir) * planes * sizeof(cx));
  return LQ_OK;
}
int lq_halo_pack(lq_ctx* c, int which, int dir, int side, void* d_buf, int64_t bytes) {
  int64_t need;
  LQ_TRY(lq_halo_bytes(c, which, dir, &need));
  if (!d_buf || bytes != need || (side != 0 && side != 1)) return LQ_E_BADARG;
  LQ_GUARD(c);
  cx* f;
  int planes;
  LQ_TRY(halo_field(c, which, &f, &planes));
  lq_i64 nface = face_sites(c, dir);
  int xh = side == 0 ? 1 : c->g.ext[dir];  // interior boundary slices
  LQ_DISPATCH(c, LQ_TRY((launch(c, nface * planes, KHaloPack<DD>{c->g, f, (cx*)d_buf, dir, xh, planes, nface}))));
  return LQ_OK;
}
int lq_halo_unpack(lq_ctx* c, int which, int dir, int side, const void* d_buf, int64_t bytes) {
  int64_t need;
  LQ_TRY(lq_halo_bytes(c, which, dir, &need));
  if (!d_buf || bytes != need || (side != 0 && side != 1)) return LQ_E_BADARG;
  LQ_GUARD(c);
  cx* f;
  int planes;
  LQ_TRY(halo_field(c, which, &f, &planes));
  lq_i64 nface = face_sites(c, dir);
  int xh = side == 0 ? 0 : c->g.ext[dir] + 1;  // ghost slices
  LQ_DISPATCH(c, LQ_TRY((launch(c, nface * planes, KHaloUnpack<DD>{c->g, f, (const cx*)d_buf, dir, xh, planes, nface}))));
  return LQ_OK;
}
int lq_halo_invalidate(lq_ctx* c, int which) {
  if (!c || which < 0 || which > 2) return LQ_E_BADARG;
  c->halo_ok[which] = false;
  return LQ_OK;
}

}  // extern "C"
