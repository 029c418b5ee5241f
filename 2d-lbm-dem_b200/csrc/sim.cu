/*
 * sim.cu -- host side of liblbmdem_gpu.so: the simulation object that owns the device state
 * and sequences the kernels the way renderScene() sequences the reference's loops
 * (src/main.c:1697-1777), plus the C ABI of include/lbmdem_gpu.h.
 *
 * Device layout (DESIGN.md "data layout"):
 *   f[2]     double-buffered populations, structure of arrays [q][x - x0][y], y contiguous, rows
 *            padded to a multiple of 32 elements (128-byte aligned rows, TMA-legal strides).
 *            Between LBM steps f[cur] holds the populations of the last step as the reference
 *            holds them just before its swap passes (holds_A: re-init, collide, ring and
 *            bounce-back sweeps applied); the reference's observable f[x][y][q] -- the same
 *            array after streaming -- is materialised into the other buffer only when somebody
 *            asks for it (get_f, fields, density).
 *   cell[2]  obstacle map [x - x0][y] (int32: -1 fluid, grain index | act bit, nbgrains = wall
 *            ring) of the last two steps, with the grain records rec[2]/R2[2]/boxes[2] that
 *            produced them: the fused kernel streams with the stored step's map and records
 *            and collides with the new ones (reinit_obst_density reads the old map, :970)
 *   grains   structure of arrays of `real`, replicated on every rank
 * Strip decomposition: rank k of P owns the global rows [xlo, xhi); with P > 1 the local arrays
 * carry GHOST = 4 extra rows on each side (x0 = xlo - 4).  The populations of the ghost rows are
 * exchanged once per LBM step, right after the fused kernel; the obstacle map of the ghost rows
 * is recomputed locally (grains are replicated).  Each rank then applies the ring and bounce-back
 * sweeps to its owned rows plus one ghost row per side, which is all the next pull reads.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lbmdem_gpu.h"
#include "kernels.h"

namespace lbmdem {

using namespace lbm;

static thread_local std::string g_create_error;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return fail(LBMDEM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

/* ---- NCCL, bound at run time so that single-GPU use has no NCCL dependency ---- */
struct NcclApi {
  struct UniqueId { char internal[128]; };
  typedef int (*GetUniqueId_t)(UniqueId *);
  typedef int (*CommInitRank_t)(void **, int, UniqueId, int);
  typedef int (*CommDestroy_t)(void *);
  typedef int (*Send_t)(const void *, size_t, int, int, void *, cudaStream_t);
  typedef int (*Recv_t)(void *, size_t, int, int, void *, cudaStream_t);
  typedef int (*Group_t)(void);
  typedef int (*AllReduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef int (*Broadcast_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef int (*AllGather_t)(const void *, void *, size_t, int, void *, cudaStream_t);
  typedef const char *(*ErrStr_t)(int);
  void *handle = nullptr;
  GetUniqueId_t GetUniqueId = nullptr;
  CommInitRank_t CommInitRank = nullptr;
  CommDestroy_t CommDestroy = nullptr;
  Send_t Send = nullptr;
  Recv_t Recv = nullptr;
  Group_t GroupStart = nullptr, GroupEnd = nullptr;
  AllReduce_t AllReduce = nullptr;
  Broadcast_t Broadcast = nullptr;
  AllGather_t AllGather = nullptr;
  ErrStr_t GetErrorString = nullptr;
  enum { Int64 = 4, Float32 = 7, Float64 = 8, Sum = 0 };
  bool load(std::string *why) {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { *why = std::string("dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name)                                  \
  field = (decltype(field))dlsym(handle, name);           \
  if (!field) { *why = std::string("dlsym ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce") SYM(Broadcast, "ncclBroadcast") SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
  }
};
static NcclApi g_nccl;

/* ---- in-process strip group: the ranks of a decomposed run as contexts of ONE process (one per GPU, or several on
 * one GPU -- SURVEY 4.5's "fake multi-GPU"), each driven by its own host thread.  No collective library: a rank PULLS
 * its ghost rows from the neighbour's device memory (cudaMemcpy2DAsync, peer copy across devices) and adds the force
 * sums with one kernel that reads every peer's partial sums.  The host threads meet at a barrier wherever one rank's
 * stream has to wait for an event another rank records in the same step (an event must be recorded before a
 * cudaStreamWaitEvent on it means anything). ---- */
struct LocalGroup {
  struct Slot {
    int device = 0;
    const void *f_cur = nullptr;    /* population buffer that holds this step's state */
    size_t plane = 0;
    int pitch = 0, x0 = 0, xlo = 0, xhi = 0;
    const void *partial = nullptr;  /* this step's partial force sums (long long, or double in the strict build) */
    cudaEvent_t ev_k1 = nullptr, ev_pulled = nullptr, ev_partial = nullptr;
    bool attached = false;
  };
  int P;
  std::vector<Slot> slot;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long gen = 0;
  bool aborted = false;
  explicit LocalGroup(int p) : P(p), slot((size_t)p) {}
  /* false when a member failed meanwhile: the caller gives up instead of waiting for a peer that will not come */
  bool barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (aborted) return false;
    const long my = gen;
    if (++arrived == P) {
      arrived = 0;
      ++gen;
      cv.notify_all();
      return true;
    }
    /* a peer that never arrives (its driver thread died, or the ranks were stepped unequally) must not hang the run */
    if (!cv.wait_for(lk, std::chrono::seconds(300), [&] { return gen != my || aborted; })) {
      aborted = true;
      cv.notify_all();
    }
    return !aborted;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(m);
    aborted = true;
    cv.notify_all();
  }
};

/* makes the context's device current for the duration of an API call, whatever thread calls */
struct DeviceGuard {
  int prev = -1, want;
  explicit DeviceGuard(int d) : want(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != want) cudaSetDevice(want);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != want) cudaSetDevice(prev);
  }
};

typedef CUresult (*EncodeTiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct SimBase {
  lbmdem_params P;
  std::string err;
  int last_code = 0;
  virtual ~SimBase() {}
  int fail(int code, const std::string &msg) {
    err = msg;
    last_code = code;
    return code;
  }
  virtual int init_device() = 0;
  virtual int load_sample(const char *path) = 0;
  virtual int set_grains(int n, const double *r, const double *x1, const double *x2) = 0;
  virtual int step(long n) = 0;
  virtual int step_capture(double *mid) = 0;
  virtual int lbm_step() = 0;
  virtual int lbm_steps(long n) = 0;
  virtual int build_verlet() = 0;
  virtual int get_scalars(double *d, long *l) = 0;
  virtual int set_nbsteps(long n) = 0;
  virtual int get_strip(int *xlo, int *xhi) = 0;
  virtual int total_density(double *sum) = 0;
  virtual int get_f(double *out) = 0;
  virtual int set_f(const double *in) = 0;
  virtual int get_obst(int *out) = 0;
  virtual int set_obst(const int *in) = 0;
  virtual int get_act(int *out) = 0;
  virtual int get_grains(double *out) = 0;
  virtual int set_grain_state(const double *in) = 0;
  virtual int get_fhf(double *out) = 0;
  virtual int set_fhf(const double *in) = 0;
  virtual int get_verlet(int *count, int *nbr, int capacity, int *wall_flags) = 0;
  virtual int save_state(const char *path) = 0;
  virtual int load_state(const char *path) = 0;
  virtual int get_fields(const double *gp, float *a, float *b, float *c, float *d, float *e) = 0;
  virtual int step_host(const void *state_in, long n, void *state_out, void *fhf_out, double *dens, bool rows_f32,
                        bool share) = 0;
  virtual int get_share(int *i0, int *i1) = 0;
  virtual int attach_nccl(const void *id) = 0;
  virtual int attach_local(LocalGroup *g) = 0;
  virtual int get_kernel_timer(double *ms, long *k1, long *all) = 0;
  virtual int reset_kernel_timer(int enable) = 0;
  virtual int get_list_counts(long *c) = 0;
  virtual int state_checksum(unsigned long long *c) = 0;
  virtual void *stream_ptr() = 0;
};

template <typename real>
struct Sim : SimBase {
  /* geometry */
  int lx = 0, ly = 0, xlo = 0, xhi = 0, x0 = 0, nxl = 0, pitch = 0;
  size_t plane = 0;
  int n = 0; /* nbgrains */
  bool ready = false;
  /* reference globals, kept in `real` like the reference keeps them (src/main.c:52-165) */
  real dx = 0, dtLB = 0, c = 0, dt = 0, dt2 = 0, Mgx = 0, Mdx = 0, Mby = 0, Mhy = 0, xG = 0, yG = 0, rmax = 0;
  real t = 0; /* src/main.c:165, advanced only while the walls vibrate */
  int npDEM = 1;
  long nbsteps = 0, nFile = 0;
  double k12 = 0, k3 = 0; /* fhf scale factors (:1329-1331) */
  /* device */
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;            /* halo exchange of a strip-decomposed run, overlapped with the interior sweep */
  cudaEvent_t ev_state = nullptr, ev_halo = nullptr;
  real *f[2] = {nullptr, nullptr};
  int *cell[2] = {nullptr, nullptr};
  unsigned char *cls[2] = {nullptr, nullptr}; /* class byte per node of cell[k] (lbm_node.cuh cell_class): what the row kernel streams */
  unsigned short *own16[2] = {nullptr, nullptr}; /* 16-bit owner per node of cell[k] (cell_own16): the stored step's map as streamed */
  CUtensorMap tmC16[2];
  int cur = 0, cur_cell = 0;
  bool holds_A = false;      /* f[cur] holds A of the last step (stream pending) instead of f */
  bool scratch_valid = false; /* f[1 - cur] holds the materialised f of the pending stream */
  Lattice<real> L_fused;      /* lattice() as the last fused launch saw it */
  bool dead_stale = false;    /* f[cur] lacks the populations the fused kernel does not write (lbm_node.cuh, node_is_dead) */
  CUtensorMap tmA[2], tmC[2], tmCh[2]; /* populations; map rows without / with the y halo */
  LinkList llist{};
  DeferList<real> defer{};
  BoundaryList blist{};
  TileBins tbins{};          /* grains binned by lattice tile, rebuilt every LBM step (kernels.h) */
  ForceFinish fin_pending{nullptr, 1, 0, 0, nullptr};
  int *range_flag_dev = nullptr; /* device view of hflags[5] */ /* fhf1..3 still have to be derived from these sums (kernels.h ForceFinish) */
  int coop_cap = 0;          /* grains the cooperative DEM kernel can take (one co-resident grid) */
  int sm_count = 148;        /* cudaDevAttrMultiProcessorCount of the context's device: sizes the persistent grids */
  int raster_step = 1;       /* counts rasteriser runs: tile stamps are compared with it (kernels.h TileBins::stamp) */
  int raster_full_until = 0; /* runs up to this one rebuild every tile (set-up, state or map set from outside) */
  int *hflags = nullptr;      /* mapped host memory: [0] Verlet capacity exceeded, [1] deferred-link list full, [2] boundary list full */
  std::vector<real *> grain_bufs;
  GrainArrays<real> g{};
  GrainRec<real> *rec[2] = {nullptr, nullptr}; /* indexed like cell[] */
  real *R2[2] = {nullptr, nullptr};
  GrainBox *boxes[2] = {nullptr, nullptr};
  EncodeTiled_t encode = nullptr;
  long long *facc = nullptr;      /* this step's fixed-point force sums: facc_buf[fslot] */
  double *fpartial = nullptr;     /* strict build: this step's fp64 partial sums: fpartial_buf[fslot] */
  long long *facc_buf[2] = {nullptr, nullptr};
  double *fpartial_buf[2] = {nullptr, nullptr};
  void *fsum = nullptr;           /* in-process group: the sums over all ranks (3 n long long / double) */
  int fslot = 0;                  /* toggles every LBM step of a group run: peers may still read the previous step's sums */
  LocalGroup *group = nullptr;
  bool peers_ready = false;
  /* NCCL transport: the force sums are added through CUDA IPC mappings of the peers' buffers (kernels.h IpcPeers)
   * instead of ncclAllReduce -- 0.464 ms against 0.483 ms per step on 8 GPUs; LBMDEM_PEER_SUMS=0 or any failure to map
   * a peer keeps the all-reduce */
  bool ipc_wanted = false, ipc_ready = false, ipc_tried = false;
  void *ipc_peer_facc[2][MAX_LOCAL_RANKS] = {};
  unsigned *ipc_peer_flags[MAX_LOCAL_RANKS] = {};
  unsigned *ipc_flags = nullptr; /* this rank's arrival words (written by the peers) */
  unsigned ipc_step = 0;
  int *ipc_timeout_dev = nullptr; /* device view of hflags[7] */
  VerletBuffers vb{};
  double *dens_partials = nullptr, *dens_out = nullptr;
  double *stage = nullptr; /* device staging for layout conversion */
  size_t stage_elems = 0;
  double *hstage = nullptr; /* pinned host staging for grain I/O */
  size_t hstage_elems = 0;
  /* NCCL */
  void *comm = nullptr;
  /* instrumentation */
  bool events_on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  size_t ev_used = 0;
  double k1_ms_acc = 0;
  long k1_launches = 0, all_launches = 0;

  ~Sim() override {
    ipc_close();
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    for (auto &e : ev_pool) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (int k = 0; k < 2; ++k) { cudaFree(f[k]); cudaFree(cell[k]); cudaFree(cls[k]); cudaFree(own16[k]); }
    for (real *p : grain_bufs) cudaFree(p);
    for (int k = 0; k < 2; ++k) { cudaFree(rec[k]); cudaFree(R2[k]); cudaFree(boxes[k]); }
    for (int k = 0; k < 2; ++k) { cudaFree(facc_buf[k]); cudaFree(fpartial_buf[k]); }
    cudaFree(fsum);
    cudaFree(vb.bucket_count); cudaFree(vb.bucket_cursor); cudaFree(vb.sorted); cudaFree(vb.gcx); cudaFree(vb.gcy);
    cudaFree(vb.nbr_count); cudaFree(vb.nbr); cudaFree(vb.wflags);
    cudaFree(dens_partials); cudaFree(dens_out); cudaFree(stage); cudaFree(mid_dev); cudaFree(gstage);
    cudaFree(defer.count); cudaFree(defer.index); cudaFree(defer.value);
    cudaFree(blist.entry); cudaFree(blist.tcount); cudaFree(llist.entry); cudaFree(llist.tcount);
    cudaFree(tbins.count); cudaFree(tbins.list); cudaFree(tbins.stamp);
    if (hflags) cudaFreeHost(hflags);
    if (hstage) cudaFreeHost(hstage);
    if (ev_state) cudaEventDestroy(ev_state);
    if (ev_halo) cudaEventDestroy(ev_halo);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    if (stream) cudaStreamDestroy(stream);
  }

  int DENS_BLOCKS = 1184;                  /* 8 x the device's SM count (init_device) */
  static constexpr int GHOST = 4;          /* extra rows per side of a strip */

  int init_device() override {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return fail(LBMDEM_ECUDA, "no CUDA device visible: liblbmdem_gpu has no CPU path");
    if (P.device < 0 || P.device >= ndev) return fail(LBMDEM_EINVAL, "device ordinal out of range");
    CK(cudaSetDevice(P.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, P.device));
    sm_count = prop.multiProcessorCount;
    DENS_BLOCKS = 8 * sm_count;
    if (prop.cooperativeLaunch) CK(dem_coop_capacity<real>(&coop_cap));
    if (prop.major < 10)
      return fail(LBMDEM_ECUDA, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                    "; the kernels are built for sm_100a only");
    lx = P.lx; ly = P.ly;
    if (lx < 8 || ly < 8) return fail(LBMDEM_EINVAL, "lattice must be at least 8 x 8");
    if (P.nranks < 1 || P.rank < 0 || P.rank >= P.nranks) return fail(LBMDEM_EINVAL, "bad rank / nranks");
    const int base = lx / P.nranks, rem = lx % P.nranks;
    xlo = P.rank * base + std::min(P.rank, rem);
    xhi = xlo + base + (P.rank < rem ? 1 : 0);
    if (P.nranks > 1 && xhi - xlo < 4) return fail(LBMDEM_EINVAL, "strips must be at least 4 rows wide");
    if (P.nranks > 1) { x0 = xlo - GHOST; nxl = xhi - xlo + 2 * GHOST; } else { x0 = 0; nxl = lx; }
    pitch = (ly + 31) / 32 * 32;
#ifndef LBMDEM_PLANE_PAD
#define LBMDEM_PLANE_PAD 0 /* tuning knob: extra elements between the population planes (de-aliases power-of-two strides) */
#endif
    plane = (size_t)nxl * pitch + LBMDEM_PLANE_PAD;
    /* node indices within a plane are 32-bit (list entries, the fused kernel's store offsets) */
    if (plane >= ((size_t)1 << 31)) return fail(LBMDEM_EINVAL, "more than 2^31 nodes per GPU: use more strips");
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (P.nranks > 1) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&comm_stream, cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&ev_state, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_halo, cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; ++k) {
      CK(cudaMalloc(&f[k], sizeof(real) * plane * NQ));
      CK(cudaMemsetAsync(f[k], 0, sizeof(real) * plane * NQ, stream));
      CK(cudaMalloc(&cell[k], sizeof(int) * plane));
      CK(cudaMalloc(&cls[k], plane));
      CK(cudaMalloc(&own16[k], sizeof(unsigned short) * plane));
    }
    CK(cudaMalloc(&dens_partials, sizeof(double) * DENS_BLOCKS));
    CK(cudaMalloc(&dens_out, sizeof(double)));
    /* TMA descriptors: 3-D (y, x, q) view of each population buffer, one lattice row per box;
     * 2-D (y, x) views of each obstacle map with and without the y halo */
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(LBMDEM_ECUDA, "cuTensorMapEncodeTiled not available");
    encode = (EncodeTiled_t)fn;
    using C = RowCfg<real>;
    for (int k = 0; k < 2; ++k) {
      const cuuint64_t dims[3] = {(cuuint64_t)ly, (cuuint64_t)nxl, (cuuint64_t)NQ};
      const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(real), (cuuint64_t)plane * sizeof(real)};
      const cuuint32_t box[3] = {(cuuint32_t)C::BY, 1, (cuuint32_t)NQ};
      const cuuint32_t estr[3] = {1, 1, 1};
#ifndef LBMDEM_K1_L2PROMO
#define LBMDEM_K1_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_256B /* tuning knob: ..._NONE / _L2_64B / _L2_128B / _L2_256B */
#endif
      CUresult r = encode(&tmA[k], sizeof(real) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                          f[k], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          LBMDEM_K1_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(LBMDEM_ECUDA, "cuTensorMapEncodeTiled(f) failed with code " + std::to_string((int)r));
      /* the stored step's map (int32, no halo) and the class bytes of this step's (one byte per node, y halo) */
      const cuuint64_t cdims[2] = {(cuuint64_t)ly, (cuuint64_t)nxl};
      const cuuint64_t cstr[1] = {(cuuint64_t)pitch * sizeof(int)};
      const cuuint32_t cbox[2] = {(cuuint32_t)C::TY, 1};
      r = encode(&tmC[k], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, cell[k], cdims, cstr, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const cuuint64_t bstr[1] = {(cuuint64_t)pitch};
      const cuuint32_t cboxh[2] = {(cuuint32_t)C::BC, 1};
      if (r == CUDA_SUCCESS)
        r = encode(&tmCh[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, cls[k], cdims, bstr, cboxh, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const cuuint64_t sstr[1] = {(cuuint64_t)pitch * sizeof(unsigned short)};
      if (r == CUDA_SUCCESS)
        r = encode(&tmC16[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, own16[k], cdims, sstr, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(LBMDEM_ECUDA, "cuTensorMapEncodeTiled(map) failed with code " + std::to_string((int)r));
    }
    CK(cudaHostAlloc(&hflags, 8 * sizeof(int), cudaHostAllocMapped));
    for (int k = 0; k < 8; ++k) hflags[k] = 0;
    CK(cudaHostGetDevicePointer(&range_flag_dev, hflags + 5, 0));
    CK(cudaHostGetDevicePointer(&ipc_timeout_dev, hflags + 7, 0));
    {
      const char *e = getenv("LBMDEM_PEER_SUMS"); /* 0: keep the force sums on ncclAllReduce */
      ipc_wanted = !(e && atoi(e) == 0);
    }
    CK(cudaStreamSynchronize(stream));
    return 0;
  }

  /* ---- grains ---- */
  /* frees and forgets: a later allocation failure must not leave a dangling pointer for the destructor */
  template <typename T>
  static void dfree(T *&p) {
    cudaFree(p);
    p = nullptr;
  }
  int alloc_grains(int n_) {
    if (n_ <= 0) return fail(LBMDEM_EINVAL, "need at least one grain (the reference reads g[0], src/main.c:220)");
    if (ipc_ready) {
      /* the peers map this rank's force-sum buffers: every rank unmaps (loading grains is collective over the ranks)
       * before any rank frees */
      ipc_close();
      long long *vd = nullptr;
      CK(cudaMalloc(&vd, sizeof(long long)));
      CK(cudaMemsetAsync(vd, 0, sizeof(long long), stream));
      const int r = g_nccl.AllReduce(vd, vd, 1, NcclApi::Int64, NcclApi::Sum, comm, stream);
      CK(cudaStreamSynchronize(stream));
      cudaFree(vd);
      if (r) return nccl_fail(r, "ncclAllReduce(ipc release)");
    }
    ipc_tried = false;
    for (real *p : grain_bufs) cudaFree(p);
    grain_bufs.clear();
    dfree(mid_dev);
    dfree(gstage);
    n = n_;
    /* one slab, so that the kinematic state (9 arrays) and state + fhf (12 arrays) move in one copy each:
     * x1 x2 x3 v1 v2 v3 a1 a2 a3 | fhf1 fhf2 fhf3 | r m It rLB */
    real *slab = nullptr;
    CK(cudaMalloc(&slab, sizeof(real) * 16 * (size_t)n));
    CK(cudaMemsetAsync(slab, 0, sizeof(real) * 16 * (size_t)n, stream));
    grain_bufs.push_back(slab);
    real **slots[] = {&g.x1, &g.x2, &g.x3, &g.v1, &g.v2, &g.v3, &g.a1, &g.a2, &g.a3, &g.fhf1, &g.fhf2, &g.fhf3,
                      &g.r, &g.m, &g.It, &g.rLB};
    for (size_t k = 0; k < 16; ++k) *slots[k] = slab + k * (size_t)n;
    for (int k = 0; k < 2; ++k) {
      dfree(rec[k]); dfree(R2[k]); dfree(boxes[k]);
      CK(cudaMalloc(&rec[k], sizeof(GrainRec<real>) * n));
      CK(cudaMalloc(&R2[k], sizeof(real) * n));
      CK(cudaMalloc(&boxes[k], sizeof(GrainBox) * n));
    }
    dfree(fsum);
    for (int k = 0; k < 2; ++k) {
      dfree(facc_buf[k]); dfree(fpartial_buf[k]);
      CK(cudaMalloc(&facc_buf[k], sizeof(long long) * 3 * n));
      CK(cudaMemsetAsync(facc_buf[k], 0, sizeof(long long) * 3 * n, stream));
      CK(cudaMalloc(&fpartial_buf[k], sizeof(double) * 3 * n));
      CK(cudaMemsetAsync(fpartial_buf[k], 0, sizeof(double) * 3 * n, stream));
    }
    CK(cudaMalloc(&fsum, sizeof(double) * 3 * n));
    fslot = 0;
    facc = facc_buf[0];
    fpartial = fpartial_buf[0];
    dfree(vb.bucket_count); dfree(vb.bucket_cursor); dfree(vb.sorted); dfree(vb.gcx); dfree(vb.gcy);
    dfree(vb.nbr_count); dfree(vb.nbr); dfree(vb.wflags);
    int nb = 1024;
    while (nb < 2 * n) nb <<= 1;
    vb.nbuckets = nb;
    vb.cap = P.neighbour_capacity > 0 ? P.neighbour_capacity : 32;
    CK(cudaMalloc(&vb.bucket_count, sizeof(int) * (nb + 1)));
    CK(cudaMalloc(&vb.bucket_cursor, sizeof(int) * nb));
    CK(cudaMalloc(&vb.sorted, sizeof(int) * n));
    CK(cudaMalloc(&vb.gcx, sizeof(int) * n));
    CK(cudaMalloc(&vb.gcy, sizeof(int) * n));
    CK(cudaMalloc(&vb.nbr_count, sizeof(int) * n));
    CK(cudaMemsetAsync(vb.nbr_count, 0, sizeof(int) * n, stream));
    CK(cudaMalloc(&vb.nbr, sizeof(int) * (size_t)n * vb.cap));
    CK(cudaMalloc(&vb.wflags, sizeof(int) * n));
    CK(cudaMemsetAsync(vb.wflags, 0, sizeof(int) * n, stream));
    CK(cudaHostGetDevicePointer(&vb.error, hflags, 0));
    dfree(defer.count); dfree(defer.index); dfree(defer.value);
    defer.capacity = std::max(65536, 64 * n);
    CK(cudaMalloc(&defer.count, sizeof(int)));
    CK(cudaMalloc(&defer.index, sizeof(size_t) * defer.capacity));
    CK(cudaMalloc(&defer.value, sizeof(real) * defer.capacity));
    CK(cudaHostGetDevicePointer(&defer.overflow, hflags + 1, 0));
    defer.range_flag = range_flag_dev;
    CK(cudaHostGetDevicePointer(&defer.seen, hflags + 6, 0));
    if (hstage) { cudaFreeHost(hstage); hstage = nullptr; }
    hstage_elems = (size_t)n * 16;
    CK(cudaMallocHost(&hstage, sizeof(double) * hstage_elems));
    return 0;
  }

  int upload(real *dst, const std::vector<real> &src) {
    CK(cudaMemcpyAsync(dst, src.data(), sizeof(real) * src.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }

  /* main():1834-1861 once r, x1, x2 (metres, `real`) are known */
  int finish_setup(const std::vector<real> &r, const std::vector<real> &x1, const std::vector<real> &x2) {
    int rc = alloc_grains((int)r.size());
    if (rc) return rc;
    fin_pending.sums = nullptr;
    const double pi = 3.14159265358979; /* src/main.c:42 */
    const real tau = (real)P.tau, nu = (real)P.nu, rho_moy = (real)P.rho_moy, reductionR = (real)P.reductionR;
    const real G = (real)P.G, angleG = (real)P.angleG, kg = (real)P.kg, iterDEM = (real)P.iterDEM;
    std::vector<real> m(n), It(n), rLB(n);
    for (int i = 0; i < n; ++i) { /* :624-626 */
      m[i] = P.rhoS * pi * r[i] * r[i];
      It[i] = m[i] * r[i] * r[i] / 2;
    }
    Mgx = 0.;
    Mdx = 1.e-3 * lx / 10;
    Mhy = 1.e-3 * ly / 10;
    Mby = 0.;
    xG = -G * sin((double)angleG);
    yG = -G * cos((double)angleG);
    dx = (1. / P.scale) * (Mdx - Mgx) / (lx - 1);
    real rMin = r[0];
    rmax = r[0];
    for (int i = 1; i < n; ++i) { rMin = fmin((double)rMin, (double)r[i]); rmax = std::max(rmax, r[i]); }
    const real dtmax = (1 / iterDEM) * pi * rMin * sqrt(pi * P.rhoS / kg);
    dtLB = dx * dx * (tau - 0.5) / (3 * nu);
    npDEM = (int)(dtLB / dtmax + 1);
    c = dx / dtLB;
    dt = dtLB / npDEM;
    dt2 = dt * dt;
    for (int i = 0; i < n; ++i) rLB[i] = reductionR * r[i] / dx;
    {
      /* tile bins: a tile's region is (RTX+2) x (RTY+2) nodes; discs of radius >= R = rMin/dx whose bounding box
       * touches it have their centre within a rectangle (RTX + 2R + 4) x (RTY + 2R + 4); a dense packing spends
       * 2 sqrt(3) R^2 nodes per disc.  Half as much again for interpenetration, and a floor for large grains. */
      const double R = std::max(0.5, (double)rMin / (double)dx);
      const double est = 1.5 * (RTX + 2 * R + 4) * (RTY + 2 * R + 4) / (3.4641 * R * R) + 8;
      dfree(tbins.count); dfree(tbins.list); dfree(tbins.stamp);
      tbins.cap = (int)std::min(4096.0, std::max(16.0, est));
      tbins.ntx = (nxl + RTX - 1) / RTX;
      tbins.nty = (ly + RTY - 1) / RTY;
      const size_t nt = (size_t)tbins.ntx * tbins.nty;
      CK(cudaMalloc(&tbins.count, sizeof(int) * nt));
      CK(cudaMemsetAsync(tbins.count, 0, sizeof(int) * nt, stream));
      /* one block: stamps, the two dirty lists, their counts, the tile kernel's exit ticket */
      CK(cudaMalloc(&tbins.stamp, sizeof(int) * (3 * nt + 3)));
      CK(cudaMemsetAsync(tbins.stamp, 0, sizeof(int) * (3 * nt + 3), stream));
      tbins.dirty = tbins.stamp + nt;
      tbins.ndirty = tbins.stamp + 3 * nt;
      tbins.ticket = tbins.stamp + 3 * nt + 2;
      tbins.resident_ctas = sm_count * 7;
      raster_step = 1;
      CK(cudaMalloc(&tbins.list, sizeof(TileEntry<real>) * nt * tbins.cap));
      CK(cudaMemsetAsync(tbins.list, 0, sizeof(TileEntry<real>) * nt * tbins.cap, stream)); /* the tile kernel prefetches entries past the count */
      CK(cudaHostGetDevicePointer(&tbins.overflow, hflags + 4, 0));
      /* the two sparse lists, one segment per tile.  A tile of RTX x RTY nodes holds about RTX RTY / (2 sqrt(3) R^2)
       * discs of a dense packing, each with a rim of 2 pi 0.85 R nodes, about 3.5 fluid links per rim node; twice
       * that for interpenetration and polydispersity, and never less than the rim of one straight wall through the tile
       * (next to the lattice ring every link of a rim node is listed: 8 per node). */
      const double rim = (double)RTX * RTY / (3.4641 * R * R) * 6.2832 * 0.85 * R;
      dfree(llist.entry); dfree(llist.tcount); dfree(blist.entry); dfree(blist.tcount);
      llist.cap = (int)std::min((double)(8 * RTX * RTY), std::max(8.0 * (RTX + RTY), 2 * 3.5 * rim));
      blist.cap = (int)std::min((double)(RTX * RTY), std::max(2.0 * (RTX + RTY), 2 * rim));
      llist.ntx = blist.ntx = tbins.ntx;
      llist.nty = blist.nty = tbins.nty;
      CK(cudaMalloc(&llist.entry, sizeof(uint2) * nt * llist.cap));
      CK(cudaMalloc(&llist.tcount, sizeof(int) * nt));
      CK(cudaMemsetAsync(llist.tcount, 0, sizeof(int) * nt, stream));
      CK(cudaMalloc(&blist.entry, sizeof(uint2) * nt * blist.cap));
      CK(cudaMalloc(&blist.tcount, sizeof(int) * nt));
      CK(cudaMemsetAsync(blist.tcount, 0, sizeof(int) * nt, stream));
      CK(cudaHostGetDevicePointer(&blist.overflow, hflags + 2, 0));
      CK(cudaHostGetDevicePointer(&llist.overflow, hflags + 3, 0));
      raster_full_until = raster_step + 2; /* both copies of the map and every list segment are built from scratch */
    }
    /* :1329-1331 with the reference's promotions ((tau - 0.5) is double) */
    k12 = (double)(real)(rho_moy * 9 * nu * nu) / ((double)dx * ((double)tau - 0.5) * ((double)tau - 0.5));
    k3 = (double)(real)(dx * rho_moy * 9 * nu * nu) / ((double)dx * ((double)tau - 0.5) * ((double)tau - 0.5));
    if ((rc = upload(g.r, r)) || (rc = upload(g.x1, x1)) || (rc = upload(g.x2, x2)) || (rc = upload(g.m, m)) ||
        (rc = upload(g.It, It)) || (rc = upload(g.rLB, rLB)))
      return rc;
    /* init_density (:716-724) and init_obst (:663-711) */
    const Lattice<real> L = lattice();
    CK(launch_fill_rest<real>(f[0], plane, L, stream));
    CK(launch_fill_rest<real>(f[1], plane, L, stream));
    for (int k = 0; k < 2; ++k) CK(launch_cell_frame(cell[k], cls[k], own16[k], lx, ly, x0, nxl, pitch, n, stream));
    cur = 0;
    cur_cell = 0;
    holds_A = false;
    scratch_valid = false;
    if ((rc = raster_into(cur_cell))) return rc;
    CK(cudaStreamSynchronize(stream));
    nbsteps = 0;
    nFile = 0;
    t = 0;
    ready = true;
    return n;
  }

  int load_sample(const char *path) override {
    FILE *fp = fopen(path, "r");
    if (!fp) return fail(LBMDEM_EIO, std::string("cannot open ") + path);
    char com[256];
    int cnt = 0;
    if (!fgets(com, sizeof com, fp) || fscanf(fp, "%d\n", &cnt) != 1 || cnt <= 0) {
      fclose(fp);
      return fail(LBMDEM_EIO, std::string("bad header in ") + path);
    }
    std::vector<real> r(cnt), x1(cnt), x2(cnt);
    const real rs = (real)P.rscale;
    for (int i = 0; i < cnt; ++i) {
      double a, b, cc; /* the reference scans with "%le" into double or "%e" into float (FLOAT_FORMAT, :34-40): parse in `real` */
      if (sizeof(real) == 8) {
        if (fscanf(fp, "%le %le %le;\n", &a, &b, &cc) != 3) { fclose(fp); return fail(LBMDEM_EIO, "short sample file"); }
        r[i] = (real)a; x1[i] = (real)b; x2[i] = (real)cc;
      } else {
        float fa, fb, fc;
        if (fscanf(fp, "%e %e %e;\n", &fa, &fb, &fc) != 3) { fclose(fp); return fail(LBMDEM_EIO, "short sample file"); }
        r[i] = (real)fa; x1[i] = (real)fb; x2[i] = (real)fc;
      }
      r[i] = r[i] * rs; /* :623-628 */
      x1[i] = x1[i] * rs;
      x2[i] = x2[i] * rs;
    }
    fclose(fp);
    return finish_setup(r, x1, x2);
  }

  int set_grains(int n_, const double *r_, const double *x1_, const double *x2_) override {
    if (n_ <= 0 || !r_ || !x1_ || !x2_) return fail(LBMDEM_EINVAL, "set_grains: bad arguments");
    std::vector<real> r(n_), x1(n_), x2(n_);
    for (int i = 0; i < n_; ++i) { r[i] = (real)r_[i]; x1[i] = (real)x1_[i]; x2[i] = (real)x2_[i]; }
    return finish_setup(r, x1, x2);
  }

  /* ---- parameter blocks ---- */
  RasterParams<real> raster_params() const {
    RasterParams<real> R;
    R.lx = lx; R.ly = ly; R.dx = dx; R.Mgx = Mgx; R.Mby = Mby;
    return R;
  }
  Lattice<real> lattice() const {
    Lattice<real> L;
    L.lx = lx; L.ly = ly; L.x0 = x0; L.nxl = nxl; L.pitch = pitch; L.plane = plane; L.ngrains = n;
    L.dx = dx; L.c = c; L.Mgx = Mgx; L.Mby = Mby;
    const real lid = (real)P.lid_u;
    L.lid6 = lid / 6;
    L.s2 = (real)P.s2; L.s3 = (real)P.s3; L.s5 = (real)P.s5; L.s7 = (real)P.s7; L.s8 = (real)P.s8; L.s9 = (real)P.s9;
    const real w0[NQ] = {4. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9, 1. / 36, 1. / 9}; /* :53-54 */
    for (int q = 0; q < NQ; ++q) L.w[q] = w0[q];
    return L;
  }
  /* the stored state: populations in f[fbuf] with the map / records of slot `cslot` */
  Stored<real> stored(int fbuf, int cslot) const {
    Stored<real> S;
    S.A = f[fbuf];
    S.cell = cell[cslot];
    S.grains = rec[cslot]; S.boxes = boxes[cslot]; S.R2 = R2[cslot];
    S.act_folded = act_folded[cslot] ? 1 : 0;
    return S;
  }
  bool act_folded[2] = {false, false};
  int raster_into(int cslot) {
    long long *fa = P.strict_fp ? nullptr : facc;
    /* params.kernel bit 1: every tile rebuilt every step (the cross-check of the incremental rasteriser) */
    const int full = (raster_step <= raster_full_until || (P.kernel & 2)) ? 1 : 0;
    CK(launch_raster_tiles<real>(raster_params(), n, g, rec[cslot], R2[cslot], boxes[cslot], rec[1 - cslot], R2[1 - cslot],
                                 boxes[1 - cslot], cell[cslot], cell[1 - cslot], cls[cslot], cls[1 - cslot], own16[cslot],
                                 own16[1 - cslot], x0, nxl, pitch, tbins, blist, llist,
                                 defer.count, fa, raster_step, raster_step == 1 ? 1 : 0, full, stream));
    ++raster_step;
    all_launches += 2; /* grain_bin, raster_tile */
    act_folded[cslot] = true;
    return 0;
  }
  /* the grain state or the map came from outside: the next two rasteriser runs rebuild everything */
  void raster_invalidate() { raster_full_until = raster_step + 2; }
  dem::Params<real> dem_params() const {
    dem::Params<real> D;
    D.kg = (real)P.kg; D.kt = (real)P.kt; D.km = (real)P.km; D.ktm = (real)P.ktm; D.nug = (real)P.nug;
    D.num = (real)P.num; D.numb = (real)P.numb; D.nugt = (real)P.nugt; D.mu = (real)P.mu; D.mum = (real)P.mum;
    D.mumb = (real)P.mumb; D.murf = (real)P.murf; D.freq = (real)P.freq; D.amp = (real)P.amp; D.t = t;
    D.distVerlet = (real)P.distVerlet; D.dt = dt; D.dt2 = dt2; D.xG = xG; D.yG = yG;
    D.Mgx = Mgx; D.Mdx = Mdx; D.Mby = Mby; D.Mhy = Mhy;
    return D;
  }

  int nccl_fail(int r, const char *what) {
    return fail(LBMDEM_ENCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  }

  /* GHOST rows of populations per side, as the fused kernel left them: the ring
   * and bounce-back sweeps that follow reach that far beyond the rows they write */
  /* in-process group: every rank PULLS its ghost rows from the neighbours' device memory */
  int halo_pull_local(cudaStream_t st) {
    LocalGroup::Slot &me = group->slot[P.rank];
    me.f_cur = f[cur];
    CK(cudaEventRecord(me.ev_k1, stream)); /* the fused kernel of this step is behind this */
    if (!group->barrier()) return fail(LBMDEM_ESTATE, "a peer of the in-process strip group failed");
    if (!peers_ready) { /* every rank has reached its first step, so every rank is attached */
      int rc = enable_peer_access();
      if (rc) return rc;
      peers_ready = true;
    }
    real *F = f[cur];
    const size_t width = (size_t)GHOST * pitch * sizeof(real); /* consecutive rows are contiguous */
    if (P.rank > 0) { /* the left neighbour's last owned rows -> the ghost rows below my first row */
      const LocalGroup::Slot &nb = group->slot[P.rank - 1];
      CK(cudaStreamWaitEvent(st, nb.ev_k1, 0));
      const real *src = static_cast<const real *>(nb.f_cur) + (size_t)(nb.xhi - GHOST - nb.x0) * nb.pitch;
      CK(cudaMemcpy2DAsync(F, plane * sizeof(real), src, nb.plane * sizeof(real), width, NQ, cudaMemcpyDefault, st));
    }
    if (P.rank < P.nranks - 1) { /* the right neighbour's first owned rows -> the ghost rows past my last row */
      const LocalGroup::Slot &nb = group->slot[P.rank + 1];
      CK(cudaStreamWaitEvent(st, nb.ev_k1, 0));
      const real *src = static_cast<const real *>(nb.f_cur) + (size_t)(nb.xlo - nb.x0) * nb.pitch;
      CK(cudaMemcpy2DAsync(F + (size_t)(xhi - x0) * pitch, plane * sizeof(real), src, nb.plane * sizeof(real), width, NQ,
                           cudaMemcpyDefault, st));
    }
    CK(cudaEventRecord(me.ev_pulled, st));
    if (!group->barrier()) return fail(LBMDEM_ESTATE, "a peer of the in-process strip group failed");
    /* the passes that follow on `stream` rewrite my first / last owned rows in place: not before the neighbours
     * have pulled them */
    if (P.rank > 0) CK(cudaStreamWaitEvent(stream, group->slot[P.rank - 1].ev_pulled, 0));
    if (P.rank < P.nranks - 1) CK(cudaStreamWaitEvent(stream, group->slot[P.rank + 1].ev_pulled, 0));
    return 0;
  }
  /* in-process group: sums over the ranks of this step's partial force sums, into fsum */
  template <typename T>
  int sum_local(const T *mine, T **total) {
    LocalGroup::Slot &me = group->slot[P.rank];
    me.partial = mine;
    CK(cudaEventRecord(me.ev_partial, stream));
    if (!group->barrier()) return fail(LBMDEM_ESTATE, "a peer of the in-process strip group failed");
    PeerPtrs pp;
    pp.count = P.nranks;
    for (int k = 0; k < P.nranks; ++k) {
      pp.p[k] = group->slot[k].partial;
      if (k != P.rank) CK(cudaStreamWaitEvent(stream, group->slot[k].ev_partial, 0));
    }
    *total = static_cast<T *>(fsum);
    CK(launch_peer_sum<T>(pp, 3 * n, *total, stream));
    ++all_launches;
    return 0;
  }

  /* ---- peer-memory sums across processes (NCCL transport) ---- */
  void ipc_peers(IpcPeers *pp) const {
    pp->nranks = P.nranks; pp->rank = P.rank; pp->lx = lx;
    for (int k = 0; k < P.nranks; ++k) {
      pp->facc[k] = static_cast<const long long *>(k == P.rank ? (const void *)facc_buf[fslot] : ipc_peer_facc[fslot][k]);
      pp->flags[k] = k == P.rank ? ipc_flags : ipc_peer_flags[k];
    }
  }
  void ipc_close() {
    for (int k = 0; k < MAX_LOCAL_RANKS; ++k) {
      for (int j = 0; j < 2; ++j)
        if (ipc_peer_facc[j][k]) { cudaIpcCloseMemHandle(ipc_peer_facc[j][k]); ipc_peer_facc[j][k] = nullptr; }
      if (ipc_peer_flags[k]) { cudaIpcCloseMemHandle(ipc_peer_flags[k]); ipc_peer_flags[k] = nullptr; }
    }
    if (ipc_flags) { cudaFree(ipc_flags); ipc_flags = nullptr; }
    ipc_ready = false;
    cudaGetLastError();
  }
  /* Collective over the ranks (every rank reaches its first LBM step): exchange the IPC handles of the two force-sum
   * buffers and of the arrival words through the communicator, map the peers' buffers, and agree -- all ranks or none --
   * on using them.  Any failure leaves the run on ncclAllReduce. */
  int ipc_open() {
    ipc_tried = true;
    struct Handles { cudaIpcMemHandle_t h[3]; };
    static_assert(sizeof(Handles) % 8 == 0, "handles travel as int64");
    int ok = P.nranks <= MAX_LOCAL_RANKS ? 1 : 0;
    Handles mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaMalloc(&ipc_flags, 256) != cudaSuccess) ok = 0;
    if (ok) CK(cudaMemsetAsync(ipc_flags, 0, 256, stream));
    if (ok && (cudaIpcGetMemHandle(&mine.h[0], facc_buf[0]) != cudaSuccess || cudaIpcGetMemHandle(&mine.h[1], facc_buf[1]) != cudaSuccess ||
               cudaIpcGetMemHandle(&mine.h[2], ipc_flags) != cudaSuccess)) ok = 0;
    cudaGetLastError();
    char *xd = nullptr;
    CK(cudaMalloc(&xd, sizeof(Handles) * P.nranks + 8));
    std::vector<Handles> all(P.nranks);
    CK(cudaMemcpyAsync(xd + sizeof(Handles) * P.rank, &mine, sizeof mine, cudaMemcpyHostToDevice, stream));
    int r = g_nccl.AllGather(xd + sizeof(Handles) * P.rank, xd, sizeof(Handles) / 8, NcclApi::Int64, comm, stream);
    if (r) { cudaFree(xd); return nccl_fail(r, "ncclAllGather(ipc handles)"); }
    CK(cudaMemcpyAsync(all.data(), xd, sizeof(Handles) * P.nranks, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int k = 0; ok && k < P.nranks; ++k) {
      if (k == P.rank) continue;
      void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;
      if (cudaIpcOpenMemHandle(&p0, all[k].h[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&p1, all[k].h[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
          cudaIpcOpenMemHandle(&p2, all[k].h[2], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
      ipc_peer_facc[0][k] = p0; ipc_peer_facc[1][k] = p1; ipc_peer_flags[k] = static_cast<unsigned *>(p2);
    }
    cudaGetLastError();
    /* all or none: the sum of the ok flags must be nranks */
    long long v = ok, *vd = reinterpret_cast<long long *>(xd + sizeof(Handles) * P.nranks);
    CK(cudaMemcpyAsync(vd, &v, sizeof v, cudaMemcpyHostToDevice, stream));
    r = g_nccl.AllReduce(vd, vd, 1, NcclApi::Int64, NcclApi::Sum, comm, stream);
    if (r) { cudaFree(xd); return nccl_fail(r, "ncclAllReduce(ipc agreement)"); }
    CK(cudaMemcpyAsync(&v, vd, sizeof v, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    cudaFree(xd);
    if (v == P.nranks) {
      ipc_ready = true;
      ipc_step = 0;
    } else {
      ipc_close();
    }
    if (getenv("LBMDEM_VERBOSE") && P.rank == 0) fprintf(stderr, "lbmdem: peer-memory force sums %s\n", ipc_ready ? "on" : "unavailable, using ncclAllReduce");
    return 0;
  }

  int halo_exchange(cudaStream_t st) {
    if (P.nranks == 1) return 0;
    if (group) return halo_pull_local(st);
    if (!comm) return fail(LBMDEM_ESTATE, "nranks > 1 but no communicator attached (lbmdem_attach_nccl / lbmdem_attach_local)");
    const int dtype = sizeof(real) == 8 ? NcclApi::Float64 : NcclApi::Float32;
    int r = g_nccl.GroupStart();
    if (r) return nccl_fail(r, "ncclGroupStart");
    real *F = f[cur];
    const size_t cnt = (size_t)GHOST * pitch; /* consecutive rows are contiguous */
    for (int q = 0; q < NQ && !r; ++q) {
      real *pl = F + (size_t)q * plane;
      if (P.rank > 0) { /* left neighbour: send the first owned rows, receive into the ghost rows below them */
        r = g_nccl.Send(pl + (size_t)GHOST * pitch, cnt, dtype, P.rank - 1, comm, st);
        if (!r) r = g_nccl.Recv(pl, cnt, dtype, P.rank - 1, comm, st);
      }
      if (!r && P.rank < P.nranks - 1) {
        const int end = xhi - x0; /* local row just past the owned rows */
        r = g_nccl.Send(pl + (size_t)(end - GHOST) * pitch, cnt, dtype, P.rank + 1, comm, st);
        if (!r) r = g_nccl.Recv(pl + (size_t)end * pitch, cnt, dtype, P.rank + 1, comm, st);
      }
    }
    const int r2 = g_nccl.GroupEnd();
    if (r) return nccl_fail(r, "ncclSend/Recv");
    if (r2) return nccl_fail(r2, "ncclGroupEnd");
    return 0;
  }

  int record_k1_begin(std::pair<cudaEvent_t, cudaEvent_t> **slot) {
    *slot = nullptr;
    if (!events_on) return 0;
    if (ev_used == ev_pool.size()) {
      if (ev_pool.size() >= 8192) { /* fold what we have so far */
        int rc = fold_events();
        if (rc) return rc;
      } else {
        std::pair<cudaEvent_t, cudaEvent_t> p;
        CK(cudaEventCreate(&p.first));
        CK(cudaEventCreate(&p.second));
        ev_pool.push_back(p);
      }
    }
    *slot = &ev_pool[ev_used++];
    CK(cudaEventRecord((*slot)->first, stream));
    return 0;
  }
  int fold_events() {
    CK(cudaStreamSynchronize(stream));
    for (size_t k = 0; k < ev_used; ++k) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, ev_pool[k].first, ev_pool[k].second));
      k1_ms_acc += ms;
    }
    ev_used = 0;
    return 0;
  }

  FusedArgs<real> fused_args(int out_buf, int stream_only) const {
    FusedArgs<real> a;
    a.L = lattice();
    a.A = f[cur];
    a.cell_prev = cell[1 - cur_cell];
    a.cell_new = cell[cur_cell];
    a.grains_new = rec[cur_cell];
    a.out = f[out_buf];
    a.xlo = xlo; a.xhi = xhi;
    a.stream_only = stream_only;
    a.prev16 = n < (int)OWN16_NONE ? 1 : 0;
    a.cls_prev = cls[1 - cur_cell];
    return a;
  }
  /* sweep 5 of the stored array (+ sweeps 1-2 of the new step unless stream_only) into f[1 - cur] */
  int launch_fused(int stream_only, bool timed) {
    const FusedArgs<real> a = fused_args(1 - cur, stream_only);
    if ((P.kernel & 1)) {
      CK(P.strict_fp ? k1_strict::launch_lbm_plain<real>(a, 0, stream) : k1_fast::launch_lbm_plain<real>(a, 0, stream));
      ++all_launches;
      return 0;
    }
    std::pair<cudaEvent_t, cudaEvent_t> *ev = nullptr;
    int rc;
    if (timed && (rc = record_k1_begin(&ev))) return rc;
    const CUtensorMap &tmPrev = a.prev16 ? tmC16[1 - cur_cell] : tmC[1 - cur_cell];
    CK(P.strict_fp ? k1_strict::launch_lbm_rows<real>(tmA[cur], tmPrev, tmCh[cur_cell], a, stream)
                   : k1_fast::launch_lbm_rows<real>(tmA[cur], tmPrev, tmCh[cur_cell], a, stream));
    if (ev) CK(cudaEventRecord(ev->second, stream));
    if (timed) ++k1_launches;
    CK(P.strict_fp ? k1_strict::launch_lbm_plain<real>(a, 1, stream) : k1_fast::launch_lbm_plain<real>(a, 1, stream));
    all_launches += 2;
    return 0;
  }

  /* the LBM part of renderScene (:1711-1717), asynchronous on `stream` */
  int lbm_step_async() {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    int rc;
    if (comm && ipc_wanted && !ipc_tried && P.nranks > 1 && !P.strict_fp && (rc = ipc_open())) return rc;
    if (group || ipc_ready) { /* peers may still be adding up the previous step's partial sums: this step fills the other buffer */
      fslot ^= 1;
      facc = facc_buf[fslot];
      fpartial = fpartial_buf[fslot];
    }
    cur_cell ^= 1; /* the rasteriser writes the other map; the previous one stays with the stored array */
    if ((rc = raster_into(cur_cell))) return rc;
    scratch_valid = false;
    if (!holds_A) {
      /* populations came from outside (init_density, set_f): sweeps 1-2 alone, in place */
      const Lattice<real> L = lattice();
      CK(P.strict_fp ? k1_strict::launch_lbm_h1<real>(L, f[cur], cell[1 - cur_cell], cell[cur_cell], rec[cur_cell], xlo, xhi, stream)
                     : k1_fast::launch_lbm_h1<real>(L, f[cur], cell[1 - cur_cell], cell[cur_cell], rec[cur_cell], xlo, xhi, stream));
      ++all_launches;
      holds_A = true;
      dead_stale = false;
    } else {
      if ((rc = launch_fused(0, true))) return rc;
      cur ^= 1;
      dead_stale = true;
      L_fused = lattice(); /* vibrating walls move the lattice origin between LBM steps: fill_dead needs this step's */
    }
    /* sweeps 3-4 in place (ring :1123-1145, grain bounce-back :1154-1222), then forces_fluid (:1285-1333) */
    const Lattice<real> L = lattice();
    const Stored<real> S = stored(cur, cur_cell);
    const bool multi = P.nranks > 1;
    long long *fa = P.strict_fp ? nullptr : facc;
    /* the deferred list and the force sums were emptied by the rasteriser's first kernel */
    if (!multi) {
      CK(launch_ring_sweep<real>(L, S, f[cur], 0, lx, stream));
      if (P.strict_fp) CK(launch_bounce_pass<real>(L, S, f[cur], 1, lx - 1, xlo, xhi, llist, defer, fa, stream));
      else {
        /* hflags[6]: the number of links the previous sweep deferred; beyond a few hundred a launch of its own applies them */
        const int apply_here = *(volatile int *)(hflags + 6) <= 512 ? 1 : 0;
        CK(launch_rim<real>(L, S, f[cur], 1, lx - 1, xlo, xhi, llist, blist, defer, fa, tbins.ticket, apply_here, stream));
        if (!apply_here) ++all_launches;
      }
    } else if (xhi - xlo < 12) {
      if ((rc = halo_exchange(stream))) return rc;
      CK(launch_ring_sweep<real>(L, S, f[cur], std::max(xlo - 3, 0), std::min(xhi + 3, lx), stream));
      CK(launch_bounce_pass<real>(L, S, f[cur], std::max(xlo - 1, 1), std::min(xhi + 1, lx - 1), xlo, xhi, llist, defer, fa, stream));
    } else {
      /* The ghost rows travel on their own stream while this one sweeps the rows that do not need them.  NCCL reads
       * the first / last GHOST owned rows while it sends them, so the interior passes WRITE nothing there: the ring
       * sweep takes rows [xlo+4, xhi-4); a link of row x reads populations of rows x-2 .. x+2 -- ring nodes among them,
       * which must be swept already -- so the bounce-back pass takes rows [xlo+6, xhi-6), and the ring nodes the edge
       * passes sweep later (rows < xlo+4) read interior nodes of rows <= xlo+4, which that pass has not touched. */
      CK(cudaEventRecord(ev_state, stream));
      CK(cudaStreamWaitEvent(comm_stream, ev_state, 0));
      if ((rc = halo_exchange(comm_stream))) return rc;
      CK(cudaEventRecord(ev_halo, comm_stream));
      CK(launch_ring_sweep<real>(L, S, f[cur], xlo + GHOST, xhi - GHOST, stream));
      CK(launch_bounce_pass<real>(L, S, f[cur], xlo + GHOST + 2, xhi - GHOST - 2, xlo, xhi, llist, defer, fa, stream));
      CK(cudaStreamWaitEvent(stream, ev_halo, 0));
      CK(launch_ring_sweep<real>(L, S, f[cur], std::max(xlo - 3, 0), xlo + GHOST, stream));
      CK(launch_ring_sweep<real>(L, S, f[cur], xhi - GHOST, std::min(xhi + 3, lx), stream));
      CK(launch_bounce_pass<real>(L, S, f[cur], std::max(xlo - 1, 1), xlo + GHOST + 2, xlo, xhi, llist, defer, fa, stream));
      CK(launch_bounce_pass<real>(L, S, f[cur], xhi - GHOST - 2, std::min(xhi + 1, lx - 1), xlo, xhi, llist, defer, fa, stream));
      all_launches += 4;
    }
    all_launches += 2;
    if (P.strict_fp) {
      CK(launch_bounce_end<real>(f[cur], defer, stream));
      ++all_launches;
      CK(launch_force_serial<real>(L, S, xlo, xhi, fpartial, stream));
      ++all_launches;
      double *ftot = fpartial;
      if (multi && group) {
        if ((rc = sum_local<double>(fpartial, &ftot))) return rc;
      } else if (multi) {
        const int r = g_nccl.AllReduce(fpartial, fpartial, (size_t)3 * n, NcclApi::Float64, NcclApi::Sum, comm, stream);
        if (r) return nccl_fail(r, "ncclAllReduce");
      }
      fin_pending = ForceFinish{ftot, 0, k12, k3, nullptr};
    } else {
      long long *ftot = facc;
      if (multi) {
        CK(launch_force_links<real>(L, S, xlo, xhi, blist, facc, f[cur], defer, stream)); /* applies the deferred links first */
        ++all_launches;
        if (group) { /* integer sum: exact, identical on every rank, independent of the decomposition */
          if ((rc = sum_local<long long>(facc, &ftot))) return rc;
        } else if (ipc_ready) {
          IpcPeers pp;
          ipc_peers(&pp);
          ++ipc_step;
          CK(launch_ipc_publish(pp, ipc_step, stream));
          ftot = static_cast<long long *>(fsum);
          CK(launch_ipc_sum(pp, ipc_step, n, boxes[cur_cell], ftot, ipc_timeout_dev, stream));
          all_launches += 2;
        } else {
          const int r = g_nccl.AllReduce(facc, facc, (size_t)3 * n, NcclApi::Int64, NcclApi::Sum, comm, stream);
          if (r) return nccl_fail(r, "ncclAllReduce");
        }
      }
      fin_pending = ForceFinish{ftot, 1, k12, k3, range_flag_dev};
    }
    /* the sums become fhf1..3 in the first DEM launch that follows (or in materialise_fhf) */
    return 0;
  }
  ForceFinish take_pending() {
    const ForceFinish f = fin_pending;
    fin_pending.sums = nullptr;
    return f;
  }
  int materialise_fhf() {
    if (!fin_pending.sums) return 0;
    CK(launch_force_finish<real>(take_pending(), n, g.fhf1, g.fhf2, g.fhf3, stream));
    ++all_launches;
    return 0;
  }

  /* The reference's f[x][y][q] as of now.  While a stream is pending it is materialised into the
   * other population buffer (sweep 5 of the stored array, nothing else); the state is untouched. */
  int observable_f(const real **out) {
    if (!holds_A) { *out = f[cur]; return 0; }
    if (!scratch_valid) {
      if (dead_stale) {
        /* the stream reads every node: first materialise what the fused kernel left unwritten, on the owned rows
         * and the row either side of them */
        const Lattice<real> &L = L_fused;
        const int xa = std::max(xlo - 1, 1), xb = std::min(xhi + 1, lx - 1);
        CK(P.strict_fp ? k1_strict::launch_lbm_fill_dead<real>(L, f[cur], cell[1 - cur_cell], cell[cur_cell], rec[cur_cell], xa, xb, stream)
                       : k1_fast::launch_lbm_fill_dead<real>(L, f[cur], cell[1 - cur_cell], cell[cur_cell], rec[cur_cell], xa, xb, stream));
        ++all_launches;
        dead_stale = false;
      }
      int rc = launch_fused(1, false);
      if (rc) return rc;
      scratch_valid = true;
    }
    *out = f[1 - cur];
    return 0;
  }
  /* make f[cur] hold plain populations again (before they or the map are overwritten from outside) */
  int settle() {
    if (!holds_A) return 0;
    const real *obs;
    int rc = observable_f(&obs);
    if (rc) return rc;
    cur ^= 1;
    holds_A = false;
    scratch_valid = false;
    return 0;
  }

  /* error flags the kernels raise in mapped host memory; valid after a stream synchronise */
  int check_flags() {
    CK(cudaStreamSynchronize(stream));
    if (hflags[0]) { hflags[0] = 0; return fail(LBMDEM_ECAP, "a grain has more Verlet neighbours than neighbour_capacity"); }
    if (hflags[1]) { hflags[1] = 0; return fail(LBMDEM_ECAP, "deferred bounce-back link list is full"); }
    if (hflags[2]) { hflags[2] = 0; return fail(LBMDEM_ECAP, "boundary-node list is full"); }
    if (hflags[3]) { hflags[3] = 0; return fail(LBMDEM_ECAP, "bounce-back link list is full"); }
    if (hflags[4]) { hflags[4] = 0; return fail(LBMDEM_ECAP, "more grains under one lattice tile than the tile bins hold"); }
    if (hflags[7]) { hflags[7] = 0; return fail(LBMDEM_ENCCL, "a peer's force sums did not arrive (peer-memory sum timed out)"); }
    if (hflags[5]) { hflags[5] = 0; return fail(LBMDEM_ERANGE, "a hydrodynamic-force sum left the range of the 64-bit fixed-point accumulators (diverged populations?)"); }
    return 0;
  }

  int verlet_async() {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    /* VerletWall (:1555-1561): the confining walls move out once nbsteps*dt >= dtt */
    if (nbsteps * dt < (real)P.dtt) {
      Mdx = 1.e-3 * lx / 10;
      Mhy = (1.e-3 * ly / 10);
    } else {
      Mdx = 1.e-3 * lx;
      Mhy = 1.e-3 * ly;
    }
    const real cell_size = 2 * rmax + (real)P.distVerlet;
    CK(launch_verlet<real>(dem_params(), n, g, cell_size, vb, stream));
    all_launches += 4;
    return 0;
  }

  int check_verlet_overflow() { return check_flags(); }

  real *mid_dev = nullptr; /* [6][n], lbmdem_step_capture */
  int step_async(long nsteps, bool *built, bool capture = false) {
    bool drift_done = false; /* the previous call's last DEM launch already did this call's kick-drift */
    for (long k = 0; k < nsteps; ++k) {
      int rc;
      if (P.vib == 1) { /* vibrating walls (:1701-1706) */
        const real freq = (real)P.freq, amp = (real)P.amp;
        t = t + dt;
        Mgx = Mgx + amp * sin((double)(freq * t));
        Mdx = Mdx + amp * sin((double)(freq * t));
      }
      if (nbsteps % npDEM == 0 && (rc = lbm_step_async())) return rc;
      if (nbsteps % P.UpdateVerlet == 0) {
        if ((rc = verlet_async())) return rc;
        *built = true;
      }
      const bool film = (nbsteps % P.stepFilm == 0);
      /* this call's sub-step and those of the following calls that do nothing else go in ONE launch: a thread-block
       * cluster for small samples, a cooperative grid for anything that fits one wave; the launch also turns the force
       * sums of the LBM step into fhf */
      long nb = 1;
      const bool coop = n <= DEM_CLUSTER_MAX || n <= coop_cap;
      if (coop && !capture && P.vib != 1 && !(P.kernel & 4)) {
        while (k + nb < nsteps && (nbsteps + nb) % npDEM != 0 && (nbsteps + nb) % P.UpdateVerlet != 0 &&
               (nbsteps + nb) % P.stepFilm != 0)
          ++nb;
        CK(launch_dem_coop<real>(dem_params(), n, (int)nb, film, g, vb, take_pending(), stream));
        all_launches += 1;
        drift_done = false;
      } else {
        if ((rc = materialise_fhf())) return rc;
        /* when the next call of this batch does nothing but its DEM sub-step (no LBM step, no list rebuild), this
         * call's closing kick and the next call's kick-drift go in one launch */
        const bool drift_next = !capture && P.vib != 1 && k + 1 < nsteps &&
                                (nbsteps + 1) % npDEM != 0 && (nbsteps + 1) % P.UpdateVerlet != 0;
        CK(launch_dem_step<real>(dem_params(), n, film, g, vb, capture ? mid_dev : nullptr, drift_done, drift_next, stream));
        all_launches += drift_done ? 2 : 3;
        drift_done = drift_next;
      }
      for (long b = 0; b < nb; ++b) {
        ++nbsteps;
        if (nbsteps % P.stepFilm == 0) ++nFile;
      }
      k += nb - 1;
    }
    return 0;
  }

  /* a rank of an in-process group that fails releases the peers waiting for it at the group's barrier */
  int done(int rc) {
    if (rc && group) group->abort();
    return rc;
  }
  int step(long nsteps) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    bool built = false;
    int rc = step_async(nsteps, &built);
    if (rc) return done(rc);
    (void)built;
    return done(check_flags());
  }
  int step_capture(double *mid) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    if (!mid) return fail(LBMDEM_EINVAL, "step_capture: null output");
    if (!mid_dev) CK(cudaMalloc(&mid_dev, sizeof(real) * 6 * n));
    bool built = false;
    int rc = step_async(1, &built, true);
    if (rc) return done(rc);
    std::vector<real> tmp((size_t)6 * n);
    CK(cudaMemcpyAsync(tmp.data(), mid_dev, sizeof(real) * 6 * n, cudaMemcpyDeviceToHost, stream));
    if ((rc = check_flags())) return rc;
    for (int k = 0; k < 6; ++k)
      for (int i = 0; i < n; ++i) mid[(size_t)i * 6 + k] = tmp[(size_t)k * n + i];
    return 0;
  }
  int lbm_step() override {
    int rc = lbm_step_async();
    if (!rc) rc = materialise_fhf();
    if (rc) return done(rc);
    return done(check_flags());
  }
  int lbm_steps(long k) override {
    for (long i = 0; i < k; ++i) {
      int rc = lbm_step_async();
      if (!rc) rc = materialise_fhf();
      if (rc) return done(rc);
    }
    return done(check_flags());
  }
  int build_verlet() override {
    int rc = verlet_async();
    if (rc) return rc;
    return check_verlet_overflow();
  }

  int get_scalars(double *d, long *l) override {
    d[0] = dx; d[1] = dtLB; d[2] = dt; d[3] = dt2; d[4] = c; d[5] = Mgx; d[6] = Mdx; d[7] = Mby; d[8] = Mhy;
    d[9] = xG; d[10] = yG;
    l[0] = npDEM; l[1] = nbsteps; l[2] = nFile; l[3] = n;
    return 0;
  }
  int set_nbsteps(long v) override { nbsteps = v; return 0; }
  int get_strip(int *a, int *b) override { *a = xlo; *b = xhi; return 0; }

  int total_density(double *sum) override {
    const real *obs;
    int rc = observable_f(&obs);
    if (rc) return rc;
    CK(launch_density<real>(obs, ly, x0, xlo, xhi, pitch, plane, dens_partials, DENS_BLOCKS, dens_out, stream));
    CK(cudaMemcpyAsync(sum, dens_out, sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }

  int ensure_stage(size_t elems) {
    if (elems <= stage_elems) return 0;
    cudaFree(stage);
    stage = nullptr;
    stage_elems = 0;
    CK(cudaMalloc(&stage, sizeof(double) * elems));
    stage_elems = elems;
    return 0;
  }
  int get_f(double *out) override {
    const int chunk = 64;
    int rc = ensure_stage((size_t)chunk * ly * NQ);
    if (rc) return rc;
    const real *obs;
    if ((rc = observable_f(&obs))) return rc;
    for (int r0 = xlo; r0 < xhi; r0 += chunk) {
      const int nr = std::min(chunk, xhi - r0);
      CK(launch_f_to_host_layout<real>(obs, ly, pitch, plane, r0 - x0, nr, stage, stream));
      CK(cudaMemcpyAsync(out + (size_t)(r0 - xlo) * ly * NQ, stage, sizeof(double) * (size_t)nr * ly * NQ,
                         cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
    }
    return 0;
  }
  int set_f(const double *in) override {
    const int chunk = 64;
    int rc = ensure_stage((size_t)chunk * ly * NQ);
    if (rc) return rc;
    holds_A = false; /* every owned population is overwritten: a pending stream is moot */
    scratch_valid = false;
    for (int r0 = xlo; r0 < xhi; r0 += chunk) {
      const int nr = std::min(chunk, xhi - r0);
      CK(cudaMemcpyAsync(stage, in + (size_t)(r0 - xlo) * ly * NQ, sizeof(double) * (size_t)nr * ly * NQ,
                         cudaMemcpyHostToDevice, stream));
      CK(launch_f_from_host_layout<real>(f[cur], ly, pitch, plane, r0 - x0, nr, stage, stream));
      CK(cudaStreamSynchronize(stream));
    }
    return 0;
  }
  int get_obst(int *out) override {
    CK(cudaMemcpy2DAsync(out, sizeof(int) * ly, cell[cur_cell] + (size_t)(xlo - x0) * pitch, sizeof(int) * pitch,
                         sizeof(int) * ly, xhi - xlo, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const size_t cnt = (size_t)(xhi - xlo) * ly;
    for (size_t k = 0; k < cnt; ++k) out[k] = cell_obst(out[k]); /* drop the act bit */
    return 0;
  }
  int set_obst(const int *in) override {
    int rc = settle(); /* a pending stream belongs to the map that is about to be replaced */
    if (rc) return rc;
    act_folded[cur_cell] = false;
    raster_invalidate(); /* this map is not what the grain records would give */
    CK(cudaMemcpy2DAsync(cell[cur_cell] + (size_t)(xlo - x0) * pitch, sizeof(int) * pitch, in, sizeof(int) * ly,
                         sizeof(int) * ly, xhi - xlo, cudaMemcpyHostToDevice, stream));
    CK(launch_cls_from_cell(cell[cur_cell], cls[cur_cell], own16[cur_cell], nxl, pitch, n, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }
  int get_act(int *out) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    int *d = nullptr;
    CK(cudaMalloc(&d, sizeof(int) * (size_t)(xhi - xlo) * ly));
    cudaError_t e = launch_act_map<real>(lattice(), stored(cur, cur_cell), xlo, xhi, d, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, sizeof(int) * (size_t)(xhi - xlo) * ly, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    CK(e);
    return 0;
  }

  /* grains: gather the structure of arrays into rows of doubles through one pinned buffer */
  int download_cols(real *const *cols, int ncols, double *out, int out_stride, int out_col0) {
    std::vector<real> tmp((size_t)n * ncols);
    for (int k = 0; k < ncols; ++k)
      CK(cudaMemcpyAsync(tmp.data() + (size_t)k * n, cols[k], sizeof(real) * n, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int k = 0; k < ncols; ++k)
      for (int i = 0; i < n; ++i) out[(size_t)i * out_stride + out_col0 + k] = tmp[(size_t)k * n + i];
    return 0;
  }
  int upload_cols(real *const *cols, int ncols, const double *in, int in_stride, int in_col0) {
    std::vector<real> tmp((size_t)n * ncols);
    for (int k = 0; k < ncols; ++k)
      for (int i = 0; i < n; ++i) tmp[(size_t)k * n + i] = (real)in[(size_t)i * in_stride + in_col0 + k];
    for (int k = 0; k < ncols; ++k)
      CK(cudaMemcpyAsync(cols[k], tmp.data() + (size_t)k * n, sizeof(real) * n, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }
  int get_grains(double *out) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    real *cols[13] = {g.x1, g.x2, g.x3, g.v1, g.v2, g.v3, g.a1, g.a2, g.a3, g.r, g.m, g.It, g.rLB};
    return download_cols(cols, 13, out, 13, 0);
  }
  int set_grain_state(const double *in) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    real *cols[9] = {g.x1, g.x2, g.x3, g.v1, g.v2, g.v3, g.a1, g.a2, g.a3};
    return upload_cols(cols, 9, in, 9, 0);
  }
  int get_fhf(double *out) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    int rc = materialise_fhf();
    if (rc) return rc;
    real *cols[3] = {g.fhf1, g.fhf2, g.fhf3};
    return download_cols(cols, 3, out, 3, 0);
  }
  int set_fhf(const double *in) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    fin_pending.sums = nullptr; /* whatever the last LBM step left is superseded */
    real *cols[3] = {g.fhf1, g.fhf2, g.fhf3};
    return upload_cols(cols, 3, in, 3, 0);
  }
  int get_verlet(int *count, int *nbr, int capacity, int *wall_flags) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    if (capacity < vb.cap) return fail(LBMDEM_EINVAL, "get_verlet: capacity smaller than the context's");
    std::vector<int> tmp((size_t)n * vb.cap);
    CK(cudaMemcpyAsync(count, vb.nbr_count, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(tmp.data(), vb.nbr, sizeof(int) * (size_t)n * vb.cap, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(wall_flags, vb.wflags, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < capacity; ++k) nbr[(size_t)i * capacity + k] = (k < count[i] && k < vb.cap) ? tmp[(size_t)i * vb.cap + k] : -1;
    return 0;
  }

  /* ---- checkpoint / restart (absent upstream): everything the next renderScene() call reads ----
   * header, grain slab (x v a fhf r m It rLB as stored), Verlet lists, the obstacle map of the last LBM step
   * (the next one treats it as "old") and the reference's f[x][y][q] of the owned rows.  Restarting from the file
   * continues bit for bit (tests/test_gpu_parity.py::test_checkpoint_restart_is_bit_exact). */
  struct CkHeader {
    char magic[8];
    int version, lx, ly, single, n, nranks, rank, cap;
    double scale;
    long nbsteps, nFile;
    double t, Mgx, Mdx, Mby, Mhy;
    unsigned long long physics; /* fingerprint of the physical parameters the run was made with (version 2) */
  };
  /* FNV-1a over the parameters that decide the continuation: a restart with other physics must fail, not drift */
  unsigned long long physics_hash() const {
    const double v[] = {P.tau, P.nu, P.rho_moy, P.reductionR, P.s2, P.s3, P.s5, P.s7, P.s8, P.s9, P.G, P.angleG, P.kg, P.kt,
                        P.km, P.ktm, P.nug, P.num, P.numb, P.nugt, P.mu, P.mum, P.mumb, P.murf, P.rscale, P.distVerlet, P.dtt,
                        P.iterDEM, P.freq, P.amp, P.rhoS, (double)P.UpdateVerlet, (double)P.stepFilm, P.lid_u,
                        (double)P.strict_fp, (double)P.vib};
    unsigned long long h = 1469598103934665603ull;
    const unsigned char *b = reinterpret_cast<const unsigned char *>(v);
    for (size_t k = 0; k < sizeof v; ++k) h = (h ^ b[k]) * 1099511628211ull;
    return h;
  }
  std::string rank_path(const char *path) const { /* one file per rank of a strip-decomposed run */
    return P.nranks > 1 ? std::string(path) + ".rank" + std::to_string(P.rank) : std::string(path);
  }
  int save_state(const char *path_) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    const std::string rp = rank_path(path_);
    const char *path = rp.c_str();
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail(LBMDEM_EIO, std::string("cannot write ") + path);
    CkHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "LBMDEMCK", 8);
    h.version = 2; h.physics = physics_hash(); h.lx = lx; h.ly = ly; h.single = sizeof(real) == 4; h.n = n; h.nranks = P.nranks; h.rank = P.rank;
    h.cap = vb.cap; h.scale = P.scale; h.nbsteps = nbsteps; h.nFile = nFile;
    h.t = t; h.Mgx = Mgx; h.Mdx = Mdx; h.Mby = Mby; h.Mhy = Mhy;
    int rc = materialise_fhf();
    if (rc) { fclose(fp); return rc; }
    const size_t rows = (size_t)(xhi - xlo);
    std::vector<real> slab((size_t)16 * n);
    std::vector<int> cnt(n), nbr((size_t)n * vb.cap), wf(n), ob(rows * ly);
    std::vector<double> fb(rows * ly * NQ);
    cudaError_t e = cudaMemcpyAsync(slab.data(), g.x1, sizeof(real) * 16 * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), vb.nbr_count, sizeof(int) * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(nbr.data(), vb.nbr, sizeof(int) * nbr.size(), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(wf.data(), vb.wflags, sizeof(int) * n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { fclose(fp); CK(e); }
    if ((rc = get_obst(ob.data())) || (rc = get_f(fb.data()))) { fclose(fp); return rc; }
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1 && fwrite(slab.data(), sizeof(real), slab.size(), fp) == slab.size() &&
              fwrite(cnt.data(), sizeof(int), cnt.size(), fp) == cnt.size() &&
              fwrite(nbr.data(), sizeof(int), nbr.size(), fp) == nbr.size() &&
              fwrite(wf.data(), sizeof(int), wf.size(), fp) == wf.size() &&
              fwrite(ob.data(), sizeof(int), ob.size(), fp) == ob.size() &&
              fwrite(fb.data(), sizeof(double), fb.size(), fp) == fb.size();
    ok = (fclose(fp) == 0) && ok;
    return ok ? 0 : fail(LBMDEM_EIO, std::string("short write to ") + path);
  }
  int load_state(const char *path_) override {
    const std::string rp = rank_path(path_);
    const char *path = rp.c_str();
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(LBMDEM_EIO, std::string("cannot open ") + path);
    CkHeader h;
    if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "LBMDEMCK", 8) || h.version != 2) {
      fclose(fp);
      return fail(LBMDEM_EIO, std::string("not a checkpoint (of this version): ") + path);
    }
    if (h.physics != physics_hash()) {
      fclose(fp);
      return fail(LBMDEM_EINVAL, "checkpoint was written with other physical parameters (tau, nu, contact constants, strict_fp ...)");
    }
    {
      /* the header is not trusted: sizes that disagree with the file itself are rejected before anything is allocated */
      const long at = ftell(fp);
      fseek(fp, 0, SEEK_END);
      const long size = ftell(fp);
      fseek(fp, at, SEEK_SET);
      const size_t rows_ = (size_t)(xhi - xlo), rb = sizeof(real);
      const bool sane = h.n > 0 && h.cap > 0 && h.cap <= 4096;
      const size_t want = sane ? sizeof h + rb * 16 * (size_t)h.n + sizeof(int) * ((size_t)h.n * (2 + (size_t)h.cap)) +
                                     sizeof(int) * rows_ * ly + sizeof(double) * rows_ * ly * NQ
                               : 0;
      if (!sane || size < 0 || (size_t)size != want) {
        fclose(fp);
        return fail(LBMDEM_EIO, std::string("corrupt or truncated checkpoint: ") + path);
      }
    }
    if (h.lx != lx || h.ly != ly || h.single != (int)(sizeof(real) == 4) || h.nranks != P.nranks || h.rank != P.rank ||
        h.scale != P.scale || h.n <= 0) {
      fclose(fp);
      return fail(LBMDEM_EINVAL, "checkpoint was written with another lattice / precision / decomposition");
    }
    const size_t rows = (size_t)(xhi - xlo);
    std::vector<real> slab((size_t)16 * h.n);
    std::vector<int> cnt(h.n), nbr((size_t)h.n * h.cap), wf(h.n), ob(rows * ly);
    std::vector<double> fb(rows * ly * NQ);
    const bool ok = fread(slab.data(), sizeof(real), slab.size(), fp) == slab.size() &&
                    fread(cnt.data(), sizeof(int), cnt.size(), fp) == cnt.size() &&
                    fread(nbr.data(), sizeof(int), nbr.size(), fp) == nbr.size() &&
                    fread(wf.data(), sizeof(int), wf.size(), fp) == wf.size() &&
                    fread(ob.data(), sizeof(int), ob.size(), fp) == ob.size() &&
                    fread(fb.data(), sizeof(double), fb.size(), fp) == fb.size();
    fclose(fp);
    if (!ok) return fail(LBMDEM_EIO, std::string("short read from ") + path);
    /* main()'s set-up from the radii (derived constants, m, It, rLB), then the saved state on top of it */
    const size_t N = (size_t)h.n;
    std::vector<real> r(slab.begin() + 12 * N, slab.begin() + 13 * N), x1(slab.begin(), slab.begin() + N),
        x2(slab.begin() + N, slab.begin() + 2 * N);
    P.neighbour_capacity = h.cap;
    int rc = finish_setup(r, x1, x2);
    if (rc < 0) return rc;
    CK(cudaMemcpyAsync(g.x1, slab.data(), sizeof(real) * 16 * N, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(vb.nbr_count, cnt.data(), sizeof(int) * N, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(vb.nbr, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(vb.wflags, wf.data(), sizeof(int) * N, cudaMemcpyHostToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    nbsteps = h.nbsteps; nFile = h.nFile;
    t = (real)h.t; Mgx = (real)h.Mgx; Mdx = (real)h.Mdx; Mby = (real)h.Mby; Mhy = (real)h.Mhy;
    if ((rc = set_obst(ob.data())) || (rc = set_f(fb.data()))) return rc;
    return n;
  }

  int get_fields(const double *gp_in, float *gpress, float *gvel, float *gacc, float *fpress, float *fvel) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    const size_t nn = (size_t)(xhi - xlo) * ly;
    float *d = nullptr;
    real *gp = nullptr;
    CK(cudaMalloc(&d, sizeof(float) * nn * 11));
    cudaError_t e = cudaSuccess;
    if (gp_in) {
      std::vector<real> tmp(n);
      for (int i = 0; i < n; ++i) tmp[i] = (real)gp_in[i];
      e = cudaMalloc(&gp, sizeof(real) * n);
      if (e == cudaSuccess) e = cudaMemcpy(gp, tmp.data(), sizeof(real) * n, cudaMemcpyHostToDevice);
    }
    float *p0 = d, *p1 = d + nn, *p2 = d + 4 * nn, *p3 = d + 7 * nn, *p4 = d + 8 * nn;
    const real *obs = nullptr;
    if (e == cudaSuccess && observable_f(&obs)) e = cudaErrorUnknown;
    if (e == cudaSuccess)
      e = launch_fields<real>(obs, cell[cur_cell], g, gp, n, ly, x0, xlo, xhi, pitch, plane, (real)P.rho_moy, p0, p1, p2,
                              p3, p4, stream);
    if (e == cudaSuccess && gpress) e = cudaMemcpyAsync(gpress, p0, sizeof(float) * nn, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && gvel) e = cudaMemcpyAsync(gvel, p1, sizeof(float) * nn * 3, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && gacc) e = cudaMemcpyAsync(gacc, p2, sizeof(float) * nn * 3, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && fpress) e = cudaMemcpyAsync(fpress, p3, sizeof(float) * nn, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && fvel) e = cudaMemcpyAsync(fvel, p4, sizeof(float) * nn * 3, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    cudaFree(gp);
    CK(e);
    return 0;
  }

  /* end-to-end step with host buffers: pinned staging, async copies on the work stream */
  /* End-to-end step with host buffers.  The rows of doubles travel as they are (one memcpy into /
   * out of pinned memory, one copy over PCIe each way); the device does the transposition and
   * the double <-> real conversion. */
  double *gstage = nullptr; /* device: [n][9] in, then [n][9] + [n][3] out (doubles, or floats in the same space) */
  /* the grains this rank moves across the host boundary in the `share` form of the end-to-end call: a balanced
   * contiguous index range per rank (all of them on one GPU) */
  void share_range(int rank, int *i0, int *i1) const {
    const int base = n / P.nranks, rem = n % P.nranks;
    *i0 = rank * base + std::min(rank, rem);
    *i1 = *i0 + base + (rank < rem ? 1 : 0);
  }
  int get_share(int *i0, int *i1) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    share_range(P.rank, i0, i1);
    return 0;
  }
  int step_host(const void *state_in, long nsteps, void *state_out, void *fhf_out, double *dens, bool rows_f32,
                bool share) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    if (rows_f32 && sizeof(real) != 4) return fail(LBMDEM_EINVAL, "float grain rows need a single-precision context");
    if (share && P.nranks > 1 && !comm) return fail(LBMDEM_ESTATE, "the share form of lbmdem_step_host needs the NCCL communicator");
    const size_t N = (size_t)n, eb = rows_f32 ? sizeof(float) : sizeof(double);
    int i0 = 0, i1 = n;
    if (share) share_range(P.rank, &i0, &i1);
    const size_t M = (size_t)(i1 - i0); /* rows that cross the host boundary */
    if (!gstage) CK(cudaMalloc(&gstage, sizeof(double) * 12 * N));
    char *gs = reinterpret_cast<char *>(gstage), *hs = reinterpret_cast<char *>(hstage); /* hstage: 16 n doubles, pinned */
    const bool pin_in = state_in && host_is_pinned(state_in);
    const bool pin_out = (!state_out || host_is_pinned(state_out)) && (!fhf_out || host_is_pinned(fhf_out));
    /* Page-locked buffers that the device can address (lbmdem_host_alloc) are read and written IN PLACE by the
     * unpack / pack kernels -- the rows cross PCIe inside those kernels, coalesced, and the two copy operations with
     * their staging slab drop out of the step (0.5 MB per step: the copies were latency, not bandwidth). */
    const void *in_dev = (state_in && pin_in && !(share && P.nranks > 1)) ? host_device_ptr(state_in) : nullptr;
    const bool one_block = state_out && fhf_out && static_cast<char *>(fhf_out) == static_cast<char *>(state_out) + eb * 9 * N;
    void *out_dev = (pin_out && !share && one_block) ? host_device_ptr(state_out) : nullptr;
    if (state_in && in_dev) {
      CK(launch_grain_unpack<real>(in_dev, rows_f32, n, 9, g.x1, stream));
    } else if (state_in) {
      const void *src = state_in;
      if (!pin_in) { memcpy(hs, state_in, eb * 9 * M); src = hs; }
      CK(cudaMemcpyAsync(gs + eb * 9 * (size_t)i0, src, eb * 9 * M, cudaMemcpyHostToDevice, stream));
      if (share && P.nranks > 1) {
        /* every rank's rows to every rank, over NVLink instead of PCIe, in place: ONE all-gather when the shares are
         * equal, else one broadcast per rank in one group */
        if (n % P.nranks == 0) {
          const int r = g_nccl.AllGather(gs + eb * 9 * (size_t)i0, gs, (size_t)9 * M, rows_f32 ? NcclApi::Float32 : NcclApi::Float64, comm, stream);
          if (r) return nccl_fail(r, "ncclAllGather");
        } else {
        int r = g_nccl.GroupStart();
        for (int k = 0; k < P.nranks && !r; ++k) {
          int a, b;
          share_range(k, &a, &b);
          char *at = gs + eb * 9 * (size_t)a;
          r = g_nccl.Broadcast(at, at, (size_t)9 * (b - a), rows_f32 ? NcclApi::Float32 : NcclApi::Float64, k, comm, stream);
        }
        const int r2 = g_nccl.GroupEnd();
        if (r) return nccl_fail(r, "ncclBroadcast");
        if (r2) return nccl_fail(r2, "ncclGroupEnd");
        }
      }
      CK(launch_grain_unpack<real>(gs, rows_f32, n, 9, g.x1, stream)); /* x1 .. a3 are contiguous in the slab */
    }
    bool built = false;
    int rc = step_async(nsteps, &built);
    if (rc) return done(rc);
    (void)built;
    if ((rc = materialise_fhf())) return rc;
    if (out_dev) {
      CK(launch_grain_pack2<real>(g.x1, 9, g.fhf1, 3, n, out_dev, rows_f32, stream)); /* [n][9] state, then [n][3] fhf, on the host */
    } else if (state_out || fhf_out) {
      CK(launch_grain_pack2<real>(g.x1, 9, g.fhf1, 3, n, gs, rows_f32, stream)); /* [n][9] state, then [n][3] fhf */
      const char *st = gs + eb * 9 * (size_t)i0, *fh = gs + eb * 9 * N + eb * 3 * (size_t)i0;
      if (pin_out && !share && one_block) {
        CK(cudaMemcpyAsync(state_out, gs, eb * 12 * N, cudaMemcpyDeviceToHost, stream)); /* one block on the host too */
      } else if (pin_out) {
        if (state_out) CK(cudaMemcpyAsync(state_out, st, eb * 9 * M, cudaMemcpyDeviceToHost, stream));
        if (fhf_out) CK(cudaMemcpyAsync(fhf_out, fh, eb * 3 * M, cudaMemcpyDeviceToHost, stream));
      } else {
        CK(cudaMemcpyAsync(hs, st, eb * 9 * M, cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(hs + eb * 9 * M, fh, eb * 3 * M, cudaMemcpyDeviceToHost, stream));
      }
    }
    if (dens) {
      const real *obs;
      if ((rc = observable_f(&obs))) return rc;
      CK(launch_density<real>(obs, ly, x0, xlo, xhi, pitch, plane, dens_partials, DENS_BLOCKS, dens_out, stream));
      CK(cudaMemcpyAsync(dens, dens_out, sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    if ((rc = check_flags())) return rc;
    if (!pin_out) {
      if (state_out) memcpy(state_out, hs, eb * 9 * M);
      if (fhf_out) memcpy(fhf_out, hs + eb * 9 * M, eb * 3 * M);
    }
    return 0;
  }
  /* the device's address of page-locked host memory, nullptr if the device cannot address it */
  static void *host_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
  }
  static bool host_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  }

  int attach_nccl(const void *id) override {
    std::string why;
    if (!g_nccl.load(&why)) return fail(LBMDEM_ENCCL, why);
    NcclApi::UniqueId uid;
    memcpy(&uid, id, sizeof uid);
    CK(cudaSetDevice(P.device));
    const int r = g_nccl.CommInitRank(&comm, P.nranks, uid, P.rank);
    if (r) return nccl_fail(r, "ncclCommInitRank");
    return 0;
  }

  int attach_local(LocalGroup *g) override {
    if (!g || g->P != P.nranks) return fail(LBMDEM_EINVAL, "attach_local: the group was created for another number of ranks");
    if (P.nranks > MAX_LOCAL_RANKS) return fail(LBMDEM_EINVAL, "attach_local: too many ranks");
    if (comm) return fail(LBMDEM_ESTATE, "attach_local: an NCCL communicator is attached already");
    LocalGroup::Slot &me = g->slot[P.rank];
    if (me.attached) return fail(LBMDEM_EINVAL, "attach_local: this rank of the group is taken");
    me.device = P.device;
    me.plane = plane; me.pitch = pitch; me.x0 = x0; me.xlo = xlo; me.xhi = xhi;
    CK(cudaEventCreateWithFlags(&me.ev_k1, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&me.ev_pulled, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&me.ev_partial, cudaEventDisableTiming));
    me.attached = true;
    group = g;
    return 0;
  }
  /* peers on other devices: their partial sums are read by this rank's kernel, their rows copied directly */
  int enable_peer_access() {
    if (!group) return 0;
    for (int k = 0; k < P.nranks; ++k) {
      const int d = group->slot[k].device;
      if (!group->slot[k].attached) return fail(LBMDEM_ESTATE, "in-process strip group: not every rank is attached");
      if (d == P.device) continue;
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, P.device, d));
      if (!can) return fail(LBMDEM_ECUDA, "in-process strip group: no peer access between the devices of two ranks");
      const cudaError_t e = cudaDeviceEnablePeerAccess(d, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else CK(e);
    }
    return 0;
  }

  int get_kernel_timer(double *ms, long *k1, long *all) override {
    int rc = fold_events();
    if (rc) return rc;
    if (ms) *ms = k1_ms_acc;
    if (k1) *k1 = k1_launches;
    if (all) *all = all_launches;
    return 0;
  }
  int reset_kernel_timer(int enable) override {
    int rc = fold_events();
    if (rc) return rc;
    k1_ms_acc = 0;
    k1_launches = 0;
    all_launches = 0;
    events_on = enable != 0;
    return 0;
  }
  int state_checksum(unsigned long long *c) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    const real *obs;
    int rc = observable_f(&obs);
    if (rc) return rc;
    unsigned long long *d = reinterpret_cast<unsigned long long *>(dens_partials); /* scratch of the density sum */
    CK(launch_checksum<real>(obs, cell[cur_cell], ly, x0, xlo, xhi, pitch, plane, d, DENS_BLOCKS, stream));
    CK(cudaMemcpyAsync(c, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
  }
  int get_list_counts(long *c) override {
    if (!ready) return fail(LBMDEM_ESTATE, "no grains loaded");
    const size_t nt = (size_t)tbins.ntx * tbins.nty;
    std::vector<int> kc(nt), bc(nt);
    int nd = 0;
    CK(cudaMemcpyAsync(kc.data(), llist.tcount, sizeof(int) * nt, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(bc.data(), blist.tcount, sizeof(int) * nt, cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&nd, defer.count, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    c[0] = c[1] = 0;
    for (size_t t = 0; t < nt; ++t) { c[0] += kc[t]; c[1] += bc[t]; }
    c[2] = nd;
    /* tiles the rasteriser rebuilt at its last run (stamped at that run or the one before) */
    std::vector<int> st(nt);
    CK(cudaMemcpy(st.data(), tbins.stamp, sizeof(int) * nt, cudaMemcpyDeviceToHost));
    c[3] = 0;
    for (size_t t = 0; t < nt; ++t) c[3] += st[t] == raster_step - 1;
    if ((P.kernel & 2) || raster_step - 1 <= raster_full_until) c[3] = (long)nt; /* that run rebuilt every tile */
    return 0;
  }
  void *stream_ptr() override { return (void *)stream; }
};

}  // namespace lbmdem

/* ================================ C ABI ================================ */
using lbmdem::SimBase;
struct lbmdem_ctx {
  SimBase *sim;
};

extern "C" {

#define API __attribute__((visibility("default")))

API int lbmdem_default_params(lbmdem_params *p) {
  if (!p) return LBMDEM_EINVAL;
  memset(p, 0, sizeof *p);
  p->lx = 7826; p->ly = 2325; p->scale = 1.; /* src/main.c:24-32 */
  p->single_precision = 0; p->device = 0; p->rank = 0; p->nranks = 1;
  p->tau = 0.504; p->nu = 1e-6; p->rho_moy = 1000; p->reductionR = 0.85; /* :74-94 */
  p->s2 = 1.5; p->s3 = 1.4; p->s5 = 1.5; p->s7 = 1.5; p->s8 = 1.9841; p->s9 = 1.9841;
  p->G = 9.81; p->angleG = 0.0; /* :97-118 */
  p->kg = 1.6e+6; p->kt = 1.0e+6; p->km = 3e+6; p->ktm = 2e+6;
  p->nug = 6.4e+1; p->num = 8.7e+1; p->numb = 8.7e+1; p->nugt = 5e-1;
  p->mu = .5317; p->mum = .466; p->mumb = .466; p->murf = 0.01;
  p->rscale = 1e-3; p->distVerlet = 5e-4; p->dtt = 0.; p->iterDEM = 100.;
  p->freq = 5; p->amp = 4.e-4; p->rhoS = 2650;
  p->UpdateVerlet = 100; p->stepFilm = 8000;
  p->lid_u = 0; p->strict_fp = 0; p->kernel = 0; p->neighbour_capacity = 32; p->vib = 0;
  return 0;
}

API int lbmdem_sizeof_params(void) { return (int)sizeof(lbmdem_params); }

API int lbmdem_create(const lbmdem_params *p, lbmdem_ctx **out) {
  if (!p || !out) return LBMDEM_EINVAL;
  SimBase *s = p->single_precision ? (SimBase *)new (std::nothrow) lbmdem::Sim<float>()
                                   : (SimBase *)new (std::nothrow) lbmdem::Sim<double>();
  if (!s) return LBMDEM_ENOMEM;
  s->P = *p;
  const int rc = s->init_device(); /* leaves the context's device current for the calling thread */
  if (rc) {
    lbmdem::g_create_error = s->err;
    delete s;
    return rc;
  }
  *out = new lbmdem_ctx{s};
  return 0;
}
API void lbmdem_destroy(lbmdem_ctx *ctx) {
  if (!ctx) return;
  if (ctx->sim) {
    lbmdem::DeviceGuard guard(ctx->sim->P.device);
    delete ctx->sim;
  }
  delete ctx;
}
API const char *lbmdem_last_error(const lbmdem_ctx *ctx) {
  return ctx ? ctx->sim->err.c_str() : lbmdem::g_create_error.c_str();
}

#define CTX_OR_FAIL if (!ctx || !ctx->sim) return LBMDEM_EINVAL
/* every entry point: the context's device made current for the call (any host thread may drive a context), and no
 * C++ exception crosses the C boundary */
#define GUARDED(expr)                                                                                    \
  do {                                                                                                   \
    CTX_OR_FAIL;                                                                                         \
    lbmdem::DeviceGuard guard__(ctx->sim->P.device);                                                     \
    try {                                                                                                \
      return (expr);                                                                                     \
    } catch (const std::bad_alloc &) {                                                                   \
      return ctx->sim->fail(LBMDEM_ENOMEM, "out of host memory");                                        \
    } catch (const std::exception &e) {                                                                  \
      return ctx->sim->fail(LBMDEM_EINVAL, std::string("unexpected: ") + e.what());                      \
    }                                                                                                    \
  } while (0)
API int lbmdem_load_sample(lbmdem_ctx *ctx, const char *path) { GUARDED(ctx->sim->load_sample(path)); }
API int lbmdem_set_grains(lbmdem_ctx *ctx, int n, const double *r, const double *x1, const double *x2) {
  GUARDED(ctx->sim->set_grains(n, r, x1, x2));
}
API int lbmdem_step(lbmdem_ctx *ctx, long n) { GUARDED(ctx->sim->step(n)); }
API int lbmdem_step_capture(lbmdem_ctx *ctx, double *mid) { GUARDED(ctx->sim->step_capture(mid)); }
API int lbmdem_lbm_step(lbmdem_ctx *ctx) { GUARDED(ctx->sim->lbm_step()); }
API int lbmdem_lbm_steps(lbmdem_ctx *ctx, long n) { GUARDED(ctx->sim->lbm_steps(n)); }
API int lbmdem_build_verlet(lbmdem_ctx *ctx) { GUARDED(ctx->sim->build_verlet()); }
API int lbmdem_get_scalars(lbmdem_ctx *ctx, double *d, long *l) { GUARDED(ctx->sim->get_scalars(d, l)); }
API int lbmdem_set_nbsteps(lbmdem_ctx *ctx, long n) { GUARDED(ctx->sim->set_nbsteps(n)); }
API int lbmdem_get_strip(lbmdem_ctx *ctx, int *a, int *b) { GUARDED(ctx->sim->get_strip(a, b)); }
API int lbmdem_total_density(lbmdem_ctx *ctx, double *s) { GUARDED(ctx->sim->total_density(s)); }
API int lbmdem_get_f(lbmdem_ctx *ctx, double *out) { GUARDED(ctx->sim->get_f(out)); }
API int lbmdem_set_f(lbmdem_ctx *ctx, const double *in) { GUARDED(ctx->sim->set_f(in)); }
API int lbmdem_get_obst(lbmdem_ctx *ctx, int *out) { GUARDED(ctx->sim->get_obst(out)); }
API int lbmdem_set_obst(lbmdem_ctx *ctx, const int *in) { GUARDED(ctx->sim->set_obst(in)); }
API int lbmdem_get_act(lbmdem_ctx *ctx, int *out) { GUARDED(ctx->sim->get_act(out)); }
API int lbmdem_get_grains(lbmdem_ctx *ctx, double *out) { GUARDED(ctx->sim->get_grains(out)); }
API int lbmdem_set_grain_state(lbmdem_ctx *ctx, const double *in) { GUARDED(ctx->sim->set_grain_state(in)); }
API int lbmdem_get_fhf(lbmdem_ctx *ctx, double *out) { GUARDED(ctx->sim->get_fhf(out)); }
API int lbmdem_set_fhf(lbmdem_ctx *ctx, const double *in) { GUARDED(ctx->sim->set_fhf(in)); }
API int lbmdem_get_verlet(lbmdem_ctx *ctx, int *count, int *nbr, int capacity, int *wf) {
  GUARDED(ctx->sim->get_verlet(count, nbr, capacity, wf));
}
API int lbmdem_save_state(lbmdem_ctx *ctx, const char *path) { GUARDED(ctx->sim->save_state(path)); }
API int lbmdem_load_state(lbmdem_ctx *ctx, const char *path) { GUARDED(ctx->sim->load_state(path)); }
API int lbmdem_get_fields(lbmdem_ctx *ctx, const double *gp, float *a, float *b, float *c, float *d, float *e) {
  GUARDED(ctx->sim->get_fields(gp, a, b, c, d, e));
}
API int lbmdem_step_host(lbmdem_ctx *ctx, const double *in, long n, double *out, double *fhf, double *dens) {
  GUARDED(ctx->sim->step_host(in, n, out, fhf, dens, false, false));
}
API int lbmdem_step_host_f32(lbmdem_ctx *ctx, const float *in, long n, float *out, float *fhf, double *dens) {
  GUARDED(ctx->sim->step_host(in, n, out, fhf, dens, true, false));
}
API int lbmdem_get_share(lbmdem_ctx *ctx, int *i0, int *i1) {
  if (!i0 || !i1) return LBMDEM_EINVAL;
  GUARDED(ctx->sim->get_share(i0, i1));
}
API int lbmdem_step_host_share(lbmdem_ctx *ctx, const double *in, long n, double *out, double *fhf, double *dens) {
  GUARDED(ctx->sim->step_host(in, n, out, fhf, dens, false, true));
}
API int lbmdem_step_host_share_f32(lbmdem_ctx *ctx, const float *in, long n, float *out, float *fhf, double *dens) {
  GUARDED(ctx->sim->step_host(in, n, out, fhf, dens, true, true));
}
API int lbmdem_nccl_unique_id(void *id128) {
  std::string why;
  if (!id128) return LBMDEM_EINVAL;
  if (!lbmdem::g_nccl.load(&why)) { lbmdem::g_create_error = why; return LBMDEM_ENCCL; }
  lbmdem::NcclApi::UniqueId uid;
  const int r = lbmdem::g_nccl.GetUniqueId(&uid);
  if (r) { lbmdem::g_create_error = lbmdem::g_nccl.GetErrorString(r); return LBMDEM_ENCCL; }
  memcpy(id128, &uid, sizeof uid);
  return 0;
}
API int lbmdem_attach_nccl(lbmdem_ctx *ctx, const void *id128) { GUARDED(ctx->sim->attach_nccl(id128)); }
API int lbmdem_local_group_create(int nranks, lbmdem_local_group **out) {
  if (!out || nranks < 1 || nranks > lbmdem::MAX_LOCAL_RANKS) return LBMDEM_EINVAL;
  *out = reinterpret_cast<lbmdem_local_group *>(new (std::nothrow) lbmdem::LocalGroup(nranks));
  return *out ? 0 : LBMDEM_ENOMEM;
}
API void lbmdem_local_group_destroy(lbmdem_local_group *group) {
  lbmdem::LocalGroup *g = reinterpret_cast<lbmdem::LocalGroup *>(group);
  if (!g) return;
  for (auto &sl : g->slot) {
    if (!sl.attached) continue;
    lbmdem::DeviceGuard guard(sl.device);
    cudaEventDestroy(sl.ev_k1); cudaEventDestroy(sl.ev_pulled); cudaEventDestroy(sl.ev_partial);
  }
  delete g;
}
API int lbmdem_attach_local(lbmdem_ctx *ctx, lbmdem_local_group *group) {
  GUARDED(ctx->sim->attach_local(reinterpret_cast<lbmdem::LocalGroup *>(group)));
}
API int lbmdem_get_kernel_timer(lbmdem_ctx *ctx, double *ms, long *k1, long *all) {
  GUARDED(ctx->sim->get_kernel_timer(ms, k1, all));
}
API int lbmdem_reset_kernel_timer(lbmdem_ctx *ctx, int enable) { GUARDED(ctx->sim->reset_kernel_timer(enable)); }
API int lbmdem_state_checksum(lbmdem_ctx *ctx, unsigned long long sums[2]) {
  if (!sums) return LBMDEM_EINVAL;
  GUARDED(ctx->sim->state_checksum(sums));
}
API int lbmdem_get_list_counts(lbmdem_ctx *ctx, long counts[4]) {
  if (!counts) return LBMDEM_EINVAL;
  GUARDED(ctx->sim->get_list_counts(counts));
}
API int lbmdem_host_alloc(size_t bytes, void **ptr) {
  if (!ptr || !bytes) return LBMDEM_EINVAL;
  *ptr = nullptr;
  /* mapped and portable: lbmdem_step_host's pack / unpack kernels read and write these buffers in place */
  return cudaHostAlloc(ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess ? 0 : LBMDEM_ECUDA;
}
API int lbmdem_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? 0 : LBMDEM_ECUDA; }
API void *lbmdem_stream(lbmdem_ctx *ctx) { return (ctx && ctx->sim) ? ctx->sim->stream_ptr() : nullptr; }

}  // extern "C"
