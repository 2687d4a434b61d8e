// C ABI + runtime of the B200-native Lagrange-leapfrog step: device mirror of
// the reference Domain (lulesh.h:148-595) in HBM, the per-cycle launch sequence
// (one CUDA graph per cycle; at several ranks a compute and a communication stream
// with peer-to-peer stores over NVLink or, as fallback, NCCL send/recv + allreduce)
// and the host-side construction of the device layouts.
//
// There is no CPU fallback anywhere in this file: without a usable sm_100
// device every compute entry point returns LULESH_B200_ECUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>   // types only; every NCCL symbol is resolved with dlsym at run time

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lulesh_b200.h"
#include "kernels.cuh"
#include "setup.cuh"
#include "host/domain.h"

using namespace lb200;

// --------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...)
{
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   g_last_error = buf;
   return code;
}

#define CK(call)                                                                      \
   do {                                                                               \
      cudaError_t e_ = (call);                                                        \
      if (e_ != cudaSuccess)                                                          \
         return fail(LULESH_B200_ECUDA, "%s failed: %s (%s:%d)", #call,               \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                     \
   } while (0)

// --------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU use has no NCCL dependency.  In a
// python process that already imported torch the dlopen returns torch's copy.
// --------------------------------------------------------------------------
struct NcclApi {
   void *lib = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                             cudaStream_t) = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
   ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
};

static bool load_nccl(NcclApi &api)
{
   const char *names[] = {"libnccl.so.2", "libnccl.so"};
   for (const char *n : names)
      if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
   if (!api.lib) return false;
#define SYM(field, name)                                                      \
   *(void **)(&api.field) = dlsym(api.lib, name);                             \
   if (!api.field) { api.lib = nullptr; return false; }
   SYM(GetUniqueId, "ncclGetUniqueId")
   SYM(CommInitRank, "ncclCommInitRank")
   SYM(CommDestroy, "ncclCommDestroy")
   SYM(GroupStart, "ncclGroupStart")
   SYM(GroupEnd, "ncclGroupEnd")
   SYM(Send, "ncclSend")
   SYM(Recv, "ncclRecv")
   SYM(AllReduce, "ncclAllReduce")
   SYM(GetErrorString, "ncclGetErrorString")
   SYM(AllGather, "ncclAllGather")
#undef SYM
   return true;
}

static NcclApi *nccl_api()
{
   // one load per process; the driver creates its handles from one thread per GPU, and the
   // initialisation of a function-local static is thread-safe
   static NcclApi api;
   static const bool ok = load_nccl(api);
   return ok ? &api : nullptr;
}

#define NK(call)                                                                      \
   do {                                                                               \
      ncclResult_t r_ = (call);                                                       \
      if (r_ != ncclSuccess)                                                          \
         return fail(LULESH_B200_ENCCL, "%s failed: %s", #call,                       \
                     h->nccl->GetErrorString(r_));                                    \
   } while (0)

// --------------------------------------------------------------------------
// handle
// --------------------------------------------------------------------------
struct Message {      // one neighbour of the node-halo exchange
   int rank, count;   // nodes shared with that neighbour
   size_t send_off, recv_off;   // doubles, into sendbuf / fhalo (3 fields each)
};

struct FaceMessage {  // one face neighbour of the MonoQ exchange
   int rank, count;
   size_t send_off;   // into mq_send (3 fields)
   size_t ghost_off;  // element offset of the ghost block inside delv_*
};

struct lulesh_b200 {
   int device = 0;
   int numRanks = 1, rank = 0;
   cudaStream_t stream = nullptr, comm_stream = nullptr;
   cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
   cudaEvent_t ev_fork = nullptr, ev_dt = nullptr;
   KParams P{};
   Ctl *h_ctl = nullptr;   // pinned mirror
   std::vector<void *> allocs;
   size_t device_bytes = 0, upload_bytes = 0;
   double *field_ptr[LULESH_F_COUNT] = {};
   size_t field_cnt[LULESH_F_COUNT] = {};
   int debug = 0;
   // graph of one cycle (single rank)
   cudaGraphExec_t graph = nullptr;
   int graph_debug = -1;
   // comm
   NcclApi *nccl = nullptr;
   ncclComm_t comm = nullptr;
   std::vector<Message> msgs;
   std::vector<FaceMessage> faces;
   double *sendbuf = nullptr;
   int *pack_idx = nullptr;
   size_t send_total = 0;
   double *mq_send = nullptr;
   int *mq_idx = nullptr;
   size_t mq_total = 0;
   int64_t launches = 0;
   // peer-to-peer exchange state (numRanks > 1 and every rank can map every other rank's arena)
   bool p2p = false;
   char *arena = nullptr;            // one allocation: flags | dt slots | fhalo[2] | delv[3][allElem]
   size_t arena_bytes = 0;
   unsigned long long *flags = nullptr;
   DtSlot *dtslots = nullptr;
   PeerCounters *counters = nullptr;
   PeerMsg *d_node_msgs = nullptr, *d_face_msgs = nullptr;
   unsigned char *d_node_slot_msg = nullptr, *d_face_slot_msg = nullptr;
   DtSlot **d_peer_slots = nullptr;
   std::vector<void *> ipc_opened;
   std::string halo_mode = "none";
   int launches_per_cycle = 5;
   int k1_grid = 0, k3_grid = 0;   // persistent grids: SMs x resident blocks (capped by the work)
};

template <typename T>
static int dev_alloc(lulesh_b200 *h, T **out, size_t count)
{
   void *p = nullptr;
   const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
   CK(cudaMalloc(&p, bytes));
   h->allocs.push_back(p);
   h->device_bytes += bytes;
   *out = static_cast<T *>(p);
   return 0;
}

template <typename T>
static int dev_upload(lulesh_b200 *h, T **out, const T *src, size_t count)
{
   int rc = dev_alloc(h, out, count);
   if (rc) return rc;
   if (count) {
      CK(cudaMemcpy(*out, src, count * sizeof(T), cudaMemcpyHostToDevice));
      h->upload_bytes += count * sizeof(T);
   }
   return 0;
}

template <typename T>
static int dev_zero(lulesh_b200 *h, T **out, size_t count)
{
   int rc = dev_alloc(h, out, count);
   if (rc) return rc;
   CK(cudaMemset(*out, 0, std::max<size_t>(count, 1) * sizeof(T)));
   return 0;
}

static inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

static int region_rep(int r, int numReg, int cost)   // lulesh.cc:2393-2400
{
   if (r < numReg / 2) return 1;
   if (r < numReg - (numReg + 15) / 20) return 1 + cost;
   return 10 * (1 + cost);
}

// the 26 neighbour directions (dcol, drow, dplane): 6 faces in the ghost-block
// order of lulesh-init.cc:582-610, then 12 edges, then 8 corners
static const int k_dirs[26][3] = {
   {0, 0, -1}, {0, 0, 1}, {0, -1, 0}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0},
   {-1, -1, 0}, {0, -1, -1}, {-1, 0, -1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1},
   {-1, 1, 0}, {0, -1, 1}, {-1, 0, 1}, {1, -1, 0}, {0, 1, -1}, {1, 0, -1},
   {-1, -1, -1}, {-1, -1, 1}, {1, -1, -1}, {1, -1, 1},
   {-1, 1, -1}, {-1, 1, 1}, {1, 1, -1}, {1, 1, 1}};

static int neighbour_rank(const lulesh_b200_host_view *v, const int dir[3])
{
   const int c = v->colLoc + dir[0], r = v->rowLoc + dir[1], p = v->planeLoc + dir[2];
   if (c < 0 || c >= v->px || r < 0 || r >= v->py || p < 0 || p >= v->pz) return -1;
   return p * v->px * v->py + r * v->px + c;
}

static void shared_nodes(const lulesh_b200_host_view *v, const int dir[3], std::vector<int> &out)
{
   const int nx1 = v->sizeX + 1, ny1 = v->sizeY + 1, nz1 = v->sizeZ + 1;
   const int i0 = dir[0] > 0 ? v->sizeX : 0, i1 = dir[0] == 0 ? nx1 : i0 + 1;
   const int j0 = dir[1] > 0 ? v->sizeY : 0, j1 = dir[1] == 0 ? ny1 : j0 + 1;
   const int k0 = dir[2] > 0 ? v->sizeZ : 0, k1 = dir[2] == 0 ? nz1 : k0 + 1;
   out.clear();
   for (int k = k0; k < k1; ++k)
      for (int j = j0; j < j1; ++j)
         for (int i = i0; i < i1; ++i) out.push_back(k * nx1 * ny1 + j * nx1 + i);
}

// boundary element layer facing `dir` (a face direction), in the (slow,fast)
// order the ghost indices of lulesh-init.cc:613-671 expect
static void face_elems(const lulesh_b200_host_view *v, const int dir[3], std::vector<int> &out)
{
   const int sx = v->sizeX, sy = v->sizeY, sz = v->sizeZ;
   out.clear();
   if (dir[2] != 0) {
      const int k = dir[2] < 0 ? 0 : sz - 1;
      for (int j = 0; j < sy; ++j) for (int i = 0; i < sx; ++i) out.push_back(k * sx * sy + j * sx + i);
   } else if (dir[1] != 0) {
      const int j = dir[1] < 0 ? 0 : sy - 1;
      for (int k = 0; k < sz; ++k) for (int i = 0; i < sx; ++i) out.push_back(k * sx * sy + j * sx + i);
   } else {
      const int i = dir[0] < 0 ? 0 : sx - 1;
      for (int k = 0; k < sz; ++k) for (int j = 0; j < sy; ++j) out.push_back(k * sx * sy + j * sx + i);
   }
}

// --------------------------------------------------------------------------
// create / destroy
// --------------------------------------------------------------------------
extern "C" const char *lulesh_b200_last_error(void) { return g_last_error.c_str(); }

extern "C" int lulesh_b200_get_unique_id(void *out_id)
{
   NcclApi *api = nccl_api();
   if (!api) return fail(LULESH_B200_ENCCL, "libnccl.so.2 could not be loaded");
   static_assert(sizeof(ncclUniqueId) == LULESH_B200_UNIQUE_ID_BYTES, "unique id size");
   ncclUniqueId id;
   ncclResult_t r = api->GetUniqueId(&id);
   if (r != ncclSuccess) return fail(LULESH_B200_ENCCL, "ncclGetUniqueId: %s", api->GetErrorString(r));
   memcpy(out_id, &id, sizeof id);
   return 0;
}

static int build_comm(lulesh_b200 *h, const lulesh_b200_host_view *v, const void *unique_id,
                      std::vector<unsigned char> &nodeFlags);

// Everything create() can check without a GPU.  The kernels index with the connectivity,
// neighbour, corner and region arrays unchecked, so anything out of range is rejected here.
static int validate_view(const lulesh_b200_host_view *v, bool generated)
{
   if (v->abi_version != LULESH_B200_ABI_VERSION)
      return fail(LULESH_B200_EINVAL, "abi_version %d != %d", v->abi_version, LULESH_B200_ABI_VERSION);
   const int ne = v->numElem, nn = v->numNode;
   if (ne <= 0 || nn <= 0 || (long long)v->sizeX * v->sizeY * v->sizeZ != ne ||
       (long long)(v->sizeX + 1) * (v->sizeY + 1) * (v->sizeZ + 1) != nn)
      return fail(LULESH_B200_EINVAL, "inconsistent sizes in host view");
   if ((long long)ne + 15 > INT_MAX / 8 || (long long)nn + 31 > INT_MAX / 8)
      return fail(LULESH_B200_EINVAL, "brick too large for int32 indices");
   if (v->numRanks < 1 || v->px < 1 || v->py < 1 || v->pz < 1 || v->px * v->py * v->pz != v->numRanks ||
       v->rank < 0 || v->rank >= v->numRanks || v->colLoc < 0 || v->colLoc >= v->px || v->rowLoc < 0 ||
       v->rowLoc >= v->py || v->planeLoc < 0 || v->planeLoc >= v->pz ||
       v->rank != v->planeLoc * v->px * v->py + v->rowLoc * v->px + v->colLoc)   // lulesh-init.cc:732-734
      return fail(LULESH_B200_EINVAL, "inconsistent decomposition in host view");
   if (!v->regElemSize || !v->regElemlist || v->numReg < 1 ||
       (!generated && (!v->x || !v->y || !v->z || !v->xd || !v->yd || !v->zd || !v->nodalMass || !v->nodelist ||
                       !v->lxim || !v->lxip || !v->letam || !v->letap || !v->lzetam || !v->lzetap || !v->elemBC ||
                       !v->e || !v->p || !v->q || !v->v || !v->volo || !v->ss || !v->elemMass ||
                       !v->nodeElemStart || !v->nodeElemCornerList)))
      return fail(LULESH_B200_EINVAL, "null array in host view");

   // region lists: a partition of the elements (lulesh-init.cc:401-510)
   std::vector<char> seen(ne, 0);
   long long total = 0;
   for (int r = 0; r < v->numReg; ++r) {
      const int n = v->regElemSize[r];
      if (n < 0 || (n > 0 && !v->regElemlist[r])) return fail(LULESH_B200_EINVAL, "bad region list %d", r);
      total += n;
      for (int t = 0; t < n; ++t) {
         const int el = v->regElemlist[r][t];
         if (el < 0 || el >= ne) return fail(LULESH_B200_EINVAL, "region list entry out of range");
         if (seen[el]++) return fail(LULESH_B200_EINVAL, "element %d is in more than one region list", el);
      }
   }
   if (total != ne) return fail(LULESH_B200_EINVAL, "region lists cover %lld of %d elements", total, ne);
   if (generated) return 0;

   const int allElem = ne + 2 * v->sizeX * v->sizeY + 2 * v->sizeX * v->sizeZ + 2 * v->sizeY * v->sizeZ;
   for (size_t i = 0; i < (size_t)8 * ne; ++i)
      if (v->nodelist[i] < 0 || v->nodelist[i] >= nn) return fail(LULESH_B200_EINVAL, "nodelist entry out of range");
   const int32_t *nbr[6] = {v->lxim, v->lxip, v->letam, v->letap, v->lzetam, v->lzetap};
   for (const int32_t *a : nbr)
      for (int i = 0; i < ne; ++i)
         if (a[i] < 0 || a[i] >= allElem) return fail(LULESH_B200_EINVAL, "face neighbour out of range");
   const struct { const int32_t *list; int n; } symm[3] = {
      {v->symmX, v->numSymmX}, {v->symmY, v->numSymmY}, {v->symmZ, v->numSymmZ}};
   for (const auto &sset : symm) {
      if (sset.n < 0 || (sset.n > 0 && !sset.list)) return fail(LULESH_B200_EINVAL, "bad symmetry node set");
      for (int i = 0; i < sset.n; ++i)
         if (sset.list[i] < 0 || sset.list[i] >= nn) return fail(LULESH_B200_EINVAL, "symmetry node out of range");
   }
   for (int n = 0; n < nn; ++n) {   // node -> corner CSR (lulesh-init.cc:295-319)
      const int b = v->nodeElemStart[n], e = v->nodeElemStart[n + 1];
      if (b < 0 || e < b || e - b > 8) return fail(LULESH_B200_EINVAL, "node %d has %d corners", n, e - b);
      for (int k = b; k < e; ++k)
         if (v->nodeElemCornerList[k] < 0 || v->nodeElemCornerList[k] >= 8 * ne)
            return fail(LULESH_B200_EINVAL, "corner list entry out of range");
   }
   return 0;
}

// `gen` != NULL: device-side Sedov setup (lulesh_b200_create_sedov): the view carries sizes,
// decomposition, constants, scalars and the region lists only; every other array is generated
// in HBM by the kernels of setup.cu.
static int create_impl(lulesh_b200 *h, const lulesh_b200_host_view *v, int device,
                       const void *unique_id, const SetupParams *gen)
{
   int rc;
   if ((rc = validate_view(v, gen != nullptr))) return rc;
   const int ne = v->numElem, nn = v->numNode;

   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return fail(LULESH_B200_ECUDA, "no CUDA device available (this library has no CPU fallback)");
   if (device < 0 || device >= ndev) return fail(LULESH_B200_EINVAL, "device %d out of range", device);
   cudaDeviceProp prop;
   CK(cudaGetDeviceProperties(&prop, device));
   if (prop.major != 10)
      return fail(LULESH_B200_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                  device, prop.major, prop.minor);
   CK(cudaSetDevice(device));
   h->device = device;
   h->numRanks = v->numRanks;
   h->rank = v->rank;
   CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
   {  // the exchange kernels are tiny and latency-critical: their blocks go first when an SM frees up
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
   }
   CK(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
   CK(cudaEventCreateWithFlags(&h->ev_dt, cudaEventDisableTiming));
   CK(cudaEventCreate(&h->ev_t0));
   CK(cudaEventCreate(&h->ev_t1));
   CK(cudaMallocHost(&h->h_ctl, sizeof(Ctl)));

   KParams &P = h->P;
   P.ne = ne; P.nn = nn;
   P.allElem = ne + 2 * v->sizeX * v->sizeY + 2 * v->sizeX * v->sizeZ + 2 * v->sizeY * v->sizeZ;
   P.ne_pad = (ne + 15) & ~15;
   P.nn_pad = (nn + 31) & ~31;
   P.c = v->constants;
   P.unit_rho0 = (v->constants.refdens == 1.0) ? 1 : 0;

   // ---- control block
   Ctl c0{};
   c0.dtcourant_bits = 0; c0.dthydro_bits = 0;
   memcpy(&c0.dtcourant_bits, &v->scalars.dtcourant, 8);
   memcpy(&c0.dthydro_bits, &v->scalars.dthydro, 8);
   c0.dtfixed = v->scalars.dtfixed; c0.time = v->scalars.time; c0.deltatime = v->scalars.deltatime;
   c0.deltatimemultlb = v->scalars.deltatimemultlb; c0.deltatimemultub = v->scalars.deltatimemultub;
   c0.dtmax = v->scalars.dtmax; c0.stoptime = v->scalars.stoptime;
   c0.gnewdt = 1.0e+20; c0.cycle = v->scalars.cycle; c0.max_cycles = INT_MAX; c0.done = 0; c0.error = 0;
   if ((rc = dev_upload(h, &P.ctl, &c0, 1))) return rc;

   // ---- node-centred state
#define UP_NODE(name, id)                                                               \
   { double *p_; if ((rc = gen ? dev_zero(h, &p_, (size_t)nn) : dev_upload(h, &p_, v->name, (size_t)nn))) return rc; \
     h->field_ptr[id] = p_; h->field_cnt[id] = nn; }
   UP_NODE(x, LULESH_F_X) UP_NODE(y, LULESH_F_Y) UP_NODE(z, LULESH_F_Z)
   UP_NODE(xd, LULESH_F_XD) UP_NODE(yd, LULESH_F_YD) UP_NODE(zd, LULESH_F_ZD)
   UP_NODE(nodalMass, LULESH_F_NODALMASS)
#undef UP_NODE
   P.x = h->field_ptr[LULESH_F_X]; P.y = h->field_ptr[LULESH_F_Y]; P.z = h->field_ptr[LULESH_F_Z];
   P.xd = h->field_ptr[LULESH_F_XD]; P.yd = h->field_ptr[LULESH_F_YD]; P.zd = h->field_ptr[LULESH_F_ZD];
   P.nodalMass = h->field_ptr[LULESH_F_NODALMASS];
   if ((rc = dev_zero(h, &P.dbg_f, (size_t)3 * nn))) return rc;
   if ((rc = dev_zero(h, &P.dbg_a, (size_t)3 * nn))) return rc;
   for (int a = 0; a < 3; ++a) {
      h->field_ptr[LULESH_F_FX + a] = P.dbg_f + (size_t)a * nn; h->field_cnt[LULESH_F_FX + a] = nn;
      h->field_ptr[LULESH_F_XDD + a] = P.dbg_a + (size_t)a * nn; h->field_cnt[LULESH_F_XDD + a] = nn;
   }

   // ---- element-centred state
#define UP_ELEM(name, id, dst)                                                          \
   { double *p_; if ((rc = gen ? dev_zero(h, &p_, (size_t)ne) : dev_upload(h, &p_, v->name, (size_t)ne))) return rc; \
     h->field_ptr[id] = p_; h->field_cnt[id] = ne; dst = p_; }
   UP_ELEM(e, LULESH_F_E, P.e) UP_ELEM(p, LULESH_F_P, P.p) UP_ELEM(q, LULESH_F_Q, P.q)
   UP_ELEM(v, LULESH_F_V, P.v) UP_ELEM(ss, LULESH_F_SS, P.ss)
   { double *p_; if ((rc = gen ? dev_zero(h, &p_, (size_t)ne) : dev_upload(h, &p_, v->volo, (size_t)ne))) return rc;
     h->field_ptr[LULESH_F_VOLO] = p_; h->field_cnt[LULESH_F_VOLO] = ne; P.volo = p_; }
   { double *p_; if ((rc = gen ? dev_zero(h, &p_, (size_t)ne) : dev_upload(h, &p_, v->elemMass, (size_t)ne))) return rc;
     h->field_ptr[LULESH_F_ELEMMASS] = p_; h->field_cnt[LULESH_F_ELEMMASS] = ne; P.elemMass = p_; }
#undef UP_ELEM
#define ZERO_ELEM(id, dst, cnt)                                                         \
   { double *p_; if ((rc = dev_zero(h, &p_, (size_t)(cnt)))) return rc;                 \
     h->field_ptr[id] = p_; h->field_cnt[id] = (cnt); dst = p_; }
   ZERO_ELEM(LULESH_F_QL, P.ql, ne) ZERO_ELEM(LULESH_F_QQ, P.qq, ne)
   ZERO_ELEM(LULESH_F_VNEW, P.vnew, ne) ZERO_ELEM(LULESH_F_DELV, P.delv, ne)
   ZERO_ELEM(LULESH_F_VDOV, P.vdov, ne) ZERO_ELEM(LULESH_F_AREALG, P.arealg, ne)
   ZERO_ELEM(LULESH_F_DELX_XI, P.delx_xi, ne) ZERO_ELEM(LULESH_F_DELX_ETA, P.delx_eta, ne)
   ZERO_ELEM(LULESH_F_DELX_ZETA, P.delx_zeta, ne)
#undef ZERO_ELEM
   if (v->numRanks == 1) {
      // delv_xi/eta/zeta share one allocation [3][allElem] (ghost blocks at the tail of each);
      // at several ranks they live in the peer-visible arena allocated by build_comm()
      double *g;
      if ((rc = dev_zero(h, &g, (size_t)3 * P.allElem))) return rc;
      P.delv_xi = g; P.delv_eta = g + P.allElem; P.delv_zeta = g + 2 * (size_t)P.allElem;
   }
   for (int a = 0; a < 3; ++a) h->field_cnt[LULESH_F_DELV_XI + a] = P.allElem;
   {
      int *p_;
#define UP_INT(name, cnt) if ((rc = gen ? dev_zero(h, &p_, (size_t)(cnt)) : dev_upload(h, &p_, v->name, (size_t)(cnt)))) return rc; P.name = p_;
      UP_INT(nodelist, 8 * (size_t)ne)
      UP_INT(lxim, ne) UP_INT(lxip, ne) UP_INT(letam, ne) UP_INT(letap, ne)
      UP_INT(lzetam, ne) UP_INT(lzetap, ne) UP_INT(elemBC, ne)
#undef UP_INT
   }
   if ((rc = dev_zero(h, &P.fcorner, (size_t)24 * P.ne_pad))) return rc;

   std::vector<unsigned char> nodeFlags(nn, 0);
   unsigned char *d_nodeFlags = nullptr;
   if (gen) {
      // ---- device-side setup: mesh, corner table, connectivity, BCs, flags, volumes, masses
      int *d_ell = nullptr;
      if ((rc = dev_alloc(h, &d_ell, (size_t)8 * P.nn_pad))) return rc;
      if ((rc = dev_alloc(h, &d_nodeFlags, (size_t)nn))) return rc;
      k_setup_nodes<<<blocks_for(nn, 256), 256, 0, h->stream>>>(*gen, P.x, P.y, P.z, d_nodeFlags, d_ell, nn,
                                                                P.nn_pad, P.ne_pad);
      k_setup_elems<<<blocks_for(ne, 256), 256, 0, h->stream>>>(
         *gen, P.x, P.y, P.z, const_cast<int *>(P.nodelist), const_cast<int *>(P.lxim),
         const_cast<int *>(P.lxip), const_cast<int *>(P.letam), const_cast<int *>(P.letap),
         const_cast<int *>(P.lzetam), const_cast<int *>(P.lzetap), const_cast<int *>(P.elemBC),
         const_cast<double *>(P.volo), const_cast<double *>(P.elemMass), P.v, P.e, ne);
      k_setup_nodal_mass<<<blocks_for(nn, 256), 256, 0, h->stream>>>(
         d_ell, P.volo, const_cast<double *>(P.nodalMass), nn, P.nn_pad, P.ne_pad);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(h->stream));
      P.cornerEll = d_ell;
   } else {
   // ---- corner gather table: CSR (lulesh-init.cc:295-319) -> 8-slot ELL, slot
   // order == CSR order == ascending element; entries re-based to the SoA planes
   {
      std::vector<int> ell((size_t)8 * P.nn_pad, -1);
      for (int n = 0; n < nn; ++n) {
         const int b = v->nodeElemStart[n], e = v->nodeElemStart[n + 1];
         if (e - b > 8 || e < b) return fail(LULESH_B200_EINVAL, "node %d has %d corners", n, e - b);
         for (int k = b; k < e; ++k) {
            const int ci = v->nodeElemCornerList[k];
            if (ci < 0 || ci >= 8 * ne) return fail(LULESH_B200_EINVAL, "corner list entry out of range");
            ell[(size_t)(k - b) * P.nn_pad + n] = (ci & 7) * P.ne_pad + (ci >> 3);
         }
      }
      int *p_;
      if ((rc = dev_upload(h, &p_, ell.data(), ell.size()))) return rc;
      P.cornerEll = p_;
   }

   // ---- per-node flags from the symmetry node sets (lulesh-init.cc:514-533)
   for (int i = 0; i < v->numSymmX; ++i) nodeFlags[v->symmX[i]] |= NODE_SYMM_X;
   for (int i = 0; i < v->numSymmY; ++i) nodeFlags[v->symmY[i]] |= NODE_SYMM_Y;
   for (int i = 0; i < v->numSymmZ; ++i) nodeFlags[v->symmZ[i]] |= NODE_SYMM_Z;
   }

   // ---- region work list.  Every cost class is padded to whole blocks so that the EOS
   // repetition count is block-uniform (no divergence in the rep loop).  Block order:
   // the blocks of the most expensive repetition class are FP64-bound, the cheap ones are
   // bound by their three dependent gathers; interleaving them lets each SM overlap the
   // two, and the expensive blocks are all issued within the first ~3/4 of the grid so
   // that they never form the tail.
   // Several ranks: the elements that read ghost values of delv_* (a *_COMM face,
   // lulesh.cc:1782-1783) form a second list at the tail, launched separately once the
   // MonoQ exchange has landed; everything else starts right behind K3 and hides the exchange.
   {
      struct Block { int first, rep; };
      std::vector<int> entries;
      std::vector<Block> heavy, light, face;
      long long total = 0;
      int max_rep = 0, min_rep = INT_MAX;
      for (int r = 0; r < v->numReg; ++r) {
         const int rep = region_rep(r, v->numReg, v->cost);
         if (v->regElemSize[r] > 0) { max_rep = std::max(max_rep, rep); min_rep = std::min(min_rep, rep); }
      }
      const int sx = v->sizeX, sy = v->sizeY, sz = v->sizeZ;
      auto reads_ghosts = [&](int el) -> bool {
         if (v->numRanks == 1) return false;
         if (!gen) return (v->elemBC[el] & (XI_M_COMM | XI_P_COMM | ETA_M_COMM | ETA_P_COMM | ZETA_M_COMM | ZETA_P_COMM)) != 0;
         const int i = el % sx, j = (el / sx) % sy, k = el / (sx * sy);   // the generated brick (setup.cu)
         return (i == 0 && v->colLoc > 0) || (i == sx - 1 && v->colLoc < v->px - 1) ||
                (j == 0 && v->rowLoc > 0) || (j == sy - 1 && v->rowLoc < v->py - 1) ||
                (k == 0 && v->planeLoc > 0) || (k == sz - 1 && v->planeLoc < v->pz - 1);
      };
      // Regions with the same repetition count are one cost class: every element is the same
      // material, only `rep` differs (lulesh.cc:2393-2400).  The lists of a class are merged
      // and sorted by element id, which lengthens the contiguous runs a warp sees (the
      // per-region lists are fragmented into runs of ~38 elements by CreateRegionIndexSets).
      std::vector<int> classes;
      for (int r = 0; r < v->numReg; ++r) {
         const int rep = region_rep(r, v->numReg, v->cost);
         if (std::find(classes.begin(), classes.end(), rep) == classes.end()) classes.push_back(rep);
      }
      std::sort(classes.begin(), classes.end(), std::greater<int>());
      for (int rep : classes) {
         std::vector<int> merged[2];   // [0] interior, [1] elements reading ghost slots
         for (int r = 0; r < v->numReg; ++r) {
            if (region_rep(r, v->numReg, v->cost) != rep) continue;
            const int n = v->regElemSize[r];
            total += n;
            for (int t = 0; t < n; ++t) {
               const int el = v->regElemlist[r][t];
               if (el < 0 || el >= ne) return fail(LULESH_B200_EINVAL, "region list entry out of range");
               merged[reads_ghosts(el) ? 1 : 0].push_back(el);
            }
         }
         for (int part = 0; part < 2; ++part) {
            std::sort(merged[part].begin(), merged[part].end());
            const int n = (int)merged[part].size();
            for (int t0 = 0; t0 < n; t0 += MAT_THREADS) {
               Block b{(int)entries.size(), rep};
               for (int t = t0; t < t0 + MAT_THREADS; ++t) entries.push_back(t < n ? merged[part][t] : -1);
               (part == 1 ? face : (rep == max_rep && max_rep > min_rep) ? heavy : light).push_back(b);
            }
         }
      }
      if (total != ne) return fail(LULESH_B200_EINVAL, "region lists cover %lld of %d elements", total, ne);
      std::vector<Block> sched;
      size_t li = 0;
      for (size_t hi = 0; hi < heavy.size(); ++hi) {
         sched.push_back(heavy[hi]);
         const size_t upto = (size_t)((double)(hi + 1) * 0.75 * (double)light.size() / (double)heavy.size());
         while (li < upto && li < light.size()) sched.push_back(light[li++]);
      }
      while (li < light.size()) sched.push_back(light[li++]);
      P.numWorkBlocksInterior = (int)sched.size();
      sched.insert(sched.end(), face.begin(), face.end());
      std::vector<int> work, reps;
      work.reserve(entries.size());
      for (const Block &b : sched) {
         work.insert(work.end(), entries.begin() + b.first, entries.begin() + b.first + MAT_THREADS);
         reps.push_back(b.rep);
      }
      int *p_;
      if ((rc = dev_upload(h, &p_, work.data(), work.size()))) return rc;
      P.workElem = p_;
      if ((rc = dev_upload(h, &p_, reps.data(), reps.size()))) return rc;
      P.workBlockRep = p_;
      P.numWorkBlocks = (int)reps.size();
   }

   if (v->numRanks > 1) {
      if ((rc = build_comm(h, v, unique_id, nodeFlags))) return rc;
   }
   for (int a = 0; a < 3; ++a) h->field_ptr[LULESH_F_DELV_XI + a] = P.delv_xi + (size_t)a * P.allElem;
   if (gen) {
      P.nodeFlags = d_nodeFlags;   // generated on the device, COMM bit included
   } else {
      unsigned char *p_;
      if ((rc = dev_upload(h, &p_, nodeFlags.data(), nodeFlags.size()))) return rc;
      P.nodeFlags = p_;
   }
   {  // persistent grids for the cp.async-pipelined element kernels: SMs x resident blocks.
      // The shared-memory carveout is the smallest one that still reaches the residency the
      // register file allows, so that as much as possible of the SM's 256 KB stays L1 for
      // the node gathers.
      CK(cudaFuncSetAttribute(k_force, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES));
      CK(cudaFuncSetAttribute(k_kinematics, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM_BYTES));
      auto residency = [&](auto kernel, int threads, int smem, int *occ_out) -> int {
         CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
         int occ = 0;
         CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
         const double need = (double)occ * (smem + 1024);   // 1 KB per block is reserved by the system
         int pct = (int)(100.0 * need / (double)prop.sharedMemPerMultiprocessor) + 1;
         pct = std::max(25, std::min(100, pct));
         CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
         *occ_out = occ;
         return 0;
      };
      int occ1 = 0, occ3 = 0;
      if ((rc = residency(k_force, K1_THREADS, K1_SMEM_BYTES, &occ1))) return rc;
      if ((rc = residency(k_kinematics, K3_THREADS, K3_SMEM_BYTES, &occ3))) return rc;
      if (occ1 < 1 || occ3 < 1) return fail(LULESH_B200_ECUDA, "element kernels do not fit on an SM");
      h->k1_grid = std::min(blocks_for(ne, K1_THREADS), prop.multiProcessorCount * occ1);
      h->k3_grid = std::min(blocks_for(ne, K3_THREADS), prop.multiProcessorCount * occ3);
      if (getenv("LULESH_B200_VERBOSE"))
         fprintf(stderr, "lulesh_b200: K1 %d x %d threads/SM, K3 %d x %d threads/SM\n", occ1, K1_THREADS, occ3,
                 K3_THREADS);
   }
   CK(cudaDeviceSynchronize());
   return 0;
}

// --------------------------------------------------------------------------
// Halo plan: everything the multi-rank path needs to know about who shares what,
// computed on the host from the view alone (no GPU, no NCCL) so that it can be
// unit-tested on CPU (tests/test_multirank_cpu.py drives it over gloo).
//
// Node exchange (replaces CommSend/CommSBN, lulesh-comm.cc:357-1257): for every
// existing neighbour one message of 3 planes x count values; receivers sum ALL
// ranks' partials of a shared node in ascending-rank order, so the result is
// bit-identical on every sharing rank and the reference's CommSyncPosVel
// (lulesh-comm.cc:1261-1680) is unnecessary.
// MonoQ exchange (CommMonoQ, lulesh-comm.cc:1684-1835): face neighbours only,
// received straight into the ghost slots of delv_xi/eta/zeta.
// --------------------------------------------------------------------------
struct lulesh_b200_halo_plan {
   std::vector<int> bnode, bsum_start, bsum_src, pack_idx;
   std::vector<int> msg_rank, msg_count, msg_send_off, msg_recv_off;
   std::vector<int> mq_idx, face_rank, face_count, face_send_off, face_ghost_off;
   int fhalo_size = 0, send_total = 0, mq_total = 0;
};

static int build_halo_plan(const lulesh_b200_host_view *v, lulesh_b200_halo_plan &pl)
{
   const int nn = v->numNode;
   const int allElem = v->numElem + 2 * v->sizeX * v->sizeY + 2 * v->sizeX * v->sizeZ +
                       2 * v->sizeY * v->sizeZ;
   std::vector<int> bmap(nn, -1);
   struct Dir { int rank; std::vector<int> nodes; };
   std::vector<Dir> dirs;
   for (int q = 0; q < 26; ++q) {
      const int nb = neighbour_rank(v, k_dirs[q]);
      if (nb < 0) continue;
      Dir d{nb, {}};
      shared_nodes(v, k_dirs[q], d.nodes);
      for (int n : d.nodes) bmap[n] = 0;
      dirs.push_back(std::move(d));
   }
   for (int n = 0; n < nn; ++n)
      if (bmap[n] == 0) { bmap[n] = (int)pl.bnode.size(); pl.bnode.push_back(n); }
   const int nb = (int)pl.bnode.size();

   long long send_off = 0, recv_off = (long long)3 * nb;
   struct Contribution { int rank, base, stride; };
   std::vector<std::vector<Contribution>> contrib(nb);
   for (int b = 0; b < nb; ++b) contrib[b].push_back({v->rank, b, nb});
   for (const Dir &d : dirs) {
      const int cnt = (int)d.nodes.size();
      if (recv_off + 3LL * cnt > INT_MAX) return fail(LULESH_B200_EINVAL, "halo too large");
      pl.msg_rank.push_back(d.rank); pl.msg_count.push_back(cnt);
      pl.msg_send_off.push_back((int)send_off); pl.msg_recv_off.push_back((int)recv_off);
      for (int a = 0; a < 3; ++a)
         for (int t = 0; t < cnt; ++t) pl.pack_idx.push_back(a * nb + bmap[d.nodes[t]]);
      for (int t = 0; t < cnt; ++t) contrib[bmap[d.nodes[t]]].push_back({d.rank, (int)(recv_off + t), cnt});
      send_off += 3LL * cnt;
      recv_off += 3LL * cnt;
   }
   pl.send_total = (int)send_off;
   pl.fhalo_size = (int)recv_off;
   pl.bsum_start.assign(nb + 1, 0);
   for (int b = 0; b < nb; ++b) {
      std::stable_sort(contrib[b].begin(), contrib[b].end(),
                       [](const Contribution &x, const Contribution &y) { return x.rank < y.rank; });
      for (const Contribution &c : contrib[b]) { pl.bsum_src.push_back(c.base); pl.bsum_src.push_back(c.stride); }
      pl.bsum_start[b + 1] = (int)(pl.bsum_src.size() / 2);
   }

   long long mq_off = 0, ghost = v->numElem;
   std::vector<int> elems;
   for (int q = 0; q < 6; ++q) {   // ghost blocks in pMin,pMax,rMin,rMax,cMin,cMax order
      const int nbr = neighbour_rank(v, k_dirs[q]);
      if (nbr < 0) continue;
      face_elems(v, k_dirs[q], elems);
      const int cnt = (int)elems.size();
      pl.face_rank.push_back(nbr); pl.face_count.push_back(cnt);
      pl.face_send_off.push_back((int)mq_off); pl.face_ghost_off.push_back((int)ghost);
      for (int a = 0; a < 3; ++a)
         for (int e : elems) pl.mq_idx.push_back(a * allElem + e);
      mq_off += 3LL * cnt;
      ghost += cnt;
   }
   pl.mq_total = (int)mq_off;
   return 0;
}

static bool view_is_sane(const lulesh_b200_host_view *v)
{
   return v && v->abi_version == LULESH_B200_ABI_VERSION && v->numElem > 0 && v->numNode > 0 &&
          (long long)v->sizeX * v->sizeY * v->sizeZ == v->numElem &&
          (long long)(v->sizeX + 1) * (v->sizeY + 1) * (v->sizeZ + 1) == v->numNode &&
          v->numRanks >= 1 && v->px * v->py * v->pz == v->numRanks && v->rank >= 0 &&
          v->rank < v->numRanks &&
          v->rank == v->planeLoc * v->px * v->py + v->rowLoc * v->px + v->colLoc;
}

extern "C" int lulesh_b200_halo_plan_create(const lulesh_b200_host_view *view,
                                            lulesh_b200_halo_plan **out)
{
   if (!out || !view_is_sane(view)) return fail(LULESH_B200_EINVAL, "bad view for halo plan");
   lulesh_b200_halo_plan *pl = new lulesh_b200_halo_plan();
   const int rc = build_halo_plan(view, *pl);
   if (rc) { delete pl; return rc; }
   *out = pl;
   return 0;
}

extern "C" int lulesh_b200_halo_plan_query(const lulesh_b200_halo_plan *pl, const char *what,
                                           const int32_t **data, size_t *count)
{
   if (!pl || !what || !data || !count) return fail(LULESH_B200_EINVAL, "null argument");
   struct Entry { const char *name; const std::vector<int> *v; };
   const Entry table[] = {
      {"bnode", &pl->bnode}, {"bsum_start", &pl->bsum_start}, {"bsum_src", &pl->bsum_src},
      {"pack_idx", &pl->pack_idx}, {"msg_rank", &pl->msg_rank}, {"msg_count", &pl->msg_count},
      {"msg_send_off", &pl->msg_send_off}, {"msg_recv_off", &pl->msg_recv_off},
      {"mq_idx", &pl->mq_idx}, {"face_rank", &pl->face_rank}, {"face_count", &pl->face_count},
      {"face_send_off", &pl->face_send_off}, {"face_ghost_off", &pl->face_ghost_off}};
   for (const Entry &e : table)
      if (!strcmp(e.name, what)) { *data = e.v->data(); *count = e.v->size(); return 0; }
   return fail(LULESH_B200_EINVAL, "unknown halo plan array '%s'", what);
}

extern "C" void lulesh_b200_halo_plan_destroy(lulesh_b200_halo_plan *pl) { delete pl; }

// --------------------------------------------------------------------------
// Peer-to-peer setup: every rank publishes where its receive slots live (one table per
// rank, all-gathered over NCCL), maps the other ranks' arenas (raw pointers + peer access
// inside one process, CUDA IPC handles across processes) and builds the device tables the
// exchange kernels use.  All ranks then agree (min-allreduce of a flag) on whether the
// peer path is usable; otherwise everybody stays on NCCL send/recv.
// --------------------------------------------------------------------------
struct PeerTable {
   int pid, device, rank, nmsg, nface, allElem, fhalo_size, ok;
   unsigned long long base;          // arena address in the owner's address space
   unsigned long long alloc_off;     // arena offset inside the cudaMalloc block the IPC handle names
   cudaIpcMemHandle_t handle;
   int msg_rank[26], msg_recv_off[26];
   int face_rank[6], face_ghost_off[6];
   unsigned long long off_fhalo, off_delv;
};

static int setup_p2p(lulesh_b200 *h, const lulesh_b200_host_view *v, const lulesh_b200_halo_plan &pl)
{
   KParams &P = h->P;
   const int n = v->numRanks, me = v->rank;
   const char *mode = getenv("LULESH_B200_HALO");
   const bool want = !(mode && !strcmp(mode, "nccl")) && n <= PEER_MAX_RANKS && h->nccl->AllGather;
   int rc;

   PeerTable mine;
   memset(&mine, 0, sizeof mine);
   mine.pid = (int)getpid(); mine.device = h->device; mine.rank = me;
   mine.nmsg = (int)pl.msg_rank.size(); mine.nface = (int)pl.face_rank.size();
   mine.allElem = P.allElem; mine.fhalo_size = pl.fhalo_size; mine.ok = want ? 1 : 0;
   mine.base = (unsigned long long)(uintptr_t)h->arena;
   mine.off_fhalo = (unsigned long long)((char *)P.fhalo - h->arena);
   mine.off_delv = (unsigned long long)((char *)P.delv_xi - h->arena);
   for (int i = 0; i < mine.nmsg; ++i) { mine.msg_rank[i] = pl.msg_rank[i]; mine.msg_recv_off[i] = pl.msg_recv_off[i]; }
   for (int i = 0; i < mine.nface; ++i) { mine.face_rank[i] = pl.face_rank[i]; mine.face_ghost_off[i] = pl.face_ghost_off[i]; }
   if (want) {
      if (cudaIpcGetMemHandle(&mine.handle, h->arena) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
      // the handle names the whole cudaMalloc block; find the arena's offset inside it
      typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      unsigned long long blockBase = mine.base;
      if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) == cudaSuccess && fn) {
         size_t sz = 0;
         if (((GetRange)fn)(&blockBase, &sz, mine.base) != 0) blockBase = mine.base;
      } else cudaGetLastError();
      mine.alloc_off = mine.base - blockBase;
   }

   // all-gather the tables
   std::vector<PeerTable> all(n);
   PeerTable *d_tab = nullptr;
   if ((rc = dev_zero(h, &d_tab, (size_t)n))) return rc;
   CK(cudaMemcpy(d_tab + me, &mine, sizeof mine, cudaMemcpyHostToDevice));
   if (h->nccl->AllGather) {
      NK(h->nccl->AllGather(d_tab + me, d_tab, sizeof(PeerTable), ncclChar, h->comm, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      CK(cudaMemcpy(all.data(), d_tab, sizeof(PeerTable) * n, cudaMemcpyDeviceToHost));
   } else all[me] = mine;

   // map every other rank's arena
   std::vector<char *> peer_arena(n, nullptr);
   int ok = want ? 1 : 0;
   for (int r = 0; r < n && ok; ++r) ok = all[r].ok;
   for (int r = 0; r < n && ok; ++r) {
      if (r == me) { peer_arena[r] = h->arena; continue; }
      if (all[r].pid == mine.pid) {
         int can = 0;
         if (cudaDeviceCanAccessPeer(&can, h->device, all[r].device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
         cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
         if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; }
         cudaGetLastError();
         peer_arena[r] = (char *)(uintptr_t)all[r].base;
      } else {
         void *blk = nullptr;
         if (cudaIpcOpenMemHandle(&blk, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError(); ok = 0; break;
         }
         h->ipc_opened.push_back(blk);
         peer_arena[r] = (char *)blk + all[r].alloc_off;
      }
   }

   // every rank must take the same path
   int *d_ok = nullptr;
   if ((rc = dev_zero(h, &d_ok, 1))) return rc;
   CK(cudaMemcpy(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice));
   NK(h->nccl->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream));
   CK(cudaStreamSynchronize(h->stream));
   CK(cudaMemcpy(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost));
   if (!ok) return 0;   // stay on NCCL send/recv

   // device tables
   std::vector<PeerMsg> node_msgs, face_msgs;
   std::vector<unsigned char> node_slot, face_slot;
   for (int i = 0; i < mine.nmsg; ++i) {
      const int r = pl.msg_rank[i];
      int j = -1;
      for (int t = 0; t < all[r].nmsg; ++t) if (all[r].msg_rank[t] == me) j = t;
      if (j < 0) return fail(LULESH_B200_EINVAL, "rank %d does not list rank %d as a neighbour", r, me);
      PeerMsg m;
      m.dst = reinterpret_cast<double *>(peer_arena[r] + all[r].off_fhalo) + all[r].msg_recv_off[j];
      m.flag = reinterpret_cast<unsigned long long *>(peer_arena[r]) + PEER_FLAG_NODE + j;
      m.send_off = pl.msg_send_off[i]; m.count = pl.msg_count[i];
      m.field_stride = pl.msg_count[i]; m.parity_stride = all[r].fhalo_size;
      node_msgs.push_back(m);
      node_slot.insert(node_slot.end(), (size_t)3 * m.count, (unsigned char)i);
   }
   for (int i = 0; i < mine.nface; ++i) {
      const int r = pl.face_rank[i];
      int j = -1;
      for (int t = 0; t < all[r].nface; ++t) if (all[r].face_rank[t] == me) j = t;
      if (j < 0) return fail(LULESH_B200_EINVAL, "rank %d does not list rank %d as a face neighbour", r, me);
      PeerMsg m;
      m.dst = reinterpret_cast<double *>(peer_arena[r] + all[r].off_delv) + all[r].face_ghost_off[j];
      m.flag = reinterpret_cast<unsigned long long *>(peer_arena[r]) + PEER_FLAG_FACE + j;
      m.send_off = pl.face_send_off[i]; m.count = pl.face_count[i];
      m.field_stride = all[r].allElem; m.parity_stride = 0;
      face_msgs.push_back(m);
      face_slot.insert(face_slot.end(), (size_t)3 * m.count, (unsigned char)i);
   }
   std::vector<DtSlot *> slots(n);
   for (int r = 0; r < n; ++r)
      slots[r] = reinterpret_cast<DtSlot *>(peer_arena[r] + PEER_NUM_FLAGS * sizeof(unsigned long long));
   if ((rc = dev_upload(h, &h->d_node_msgs, node_msgs.data(), node_msgs.size()))) return rc;
   if ((rc = dev_upload(h, &h->d_face_msgs, face_msgs.data(), face_msgs.size()))) return rc;
   if ((rc = dev_upload(h, &h->d_node_slot_msg, node_slot.data(), node_slot.size()))) return rc;
   if ((rc = dev_upload(h, &h->d_face_slot_msg, face_slot.data(), face_slot.size()))) return rc;
   if ((rc = dev_upload(h, &h->d_peer_slots, slots.data(), slots.size()))) return rc;
   if ((rc = dev_zero(h, &h->counters, 1))) return rc;
   P.peer_cnt = h->counters;
   h->p2p = true;
   h->halo_mode = "p2p";
   // nobody may start writing into a peer before that peer has finished zeroing its arena
   NK(h->nccl->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, h->comm, h->stream));
   CK(cudaStreamSynchronize(h->stream));
   return 0;
}

static int build_comm(lulesh_b200 *h, const lulesh_b200_host_view *v, const void *unique_id,
                      std::vector<unsigned char> &nodeFlags)
{
   KParams &P = h->P;
   int rc;
   if (!unique_id) return fail(LULESH_B200_EINVAL, "numRanks > 1 needs an NCCL unique id");
   h->nccl = nccl_api();
   if (!h->nccl) return fail(LULESH_B200_ENCCL, "libnccl.so.2 could not be loaded");

   lulesh_b200_halo_plan pl;
   if ((rc = build_halo_plan(v, pl))) return rc;
   for (int n : pl.bnode) nodeFlags[n] |= NODE_COMM;
   P.nbnode = (int)pl.bnode.size();
   P.fhalo_stride = pl.fhalo_size;
   h->send_total = pl.send_total;
   h->mq_total = pl.mq_total;
   for (size_t i = 0; i < pl.msg_rank.size(); ++i)
      h->msgs.push_back({pl.msg_rank[i], pl.msg_count[i], (size_t)pl.msg_send_off[i], (size_t)pl.msg_recv_off[i]});
   for (size_t i = 0; i < pl.face_rank.size(); ++i)
      h->faces.push_back({pl.face_rank[i], pl.face_count[i], (size_t)pl.face_send_off[i], (size_t)pl.face_ghost_off[i]});
   int *pi;
   if ((rc = dev_upload(h, &pi, pl.bnode.data(), pl.bnode.size()))) return rc;
   P.bnode = pi;
   if ((rc = dev_upload(h, &pi, pl.bsum_start.data(), pl.bsum_start.size()))) return rc;
   P.bsum_start = pi;
   if ((rc = dev_upload(h, &pi, pl.bsum_src.data(), pl.bsum_src.size()))) return rc;
   P.bsum_src = pi;
   if ((rc = dev_upload(h, &h->pack_idx, pl.pack_idx.data(), pl.pack_idx.size()))) return rc;
   {  // peer-visible arena: flags | dt slots | fhalo[2] (double-buffered by sequence parity) | delv
      const size_t off_fhalo = 4096;
      const size_t fhalo_bytes = (((size_t)2 * pl.fhalo_size * sizeof(double)) + 255) & ~(size_t)255;
      const size_t off_delv = off_fhalo + fhalo_bytes;
      h->arena_bytes = off_delv + (size_t)3 * P.allElem * sizeof(double);
      if ((rc = dev_zero(h, &h->arena, h->arena_bytes))) return rc;
      h->flags = reinterpret_cast<unsigned long long *>(h->arena);
      h->dtslots = reinterpret_cast<DtSlot *>(h->arena + PEER_NUM_FLAGS * sizeof(unsigned long long));
      static_assert(PEER_NUM_FLAGS * 8 + 2 * PEER_MAX_RANKS * sizeof(DtSlot) <= 4096, "arena header");
      P.fhalo = reinterpret_cast<double *>(h->arena + off_fhalo);
      double *g = reinterpret_cast<double *>(h->arena + off_delv);
      P.delv_xi = g; P.delv_eta = g + P.allElem; P.delv_zeta = g + 2 * (size_t)P.allElem;
   }
   if ((rc = dev_zero(h, &h->sendbuf, (size_t)pl.send_total))) return rc;
   if ((rc = dev_upload(h, &h->mq_idx, pl.mq_idx.data(), pl.mq_idx.size()))) return rc;
   if ((rc = dev_zero(h, &h->mq_send, (size_t)pl.mq_total))) return rc;

   ncclUniqueId id;
   memcpy(&id, unique_id, sizeof id);
   NK(h->nccl->CommInitRank(&h->comm, v->numRanks, id, v->rank));
   h->halo_mode = "nccl";
   return setup_p2p(h, v, pl);
}

extern "C" int lulesh_b200_create(const lulesh_b200_host_view *view, int device,
                                  const void *unique_id, lulesh_b200 **out)
{
   if (!view || !out) return fail(LULESH_B200_EINVAL, "null argument");
   *out = nullptr;
   lulesh_b200 *h = new lulesh_b200();
   const int rc = create_impl(h, view, device, unique_id, nullptr);
   if (rc) { lulesh_b200_destroy(h); return rc; }
   *out = h;
   return 0;
}

extern "C" int lulesh_b200_create_sedov(const lulesh_b200_sedov_params *p, int device,
                                        const void *unique_id, lulesh_b200 **out)
{
   if (!p || !out) return fail(LULESH_B200_EINVAL, "null argument");
   *out = nullptr;
   if (p->abi_version != LULESH_B200_ABI_VERSION) return fail(LULESH_B200_EINVAL, "abi_version mismatch");
   if (p->px < 1 || p->py < 1 || p->pz < 1 || p->numRanks != p->px * p->py * p->pz || p->rank < 0 ||
       p->rank >= p->numRanks || p->sx < 1 || p->sy < 1 || p->sz < 1 || p->numReg < 1)
      return fail(LULESH_B200_EINVAL, "inconsistent Sedov parameters");
   const long long ne = (long long)p->sx * p->sy * p->sz;
   if (8 * ne > INT_MAX) return fail(LULESH_B200_EINVAL, "brick too large for int32 indices");

   lulesh_b200_host_view v;
   memset(&v, 0, sizeof v);
   v.abi_version = LULESH_B200_ABI_VERSION;
   v.sizeX = p->sx; v.sizeY = p->sy; v.sizeZ = p->sz;
   v.numElem = (int)ne;
   v.numNode = (p->sx + 1) * (p->sy + 1) * (p->sz + 1);
   v.numRanks = p->numRanks; v.rank = p->rank;
   v.px = p->px; v.py = p->py; v.pz = p->pz;
   v.colLoc = p->rank % p->px; v.rowLoc = (p->rank / p->px) % p->py; v.planeLoc = p->rank / (p->px * p->py);
   v.numReg = p->numReg; v.cost = p->cost;

   SetupParams S;
   S.sx = p->sx; S.sy = p->sy; S.sz = p->sz; S.px = p->px; S.py = p->py; S.pz = p->pz;
   S.col = v.colLoc; S.row = v.rowLoc; S.plane = v.planeLoc; S.numRanks = p->numRanks;
   S.G = std::max(p->px * p->sx, std::max(p->py * p->sy, p->pz * p->sz));
   SedovInitialScalars(S.G, &v.scalars, &v.constants, &S.einit);

   // region index sets: sequential glibc rand(), host only (lulesh-init.cc:401-510)
   std::vector<Index_t> regNumList, regElemSize;
   std::vector<std::vector<Index_t>> lists;
   CreateRegionIndexSetsHost(p->rank, v.numElem, p->numReg, p->balance, regNumList, regElemSize, lists);
   std::vector<const int32_t *> ptrs(p->numReg);
   for (int r = 0; r < p->numReg; ++r) ptrs[r] = lists[r].data();
   v.regElemSize = regElemSize.data();
   v.regElemlist = ptrs.data();

   lulesh_b200 *h = new lulesh_b200();
   const int rc = create_impl(h, &v, device, unique_id, &S);
   if (rc) { lulesh_b200_destroy(h); return rc; }
   *out = h;
   return 0;
}

extern "C" void lulesh_b200_destroy(lulesh_b200 *h)
{
   if (!h) return;
   cudaSetDevice(h->device);
   if (h->stream) cudaStreamSynchronize(h->stream);
   if (h->comm && h->nccl) h->nccl->CommDestroy(h->comm);
   if (h->graph) cudaGraphExecDestroy(h->graph);
   for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
   for (void *p : h->allocs) cudaFree(p);
   if (h->h_ctl) cudaFreeHost(h->h_ctl);
   if (h->ev_a) cudaEventDestroy(h->ev_a);
   if (h->ev_b) cudaEventDestroy(h->ev_b);
   if (h->ev_fork) cudaEventDestroy(h->ev_fork);
   if (h->ev_dt) cudaEventDestroy(h->ev_dt);
   if (h->ev_t0) cudaEventDestroy(h->ev_t0);
   if (h->ev_t1) cudaEventDestroy(h->ev_t1);
   if (h->stream) cudaStreamDestroy(h->stream);
   if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
   delete h;
}

// --------------------------------------------------------------------------
// the cycle
// --------------------------------------------------------------------------

// NCCL fallback of the node-halo exchange of the three force planes (or the replicated
// mass), entirely on stream `cs`: pack, one group of send/recv per neighbour.
static int exchange_nodes(lulesh_b200 *h, cudaStream_t cs)
{
   const KParams &P = h->P;
   k_gather_index<<<blocks_for((int)h->send_total, 256), 256, 0, cs>>>(
      h->sendbuf, P.fhalo, h->pack_idx, (int)h->send_total);
   NK(h->nccl->GroupStart());
   for (const Message &m : h->msgs) {
      NK(h->nccl->Recv(P.fhalo + m.recv_off, (size_t)3 * m.count, ncclDouble, m.rank, h->comm, cs));
      NK(h->nccl->Send(h->sendbuf + m.send_off, (size_t)3 * m.count, ncclDouble, m.rank, h->comm, cs));
   }
   NK(h->nccl->GroupEnd());
   h->launches += 1;
   return 0;
}

static int exchange_monoq(lulesh_b200 *h, cudaStream_t cs)
{
   const KParams &P = h->P;
   k_gather_index<<<blocks_for((int)h->mq_total, 256), 256, 0, cs>>>(
      h->mq_send, P.delv_xi, h->mq_idx, (int)h->mq_total);
   NK(h->nccl->GroupStart());
   for (const FaceMessage &f : h->faces)
      for (int a = 0; a < 3; ++a) {   // receive straight into the ghost slots: zero unpack
         NK(h->nccl->Recv(P.delv_xi + (size_t)a * P.allElem + f.ghost_off, f.count, ncclDouble, f.rank,
                          h->comm, cs));
         NK(h->nccl->Send(h->mq_send + f.send_off + (size_t)a * f.count, f.count, ncclDouble, f.rank,
                          h->comm, cs));
      }
   NK(h->nccl->GroupEnd());
   h->launches += 1;
   return 0;
}

// Timeline events of one cycle (lulesh_b200_timeline / per-kernel timing).  M_* are recorded on
// the main stream, C_* on the comm stream (several ranks only).
enum {
   M_START = 0, M_K1_START, M_K1_END, M_K2_END, M_NODE_JOIN, M_K3_END, M_K45I_END, M_MONOQ_JOIN, M_END,
   C_DT_START, C_DT_END, C_NODE_START, C_NODE_END, C_MQ_START, C_MQ_END, TL_EVENTS
};

// enqueue one TimeIncrement + LagrangeLeapFrog on h->stream (no host sync).
// `tl` (TL_EVENTS events or null) receives the timeline; the schedule is the same either way.
static int enqueue_cycle(lulesh_b200 *h, cudaEvent_t *tl)
{
   const KParams &P = h->P;
   cudaStream_t s = h->stream, cs = h->comm_stream;
   const int dbg = h->debug;
   const bool multi = h->numRanks > 1;
#define MARK(id, stream) do { if (tl) CK(cudaEventRecord(tl[id], stream)); } while (0)
   MARK(M_START, s);
   // K1 does not use dt (only K2/K3 do, lulesh.cc:1230,1577), so at several ranks the
   // TimeIncrement chain with its min-allreduce runs on the comm stream underneath K1.
   if (!multi) {
      k_time_increment<<<1, 32, 0, s>>>(P.ctl, 0);
   } else {
      CK(cudaEventRecord(h->ev_fork, s));
      CK(cudaStreamWaitEvent(cs, h->ev_fork, 0));
      MARK(C_DT_START, cs);
      if (h->p2p) {   // candidates are written into every rank's slot table over NVLink
         k_peer_dt_post<<<1, PEER_MAX_RANKS, 0, cs>>>(P.ctl, h->d_peer_slots, h->rank, h->numRanks,
                                                      &h->counters->dt_seq);
         k_peer_dt_wait<<<1, PEER_MAX_RANKS, 0, cs>>>(P.ctl, h->dtslots, h->numRanks, &h->counters->dt_expect);
      } else {
         k_time_increment<<<1, 32, 0, cs>>>(P.ctl, 1);
         NK(h->nccl->AllReduce(&P.ctl->gnewdt, &P.ctl->gnewdt, 1, ncclDouble, ncclMin, h->comm, cs));
      }
      k_time_increment<<<1, 32, 0, cs>>>(P.ctl, 2);
      CK(cudaEventRecord(h->ev_dt, cs));
      MARK(C_DT_END, cs);
      h->launches += 2;
   }
   MARK(M_K1_START, s);
   k_force<<<h->k1_grid, K1_THREADS, K1_SMEM_BYTES, s>>>(P);
   MARK(M_K1_END, s);
   if (multi) {
      // Shared nodes: gather own partials -> ship them to the neighbours -> wait for theirs ->
      // sum in rank order and advance.  The whole chain runs on the comm stream underneath the
      // interior node update (it follows the dt chain there, which boundary_update needs anyway).
      CK(cudaEventRecord(h->ev_a, s));                 // K1 finished
      CK(cudaStreamWaitEvent(cs, h->ev_a, 0));
      MARK(C_NODE_START, cs);
      k_node_boundary_gather<<<blocks_for(P.nbnode, 128), 128, 0, cs>>>(P);
      if (h->p2p) {
         k_peer_pack<<<blocks_for((int)h->send_total, 256), 256, 0, cs>>>(
            P.fhalo, P.fhalo_stride, h->pack_idx, h->d_node_slot_msg, (int)h->send_total, h->d_node_msgs,
            (int)h->msgs.size(), &h->counters->node_done, &h->counters->node_seq);
         k_peer_wait<<<1, 32, 0, cs>>>(h->flags, PEER_FLAG_NODE, (int)h->msgs.size(), &h->counters->node_expect, P.ctl);
         h->launches += 2;
      } else {
         int rc;
         if ((rc = exchange_nodes(h, cs))) return rc;
         h->launches += 1;
      }
      k_node_boundary_update<<<blocks_for(P.nbnode, 128), 128, 0, cs>>>(P, dbg);
      CK(cudaEventRecord(h->ev_b, cs));
      MARK(C_NODE_END, cs);
      CK(cudaStreamWaitEvent(s, h->ev_dt, 0));
      h->launches += 2;
   }
   k_node<<<blocks_for(P.nn, K2_THREADS), K2_THREADS, 0, s>>>(P, dbg);   // interior nodes at several ranks
   MARK(M_K2_END, s);
   if (multi) CK(cudaStreamWaitEvent(s, h->ev_b, 0));
   MARK(M_NODE_JOIN, s);
   k_kinematics<<<h->k3_grid, K3_THREADS, K3_SMEM_BYTES, s>>>(P);
   MARK(M_K3_END, s);
   if (multi) {
      // MonoQ exchange (CommMonoQ, lulesh-comm.cc:1684) on the comm stream; the elements that read
      // no ghost slot start right away on the main stream and hide it.
      CK(cudaEventRecord(h->ev_a, s));                 // K3 finished
      CK(cudaStreamWaitEvent(cs, h->ev_a, 0));
      MARK(C_MQ_START, cs);
      if (h->p2p) {
         k_peer_pack<<<blocks_for((int)h->mq_total, 256), 256, 0, cs>>>(
            P.delv_xi, 0, h->mq_idx, h->d_face_slot_msg, (int)h->mq_total, h->d_face_msgs,
            (int)h->faces.size(), &h->counters->face_done, &h->counters->face_seq);
         k_peer_wait<<<1, 32, 0, cs>>>(h->flags, PEER_FLAG_FACE, (int)h->faces.size(), &h->counters->face_expect, P.ctl);
         h->launches += 2;
      } else {
         int rc;
         if ((rc = exchange_monoq(h, cs))) return rc;
         h->launches += 1;
      }
      // The face-layer elements follow the exchange on the comm stream itself: a few hundred blocks
      // whose run time is the latency of one block (up to 10 (1 + cost) EOS repetitions), hidden
      // under the interior launch instead of trailing it.  Both launches merge their dt minima
      // with the same integer atomicMin.
      if (P.numWorkBlocks > P.numWorkBlocksInterior)
         (P.unit_rho0 ? k_material : k_material_rho0)<<<P.numWorkBlocks - P.numWorkBlocksInterior, MAT_THREADS, 0, cs>>>(P, dbg, P.numWorkBlocksInterior);
      CK(cudaEventRecord(h->ev_b, cs));
      MARK(C_MQ_END, cs);
      if (P.numWorkBlocksInterior > 0)
         (P.unit_rho0 ? k_material : k_material_rho0)<<<P.numWorkBlocksInterior, MAT_THREADS, 0, s>>>(P, dbg, 0);
      MARK(M_K45I_END, s);
      CK(cudaStreamWaitEvent(s, h->ev_b, 0));
      MARK(M_MONOQ_JOIN, s);
      h->launches += 1;
   } else {
      MARK(M_K45I_END, s);
      MARK(M_MONOQ_JOIN, s);
      (P.unit_rho0 ? k_material : k_material_rho0)<<<P.numWorkBlocks, MAT_THREADS, 0, s>>>(P, dbg, 0);
   }
   MARK(M_END, s);
#undef MARK
   h->launches += 5;
   CK(cudaGetLastError());
   return 0;
}

static int ensure_graph(lulesh_b200 *h)
{
   if (h->graph && h->graph_debug == h->debug) return 0;
   if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
   cudaGraph_t g;
   const int64_t saved = h->launches;
   CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
   int rc = enqueue_cycle(h, nullptr);
   cudaError_t e = cudaStreamEndCapture(h->stream, &g);
   h->launches_per_cycle = (int)(h->launches - saved);
   h->launches = saved;
   if (rc) return rc;
   if (e != cudaSuccess) return fail(LULESH_B200_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
   CK(cudaGraphInstantiate(&h->graph, g, 0));
   CK(cudaGraphDestroy(g));
   h->graph_debug = h->debug;
   return 0;
}

static bool use_graph(const lulesh_b200 *h)
{
   // A cycle is captured once and replayed: 5 kernels at one rank; kernels + peer-to-peer exchange
   // kernels on two streams at several ranks.  The NCCL fallback launches its cycles eagerly:
   // a captured cycle with ncclAllReduce / ncclSend / ncclRecv on the forked stream reported an
   // NCCL internal error in round 1 and, with an eager warm-up pass and relaxed capture mode,
   // dead-locked in round 2 (NCCL 2.27.3 and 2.28.9, threads and separate processes, with and
   // without NCCL_GRAPH_REGISTER).  LULESH_B200_NO_GRAPH=1 launches every cycle eagerly.
   static const bool disabled = getenv("LULESH_B200_NO_GRAPH") != nullptr;
   if (disabled) return false;
   return h->numRanks == 1 || h->p2p;
}

static int enqueue_cycles(lulesh_b200 *h, int n)
{
   int rc;
   if (use_graph(h)) {
      if ((rc = ensure_graph(h))) return rc;
      for (int i = 0; i < n; ++i) CK(cudaGraphLaunch(h->graph, h->stream));
      h->launches += (int64_t)h->launches_per_cycle * n;
   } else {
      for (int i = 0; i < n; ++i)
         if ((rc = enqueue_cycle(h, nullptr))) return rc;
   }
   return 0;
}

static int fetch_ctl(lulesh_b200 *h)
{
   CK(cudaMemcpyAsync(h->h_ctl, h->P.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
   CK(cudaStreamSynchronize(h->stream));
   return 0;
}

// New cycle limit for the device-side loop condition (lulesh.cc:2745).  If the loop can go on
// under the new limit, `done` is cleared here, before any cycle is enqueued: the force kernel
// reads it concurrently with the dt chain that would otherwise clear it (see k_force).
static int set_max_cycles(lulesh_b200 *h, int max_cycles)
{
   int rc;
   if ((rc = fetch_ctl(h))) return rc;
   const Ctl &c = *h->h_ctl;
   CK(cudaMemcpyAsync(&h->P.ctl->max_cycles, &max_cycles, sizeof(int), cudaMemcpyHostToDevice, h->stream));
   if (c.error == 0 && c.time < c.stoptime && c.cycle < max_cycles)
      CK(cudaMemsetAsync(&h->P.ctl->done, 0, sizeof(int), h->stream));
   CK(cudaStreamSynchronize(h->stream));   // max_cycles is a stack variable
   return 0;
}

extern "C" int lulesh_b200_sum_nodal_mass(lulesh_b200 *h)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   if (h->numRanks == 1) return 0;
   CK(cudaSetDevice(h->device));
   const KParams &P = h->P;
   double *mass = h->field_ptr[LULESH_F_NODALMASS];
   // own partial mass into the first plane of fhalo (the exchange ships 3 planes;
   // the other two carry the same values and are ignored)
   for (int a = 0; a < 3; ++a)
      k_gather_index<<<blocks_for(P.nbnode, 256), 256, 0, h->stream>>>(
         P.fhalo + (size_t)a * P.nbnode, mass, P.bnode, P.nbnode);
   int rc;
   if ((rc = exchange_nodes(h, h->stream))) return rc;
   k_boundary_mass<<<blocks_for(P.nbnode, 128), 128, 0, h->stream>>>(P, mass);
   CK(cudaStreamSynchronize(h->stream));   // doubles as the MPI_Barrier of lulesh.cc:2732
   return 0;
}

extern "C" int lulesh_b200_run(lulesh_b200 *h, int32_t max_cycles, int32_t sync_every,
                               lulesh_b200_progress_cb cb, void *user)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = set_max_cycles(h, max_cycles))) return rc;
   if (sync_every <= 0) sync_every = 64;
   if (cb) sync_every = 1;
   if ((rc = fetch_ctl(h))) return rc;
   while (h->h_ctl->error == 0 && h->h_ctl->time < h->h_ctl->stoptime && h->h_ctl->cycle < max_cycles) {
      const int before = h->h_ctl->cycle;
      const int batch = std::min<int>(sync_every, max_cycles - before);
      if ((rc = enqueue_cycles(h, batch))) return rc;
      if ((rc = fetch_ctl(h))) return rc;
      if (cb && h->h_ctl->cycle != before) cb(h->h_ctl->cycle, h->h_ctl->time, h->h_ctl->deltatime, user);
   }
   return h->h_ctl->error;
}

extern "C" int lulesh_b200_step(lulesh_b200 *h)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = set_max_cycles(h, INT_MAX))) return rc;
   if ((rc = enqueue_cycles(h, 1))) return rc;
   if ((rc = fetch_ctl(h))) return rc;
   return h->h_ctl->error;
}

// Average per-cycle timeline over `cycles` eagerly launched cycles (same schedule and stream
// layout as the graph: nothing is serialised for the measurement).
static int run_timeline(lulesh_b200 *h, int cycles, float *out)
{
   int rc;
   cudaEvent_t ev[TL_EVENTS];
   for (auto &e : ev) CK(cudaEventCreate(&e));
   double acc[LULESH_B200_TIMELINE_N] = {0};
   const bool multi = h->numRanks > 1;
   auto span = [&](int a, int b, double *dst) -> int {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, ev[a], ev[b]));
      *dst += ms;
      return 0;
   };
   for (int i = 0; i < cycles; ++i) {
      if ((rc = enqueue_cycle(h, ev))) return rc;
      CK(cudaStreamSynchronize(h->stream));
      if (multi) CK(cudaStreamSynchronize(h->comm_stream));
      if ((rc = span(M_START, M_K1_START, &acc[LULESH_TL_TIME_INCREMENT]))) return rc;
      if ((rc = span(M_K1_START, M_K1_END, &acc[LULESH_TL_K1]))) return rc;
      if ((rc = span(M_K1_END, M_K2_END, &acc[LULESH_TL_K2]))) return rc;
      if ((rc = span(M_K2_END, M_NODE_JOIN, &acc[LULESH_TL_NODE_JOIN_WAIT]))) return rc;
      if ((rc = span(M_NODE_JOIN, M_K3_END, &acc[LULESH_TL_K3]))) return rc;
      if ((rc = span(M_K3_END, M_K45I_END, &acc[LULESH_TL_K45_INTERIOR]))) return rc;
      if ((rc = span(M_K45I_END, M_MONOQ_JOIN, &acc[LULESH_TL_MONOQ_JOIN_WAIT]))) return rc;
      if ((rc = span(M_MONOQ_JOIN, M_END, &acc[LULESH_TL_K45_TAIL]))) return rc;
      if ((rc = span(M_START, M_END, &acc[LULESH_TL_CYCLE]))) return rc;
      if (multi) {
         if ((rc = span(C_DT_START, C_DT_END, &acc[LULESH_TL_COMM_DT]))) return rc;
         if ((rc = span(C_NODE_START, C_NODE_END, &acc[LULESH_TL_COMM_NODE]))) return rc;
         if ((rc = span(C_MQ_START, C_MQ_END, &acc[LULESH_TL_COMM_MONOQ]))) return rc;
      }
   }
   for (auto &e : ev) cudaEventDestroy(e);
   for (int k = 0; k < LULESH_B200_TIMELINE_N; ++k) out[k] = cycles > 0 ? (float)(acc[k] / cycles) : 0.f;
   return 0;
}

extern "C" int lulesh_b200_timeline(lulesh_b200 *h, int32_t cycles, float *out_ms)
{
   if (!h || cycles < 1 || !out_ms) return fail(LULESH_B200_EINVAL, "bad argument");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = set_max_cycles(h, INT_MAX))) return rc;
   if ((rc = run_timeline(h, cycles, out_ms))) return rc;
   if ((rc = fetch_ctl(h))) return rc;
   return h->h_ctl->error;
}

extern "C" int lulesh_b200_time_cycles(lulesh_b200 *h, int32_t cycles, float *total_ms,
                                       float *per_kernel_ms, int64_t *launches)
{
   if (!h || cycles < 0) return fail(LULESH_B200_EINVAL, "bad argument");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = set_max_cycles(h, INT_MAX))) return rc;
   const int64_t l0 = h->launches;
   if (!per_kernel_ms) {
      if (use_graph(h) && (rc = ensure_graph(h))) return rc;
      CK(cudaStreamSynchronize(h->stream));
      CK(cudaEventRecord(h->ev_t0, h->stream));
      if ((rc = enqueue_cycles(h, cycles))) return rc;
      CK(cudaEventRecord(h->ev_t1, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      if (total_ms) CK(cudaEventElapsedTime(total_ms, h->ev_t0, h->ev_t1));
   } else {
      float tl[LULESH_B200_TIMELINE_N];
      if ((rc = run_timeline(h, cycles, tl))) return rc;
      // the five main-stream intervals of a cycle (waits for the comm stream included)
      per_kernel_ms[0] = tl[LULESH_TL_TIME_INCREMENT] * cycles;
      per_kernel_ms[1] = tl[LULESH_TL_K1] * cycles;
      per_kernel_ms[2] = (tl[LULESH_TL_K2] + tl[LULESH_TL_NODE_JOIN_WAIT]) * cycles;
      per_kernel_ms[3] = tl[LULESH_TL_K3] * cycles;
      per_kernel_ms[4] = (tl[LULESH_TL_K45_INTERIOR] + tl[LULESH_TL_MONOQ_JOIN_WAIT] + tl[LULESH_TL_K45_TAIL]) * cycles;
      if (total_ms) *total_ms = tl[LULESH_TL_CYCLE] * cycles;
   }
   if (launches) *launches = h->launches - l0;
   if ((rc = fetch_ctl(h))) return rc;
   return h->h_ctl->error;
}

// --------------------------------------------------------------------------
// scalars, fields, per-kernel entry points
// --------------------------------------------------------------------------
extern "C" int lulesh_b200_get_scalars(lulesh_b200 *h, lulesh_b200_scalars *out)
{
   if (!h || !out) return fail(LULESH_B200_EINVAL, "null argument");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = fetch_ctl(h))) return rc;
   const Ctl &c = *h->h_ctl;
   memcpy(&out->dtcourant, &c.dtcourant_bits, 8);
   memcpy(&out->dthydro, &c.dthydro_bits, 8);
   out->dtfixed = c.dtfixed; out->time = c.time; out->deltatime = c.deltatime;
   out->deltatimemultlb = c.deltatimemultlb; out->deltatimemultub = c.deltatimemultub;
   out->dtmax = c.dtmax; out->stoptime = c.stoptime; out->cycle = c.cycle; out->error = c.error;
   return 0;
}

extern "C" int lulesh_b200_set_scalars(lulesh_b200 *h, const lulesh_b200_scalars *in)
{
   if (!h || !in) return fail(LULESH_B200_EINVAL, "null argument");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = fetch_ctl(h))) return rc;
   Ctl c = *h->h_ctl;
   memcpy(&c.dtcourant_bits, &in->dtcourant, 8);
   memcpy(&c.dthydro_bits, &in->dthydro, 8);
   c.dtfixed = in->dtfixed; c.time = in->time; c.deltatime = in->deltatime;
   c.deltatimemultlb = in->deltatimemultlb; c.deltatimemultub = in->deltatimemultub;
   c.dtmax = in->dtmax; c.stoptime = in->stoptime; c.cycle = in->cycle; c.error = in->error;
   c.done = 0;
   c.pending_error = 0;
   CK(cudaMemcpy(h->P.ctl, &c, sizeof c, cudaMemcpyHostToDevice));
   return 0;
}

extern "C" size_t lulesh_b200_field_count(lulesh_b200 *h, int field)
{
   if (!h || field < 0 || field >= LULESH_F_COUNT) return 0;
   return h->field_cnt[field];
}

extern "C" int lulesh_b200_download(lulesh_b200 *h, int field, double *dst, size_t count)
{
   if (!h || !dst || field < 0 || field >= LULESH_F_COUNT) return fail(LULESH_B200_EINVAL, "bad argument");
   if (count != h->field_cnt[field])
      return fail(LULESH_B200_EINVAL, "field %d has %zu entries, not %zu", field, h->field_cnt[field], count);
   CK(cudaSetDevice(h->device));
   CK(cudaStreamSynchronize(h->stream));
   CK(cudaMemcpy(dst, h->field_ptr[field], count * sizeof(double), cudaMemcpyDeviceToHost));
   return 0;
}

extern "C" int lulesh_b200_upload(lulesh_b200 *h, int field, const double *src, size_t count)
{
   if (!h || !src || field < 0 || field >= LULESH_F_COUNT) return fail(LULESH_B200_EINVAL, "bad argument");
   if (count != h->field_cnt[field])
      return fail(LULESH_B200_EINVAL, "field %d has %zu entries, not %zu", field, h->field_cnt[field], count);
   CK(cudaSetDevice(h->device));
   CK(cudaStreamSynchronize(h->stream));
   CK(cudaMemcpy(h->field_ptr[field], src, count * sizeof(double), cudaMemcpyHostToDevice));
   return 0;
}

extern "C" int lulesh_b200_set_debug(lulesh_b200 *h, int on)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   h->debug = on ? 1 : 0;
   return 0;
}

static int finish_kernel(lulesh_b200 *h)
{
   CK(cudaGetLastError());
   int rc;
   if ((rc = fetch_ctl(h))) return rc;
   return h->h_ctl->error;
}

extern "C" int lulesh_b200_kernel_time_increment(lulesh_b200 *h)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   if (h->numRanks != 1) return fail(LULESH_B200_EINVAL, "per-kernel entry points are single-rank");
   CK(cudaSetDevice(h->device));
   int rc;
   if ((rc = set_max_cycles(h, INT_MAX))) return rc;
   k_time_increment<<<1, 32, 0, h->stream>>>(h->P.ctl, 0);
   return finish_kernel(h);
}

extern "C" int lulesh_b200_kernel_force(lulesh_b200 *h)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   CK(cudaSetDevice(h->device));
   k_force<<<h->k1_grid, K1_THREADS, K1_SMEM_BYTES, h->stream>>>(h->P);
   int rc = finish_kernel(h);
   if (rc == 0 && h->h_ctl->pending_error != 0) {   // standalone launch: promote K1's abort test here (k_node does it in a cycle)
      rc = h->h_ctl->pending_error;
      const int v[2] = {rc, 0};
      CK(cudaMemcpy(&h->P.ctl->error, &v[0], sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(&h->P.ctl->pending_error, &v[1], sizeof(int), cudaMemcpyHostToDevice));
   }
   return rc;
}

extern "C" int lulesh_b200_kernel_node(lulesh_b200 *h, int materialise_debug)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   if (h->numRanks != 1) return fail(LULESH_B200_EINVAL, "per-kernel entry points are single-rank");
   CK(cudaSetDevice(h->device));
   k_node<<<blocks_for(h->P.nn, K2_THREADS), K2_THREADS, 0, h->stream>>>(h->P, materialise_debug);
   return finish_kernel(h);
}

extern "C" int lulesh_b200_kernel_kinematics(lulesh_b200 *h)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   CK(cudaSetDevice(h->device));
   k_kinematics<<<h->k3_grid, K3_THREADS, K3_SMEM_BYTES, h->stream>>>(h->P);
   return finish_kernel(h);
}

extern "C" int lulesh_b200_kernel_material(lulesh_b200 *h)
{
   if (!h) return fail(LULESH_B200_EINVAL, "null handle");
   if (h->numRanks != 1) return fail(LULESH_B200_EINVAL, "per-kernel entry points are single-rank");
   CK(cudaSetDevice(h->device));
   (h->P.unit_rho0 ? k_material : k_material_rho0)<<<h->P.numWorkBlocks, MAT_THREADS, 0, h->stream>>>(h->P, 1, 0);
   return finish_kernel(h);
}

extern "C" const char *lulesh_b200_halo_mode(lulesh_b200 *h) { return h ? h->halo_mode.c_str() : ""; }
extern "C" size_t lulesh_b200_device_bytes(lulesh_b200 *h) { return h ? h->device_bytes : 0; }
extern "C" size_t lulesh_b200_upload_bytes(lulesh_b200 *h) { return h ? h->upload_bytes : 0; }
