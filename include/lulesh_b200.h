/*
 * lulesh_b200.h -- C ABI of the B200-native Lagrange-leapfrog time step.
 *
 * This is the drop-in boundary for ONE path of LLNL/LULESH 2.0: the body of
 * main()'s timed loop,
 *
 *     while (time < stoptime && cycle < its) {      // lulesh.cc:2745
 *        TimeIncrement(domain);                     // lulesh.cc:2747 (167-222)
 *        LagrangeLeapFrog(domain);                  // lulesh.cc:2748 (2601-2645)
 *     }
 *
 * The reference has no plugin/FFI interface; the host driver owns a `Domain`
 * (lulesh.h:148-595), hands read-only views of its arrays to
 * lulesh_b200_create(), runs the loop on the device, and reads the results
 * back through lulesh_b200_download() for VerifyAndWriteFinalOutput
 * (lulesh-util.cc:175-230).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions (mirroring the reference):
 *   - Real_t  = double  (lulesh.h:39), Index_t = int32_t (lulesh.h:38).
 *   - status: 0 ok, -1 VolumeError, -2 QStopError (lulesh.h:42); other
 *     negative values are infrastructure failures (see enum below).
 *   - one handle per GPU, used by one host thread at a time
 *     (MPI_THREAD_FUNNELED in the reference, lulesh.cc:2663).
 *   - the library owns all device memory, the caller owns all host memory.
 *   - there is NO CPU fallback: every entry point that computes fails with
 *     LULESH_B200_ECUDA when no sm_100 device is usable.
 */
#ifndef LULESH_B200_H
#define LULESH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LULESH_B200_ABI_VERSION 1

typedef double  lulesh_real_t;   /* lulesh.h:39 Real_t  */
typedef int32_t lulesh_index_t;  /* lulesh.h:38 Index_t */

enum lulesh_b200_status {
   LULESH_B200_OK           =  0,
   LULESH_B200_VOLUME_ERROR = -1,  /* lulesh.h:42 VolumeError */
   LULESH_B200_QSTOP_ERROR  = -2,  /* lulesh.h:42 QStopError  */
   LULESH_B200_EINVAL       = -10, /* bad argument / inconsistent view */
   LULESH_B200_ECUDA        = -11, /* no device, launch or runtime failure */
   LULESH_B200_ENCCL        = -12, /* NCCL missing or failed */
   LULESH_B200_ENOMEM       = -13
};

/* Field ids for download/upload.  Names are the reference accessors
 * (lulesh.h:266-373).  N = per node, E = per element, A = allElem (elements +
 * ghost slots, lulesh.cc:1955-1958). */
enum lulesh_b200_field {
   LULESH_F_X = 0, LULESH_F_Y, LULESH_F_Z,               /* N  lulesh.h:266-268 */
   LULESH_F_XD, LULESH_F_YD, LULESH_F_ZD,                /* N  lulesh.h:271-273 */
   LULESH_F_XDD, LULESH_F_YDD, LULESH_F_ZDD,             /* N  lulesh.h:276-278 (debug mirror) */
   LULESH_F_FX, LULESH_F_FY, LULESH_F_FZ,                /* N  lulesh.h:281-283 (debug mirror) */
   LULESH_F_NODALMASS,                                   /* N  lulesh.h:286 */
   LULESH_F_E, LULESH_F_P, LULESH_F_Q,                   /* E  lulesh.h:337-343 */
   LULESH_F_QL, LULESH_F_QQ,                             /* E  lulesh.h:346-348 */
   LULESH_F_V, LULESH_F_VOLO, LULESH_F_VNEW,             /* E  lulesh.h:351,355,324 */
   LULESH_F_DELV, LULESH_F_VDOV, LULESH_F_AREALG,        /* E  lulesh.h:352,358,361 */
   LULESH_F_SS, LULESH_F_ELEMMASS,                       /* E  lulesh.h:364,367 */
   LULESH_F_DELV_XI, LULESH_F_DELV_ETA, LULESH_F_DELV_ZETA, /* A lulesh.h:327-329 */
   LULESH_F_DELX_XI, LULESH_F_DELX_ETA, LULESH_F_DELX_ZETA, /* E lulesh.h:332-334 */
   LULESH_F_COUNT
};

/* The 19 material/cut-off constants of Domain (lulesh.h:534-555,
 * values lulesh-init.cc:20-38). */
typedef struct lulesh_b200_constants {
   double e_cut, p_cut, q_cut, v_cut, u_cut;
   double hgcoef, ss4o3, qstop, monoq_max_slope, monoq_limiter_mult;
   double qlc_monoq, qqc_monoq, qqc, eosvmax, eosvmin;
   double pmin, emin, dvovmax, refdens;
} lulesh_b200_constants;

/* Time-step control block (lulesh.h:558-567). */
typedef struct lulesh_b200_scalars {
   double  dtcourant, dthydro;
   double  dtfixed, time, deltatime;
   double  deltatimemultlb, deltatimemultub, dtmax, stoptime;
   int32_t cycle;
   int32_t error;      /* sticky device error word: 0 / -1 / -2 */
} lulesh_b200_scalars;

/* Read-only view of a host Domain, copied to HBM by lulesh_b200_create.
 * Every pointer is a plain host array with the reference's layout. */
typedef struct lulesh_b200_host_view {
   int32_t abi_version;                 /* LULESH_B200_ABI_VERSION */

   /* local brick (lulesh.h:420-426) */
   int32_t sizeX, sizeY, sizeZ;         /* elements per edge of this rank's box */
   int32_t numElem, numNode;

   /* decomposition; generalises m_tp (lulesh-init.cc:59) to (px,py,pz) */
   int32_t numRanks, rank;
   int32_t px, py, pz;                  /* ranks per axis (col,row,plane) */
   int32_t colLoc, rowLoc, planeLoc;    /* lulesh.h:415-417 */

   /* node-centred state, double[numNode] (lulesh.h:460-476) */
   const double *x, *y, *z, *xd, *yd, *zd, *nodalMass;

   /* symmetry-plane node sets (lulesh.h:478-480); NULL/0 when absent */
   const int32_t *symmX, *symmY, *symmZ;
   int32_t numSymmX, numSymmY, numSymmZ;

   /* element-centred, [numElem] unless noted (lulesh.h:491-531) */
   const int32_t *nodelist;             /* int[8*numElem], element-major */
   const int32_t *lxim, *lxip, *letam, *letap, *lzetam, *lzetap;
   const int32_t *elemBC;
   const double  *e, *p, *q, *v, *volo, *ss, *elemMass;

   /* regions (lulesh.h:485-489): region r owns regElemlist[r][0..regElemSize[r]) */
   int32_t numReg, cost;
   const int32_t *regElemSize;
   const int32_t *const *regElemlist;

   /* node -> element-corner CSR (lulesh.h:587-588); entries are elem*8+corner
    * in ascending element order (lulesh-init.cc:310-319).  This fixes the
    * deterministic force-gather order. */
   const int32_t *nodeElemStart;        /* int[numNode+1] */
   const int32_t *nodeElemCornerList;   /* int[nodeElemStart[numNode]] */

   lulesh_b200_constants constants;
   lulesh_b200_scalars   scalars;       /* initial time controls */
} lulesh_b200_host_view;

typedef struct lulesh_b200 lulesh_b200;           /* opaque per-GPU handle */

/* -p progress callback: called once per cycle on the calling thread
 * (lulesh.cc:2750-2756 prints cycle, time, dt). */
typedef void (*lulesh_b200_progress_cb)(int32_t cycle, double time, double dt,
                                        void *user);

/* 128-byte NCCL unique id, produced on rank 0 and shipped to the other ranks
 * by whatever launcher the host uses (torch.distributed, threads, files). */
#define LULESH_B200_UNIQUE_ID_BYTES 128
int lulesh_b200_get_unique_id(void *out_id /* 128 bytes */);

/* Replaces the tail of Domain::Domain + SetupCommBuffers (lulesh-init.cc:
 * 16-194, 342-396): uploads the view to device `device`, builds the device
 * layouts (corner-gather table, region work list, halo index lists) and, when
 * view->numRanks > 1, joins the NCCL communicator identified by `unique_id`.
 * `unique_id` may be NULL when numRanks == 1. */
int lulesh_b200_create(const lulesh_b200_host_view *view, int device,
                       const void *unique_id, lulesh_b200 **out);

/* Device-side alternative to building a host Domain first: replaces
 * Domain::Domain (lulesh-init.cc:16-194) for the Sedov problem.  BuildMesh (218-267), the
 * node->corner lists (272-337), symmetry planes, connectivity and boundary conditions
 * (514-673), volo / elemMass / nodalMass (159-178), the energy deposit and dt0 (183-192) are
 * generated by setup kernels directly in HBM, bit-identical to the host Domain; only the
 * region index sets (sequential glibc rand(), 401-510) are built on the host. */
typedef struct lulesh_b200_sedov_params {
   int32_t abi_version;                 /* LULESH_B200_ABI_VERSION */
   int32_t numRanks, rank;
   int32_t px, py, pz;                  /* ranks per axis; the reference: px = py = pz = tp */
   int32_t sx, sy, sz;                  /* elements per edge of this rank's brick; the reference: nx */
   int32_t numReg, balance, cost;       /* -r -b -c */
} lulesh_b200_sedov_params;
int lulesh_b200_create_sedov(const lulesh_b200_sedov_params *params, int device,
                             const void *unique_id, lulesh_b200 **out);

/* Replaces the initial nodalMass halo sum (lulesh.cc:2720-2729) and the
 * MPI_Barrier that follows (lulesh.cc:2732).  No-op at numRanks == 1. */
int lulesh_b200_sum_nodal_mass(lulesh_b200 *h);

/* Replaces the timed while loop (lulesh.cc:2745-2757): advances until
 * time >= stoptime or cycle >= max_cycles.  `sync_every` > 0 bounds how many
 * cycles are enqueued between host polls of the control block (cycles issued
 * after termination are device-side no-ops).  `cb` (may be NULL) forces a
 * per-cycle poll, like -p.  Returns 0 / -1 / -2 as the reference's exit
 * codes (lulesh.h:42).
 * Several ranks: the ranks of one communicator advance in lockstep (every cycle
 * contains exchanges that all of them take part in), so EVERY rank must call
 * this with the same `max_cycles` and the same `sync_every`; a rank with a
 * callback polls every cycle, so when any rank passes `cb` the others pass
 * sync_every = 1.  A device error on one rank (-1 / -2) reaches all ranks with
 * the next dt reduction, and every rank returns it from the same cycle. */
int lulesh_b200_run(lulesh_b200 *h, int32_t max_cycles, int32_t sync_every,
                    lulesh_b200_progress_cb cb, void *user);

/* One TimeIncrement + LagrangeLeapFrog (lulesh.cc:2747-2748), synchronous. */
int lulesh_b200_step(lulesh_b200 *h);

/* time(), deltatime(), cycle(), dtcourant(), dthydro() ... (lulesh.h:402-412) */
int lulesh_b200_get_scalars(lulesh_b200 *h, lulesh_b200_scalars *out);
int lulesh_b200_set_scalars(lulesh_b200 *h, const lulesh_b200_scalars *in);

/* Accessor reads after the loop / test injection.  `count` is in elements of
 * the field's type and must equal the field's length. */
int lulesh_b200_download(lulesh_b200 *h, int field, double *dst, size_t count);
int lulesh_b200_upload(lulesh_b200 *h, int field, const double *src, size_t count);
size_t lulesh_b200_field_count(lulesh_b200 *h, int field);

/* When on, run/step also store the fx..fz / xdd..zdd mirrors and ql/qq, which
 * the fused kernels otherwise keep in registers (tests, -v style dumps). */
int lulesh_b200_set_debug(lulesh_b200 *h, int on);

/* Per-kernel entry points (unit tests, ncu).  Each runs synchronously on the
 * handle's stream and returns the sticky error word.
 *   force     : K1  InitStressTerms + IntegrateStress + HourglassControl +
 *                   FBHourglassForce per element -> corner forces
 *                   (lulesh.cc:274-286, 495-560, 996-1041, 711-965)
 *   node      : K2  corner gather + CalcAcceleration + BCs + CalcVelocity +
 *                   CalcPosition (lulesh.cc:565-582, 969-986, 1139-1219)
 *   kinematics: K3  CalcKinematicsForElems + vdov + CalcMonotonicQGradients
 *                   (lulesh.cc:1505-1606, 1614-1757)
 *   material  : K4+K5 CalcMonotonicQRegion + qstop + EvalEOS + sound speed +
 *                   UpdateVolumes + Courant/hydro constraints
 *                   (lulesh.cc:1762-1921, 1994-2008, 2205-2427, 2448-2596)
 *   time_increment: K6 TimeIncrement (lulesh.cc:167-222)
 * `materialise_debug` != 0 additionally stores fx..fz / xdd..zdd mirrors. */
int lulesh_b200_kernel_time_increment(lulesh_b200 *h);
int lulesh_b200_kernel_force(lulesh_b200 *h);
int lulesh_b200_kernel_node(lulesh_b200 *h, int materialise_debug);
int lulesh_b200_kernel_kinematics(lulesh_b200 *h);
int lulesh_b200_kernel_material(lulesh_b200 *h);

/* Timing support for bench.py: runs `cycles` full cycles (no termination
 * test against stoptime beyond the device-side one) bracketed by CUDA events
 * on the handle's stream and returns the elapsed milliseconds; when
 * `per_kernel_ms` != NULL it receives LULESH_B200_NUM_KERNELS accumulated
 * main-stream intervals K6 | K1 | K2 | K3 | K45 (this mode launches the kernels
 * eagerly instead of replaying the CUDA graph; streams and overlap are the
 * same; at several ranks an interval includes the time its kernel's stream
 * waited for the exchange it depends on, see lulesh_b200_timeline). */
#define LULESH_B200_NUM_KERNELS 5
int lulesh_b200_time_cycles(lulesh_b200 *h, int32_t cycles, float *total_ms,
                            float *per_kernel_ms, int64_t *launches);

/* Per-cycle timeline (milliseconds, averaged over `cycles` cycles launched exactly like the
 * cycles of lulesh_b200_run: same streams, same overlap).  Main-stream intervals first; at
 * several ranks the three chains that run on the communication stream underneath them follow
 * (zero at one rank).  "join wait" = time the main stream sat waiting for the comm stream. */
enum lulesh_b200_timeline_slot {
   LULESH_TL_TIME_INCREMENT = 0,  /* K6 on the main stream (one rank only)              */
   LULESH_TL_K1,                  /* force kernel                                       */
   LULESH_TL_K2,                  /* node kernel (interior nodes at several ranks)      */
   LULESH_TL_NODE_JOIN_WAIT,      /* main stream waiting for the shared-node chain      */
   LULESH_TL_K3,                  /* kinematics + monotonic-Q gradients                 */
   LULESH_TL_K45_INTERIOR,        /* material kernel over elements without ghost reads  */
   LULESH_TL_MONOQ_JOIN_WAIT,     /* main stream waiting for the MonoQ exchange         */
   LULESH_TL_K45_TAIL,            /* 0: the face-layer launch runs on the comm stream   */
   LULESH_TL_CYCLE,               /* whole cycle, main stream                           */
   LULESH_TL_COMM_DT,             /* comm stream: dt candidate, min over ranks, K6      */
   LULESH_TL_COMM_NODE,           /* comm stream: shared-node gather, exchange, update  */
   LULESH_TL_COMM_MONOQ,          /* comm stream: MonoQ pack, exchange, K45 face layer  */
   LULESH_B200_TIMELINE_N
};
int lulesh_b200_timeline(lulesh_b200 *h, int32_t cycles, float *out_ms /* [LULESH_B200_TIMELINE_N] */);

/* Bytes resident in HBM for this handle / bytes uploaded by create. */
size_t lulesh_b200_device_bytes(lulesh_b200 *h);
size_t lulesh_b200_upload_bytes(lulesh_b200 *h);

/* Host-only description of the multi-rank exchanges of one rank (no GPU, no NCCL
 * needed): who shares which nodes / face layers and in which canonical order
 * the partial sums are added.  Replaces the index arithmetic of
 * lulesh-comm.cc:59-1835; exposed so it can be tested on CPU.  Arrays (int32):
 *   bnode[nb]            node ids shared with at least one other rank
 *   pack_idx[send_total] send slot -> index into the rank's own [3][nb] partials
 *   msg_rank/count/send_off/recv_off  one entry per neighbour (26 max)
 *   bsum_start[nb+1], bsum_src[2*k]   CSR of (base, stride) pairs into the halo
 *                        buffer [own 3*nb | received messages], ascending source rank
 *   mq_idx[mq_total]     MonoQ send slot -> index into delv_xi|eta|zeta ([3][allElem])
 *   face_rank/count/send_off/ghost_off  one entry per face neighbour (6 max) */
typedef struct lulesh_b200_halo_plan lulesh_b200_halo_plan;
int lulesh_b200_halo_plan_create(const lulesh_b200_host_view *view, lulesh_b200_halo_plan **out);
int lulesh_b200_halo_plan_query(const lulesh_b200_halo_plan *plan, const char *what,
                                const int32_t **data, size_t *count);
void lulesh_b200_halo_plan_destroy(lulesh_b200_halo_plan *plan);

/* "none" (1 rank), "p2p" (NVLink peer stores + flags) or "nccl" (send/recv fallback;
 * forced with LULESH_B200_HALO=nccl). */
const char *lulesh_b200_halo_mode(lulesh_b200 *h);

const char *lulesh_b200_last_error(void);

/* Replaces ~Domain (lulesh-init.cc:198-213) for the device side. */
void lulesh_b200_destroy(lulesh_b200 *h);

#ifdef __cplusplus
}
#endif
#endif /* LULESH_B200_H */
