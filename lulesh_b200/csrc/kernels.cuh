// Device kernels of the B200-native Lagrange-leapfrog step (sm_100a, FP64).
//
// One cycle = K6 time_increment -> K1 force_elem -> K2 node_update ->
//             K3 kinematics_grad -> K45 material (monoq + EOS + dt minima).
// Several ranks: K6 with its min-reduction, the shared-node chain (k_node_boundary_*,
// k_peer_pack/wait) and the MonoQ exchange + the face-layer part of K45 run on a second,
// high-priority stream underneath K1, K2 and the interior part of K45 (api.cu: enqueue_cycle).
// All index arrays (nodelist, face neighbours, region work list, corner gather
// table) are READ FROM MEMORY; nothing is synthesised from (i,j,k).
//
// Reference map (file:line are /root/reference):
//   K6  lulesh.cc:167-222
//   K1  lulesh.cc:274-286, 495-560, 1082-1091, 996-1041, 711-965
//   K2  lulesh.cc:565-582, 969-986, 1139-1219
//   K3  lulesh.cc:1505-1606, 1614-1757
//   K45 lulesh.cc:1762-1921, 1994-2008, 2329-2427, 2448-2596
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lulesh_b200.h"

namespace lb200 {

// elemBC bit layout (lulesh.h:59-87)
enum : int {
   XI_M = 0x00007, XI_M_SYMM = 0x00001, XI_M_FREE = 0x00002,
   XI_P = 0x00038, XI_P_SYMM = 0x00008, XI_P_FREE = 0x00010,
   ETA_M = 0x001c0, ETA_M_SYMM = 0x00040, ETA_M_FREE = 0x00080,
   ETA_P = 0x00e00, ETA_P_SYMM = 0x00200, ETA_P_FREE = 0x00400,
   ZETA_M = 0x07000, ZETA_M_SYMM = 0x01000, ZETA_M_FREE = 0x02000,
   ZETA_P = 0x38000, ZETA_P_SYMM = 0x08000, ZETA_P_FREE = 0x10000,
   XI_M_COMM = 0x00004, XI_P_COMM = 0x00020, ETA_M_COMM = 0x00100, ETA_P_COMM = 0x00800,
   ZETA_M_COMM = 0x04000, ZETA_P_COMM = 0x20000
};

// per-node flag byte built at create() from symmX/Y/Z and the halo layout
enum : unsigned { NODE_SYMM_X = 1u, NODE_SYMM_Y = 2u, NODE_SYMM_Z = 4u, NODE_COMM = 8u };

// Device-resident control block (lulesh.h:558-567 + run control).  The two dt
// minima are kept as the bit patterns of positive doubles so that the block
// minima can be merged with a 64-bit integer atomicMin (order independent,
// hence deterministic; no FP atomics anywhere).
struct Ctl {
   unsigned long long dtcourant_bits, dthydro_bits;
   double dtfixed, time, deltatime, deltatimemultlb, deltatimemultub, dtmax, stoptime;
   double gnewdt;        // this rank's candidate, input/output of the min-allreduce
   int cycle, max_cycles;
   int done;             // 1: time >= stoptime or cycle >= max_cycles -> kernels no-op
   int error;            // sticky, first one wins: 0 / VolumeError -1 / QStopError -2 / infrastructure
   int pending_error;    // K1's abort test, promoted to `error` by K2 if the cycle is live
   int zero;             // always 0; loaded by the EOS loop so that its repetitions cannot be folded
};

struct KParams {
   int ne, nn, allElem;
   int ne_pad, nn_pad;            // plane strides of the SoA corner buffers
   Ctl *ctl;
   // node-centred
   double *x, *y, *z, *xd, *yd, *zd;
   const double *nodalMass;
   const unsigned char *nodeFlags;
   const int *cornerEll;          // [8][nn_pad] -> index into one force plane set, -1 = none
   double *dbg_f, *dbg_a;         // [3][nn] debug mirrors of fx.. / xdd.. (may be null)
   // element-centred
   const int *nodelist;           // [ne][8]
   const int *lxim, *lxip, *letam, *letap, *lzetam, *lzetap, *elemBC;
   double *e, *p, *q, *ql, *qq, *v, *ss, *vnew, *delv, *vdov, *arealg;
   const double *volo, *elemMass;
   double *delv_xi, *delv_eta, *delv_zeta;   // [allElem]
   double *delx_xi, *delx_eta, *delx_zeta;   // [ne]
   double *fcorner;               // [3][8][ne_pad] per-corner forces, SoA by (axis,corner)
   // region work list: blocks of MAT_THREADS entries, one region per block
   const int *workElem;           // [numWorkBlocks*MAT_THREADS], -1 = padding
   const int *workBlockRep;       // [numWorkBlocks] EOS repetition count of the block's region
   int numWorkBlocks;
   int numWorkBlocksInterior;     // blocks [0, interior) never read ghost slots; the rest wait for the MonoQ exchange
   // multi-rank boundary-node machinery (null/0 at numRanks==1)
   int nbnode;                    // boundary (shared) nodes on this rank
   const int *bnode;              // [nbnode] node id
   const int *bsum_start;         // [nbnode+1] CSR over contributions, canonical rank order
   const int *bsum_src;           // index into fhalo planes (own slot or recv slot)
   double *fhalo;                 // [3][fhalo_stride]: own partials [0,nbnode) then recv slots
   int fhalo_stride;
   const struct PeerCounters *peer_cnt;   // non-null in peer-to-peer mode: fhalo is double-buffered by sequence parity
   int unit_rho0;                 // refdens == 1.0: k_material (no division by rho0) instead of k_material_rho0
   lulesh_b200_constants c;
};

#ifndef LB_K1_THREADS
#define LB_K1_THREADS 128
#endif
#ifndef LB_K3_THREADS
#define LB_K3_THREADS 128
#endif
constexpr int K1_THREADS = LB_K1_THREADS;
#ifndef LB_K1_BPS
#define LB_K1_BPS 2
#endif
#ifndef LB_K3_BPS
#define LB_K3_BPS 3
#endif
constexpr int K1_BLOCKS_PER_SM = LB_K1_BPS;
constexpr int K2_THREADS = 256;
constexpr int K3_THREADS = LB_K3_THREADS;
constexpr int K3_BLOCKS_PER_SM = LB_K3_BPS;
constexpr int K1_SLOTS = 54;   // cp.async staging tile: 48 node values + 6 scalars
constexpr int K1_SMEM_BYTES = K1_SLOTS * K1_THREADS * 8;
constexpr int K3_SMEM_BYTES = 50 * K3_THREADS * 8;   // 48 node values + volo, v
constexpr int MAT_THREADS = 128;
constexpr int MAT_BLOCKS_PER_SM = 8;

__global__ void k_time_increment(Ctl *ctl, int phase);
__global__ void k_force(const KParams P);
__global__ void k_node(const KParams P, int storeDebug);
__global__ void k_node_boundary_gather(const KParams P);
__global__ void k_node_boundary_update(const KParams P, int storeDebug);
__global__ void k_kinematics(const KParams P);
__global__ void k_material(const KParams P, int storeQ, int firstBlock);
__global__ void k_material_rho0(const KParams P, int storeQ, int firstBlock);
__global__ void k_gather_index(double *dst, const double *src, const int *idx, int n);

// ---- peer-to-peer halo exchange over NVLink (stores into the neighbour's HBM + flags)
constexpr int PEER_MAX_RANKS = 64;
constexpr int PEER_FLAG_NODE = 0;     // flags[0..25]: node-halo messages, receiver's message index
constexpr int PEER_FLAG_FACE = 32;    // flags[32..37]: MonoQ face messages, receiver's face index
constexpr int PEER_NUM_FLAGS = 64;

struct DtSlot { double val; unsigned long long seq; };

struct PeerMsg {                       // one outgoing message
   double *dst;                        // peer memory: where field 0 of this message starts
   unsigned long long *flag;           // peer memory: the receiver's flag for this message
   int send_off, count;                // slot range [send_off, send_off + 3*count) of the pack list
   int field_stride;                   // distance between the three fields at the destination
   int parity_stride;                  // destination offset of the odd-sequence buffer (0: single buffer)
};

struct PeerCounters {                  // device-resident sequence numbers, one pair per exchange kind
   unsigned long long node_seq, node_expect, face_seq, face_expect, dt_seq, dt_expect;
   unsigned int node_done, face_done;  // "last block" detection of the pack kernels
};

__global__ void k_peer_pack(const double *src, int src_parity_stride, const int *idx,
                            const unsigned char *slot_msg, int n, const PeerMsg *msgs, int nmsg,
                            unsigned int *done_counter, unsigned long long *seq);
__global__ void k_peer_wait(const unsigned long long *flags, int first, int n,
                            unsigned long long *expect, Ctl *ctl);
__global__ void k_peer_dt_post(Ctl *ctl, DtSlot *const *peer_slots, int me, int nranks,
                               unsigned long long *dt_seq);
__global__ void k_peer_dt_wait(Ctl *ctl, const DtSlot *my_slots, int nranks,
                               unsigned long long *dt_expect);
__global__ void k_boundary_mass(const KParams P, double *nodalMass);

}  // namespace lb200
