/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the LULESH 2.0
 * Lagrange-leapfrog step and of the mesh/region/BC setup that feeds it.
 *
 * Parity status: PINNED.  tests/test_oracle_pinning.py checks this restatement
 * bit-for-bit against (a) the goldens in tests/golden/ref_goldens.json that
 * were produced by the unmodified reference (oracle/_ref, built by
 * oracle/Makefile from /root/reference) and (b) per-array dumps of the
 * reference Domain committed under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this; the product library never does.
 */
#ifndef LULESH_ORACLE_H
#define LULESH_ORACLE_H

#include "../include/lulesh_b200.h" /* field ids, constants/scalars structs */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_domain ora_domain;

/* Domain::Domain (lulesh-init.cc:16-194) generalised from (tp,nx) to a
 * (px,py,pz) rank grid of (sx,sy,sz) bricks. */
ora_domain *ora_new(int numRanks, int rank, int px, int py, int pz,
                    int sx, int sy, int sz, int numReg, int balance, int cost);
void ora_free(ora_domain *d);
void ora_use_reference_dt0(ora_domain *d);              /* lulesh-init.cc:192 as written (per-rank) */

double *ora_real(ora_domain *d, int field);          /* field ids of lulesh_b200.h */
size_t  ora_real_count(ora_domain *d, int field);
int    *ora_int(ora_domain *d, const char *name, int *count);
int    *ora_region_list(ora_domain *d, int r, int *count);
lulesh_b200_scalars   *ora_scalars(ora_domain *d);
lulesh_b200_constants *ora_constants(ora_domain *d);

/* the step, phase by phase (return 0 / VolumeError -1 / QStopError -2) */
double ora_dt_candidate(ora_domain *d);               /* lulesh.cc:176-183 */
void   ora_time_increment(ora_domain *d, double newdt_reduced); /* lulesh.cc:167-222 */
int    ora_calc_force(ora_domain *d);                 /* lulesh.cc:1104-1135 (no comm) */
void   ora_node_update(ora_domain *d);                /* lulesh.cc:1139-1219 */
int    ora_kinematics(ora_domain *d);                 /* lulesh.cc:1573-1609 */
void   ora_monoq_gradients(ora_domain *d);            /* lulesh.cc:1614-1757 */
int    ora_monoq_regions(ora_domain *d);              /* lulesh.cc:1926-1941,1994-2008 */
int    ora_material(ora_domain *d);                   /* lulesh.cc:2329-2427 */
void   ora_time_constraints(ora_domain *d);           /* lulesh.cc:2577-2596 */
int    ora_step(ora_domain *d);                       /* one TimeIncrement + LagrangeLeapFrog */
int    ora_run(ora_domain *d, int max_cycles);        /* lulesh.cc:2745-2757 */

/* In-process emulation of a (px,py,pz) rank grid: one ora_domain per rank,
 * halo exchanges by memcpy with the reference's semantics
 * (lulesh-comm.cc: CommSBN add, CommSyncPosVel overwrite, CommMonoQ ghost
 * copy; dt = min over ranks, lulesh.cc:186).  Oracle for 2/4/8-GPU runs. */
typedef struct ora_multi ora_multi;
ora_multi  *ora_multi_new(int px, int py, int pz, int sx, int sy, int sz,
                          int numReg, int balance, int cost);
void        ora_multi_free(ora_multi *m);
ora_domain *ora_multi_rank(ora_multi *m, int rank);
int         ora_multi_step(ora_multi *m);
int         ora_multi_run(ora_multi *m, int max_cycles);

/* lulesh-util.cc:197-218 symmetry triple on plane 0 (n = square edge) */
void ora_symmetry(ora_domain *d, int n, double out3[3]);

#ifdef __cplusplus
}
#endif
#endif
