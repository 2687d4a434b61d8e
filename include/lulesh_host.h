/*
 * lulesh_host.h -- C access to the host-side Domain (lulesh_b200/csrc/host),
 * the C++ mirror of the reference's `class Domain` (lulesh.h:148-595) and of its
 * setup in lulesh-init.cc.  This is what a test or a foreign-language host uses
 * to obtain the arrays that lulesh_b200_create() consumes; the C++ driver uses
 * the class directly.
 */
#ifndef LULESH_HOST_H
#define LULESH_HOST_H

#include "lulesh_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lulesh_host_domain lulesh_host_domain;

/* Domain::Domain (lulesh-init.cc:16-194) for rank `rank` of a (px,py,pz) grid of
 * (sx,sy,sz)-element bricks.  The reference's constructor is the special case
 * px=py=pz=tp, sx=sy=sz=nx.  Returns NULL on invalid arguments. */
lulesh_host_domain *lulesh_host_domain_new(int numRanks, int rank, int px, int py, int pz,
                                           int sx, int sy, int sz, int numReg, int balance,
                                           int cost);
void lulesh_host_domain_free(lulesh_host_domain *d);

/* Fills `out` with pointers into the Domain (valid until it is freed). */
void lulesh_host_domain_view(lulesh_host_domain *d, lulesh_b200_host_view *out);

/* Host storage of a real-valued field (ids of lulesh_b200.h); NULL if the host
 * Domain does not keep that field. */
double *lulesh_host_domain_field(lulesh_host_domain *d, int field, size_t *count);
const int32_t *lulesh_host_domain_ints(lulesh_host_domain *d, const char *name, size_t *count);
lulesh_b200_scalars *lulesh_host_domain_scalars(lulesh_host_domain *d);

/* The `-v` dump of one rank's Domain (DumpToVisit / DumpDomainToVisit,
 * lulesh-viz.cc:56-258): mesh, connectivity, region numbers, zone fields e p v q,
 * node fields speed xd yd zd -- as a binary legacy-VTK unstructured grid instead of
 * Silo.  The Domain's host arrays are written as they are (the driver downloads the
 * device state into them first).  Returns 0 on success, -1 on I/O errors. */
int lulesh_host_domain_write_vtk(lulesh_host_domain *d, int rank, const char *path);

/* InitMeshDecomp (lulesh-init.cc:676-738) generalised: picks (px,py,pz) for
 * numRanks in {1,2,4,8,27,...}: cubes as the reference, 2 -> 1x1x2, 4 -> 1x2x2.
 * Returns 0 on success. */
int lulesh_host_decompose(int numRanks, int *px, int *py, int *pz);

/* The drop-in driver's main() (lulesh.cc:2650-2792), callable from tests. */
int lulesh_host_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif
