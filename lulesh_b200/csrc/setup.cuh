// Device-side Sedov setup kernels, see setup.cu.
#pragma once

namespace lb200 {

struct SetupParams {
   int sx, sy, sz, px, py, pz, col, row, plane, numRanks;
   int G;            // longest global edge in elements
   double einit;     // energy deposited in the global origin element (lulesh-init.cc:183-185)
};

__global__ void k_setup_nodes(SetupParams S, double *x, double *y, double *z, unsigned char *nodeFlags,
                              int *cornerEll, int nn, int nn_pad, int ne_pad);
__global__ void k_setup_elems(SetupParams S, const double *x, const double *y, const double *z,
                              int *nodelist, int *lxim, int *lxip, int *letam, int *letap, int *lzetam,
                              int *lzetap, int *elemBC, double *volo, double *elemMass, double *v,
                              double *e, int ne);
__global__ void k_setup_nodal_mass(const int *cornerEll, const double *volo, double *nodalMass, int nn,
                                   int nn_pad, int ne_pad);

}  // namespace lb200
