/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of LULESH 2.0's Lagrange
 * leapfrog step and Domain setup (see lulesh_oracle.h for the contract and the
 * pinning status).  Plain C11, compiled with -ffp-contract=off and without
 * -march so that, like the reference's Makefile:24 build, no FMA is formed.
 *
 * Every floating-point expression keeps the reference's association order so
 * the restatement is bit-identical to the reference's *threaded* path (corner
 * scratch + gather in nodeElemCornerList order, lulesh.cc:540-547, 565-582,
 * 901-931, 969-986), which is the canonical summation order of this project.
 * The code is organised around small tables (hexahedron faces, volume-
 * derivative stencils, hourglass base vectors) instead of the reference's
 * unrolled scalar code, and works on per-element scalars where the reference
 * streams through region-sized temporaries; arithmetic per element is the same.
 */
#define _GNU_SOURCE
#include "lulesh_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* elemBC bit layout, lulesh.h:59-87 */
enum {
   XI_M = 0x00007, XI_M_SYMM = 0x00001, XI_M_FREE = 0x00002, XI_M_COMM = 0x00004,
   XI_P = 0x00038, XI_P_SYMM = 0x00008, XI_P_FREE = 0x00010, XI_P_COMM = 0x00020,
   ETA_M = 0x001c0, ETA_M_SYMM = 0x00040, ETA_M_FREE = 0x00080, ETA_M_COMM = 0x00100,
   ETA_P = 0x00e00, ETA_P_SYMM = 0x00200, ETA_P_FREE = 0x00400, ETA_P_COMM = 0x00800,
   ZETA_M = 0x07000, ZETA_M_SYMM = 0x01000, ZETA_M_FREE = 0x02000, ZETA_M_COMM = 0x04000,
   ZETA_P = 0x38000, ZETA_P_SYMM = 0x08000, ZETA_P_FREE = 0x10000, ZETA_P_COMM = 0x20000
};

struct ora_domain {
   int sx, sy, sz, px, py, pz, col, row, plane, numRanks, rank;
   int numElem, numNode, allElem;
   int cMin, cMax, rMin, rMax, pMin, pMax; /* neighbour exists? (lulesh-init.cc:350-355) */
   double *f[LULESH_F_COUNT];
   int *nodelist, *lxim, *lxip, *letam, *letap, *lzetam, *lzetap, *elemBC;
   int *symmX, *symmY, *symmZ, nsymmX, nsymmY, nsymmZ;
   int numReg, cost, *regElemSize, *regNumList, **regElemlist;
   int *nodeElemStart, *nodeElemCornerList;
   double *cfx, *cfy, *cfz; /* per-corner scratch, [8*numElem] (lulesh.cc:515-517) */
   lulesh_b200_constants c;
   lulesh_b200_scalars s;
};

#define F(d, id) ((d)->f[LULESH_F_##id])

/* ------------------------------------------------------------------------ */
/* element geometry helpers                                                  */
/* ------------------------------------------------------------------------ */

static void gather8(const double *a, const int *nl, double out[8])
{
   for (int c = 0; c < 8; ++c) out[c] = a[nl[c]];
}

/* lulesh.cc:1337-1338: TRIPLE_PRODUCT with its nine positional arguments */
static inline double triple(double a1, double a2, double a3, double b1, double b2,
                            double b3, double c1, double c2, double c3)
{
   return a1 * (b2 * c3 - b3 * c2) + b1 * (a3 * c2 - a2 * c3) + c1 * (a2 * b3 - a3 * b2);
}

/* CalcElemVolume, lulesh.cc:1274-1356 */
static double elem_volume(const double x[8], const double y[8], const double z[8])
{
#define D(a, i, j) (a[i] - a[j])
   double v =
      triple(D(x,3,1) + D(x,7,2), D(x,6,3), D(x,2,0),
             D(y,3,1) + D(y,7,2), D(y,6,3), D(y,2,0),
             D(z,3,1) + D(z,7,2), D(z,6,3), D(z,2,0)) +
      triple(D(x,4,3) + D(x,5,7), D(x,6,4), D(x,7,0),
             D(y,4,3) + D(y,5,7), D(y,6,4), D(y,7,0),
             D(z,4,3) + D(z,5,7), D(z,6,4), D(z,7,0)) +
      triple(D(x,1,4) + D(x,2,5), D(x,6,1), D(x,5,0),
             D(y,1,4) + D(y,2,5), D(y,6,1), D(y,5,0),
             D(z,1,4) + D(z,2,5), D(z,6,1), D(z,5,0));
#undef D
   return v * (1.0 / 12.0);
}

/* CalcElemShapeFunctionDerivatives, lulesh.cc:291-377.  fj[a][k]: derivative
 * of coordinate a w.r.t. (xi, eta, zeta); cj = cofactors. */
static void shape_derivs(const double x[8], const double y[8], const double z[8],
                         double b[3][8], double *vol)
{
   const double *co[3] = {x, y, z};
   double fj[3][3], cj[3][3];
   for (int a = 0; a < 3; ++a) {
      const double *q = co[a];
      double d60 = q[6] - q[0], d53 = q[5] - q[3], d71 = q[7] - q[1], d42 = q[4] - q[2];
      fj[a][0] = .125 * (d60 + d53 - d71 - d42);
      fj[a][1] = .125 * (d60 - d53 + d71 - d42);
      fj[a][2] = .125 * (d60 + d53 + d71 + d42);
   }
   /* cofactors, lulesh.cc:332-342.  With (u,w) the other two coordinates in
    * cyclic order all nine entries follow one pattern; the reference's
    * "-(p) + q" forms are evaluated as q - p, which is identical in IEEE. */
   for (int a = 0; a < 3; ++a) {
      const int u = (a + 1) % 3, w = (a + 2) % 3;
      cj[a][0] = (fj[u][1] * fj[w][2]) - (fj[w][1] * fj[u][2]);
      cj[a][1] = (fj[w][0] * fj[u][2]) - (fj[u][0] * fj[w][2]);
      cj[a][2] = (fj[u][0] * fj[w][1]) - (fj[w][0] * fj[u][1]);
   }
   for (int a = 0; a < 3; ++a) {
      double c0 = cj[a][0], c1 = cj[a][1], c2 = cj[a][2];
      b[a][0] = -c0 - c1 - c2;
      b[a][1] =  c0 - c1 - c2;
      b[a][2] =  c0 + c1 - c2;
      b[a][3] = -c0 + c1 - c2;
      b[a][4] = -b[a][2];
      b[a][5] = -b[a][3];
      b[a][6] = -b[a][0];
      b[a][7] = -b[a][1];
   }
   *vol = 8. * (fj[0][1] * cj[0][1] + fj[1][1] * cj[1][1] + fj[2][1] * cj[2][1]);
}

/* the six faces in the reference's visiting order, lulesh.cc:432-473 */
static const int k_face_nodes[6][4] = {
   {0, 1, 2, 3}, {0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {4, 7, 6, 5}};

/* CalcElemNodeNormals + SumElemFaceNormal, lulesh.cc:382-474 */
static void node_normals(const double x[8], const double y[8], const double z[8],
                         double pf[3][8])
{
   const double *co[3] = {x, y, z};
   for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 8; ++c) pf[a][c] = 0.0;
   for (int f = 0; f < 6; ++f) {
      const int *n = k_face_nodes[f];
      double b0[3], b1[3], area[3];
      for (int a = 0; a < 3; ++a) {
         const double *q = co[a];
         b0[a] = 0.5 * (q[n[3]] + q[n[2]] - q[n[1]] - q[n[0]]);
         b1[a] = 0.5 * (q[n[2]] + q[n[1]] - q[n[3]] - q[n[0]]);
      }
      area[0] = 0.25 * (b0[1] * b1[2] - b0[2] * b1[1]);
      area[1] = 0.25 * (b0[2] * b1[0] - b0[0] * b1[2]);
      area[2] = 0.25 * (b0[0] * b1[1] - b0[1] * b1[0]);
      for (int a = 0; a < 3; ++a)
         for (int k = 0; k < 4; ++k) pf[a][n[k]] += area[a];
   }
}

/* VoluDer stencil rows, lulesh.cc:631-662: output corner, then the six corners
 * passed as (0..5) */
static const int k_voluder[8][7] = {
   {0, 1, 2, 3, 4, 5, 7}, {3, 0, 1, 2, 7, 4, 6}, {2, 3, 0, 1, 6, 7, 5}, {1, 2, 3, 0, 5, 6, 4},
   {4, 7, 6, 5, 0, 3, 1}, {5, 4, 7, 6, 1, 0, 2}, {6, 5, 4, 7, 2, 1, 3}, {7, 6, 5, 4, 3, 2, 0}};

/* one component of VoluDer (lulesh.cc:602-605) for coordinate pair (p,q) */
static inline double voluder_term(const double p[6], const double q[6])
{
   return (p[1] + p[2]) * (q[0] + q[1]) - (p[0] + p[1]) * (q[1] + q[2]) +
          (p[0] + p[4]) * (q[3] + q[4]) - (p[3] + p[4]) * (q[0] + q[4]) -
          (p[2] + p[5]) * (q[3] + q[5]) + (p[3] + p[5]) * (q[2] + q[5]);
}

/* CalcElemVolumeDerivative, lulesh.cc:592-663.  dvdy and dvdz are the exact
 * negatives of the dvdx pattern applied to (x,z) and (y,x): negating every
 * term of a sum negates the rounded result exactly. */
static void volume_derivs(const double x[8], const double y[8], const double z[8],
                          double dv[3][8])
{
   const double twelfth = 1.0 / 12.0;
   for (int r = 0; r < 8; ++r) {
      const int *s = k_voluder[r];
      double xs[6], ys[6], zs[6];
      for (int k = 0; k < 6; ++k) { xs[k] = x[s[k + 1]]; ys[k] = y[s[k + 1]]; zs[k] = z[s[k + 1]]; }
      dv[0][s[0]] = voluder_term(ys, zs) * twelfth;
      dv[1][s[0]] = -voluder_term(xs, zs) * twelfth;
      dv[2][s[0]] = -voluder_term(ys, xs) * twelfth;
   }
}

/* hourglass base vectors, lulesh.cc:745-776 */
static const double k_gamma[4][8] = {
   { 1,  1, -1, -1, -1, -1,  1,  1},
   { 1, -1, -1,  1, -1,  1,  1, -1},
   { 1, -1,  1, -1,  1, -1,  1, -1},
   {-1,  1, -1,  1,  1, -1,  1, -1}};

/* AreaFace, lulesh.cc:1371-1390 */
static inline double area_face(const double x[8], const double y[8], const double z[8],
                               int n0, int n1, int n2, int n3)
{
   double fx = (x[n2] - x[n0]) - (x[n3] - x[n1]);
   double fy = (y[n2] - y[n0]) - (y[n3] - y[n1]);
   double fz = (z[n2] - z[n0]) - (z[n3] - z[n1]);
   double gx = (x[n2] - x[n0]) + (x[n3] - x[n1]);
   double gy = (y[n2] - y[n0]) + (y[n3] - y[n1]);
   double gz = (z[n2] - z[n0]) + (z[n3] - z[n1]);
   return (fx * fx + fy * fy + fz * fz) * (gx * gx + gy * gy + gz * gz) -
          (fx * gx + fy * gy + fz * gz) * (fx * gx + fy * gy + fz * gz);
}

/* CalcElemCharacteristicLength, lulesh.cc:1395-1435 */
static double char_length(const double x[8], const double y[8], const double z[8], double volume)
{
   static const int faces[6][4] = {
      {0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}};
   double m = 0.0;
   for (int f = 0; f < 6; ++f) {
      double a = area_face(x, y, z, faces[f][0], faces[f][1], faces[f][2], faces[f][3]);
      if (m < a) m = a; /* std::max(a, charLength) */
   }
   return 4.0 * volume / sqrt(m);
}

/* ------------------------------------------------------------------------ */
/* A1 TimeIncrement                                                          */
/* ------------------------------------------------------------------------ */

double ora_dt_candidate(ora_domain *d)
{
   double g = 1.0e+20;
   if (d->s.dtcourant < g) g = d->s.dtcourant / 2.0;
   if (d->s.dthydro < g) g = d->s.dthydro * 2.0 / 3.0;
   return g;
}

void ora_time_increment(ora_domain *d, double newdt)
{
   lulesh_b200_scalars *s = &d->s;
   double targetdt = s->stoptime - s->time;
   if (s->dtfixed <= 0.0 && s->cycle != 0) {
      double olddt = s->deltatime;
      double ratio = newdt / olddt;
      if (ratio >= 1.0) {
         if (ratio < s->deltatimemultlb) newdt = olddt;
         else if (ratio > s->deltatimemultub) newdt = olddt * s->deltatimemultub;
      }
      if (newdt > s->dtmax) newdt = s->dtmax;
      s->deltatime = newdt;
   }
   if (targetdt > s->deltatime && targetdt < 4.0 * s->deltatime / 3.0)
      targetdt = 2.0 * s->deltatime / 3.0;
   if (targetdt < s->deltatime) s->deltatime = targetdt;
   s->time += s->deltatime;
   ++s->cycle;
}

/* ------------------------------------------------------------------------ */
/* A2-A6 nodal forces (threaded-path order)                                  */
/* ------------------------------------------------------------------------ */

static void gather_corners(ora_domain *d, int accumulate)
{
   double *fx = F(d, FX), *fy = F(d, FY), *fz = F(d, FZ);
#pragma omp parallel for
   for (int n = 0; n < d->numNode; ++n) {
      double tx = 0.0, ty = 0.0, tz = 0.0;
      for (int k = d->nodeElemStart[n]; k < d->nodeElemStart[n + 1]; ++k) {
         int ci = d->nodeElemCornerList[k];
         tx += d->cfx[ci]; ty += d->cfy[ci]; tz += d->cfz[ci];
      }
      if (accumulate) { fx[n] += tx; fy[n] += ty; fz[n] += tz; }
      else            { fx[n]  = tx; fy[n]  = ty; fz[n]  = tz; }
   }
}

int ora_calc_force(ora_domain *d)
{
   const int ne = d->numElem;
   const double *X = F(d, X), *Y = F(d, Y), *Z = F(d, Z);
   const double *XD = F(d, XD), *YD = F(d, YD), *ZD = F(d, ZD);
   int err = 0;

   memset(F(d, FX), 0, sizeof(double) * d->numNode);
   memset(F(d, FY), 0, sizeof(double) * d->numNode);
   memset(F(d, FZ), 0, sizeof(double) * d->numNode);
   if (ne == 0) return 0;

   /* IntegrateStressForElems, lulesh.cc:495-587 with sig = -p-q (274-286) */
#pragma omp parallel for reduction(| : err)
   for (int k = 0; k < ne; ++k) {
      const int *nl = &d->nodelist[8 * k];
      double x[8], y[8], z[8], B[3][8], determ;
      gather8(X, nl, x); gather8(Y, nl, y); gather8(Z, nl, z);
      shape_derivs(x, y, z, B, &determ); /* only determ survives (539) */
      node_normals(x, y, z, B);
      double sig = -F(d, P)[k] - F(d, Q)[k];
      for (int c = 0; c < 8; ++c) {
         d->cfx[8 * k + c] = -(sig * B[0][c]);
         d->cfy[8 * k + c] = -(sig * B[1][c]);
         d->cfz[8 * k + c] = -(sig * B[2][c]);
      }
      if (determ <= 0.0) err |= 1; /* lulesh.cc:1082-1091 */
   }
   gather_corners(d, 0);
   if (err) return LULESH_B200_VOLUME_ERROR;

   /* CalcHourglassControlForElems + CalcFBHourglassForceForElems,
    * lulesh.cc:996-1057, 711-991 */
   const double hourg = d->c.hgcoef;
#pragma omp parallel for reduction(| : err)
   for (int k = 0; k < ne; ++k) {
      const int *nl = &d->nodelist[8 * k];
      double x[8], y[8], z[8], dv[3][8];
      gather8(X, nl, x); gather8(Y, nl, y); gather8(Z, nl, z);
      volume_derivs(x, y, z, dv);
      double determ = F(d, VOLO)[k] * F(d, V)[k];
      if (F(d, V)[k] <= 0.0) err |= 1; /* lulesh.cc:1034 */
      if (!(hourg > 0.0)) {
         for (int c = 0; c < 8; ++c) d->cfx[8*k+c] = d->cfy[8*k+c] = d->cfz[8*k+c] = 0.0;
         continue;
      }
      double volinv = 1.0 / determ;
      double hourgam[8][4];
      for (int m = 0; m < 4; ++m) {
         const double *g = k_gamma[m];
         double hx = x[0] * g[0], hy = y[0] * g[0], hz = z[0] * g[0];
         for (int c = 1; c < 8; ++c) { hx += x[c] * g[c]; hy += y[c] * g[c]; hz += z[c] * g[c]; }
         for (int c = 0; c < 8; ++c)
            hourgam[c][m] = g[c] - volinv * (dv[0][c] * hx + dv[1][c] * hy + dv[2][c] * hz);
      }
      double ss1 = F(d, SS)[k], mass1 = F(d, ELEMMASS)[k], volume13 = cbrt(determ);
      double coefficient = -hourg * 0.01 * ss1 * mass1 / volume13;
      double vel[3][8];
      gather8(XD, nl, vel[0]); gather8(YD, nl, vel[1]); gather8(ZD, nl, vel[2]);
      double *out[3] = {&d->cfx[8 * k], &d->cfy[8 * k], &d->cfz[8 * k]};
      for (int a = 0; a < 3; ++a) { /* CalcElemFBHourglassForce, lulesh.cc:668-706 */
         double h[4];
         for (int m = 0; m < 4; ++m) {
            double s = hourgam[0][m] * vel[a][0];
            for (int c = 1; c < 8; ++c) s += hourgam[c][m] * vel[a][c];
            h[m] = s;
         }
         for (int c = 0; c < 8; ++c)
            out[a][c] = coefficient * (hourgam[c][0] * h[0] + hourgam[c][1] * h[1] +
                                       hourgam[c][2] * h[2] + hourgam[c][3] * h[3]);
      }
   }
   if (err) return LULESH_B200_VOLUME_ERROR;
   if (hourg > 0.0) gather_corners(d, 1);
   return 0;
}

/* ------------------------------------------------------------------------ */
/* A7-A10 nodal update                                                       */
/* ------------------------------------------------------------------------ */

void ora_node_update(ora_domain *d)
{
   const int nn = d->numNode;
   const double dt = d->s.deltatime, u_cut = d->c.u_cut;
   double *xdd = F(d, XDD), *ydd = F(d, YDD), *zdd = F(d, ZDD);
#pragma omp parallel for
   for (int i = 0; i < nn; ++i) { /* lulesh.cc:1143-1148 */
      double m = F(d, NODALMASS)[i];
      xdd[i] = F(d, FX)[i] / m; ydd[i] = F(d, FY)[i] / m; zdd[i] = F(d, FZ)[i] / m;
   }
   for (int i = 0; i < d->nsymmX; ++i) xdd[d->symmX[i]] = 0.0; /* lulesh.cc:1159-1178 */
   for (int i = 0; i < d->nsymmY; ++i) ydd[d->symmY[i]] = 0.0;
   for (int i = 0; i < d->nsymmZ; ++i) zdd[d->symmZ[i]] = 0.0;
   double *vel[3] = {F(d, XD), F(d, YD), F(d, ZD)};
   double *acc[3] = {xdd, ydd, zdd};
   double *pos[3] = {F(d, X), F(d, Y), F(d, Z)};
#pragma omp parallel for
   for (int i = 0; i < nn; ++i)
      for (int a = 0; a < 3; ++a) {
         double t = vel[a][i] + acc[a][i] * dt; /* lulesh.cc:1193-1203 */
         if (fabs(t) < u_cut) t = 0.0;
         vel[a][i] = t;
         pos[a][i] += t * dt;                   /* lulesh.cc:1215-1217 */
      }
}

/* ------------------------------------------------------------------------ */
/* A11 kinematics, A12 monotonic-Q gradients                                 */
/* ------------------------------------------------------------------------ */

int ora_kinematics(ora_domain *d)
{
   const int ne = d->numElem;
   const double dt = d->s.deltatime;
   int err = 0;
#pragma omp parallel for reduction(| : err)
   for (int k = 0; k < ne; ++k) {
      const int *nl = &d->nodelist[8 * k];
      double x[8], y[8], z[8], xd[8], yd[8], zd[8], B[3][8], detJ;
      gather8(F(d, X), nl, x); gather8(F(d, Y), nl, y); gather8(F(d, Z), nl, z);
      double volume = elem_volume(x, y, z);
      double rel = volume / F(d, VOLO)[k];
      F(d, VNEW)[k] = rel;
      F(d, DELV)[k] = rel - F(d, V)[k];
      F(d, AREALG)[k] = char_length(x, y, z, volume);
      gather8(F(d, XD), nl, xd); gather8(F(d, YD), nl, yd); gather8(F(d, ZD), nl, zd);
      double dt2 = 0.5 * dt;
      for (int c = 0; c < 8; ++c) { x[c] -= dt2 * xd[c]; y[c] -= dt2 * yd[c]; z[c] -= dt2 * zd[c]; }
      shape_derivs(x, y, z, B, &detJ);
      /* CalcElemVelocityGradient diagonal, lulesh.cc:1447-1466 */
      double inv = 1.0 / detJ;
      const double *vv[3] = {xd, yd, zd};
      double D[3];
      for (int a = 0; a < 3; ++a)
         D[a] = inv * (B[a][0] * (vv[a][0] - vv[a][6]) + B[a][1] * (vv[a][1] - vv[a][7]) +
                       B[a][2] * (vv[a][2] - vv[a][4]) + B[a][3] * (vv[a][3] - vv[a][5]));
      F(d, VDOV)[k] = D[0] + D[1] + D[2]; /* lulesh.cc:1588-1592 */
      if (rel <= 0.0) err |= 1;           /* lulesh.cc:1598 */
   }
   return err ? LULESH_B200_VOLUME_ERROR : 0;
}

static inline double s4(const double *q, int a, int b, int c, int e)
{
   return q[a] + q[b] + q[c] + q[e];
}

void ora_monoq_gradients(ora_domain *d)
{
   const int ne = d->numElem;
   const double ptiny = 1.e-36;
#pragma omp parallel for
   for (int i = 0; i < ne; ++i) {
      const int *nl = &d->nodelist[8 * i];
      double p[3][8], u[3][8];
      gather8(F(d, X), nl, p[0]); gather8(F(d, Y), nl, p[1]); gather8(F(d, Z), nl, p[2]);
      gather8(F(d, XD), nl, u[0]); gather8(F(d, YD), nl, u[1]); gather8(F(d, ZD), nl, u[2]);
      double vol = F(d, VOLO)[i] * F(d, VNEW)[i];
      double norm = 1.0 / (vol + ptiny);
      double dj[3], di[3], dk[3], vj[3], vi[3], vk[3];
      for (int a = 0; a < 3; ++a) { /* lulesh.cc:1691-1701, 1715-1753 */
         dj[a] = -0.25 * (s4(p[a], 0, 1, 5, 4) - s4(p[a], 3, 2, 6, 7));
         di[a] =  0.25 * (s4(p[a], 1, 2, 6, 5) - s4(p[a], 0, 3, 7, 4));
         dk[a] =  0.25 * (s4(p[a], 4, 5, 6, 7) - s4(p[a], 0, 1, 2, 3));
         vk[a] =  0.25 * (s4(u[a], 4, 5, 6, 7) - s4(u[a], 0, 1, 2, 3));
         vi[a] =  0.25 * (s4(u[a], 1, 2, 6, 5) - s4(u[a], 0, 3, 7, 4));
         vj[a] = -0.25 * (s4(u[a], 0, 1, 5, 4) - s4(u[a], 3, 2, 6, 7));
      }
      const double *L[3] = {di, dj, dk}, *R[3] = {dj, dk, di}, *V[3] = {vk, vi, vj};
      double *delx[3] = {F(d, DELX_ZETA), F(d, DELX_XI), F(d, DELX_ETA)};
      double *delv[3] = {F(d, DELV_ZETA), F(d, DELV_XI), F(d, DELV_ETA)};
      for (int t = 0; t < 3; ++t) { /* zeta = i x j, xi = j x k, eta = k x i */
         const double *l = L[t], *r = R[t];
         double ax = l[1] * r[2] - l[2] * r[1];
         double ay = l[2] * r[0] - l[0] * r[2];
         double az = l[0] * r[1] - l[1] * r[0];
         delx[t][i] = vol / sqrt(ax * ax + ay * ay + az * az + ptiny);
         ax *= norm; ay *= norm; az *= norm;
         delv[t][i] = ax * V[t][0] + ay * V[t][1] + az * V[t][2];
      }
   }
}

/* ------------------------------------------------------------------------ */
/* A13 monotonic-Q limiter per region + q-stop                               */
/* ------------------------------------------------------------------------ */

static inline double limiter(double self, double dm, double dp, double mult, double maxs)
{
   double norm = 1. / (self + 1.e-36);
   dm = dm * norm; dp = dp * norm;
   double phi = .5 * (dm + dp);
   dm *= mult; dp *= mult;
   if (dm < phi) phi = dm;
   if (dp < phi) phi = dp;
   if (phi < 0.) phi = 0.;
   if (phi > maxs) phi = maxs;
   return phi;
}

static inline double nbr(const double *a, int self, int other, int bc, int symm, int fre)
{
   /* lulesh.cc:1781-1800: COMM and interior read the neighbour slot */
   if (bc == symm) return a[self];
   if (bc == fre) return 0.0;
   return a[other];
}

int ora_monoq_regions(ora_domain *d)
{
   const lulesh_b200_constants *c = &d->c;
   const double *dvx = F(d, DELV_XI), *dve = F(d, DELV_ETA), *dvz = F(d, DELV_ZETA);
   for (int r = 0; r < d->numReg; ++r) {
      const int *list = d->regElemlist[r];
#pragma omp parallel for
      for (int t = 0; t < d->regElemSize[r]; ++t) {
         int i = list[t];
         int bc = d->elemBC[i];
         double phixi = limiter(dvx[i],
            nbr(dvx, i, d->lxim[i], bc & XI_M, XI_M_SYMM, XI_M_FREE),
            nbr(dvx, i, d->lxip[i], bc & XI_P, XI_P_SYMM, XI_P_FREE),
            c->monoq_limiter_mult, c->monoq_max_slope);
         double phieta = limiter(dve[i],
            nbr(dve, i, d->letam[i], bc & ETA_M, ETA_M_SYMM, ETA_M_FREE),
            nbr(dve, i, d->letap[i], bc & ETA_P, ETA_P_SYMM, ETA_P_FREE),
            c->monoq_limiter_mult, c->monoq_max_slope);
         double phizeta = limiter(dvz[i],
            nbr(dvz, i, d->lzetam[i], bc & ZETA_M, ZETA_M_SYMM, ZETA_M_FREE),
            nbr(dvz, i, d->lzetap[i], bc & ZETA_P, ZETA_P_SYMM, ZETA_P_FREE),
            c->monoq_limiter_mult, c->monoq_max_slope);
         double qlin, qquad;
         if (F(d, VDOV)[i] > 0.) { qlin = 0.; qquad = 0.; }
         else { /* lulesh.cc:1897-1915 */
            double a = dvx[i] * F(d, DELX_XI)[i];
            double b = dve[i] * F(d, DELX_ETA)[i];
            double g = dvz[i] * F(d, DELX_ZETA)[i];
            if (a > 0.) a = 0.;
            if (b > 0.) b = 0.;
            if (g > 0.) g = 0.;
            double rho = F(d, ELEMMASS)[i] / (F(d, VOLO)[i] * F(d, VNEW)[i]);
            qlin = -c->qlc_monoq * rho *
                   (a * (1. - phixi) + b * (1. - phieta) + g * (1. - phizeta));
            qquad = c->qqc_monoq * rho *
                    (a * a * (1. - phixi * phixi) + b * b * (1. - phieta * phieta) +
                     g * g * (1. - phizeta * phizeta));
         }
         F(d, QQ)[i] = qquad;
         F(d, QL)[i] = qlin;
      }
   }
   for (int i = 0; i < d->numElem; ++i) /* lulesh.cc:1994-2008, tests the OLD q */
      if (F(d, Q)[i] > c->qstop) return LULESH_B200_QSTOP_ERROR;
   return 0;
}

/* ------------------------------------------------------------------------ */
/* A14-A16 material update                                                   */
/* ------------------------------------------------------------------------ */

/* CalcPressureForElems on one element, lulesh.cc:2015-2044 */
static inline double eos_pressure(double *bvc, double *pbvc, double e, double comp,
                                  double vnewc, const lulesh_b200_constants *c)
{
   const double c1s = 2.0 / 3.0;
   *bvc = c1s * (comp + 1.);
   *pbvc = c1s;
   double p = *bvc * e;
   if (fabs(p) < c->p_cut) p = 0.0;
   if (vnewc >= c->eosvmax) p = 0.0;
   if (p < c->pmin) p = c->pmin;
   return p;
}

static inline double eos_ssc(double pbvc, double e, double vol, double bvc, double p, double rho0)
{
   double ssc = (pbvc * e + vol * vol * bvc * p) / rho0; /* lulesh.cc:2083-2090 */
   if (ssc <= .1111111e-36) ssc = .3333333e-18;
   else ssc = sqrt(ssc);
   return ssc;
}

static int region_rep(int r, int numReg, int cost) /* lulesh.cc:2393-2400 */
{
   if (r < numReg / 2) return 1;
   if (r < numReg - (numReg + 15) / 20) return 1 + cost;
   return 10 * (1 + cost);
}

int ora_material(ora_domain *d)
{
   const lulesh_b200_constants *c = &d->c;
   const int ne = d->numElem;
   if (ne == 0) return 0;
   int err = 0;
   for (int i = 0; i < ne; ++i) { /* lulesh.cc:2366-2384 */
      double vc = F(d, V)[i];
      if (c->eosvmin != 0. && vc < c->eosvmin) vc = c->eosvmin;
      if (c->eosvmax != 0. && vc > c->eosvmax) vc = c->eosvmax;
      if (vc <= 0.) err = 1;
   }
   if (err) return LULESH_B200_VOLUME_ERROR;

   const double rho0 = c->refdens;
   for (int r = 0; r < d->numReg; ++r) {
      const int *list = d->regElemlist[r];
      const int rep = region_rep(r, d->numReg, d->cost);
#pragma omp parallel for
      for (int t = 0; t < d->regElemSize[r]; ++t) {
         const int i = list[t];
         double vnewc = F(d, VNEW)[i]; /* lulesh.cc:2342-2361 */
         if (c->eosvmin != 0. && vnewc < c->eosvmin) vnewc = c->eosvmin;
         if (c->eosvmax != 0. && vnewc > c->eosvmax) vnewc = c->eosvmax;
         double p_new = 0, e_new = 0, q_new = 0, bvc = 0, pbvc = 0;
         for (int j = 0; j < rep; ++j) { /* lulesh.cc:2238-2295 */
            double e_old = F(d, E)[i], delvc = F(d, DELV)[i];
            double p_old = F(d, P)[i], q_old = F(d, Q)[i];
            double qq_old = F(d, QQ)[i], ql_old = F(d, QL)[i];
            double comp = 1. / vnewc - 1.;
            double vchalf = vnewc - delvc * .5;
            double compHalf = 1. / vchalf - 1.;
            if (c->eosvmin != 0. && vnewc <= c->eosvmin) compHalf = comp;
            if (c->eosvmax != 0. && vnewc >= c->eosvmax) { p_old = 0.; comp = 0.; compHalf = 0.; }
            const double work = 0.;
            /* CalcEnergyForElems, lulesh.cc:2049-2176 */
            e_new = e_old - 0.5 * delvc * (p_old + q_old) + 0.5 * work;
            if (e_new < c->emin) e_new = c->emin;
            double pHalf = eos_pressure(&bvc, &pbvc, e_new, compHalf, vnewc, c);
            double vhalf = 1. / (1. + compHalf);
            if (delvc > 0.) q_new = 0.;
            else q_new = eos_ssc(pbvc, e_new, vhalf, bvc, pHalf, rho0) * ql_old + qq_old;
            e_new = e_new + 0.5 * delvc * (3.0 * (p_old + q_old) - 4.0 * (pHalf + q_new));
            e_new += 0.5 * work;
            if (fabs(e_new) < c->e_cut) e_new = 0.;
            if (e_new < c->emin) e_new = c->emin;
            p_new = eos_pressure(&bvc, &pbvc, e_new, comp, vnewc, c);
            double q_tilde;
            if (delvc > 0.) q_tilde = 0.;
            else q_tilde = eos_ssc(pbvc, e_new, vnewc, bvc, p_new, rho0) * ql_old + qq_old;
            const double sixth = 1.0 / 6.0;
            e_new = e_new - (7.0 * (p_old + q_old) - 8.0 * (pHalf + q_new) + (p_new + q_tilde)) *
                               delvc * sixth;
            if (fabs(e_new) < c->e_cut) e_new = 0.;
            if (e_new < c->emin) e_new = c->emin;
            p_new = eos_pressure(&bvc, &pbvc, e_new, comp, vnewc, c);
            if (delvc <= 0.) {
               q_new = eos_ssc(pbvc, e_new, vnewc, bvc, p_new, rho0) * ql_old + qq_old;
               if (fabs(q_new) < c->q_cut) q_new = 0.;
            }
         }
         F(d, P)[i] = p_new; F(d, E)[i] = e_new; F(d, Q)[i] = q_new; /* lulesh.cc:2297-2303 */
         F(d, SS)[i] = eos_ssc(pbvc, e_new, vnewc, bvc, p_new, rho0); /* lulesh.cc:2187-2199 */
      }
   }
#pragma omp parallel for
   for (int i = 0; i < ne; ++i) { /* UpdateVolumesForElems, lulesh.cc:2411-2427 */
      double t = F(d, VNEW)[i];
      if (fabs(t - 1.0) < c->v_cut) t = 1.0;
      F(d, V)[i] = t;
   }
   return 0;
}

/* ------------------------------------------------------------------------ */
/* A17 time constraints                                                      */
/* ------------------------------------------------------------------------ */

void ora_time_constraints(ora_domain *d)
{
   const double qqc2 = 64.0 * d->c.qqc * d->c.qqc;
   double dtc = 1.0e+20, dth = 1.0e+20;
   for (int r = 0; r < d->numReg; ++r) {
      const int *list = d->regElemlist[r];
      for (int t = 0; t < d->regElemSize[r]; ++t) {
         int i = list[t];
         double ss = F(d, SS)[i], vdov = F(d, VDOV)[i], al = F(d, AREALG)[i];
         double dtf = ss * ss; /* lulesh.cc:2477-2493 */
         if (vdov < 0.) dtf = dtf + qqc2 * al * al * vdov * vdov;
         dtf = sqrt(dtf);
         dtf = al / dtf;
         if (vdov != 0.) {
            if (dtf < dtc) dtc = dtf;
            double dtdvov = d->c.dvovmax / (fabs(vdov) + 1.e-20); /* lulesh.cc:2546-2553 */
            if (dth > dtdvov) dth = dtdvov;
         }
      }
   }
   d->s.dtcourant = dtc;
   d->s.dthydro = dth;
}

/* ------------------------------------------------------------------------ */
/* the cycle                                                                 */
/* ------------------------------------------------------------------------ */

int ora_step(ora_domain *d)
{
   int rc;
   ora_time_increment(d, ora_dt_candidate(d));
   if ((rc = ora_calc_force(d))) return rc;
   ora_node_update(d);
   if ((rc = ora_kinematics(d))) return rc;
   ora_monoq_gradients(d);
   if ((rc = ora_monoq_regions(d))) return rc;
   if ((rc = ora_material(d))) return rc;
   ora_time_constraints(d);
   return 0;
}

int ora_run(ora_domain *d, int max_cycles)
{
   while (d->s.time < d->s.stoptime && d->s.cycle < max_cycles) {
      int rc = ora_step(d);
      if (rc) return rc;
   }
   return 0;
}

void ora_symmetry(ora_domain *d, int n, double out3[3])
{
   const double *e = F(d, E);
   double maxAbs = 0, tot = 0, maxRel = 0;
   for (int j = 0; j < n; ++j)
      for (int k = j + 1; k < n; ++k) {
         double a = fabs(e[j * d->sx + k] - e[k * d->sx + j]);
         tot += a;
         if (maxAbs < a) maxAbs = a;
         double r = a / e[k * d->sx + j];
         if (maxRel < r) maxRel = r;
      }
   out3[0] = maxAbs; out3[1] = tot; out3[2] = maxRel;
}

/* ------------------------------------------------------------------------ */
/* setup (lulesh-init.cc), generalised to (px,py,pz) x (sx,sy,sz)            */
/* ------------------------------------------------------------------------ */

static void *zalloc(size_t n, size_t w) { return calloc(n ? n : 1, w); }

static void create_regions(ora_domain *d, int nr, int balance)
{
   /* lulesh-init.cc:401-510; glibc rand() seeded with the rank */
   const int ne = d->numElem;
   srand(d->rank);
   d->numReg = nr;
   d->regElemSize = zalloc(nr, sizeof(int));
   d->regElemlist = zalloc(nr, sizeof(int *));
   d->regNumList = zalloc(ne, sizeof(int));
   int next = 0;
   if (nr == 1) {
      for (; next < ne; ++next) d->regNumList[next] = 1;
   } else {
      int lastReg = -1, costDenominator = 0;
      int *binEnd = zalloc(nr, sizeof(int));
      for (int i = 0; i < nr; ++i) {
         costDenominator += pow((i + 1), balance);
         binEnd[i] = costDenominator;
      }
      while (next < ne) {
         int regionNum;
         do {
            int var = rand() % costDenominator, i = 0;
            while (var >= binEnd[i]) i++;
            regionNum = ((i + d->rank) % nr) + 1;
         } while (regionNum == lastReg);
         int binSize = rand() % 1000, elements;
         if (binSize < 773) elements = rand() % 15 + 1;
         else if (binSize < 937) elements = rand() % 16 + 16;
         else if (binSize < 970) elements = rand() % 32 + 32;
         else if (binSize < 974) elements = rand() % 64 + 64;
         else if (binSize < 978) elements = rand() % 128 + 128;
         else if (binSize < 981) elements = rand() % 256 + 256;
         else elements = rand() % 1537 + 512;
         int runto = elements + next;
         while (next < runto && next < ne) d->regNumList[next++] = regionNum;
         lastReg = regionNum;
      }
      free(binEnd);
   }
   for (int i = 0; i < ne; ++i) d->regElemSize[d->regNumList[i] - 1]++;
   for (int r = 0; r < nr; ++r) {
      d->regElemlist[r] = zalloc(d->regElemSize[r], sizeof(int));
      d->regElemSize[r] = 0;
   }
   for (int i = 0; i < ne; ++i) {
      int r = d->regNumList[i] - 1;
      d->regElemlist[r][d->regElemSize[r]++] = i;
   }
}

static int imax3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }

ora_domain *ora_new(int numRanks, int rank, int px, int py, int pz, int sx, int sy, int sz,
                    int nr, int balance, int cost)
{
   if (numRanks != px * py * pz || rank < 0 || rank >= numRanks || sx < 1 || sy < 1 || sz < 1 ||
       nr < 1)
      return NULL;
   ora_domain *d = calloc(1, sizeof(*d));
   d->numRanks = numRanks; d->rank = rank;
   d->px = px; d->py = py; d->pz = pz;
   d->sx = sx; d->sy = sy; d->sz = sz;
   d->col = rank % px; d->row = (rank / px) % py; d->plane = rank / (px * py);
   d->cost = cost;
   const int nx1 = sx + 1, ny1 = sy + 1, nz1 = sz + 1;
   const int ne = d->numElem = sx * sy * sz;
   const int nn = d->numNode = nx1 * ny1 * nz1;
   d->allElem = ne + 2 * sx * sy + 2 * sx * sz + 2 * sy * sz; /* lulesh.cc:1955-1958 */
   d->cMin = d->col != 0; d->cMax = d->col != px - 1;
   d->rMin = d->row != 0; d->rMax = d->row != py - 1;
   d->pMin = d->plane != 0; d->pMax = d->plane != pz - 1;

   d->c = (lulesh_b200_constants){ /* lulesh-init.cc:20-38 */
      .e_cut = 1.0e-7, .p_cut = 1.0e-7, .q_cut = 1.0e-7, .v_cut = 1.0e-10, .u_cut = 1.0e-7,
      .hgcoef = 3.0, .ss4o3 = 4.0 / 3.0, .qstop = 1.0e+12, .monoq_max_slope = 1.0,
      .monoq_limiter_mult = 2.0, .qlc_monoq = 0.5, .qqc_monoq = 2.0 / 3.0, .qqc = 2.0,
      .eosvmax = 1.0e+9, .eosvmin = 1.0e-9, .pmin = 0., .emin = -1.0e+15, .dvovmax = 0.1,
      .refdens = 1.0};

   for (int id = 0; id < LULESH_F_COUNT; ++id) d->f[id] = zalloc(ora_real_count(d, id), sizeof(double));
   d->nodelist = zalloc(8 * (size_t)ne, sizeof(int));
   d->lxim = zalloc(ne, sizeof(int)); d->lxip = zalloc(ne, sizeof(int));
   d->letam = zalloc(ne, sizeof(int)); d->letap = zalloc(ne, sizeof(int));
   d->lzetam = zalloc(ne, sizeof(int)); d->lzetap = zalloc(ne, sizeof(int));
   d->elemBC = zalloc(ne, sizeof(int));
   d->cfx = zalloc(8 * (size_t)ne, sizeof(double));
   d->cfy = zalloc(8 * (size_t)ne, sizeof(double));
   d->cfz = zalloc(8 * (size_t)ne, sizeof(double));
   for (int i = 0; i < ne; ++i) F(d, V)[i] = 1.0; /* lulesh-init.cc:98-100 */

   /* BuildMesh, lulesh-init.cc:218-267.  G = longest global edge in elements;
    * equals tp*nx for the reference's cubic layouts. */
   const int G = imax3(px * sx, py * sy, pz * sz);
   for (int k = 0, n = 0; k < nz1; ++k)
      for (int j = 0; j < ny1; ++j)
         for (int i = 0; i < nx1; ++i, ++n) {
            F(d, X)[n] = 1.125 * (double)(d->col * sx + i) / (double)G;
            F(d, Y)[n] = 1.125 * (double)(d->row * sy + j) / (double)G;
            F(d, Z)[n] = 1.125 * (double)(d->plane * sz + k) / (double)G;
         }
   for (int k = 0, e = 0; k < sz; ++k)
      for (int j = 0; j < sy; ++j)
         for (int i = 0; i < sx; ++i, ++e) {
            int n = k * nx1 * ny1 + j * nx1 + i;
            int *nl = &d->nodelist[8 * e];
            nl[0] = n; nl[1] = n + 1; nl[2] = n + nx1 + 1; nl[3] = n + nx1;
            for (int c = 0; c < 4; ++c) nl[4 + c] = nl[c] + nx1 * ny1;
         }

   /* SetupThreadSupportStructures, lulesh-init.cc:272-337 (always built here) */
   d->nodeElemStart = zalloc(nn + 1, sizeof(int));
   for (int i = 0; i < 8 * ne; ++i) d->nodeElemStart[d->nodelist[i] + 1]++;
   for (int n = 0; n < nn; ++n) d->nodeElemStart[n + 1] += d->nodeElemStart[n];
   d->nodeElemCornerList = zalloc(8 * (size_t)ne, sizeof(int));
   {
      int *fill = zalloc(nn, sizeof(int));
      for (int i = 0; i < 8 * ne; ++i) {
         int n = d->nodelist[i];
         d->nodeElemCornerList[d->nodeElemStart[n] + fill[n]++] = i;
      }
      free(fill);
   }

   create_regions(d, nr, balance);

   /* SetupSymmetryPlanes, lulesh-init.cc:514-533 (sizes per appendix C) */
   if (d->plane == 0) {
      d->nsymmZ = nx1 * ny1; d->symmZ = zalloc(d->nsymmZ, sizeof(int));
      for (int j = 0, t = 0; j < ny1; ++j) for (int i = 0; i < nx1; ++i) d->symmZ[t++] = j * nx1 + i;
   }
   if (d->row == 0) {
      d->nsymmY = nx1 * nz1; d->symmY = zalloc(d->nsymmY, sizeof(int));
      for (int k = 0, t = 0; k < nz1; ++k) for (int i = 0; i < nx1; ++i) d->symmY[t++] = k * nx1 * ny1 + i;
   }
   if (d->col == 0) {
      d->nsymmX = ny1 * nz1; d->symmX = zalloc(d->nsymmX, sizeof(int));
      for (int k = 0, t = 0; k < nz1; ++k) for (int j = 0; j < ny1; ++j) d->symmX[t++] = k * nx1 * ny1 + j * nx1;
   }

   /* SetupElementConnectivities + SetupBoundaryConditions, lulesh-init.cc:539-673 */
   int ghost[6], pidx = ne;
   for (int f = 0; f < 6; ++f) ghost[f] = INT_MIN;
   if (d->pMin) { ghost[0] = pidx; pidx += sx * sy; }
   if (d->pMax) { ghost[1] = pidx; pidx += sx * sy; }
   if (d->rMin) { ghost[2] = pidx; pidx += sx * sz; }
   if (d->rMax) { ghost[3] = pidx; pidx += sx * sz; }
   if (d->cMin) { ghost[4] = pidx; pidx += sy * sz; }
   if (d->cMax) { ghost[5] = pidx; }
   for (int k = 0, e = 0; k < sz; ++k)
      for (int j = 0; j < sy; ++j)
         for (int i = 0; i < sx; ++i, ++e) {
            int bc = 0;
            /* lulesh-init.cc:541-564: index arithmetic that "wraps" at brick faces;
             * those entries are masked by elemBC and never dereferenced */
            d->lxim[e] = (e >= 1) ? e - 1 : e;
            d->lxip[e] = (e < ne - 1) ? e + 1 : e;
            d->letam[e] = (e >= sx) ? e - sx : e;
            d->letap[e] = (e < ne - sx) ? e + sx : e;
            d->lzetam[e] = (e >= sx * sy) ? e - sx * sy : e;
            d->lzetap[e] = (e < ne - sx * sy) ? e + sx * sy : e;
            if (k == 0) {
               if (d->plane == 0) bc |= ZETA_M_SYMM;
               else { bc |= ZETA_M_COMM; d->lzetam[e] = ghost[0] + j * sx + i; }
            }
            if (k == sz - 1) {
               if (d->plane == pz - 1) bc |= ZETA_P_FREE;
               else { bc |= ZETA_P_COMM; d->lzetap[e] = ghost[1] + j * sx + i; }
            }
            if (j == 0) {
               if (d->row == 0) bc |= ETA_M_SYMM;
               else { bc |= ETA_M_COMM; d->letam[e] = ghost[2] + k * sx + i; }
            }
            if (j == sy - 1) {
               if (d->row == py - 1) bc |= ETA_P_FREE;
               else { bc |= ETA_P_COMM; d->letap[e] = ghost[3] + k * sx + i; }
            }
            if (i == 0) {
               if (d->col == 0) bc |= XI_M_SYMM;
               else { bc |= XI_M_COMM; d->lxim[e] = ghost[4] + k * sy + j; }
            }
            if (i == sx - 1) {
               if (d->col == px - 1) bc |= XI_P_FREE;
               else { bc |= XI_P_COMM; d->lxip[e] = ghost[5] + k * sy + j; }
            }
            d->elemBC[e] = bc;
         }

   /* time controls, lulesh-init.cc:146-156 */
   d->s = (lulesh_b200_scalars){.dtcourant = 1.0e+20, .dthydro = 1.0e+20, .dtfixed = -1.0e-6,
                                .time = 0., .deltatime = 0., .deltatimemultlb = 1.1,
                                .deltatimemultub = 1.2, .dtmax = 1.0e-2, .stoptime = 1.0e-2,
                                .cycle = 0, .error = 0};

   /* volo, elemMass, nodalMass, lulesh-init.cc:159-178 */
   for (int e = 0; e < ne; ++e) {
      const int *nl = &d->nodelist[8 * e];
      double x[8], y[8], z[8];
      gather8(F(d, X), nl, x); gather8(F(d, Y), nl, y); gather8(F(d, Z), nl, z);
      double volume = elem_volume(x, y, z);
      F(d, VOLO)[e] = volume;
      F(d, ELEMMASS)[e] = volume;
      for (int c = 0; c < 8; ++c) F(d, NODALMASS)[nl[c]] += volume / 8.0;
   }

   /* energy deposit and dt0, lulesh-init.cc:183-192.  dt0 uses the volume of the
    * GLOBAL origin element on every rank -- on the origin rank this
    * is volo(0), i.e. the reference's value. */
   const double ebase = 3.948746e+7;
   double scale = (double)G / 45.0;
   double einit = ebase * scale * scale * scale;
   if (d->row + d->col + d->plane == 0) F(d, E)[0] = einit;
   {
      double x[8], y[8], z[8];
      for (int c = 0; c < 8; ++c) {
         int i = (c == 1 || c == 2 || c == 5 || c == 6), j = (c == 2 || c == 3 || c == 6 || c == 7),
             k = c >= 4;
         x[c] = 1.125 * (double)i / (double)G;
         y[c] = 1.125 * (double)j / (double)G;
         z[c] = 1.125 * (double)k / (double)G;
      }
      d->s.deltatime = (.5 * cbrt(elem_volume(x, y, z))) / sqrt(2.0 * einit);
   }
   return d;
}

/* The reference derives deltatime from the rank's OWN volo(0) (lulesh-init.cc:192), which
 * differs by an ulp between ranks for sizes whose coordinates are not exact in binary
 * (SURVEY F10).  The product and ora_new() use the global-origin element on every rank; this
 * switch restores the reference's per-rank value so that the multi-rank emulation can be
 * pinned bit-for-bit against the reference's MPI build (oracle/_ref/lulesh_mpi). */
void ora_use_reference_dt0(ora_domain *d)
{
   const int G = imax3(d->px * d->sx, d->py * d->sy, d->pz * d->sz);
   const double scale = (double)G / 45.0;
   const double einit = 3.948746e+7 * scale * scale * scale;
   d->s.deltatime = (.5 * cbrt(F(d, VOLO)[0])) / sqrt(2.0 * einit);
}

void ora_free(ora_domain *d)
{
   if (!d) return;
   for (int id = 0; id < LULESH_F_COUNT; ++id) free(d->f[id]);
   free(d->nodelist); free(d->lxim); free(d->lxip); free(d->letam); free(d->letap);
   free(d->lzetam); free(d->lzetap); free(d->elemBC);
   free(d->symmX); free(d->symmY); free(d->symmZ);
   for (int r = 0; r < d->numReg; ++r) free(d->regElemlist[r]);
   free(d->regElemlist); free(d->regElemSize); free(d->regNumList);
   free(d->nodeElemStart); free(d->nodeElemCornerList);
   free(d->cfx); free(d->cfy); free(d->cfz);
   free(d);
}

size_t ora_real_count(ora_domain *d, int field)
{
   if (field < 0 || field >= LULESH_F_COUNT) return 0;
   if (field <= LULESH_F_NODALMASS) return (size_t)d->numNode;
   if (field >= LULESH_F_DELV_XI && field <= LULESH_F_DELV_ZETA) return (size_t)d->allElem;
   return (size_t)d->numElem;
}

double *ora_real(ora_domain *d, int field)
{
   return (field < 0 || field >= LULESH_F_COUNT) ? NULL : d->f[field];
}

int *ora_int(ora_domain *d, const char *name, int *count)
{
#define RET(n, p, c) if (!strcmp(name, n)) { if (count) *count = (c); return (p); }
   RET("nodelist", d->nodelist, 8 * d->numElem)
   RET("lxim", d->lxim, d->numElem) RET("lxip", d->lxip, d->numElem)
   RET("letam", d->letam, d->numElem) RET("letap", d->letap, d->numElem)
   RET("lzetam", d->lzetam, d->numElem) RET("lzetap", d->lzetap, d->numElem)
   RET("elemBC", d->elemBC, d->numElem) RET("regNumList", d->regNumList, d->numElem)
   RET("regElemSize", d->regElemSize, d->numReg)
   RET("symmX", d->symmX, d->nsymmX) RET("symmY", d->symmY, d->nsymmY)
   RET("symmZ", d->symmZ, d->nsymmZ)
   RET("nodeElemStart", d->nodeElemStart, d->numNode + 1)
   RET("nodeElemCornerList", d->nodeElemCornerList, 8 * d->numElem)
#undef RET
   if (count) *count = 0;
   return NULL;
}

int *ora_region_list(ora_domain *d, int r, int *count)
{
   if (r < 0 || r >= d->numReg) { if (count) *count = 0; return NULL; }
   if (count) *count = d->regElemSize[r];
   return d->regElemlist[r];
}

lulesh_b200_scalars *ora_scalars(ora_domain *d) { return &d->s; }
lulesh_b200_constants *ora_constants(ora_domain *d) { return &d->c; }

/* ------------------------------------------------------------------------ */
/* multi-rank emulation (lulesh-comm.cc semantics, in-process)               */
/* ------------------------------------------------------------------------ */

struct ora_multi {
   int px, py, pz, n;
   ora_domain **dom;
};

/* the 26 neighbour directions (dcol,drow,dplane) in CommSBN's unpack order,
 * lulesh-comm.cc:891-1256: 6 faces, 12 edges, 8 corners */
static const int k_dirs[26][3] = {
   {0, 0, -1}, {0, 0, 1}, {0, -1, 0}, {0, 1, 0}, {-1, 0, 0}, {1, 0, 0},
   {-1, -1, 0}, {0, -1, -1}, {-1, 0, -1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1},
   {-1, 1, 0}, {0, -1, 1}, {-1, 0, 1}, {1, -1, 0}, {0, 1, -1}, {1, 0, -1},
   {-1, -1, -1}, {-1, -1, 1}, {1, -1, -1}, {1, -1, 1},
   {-1, 1, -1}, {-1, 1, 1}, {1, 1, -1}, {1, 1, 1}};

static int neighbour_rank(const ora_domain *d, const int dir[3])
{
   int c = d->col + dir[0], r = d->row + dir[1], p = d->plane + dir[2];
   if (c < 0 || c >= d->px || r < 0 || r >= d->py || p < 0 || p >= d->pz) return -1;
   return p * d->px * d->py + r * d->px + c;
}

/* node ids on the part of the boundary shared with the neighbour in `dir`,
 * plane-major / row / col order (both sides enumerate the same global nodes
 * in the same order) */
static int shared_nodes(const ora_domain *d, const int dir[3], int *out)
{
   const int nx1 = d->sx + 1, ny1 = d->sy + 1, nz1 = d->sz + 1;
   int i0 = dir[0] < 0 ? 0 : (dir[0] > 0 ? d->sx : 0), i1 = dir[0] == 0 ? nx1 : i0 + 1;
   int j0 = dir[1] < 0 ? 0 : (dir[1] > 0 ? d->sy : 0), j1 = dir[1] == 0 ? ny1 : j0 + 1;
   int k0 = dir[2] < 0 ? 0 : (dir[2] > 0 ? d->sz : 0), k1 = dir[2] == 0 ? nz1 : k0 + 1;
   int t = 0;
   for (int k = k0; k < k1; ++k)
      for (int j = j0; j < j1; ++j)
         for (int i = i0; i < i1; ++i) out[t++] = k * nx1 * ny1 + j * nx1 + i;
   return t;
}

/* exchange node fields: mode 0 = add from all 26 (CommSBN), mode 1 = overwrite
 * from lexicographically-higher (plane,row,col) neighbours (CommSyncPosVel) */
static void exchange_nodes(ora_multi *m, const int *fields, int nf, int mode)
{
   /* snapshot ("send buffers are packed before any unpack") */
   double **snap = calloc((size_t)m->n * nf, sizeof(double *));
   for (int r = 0; r < m->n; ++r)
      for (int f = 0; f < nf; ++f) {
         size_t cnt = m->dom[r]->numNode;
         snap[r * nf + f] = malloc(cnt * sizeof(double));
         memcpy(snap[r * nf + f], m->dom[r]->f[fields[f]], cnt * sizeof(double));
      }
   for (int r = 0; r < m->n; ++r) {
      ora_domain *d = m->dom[r];
      int cap = imax3((d->sx + 1) * (d->sy + 1), (d->sx + 1) * (d->sz + 1), (d->sy + 1) * (d->sz + 1));
      int *mine = malloc(cap * sizeof(int)), *theirs = malloc(cap * sizeof(int));
      for (int q = 0; q < 26; ++q) {
         const int *dir = k_dirs[q];
         int nb = neighbour_rank(d, dir);
         if (nb < 0) continue;
         if (mode == 1) { /* receive only from (dp,dr,dc) lexicographically > 0 */
            int lex = dir[2] != 0 ? dir[2] : (dir[1] != 0 ? dir[1] : dir[0]);
            if (lex < 0) continue;
         }
         int opp[3] = {-dir[0], -dir[1], -dir[2]};
         int cnt = shared_nodes(d, dir, mine);
         int cnt2 = shared_nodes(m->dom[nb], opp, theirs);
         if (cnt != cnt2) { fprintf(stderr, "oracle: halo size mismatch\n"); abort(); }
         for (int f = 0; f < nf; ++f) {
            double *dst = d->f[fields[f]];
            const double *src = snap[nb * nf + f];
            if (mode == 0) for (int t = 0; t < cnt; ++t) dst[mine[t]] += src[theirs[t]];
            else           for (int t = 0; t < cnt; ++t) dst[mine[t]]  = src[theirs[t]];
         }
      }
      free(mine); free(theirs);
   }
   for (int i = 0; i < m->n * nf; ++i) free(snap[i]);
   free(snap);
}

/* CommMonoQ, lulesh-comm.cc:1684-1835: face-neighbour boundary layers of
 * delv_* copied into the ghost slots in the order pMin,pMax,rMin,rMax,cMin,cMax */
static void exchange_monoq(ora_multi *m)
{
   static const int fid[3] = {LULESH_F_DELV_XI, LULESH_F_DELV_ETA, LULESH_F_DELV_ZETA};
   for (int r = 0; r < m->n; ++r) {
      ora_domain *d = m->dom[r];
      int off = d->numElem;
      for (int face = 0; face < 6; ++face) {
         const int *dir = k_dirs[face];
         int nb = neighbour_rank(d, dir);
         if (nb < 0) continue;
         const ora_domain *s = m->dom[nb];
         int cnt = 0;
         for (int f = 0; f < 3; ++f) {
            const double *src = s->f[fid[f]];
            double *dst = d->f[fid[f]] + off;
            int t = 0;
            if (dir[2] != 0) { /* sender's opposite plane layer */
               int k = dir[2] < 0 ? s->sz - 1 : 0;
               for (int j = 0; j < s->sy; ++j) for (int i = 0; i < s->sx; ++i)
                  dst[t++] = src[k * s->sx * s->sy + j * s->sx + i];
            } else if (dir[1] != 0) {
               int j = dir[1] < 0 ? s->sy - 1 : 0;
               for (int k = 0; k < s->sz; ++k) for (int i = 0; i < s->sx; ++i)
                  dst[t++] = src[k * s->sx * s->sy + j * s->sx + i];
            } else {
               int i = dir[0] < 0 ? s->sx - 1 : 0;
               for (int k = 0; k < s->sz; ++k) for (int j = 0; j < s->sy; ++j)
                  dst[t++] = src[k * s->sx * s->sy + j * s->sx + i];
            }
            cnt = t;
         }
         off += cnt;
      }
   }
}

ora_multi *ora_multi_new(int px, int py, int pz, int sx, int sy, int sz, int nr, int balance,
                         int cost)
{
   ora_multi *m = calloc(1, sizeof(*m));
   m->px = px; m->py = py; m->pz = pz; m->n = px * py * pz;
   m->dom = calloc(m->n, sizeof(ora_domain *));
   for (int r = 0; r < m->n; ++r) {
      m->dom[r] = ora_new(m->n, r, px, py, pz, sx, sy, sz, nr, balance, cost);
      if (!m->dom[r]) { ora_multi_free(m); return NULL; }
   }
   if (m->n > 1) { /* lulesh.cc:2720-2729 */
      const int f[1] = {LULESH_F_NODALMASS};
      exchange_nodes(m, f, 1, 0);
   }
   return m;
}

void ora_multi_free(ora_multi *m)
{
   if (!m) return;
   for (int r = 0; r < m->n; ++r) ora_free(m->dom[r]);
   free(m->dom);
   free(m);
}

ora_domain *ora_multi_rank(ora_multi *m, int rank)
{
   return (rank < 0 || rank >= m->n) ? NULL : m->dom[rank];
}

int ora_multi_step(ora_multi *m)
{
   int rc = 0;
   double newdt = 1.0e+300;
   for (int r = 0; r < m->n; ++r) { /* per-rank rule, then MIN (lulesh.cc:176-188) */
      double g = ora_dt_candidate(m->dom[r]);
      if (g < newdt) newdt = g;
   }
   for (int r = 0; r < m->n; ++r) ora_time_increment(m->dom[r], newdt);
   for (int r = 0; r < m->n; ++r) if ((rc = ora_calc_force(m->dom[r]))) return rc;
   if (m->n > 1) {
      const int f[3] = {LULESH_F_FX, LULESH_F_FY, LULESH_F_FZ};
      exchange_nodes(m, f, 3, 0);
   }
   for (int r = 0; r < m->n; ++r) ora_node_update(m->dom[r]);
   if (m->n > 1) {
      const int f[6] = {LULESH_F_X, LULESH_F_Y, LULESH_F_Z, LULESH_F_XD, LULESH_F_YD, LULESH_F_ZD};
      exchange_nodes(m, f, 6, 1);
   }
   for (int r = 0; r < m->n; ++r) {
      if ((rc = ora_kinematics(m->dom[r]))) return rc;
      ora_monoq_gradients(m->dom[r]);
   }
   if (m->n > 1) exchange_monoq(m);
   for (int r = 0; r < m->n; ++r) {
      if ((rc = ora_monoq_regions(m->dom[r]))) return rc;
      if ((rc = ora_material(m->dom[r]))) return rc;
      ora_time_constraints(m->dom[r]);
   }
   return 0;
}

int ora_multi_run(ora_multi *m, int max_cycles)
{
   ora_domain *d0 = m->dom[0];
   while (d0->s.time < d0->s.stoptime && d0->s.cycle < max_cycles) {
      int rc = ora_multi_step(m);
      if (rc) return rc;
   }
   return 0;
}
