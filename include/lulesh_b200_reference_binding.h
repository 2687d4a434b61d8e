/*
 * lulesh_b200_reference_binding.h -- the reference-side binding of the B200 step.
 *
 * LULESH 2.0 has no plugin or FFI layer; this header is what a maintainer adds to the
 * reference's lulesh.cc (after `#include "lulesh.h"`) to run the timed loop of main()
 * (lulesh.cc:2745-2757) on a B200 through the C ABI of lulesh_b200.h.  It uses nothing but
 * the reference's public `Domain` accessors (lulesh.h:266-429) and `cmdLineOpts`
 * (lulesh.h:600-611): no private member is touched.
 *
 *     // lulesh.cc, in main(), instead of the while loop at 2745-2757:
 *     B200TimedLoop(*locDom, opts, myRank, numRanks);
 *
 * Everything before (command line, banner, InitMeshDecomp, `new Domain`, timer start) and
 * after (timer stop, VerifyAndWriteFinalOutput) stays the reference's own code.
 * oracle/Makefile builds exactly that program as oracle/_ref/lulesh_patched (test
 * infrastructure: the patched copy of lulesh.cc is generated into the git-ignored
 * oracle/_ref/ and never committed), and tests/test_gpu_parity.py runs it against the goldens.
 *
 * One rank only: the USE_MPI=0 build.  With MPI the same code applies per rank plus the
 * unique-id broadcast shown in INTEGRATION.md.
 */
#ifndef LULESH_B200_REFERENCE_BINDING_H
#define LULESH_B200_REFERENCE_BINDING_H

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "lulesh_b200.h"

/* Keeps the arrays alive that the view points to but the reference does not expose as arrays. */
struct B200ViewStorage {
   std::vector<Index_t> symmX, symmY, symmZ;          /* lulesh.h:289-291 return by value */
   std::vector<const Index_t *> regLists;             /* lulesh.h:302 */
   std::vector<Index_t> nodeElemStart, nodeElemCornerList;
};

/* Domain -> lulesh_b200_host_view (the arguments of lulesh_b200_create). */
static inline lulesh_b200_host_view B200MakeView(Domain &d, int numRanks, int myRank, B200ViewStorage &st)
{
   lulesh_b200_host_view v = lulesh_b200_host_view();
   v.abi_version = LULESH_B200_ABI_VERSION;
   v.sizeX = d.sizeX(); v.sizeY = d.sizeY(); v.sizeZ = d.sizeZ();
   v.numElem = d.numElem(); v.numNode = d.numNode();
   v.numRanks = numRanks; v.rank = myRank;
   v.px = v.py = v.pz = d.tp();                                   /* lulesh-init.cc:59 */
   v.colLoc = d.colLoc(); v.rowLoc = d.rowLoc(); v.planeLoc = d.planeLoc();
   v.x = &d.x(0); v.y = &d.y(0); v.z = &d.z(0);                   /* lulesh.h:266-268 */
   v.xd = &d.xd(0); v.yd = &d.yd(0); v.zd = &d.zd(0);             /* lulesh.h:271-273 */
   v.nodalMass = &d.nodalMass(0);                                 /* lulesh.h:286 */

   const Index_t edgeNodes = d.sizeX() + 1;                       /* lulesh-init.cc:390-395, 514-533 */
   const Index_t planeNodes = edgeNodes * edgeNodes;
   if (!d.symmXempty()) for (Index_t i = 0; i < planeNodes; ++i) st.symmX.push_back(d.symmX(i));
   if (!d.symmYempty()) for (Index_t i = 0; i < planeNodes; ++i) st.symmY.push_back(d.symmY(i));
   if (!d.symmZempty()) for (Index_t i = 0; i < planeNodes; ++i) st.symmZ.push_back(d.symmZ(i));
   v.symmX = st.symmX.data(); v.numSymmX = (int32_t)st.symmX.size();
   v.symmY = st.symmY.data(); v.numSymmY = (int32_t)st.symmY.size();
   v.symmZ = st.symmZ.data(); v.numSymmZ = (int32_t)st.symmZ.size();

   v.nodelist = d.nodelist(0);                                    /* lulesh.h:305 */
   v.lxim = &d.lxim(0); v.lxip = &d.lxip(0); v.letam = &d.letam(0); v.letap = &d.letap(0);
   v.lzetam = &d.lzetam(0); v.lzetap = &d.lzetap(0); v.elemBC = &d.elemBC(0);
   v.e = &d.e(0); v.p = &d.p(0); v.q = &d.q(0); v.v = &d.v(0);
   v.volo = &d.volo(0); v.ss = &d.ss(0); v.elemMass = &d.elemMass(0);

   v.numReg = d.numReg(); v.cost = d.cost();
   v.regElemSize = &d.regElemSize(0);
   for (Int_t r = 0; r < d.numReg(); ++r) st.regLists.push_back(d.regElemlist(r));
   v.regElemlist = st.regLists.data();

   /* Node -> element-corner lists in ascending element order: the deterministic force-gather
    * order of the reference's threaded path (lulesh.cc:565-582).  The reference only builds
    * them when it runs with more than one thread (lulesh-init.cc:280) and keeps the offsets
    * private, so they are rebuilt here from nodelist exactly as lulesh-init.cc:295-319 does. */
   const Index_t numElem = d.numElem(), numNode = d.numNode();
   st.nodeElemStart.assign(numNode + 1, 0);
   for (Index_t k = 0; k < numElem; ++k) {
      const Index_t *nl = d.nodelist(k);
      for (Index_t c = 0; c < 8; ++c) ++st.nodeElemStart[nl[c] + 1];
   }
   for (Index_t n = 0; n < numNode; ++n) st.nodeElemStart[n + 1] += st.nodeElemStart[n];
   st.nodeElemCornerList.resize(st.nodeElemStart[numNode]);
   std::vector<Index_t> fill(st.nodeElemStart.begin(), st.nodeElemStart.end() - 1);
   for (Index_t k = 0; k < numElem; ++k) {
      const Index_t *nl = d.nodelist(k);
      for (Index_t c = 0; c < 8; ++c) st.nodeElemCornerList[fill[nl[c]]++] = k * 8 + c;
   }
   v.nodeElemStart = st.nodeElemStart.data();
   v.nodeElemCornerList = st.nodeElemCornerList.data();

   const lulesh_b200_constants c = {d.e_cut(), d.p_cut(), d.q_cut(), d.v_cut(), d.u_cut(), d.hgcoef(),
                                    d.ss4o3(), d.qstop(), d.monoq_max_slope(), d.monoq_limiter_mult(),
                                    d.qlc_monoq(), d.qqc_monoq(), d.qqc(), d.eosvmax(), d.eosvmin(),
                                    d.pmin(), d.emin(), d.dvovmax(), d.refdens()};   /* lulesh.h:378-399 */
   v.constants = c;
   const lulesh_b200_scalars s = {d.dtcourant(), d.dthydro(), d.dtfixed(), d.time(), d.deltatime(),
                                  d.deltatimemultlb(), d.deltatimemultub(), d.dtmax(), d.stoptime(),
                                  d.cycle(), 0};                                     /* lulesh.h:402-412 */
   v.scalars = s;
   return v;
}

/* -p line, lulesh.cc:2750-2756 */
static inline void B200PrintCycle(int32_t cycle, double time, double dt, void *)
{
   std::cout << "cycle = " << cycle << ", " << std::scientific << "time = " << time << ", "
             << "dt=" << dt << "\n";
   std::cout.unsetf(std::ios_base::floatfield);
}

/* Replaces the while loop of main() (lulesh.cc:2745-2757).  On return the Domain holds what the
 * loop would have left in it: the time controls and every state array. */
static inline void B200TimedLoop(Domain &d, const cmdLineOpts &opts, int myRank, int numRanks)
{
   if (numRanks != 1) {
      fprintf(stderr, "B200TimedLoop: this binding is the one-rank (USE_MPI=0) form\n");
      exit(1);
   }
   B200ViewStorage storage;
   const lulesh_b200_host_view view = B200MakeView(d, numRanks, myRank, storage);
   lulesh_b200 *gpu = NULL;
   if (lulesh_b200_create(&view, 0, NULL, &gpu) != 0) {
      fprintf(stderr, "lulesh_b200: %s\n", lulesh_b200_last_error());
      exit(1);
   }
   const bool show = (opts.showProg != 0) && (opts.quiet == 0) && (myRank == 0);
   const int rc = lulesh_b200_run(gpu, opts.its, show ? 1 : 64, show ? B200PrintCycle : NULL, NULL);
   if (rc == LULESH_B200_VOLUME_ERROR) exit(VolumeError);          /* lulesh.h:42, lulesh.cc:1038 */
   if (rc == LULESH_B200_QSTOP_ERROR) exit(QStopError);            /* lulesh.cc:2007 */
   if (rc != 0) {
      fprintf(stderr, "lulesh_b200: %s\n", lulesh_b200_last_error());
      exit(1);
   }
   lulesh_b200_scalars s;
   lulesh_b200_get_scalars(gpu, &s);
   d.time() = s.time; d.deltatime() = s.deltatime; d.cycle() = s.cycle;
   d.dtcourant() = s.dtcourant; d.dthydro() = s.dthydro;
   const struct { int field; Real_t *dst; size_t n; } back[] = {
      {LULESH_F_X, &d.x(0), (size_t)d.numNode()},   {LULESH_F_Y, &d.y(0), (size_t)d.numNode()},
      {LULESH_F_Z, &d.z(0), (size_t)d.numNode()},   {LULESH_F_XD, &d.xd(0), (size_t)d.numNode()},
      {LULESH_F_YD, &d.yd(0), (size_t)d.numNode()}, {LULESH_F_ZD, &d.zd(0), (size_t)d.numNode()},
      {LULESH_F_E, &d.e(0), (size_t)d.numElem()},   {LULESH_F_P, &d.p(0), (size_t)d.numElem()},
      {LULESH_F_Q, &d.q(0), (size_t)d.numElem()},   {LULESH_F_V, &d.v(0), (size_t)d.numElem()},
      {LULESH_F_SS, &d.ss(0), (size_t)d.numElem()}};
   for (size_t i = 0; i < sizeof back / sizeof back[0]; ++i)
      if (lulesh_b200_download(gpu, back[i].field, back[i].dst, back[i].n) != 0) {
         fprintf(stderr, "lulesh_b200: %s\n", lulesh_b200_last_error());
         exit(1);
      }
   lulesh_b200_destroy(gpu);
}

#endif /* LULESH_B200_REFERENCE_BINDING_H */
