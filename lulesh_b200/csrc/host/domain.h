// Host-side Domain: the data model of the reference (`class Domain`,
// lulesh.h:148-595) with the same accessor names, argument meaning and field
// layout (SoA double / int32), plus the setup of lulesh-init.cc generalised from
// a tp^3 grid of nx^3 cubes to a (px,py,pz) grid of (sx,sy,sz) bricks
// (SURVEY appendix C).  The device library never sees this class, only the
// lulesh_b200_host_view produced by view().
#pragma once
#include <cstdint>
#include <vector>

#include "../../../include/lulesh_b200.h"

typedef double  Real_t;    // lulesh.h:39
typedef int32_t Index_t;   // lulesh.h:38
typedef int32_t Int_t;     // lulesh.h:40

class Domain {
 public:
   // Reference signature (lulesh.h:153-155): cubic layout, tp ranks per axis.
   Domain(Int_t numRanks, Index_t colLoc, Index_t rowLoc, Index_t planeLoc, Index_t nx, Int_t tp,
          Int_t nr, Int_t balance, Int_t cost);
   // Generalised layout; `rank` seeds the region generator (lulesh-init.cc:406).
   Domain(Int_t numRanks, Int_t rank, Int_t px, Int_t py, Int_t pz, Index_t sx, Index_t sy,
          Index_t sz, Int_t nr, Int_t balance, Int_t cost);

   // ---- node-centred (lulesh.h:266-294)
   Real_t &x(Index_t i) { return m_x[i]; }
   Real_t &y(Index_t i) { return m_y[i]; }
   Real_t &z(Index_t i) { return m_z[i]; }
   Real_t &xd(Index_t i) { return m_xd[i]; }
   Real_t &yd(Index_t i) { return m_yd[i]; }
   Real_t &zd(Index_t i) { return m_zd[i]; }
   Real_t &nodalMass(Index_t i) { return m_nodalMass[i]; }
   Index_t symmX(Index_t i) { return m_symmX[i]; }
   Index_t symmY(Index_t i) { return m_symmY[i]; }
   Index_t symmZ(Index_t i) { return m_symmZ[i]; }
   bool symmXempty() { return m_symmX.empty(); }
   bool symmYempty() { return m_symmY.empty(); }
   bool symmZempty() { return m_symmZ.empty(); }

   // ---- element-centred (lulesh.h:299-367)
   Index_t &regElemSize(Index_t r) { return m_regElemSize[r]; }
   Index_t &regNumList(Index_t i) { return m_regNumList[i]; }
   Index_t *regElemlist(Int_t r) { return m_regElemlist[r].data(); }
   Index_t &regElemlist(Int_t r, Index_t i) { return m_regElemlist[r][i]; }
   Index_t *nodelist(Index_t i) { return &m_nodelist[Index_t(8) * i]; }
   Index_t &lxim(Index_t i) { return m_lxim[i]; }
   Index_t &lxip(Index_t i) { return m_lxip[i]; }
   Index_t &letam(Index_t i) { return m_letam[i]; }
   Index_t &letap(Index_t i) { return m_letap[i]; }
   Index_t &lzetam(Index_t i) { return m_lzetam[i]; }
   Index_t &lzetap(Index_t i) { return m_lzetap[i]; }
   Int_t &elemBC(Index_t i) { return m_elemBC[i]; }
   Real_t &e(Index_t i) { return m_e[i]; }
   Real_t &p(Index_t i) { return m_p[i]; }
   Real_t &q(Index_t i) { return m_q[i]; }
   Real_t &v(Index_t i) { return m_v[i]; }
   Real_t &volo(Index_t i) { return m_volo[i]; }
   Real_t &ss(Index_t i) { return m_ss[i]; }
   Real_t &elemMass(Index_t i) { return m_elemMass[i]; }
   Index_t nodeElemCount(Index_t i) { return m_nodeElemStart[i + 1] - m_nodeElemStart[i]; }
   Index_t *nodeElemCornerList(Index_t i) { return &m_nodeElemCornerList[m_nodeElemStart[i]]; }

   // ---- parameters (lulesh.h:378-399)
   Real_t u_cut() const { return m_c.u_cut; }
   Real_t e_cut() const { return m_c.e_cut; }
   Real_t p_cut() const { return m_c.p_cut; }
   Real_t q_cut() const { return m_c.q_cut; }
   Real_t v_cut() const { return m_c.v_cut; }
   Real_t hgcoef() const { return m_c.hgcoef; }
   Real_t qstop() const { return m_c.qstop; }
   Real_t monoq_max_slope() const { return m_c.monoq_max_slope; }
   Real_t monoq_limiter_mult() const { return m_c.monoq_limiter_mult; }
   Real_t ss4o3() const { return m_c.ss4o3; }
   Real_t qlc_monoq() const { return m_c.qlc_monoq; }
   Real_t qqc_monoq() const { return m_c.qqc_monoq; }
   Real_t qqc() const { return m_c.qqc; }
   Real_t eosvmax() const { return m_c.eosvmax; }
   Real_t eosvmin() const { return m_c.eosvmin; }
   Real_t pmin() const { return m_c.pmin; }
   Real_t emin() const { return m_c.emin; }
   Real_t dvovmax() const { return m_c.dvovmax; }
   Real_t refdens() const { return m_c.refdens; }

   // ---- time-step controls (lulesh.h:402-412)
   Real_t &time() { return m_s.time; }
   Real_t &deltatime() { return m_s.deltatime; }
   Real_t &deltatimemultlb() { return m_s.deltatimemultlb; }
   Real_t &deltatimemultub() { return m_s.deltatimemultub; }
   Real_t &stoptime() { return m_s.stoptime; }
   Real_t &dtcourant() { return m_s.dtcourant; }
   Real_t &dthydro() { return m_s.dthydro; }
   Real_t &dtmax() { return m_s.dtmax; }
   Real_t &dtfixed() { return m_s.dtfixed; }
   Int_t &cycle() { return m_s.cycle; }

   // ---- layout (lulesh.h:413-426)
   Index_t &numRanks() { return m_numRanks; }
   Index_t &colLoc() { return m_colLoc; }
   Index_t &rowLoc() { return m_rowLoc; }
   Index_t &planeLoc() { return m_planeLoc; }
   Index_t &tp() { return m_px; }
   Index_t &sizeX() { return m_sizeX; }
   Index_t &sizeY() { return m_sizeY; }
   Index_t &sizeZ() { return m_sizeZ; }
   Index_t &numReg() { return m_numReg; }
   Int_t &cost() { return m_cost; }
   Index_t &numElem() { return m_numElem; }
   Index_t &numNode() { return m_numNode; }
   Int_t rank() const { return m_rank; }

   // read-only view for lulesh_b200_create (pointers stay valid for the Domain's lifetime)
   lulesh_b200_host_view view();
   lulesh_b200_scalars &scalars() { return m_s; }
   std::vector<Real_t> *realField(int field);
   const std::vector<Index_t> *intField(const char *name);

 private:
   void Init(Int_t nr, Int_t balance);
   void BuildMesh();
   void SetupThreadSupportStructures();
   void CreateRegionIndexSets(Int_t nreg, Int_t balance);
   void SetupSymmetryPlanes();
   void SetupElementConnectivitiesAndBCs();
   void InitializeFieldData();

   std::vector<Real_t> m_x, m_y, m_z, m_xd, m_yd, m_zd, m_nodalMass;
   std::vector<Index_t> m_symmX, m_symmY, m_symmZ;
   Int_t m_numReg = 0, m_cost = 0;
   std::vector<Index_t> m_regElemSize, m_regNumList;
   std::vector<std::vector<Index_t>> m_regElemlist;
   std::vector<const Index_t *> m_regElemlistPtrs;
   std::vector<Index_t> m_nodelist, m_lxim, m_lxip, m_letam, m_letap, m_lzetam, m_lzetap;
   std::vector<Int_t> m_elemBC;
   std::vector<Real_t> m_e, m_p, m_q, m_v, m_volo, m_ss, m_elemMass;
   std::vector<Index_t> m_nodeElemStart, m_nodeElemCornerList;
   lulesh_b200_constants m_c;
   lulesh_b200_scalars m_s;
   Int_t m_numRanks, m_rank;
   Index_t m_px, m_py, m_pz, m_colLoc, m_rowLoc, m_planeLoc;
   Index_t m_sizeX, m_sizeY, m_sizeZ, m_numElem, m_numNode;
};

// Pieces of the setup that stay on the host even with device-side mesh generation
// (lulesh_b200_create_sedov): the region index sets need glibc's sequential rand()
// (lulesh-init.cc:401-510); the initial time controls and energy are a handful of scalars.
void CreateRegionIndexSetsHost(Int_t rank, Index_t numElem, Int_t nr, Int_t balance,
                               std::vector<Index_t> &regNumList, std::vector<Index_t> &regElemSize,
                               std::vector<std::vector<Index_t>> &regElemlist);
void SedovInitialScalars(Index_t globalEdge, lulesh_b200_scalars *s, lulesh_b200_constants *c,
                         Real_t *einit);

// CalcElemVolume (lulesh.cc:1274-1366), needed by the setup for volo/elemMass.
Real_t CalcElemVolume(const Real_t x[8], const Real_t y[8], const Real_t z[8]);

// `-v` field dump (lulesh-viz.cc:56-258 with VTK in place of Silo), vizdump.cc
int DumpDomainToVTK(Domain &d, Int_t myRank, const char *path);
int WriteVisitIndex(const char *path, const char *basename, Int_t numRanks);

struct cmdLineOpts {   // lulesh.h:599-609 plus the additive multi-GPU flags
   Int_t its, nx, numReg, numFiles, showProg, quiet, viz, cost, balance;
   Int_t gpus;            // --gpus N   (ranks = GPUs of this node, one host thread each)
   Int_t px, py, pz;      // --decomp PXxPYxPZ (default from lulesh_host_decompose)
   Int_t global;          // --global G: strong scaling, local brick = G/p per axis
   Int_t syncEvery;       // --sync-every K cycles between host polls
   Int_t deviceSetup;     // --device-setup: lulesh_b200_create_sedov instead of a host Domain
};
