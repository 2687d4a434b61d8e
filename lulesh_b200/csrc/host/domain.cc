// Host Domain setup: mesh, node->corner CSR, region index sets, symmetry node
// sets, face connectivity + boundary-condition masks, initial field data.
// Behaviour follows lulesh-init.cc:16-673; structure is written around (i,j,k)
// loops over a general (sx,sy,sz) brick of a (px,py,pz) rank grid.
#include "domain.h"

#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace {

// elemBC bits, lulesh.h:59-87
enum : int {
   XI_M_SYMM = 0x00001, XI_M_COMM = 0x00004, XI_P_FREE = 0x00010, XI_P_COMM = 0x00020,
   ETA_M_SYMM = 0x00040, ETA_M_COMM = 0x00100, ETA_P_FREE = 0x00400, ETA_P_COMM = 0x00800,
   ZETA_M_SYMM = 0x01000, ZETA_M_COMM = 0x04000, ZETA_P_FREE = 0x10000, ZETA_P_COMM = 0x20000
};

inline Real_t det3(Real_t a1, Real_t a2, Real_t a3, Real_t b1, Real_t b2, Real_t b3, Real_t c1,
                   Real_t c2, Real_t c3)
{
   // lulesh.cc:1337-1338 with its nine positional arguments
   return a1 * (b2 * c3 - b3 * c2) + b1 * (a3 * c2 - a2 * c3) + c1 * (a2 * b3 - a3 * b2);
}

inline Index_t longestEdge(Index_t a, Index_t b, Index_t c) { return std::max(a, std::max(b, c)); }

}  // namespace

Real_t CalcElemVolume(const Real_t x[8], const Real_t y[8], const Real_t z[8])
{
   auto d = [](const Real_t *a, int i, int j) { return a[i] - a[j]; };
   Real_t volume =
      det3(d(x, 3, 1) + d(x, 7, 2), d(x, 6, 3), d(x, 2, 0), d(y, 3, 1) + d(y, 7, 2), d(y, 6, 3),
           d(y, 2, 0), d(z, 3, 1) + d(z, 7, 2), d(z, 6, 3), d(z, 2, 0)) +
      det3(d(x, 4, 3) + d(x, 5, 7), d(x, 6, 4), d(x, 7, 0), d(y, 4, 3) + d(y, 5, 7), d(y, 6, 4),
           d(y, 7, 0), d(z, 4, 3) + d(z, 5, 7), d(z, 6, 4), d(z, 7, 0)) +
      det3(d(x, 1, 4) + d(x, 2, 5), d(x, 6, 1), d(x, 5, 0), d(y, 1, 4) + d(y, 2, 5), d(y, 6, 1),
           d(y, 5, 0), d(z, 1, 4) + d(z, 2, 5), d(z, 6, 1), d(z, 5, 0));
   return volume * (Real_t(1.0) / Real_t(12.0));
}

Domain::Domain(Int_t numRanks, Index_t colLoc, Index_t rowLoc, Index_t planeLoc, Index_t nx,
               Int_t tp, Int_t nr, Int_t balance, Int_t cost)
   : m_cost(cost), m_numRanks(numRanks), m_rank(planeLoc * tp * tp + rowLoc * tp + colLoc),
     m_px(tp), m_py(tp), m_pz(tp), m_colLoc(colLoc), m_rowLoc(rowLoc), m_planeLoc(planeLoc),
     m_sizeX(nx), m_sizeY(nx), m_sizeZ(nx)
{
   Init(nr, balance);
}

Domain::Domain(Int_t numRanks, Int_t rank, Int_t px, Int_t py, Int_t pz, Index_t sx, Index_t sy,
               Index_t sz, Int_t nr, Int_t balance, Int_t cost)
   : m_cost(cost), m_numRanks(numRanks), m_rank(rank), m_px(px), m_py(py), m_pz(pz),
     m_colLoc(rank % px), m_rowLoc((rank / px) % py), m_planeLoc(rank / (px * py)),
     m_sizeX(sx), m_sizeY(sy), m_sizeZ(sz)
{
   Init(nr, balance);
}

void Domain::Init(Int_t nr, Int_t balance)
{
   if (m_px < 1 || m_py < 1 || m_pz < 1 || m_px * m_py * m_pz != m_numRanks || m_rank < 0 ||
       m_rank >= m_numRanks || m_sizeX < 1 || m_sizeY < 1 || m_sizeZ < 1 || nr < 1)
      throw std::invalid_argument("Domain: inconsistent layout");
   const long long ne = (long long)m_sizeX * m_sizeY * m_sizeZ;
   const long long nn = (long long)(m_sizeX + 1) * (m_sizeY + 1) * (m_sizeZ + 1);
   if (8 * ne > INT_MAX) throw std::invalid_argument("Domain: brick too large for int32 indices");
   m_numElem = (Index_t)ne;
   m_numNode = (Index_t)nn;

   Real_t einit_unused;
   SedovInitialScalars(longestEdge(m_px * m_sizeX, m_py * m_sizeY, m_pz * m_sizeZ), &m_s, &m_c, &einit_unused);

   // AllocateElemPersistent / AllocateNodePersistent + basic field init
   // (lulesh.h:164-219, lulesh-init.cc:90-116): e = p = q = ss = 0, v = 1, velocities 0
   m_e.assign(ne, 0.0); m_p.assign(ne, 0.0); m_q.assign(ne, 0.0); m_ss.assign(ne, 0.0);
   m_v.assign(ne, 1.0); m_volo.assign(ne, 0.0); m_elemMass.assign(ne, 0.0);
   m_x.assign(nn, 0.0); m_y.assign(nn, 0.0); m_z.assign(nn, 0.0);
   m_xd.assign(nn, 0.0); m_yd.assign(nn, 0.0); m_zd.assign(nn, 0.0);
   m_nodalMass.assign(nn, 0.0);

   BuildMesh();
   SetupThreadSupportStructures();
   CreateRegionIndexSets(nr, balance);
   SetupSymmetryPlanes();
   SetupElementConnectivitiesAndBCs();

   InitializeFieldData();
}

// lulesh-init.cc:218-267.  Coordinates are 1.125*i/G with G the longest global
// edge in elements (= tp*nx for the reference's cubic layouts); each one is a
// single multiply and divide, never an accumulation.
void Domain::BuildMesh()
{
   const Index_t nx1 = m_sizeX + 1, ny1 = m_sizeY + 1, nz1 = m_sizeZ + 1;
   const Index_t G = longestEdge(m_px * m_sizeX, m_py * m_sizeY, m_pz * m_sizeZ);
   Index_t n = 0;
   for (Index_t k = 0; k < nz1; ++k)
      for (Index_t j = 0; j < ny1; ++j)
         for (Index_t i = 0; i < nx1; ++i, ++n) {
            m_x[n] = Real_t(1.125) * Real_t(m_colLoc * m_sizeX + i) / Real_t(G);
            m_y[n] = Real_t(1.125) * Real_t(m_rowLoc * m_sizeY + j) / Real_t(G);
            m_z[n] = Real_t(1.125) * Real_t(m_planeLoc * m_sizeZ + k) / Real_t(G);
         }
   m_nodelist.resize(8 * (size_t)m_numElem);
   Index_t e = 0;
   for (Index_t k = 0; k < m_sizeZ; ++k)
      for (Index_t j = 0; j < m_sizeY; ++j)
         for (Index_t i = 0; i < m_sizeX; ++i, ++e) {
            const Index_t base = k * nx1 * ny1 + j * nx1 + i;
            Index_t *nl = nodelist(e);
            nl[0] = base; nl[1] = base + 1; nl[2] = base + nx1 + 1; nl[3] = base + nx1;
            for (int c = 0; c < 4; ++c) nl[4 + c] = nl[c] + nx1 * ny1;
         }
}

// lulesh-init.cc:272-337.  The reference builds this CSR only for threaded
// runs; here it is always built because it defines the deterministic order in
// which corner forces are summed into a node (ascending element index).
void Domain::SetupThreadSupportStructures()
{
   m_nodeElemStart.assign((size_t)m_numNode + 1, 0);
   for (size_t t = 0; t < m_nodelist.size(); ++t) ++m_nodeElemStart[m_nodelist[t] + 1];
   for (Index_t n = 0; n < m_numNode; ++n) m_nodeElemStart[n + 1] += m_nodeElemStart[n];
   m_nodeElemCornerList.resize(m_nodelist.size());
   std::vector<Index_t> fill(m_numNode, 0);
   for (size_t t = 0; t < m_nodelist.size(); ++t) {
      const Index_t n = m_nodelist[t];
      m_nodeElemCornerList[m_nodeElemStart[n] + fill[n]++] = (Index_t)t;   // t == elem*8 + corner
   }
}

// lulesh-init.cc:401-510: weighted random runs of elements per region, drawn
// from glibc rand() seeded with the rank; region ids rotate with the rank.
void CreateRegionIndexSetsHost(Int_t rank, Index_t numElem, Int_t nr, Int_t balance,
                               std::vector<Index_t> &regNumList, std::vector<Index_t> &regElemSize,
                               std::vector<std::vector<Index_t>> &regElemlist)
{
   srand(rank);
   regElemSize.assign(nr, 0);
   regNumList.assign(numElem, 0);
   Index_t next = 0;
   if (nr == 1) {
      std::fill(regNumList.begin(), regNumList.end(), 1);
   } else {
      std::vector<Int_t> binEnd(nr);
      Int_t costDenominator = 0, lastReg = -1;
      for (Int_t i = 0; i < nr; ++i) {
         costDenominator += pow((i + 1), balance);
         binEnd[i] = costDenominator;
      }
      auto draw = [&]() {
         const Int_t var = rand() % costDenominator;
         Int_t i = 0;
         while (var >= binEnd[i]) ++i;
         return ((i + rank) % nr) + 1;
      };
      while (next < numElem) {
         Int_t regionNum = draw();
         while (regionNum == lastReg) regionNum = draw();
         const Int_t bin = rand() % 1000;
         Index_t run;
         if (bin < 773) run = rand() % 15 + 1;
         else if (bin < 937) run = rand() % 16 + 16;
         else if (bin < 970) run = rand() % 32 + 32;
         else if (bin < 974) run = rand() % 64 + 64;
         else if (bin < 978) run = rand() % 128 + 128;
         else if (bin < 981) run = rand() % 256 + 256;
         else run = rand() % 1537 + 512;
         const Index_t stop = std::min<long long>((long long)next + run, numElem);
         while (next < stop) regNumList[next++] = regionNum;
         lastReg = regionNum;
      }
   }
   for (Index_t i = 0; i < numElem; ++i) ++regElemSize[regNumList[i] - 1];
   regElemlist.assign(nr, {});
   for (Int_t r = 0; r < nr; ++r) regElemlist[r].reserve(regElemSize[r]);
   for (Index_t i = 0; i < numElem; ++i) regElemlist[regNumList[i] - 1].push_back(i);
}

void Domain::CreateRegionIndexSets(Int_t nr, Int_t balance)
{
   m_numReg = nr;
   CreateRegionIndexSetsHost(m_rank, m_numElem, nr, balance, m_regNumList, m_regElemSize, m_regElemlist);
   m_regElemlistPtrs.resize(nr);
   for (Int_t r = 0; r < nr; ++r) m_regElemlistPtrs[r] = m_regElemlist[r].data();
}

// constants (lulesh-init.cc:20-38), time controls (146-156), deposited energy (183-185) and
// the initial time step (192) from the volume of the GLOBAL origin element
void SedovInitialScalars(Index_t G, lulesh_b200_scalars *s, lulesh_b200_constants *c, Real_t *einit)
{
   c->e_cut = 1.0e-7; c->p_cut = 1.0e-7; c->q_cut = 1.0e-7; c->v_cut = 1.0e-10; c->u_cut = 1.0e-7;
   c->hgcoef = 3.0; c->ss4o3 = 4.0 / 3.0; c->qstop = 1.0e+12; c->monoq_max_slope = 1.0;
   c->monoq_limiter_mult = 2.0; c->qlc_monoq = 0.5; c->qqc_monoq = 2.0 / 3.0; c->qqc = 2.0;
   c->eosvmax = 1.0e+9; c->eosvmin = 1.0e-9; c->pmin = 0.; c->emin = -1.0e+15;
   c->dvovmax = 0.1; c->refdens = 1.0;
   s->dtfixed = -1.0e-6; s->stoptime = 1.0e-2;
   s->deltatimemultlb = 1.1; s->deltatimemultub = 1.2;
   s->dtcourant = 1.0e+20; s->dthydro = 1.0e+20; s->dtmax = 1.0e-2;
   s->time = 0.; s->cycle = 0; s->error = 0;
   const Real_t scale = Real_t(G) / Real_t(45.0);
   *einit = Real_t(3.948746e+7) * scale * scale * scale;
   Real_t xo[8], yo[8], zo[8];
   for (int k = 0; k < 8; ++k) {
      const int i = (k == 1 || k == 2 || k == 5 || k == 6);
      const int j = (k == 2 || k == 3 || k == 6 || k == 7);
      const int l = (k >= 4);
      xo[k] = Real_t(1.125) * Real_t(i) / Real_t(G);
      yo[k] = Real_t(1.125) * Real_t(j) / Real_t(G);
      zo[k] = Real_t(1.125) * Real_t(l) / Real_t(G);
   }
   s->deltatime = (Real_t(.5) * cbrt(CalcElemVolume(xo, yo, zo))) / sqrt(Real_t(2.0) * (*einit));
}

// lulesh-init.cc:514-533; a set exists only on ranks touching the global min plane
void Domain::SetupSymmetryPlanes()
{
   const Index_t nx1 = m_sizeX + 1, ny1 = m_sizeY + 1, nz1 = m_sizeZ + 1;
   if (m_planeLoc == 0)
      for (Index_t j = 0; j < ny1; ++j)
         for (Index_t i = 0; i < nx1; ++i) m_symmZ.push_back(j * nx1 + i);
   if (m_rowLoc == 0)
      for (Index_t k = 0; k < nz1; ++k)
         for (Index_t i = 0; i < nx1; ++i) m_symmY.push_back(k * nx1 * ny1 + i);
   if (m_colLoc == 0)
      for (Index_t k = 0; k < nz1; ++k)
         for (Index_t j = 0; j < ny1; ++j) m_symmX.push_back(k * nx1 * ny1 + j * nx1);
}

// lulesh-init.cc:539-673: face neighbours (self at brick ends), then per face:
// SYMM on the global min planes, FREE on the global max planes, COMM with the
// neighbour index redirected into the ghost block otherwise.
void Domain::SetupElementConnectivitiesAndBCs()
{
   const Index_t sx = m_sizeX, sy = m_sizeY, sz = m_sizeZ, ne = m_numElem;
   m_lxim.resize(ne); m_lxip.resize(ne); m_letam.resize(ne); m_letap.resize(ne);
   m_lzetam.resize(ne); m_lzetap.resize(ne); m_elemBC.assign(ne, 0);

   Index_t ghost[6], next = ne;   // lulesh-init.cc:582-610
   const bool has[6] = {m_planeLoc != 0, m_planeLoc != m_pz - 1, m_rowLoc != 0,
                        m_rowLoc != m_py - 1, m_colLoc != 0, m_colLoc != m_px - 1};
   const Index_t faceSize[6] = {sx * sy, sx * sy, sx * sz, sx * sz, sy * sz, sy * sz};
   for (int f = 0; f < 6; ++f) {
      ghost[f] = has[f] ? next : INT_MIN;
      if (has[f]) next += faceSize[f];
   }

   Index_t e = 0;
   for (Index_t k = 0; k < sz; ++k)
      for (Index_t j = 0; j < sy; ++j)
         for (Index_t i = 0; i < sx; ++i, ++e) {
            Int_t bc = 0;
            // lulesh-init.cc:541-564: plain index arithmetic; only the very first /
            // last entries point to themselves.  Entries on a brick face "wrap" into
            // the neighbouring row/plane exactly as in the reference; they are never
            // dereferenced because the face's elemBC bit (below) takes precedence.
            m_lxim[e] = (e >= 1) ? e - 1 : e;
            m_lxip[e] = (e < ne - 1) ? e + 1 : e;
            m_letam[e] = (e >= sx) ? e - sx : e;
            m_letap[e] = (e < ne - sx) ? e + sx : e;
            m_lzetam[e] = (e >= sx * sy) ? e - sx * sy : e;
            m_lzetap[e] = (e < ne - sx * sy) ? e + sx * sy : e;
            if (k == 0) {
               if (!has[0]) bc |= ZETA_M_SYMM;
               else { bc |= ZETA_M_COMM; m_lzetam[e] = ghost[0] + j * sx + i; }
            }
            if (k == sz - 1) {
               if (!has[1]) bc |= ZETA_P_FREE;
               else { bc |= ZETA_P_COMM; m_lzetap[e] = ghost[1] + j * sx + i; }
            }
            if (j == 0) {
               if (!has[2]) bc |= ETA_M_SYMM;
               else { bc |= ETA_M_COMM; m_letam[e] = ghost[2] + k * sx + i; }
            }
            if (j == sy - 1) {
               if (!has[3]) bc |= ETA_P_FREE;
               else { bc |= ETA_P_COMM; m_letap[e] = ghost[3] + k * sx + i; }
            }
            if (i == 0) {
               if (!has[4]) bc |= XI_M_SYMM;
               else { bc |= XI_M_COMM; m_lxim[e] = ghost[4] + k * sy + j; }
            }
            if (i == sx - 1) {
               if (!has[5]) bc |= XI_P_FREE;
               else { bc |= XI_P_COMM; m_lxip[e] = ghost[5] + k * sy + j; }
            }
            m_elemBC[e] = bc;
         }
}

// lulesh-init.cc:159-192: reference volumes and masses, energy deposit in the
// origin element, initial time step.
void Domain::InitializeFieldData()
{
   for (Index_t i = 0; i < m_numElem; ++i) {
      Real_t xl[8], yl[8], zl[8];
      const Index_t *nl = nodelist(i);
      for (int c = 0; c < 8; ++c) { xl[c] = m_x[nl[c]]; yl[c] = m_y[nl[c]]; zl[c] = m_z[nl[c]]; }
      const Real_t volume = CalcElemVolume(xl, yl, zl);
      m_volo[i] = volume;
      m_elemMass[i] = volume;
      for (int c = 0; c < 8; ++c) m_nodalMass[nl[c]] += volume / Real_t(8.0);
   }
   // energy deposit (lulesh-init.cc:183-190); deltatime was set by SedovInitialScalars from the
   // GLOBAL origin element: the reference's per-rank volo(0) (lulesh-init.cc:192) is not
   // bit-identical across ranks for sizes such as 640, and on the origin rank the
   // two coincide.
   lulesh_b200_scalars s_unused;
   lulesh_b200_constants c_unused;
   Real_t einit;
   SedovInitialScalars(longestEdge(m_px * m_sizeX, m_py * m_sizeY, m_pz * m_sizeZ), &s_unused, &c_unused, &einit);
   if (m_rowLoc + m_colLoc + m_planeLoc == 0) m_e[0] = einit;
}

lulesh_b200_host_view Domain::view()
{
   lulesh_b200_host_view v;
   memset(&v, 0, sizeof v);
   v.abi_version = LULESH_B200_ABI_VERSION;
   v.sizeX = m_sizeX; v.sizeY = m_sizeY; v.sizeZ = m_sizeZ;
   v.numElem = m_numElem; v.numNode = m_numNode;
   v.numRanks = m_numRanks; v.rank = m_rank;
   v.px = m_px; v.py = m_py; v.pz = m_pz;
   v.colLoc = m_colLoc; v.rowLoc = m_rowLoc; v.planeLoc = m_planeLoc;
   v.x = m_x.data(); v.y = m_y.data(); v.z = m_z.data();
   v.xd = m_xd.data(); v.yd = m_yd.data(); v.zd = m_zd.data();
   v.nodalMass = m_nodalMass.data();
   v.symmX = m_symmX.empty() ? nullptr : m_symmX.data(); v.numSymmX = (int32_t)m_symmX.size();
   v.symmY = m_symmY.empty() ? nullptr : m_symmY.data(); v.numSymmY = (int32_t)m_symmY.size();
   v.symmZ = m_symmZ.empty() ? nullptr : m_symmZ.data(); v.numSymmZ = (int32_t)m_symmZ.size();
   v.nodelist = m_nodelist.data();
   v.lxim = m_lxim.data(); v.lxip = m_lxip.data(); v.letam = m_letam.data();
   v.letap = m_letap.data(); v.lzetam = m_lzetam.data(); v.lzetap = m_lzetap.data();
   v.elemBC = m_elemBC.data();
   v.e = m_e.data(); v.p = m_p.data(); v.q = m_q.data(); v.v = m_v.data();
   v.volo = m_volo.data(); v.ss = m_ss.data(); v.elemMass = m_elemMass.data();
   v.numReg = m_numReg; v.cost = m_cost;
   v.regElemSize = m_regElemSize.data();
   v.regElemlist = m_regElemlistPtrs.data();
   v.nodeElemStart = m_nodeElemStart.data();
   v.nodeElemCornerList = m_nodeElemCornerList.data();
   v.constants = m_c;
   v.scalars = m_s;
   return v;
}

std::vector<Real_t> *Domain::realField(int field)
{
   switch (field) {
      case LULESH_F_X: return &m_x;   case LULESH_F_Y: return &m_y;   case LULESH_F_Z: return &m_z;
      case LULESH_F_XD: return &m_xd; case LULESH_F_YD: return &m_yd; case LULESH_F_ZD: return &m_zd;
      case LULESH_F_NODALMASS: return &m_nodalMass;
      case LULESH_F_E: return &m_e;   case LULESH_F_P: return &m_p;   case LULESH_F_Q: return &m_q;
      case LULESH_F_V: return &m_v;   case LULESH_F_VOLO: return &m_volo;
      case LULESH_F_SS: return &m_ss; case LULESH_F_ELEMMASS: return &m_elemMass;
      default: return nullptr;
   }
}

const std::vector<Index_t> *Domain::intField(const char *name)
{
   struct Entry { const char *name; const std::vector<Index_t> *v; };
   const Entry table[] = {
      {"nodelist", &m_nodelist}, {"lxim", &m_lxim}, {"lxip", &m_lxip}, {"letam", &m_letam},
      {"letap", &m_letap}, {"lzetam", &m_lzetam}, {"lzetap", &m_lzetap}, {"elemBC", &m_elemBC},
      {"regNumList", &m_regNumList}, {"regElemSize", &m_regElemSize}, {"symmX", &m_symmX},
      {"symmY", &m_symmY}, {"symmZ", &m_symmZ}, {"nodeElemStart", &m_nodeElemStart},
      {"nodeElemCornerList", &m_nodeElemCornerList}};
   for (const Entry &t : table)
      if (!strcmp(t.name, name)) return t.v;
   return nullptr;
}
