// TEST INFRASTRUCTURE ONLY.  CPU check of the reference-side binding
// (include/lulesh_b200_reference_binding.h): builds the reference's own Domain
// (lulesh-init.cc, unmodified, from /root/reference), lets B200MakeView() turn it into a
// lulesh_b200_host_view, and verifies every pointer / size / scalar of the view against the
// reference's public accessors -- in particular the node -> element-corner lists the binding
// rebuilds from nodelist against the lists the reference itself builds when it runs threaded
// (SetupThreadSupportStructures, lulesh-init.cc:272-337; run with OMP_NUM_THREADS >= 2).
// No GPU call is made.  Exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if _OPENMP
#include <omp.h>
#endif

#include "lulesh.h"
#include "lulesh_b200_reference_binding.h"

#define CHECK(cond)                                                          \
   do {                                                                      \
      if (!(cond)) { printf("BINDING_CHECK FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } \
   } while (0)

int main(int argc, char **argv)
{
   const int nx = argc > 1 ? atoi(argv[1]) : 6;
   const int numReg = argc > 2 ? atoi(argv[2]) : 11, balance = argc > 3 ? atoi(argv[3]) : 1;
   const int cost = argc > 4 ? atoi(argv[4]) : 1;
   Domain d(1, 0, 0, 0, nx, 1, numReg, balance, cost);    // lulesh.cc:2715 with one rank
   B200ViewStorage st;
   const lulesh_b200_host_view v = B200MakeView(d, 1, 0, st);

   CHECK(v.abi_version == LULESH_B200_ABI_VERSION);
   CHECK(v.numElem == nx * nx * nx && v.numNode == (nx + 1) * (nx + 1) * (nx + 1));
   CHECK(v.sizeX == nx && v.sizeY == nx && v.sizeZ == nx && v.px == 1 && v.numRanks == 1 && v.rank == 0);
   CHECK(v.x == &d.x(0) && v.zd == &d.zd(0) && v.nodalMass == &d.nodalMass(0));
   CHECK(v.e == &d.e(0) && v.ss == &d.ss(0) && v.elemMass == &d.elemMass(0) && v.nodelist == d.nodelist(0));
   CHECK(v.numSymmX == (nx + 1) * (nx + 1) && v.numSymmY == v.numSymmX && v.numSymmZ == v.numSymmX);
   for (int i = 0; i < v.numSymmX; ++i)
      CHECK(v.symmX[i] == d.symmX(i) && v.symmY[i] == d.symmY(i) && v.symmZ[i] == d.symmZ(i));
   CHECK(v.numReg == d.numReg() && v.cost == d.cost());
   long long total = 0;
   for (int r = 0; r < v.numReg; ++r) {
      CHECK(v.regElemSize[r] == d.regElemSize(r) && v.regElemlist[r] == d.regElemlist(r));
      total += v.regElemSize[r];
   }
   CHECK(total == v.numElem);
   CHECK(v.scalars.deltatime == d.deltatime() && v.scalars.stoptime == d.stoptime() && v.scalars.cycle == d.cycle());
   CHECK(v.constants.hgcoef == d.hgcoef() && v.constants.refdens == d.refdens() && v.constants.qstop == d.qstop());

   // node -> corner lists: ours (rebuilt from nodelist) against the reference's own
   CHECK(v.nodeElemStart[0] == 0 && v.nodeElemStart[v.numNode] == 8 * v.numElem);
   int compared = 0;
#if _OPENMP
   if (omp_get_max_threads() > 1) {
      for (Index_t n = 0; n < d.numNode(); ++n) {
         const Index_t cnt = d.nodeElemCount(n);
         CHECK(cnt == v.nodeElemStart[n + 1] - v.nodeElemStart[n]);
         const Index_t *ref = d.nodeElemCornerList(n);
         for (Index_t k = 0; k < cnt; ++k) CHECK(ref[k] == v.nodeElemCornerList[v.nodeElemStart[n] + k]);
         compared += cnt;
      }
   }
#endif
   printf("BINDING_CHECK ok nx=%d regions=%d corner_entries_compared=%d\n", nx, numReg, compared);
   return 0;
}
