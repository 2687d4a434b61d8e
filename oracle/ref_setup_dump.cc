// TEST INFRASTRUCTURE ONLY.
//
// Constructs the reference's own Domain (lulesh-init.cc:16-194) for ONE rank
// location of a tp^3 layout inside a non-MPI build and dumps every setup array
// (coordinates, nodelist, face neighbours incl. ghost indices, elemBC, symmetry
// lists, volo, pre-exchange nodalMass, e, deltatime).  In a USE_MPI=0 build the
// constructor never touches MPI, so all ranks of a layout can be produced in
// one process (SURVEY 8(c) "Setup oracle for multi-rank layouts").  lulesh.cc
// is linked with -Dmain=lulesh_reference_main only to obtain CalcElemVolume.
#include <cstdio>
#include <cstdlib>
#include "lulesh.h"

void RefDumpDomain(Domain& d, Int_t nx, Int_t numRanks, const char* path);

int main(int argc, char** argv)
{
   if (argc != 10) {
      fprintf(stderr, "usage: %s tp nx col row plane numReg balance cost out.bin\n", argv[0]);
      return 2;
   }
   int tp = atoi(argv[1]), nx = atoi(argv[2]);
   int col = atoi(argv[3]), row = atoi(argv[4]), plane = atoi(argv[5]);
   int nr = atoi(argv[6]), balance = atoi(argv[7]), cost = atoi(argv[8]);
   Domain* d = new Domain(tp * tp * tp, col, row, plane, nx, tp, nr, balance, cost);
   RefDumpDomain(*d, nx, tp * tp * tp, argv[9]);
   delete d;
   return 0;
}
