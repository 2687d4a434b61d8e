// Field dump behind `-v`.  The reference writes a Silo/HDF5 file per
// group of ranks (`DumpToVisit`, lulesh-viz.cc:56-258, compiled only with -DVIZ_MESH):
// the hex mesh, the region number of every element as a material, the zone fields
// e, p, v, q and the node fields speed, xd, yd, zd.  Silo is not a dependency here:
// the same content goes into one legacy-VTK unstructured-grid file per rank (binary,
// big-endian as the format demands; LULESH's node order of a hex is VTK_HEXAHEDRON's),
// plus a `.visit` block index written by the driver, both readable by VisIt/ParaView.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../../include/lulesh_host.h"
#include "domain.h"

namespace {

template <typename T>
void put_be(std::vector<unsigned char> &buf, T v)
{
   unsigned char b[sizeof(T)];
   memcpy(b, &v, sizeof(T));
   for (size_t i = 0; i < sizeof(T); ++i) buf.push_back(b[sizeof(T) - 1 - i]);   // host is little-endian
}

bool flush(FILE *f, std::vector<unsigned char> &buf)
{
   const bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size() && fputc('\n', f) != EOF;
   buf.clear();
   return ok;
}

bool put_scalars(FILE *f, const char *name, const Real_t *a, size_t n, std::vector<unsigned char> &buf)
{
   fprintf(f, "SCALARS %s double 1\nLOOKUP_TABLE default\n", name);
   buf.reserve(n * 8);
   for (size_t i = 0; i < n; ++i) put_be(buf, a[i]);
   return flush(f, buf);
}

}  // namespace

// lulesh-viz.cc:121-258 (DumpDomainToVisit): same objects, VTK container.
int DumpDomainToVTK(Domain &d, Int_t myRank, const char *path)
{
   FILE *f = fopen(path, "wb");
   if (!f) return -1;
   const size_t nn = (size_t)d.numNode(), ne = (size_t)d.numElem();
   std::vector<unsigned char> buf;
   bool ok = true;
   fprintf(f, "# vtk DataFile Version 3.0\n");
   fprintf(f, "LULESH cycle %d time %.17g rank %d\n", d.cycle(), d.time(), myRank);
   fprintf(f, "BINARY\nDATASET UNSTRUCTURED_GRID\n");

   fprintf(f, "POINTS %zu double\n", nn);                       // lulesh-viz.cc:148-165 ("mesh")
   buf.reserve(nn * 24);
   for (size_t i = 0; i < nn; ++i) { put_be(buf, d.x(i)); put_be(buf, d.y(i)); put_be(buf, d.z(i)); }
   ok = ok && flush(f, buf);

   fprintf(f, "CELLS %zu %zu\n", ne, ne * 9);                   // lulesh-viz.cc:131-146 ("connectivity")
   buf.reserve(ne * 36);
   for (size_t k = 0; k < ne; ++k) {
      put_be<int32_t>(buf, 8);
      const Index_t *nl = d.nodelist(k);
      for (int c = 0; c < 8; ++c) put_be<int32_t>(buf, nl[c]);
   }
   ok = ok && flush(f, buf);
   fprintf(f, "CELL_TYPES %zu\n", ne);
   for (size_t k = 0; k < ne; ++k) put_be<int32_t>(buf, 12);    // VTK_HEXAHEDRON
   ok = ok && flush(f, buf);

   fprintf(f, "CELL_DATA %zu\n", ne);
   ok = ok && put_scalars(f, "e", &d.e(0), ne, buf);            // lulesh-viz.cc:183-220
   ok = ok && put_scalars(f, "p", &d.p(0), ne, buf);
   ok = ok && put_scalars(f, "v", &d.v(0), ne, buf);
   ok = ok && put_scalars(f, "q", &d.q(0), ne, buf);
   fprintf(f, "SCALARS regions int 1\nLOOKUP_TABLE default\n"); // lulesh-viz.cc:167-181 (material numbers)
   for (size_t k = 0; k < ne; ++k) put_be<int32_t>(buf, d.regNumList(k));
   ok = ok && flush(f, buf);

   fprintf(f, "POINT_DATA %zu\n", nn);
   std::vector<Real_t> speed(nn);                               // lulesh-viz.cc:222-236
   for (size_t i = 0; i < nn; ++i)
      speed[i] = sqrt(d.xd(i) * d.xd(i) + d.yd(i) * d.yd(i) + d.zd(i) * d.zd(i));
   ok = ok && put_scalars(f, "speed", speed.data(), nn, buf);
   ok = ok && put_scalars(f, "xd", &d.xd(0), nn, buf);          // lulesh-viz.cc:238-252
   ok = ok && put_scalars(f, "yd", &d.yd(0), nn, buf);
   ok = ok && put_scalars(f, "zd", &d.zd(0), nn, buf);
   ok = (fclose(f) == 0) && ok;
   return ok ? 0 : -1;
}

// Block index for VisIt (the role of DumpMultiblockObjects, lulesh-viz.cc:262-356).
int WriteVisitIndex(const char *path, const char *basename, Int_t numRanks)
{
   FILE *f = fopen(path, "w");
   if (!f) return -1;
   fprintf(f, "!NBLOCKS %d\n", numRanks);
   for (Int_t r = 0; r < numRanks; ++r) fprintf(f, "%s.%03d.vtk\n", basename, r);
   return fclose(f) == 0 ? 0 : -1;
}
