// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Full-precision reporting shim for the UNMODIFIED reference build kept in
// oracle/_ref/.  The reference prints its final origin energy with 7 digits
// (lulesh-util.cc:194-195) which is too coarse for a 1e-8 parity bar, so this
// translation unit textually includes the reference's own lulesh-util.cc from
// where it lies (the Makefile passes -I$(REF); nothing is copied into this
// repo), renames its VerifyAndWriteFinalOutput, and supplies a replacement that
// first calls the original and then also emits
//   * one "REFJSON {...}" line with %.17g scalars/checksums, and
//   * (if LULESH_REF_DUMP is set) every accessor-reachable Domain array in a
//     small self-describing binary file used to make tests/golden fixtures.
// The timed hot path (lulesh.cc) is compiled untouched.

#define VerifyAndWriteFinalOutput VerifyAndWriteFinalOutput_reference
#include "lulesh-util.cc"
#undef VerifyAndWriteFinalOutput

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {

void put(FILE* f, const char* name, const char* dtype, const void* p,
         size_t count, size_t width) {
   fprintf(f, "%s %s %zu\n", name, dtype, count);
   fwrite(p, width, count, f);
}

template <typename F>
void put_real(FILE* f, const char* name, Index_t n, F get) {
   std::vector<double> tmp(n);
   for (Index_t i = 0; i < n; ++i) tmp[i] = get(i);
   put(f, name, "f8", tmp.data(), tmp.size(), sizeof(double));
}

template <typename F>
void put_int(FILE* f, const char* name, Index_t n, F get) {
   std::vector<int> tmp(n);
   for (Index_t i = 0; i < n; ++i) tmp[i] = get(i);
   put(f, name, "i4", tmp.data(), tmp.size(), sizeof(int));
}

template <typename F>
double ksum(Index_t n, F get) {
   double s = 0.0;
   for (Index_t i = 0; i < n; ++i) s += get(i);
   return s;
}

}  // namespace

void RefDumpDomain(Domain& d, Int_t nx, Int_t numRanks, const char* path);

void VerifyAndWriteFinalOutput(Real_t elapsed_time, Domain& d, Int_t nx,
                               Int_t numRanks)
{
   VerifyAndWriteFinalOutput_reference(elapsed_time, d, nx, numRanks);

   const Index_t ne = d.numElem();
   const Index_t nn = d.numNode();

   // symmetry triple, recomputed here at full precision
   double maxAbs = 0.0, totAbs = 0.0, maxRel = 0.0;
   for (Index_t j = 0; j < nx; ++j)
      for (Index_t k = j + 1; k < nx; ++k) {
         double a = fabs(d.e(j * nx + k) - d.e(k * nx + j));
         totAbs += a;
         if (maxAbs < a) maxAbs = a;
         double r = a / d.e(k * nx + j);
         if (maxRel < r) maxRel = r;
      }

   printf("REFJSON {\"nx\": %d, \"numRanks\": %d, \"cycles\": %d, "
          "\"e0\": %.17g, \"time\": %.17g, \"dt\": %.17g, "
          "\"dtcourant\": %.17g, \"dthydro\": %.17g, "
          "\"sum_e\": %.17g, \"sum_p\": %.17g, \"sum_q\": %.17g, "
          "\"sum_v\": %.17g, \"sum_ss\": %.17g, \"sum_xyz\": %.17g, "
          "\"sum_absvel\": %.17g, \"max_abs_diff\": %.17g, "
          "\"total_abs_diff\": %.17g, \"max_rel_diff\": %.17g, "
          "\"elapsed\": %.9g, \"regions\": [",
          (int)nx, (int)numRanks, (int)d.cycle(), d.e(0), d.time(),
          d.deltatime(), d.dtcourant(), d.dthydro(),
          ksum(ne, [&](Index_t i) { return d.e(i); }),
          ksum(ne, [&](Index_t i) { return d.p(i); }),
          ksum(ne, [&](Index_t i) { return d.q(i); }),
          ksum(ne, [&](Index_t i) { return d.v(i); }),
          ksum(ne, [&](Index_t i) { return d.ss(i); }),
          ksum(nn, [&](Index_t i) { return d.x(i) + d.y(i) + d.z(i); }),
          ksum(nn, [&](Index_t i) {
             return fabs(d.xd(i)) + fabs(d.yd(i)) + fabs(d.zd(i)); }),
          maxAbs, totAbs, maxRel, (double)elapsed_time);
   for (Index_t r = 0; r < d.numReg(); ++r)
      printf("%s%d", r ? ", " : "", (int)d.regElemSize(r));
   printf("]}\n");

   const char* path = getenv("LULESH_REF_DUMP");
   if (path == NULL || *path == '\0') return;
   RefDumpDomain(d, nx, numRanks, path);
}

// Also used by ref_setup_dump.cc (multi-rank setup fixtures).
void RefDumpDomain(Domain& d, Int_t nx, Int_t numRanks, const char* path)
{
   const Index_t ne = d.numElem();
   const Index_t nn = d.numNode();
   FILE* f = fopen(path, "wb");
   if (!f) { perror("LULESH_REF_DUMP"); return; }

   int hdr[8] = {(int)nx, (int)numRanks, (int)d.cycle(), (int)ne, (int)nn,
                 (int)d.numReg(), (int)d.cost(), 0};
   put(f, "header", "i4", hdr, 8, sizeof(int));
   double sc[5] = {d.time(), d.deltatime(), d.dtcourant(), d.dthydro(),
                   d.stoptime()};
   put(f, "scalars", "f8", sc, 5, sizeof(double));

#define NODE_FIELD(n) put_real(f, #n, nn, [&](Index_t i) { return d.n(i); })
   NODE_FIELD(x); NODE_FIELD(y); NODE_FIELD(z);
   NODE_FIELD(xd); NODE_FIELD(yd); NODE_FIELD(zd);
   NODE_FIELD(xdd); NODE_FIELD(ydd); NODE_FIELD(zdd);
   NODE_FIELD(fx); NODE_FIELD(fy); NODE_FIELD(fz);
   NODE_FIELD(nodalMass);
#undef NODE_FIELD
#define ELEM_FIELD(n) put_real(f, #n, ne, [&](Index_t i) { return d.n(i); })
   ELEM_FIELD(e); ELEM_FIELD(p); ELEM_FIELD(q); ELEM_FIELD(ql); ELEM_FIELD(qq);
   ELEM_FIELD(v); ELEM_FIELD(volo); ELEM_FIELD(vnew); ELEM_FIELD(delv);
   ELEM_FIELD(vdov); ELEM_FIELD(arealg); ELEM_FIELD(ss); ELEM_FIELD(elemMass);
#undef ELEM_FIELD
#define ELEM_INT(n) put_int(f, #n, ne, [&](Index_t i) { return d.n(i); })
   ELEM_INT(lxim); ELEM_INT(lxip); ELEM_INT(letam); ELEM_INT(letap);
   ELEM_INT(lzetam); ELEM_INT(lzetap); ELEM_INT(elemBC); ELEM_INT(regNumList);
#undef ELEM_INT
   put_int(f, "nodelist", 8 * ne,
           [&](Index_t i) { return d.nodelist(i / 8)[i % 8]; });
   put_int(f, "regElemSize", d.numReg(),
           [&](Index_t r) { return d.regElemSize(r); });
   if (!d.symmXempty())
      put_int(f, "symmX", (nx + 1) * (nx + 1), [&](Index_t i) { return d.symmX(i); });
   if (!d.symmYempty())
      put_int(f, "symmY", (nx + 1) * (nx + 1), [&](Index_t i) { return d.symmY(i); });
   if (!d.symmZempty())
      put_int(f, "symmZ", (nx + 1) * (nx + 1), [&](Index_t i) { return d.symmZ(i); });
   fclose(f);
}
