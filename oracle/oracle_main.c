/* TEST INFRASTRUCTURE ONLY -- command-line front end of the CPU restatement.
 * Same flags as the reference (-s -i -r -b -c), plus --decomp PXxPYxPZ to run
 * the in-process multi-rank emulation.  Prints one JSON line with the same
 * keys as oracle/ref_util_wrap.cc so the two can be diffed directly. */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "lulesh_oracle.h"

static double now(void)
{
   struct timespec t;
   clock_gettime(CLOCK_MONOTONIC, &t);
   return t.tv_sec + 1e-9 * t.tv_nsec;
}

static double ksum(const double *a, size_t n)
{
   double s = 0.0;
   for (size_t i = 0; i < n; ++i) s += a[i];
   return s;
}

int main(int argc, char **argv)
{
   int nx = 30, its = 9999999, nr = 11, balance = 1, cost = 1, px = 1, py = 1, pz = 1, refdt0 = 0;
   for (int i = 1; i < argc; ++i) {
      if (!strcmp(argv[i], "-s") && i + 1 < argc) nx = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-i") && i + 1 < argc) its = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-r") && i + 1 < argc) nr = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-b") && i + 1 < argc) balance = atoi(argv[++i]);
      else if (!strcmp(argv[i], "-c") && i + 1 < argc) cost = atoi(argv[++i]);
      else if (!strcmp(argv[i], "--decomp") && i + 1 < argc) sscanf(argv[++i], "%dx%dx%d", &px, &py, &pz);
      else if (!strcmp(argv[i], "--reference-dt0")) refdt0 = 1;
      else if (!strcmp(argv[i], "-q")) {}
      else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 255; }
   }
   ora_multi *m = ora_multi_new(px, py, pz, nx, nx, nx, nr, balance, cost);
   if (!m) { fprintf(stderr, "bad configuration\n"); return 255; }
   if (refdt0)
      for (int r = 0; r < px * py * pz; ++r) ora_use_reference_dt0(ora_multi_rank(m, r));
   double t0 = now();
   int rc = ora_multi_run(m, its);
   double el = now() - t0;
   if (rc) { fprintf(stderr, "abort %d\n", rc); return rc == -1 ? 255 : 254; }
   ora_domain *d = ora_multi_rank(m, 0);
   lulesh_b200_scalars *s = ora_scalars(d);
   size_t ne = ora_real_count(d, LULESH_F_E), nn = ora_real_count(d, LULESH_F_X);
   double sym[3];
   ora_symmetry(d, nx, sym);
   double sxyz = 0, svel = 0;
   for (size_t i = 0; i < nn; ++i) {
      sxyz += ora_real(d, LULESH_F_X)[i] + ora_real(d, LULESH_F_Y)[i] + ora_real(d, LULESH_F_Z)[i];
      svel += fabs(ora_real(d, LULESH_F_XD)[i]) + fabs(ora_real(d, LULESH_F_YD)[i]) +
              fabs(ora_real(d, LULESH_F_ZD)[i]);
   }
   int nranks = px * py * pz;
   double zc = (double)ne * nranks * s->cycle;
   printf("ORAJSON {\"nx\": %d, \"numRanks\": %d, \"cycles\": %d, \"e0\": %.17g, \"time\": %.17g, "
          "\"dt\": %.17g, \"dtcourant\": %.17g, \"dthydro\": %.17g, \"sum_e\": %.17g, "
          "\"sum_p\": %.17g, \"sum_q\": %.17g, \"sum_v\": %.17g, \"sum_ss\": %.17g, "
          "\"sum_xyz\": %.17g, \"sum_absvel\": %.17g, \"max_abs_diff\": %.17g, "
          "\"total_abs_diff\": %.17g, \"max_rel_diff\": %.17g, \"elapsed\": %.9g, "
          "\"zones_per_s\": %.9g, \"regions\": [",
          nx, nranks, s->cycle, ora_real(d, LULESH_F_E)[0], s->time, s->deltatime, s->dtcourant,
          s->dthydro, ksum(ora_real(d, LULESH_F_E), ne), ksum(ora_real(d, LULESH_F_P), ne),
          ksum(ora_real(d, LULESH_F_Q), ne), ksum(ora_real(d, LULESH_F_V), ne),
          ksum(ora_real(d, LULESH_F_SS), ne), sxyz, svel, sym[0], sym[1], sym[2], el, zc / el);
   int cnt;
   int *sizes = ora_int(d, "regElemSize", &cnt);
   for (int r = 0; r < cnt; ++r) printf("%s%d", r ? ", " : "", sizes[r]);
   printf("]}\n");
   ora_multi_free(m);
   return 0;
}
