// Drop-in host driver: the reference's command line, banner, timed loop and
// final report (lulesh.cc:2650-2792, lulesh-util.cc:13-230) around the C ABI of
// include/lulesh_b200.h.  MPI ranks become one host thread per GPU of this node;
// MPI_Allreduce / halo exchanges happen inside the library over NCCL.
#include <sys/time.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>
#include <thread>
#include <vector>

#include <pthread.h>

#include "../../../include/lulesh_host.h"
#include "domain.h"

namespace {

// ---- lulesh-util.cc:13-29
template <typename IntT>
int StrToInt(const char *token, IntT *retVal)
{
   if (token == NULL) return 0;
   char *endptr;
   *retVal = strtol(token, &endptr, 10);
   return (endptr != token) && ((*endptr == ' ') || (*endptr == '\0'));
}

void PrintCommandLineOptions(char *execname)
{
   printf("Usage: %s [opts]\n", execname);
   printf(" where [opts] is one or more of:\n");
   printf(" -q              : quiet mode - suppress all stdout\n");
   printf(" -i <iterations> : number of cycles to run\n");
   printf(" -s <size>       : length of cube mesh along side\n");
   printf(" -r <numregions> : Number of distinct regions (def: 11)\n");
   printf(" -b <balance>    : Load balance between regions of a domain (def: 1)\n");
   printf(" -c <cost>       : Extra cost of more expensive regions (def: 1)\n");
   printf(" -f <numfiles>   : Number of files to split viz dump into (def: (np+10)/9)\n");
   printf(" -p              : Print out progress\n");
   printf(" -v              : Output viz file (VTK, one per rank; no -DVIZ_MESH/Silo needed)\n");
   printf(" -h              : This message\n");
   printf(" --gpus <n>      : B200 extension: ranks = GPUs of this node (def: 1)\n");
   printf(" --decomp AxBxC  : B200 extension: ranks per axis (col x row x plane)\n");
   printf(" --global <size> : B200 extension: -s is the GLOBAL edge, split over the ranks\n");
   printf(" --sync-every <k>: B200 extension: cycles enqueued between host polls (def: 64)\n");
   printf(" --device-setup  : B200 extension: generate the mesh in HBM instead of on the host\n");
   printf("\n\n");
}

[[noreturn]] void ParseError(const char *message)
{
   printf("%s\n", message);
   exit(-1);   // lulesh-util.cc:51-61
}

struct IntFlag { const char *flag; Int_t cmdLineOpts::*field; };

void ParseCommandLineOptions(int argc, char *argv[], cmdLineOpts *opts)
{
   static const IntFlag intFlags[] = {
      {"-i", &cmdLineOpts::its},       {"-s", &cmdLineOpts::nx},   {"-r", &cmdLineOpts::numReg},
      {"-f", &cmdLineOpts::numFiles},  {"-b", &cmdLineOpts::balance}, {"-c", &cmdLineOpts::cost},
      {"--gpus", &cmdLineOpts::gpus},  {"--global", &cmdLineOpts::global},
      {"--sync-every", &cmdLineOpts::syncEvery}};
   for (int i = 1; i < argc;) {
      const IntFlag *hit = nullptr;
      for (const IntFlag &f : intFlags)
         if (strcmp(argv[i], f.flag) == 0) hit = &f;
      if (hit) {
         char msg[160];
         if (i + 1 >= argc) {
            // the reference's -i message has no trailing newline, the others do
            snprintf(msg, sizeof msg, "Missing integer argument to %s%s", hit->flag,
                     strcmp(hit->flag, "-i") ? "\n" : "");
            ParseError(msg);
         }
         if (!StrToInt(argv[i + 1], &(opts->*(hit->field)))) {
            snprintf(msg, sizeof msg,
                     "Parse Error on option %s integer value required after argument\n", hit->flag);
            ParseError(msg);
         }
         i += 2;
      } else if (strcmp(argv[i], "--decomp") == 0) {
         if (i + 1 >= argc || sscanf(argv[i + 1], "%dx%dx%d", &opts->px, &opts->py, &opts->pz) != 3)
            ParseError("Parse Error on option --decomp AxBxC required after argument\n");
         i += 2;
      } else if (strcmp(argv[i], "--device-setup") == 0) { opts->deviceSetup = 1; i++; }
      else if (strcmp(argv[i], "-p") == 0) { opts->showProg = 1; i++; }
      else if (strcmp(argv[i], "-q") == 0) { opts->quiet = 1; i++; }
      else if (strcmp(argv[i], "-v") == 0) { opts->viz = 1; i++; }   // lulesh-util.cc:146-152 with VIZ_MESH
      else if (strcmp(argv[i], "-h") == 0) {
         PrintCommandLineOptions(argv[0]);
         exit(0);
      } else {
         char msg[80];
         PrintCommandLineOptions(argv[0]);
         snprintf(msg, sizeof msg, "ERROR: Unknown command line argument: %s\n", argv[i]);
         ParseError(msg);
      }
   }
}

// ---- lulesh-util.cc:175-230, reading e() from the host Domain after download.
// `zones` is the true global zone count (nx^3*numRanks for the reference's
// cubic layouts); `n` is the edge of the square part of rank 0's plane 0.
struct ReportView {   // what the final report reads from rank 0's Domain
   Int_t cycles;
   const Real_t *energy;   // e(i), i < sizeX*sizeY is enough
   Index_t sx, sy;
   Int_t cycle() const { return cycles; }
   Real_t e(Index_t i) const { return energy[i]; }
   Index_t sizeX() const { return sx; }
   Index_t sizeY() const { return sy; }
};

void VerifyAndWriteFinalOutput(Real_t elapsed_time, const ReportView &locDom, Int_t nx, Int_t numRanks,
                               long long zones)
{
   const Real_t perDom = Real_t(zones) / Real_t(numRanks);
   Real_t grindTime1 = ((elapsed_time * 1e6) / locDom.cycle()) / perDom;
   Real_t grindTime2 = ((elapsed_time * 1e6) / locDom.cycle()) / Real_t(zones);

   Index_t ElemId = 0;
   std::cout << "Run completed:\n";
   std::cout << "   Problem size        =  " << nx << "\n";
   std::cout << "   MPI tasks           =  " << numRanks << "\n";
   std::cout << "   Iteration count     =  " << locDom.cycle() << "\n";
   std::cout << "   Final Origin Energy =  ";
   std::cout << std::scientific << std::setprecision(6);
   std::cout << std::setw(12) << locDom.e(ElemId) << "\n";

   Real_t MaxAbsDiff = Real_t(0.0), TotalAbsDiff = Real_t(0.0), MaxRelDiff = Real_t(0.0);
   const Index_t stride = locDom.sizeX();
   const Index_t n = std::min(locDom.sizeX(), locDom.sizeY());
   for (Index_t j = 0; j < n; ++j) {
      for (Index_t k = j + 1; k < n; ++k) {
         Real_t AbsDiff = fabs(locDom.e(j * stride + k) - locDom.e(k * stride + j));
         TotalAbsDiff += AbsDiff;
         if (MaxAbsDiff < AbsDiff) MaxAbsDiff = AbsDiff;
         Real_t RelDiff = AbsDiff / locDom.e(k * stride + j);
         if (MaxRelDiff < RelDiff) MaxRelDiff = RelDiff;
      }
   }
   std::cout << "   Testing Plane 0 of Energy Array on rank 0:\n";
   std::cout << "        MaxAbsDiff   = " << std::setw(12) << MaxAbsDiff << "\n";
   std::cout << "        TotalAbsDiff = " << std::setw(12) << TotalAbsDiff << "\n";
   std::cout << "        MaxRelDiff   = " << std::setw(12) << MaxRelDiff << "\n";

   std::cout.unsetf(std::ios_base::floatfield);
   std::cout << std::setprecision(2);
   std::cout << "\nElapsed time         = " << std::setw(10) << elapsed_time << " (s)\n";
   std::cout << std::setprecision(8);
   std::cout << "Grind time (us/z/c)  = " << std::setw(10) << grindTime1 << " (per dom)  ("
             << std::setw(10) << elapsed_time << " overall)\n";
   std::cout << "FOM                  = " << std::setw(10) << 1000.0 / grindTime2 << " (z/s)\n\n";
   if (getenv("LULESH_B200_FULL_PRECISION"))   // test hook: %.17g record like oracle/ref_util_wrap.cc
      printf("B200JSON {\"cycles\": %d, \"e0\": %.17g, \"max_abs_diff\": %.17g, "
             "\"total_abs_diff\": %.17g, \"max_rel_diff\": %.17g, \"zones_per_s\": %.9g}\n",
             (int)locDom.cycle(), locDom.e(0), MaxAbsDiff, TotalAbsDiff, MaxRelDiff,
             (double)zones * locDom.cycle() / elapsed_time);
}

void progress(int32_t cycle, double time, double dt, void *)
{
   std::cout << "cycle = " << cycle << ", " << std::scientific << "time = " << time << ", "
             << "dt=" << dt << "\n";   // lulesh.cc:2750-2756
   std::cout.unsetf(std::ios_base::floatfield);
}

double wallclock()
{
   timeval t;
   gettimeofday(&t, NULL);
   return (double)t.tv_sec + (double)t.tv_usec / 1000000;
}

struct RankState {
   std::unique_ptr<Domain> dom;
   lulesh_b200 *handle = nullptr;
   int status = 0;
   double elapsed = 0.0;
};

}  // namespace

extern "C" int lulesh_host_decompose(int numRanks, int *px, int *py, int *pz)
{
   if (numRanks < 1) return -1;
   const int c = (int)(cbrt((double)numRanks) + 0.5);   // lulesh-init.cc:684
   if (c * c * c == numRanks) { *px = *py = *pz = c; return 0; }
   if (numRanks == 2) { *px = 1; *py = 1; *pz = 2; return 0; }   // cut z: faces are contiguous
   if (numRanks == 4) { *px = 1; *py = 2; *pz = 2; return 0; }
   return -1;
}

extern "C" int lulesh_host_main(int argc, char **argv)
{
   cmdLineOpts opts;
   memset(&opts, 0, sizeof opts);
   // defaults, lulesh.cc:2682-2690
   opts.its = 9999999; opts.nx = 30; opts.numReg = 11; opts.balance = 1; opts.cost = 1;
   opts.gpus = 1; opts.syncEvery = 64;
   ParseCommandLineOptions(argc, argv, &opts);
   int numRanks = opts.gpus;
   if (opts.px > 0) numRanks = opts.px * opts.py * opts.pz;
   else if (lulesh_host_decompose(numRanks, &opts.px, &opts.py, &opts.pz) != 0) {
      printf("Num processors must be a cube of an integer (1, 8, 27, ...) or 2 or 4\n");
      exit(-1);
   }
   opts.numFiles = (int)(numRanks + 10) / 9;

   Index_t sx = opts.nx, sy = opts.nx, sz = opts.nx;
   if (opts.global > 0) {
      if (opts.global % opts.px || opts.global % opts.py || opts.global % opts.pz) {
         printf("--global %d is not divisible by the decomposition %dx%dx%d\n", opts.global,
                opts.px, opts.py, opts.pz);
         exit(-1);
      }
      sx = opts.global / opts.px; sy = opts.global / opts.py; sz = opts.global / opts.pz;
      opts.nx = opts.global;
   }
   const long long zones = (long long)numRanks * sx * sy * sz;

   if (opts.quiet == 0) {   // lulesh.cc:2694-2708
      std::cout << "Running problem size " << opts.nx << "^3 per domain until completion\n";
      std::cout << "Num processors: " << numRanks << "\n";
      std::cout << "Total number of elements: " << zones << " \n\n";
      std::cout << "To run other sizes, use -s <integer>.\n";
      std::cout << "To run a fixed number of iterations, use -i <integer>.\n";
      std::cout << "To run a more or less balanced region set, use -b <integer>.\n";
      std::cout << "To change the relative costs of regions, use -c <integer>.\n";
      std::cout << "To print out progress, use -p\n";
      std::cout << "To write an output file for VisIt, use -v\n";
      std::cout << "See help (-h) for more options\n\n";
   }

   unsigned char uid[LULESH_B200_UNIQUE_ID_BYTES] = {0};
   if (numRanks > 1 && lulesh_b200_get_unique_id(uid) != 0) {
      fprintf(stderr, "lulesh_b200: %s\n", lulesh_b200_last_error());
      return 1;
   }

   std::vector<RankState> ranks(numRanks);
   pthread_barrier_t barrier;
   pthread_barrier_init(&barrier, NULL, numRanks);
   double t_start = 0.0;

   auto body = [&](int r) {
      RankState &st = ranks[r];
      // lulesh.cc:2712-2716 (InitMeshDecomp + new Domain)
      if (opts.deviceSetup) {   // mesh, connectivity, BCs, masses generated by kernels in HBM
         lulesh_b200_sedov_params sp;
         memset(&sp, 0, sizeof sp);
         sp.abi_version = LULESH_B200_ABI_VERSION;
         sp.numRanks = numRanks; sp.rank = r;
         sp.px = opts.px; sp.py = opts.py; sp.pz = opts.pz;
         sp.sx = sx; sp.sy = sy; sp.sz = sz;
         sp.numReg = opts.numReg; sp.balance = opts.balance; sp.cost = opts.cost;
         st.status = lulesh_b200_create_sedov(&sp, r, numRanks > 1 ? uid : NULL, &st.handle);
      } else {
         st.dom.reset(new Domain(numRanks, r, opts.px, opts.py, opts.pz, sx, sy, sz, opts.numReg,
                                 opts.balance, opts.cost));
         lulesh_b200_host_view view = st.dom->view();
         st.status = lulesh_b200_create(&view, r, numRanks > 1 ? uid : NULL, &st.handle);
      }
      if (st.status == 0) st.status = lulesh_b200_sum_nodal_mass(st.handle);   // lulesh.cc:2720-2732
      if (st.status != 0) fprintf(stderr, "lulesh_b200 (rank %d): %s\n", r, lulesh_b200_last_error());
      pthread_barrier_wait(&barrier);
      bool ok = true;
      for (const RankState &o : ranks) ok = ok && (o.status == 0);
      if (!ok) return;
      if (r == 0) t_start = wallclock();   // lulesh.cc:2737-2741
      pthread_barrier_wait(&barrier);
      // -p: every rank polls every cycle (the ranks of one run must enqueue the same number of
      // cycles, see lulesh_b200_run); only rank 0 prints (lulesh.cc:2750).
      const bool show = (opts.showProg != 0) && (opts.quiet == 0);
      st.status = lulesh_b200_run(st.handle, opts.its, show ? 1 : opts.syncEvery,
                                  (show && r == 0) ? progress : NULL, NULL);   // lulesh.cc:2745-2757
      st.elapsed = wallclock() - t_start;
      if (st.status != 0 && st.status != LULESH_B200_VOLUME_ERROR && st.status != LULESH_B200_QSTOP_ERROR)
         fprintf(stderr, "lulesh_b200 (rank %d): run failed with status %d: %s\n", r, st.status, lulesh_b200_last_error());
      pthread_barrier_wait(&barrier);
   };
   if (numRanks == 1) body(0);
   else {
      std::vector<std::thread> threads;
      for (int r = 0; r < numRanks; ++r) threads.emplace_back(body, r);
      for (auto &t : threads) t.join();
   }

   int status = 0;
   double elapsed = 0.0;   // MPI_Reduce(MAX), lulesh.cc:2770-2771
   for (const RankState &st : ranks) {   // the reference's exit codes outrank infrastructure failures
      const bool physics = (st.status == LULESH_B200_VOLUME_ERROR || st.status == LULESH_B200_QSTOP_ERROR);
      const bool have = (status == LULESH_B200_VOLUME_ERROR || status == LULESH_B200_QSTOP_ERROR);
      if (st.status != 0 && (status == 0 || (physics && !have))) status = st.status;
      elapsed = std::max(elapsed, st.elapsed);
   }
   if (status == LULESH_B200_VOLUME_ERROR) exit(-1);   // lulesh.h:42, lulesh.cc:1038
   if (status == LULESH_B200_QSTOP_ERROR) exit(-2);
   if (status != 0) return 1;

   if (opts.quiet == 0) {   // lulesh.cc:2781-2783
      lulesh_b200_scalars s;
      lulesh_b200_get_scalars(ranks[0].handle, &s);
      std::vector<Real_t> e((size_t)sx * sy * sz);
      lulesh_b200_download(ranks[0].handle, LULESH_F_E, e.data(), e.size());
      if (ranks[0].dom) {   // keep the host Domain in step with the device, as a caller would
         ranks[0].dom->scalars() = s;
         std::copy(e.begin(), e.end(), &ranks[0].dom->e(0));
      }
      const ReportView rv = {s.cycle, e.data(), sx, sy};
      VerifyAndWriteFinalOutput(elapsed, rv, opts.nx, numRanks, zones);
   }
   if (opts.viz) {   // lulesh.cc:2776-2778 (DumpToVisit); -f is accepted, the dump is one block per rank
      lulesh_b200_scalars s;
      lulesh_b200_get_scalars(ranks[0].handle, &s);
      char basename[32], name[64];
      snprintf(basename, sizeof basename, "lulesh_plot_c%d", s.cycle);   // lulesh-viz.cc:62
      static const int dumped[] = {LULESH_F_X, LULESH_F_Y, LULESH_F_Z, LULESH_F_XD, LULESH_F_YD,
                                   LULESH_F_ZD, LULESH_F_E, LULESH_F_P, LULESH_F_V, LULESH_F_Q};
      for (int r = 0; r < numRanks; ++r) {
         RankState &st = ranks[r];
         if (!st.dom)   // device-side setup keeps no host mesh: rebuild it for the connectivity
            st.dom.reset(new Domain(numRanks, r, opts.px, opts.py, opts.pz, sx, sy, sz, opts.numReg,
                                    opts.balance, opts.cost));
         st.dom->scalars() = s;
         for (int fld : dumped) {
            std::vector<Real_t> *a = st.dom->realField(fld);
            if (lulesh_b200_download(st.handle, fld, a->data(), a->size()) != 0) {
               fprintf(stderr, "lulesh_b200 (rank %d): %s\n", r, lulesh_b200_last_error());
               return 1;
            }
         }
         snprintf(name, sizeof name, "%s.%03d.vtk", basename, r);
         if (DumpDomainToVTK(*st.dom, r, name) != 0) {
            fprintf(stderr, "lulesh_b200: cannot write %s\n", name);
            return 1;
         }
      }
      snprintf(name, sizeof name, "%s.visit", basename);
      if (WriteVisitIndex(name, basename, numRanks) != 0) {
         fprintf(stderr, "lulesh_b200: cannot write %s\n", name);
         return 1;
      }
   }
   for (RankState &st : ranks) lulesh_b200_destroy(st.handle);
   pthread_barrier_destroy(&barrier);
   return 0;
}
