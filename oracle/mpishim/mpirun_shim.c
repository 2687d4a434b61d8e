/* TEST INFRASTRUCTURE ONLY -- launcher for the MPI stand-in: mpirun_shim -np N prog args...
 * Creates the shared segment, forks N ranks, returns the first non-zero exit status. */
#define _GNU_SOURCE
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

size_t mpishim_segment_bytes(int nranks, long slot_bytes);

int main(int argc, char **argv)
{
   if (argc < 4 || strcmp(argv[1], "-np")) {
      fprintf(stderr, "usage: %s -np N prog [args...]\n", argv[0]);
      return 2;
   }
   const int n = atoi(argv[2]);
   if (n < 1 || n > 64) { fprintf(stderr, "mpirun_shim: bad rank count\n"); return 2; }
   const char *mb = getenv("MPISHIM_SLOT_MB");
   const long slot_bytes = (mb ? atol(mb) : 4) * 1024L * 1024L;
   char name[64];
   snprintf(name, sizeof name, "/mpishim_%d", (int)getpid());
   int fd = shm_open(name, O_CREAT | O_RDWR | O_EXCL, 0600);
   if (fd < 0) { perror("shm_open"); return 1; }
   const size_t bytes = mpishim_segment_bytes(n, slot_bytes);
   if (ftruncate(fd, (off_t)bytes) != 0) { perror("ftruncate"); shm_unlink(name); return 1; }
   int *hdr = mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
   if (hdr == MAP_FAILED) { perror("mmap"); shm_unlink(name); return 1; }
   hdr[0] = n;                       /* shim_header.nranks */
   *(long *)(hdr + 2) = slot_bytes;  /* shim_header.slot_bytes (after padding) */
   close(fd);

   pid_t pids[64];
   char buf[32];
   for (int r = 0; r < n; ++r) {
      pids[r] = fork();
      if (pids[r] == 0) {
         setenv("MPISHIM_SHM", name, 1);
         snprintf(buf, sizeof buf, "%d", r); setenv("MPISHIM_RANK", buf, 1);
         snprintf(buf, sizeof buf, "%d", n); setenv("MPISHIM_SIZE", buf, 1);
         snprintf(buf, sizeof buf, "%ld", slot_bytes); setenv("MPISHIM_SLOT_BYTES", buf, 1);
         execvp(argv[3], argv + 3);
         perror("execvp");
         _exit(127);
      }
   }
   int result = 0;
   for (int left = n; left > 0; --left) {
      int st;
      pid_t p = wait(&st);
      if (p < 0) break;
      const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
      if (code != 0 && result == 0) {
         result = code;
         for (int r = 0; r < n; ++r) if (pids[r] != p) kill(pids[r], SIGTERM);
      }
   }
   shm_unlink(name);
   return result;
}
