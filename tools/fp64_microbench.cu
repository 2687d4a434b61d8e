// FP64 pipe micro-benchmark for sm_100a: DFMA/DADD throughput as a function of resident
// warps per SM sub-partition and of independent chains per warp (ILP).  Used to size the
// occupancy the FP64-bound element kernels need.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool ADD>
__global__ void chain(double *out, int iters, double a, double b)
{
   double v[ILP];
#pragma unroll
   for (int j = 0; j < ILP; ++j) v[j] = a + j + threadIdx.x * 1e-9;
   for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
         for (int j = 0; j < ILP; ++j) v[j] = ADD ? (v[j] + b) : fma(v[j], b, a);
   }
   double s = 0;
#pragma unroll
   for (int j = 0; j < ILP; ++j) s += v[j];
   if (s == 12345.678) out[0] = s;
}

template <int ILP, bool ADD>
void run(int warps_per_sm, int sms, double *d)
{
   const int iters = 4096;
   const int threads = warps_per_sm * 32;   // one block per SM
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   chain<ILP, ADD><<<sms, threads>>>(d, 16, 1.0, 0.999999);
   cudaEventRecord(e0);
   chain<ILP, ADD><<<sms, threads>>>(d, iters, 1.0, 0.999999);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   const double inst = (double)sms * warps_per_sm * 32.0 * iters * 8 * ILP;   // thread-level instructions
   printf("%s ilp=%d warps/SM=%2d (%.2f/SMSP): %7.2f G thread-inst/s per SM-clk-normalised: %6.2f T inst/s  -> %5.1f%% of 64 lanes/clk/SM @1.965GHz\n",
          ADD ? "DADD" : "DFMA", ILP, warps_per_sm, warps_per_sm / 4.0, inst / ms / 1e6, inst / ms / 1e9,
          100.0 * (inst / (ms * 1e-3)) / (sms * 64.0 * 1.965e9));
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   const int sms = p.multiProcessorCount;
   double *d;
   cudaMalloc(&d, 64);
   printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
   const int ws[] = {4, 8, 12, 16, 24, 32};
   for (int w : ws) { run<1, false>(w, sms, d); }
   for (int w : ws) { run<2, false>(w, sms, d); }
   for (int w : ws) { run<4, false>(w, sms, d); }
   for (int w : ws) { run<8, false>(w, sms, d); }
   for (int w : ws) { run<1, true>(w, sms, d); }
   for (int w : ws) { run<4, true>(w, sms, d); }
   return 0;
}
