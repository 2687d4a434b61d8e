// FP64 pipe micro-benchmark, part 2: does a DFMA with three DISTINCT register operands
// issue as fast as one that re-uses operands?  (register-file bank / operand-collector effects)
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void chain(double *out, const double *in, int iters)
{
   double v[ILP], w[ILP], u[ILP];
#pragma unroll
   for (int j = 0; j < ILP; ++j) { v[j] = in[j] + threadIdx.x * 1e-9; w[j] = in[8 + j]; u[j] = in[16 + j]; }
   for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
         for (int j = 0; j < ILP; ++j) {
            if (MODE == 0) v[j] = fma(v[j], w[0], u[0]);          // two shared operands (reuse cache)
            else if (MODE == 1) v[j] = fma(v[j], w[j], u[j]);     // three distinct registers per chain
            else if (MODE == 2) v[j] = v[j] + w[j];               // DADD, two distinct
            else v[j] = fma(w[j], u[(j + 1) % ILP], v[j]);        // accumulate form, three distinct
         }
   }
   double s = 0;
#pragma unroll
   for (int j = 0; j < ILP; ++j) s += v[j] + w[j] + u[j];
   if (s == 12345.678) out[0] = s;
}

template <int ILP, int MODE>
void run(int warps_per_sm, int sms, double *d, const double *in)
{
   const int iters = 4096, threads = warps_per_sm * 32;
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   chain<ILP, MODE><<<sms, threads>>>(d, in, 16);
   cudaEventRecord(e0);
   chain<ILP, MODE><<<sms, threads>>>(d, in, iters);
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms;
   cudaEventElapsedTime(&ms, e0, e1);
   const double inst = (double)sms * warps_per_sm * 32.0 * iters * 8 * ILP;
   const char *names[] = {"DFMA shared-operands", "DFMA 3-distinct", "DADD 2-distinct", "DFMA accumulate 3-distinct"};
   printf("%-28s ilp=%d warps/SM=%2d: %6.2f T inst/s -> %5.1f%% of 64 lanes/clk/SM\n", names[MODE], ILP, warps_per_sm,
          inst / ms / 1e9, 100.0 * (inst / (ms * 1e-3)) / (sms * 64.0 * 1.965e9));
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   const int sms = p.multiProcessorCount;
   double *d, *in;
   cudaMalloc(&d, 64);
   cudaMalloc(&in, 32 * 8);
   double h[32];
   for (int i = 0; i < 32; ++i) h[i] = 0.999 + 1e-6 * i;
   cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
   const int ws[] = {4, 8, 12, 16, 32};
   for (int w : ws) run<4, 0>(w, sms, d, in);
   for (int w : ws) run<4, 1>(w, sms, d, in);
   for (int w : ws) run<8, 1>(w, sms, d, in);
   for (int w : ws) run<8, 3>(w, sms, d, in);
   for (int w : ws) run<8, 2>(w, sms, d, in);
   return 0;
}
