// Micro-benchmark: fp64 dependent-issue latency and multi-warp throughput on one SM partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scripts/ubench/fp64_lat scripts/ubench/fp64_lat.cu
// Prints cycles per instruction for chains of DFMA / DMUL / DADD / MUFU.RCP64H+DFMA with ILP 1..4 and 1..8 warps
// per scheduler (one block per SM, block = 128 * warps_per_scheduler threads).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void chain(double* out, int iters, double a, double b, long long* cyc)
{
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x + k;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
#pragma unroll
      for (int k = 0; k < ILP; ++k)
      {
        if (OP == 0) x[k] = __fma_rn(x[k], a, b);
        if (OP == 1) x[k] = __dmul_rn(x[k], a);
        if (OP == 2) x[k] = __dadd_rn(x[k], b);
        if (OP == 3)
        {
          double y;
          asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[k]));
          x[k] = __fma_rn(y, a, b);
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int OP>
void run(const char* name, int wps)
{
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  chain<ILP, OP><<<148, 128 * wps>>>(out, iters, 1.0000001, 1e-9, cyc);
  chain<ILP, OP><<<148, 128 * wps>>>(out, iters, 1.0000001, 1e-9, cyc);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double n = 8.0 * iters * ILP * (OP == 3 ? 2 : 1);  // instructions per warp
  printf("%-6s ILP %d warps/sched %d : %.2f cycles per warp-instruction per warp, %.3f inst/cycle/scheduler\n", name, ILP,
         wps, h / n, n * wps / h);
  cudaFree(out);
  cudaFree(cyc);
}

int main()
{
  for (int wps : {1, 2, 4, 8})
  {
    run<1, 0>("DFMA", wps);
    run<2, 0>("DFMA", wps);
    run<4, 0>("DFMA", wps);
  }
  run<1, 1>("DMUL", 1);
  run<1, 2>("DADD", 1);
  run<1, 3>("RCP+F", 1);
  run<1, 3>("RCP+F", 4);
  run<2, 3>("RCP+F", 4);
  return 0;
}
