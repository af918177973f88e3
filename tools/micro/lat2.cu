// Warp-level dependent-chain latencies behind k_qp's triangular solves and sweeps: 64-bit shuffle, shuffle + FMA
// step, block barrier with 4 / 8 warps, independent-instruction issue rate of one warp, rcp sequence.
// Measurement helper (tools/micro): nvcc -arch=sm_100a -o lat2 lat2.cu && ./lat2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp2(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__global__ void k(double* out, long long* cyc, double seed)
{
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double x = seed + lane * 1e-3, l = 1e-3 * (lane + 1);
  long long t0, t1;
  const int N = 1024;
  // 0: shuffle chain (64-bit), whole block converged
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 1: shuffle + fma chain
  t0 = clock64();
  for (int i = 0; i < N; i++)
  {
    const double xk = __shfl_sync(0xffffffffu, x, i & 31);
    x = fma(-l, xk, x);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 2: the same inside a warp-0-only region (divergent for the compiler)
  __syncthreads();
  if (threadIdx.x < 32)
  {
    t0 = clock64();
    for (int i = 0; i < N; i++)
    {
      const double xk = __shfl_sync(0xffffffffu, x, i & 31);
      x = fma(-l, xk, x);
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
  }
  __syncthreads();
  // 3: block barrier chain
  t0 = clock64();
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 4: 8 independent fma chains (issue rate of one warp)
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
  for (int i = 0; i < N; i++)
  {
    a0 = fma(a0, 0.999, l), a1 = fma(a1, 0.999, l), a2 = fma(a2, 0.999, l), a3 = fma(a3, 0.999, l);
    a4 = fma(a4, 0.999, l), a5 = fma(a5, 0.999, l), a6 = fma(a6, 0.999, l), a7 = fma(a7, 0.999, l);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 8;
  x = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  // 5: rcp sequence chain
  t0 = clock64();
  for (int i = 0; i < N; i++) x = rcp2(x) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 6: shared store -> barrier -> load by another thread (broadcast through shared memory)
  t0 = clock64();
  for (int i = 0; i < N; i++)
  {
    if (threadIdx.x == (i & 31)) sm[i & 1023] = x;
    __syncthreads();
    x += sm[i & 1023];
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 7: 8 independent integer chains
  int b0 = lane, b1 = lane + 1, b2 = lane + 2, b3 = lane + 3, b4 = lane + 4, b5 = lane + 5, b6 = lane + 6, b7 = lane + 7;
  t0 = clock64();
  for (int i = 0; i < N; i++)
  {
    b0 = b0 * 3 + i, b1 = b1 * 3 + i, b2 = b2 * 3 + i, b3 = b3 * 3 + i, b4 = b4 * 3 + i, b5 = b5 * 3 + i, b6 = b6 * 3 + i, b7 = b7 * 3 + i;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = (t1 - t0) / 8;
  // 8: dependent integer chain
  t0 = clock64();
  for (int i = 0; i < N; i++) b0 = b0 * 3 + i;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  // 9: shared-memory double load chain (address depends on the value)
  int p = lane;
  t0 = clock64();
  for (int i = 0; i < N; i++) p = (int)sm[p & 1023] & 1023;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[9] = t1 - t0;
  out[threadIdx.x] = x + b0 + b1 + b2 + b3 + b4 + b5 + b6 + b7 + p + warp;
}
int main()
{
  double* o;
  long long* c;
  cudaMalloc(&o, 8 * 1024);
  cudaMalloc(&c, 16 * 8);
  const char* nm[10] = { "shfl64 chain", "shfl64 + dfma step", "same, warp-0-only region", "__syncthreads", "dfma, 8 independent chains (per op)",
                         "rcp (approx + 2 Newton) + add", "sts -> barrier -> lds", "imad, 8 independent (per op)", "imad dependent", "lds + cvt chain" };
  for (int nt : { 32, 128, 256 })
  {
    for (int rep = 0; rep < 2; rep++) k<<<1, nt>>>(o, c, 1.25);
    long long hc[16];
    cudaMemcpy(hc, c, 128, cudaMemcpyDeviceToHost);
    printf("block of %d threads\n", nt);
    for (int i = 0; i < 10; i++) printf("  %-40s %.1f cycles\n", nm[i], hc[i] / 1024.0);
  }
  return 0;
}
