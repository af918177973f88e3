// Dependent-chain latencies on one thread (cycles per op): FP64 add / mul / div / sqrt, shared-memory load through a
// generic pointer, L1-hit and L2-hit global loads.  Measurement helper for the latency-bound kernels (k_search, k_qp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed, const int* chase_g, int n_chase)
{
  __shared__ int chase_s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) chase_s[i] = (i * 17 + 1) & 1023;
  __syncthreads();
  if (threadIdx.x != 0) return;
  double x = seed, y = seed * 0.5;
  long long t0, t1;
  const int N = 2048;
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __dadd_rn(x, y);
  t1 = clock64(); cyc[0] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < N; i++) x = __dmul_rn(x, 1.0000001);
  t1 = clock64(); cyc[1] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < N; i++) x = x / 1.0000001;
  t1 = clock64(); cyc[2] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < N; i++) x = sqrt(x) + 1.0;
  t1 = clock64(); cyc[3] = (t1 - t0);
  int p = 0;
  const int* gs = chase_s;   // generic pointer to shared memory
  t0 = clock64();
  for (int i = 0; i < N; i++) p = gs[p];
  t1 = clock64(); cyc[4] = (t1 - t0);
  int q = 0;
  for (int i = 0; i < 256; i++) q = chase_g[q & 255];   // warm L1 for a small ring
  t0 = clock64();
  for (int i = 0; i < N; i++) q = chase_g[q & 255];
  t1 = clock64(); cyc[5] = (t1 - t0);
  int r = 0;
  t0 = clock64();
  for (int i = 0; i < N; i++) r = __ldcg(chase_g + r);   // L2 (bypass L1), big ring
  t1 = clock64(); cyc[6] = (t1 - t0);
  t0 = clock64();
  for (int i = 0; i < N; i++) x = fma(x, 0.999, y);
  t1 = clock64(); cyc[7] = (t1 - t0);
  out[0] = x + p + q + r;
}
int main()
{
  const int n = 1 << 20;
  int* h = new int[n];
  for (int i = 0; i < n; i++) h[i] = (int)(((long long)i * 7919 + 12345) % n);
  for (int i = 0; i < 256; i++) h[i] = (i * 5 + 1) & 255;
  int* d; double* o; long long* c;
  cudaMalloc(&d, n * sizeof(int)); cudaMalloc(&o, 8); cudaMalloc(&c, 8 * 8);
  cudaMemcpy(d, h, n * sizeof(int), cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) k<<<1, 64>>>(o, c, 1.25, d, n);
  long long hc[8];
  cudaMemcpy(hc, c, 64, cudaMemcpyDeviceToHost);
  const char* nm[8] = { "dadd", "dmul", "ddiv", "dsqrt+add", "ld generic->shared", "ld global L1 hit", "ld.cg L2", "dfma" };
  for (int i = 0; i < 8; i++) printf("%-20s %.1f cycles/op\n", nm[i], hc[i] / 2048.0);
  return 0;
}
