// Micro-probe: FP64 issue-rate ceilings on sm_100a (DFMA vs DMMA shapes). Not product code.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters) {
  double a[16]; double b = threadIdx.x * 1e-9 + 1.0, c = 0.999999;
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = i + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0; for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma884(double* out, int iters) {
  double c[8][2]; double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1684(double* out, int iters) {
  double c[4][4]; double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, b = 1.0 + threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
  }
  double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma1688(double* out, int iters) {
  double c[4][4]; double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 1.0 + threadIdx.x * 1e-9, b1 = b0 * 2;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
  }
  double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma16816(double* out, int iters) {
  double c[4][4]; double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-9 * i;
#pragma unroll
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
  int nsm = p.multiProcessorCount; double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32; int blocks = nsm * 2;
    if (threads * 2 > 2048) blocks = nsm;
    double nthr = (double)blocks * threads, nw = nthr / 32;
    float ms;
    ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("warps/blk=%2d blocks=%d  DFMA        %.2f TFLOP/s\n", warps, blocks, nthr * iters * 16 * 2 / ms / 1e9);
    ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
    printf("warps/blk=%2d blocks=%d  DMMA.884    %.2f TFLOP/s\n", warps, blocks, nw * iters * 8 * (8 * 8 * 4 * 2.0) / ms / 1e9);
    ms = timeit([&] { k_dmma1684<<<blocks, threads>>>(out, iters); });
    printf("warps/blk=%2d blocks=%d  DMMA.1684   %.2f TFLOP/s\n", warps, blocks, nw * iters * 4 * (16 * 8 * 4 * 2.0) / ms / 1e9);
    ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
    printf("warps/blk=%2d blocks=%d  DMMA.1688   %.2f TFLOP/s\n", warps, blocks, nw * iters * 4 * (16 * 8 * 8 * 2.0) / ms / 1e9);
    ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters); });
    printf("warps/blk=%2d blocks=%d  DMMA.16816  %.2f TFLOP/s\n", warps, blocks, nw * iters * 4 * (16 * 8 * 16 * 2.0) / ms / 1e9);
  }
  CK(cudaGetLastError());
  return 0;
}
