// FP64 dependent-issue latency / throughput probe for B200 (sm_100a). nvcc --fmad=false.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_add(double* out, double a, int n, long long* cyc) {
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_mul(double* out, double a, int n, long long* cyc) {
  double x = 1.0 + 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = x * a; x = x * a; x = x * a; x = x * a; x = x * a; x = x * a; x = x * a; x = x * a; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_div(double* out, double a, int n, long long* cyc) {
  double x = 1.0 + 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = a / x; x = a / x; x = a / x; x = a / x; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_sqrt(double* out, double a, int n, long long* cyc) {
  double x = 2.0 + 1e-9 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = sqrt(x) + a; x = sqrt(x) + a; x = sqrt(x) + a; x = sqrt(x) + a; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ilp_add(double* out, double a, int n, long long* cyc) {  // 8 independent chains
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_fadd(float* out, float a, int n, long long* cyc) {
  float x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; x = x + a; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x; if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  long long h; int n = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    int thr = 32 * warps;
    chain_add<<<1, thr>>>(out, 1e-3, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("warps/SM %2d: DADD chain %.1f cyc/op", warps, (double)h / (8.0 * n));
    chain_mul<<<1, thr>>>(out, 1.0000001, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("  DMUL %.1f", (double)h / (8.0 * n));
    chain_div<<<1, thr>>>(out, 1.5, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("  DDIV %.1f", (double)h / (4.0 * n));
    chain_sqrt<<<1, thr>>>(out, 1.5, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("  DSQRT+DADD %.1f", (double)h / (4.0 * n));
    ilp_add<<<1, thr>>>(out, 1e-3, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("  8xILP DADD %.2f cyc/op", (double)h / (8.0 * n));
    chain_fadd<<<1, thr>>>((float*)out, 1e-3f, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("  FADD chain %.1f\n", (double)h / (8.0 * n));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
