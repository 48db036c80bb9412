// Latency / issue-rate probes of the FP64 building blocks the factorisation kernels are made of (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double *out, long long *clk, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, 1e-3);
  long long t1 = clock64();
  out[2] = a;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_dfma_tput(double *out, long long *clk, int n) {
  double a[ILP];
  for (int k = 0; k < ILP; ++k) a[k] = out[k];
  const double b = out[20];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, 1e-3);
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += a[k];
  out[32 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl_chain(double *out, long long *clk, int n) {
  double a = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 << (i % 5));
  long long t1 = clock64();
  out[64 + threadIdx.x] = a;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_chain(double *out, long long *clk, int n) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 17 + 5) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) p = nxt[p];
  long long t1 = clock64();
  out[128 + threadIdx.x] = p;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_rsqrt_chain(double *out, long long *clk, int n) {
  double a = out[0] + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = rsqrt(a) + 1.5;
  long long t1 = clock64();
  out[3] = a;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_div_chain(double *out, long long *clk, int n) {
  double a = out[0] + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) a = 1.0 / a + 1.5;
  long long t1 = clock64();
  out[4] = a;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_nanosleep(long long *clk, int n, unsigned ns) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) __nanosleep(ns);
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_flag_pingpong(long long *clk, int n, int use_sleep) {
  // warp 0 and warp 1 hand a shared-memory flag back and forth: round-trip cost of the publish/poll protocol
  __shared__ volatile int flag;
  if (threadIdx.x == 0) flag = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    const int want = 2 * i + warp;
    while (flag != want) { if (use_sleep) __nanosleep(40); }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) flag = want + 1;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_fence(long long *clk, int n) {
  __shared__ volatile int x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = i; __threadfence_block(); }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_dmma_chain(double *out, long long *clk, int n) {
  double c0 = 0, c1 = 0, a = out[threadIdx.x & 31], b = out[(threadIdx.x + 7) & 31];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  out[256 + threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
__global__ void k_dmma_tput(double *out, long long *clk, int n) {
  double c[ILP][2], a = out[threadIdx.x & 31], b = out[(threadIdx.x + 7) & 31];
  for (int k = 0; k < ILP; ++k) c[k][0] = c[k][1] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1];
  out[512 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  double *out; long long *clk, h;
  cudaMalloc(&out, 8192 * 8); cudaMemset(out, 0, 8192 * 8);
  cudaMalloc(&clk, 1024 * 8);
  const int n = 4096;
  auto rd = [&]() { cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); return (double)h; };
  k_dfma_chain<<<1, 32>>>(out, clk, n); printf("DFMA dependent chain            : %.1f cycles/op\n", rd() / n);
  for (int warps : {1, 2, 4, 8, 16}) {
    k_dfma_tput<8><<<1, 32 * warps>>>(out, clk, n);
    printf("DFMA ILP8, %2d warps on one SM    : %.2f cycles per warp-instruction (SM-wide)\n", warps, rd() / (n * 8.0 * warps));
  }
  k_shfl_chain<<<1, 32>>>(out, clk, n); printf("shfl.f64 + DADD dependent chain : %.1f cycles/step\n", rd() / n);
  k_lds_chain<<<1, 32>>>(out, clk, n); printf("LDS pointer chase               : %.1f cycles/load\n", rd() / n);
  k_rsqrt_chain<<<1, 32>>>(out, clk, n); printf("rsqrt(double)+DADD chain        : %.1f cycles/op\n", rd() / n);
  k_div_chain<<<1, 32>>>(out, clk, n); printf("1.0/x (double)+DADD chain       : %.1f cycles/op\n", rd() / n);
  for (unsigned ns : {0u, 20u, 40u, 100u}) { k_nanosleep<<<1, 32>>>(clk, 1024, ns); printf("__nanosleep(%3u)                : %.1f cycles\n", ns, rd() / 1024); }
  k_fence<<<1, 32>>>(clk, n); printf("st.volatile + membar.cta        : %.1f cycles\n", rd() / n);
  k_flag_pingpong<<<1, 64>>>(clk, 1024, 0); printf("flag hand-off (spin)            : %.1f cycles per hand-off\n", rd() / 2048);
  k_flag_pingpong<<<1, 64>>>(clk, 1024, 1); printf("flag hand-off (nanosleep 40)    : %.1f cycles per hand-off\n", rd() / 2048);
  k_dmma_chain<<<1, 32>>>(out, clk, n); printf("DMMA m8n8k4 dependent chain     : %.1f cycles/op\n", rd() / n);
  for (int warps : {1, 2, 4, 8, 16}) {
    k_dmma_tput<8><<<1, 32 * warps>>>(out, clk, n);
    printf("DMMA ILP8, %2d warps on one SM    : %.2f cycles per warp-instruction (SM-wide)\n", warps, rd() / (n * 8.0 * warps));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
