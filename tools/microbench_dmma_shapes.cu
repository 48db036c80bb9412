// FP64 tensor-core instruction shapes on sm_100a: sustained throughput of mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16
// (.f64) per SM, ILP independent accumulator sets per warp. Development aid:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb/shapes tools/microbench_dmma_shapes.cu && /tmp/mb/shapes
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE> struct Mma;
template <> struct Mma<0> {   // m8n8k4
  static constexpr int NC = 2, FLOP = 2 * 8 * 8 * 4;
  __device__ static void run(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
  }
};
template <> struct Mma<1> {   // m16n8k4
  static constexpr int NC = 4, FLOP = 2 * 16 * 8 * 4;
  __device__ static void run(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  }
};
template <> struct Mma<2> {   // m16n8k8
  static constexpr int NC = 4, FLOP = 2 * 16 * 8 * 8;
  __device__ static void run(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  }
};
template <> struct Mma<3> {   // m16n8k16
  static constexpr int NC = 4, FLOP = 2 * 16 * 8 * 16;
  __device__ static void run(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
};

template <int SHAPE, int ILP>
__global__ void k_tput(double *out, long long *clk, int n) {
  double c[ILP][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = out[(threadIdx.x + i) & 31];
  for (int i = 0; i < 4; ++i) b[i] = out[(threadIdx.x + 7 + i) & 31];
  for (int k = 0; k < ILP; ++k) for (int i = 0; i < 4; ++i) c[k][i] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) Mma<SHAPE>::run(c[k], a, b);
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; ++k) for (int i = 0; i < 4; ++i) s += c[k][i];
  out[512 + threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int SHAPE>
void run(const char *name, double *out, long long *clk) {
  const int n = 2048;
  for (int warps : {4, 8, 16}) {
    k_tput<SHAPE, 4><<<148, 32 * warps>>>(out, clk, n);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (long long x : h) mx = x > mx ? (double)x : mx;
    const double flop_per_clk = (double)n * 4 * warps * Mma<SHAPE>::FLOP / mx;
    printf("%-9s ILP4 %2d warps/SM: %7.1f flop/clk/SM  (%.1f TF/s at 1.965 GHz x 148 SMs)\n", name, warps, flop_per_clk, flop_per_clk * 1.965e9 * 148 / 1e12);
  }
}

int main() {
  double *out; long long *clk;
  cudaMalloc(&out, (512 + 148 * 512) * 8); cudaMemset(out, 0, (512 + 148 * 512) * 8);
  cudaMalloc(&clk, 1024 * 8);
  run<0>("m8n8k4", out, clk);
  run<1>("m16n8k4", out, clk);
  run<2>("m16n8k8", out, clk);
  run<3>("m16n8k16", out, clk);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
