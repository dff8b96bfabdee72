// Roofline denominators that MEASURED_PEAKS.json lacks: FP32 FMA, FP64 FMA issue rate and
// L2-resident read bandwidth.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void fmaKernel(T* out, int iters) {
  T a0 = threadIdx.x * (T)1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T b = (T)1.0000001, c = (T)1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
    a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void l2Read(const float4* __restrict__ in, size_t n, float4* out, int reps) {
  float4 acc = make_float4(0, 0, 0, 0);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      float4 v = __ldg(in + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if (acc.x == 12345.678f) out[0] = acc;
}

template <typename T>
double runFma(const char* name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  T* out; cudaMalloc(&out, sizeof(T) * sms * 8 * 256);
  const int iters = 1 << 16;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  fmaKernel<T><<<sms * 8, 256>>>(out, 1024);
  double best = 0;
  for (int k = 0; k < 3; ++k) {
    cudaEventRecord(e0);
    fmaKernel<T><<<sms * 8, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * 8 * 256 * 8.0 * iters;
    double rate = fma / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  printf("{\"bench\": \"%s\", \"fma_per_s\": %.4e, \"tflops\": %.2f}\n", name, best, 2 * best / 1e12);
  cudaFree(out);
  return best;
}

int main() {
  runFma<float>("fp32_fma");
  runFma<double>("fp64_fma");
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (size_t mb : {32, 48, 64, 78, 96, 114}) {  // MiB; 78 MiB = the soup_1m BVH with 64-byte quantised nodes, 114 with the float32 nodes
    size_t n = mb * 1024 * 1024 / 16;
    float4 *in, *out; cudaMalloc(&in, n * 16); cudaMalloc(&out, 16);
    cudaMemset(in, 0, n * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    l2Read<<<sms * 8, 256>>>(in, n, out, 2);
    const int reps = 20;
    cudaEventRecord(e0);
    l2Read<<<sms * 8, 256>>>(in, n, out, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"bench\": \"l2_read\", \"working_set_mb\": %zu, \"gb_per_s\": %.1f}\n", mb, (double)n * 16 * reps / (ms * 1e-3) / 1e9);
    cudaFree(in); cudaFree(out);
  }
  return 0;
}
