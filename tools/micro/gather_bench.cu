// gather_bench.cu -- microbenchmark behind the node-fetch design: divergent per-lane fetches of 128-byte records that sit in
// L2/L1, with 8 x LDG.128 (what k_trace does today) against 4 x LDG.256 (sm_100's 32-byte loads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void ld256(const void *p, float (&a)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
               : "l"(p));
}

// every thread chases `steps` records: next index depends on the data (like a traversal)
template <int MODE> __global__ void __launch_bounds__(128, 6) k_gather(const float4 *rec, unsigned n_rec, unsigned steps, unsigned coh, float *out) {
  unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  // coh = lanes per group that walk together (32 = coherent warp, 1 = fully divergent)
  unsigned walker = tid / coh;
  unsigned idx = (walker * 2654435761u) % n_rec;
  float acc = 0.f;
  for (unsigned s = 0; s < steps; ++s) {
    const float4 *p = rec + 8 * (size_t)idx;
    float sum = 0.f;
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 v = __ldg(p + k);
        sum += v.x + v.y + v.z + v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a[8];
        ld256(p + 2 * k, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += a[j];
      }
    }
    acc += sum;
    idx = (idx * 1664525u + 1013904223u + (unsigned)(sum != 12345.f)) % n_rec;
  }
  out[tid] = acc;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const unsigned steps = 256;
  const int grid = sms * 6, block = 128;
  float *out;
  cudaMalloc(&out, (size_t)grid * block * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (unsigned n_rec : { 1024u, 16384u, 65536u, 1048576u }) {
    float4 *rec;
    cudaMalloc(&rec, (size_t)n_rec * 128);
    cudaMemset(rec, 0, (size_t)n_rec * 128);
    for (unsigned coh : { 32u, 4u, 1u }) {
      for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          cudaEventRecord(e0);
          if (mode == 0)
            k_gather<0><<<grid, block>>>(rec, n_rec, steps, coh, out);
          else
            k_gather<1><<<grid, block>>>(rec, n_rec, steps, coh, out);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        double recs = (double)grid * block * steps;
        printf("{\"records\": %u, \"MB\": %.2f, \"lanes_together\": %u, \"load\": \"%s\", \"ms\": %.4f, \"Grec_s\": %.3f, \"TB_s\": %.3f}\n", n_rec,
               n_rec * 128.0 / 1e6, coh, mode ? "4xLDG.256" : "8xLDG.128", best, recs / best / 1e6, recs * 128 / best / 1e9);
      }
    }
    cudaFree(rec);
  }
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(err));
  return 0;
}
