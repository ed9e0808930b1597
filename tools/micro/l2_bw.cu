// l2_bw.cu -- read bandwidth of the B200's L2 and HBM as a kernel sees it ("L2 bandwidth is not in MEASURED_PEAKS.json:
// measure it with a read microbenchmark first").  A buffer is read over and over with 16-byte loads (LDG.E.128, ld.global.cg so
// that L1 does not serve it) by a grid that fills the chip; buffers up to ~64 MB stay resident in the 126 MB L2, larger ones
// stream from HBM.  Prints one JSON line per buffer size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/l2_bw tools/micro/l2_bw.cu && tools/micro/l2_bw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_read(const uint4 *p, size_t n, int reps, unsigned *sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
      const uint4 v = __ldcg(p + i);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x12345678u) *sink = acc; // never true in practice: keeps the loads alive
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  unsigned *sink;
  cudaMalloc(&sink, 4);
  const size_t sizes_mb[] = { 8, 16, 32, 48, 64, 96, 128, 256, 1024, 4096 };
  for (size_t mb : sizes_mb) {
    const size_t bytes = mb << 20, n = bytes / 16;
    uint4 *buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) break;
    cudaMemset(buf, 1, bytes);
    const int reps = (int)(mb <= 128 ? (8192 / mb) : (mb <= 1024 ? 8 : 2));
    const int grid = prop.multiProcessorCount * 8, block = 256;
    k_read<<<grid, block>>>(buf, n, 2, sink); // warm: the buffer is in L2 afterwards if it fits
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int t = 0; t < 5; ++t) {
      cudaEventRecord(e0);
      k_read<<<grid, block>>>(buf, n, reps, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("{\"buffer_MB\": %zu, \"reps\": %d, \"ms\": %.4f, \"read_GBps\": %.1f, \"sms\": %d}\n", mb, reps, best,
           (double)bytes * reps / (best * 1e-3) / 1e9, prop.multiProcessorCount);
    cudaFree(buf);
  }
  return 0;
}
