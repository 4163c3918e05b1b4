// Measurement helper (bench.py only; not part of the render path or of include/voxpop.h): the L2 read
// bandwidth this B200 sustains, the denominator for the sector-granular traffic figure of SURVEY.md 8(d).
// A buffer that fits the L2 is read over and over with 128-bit loads that bypass L1 (ld.global.cg), by a
// grid of a multiple of the SM count; time is taken with CUDA events on the launching stream.
#include <cuda_runtime.h>
#include <cstdint>

namespace {

__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ buf, size_t n16, int reps, uint32_t* sink)
{
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent loads in flight per thread
    for (; i + 3 * stride < n16; i += 4 * stride) {
      const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
      acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n16; i += stride) { const uint4 a = __ldcg(buf + i); acc += a.x ^ a.y ^ a.z ^ a.w; }
  }
  if (acc == 0x9E3779B9u) *sink = acc;  // keeps the loads alive; practically never true
}

}  // namespace

extern "C" int vp_probe_l2_read(int device, size_t bytes, int reps, double* gb_per_s)
{
  if (!gb_per_s || bytes < (1u << 20) || reps < 1) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 2;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 2;
  uint4* buf = nullptr; uint32_t* sink = nullptr;
  const size_t n16 = bytes / 16;
  if (cudaMalloc(&buf, n16 * 16) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { cudaFree(buf); return 3; }
  cudaMemset(buf, 1, n16 * 16);
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8;
  k_l2_read<<<blocks, 256, 0, st>>>(buf, n16, 2, sink);  // warm: pulls the buffer into L2
  double best = 0;
  for (int t = 0; t < 5; t++) {
    cudaEventRecord(e0, st);
    k_l2_read<<<blocks, 256, 0, st>>>(buf, n16, reps, sink);
    cudaEventRecord(e1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { best = 0; break; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)n16 * 16.0 * reps / (ms * 1e-3) / 1e9;
    if (g > best) best = g;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaStreamDestroy(st); cudaFree(buf); cudaFree(sink);
  *gb_per_s = best;
  return best > 0 ? 0 : 4;
}
