/* hx_diag.cu -- device microbenchmarks behind the roofline numbers bench.py reports.
 *
 * hx_measure_fp64_peak: DFMA throughput of the device.  SURVEY.md section 8(d) ("Which roofline
 * binds"): the path is FP64 SIMT work, the FP64 peak is not in MEASURED_PEAKS.json and has to be
 * measured.  Every thread runs HX_FMA_CHAINS independent dependent-FMA chains (enough
 * instruction-level parallelism to cover the pipe's latency at any occupancy), the grid fills
 * every SM with as many CTAs as fit, timing by CUDA events around the launch.
 *
 * hx_measure_hbm_copy: device-to-device copy bandwidth (read + write), the same quantity as
 * MEASURED_PEAKS.json's hbm_gbs -- a live cross-check of the denominator on the box at hand.
 */
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/hector_b200.h"

namespace {

#define HX_FMA_CHAINS 8

__global__ void __launch_bounds__(256)
fp64_fma_kernel(double *sink, int iters, double a, double b) {
  double x[HX_FMA_CHAINS];
#pragma unroll
  for (int k = 0; k < HX_FMA_CHAINS; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int k = 0; k < HX_FMA_CHAINS; ++k) x[k] = fma(x[k], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < HX_FMA_CHAINS; ++k) s += x[k];
  if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s; /* never true: keeps the chains */
}

__global__ void copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}

} // namespace

extern "C" int hx_measure_fp64_peak(int32_t device, double *tflops, double *fma_per_clk_per_sm) {
  if (!tflops) return HX_ERR_ARG;
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess) return HX_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return HX_ERR_CUDA;
  int sms = 0, per_sm = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fp64_fma_kernel, 256, 0);
  if (per_sm < 1) per_sm = 1;
  const int grid = sms * per_sm;
  double *sink = nullptr;
  if (cudaMalloc(&sink, (size_t)grid * 256 * sizeof(double)) != cudaSuccess) return HX_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 4096; /* x 8 unrolled x 8 chains = 262 144 FMAs per thread, ~20 ms */
  fp64_fma_kernel<<<grid, 256>>>(sink, 64, 1.0000001, 1e-9); /* warm-up */
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fp64_fma_kernel<<<grid, 256>>>(sink, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)grid * 256.0 * iters * 8.0 * HX_FMA_CHAINS;
    const double tf = 2.0 * fmas / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  const cudaError_t err = cudaGetLastError();
  cudaSetDevice(prev);
  if (err != cudaSuccess || best == 0.0) return HX_ERR_CUDA;
  *tflops = best;
  /* per SM and clock at the nominal maximum SM clock (the run may clock lower: lower bound) */
  if (fma_per_clk_per_sm) *fma_per_clk_per_sm = best * 1e12 / 2.0 / ((double)khz * 1e3) / sms;
  return HX_OK;
}

extern "C" int hx_measure_hbm_copy(int32_t device, double *gbs) {
  if (!gbs) return HX_ERR_ARG;
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess) return HX_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return HX_ERR_CUDA;
  const size_t bytes = (size_t)1 << 30; /* 1 GiB each way, far beyond the 126 MB L2 */
  double2 *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
    cudaFree(a);
    cudaSetDevice(prev);
    return HX_ERR_CUDA;
  }
  cudaMemset(a, 0, bytes);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t n = bytes / sizeof(double2);
  copy_kernel<<<sms * 8, 256>>>(a, b, n);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    copy_kernel<<<sms * 8, 256>>>(a, b, n);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
    if (g > best) best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  const cudaError_t err = cudaGetLastError();
  cudaSetDevice(prev);
  if (err != cudaSuccess || best == 0.0) return HX_ERR_CUDA;
  *gbs = best;
  return HX_OK;
}
