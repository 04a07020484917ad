// C-ABI implementation, part 2: resident step state (the role of Disc::primal(step) fields
// for the step being solved, src/disc.cpp:643-683) and device micro-benchmarks used to
// measure the roofline denominators on the box (bench.py).
#include "c8b200.h"
#include "context.cuh"

namespace c8 {

// fp64 FMA throughput: 8 independent chains per thread
__global__ void k_dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4.,
         a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
  const double m = 1.0000001, c = 1e-7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// STREAM-style copy for an HBM reference measured with the same timer as the kernels
__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

}  // namespace c8

using namespace c8;

extern "C" {

int c8_state_set_prev(c8_ctx* ctx, const double* u_prev, const double* p_prev,
                      const double* xi_prev) {
  C8_REQUIRE(ctx, ctx->kt != nullptr, "c8_set_model must be called first");
  int rc;
  if ((rc = c8_pack_x(ctx, u_prev, p_prev, ctx->d_xp)) != C8_OK) return rc;
  if ((rc = c8_pack_xi(ctx, xi_prev, ctx->d_xip)) != C8_OK) return rc;
  // create_primal(step) starts the new step's fields as a copy of step-1 (src/disc.cpp:643-683)
  C8_CUDA(ctx, cudaMemcpyAsync(ctx->d_xi, ctx->d_xip,
                               size_t(ctx->xi_ld) * ctx->kt->nxi * sizeof(double),
                               cudaMemcpyDeviceToDevice, ctx->stream));
  C8_CUDA(ctx, cudaMemcpyAsync(ctx->d_x, ctx->d_xp, size_t(ctx->n_nodes) * ctx->kt->nb * sizeof(double),
                               cudaMemcpyDeviceToDevice, ctx->stream));
  return C8_OK;
}

int c8_state_forward_jacobian(c8_ctx* ctx, const double* u, const double* p, double* b_u,
                              double* b_p, int* n_failed) {
  return forward_state_host(ctx, u, p, b_u, b_p, n_failed);
}

int c8_state_get_xi(c8_ctx* ctx, double* xi_host) { return c8_unpack_xi(ctx, ctx->d_xi, xi_host); }

int c8_state_ptrs(c8_ctx* ctx, double** x, double** x_prev, double** xi, double** xi_prev,
                  double** A, double** b) {
  if (x) *x = ctx->d_x;
  if (x_prev) *x_prev = ctx->d_xp;
  if (xi) *xi = ctx->d_xi;
  if (xi_prev) *xi_prev = ctx->d_xip;
  if (A) *A = ctx->d_A;
  if (b) *b = ctx->d_b;
  return C8_OK;
}

int c8_bench_dfma(c8_ctx* ctx, int iters, double* tflops) {
  int sms = 0;
  C8_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  const int blocks = sms * 8, threads = 256;
  double* out = stage(ctx, size_t(blocks) * threads * sizeof(double));
  C8_REQUIRE(ctx, out != nullptr, "staging allocation failed");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(out, 16);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(out, iters);
    cudaEventRecord(e1, ctx->stream);
    C8_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * double(iters) * blocks * threads;
    best = std::max(best, flops / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = best;
  return C8_OK;
}

int c8_bench_copy(c8_ctx* ctx, double* gbs) {
  const size_t n = size_t(1) << 27;  // 2 x 2 GiB of double2
  double2 *a = nullptr, *b = nullptr;
  C8_CUDA(ctx, cudaMalloc(&a, n * sizeof(double2)));
  C8_CUDA(ctx, cudaMalloc(&b, n * sizeof(double2)));
  cudaMemsetAsync(a, 0, n * sizeof(double2), ctx->stream);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    k_copy<<<sms * 16, 512, 0, ctx->stream>>>(a, b, n);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::max(best, 2.0 * n * sizeof(double2) / (ms * 1e-3) * 1e-9);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(a); cudaFree(b);
  *gbs = best;
  return C8_OK;
}

}  // extern "C"

