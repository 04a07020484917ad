// Instantiates every kernel of the hot path for one (DIM, MECH, Model, G) combination.
#pragma once
// threads per quadrature point for the 3-D mixed u-p combinations (NX = 16 element dofs): 4 threads
// x 4 derivative lanes measured fastest on B200 (profiles/README.md: fewer replicated value
// instructions than 8 x 2, and a smaller program)
#ifndef C8_G3D
#define C8_G3D 4
#endif
#include <cstdio>
#include <cstdlib>
#include "vfm.cuh"
#include "kernel_table.h"

namespace c8 {

template <class C>
__global__ void k_init_xi(double* xi, long long xi_ld, int n_elems) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  double v[C::NXI];
  C::Model::init(v);
#pragma unroll
  for (int q = 0; q < C::NXI; ++q) xi[size_t(q) * xi_ld + e] = v[q];
}

// C8_K1_PERSISTENT=0 in the environment selects the one-tile-per-CTA element kernel (read once)
inline bool k1_persistent_enabled() {
  static const int on = [] { const char* v = getenv("C8_K1_PERSISTENT"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}

// C8_K3_PERSISTENT=0 selects the one-tile-per-CTA K3
inline bool k3_persistent_enabled() {
  static const int on = [] { const char* v = getenv("C8_K3_PERSISTENT"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}

template <class C>
struct Launch {
  // phase 2 of the assembly (forward.cuh): BSR values of the owned rows <- element matrices
  template <bool TRANSPOSE>
  static void gather(const MeshArgs& m, const double* emat, double* vals, cudaStream_t s) {
    constexpr int EPT = (C::NB % 2 == 0) ? 2 : 1;  // entries per thread (forward.cuh)
    const long long work = (long long)m.n_row_blocks * (C::NB * C::NB / EPT);
    if (work == 0) return;
    k_bsr_gather<C::NB, C::NN, TRANSPOSE><<<(unsigned)((work + 255) / 256), 256, 0, s>>>(
        m.gptr, m.gsrc, emat, vals, m.n_row_blocks);
  }
  static void forward_jacobian(const FwdArgs& a0, cudaStream_t s) {
    if (a0.mesh.n_elems == 0) return;
    FwdArgs a = a0;
    const int block = C8_K1_BLOCK;
    // FAST: the production call (matrix + residual, no element-level output) is branch-free
    const bool fast = a.vals && a.b && !a.elem_J && !a.elem_R;
    if constexpr (C::NB % 2 == 0) {
      if (a.vals && a.mesh.chunk_elems > 0 && a.mesh.n_chunks > 1 && a.cg_ptr_host) {
        // chunked two-phase assembly: element kernel of chunk c, then its gather while the chunk's
        // element matrices are still in L2 (args.h)
        for (int c = 0; c < a.mesh.n_chunks; ++c) {
          a.elem_begin = c * a.mesh.chunk_elems;
          a.elem_end = a.elem_begin + a.mesh.chunk_elems < a.mesh.n_elems ? a.elem_begin + a.mesh.chunk_elems : a.mesh.n_elems;
          const long long threads = (long long)(a.elem_end - a.elem_begin) * C::G;
          const unsigned grid = (unsigned)((threads + block - 1) / block);
          if (fast) k_forward_jacobian<C, true><<<grid, block, 0, s>>>(a);
          else k_forward_jacobian<C, false><<<grid, block, 0, s>>>(a);
          if (c + 1 == a.mesh.n_chunks && a.elements_done) cudaEventRecord(a.elements_done, s);
          const int en0 = a.cg_ptr_host[c], nen = a.cg_end_host[c] - en0;
          if (nen > 0) {
            const long long work = (long long)nen * (C::NB * C::NB / 2);
            k_bsr_gather_chunk<C::NB, C::NN><<<(unsigned)((work + 255) / 256), 256, 0, s>>>(
                a.mesh.cg_blk, a.mesh.cg_k, a.mesh.gsrc, a.emat, a.vals, en0, nen, a.elem_begin);
          }
        }
        return;
      }
    }
    a.elem_begin = 0; a.elem_end = a.mesh.n_elems;
    const long long threads = (long long)a.mesh.n_elems * C::G;
    const unsigned grid = (unsigned)((threads + block - 1) / block);
    if (fast && k1_persistent_enabled()) {
      // production call: persistent CTAs (one per SM) that prefetch the next tile's element records.
      // The shared-memory opt-in and the occupancy are per DEVICE (a process may hold contexts on several).
      constexpr int MAXDEV = 64;
      static int n_sm_dev[MAXDEV] = {0}, per_sm_dev[MAXDEV] = {0};
      int dev = 0;
      cudaGetDevice(&dev);
      const int di = dev < MAXDEV ? dev : MAXDEV - 1;
      constexpr int smem = k1_smem_bytes<C>();
      constexpr int pblock = K1PBlock<C>::value;
      if (!n_sm_dev[di] || dev >= MAXDEV) {
        int sm = 0, per = 1;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_forward_jacobian_persistent<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_forward_jacobian_persistent<C, true>, pblock, smem) !=
                cudaSuccess || per < 1)
          per = 1;
        n_sm_dev[di] = sm; per_sm_dev[di] = per;
      }
      const int n_sm = n_sm_dev[di], per_sm = per_sm_dev[di];
      const unsigned resident = (unsigned)(n_sm * per_sm);
      const unsigned tiles = (unsigned)((threads + pblock - 1) / pblock);
      const unsigned pgrid = tiles < resident ? tiles : resident;
      k_forward_jacobian_persistent<C, true><<<pgrid, pblock, smem, s>>>(a);
    } else if (fast) k_forward_jacobian<C, true><<<grid, block, 0, s>>>(a);
    else k_forward_jacobian<C, false><<<grid, block, 0, s>>>(a);
    if (a.elements_done) cudaEventRecord(a.elements_done, s);  // xi, b, path and the status are final here
    if (a.vals) gather<false>(a.mesh, a.emat, a.vals, s);
#ifdef C8_K1_PHASE_CLOCKS
    if (fast) {  // tuning builds: print and reset the per-phase warp-cycle counters
      unsigned long long h[8], z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      cudaStreamSynchronize(s);
      cudaMemcpyFromSymbol(h, g_k1_phase_clocks, sizeof(h));
      cudaMemcpyToSymbol(g_k1_phase_clocks, z, sizeof(z));
      double tot = 0;
      for (int k = 0; k < 5; ++k) tot += double(h[k]);
      fprintf(stderr, "K1 phases [load, P1 newton, P2 dC/dx+sens, P3 momentum, P3 pressure+scatter] %%:");
      for (int k = 0; k < 5; ++k) fprintf(stderr, " %.1f", 100.0 * double(h[k]) / tot);
      fprintf(stderr, "\n");
    }
#endif
  }
  static void global_residual(const FwdArgs& a, cudaStream_t s) {
    if (a.mesh.n_elems == 0) return;
    const int block = 128;
    k_global_residual<C><<<(a.mesh.n_elems + block - 1) / block, block, 0, s>>>(a);
  }
  static void init_xi(double* xi, long long xi_ld, int n_elems, cudaStream_t s) {
    if (n_elems == 0) return;
    k_init_xi<C><<<(n_elems + 255) / 256, 256, 0, s>>>(xi, xi_ld, n_elems);
  }
  static void adjoint_jacobian(const AdjArgs& a, cudaStream_t s) {
    if (a.mesh.n_elems == 0) return;
    const long long threads = (long long)a.mesh.n_elems * C::G;
    // 3-D combinations only: their one-tile K3 holds 8 warps per SM anyway (250+ registers); the 2-D ones run
    // at 116-230 registers with more resident warps than one 256-thread CTA would give
    // (measured at 1 M tets: hyper-J2 2.28 -> 2.04 ms, small-J2 1.78 -> 1.75, small-Hill 1.89 -> 1.86; the elastic model,
    // with no local solve, is faster with many small CTAs: 1.43 against 1.52)
    if (C::D == 3 && C::Model::HAS_NEWTON && k3_persistent_enabled() && a.tile_counter) {
      constexpr int MAXDEV = 64;
      static int n_sm_dev[MAXDEV] = {0};
      int dev = 0;
      cudaGetDevice(&dev);
      const int di = dev < MAXDEV ? dev : MAXDEV - 1;
      constexpr int smem = (int)sizeof(K3Smem<C>);
      if (!n_sm_dev[di] || dev >= MAXDEV) {
        int sm = 0;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_adjoint_jacobian_persistent<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        n_sm_dev[di] = sm;
      }
      const unsigned tiles = (unsigned)((threads + C8_K1_BLOCK - 1) / C8_K1_BLOCK);
      const unsigned pgrid = tiles < (unsigned)n_sm_dev[di] ? tiles : (unsigned)n_sm_dev[di];
      k_adjoint_jacobian_persistent<C><<<pgrid, C8_K1_BLOCK, smem, s>>>(a);
    } else {
      k_adjoint_jacobian<C><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
    }
    if (a.vals) gather<true>(a.mesh, a.emat, a.vals, s);
  }
  static void adjoint_local(const AdjArgs& a, cudaStream_t s) {
    if (a.mesh.n_elems == 0) return;
    const long long threads = (long long)a.mesh.n_elems * C::G;
    k_adjoint_local<C><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
  }
  static void qoi_gradient(const AdjArgs& a, cudaStream_t s) {
    if (a.mesh.n_elems == 0) return;
    const long long threads = (long long)a.mesh.n_elems * C::G;
    k_qoi_gradient<C><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
  }
  static void qoi_value(const AdjArgs& a, int mode, cudaStream_t s) {
    if (a.mesh.n_elems == 0) return;
    k_qoi_value<C><<<(a.mesh.n_elems + 127) / 128, 128, 0, s>>>(a, mode);
  }
  static void vfm_forward(const VfmArgs& a, cudaStream_t s) {
    if constexpr (C::M == MECH_PLANE_STRESS) {
      if (a.mesh.n_elems == 0) return;
      const long long threads = (long long)a.mesh.n_elems * C::G;
      k_vfm_forward<C><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
    }
  }
  static void vfm_adjoint(const VfmArgs& a, cudaStream_t s) {
    if constexpr (C::M == MECH_PLANE_STRESS) {
      if (a.mesh.n_elems == 0) return;
      const long long threads = (long long)a.mesh.n_elems * C::G;
      k_vfm_adjoint<C><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(a);
    }
  }
  static KernelTable table() {
    KernelTable t;
    t.dim = C::D; t.mech = C::M; t.local_type = C::Model::TYPE;
    t.nn = C::NN; t.nb = C::NB; t.nx = C::NX; t.nxi = C::NXI; t.npar = C::NPAR; t.group = C::G;
    t.finite = C::Model::FINITE;
    t.forward_jacobian = &forward_jacobian;
    t.global_residual = &global_residual;
    t.init_xi = &init_xi;
    t.adjoint_jacobian = &adjoint_jacobian;
    t.adjoint_local = &adjoint_local;
    t.qoi_gradient = &qoi_gradient;
    t.qoi_value = &qoi_value;
    t.vfm_forward = (C::M == MECH_PLANE_STRESS) ? &vfm_forward : nullptr;
    t.vfm_adjoint = (C::M == MECH_PLANE_STRESS) ? &vfm_adjoint : nullptr;
    return t;
  }
};

}  // namespace c8

#define C8_DEFINE_COMBO(NAME, DIM, MECH, MODEL, G)                                   \
  namespace c8 {                                                                     \
  const KernelTable* table_##NAME() {                                                \
    static const KernelTable t = Launch<Cfg<DIM, MECH, MODEL<DIM>, G>>::table();     \
    return &t;                                                                       \
  }                                                                                  \
  }
