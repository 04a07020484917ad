// Aggregation multigrid preconditioner on the node graph (BSR, NB x NB blocks kept at every level).
//
// Replaces, for this path, the Teko block-Gauss-Seidel / MueLu smoothed-aggregation stack the
// reference configures through Belos (src/linear_solve.cpp:74-105, decks
// test/primal/notch_small_J2.yaml.in:54-98).  The hierarchy is plain (unsmoothed) aggregation:
//   P        piecewise-constant per node block (every fine node injects into its aggregate with an
//            identity NB x NB block), so the Galerkin operator P^T A P is the sum of the fine blocks
//            of each aggregate pair and keeps the BSR format;
//   pattern  of every level depends on the mesh only -> built once on the host; the numeric set-up
//            per matrix is a gather-sum kernel per level (HBM-bound) + block-Jacobi inverses + a
//            dense inverse of the coarsest operator (one CTA);
//   smoother damped block-Jacobi (the NB x NB nodal block couples u and p of the stabilised mixed
//            formulation), nu_pre / nu_post sweeps;
//   cycle    V, optionally over-corrected.
// The preconditioner is a fixed linear operator, so plain right-preconditioned GMRES applies.
// In a partitioned run with one of the library's transports (NCCL / host-staged) the hierarchy
// spans the parts: aggregates stay inside a part, every level keeps an owned + ghost numbering with
// its own halo plan, and once a level is small it is replicated on every part (amg_host.hpp).  With
// the caller's own communication hooks it acts on the owned x owned block of the part.
#pragma once
#include <vector>

#include "linalg.cuh"

namespace c8 {

struct AmgLevel {
  int n = 0;        // nodes = block rows
  int nnzb = 0;
  int ld = 0;       // nodes in a vector of this level (owned + ghost)
  int halo_level = -1;   // >= 0: partitioned level, plan id for comm_halo_level
  const int* rowptr = nullptr;   // device
  const int* colind = nullptr;
  const double* vals = nullptr;  // level 0: the caller's matrix
  int* own_rowptr = nullptr;
  int* own_colind = nullptr;
  double* own_vals = nullptr;
  double* dinv = nullptr;
  double *x = nullptr, *b = nullptr, *r = nullptr;  // level >= 1 (level 0 uses the caller's x, b)
  double* xt = nullptr;          // ping-pong partner of x for the out-of-place sweeps
  float* vals32 = nullptr;       // fp32 copy of the fine-level values (level 0 only)
  // transfer to the next coarser level
  int nc = 0;                              // rows of the coarse level
  int npc = 0;                             // columns with an aggregate (n, or ld when the hierarchy spans the parts)
  bool coarse_replicated = false;          // coarse level global + replicated: values and rhs summed over the parts
  int* agg = nullptr;                      // [npc] aggregate of a node = index into the coarse vector
  int *aggptr = nullptr, *aggmem = nullptr;  // members of an aggregate
  int *cptr = nullptr, *cmem = nullptr;      // fine blocks summed into a coarse block
  int n_cmem = 0;
};

struct AmgOptions {
  int nu_pre = 2, nu_post = 2;
  // block-Jacobi damping.  Time to solution of the 1 M-tet hyper-J2 forward solve (tools/amg_sweep.py, ms per load
  // step): 0.6: 124.9, 0.7: 118.9, 0.8: 112.4, 0.9: 108.5, 1.0: the smoother amplifies the highest modes and
  // GMRES stalls.  0.8 keeps a margin to that edge on other meshes / materials.
  double omega = 0.8;
  double over_correction = 1.6; // plain aggregation under-corrects; measured best 1.5-1.8
  int coarsest_max_nodes = 40;
  int max_levels = 12;
  int coarse_aggregate_size = 12;  // aggregate bound on the coarse levels (0: root + all neighbours); 8 -> 12: -3 %
  int coarse_nu = 0;            // sweeps per side on levels >= 1 (0: same as the fine level)
  bool fp32_fine_level = true;  // the fine-level sweeps read an fp32 copy of the matrix
  int max_aggregate_size = 8;  // bounded compact aggregates (0: root + all neighbours, ~25 nodes in 3-D)
  bool distributed = true;      // partitioned run: hierarchy across the parts (library transports only)
  int replicate_max_nodes = 30000;  // a level with at most this many nodes GLOBALLY is replicated
};

class Amg {
 public:
  explicit Amg(c8_ctx* ctx) : ctx_(ctx) {}
  ~Amg();
  AmgOptions opt;
  int build();                   // host: aggregates + coarse patterns from the context's BSR pattern
  int setup(const double* A);    // device: Galerkin sums, block-Jacobi inverses, coarsest inverse
  void apply(const double* r, double* z);  // z = M^-1 r, both [n_nodes_local][NB]
  int num_levels() const { return int(lv_.size()); }
  const std::vector<AmgLevel>& levels() const { return lv_; }
  double operator_complexity() const;
  bool distributed() const { return dist_; }

 private:
  void cycle(int l, const double* b, double* x);
  void smooth(int l, const double* b, double* x, int sweeps, bool zero_guess);
  void sweep(int l, const double* b, const double* xin, double* xout, const double* xc,
             const double* xprev = nullptr, double c1 = 0.0, double c2 = 0.0);
  void halo(int l, const double* v);
  c8_ctx* ctx_;
  int nb_ = 0;
  bool dist_ = false;
  std::vector<AmgLevel> lv_;
  int nd_ = 0;                 // coarsest dense size (dofs)
  double* dense_ = nullptr;    // [nd][2 nd] work, inverse in the right half
  double* r0_ = nullptr;       // level-0 residual
};

}  // namespace c8
