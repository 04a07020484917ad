// POD argument blocks of the adjoint / QoI kernels (shared with the host-side context).
#pragma once
#include "args.h"

namespace c8 {

// QoI<T> classes of the reference: avg_disp.cpp, calibration.cpp, reaction_mismatch.cpp,
// load_mismatch.cpp, surface_mismatch.cpp
enum QoiType { QOI_AVG_DISP = 0, QOI_CALIBRATION = 1, QOI_REACTION_MISMATCH = 2, QOI_LOAD_MISMATCH = 3,
               QOI_SURFACE_MISMATCH = 4 };

struct QoiArgs {
  int type;
  // calibration
  double weights[3];
  double balance_factor;
  double dt_over_T;        // m_dt / m_total_time
  double inv_area;         // 1 / m_area
  double load_mismatch;    // total_load - load_meas of the step (set after the preprocess pass)
  int coord_idx;
  double coord_value, coord_tol;
  int reaction_force_comp;
  const double* measured;      // [n_nodes][DIM] measured displacement of the step
  const signed char* facet;    // [n_elems][3] local vertex ids of the facet on the side set, -1 none (2-D: 2 ids)
  int compute_torque;          // reaction mismatch "compute torque": moment about axis reaction_force_comp
  double normal_2d[2];         // load mismatch "2D surface normal"
  // which integrands the type carries
  __host__ __device__ bool has_disp() const { return type == QOI_CALIBRATION || type == QOI_SURFACE_MISMATCH; }
  __host__ __device__ bool has_node_load() const { return type == QOI_CALIBRATION || type == QOI_REACTION_MISMATCH; }
  __host__ __device__ bool has_face_load() const { return type == QOI_LOAD_MISMATCH; }
};

struct AdjArgs {
  MeshArgs mesh;
  ModelArgs model;
  QoiArgs qoi;
  const double* x;
  const double* x_prev;
  const double* xi;
  const double* xi_prev;
  long long xi_ld;
  double* g;          // K3: in/out ; K4: in (g) / out (new g)
  double* f;          // K3: in ; K4: out        [NX][xi_ld], element dofs node-interleaved
  double* vals;       // K3: A^T (overwritten)
  double* emat;       // K3: element-matrix scratch of the two-phase assembly
  double* b;          // K3: rhs (+=)
  const double* z;    // K4/K6: nodal adjoint [n_nodes][NB]
  double* phi;        // K4: out ; K6: in
  double* grad;       // K6: [n_es][NPAR] (+=), derivative w.r.t. EVERY parameter of the model
  double* scalars;    // K5: [0] += J, [1] += total load
  int* tile_counter;  // K3 (persistent): device counter of the dynamic tile list, zeroed before the launch
};

struct VfmArgs {
  MeshArgs mesh;
  ModelArgs model;
  const double* x;        // measured displacement of the step
  const double* x_prev;   // measured displacement of step-1
  const double* xi_prev;
  double* xi;             // K7: in (current-field values) / out ; K8: in
  long long xi_ld;
  double* b;              // K7: internal force residual (+=)
  double* dR;             // K7 (grad): [NPAR][n_dofs] (+=) or nullptr
  double* local_sens;     // K7 (grad): [NXI*NPAR][ld] in/out
  int* n_failed;
  const double* w;        // K8: virtual field [n_nodes][NB]
  double* hist;           // K8: [NXI][ld] in/out
  double s;               // K8: scaled virtual power mismatch
  double* grad;           // K8: [NPAR] (+=)
};

}  // namespace c8
