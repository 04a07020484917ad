// Host-side descriptors of the residual objects, with the reference's names and factories:
//   LocalResidual  metadata  src/local_residual.hpp:41-53 (num_residuals, var_type, num_eqs, resid_name,
//                            param_names) as set by each model's constructor (e.g. src/small_J2.cpp:36-60)
//   create_local_residual    src/local_residual.cpp:892-933  (YAML `type` string -> model)
//   GlobalResidual metadata  src/global_residual.hpp, src/mechanics.cpp:18-47, src/mechanics_plane_stress.cpp
//   create_global_residual   src/global_residual.cpp:619-630
// The arithmetic of these objects lives in the device templates of csrc/models.cuh / mechanics.cuh;
// here is what the host needs to size, name and pack their variables (get_num_eqs: src/fields.cpp:11-18).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "c8b200.h"

namespace c8host {

enum VarType { SCALAR = 0, VECTOR = 1, SYM_TENSOR = 2, TENSOR = 3 };  // src/defines.hpp

inline int get_num_eqs(int var_type, int ndims) {  // src/fields.cpp:11-18
  switch (var_type) {
    case SCALAR: return 1;
    case VECTOR: return ndims;
    case SYM_TENSOR: return ndims * (ndims + 1) / 2;
    case TENSOR: return ndims * ndims;
  }
  throw std::runtime_error("get_num_eqs: unknown variable type");
}

struct LocalResidual {
  std::string type;
  int c8_type = -1;                       // C8_* constant of include/c8b200.h
  int ndims = 0;
  std::vector<std::string> resid_names;   // resid_name(i)
  std::vector<int> var_types, num_eqs;    // var_type(i), num_eqs(i)
  std::vector<std::string> param_names;   // param_names(), the order of init_params
  bool finite_deformation = false;        // is_finite_deformation()
  int z_stretch_idx = -1;                 // hyper_J2_plane_stress: the lambda_z residual
  int num_residuals() const { return int(resid_names.size()); }
  int num_dofs() const { int n = 0; for (int e : num_eqs) n += e; return n; }
  int num_params() const { return int(param_names.size()); }
};

inline LocalResidual create_local_residual(const std::string& type, int ndims) {
  LocalResidual r;
  r.type = type; r.ndims = ndims;
  auto add = [&](const char* name, int vt) {
    r.resid_names.push_back(name); r.var_types.push_back(vt); r.num_eqs.push_back(get_num_eqs(vt, ndims));
  };
  if (type == "elastic") {
    r.c8_type = C8_ELASTIC; add("dummy", SCALAR);
    r.param_names = {"E", "nu", "cte", "delta_T"};
  } else if (type == "small_J2") {
    r.c8_type = C8_SMALL_J2; add("pstrain", SYM_TENSOR); add("alpha", SCALAR);
    r.param_names = {"E", "nu", "K", "Y", "cte", "delta_T"};
  } else if (type == "small_hill") {
    r.c8_type = C8_SMALL_HILL; add("pstrain", SYM_TENSOR); add("alpha", SCALAR);
    r.param_names = {"E", "nu", "Y", "R00", "R11", "R22", "R01", "R02", "R12", "S", "D"};
  } else if (type == "small_hill_plane_stress" || type == "small_hill_plane_strain") {
    r.c8_type = type == "small_hill_plane_stress" ? C8_SMALL_HILL_PLANE_STRESS : C8_SMALL_HILL_PLANE_STRAIN;
    add("pstrain", SYM_TENSOR); add("alpha", SCALAR);
    r.param_names = {"E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01"};
  } else if (type == "hyper_J2") {
    r.c8_type = C8_HYPER_J2; r.finite_deformation = true;
    add("zeta", SYM_TENSOR); add("Ie", SCALAR); add("alpha", SCALAR);
    r.param_names = {"E", "nu", "Y", "S", "D", "A", "n", "K"};
  } else if (type == "hyper_J2_plane_stress") {
    r.c8_type = C8_HYPER_J2_PLANE_STRESS; r.finite_deformation = true;
    add("zeta", SYM_TENSOR); add("Ie", SCALAR); add("lambda_z", SCALAR); add("alpha", SCALAR);
    r.z_stretch_idx = 2;
    r.param_names = {"E", "nu", "Y", "S", "D", "A", "n", "K"};
  } else if (type == "hyper_J2_plane_strain") {
    r.c8_type = C8_HYPER_J2_PLANE_STRAIN; r.finite_deformation = true;
    add("zeta", SYM_TENSOR); add("Ie", SCALAR); add("alpha", SCALAR);
    r.param_names = {"E", "nu", "K", "Y", "Y_inf", "delta"};
  } else if (type == "hypo_hill") {                      // src/hypo_hill.cpp:43-70
    r.c8_type = C8_HYPO_HILL; r.finite_deformation = true;
    add("TC", SYM_TENSOR); add("alpha", SCALAR);
    r.param_names = {"E", "nu", "Y", "R00", "R11", "R22", "R01", "R02", "R12", "S", "D"};
  } else if (type == "hypo_hill_plane_strain") {         // src/hypo_hill_plane_strain.cpp:40-73
    r.c8_type = C8_HYPO_HILL_PLANE_STRAIN; r.finite_deformation = true;
    add("TC", SYM_TENSOR); add("alpha", SCALAR); add("TC_zz", SCALAR);
    r.param_names = {"E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01"};
  } else if (type == "hypo_hill_plane_stress") {         // src/hypo_hill_plane_stress.cpp:44-77
    r.c8_type = C8_HYPO_HILL_PLANE_STRESS; r.finite_deformation = true;
    add("TC", SYM_TENSOR); add("alpha", SCALAR); add("lambda_z", SCALAR);
    r.z_stretch_idx = 2;
    r.param_names = {"E", "nu", "Y", "S", "D", "R00", "R11", "R22", "R01", "Q00", "Q01", "Q10", "Q11"};
  } else {
    throw std::runtime_error("create_local_residual: type '" + type + "' is not in the hot-path scope");
  }
  if ((type == "small_hill" || type == "hypo_hill") && ndims != 3)
    throw std::runtime_error("create_local_residual: " + type + " needs a 3-D mesh");
  if ((type.find("plane") != std::string::npos) && ndims != 2)
    throw std::runtime_error("create_local_residual: " + type + " needs a 2-D mesh");
  return r;
}

struct GlobalResidual {
  std::string type;
  int c8_type = -1;
  int ndims = 0;
  std::vector<std::string> resid_names;   // "u"[, "p"]
  std::vector<int> var_types, num_eqs;
  int num_ip_sets = 1;                    // mixed: the coupled point + the order-2 pressure-mass set
  int num_residuals() const { return int(resid_names.size()); }
  int num_node_dofs() const { int n = 0; for (int e : num_eqs) n += e; return n; }   // NB
};

inline GlobalResidual create_global_residual(const std::string& type, int ndims, bool mixed = true) {
  GlobalResidual g;
  g.type = type; g.ndims = ndims;
  if (type == "mechanics") {
    if (!mixed) {
      // "mixed formulation: false" (src/mechanics.cpp:18-54): ONE residual u, one ip set, the momentum
      // balance with local->cauchy().  The cauchy() of the mixed-type models of this path reads the
      // pressure residual (src/small_J2.cpp:252-263 ...), which does not exist in this mode, so the only
      // in-scope local residuals it can drive are the plane-stress ones, whose cauchy() is the full
      // stress: sum_j sigma_ij dN/dX_j w dv == mechanics_plane_stress with unit thickness (small strain;
      // the finite-strain plane-stress model differs by the z stretch and is rejected by the caller).
      if (ndims != 2)
        throw std::runtime_error("create_global_residual: displacement-only mechanics needs a local residual whose "
                                 "cauchy() does not read the pressure (the plane-stress models, 2-D)");
      g.c8_type = C8_MECHANICS_PLANE_STRESS;
      g.resid_names = {"u"}; g.var_types = {VECTOR}; g.num_eqs = {2};
      return g;
    }
    g.c8_type = C8_MECHANICS;
    g.resid_names = {"u", "p"}; g.var_types = {VECTOR, SCALAR};
    g.num_eqs = {get_num_eqs(VECTOR, ndims), 1};
    g.num_ip_sets = 2;
  } else if (type == "mechanics_plane_stress") {
    if (ndims != 2) throw std::runtime_error("create_global_residual: mechanics_plane_stress needs a 2-D mesh");
    g.c8_type = C8_MECHANICS_PLANE_STRESS;
    g.resid_names = {"u"}; g.var_types = {VECTOR}; g.num_eqs = {2};
  } else {
    throw std::runtime_error("create_global_residual: type '" + type + "' is not in the hot-path scope");
  }
  return g;
}

}  // namespace c8host
