// Services of comm.cu for the partitioned multigrid (amg.cu): the level-0 halo plan given to
// c8_set_halo_plan, plans of the coarser levels, and the halo copy of any level through the bound
// library transport (NCCL or host-staged).
#pragma once
#include "amg_host.hpp"

struct c8_ctx;

namespace c8 {
// true when a plan and one of the library's own transports are bound on a run with > 1 parts
bool comm_library_transport(c8_ctx* ctx);
int comm_rank(c8_ctx* ctx);
int comm_nranks(c8_ctx* ctx);
HaloPlanHost comm_plan(c8_ctx* ctx, int level);
int comm_add_level(c8_ctx* ctx, const HaloPlanHost& plan);   // returns the level id (>= 1), -1 on error
void comm_drop_levels(c8_ctx* ctx);
void comm_halo_level(c8_ctx* ctx, int level, double* vec_dev, int nb);
}  // namespace c8
