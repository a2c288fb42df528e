// Multi-GPU exchange of the sparse posteriors -- filled in by exchange step; see DESIGN.md.
#include "../../include/mlprobs_b200.h"
extern "C" int mlp_nccl_unique_id(uint8_t id128[128]) { (void)id128; return MLP_E_UNSUPPORTED; }
extern "C" int mlp_comm_init(mlp_ctx* ctx, const uint8_t id128[128], int rank, int world) { (void)ctx; (void)id128; (void)rank; (void)world; return MLP_E_UNSUPPORTED; }
extern "C" int mlp_exchange(mlp_ctx* ctx) { (void)ctx; return MLP_E_UNSUPPORTED; }
