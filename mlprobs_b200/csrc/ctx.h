// Internal: the context object behind the opaque mlp_ctx handle (shared by capi.cu and exchange.cu).
#pragma once
#include "../../include/mlprobs_b200.h"
#include "posterior.cuh"
#include <string>
#include <vector>

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (void)cudaGetLastError();   /* clear the per-thread error so that it cannot surface at a later launch */ \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return MLP_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

static const int kCmaxLimit = 8;    // columns per lane (= MLP_SWEEP_MAXC, the unroll width of sweep.cuh); 32*8 = 256 columns per column block

struct mlp_ctx {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: partition-function sweeps, overlapped with the HMM sweeps
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // streamed posterior stage (mlp_stream_begin, large families): every batch's matrices are finished (QuickProbs' re-quantisation of a
    // consistency repetition without third sequences), digested and dropped, so the cell pool only ever holds one batch
    bool stream_mode = false; int stream_reps = 1; unsigned long long* d_sdigest = nullptr; long long sdigest_cap = 0; int* d_len = nullptr; int len_cap = 0;
    bool restricted = false;                             // shard cut down by mlp_restrict_pairs (a stage over it keeps the resident guide tree)
    long long own_cells = 0;                             // cells of the set before mlp_exchange_needed appended its imports
    bool loc_old = false;                                // local model: fall back to the round-1 kernels (set when loc_c.cu's checked bound fails)
    int overlap = 0, bps_part = 0, bps_hmm = 0;          // tuning knobs (MLP_OVERLAP, MLP_BPS_PART, MLP_BPS_HMM); measured: no gain, off
    std::string err;
    // configuration
    int64_t scratch_budget = 0, cell_capacity_req = 0;
    // sequences
    int n = 0;
    std::vector<int> len;
    std::vector<long long> seq_off;
    long long total_res = 0;
    std::vector<uint8_t> codes_h;
    uint8_t* d_res = nullptr;
    long long* d_seq_off = nullptr;
    // tables
    bool have_tables = false;
    mlp_hmm_tables hmm;
    mlp_part_tables part;
    float* d_match = nullptr; float* d_ins = nullptr; double* d_sub = nullptr;
    // pairs
    std::vector<PairTask> all_pairs;     // cost-sorted (descending)
    std::vector<PairTask> owned;         // this shard
    std::vector<PairTask> relax_tasks;   // owned pairs in the tile order mlp_relax processes them (cached)
    int relax_tasks_n = 0;
    PairTask* d_relax_tasks = nullptr; size_t relax_tasks_cap = 0; bool relax_tasks_on_device = false;
    int rank = 0, world = 1;
    // sparse sets (double buffered for relax)
    std::vector<long long> rp_off_h;
    long long rp_total = 0;
    long long* d_rp_off = nullptr;
    CsrSetDev set[2] = {};
    long long rp_cap = 0, nn_cap = 0;    // allocated sizes of the pooled arrays (ints per rp_pool, entries per n*n array): a context
                                         // that sees many families keeps its pools and only re-allocates when one does not fit
    long long res_cap = 0; int seqoff_cap = 0;
    void* tail_prov = nullptr;           // the tail's profile-posterior provider with its staging buffers (qp_tail_dev.cu), kept across families
    void (*tail_prov_free)(void*) = nullptr;
    int cur = 0;
    bool have_sets = false;
    int flavour_of_set = -1;
    float* d_dist = nullptr;
    // per-launch scratch
    void* d_scratch = nullptr; size_t scratch_bytes = 0;
    PairTask* d_tasks = nullptr; PairOut* d_pout = nullptr; size_t tasks_cap = 0;
    // host-mapped words the device publishes the error word and a set's cursor into: the host reads them after a stream sync without
    // a device->host copy (a small copy would queue behind a split read-back on the DMA engine and stall the stage for its whole length)
    unsigned long long* h_flags = nullptr; unsigned long long* d_hflags = nullptr;
    int* d_counter = nullptr; int* d_err = nullptr;   // d_counter: 16 work-queue heads, one per kernel id
    int4* d_stage = nullptr; int stage_cap = 0; long long stage_warps = 0;
    int* d_rowexp = nullptr; size_t rowexp_cap = 0;
    float* d_rowaux = nullptr; size_t rowaux_cap = 0;   // local model: per task row sums / prefix bounds / candidate counts of the Z chain (loc_c.cu)
    int* d_tfill = nullptr; long long tfill_stride = 0, tfill_warps = 0;
    void* d_edge = nullptr; long long edge_stride = 0, edge_warps = 0;
    float* d_wk = nullptr; long long wk_warps = 0;
    float* d_weights = nullptr; float* d_seldist = nullptr; int weights_cap = 0;
    void* d_tree = nullptr; int tree_cap = 0;            // device guide tree (tree_dev.cu): scratch matrix + tree arrays
    bool tree_resident = false;                          // d_weights / d_seldist hold the tree of the current distance matrix (mlp_qp_guide_tree_device)
    // nccl
    void* nccl_comm = nullptr; int comm_rank = 0, comm_world = 1;
    int* d_xcnt = nullptr; size_t xcnt_cap = 0;          // gathered per-matrix cell counts (selective exchange)
    void* d_ximp = nullptr; size_t ximp_cap = 0;         // import list of the selective exchange
    bool imported = false;                               // set holds foreign matrices imported by mlp_exchange_needed (only the owned pairs are this rank's result)
    unsigned long long* d_xused = nullptr;               // per-rank cell counts of an exchange
    unsigned* d_xq = nullptr; size_t xq_cap = 0;          // packed wire buffer of the QuickProbs exchange
    bool set_partial = false, dist_partial = false;       // sharded stage output not yet exchanged (set by posterior / relax, cleared by mlp_exchange)
    // split read-back (mlp_get_csr_packed_begin / _end): own stream and pack buffers, so that the copy of one step's result runs
    // beside the posterior stage of the next; rb_set = the set being read (-1: none)
    cudaStream_t stream_rb = nullptr; cudaEvent_t ev_rb = nullptr; int rb_set = -1;
    void* d_pack = nullptr; size_t pack_bytes = 0;
    // sharded sets: only the row sizes of the owned matrices cross PCIe (compact, in owned-list order) and are scattered into the
    // caller's fixed-layout table by _end
    unsigned short* h_rs_stage = nullptr; size_t rs_stage_cap = 0; long long* d_rs_off = nullptr; size_t rs_off_cap = 0; PairTask* d_rs_tasks = nullptr; size_t rs_tasks_cap = 0;
    std::vector<long long> rs_off_h; uint16_t* rb_row_sizes = nullptr; long long rb_rs_total = 0;
    cudaEvent_t ev_dist = nullptr; bool exch_pending = false; unsigned long long exch_total = 0;   // split exchange (mlp_exchange_begin / _end)
    // stats
    mlp_stage_stats stats = {};
    cudaEvent_t ev[2] = {nullptr, nullptr};
};


void free_dev(void* p);
int grow_cells(mlp_ctx* ctx, int which, long long new_cap, unsigned long long keep);
