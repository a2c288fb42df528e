// Shared between posterior.cu (kernels), relax.cu and capi.cu (host driver).
#pragma once
#include "dev_common.cuh"
#include "../../include/mlprobs_b200.h"   // MLP_K_* kernel ids

#define MLP_BLOCK 128                  // 4 warps per CTA, one pair per warp
#define MLP_HMM_TABLE_BYTES 3072       // 676 match + 26 ins floats padded to 2816 B, then the 256-byte LogAddLut
#define MLP_PART_TABLE_BYTES 5408      // 676 doubles
#define MLP_FINAL_TABLE_BYTES 288      // ExpLut: 6 pieces x 6 doubles
#define MLP_K_VITERBI 9                // all-pairs 3-state Viterbi (model selection)
#define MLP_K_TRANSPOSE 8              // extends the MLP_K_* kernel ids of mlprobs_b200.h
#define MLP_K_RELAX_ID 7
#define MLP_K_LOCAL_CAND 10            // candidate lists of the local model's Z chain (loc_c.cu)
#define MLP_K_LOCAL_REPLAY 11          // the chain itself, one thread per pair


// per-pair scalars handed from one sweep kernel to the next
struct PairOut {
    double Zpart;          // partition function Z (Zfm[0][0])
    float tF5, total5;     // 5-state: forward total, (forward+backward)/2
    float tFL, totalL;     // local model
    float mea;             // MEA score
    int zexp;              // binary exponent that scales Zpart (rescaled FP64 partition function, cpnp flavour)
};

// Device view of a sparse-posterior set: every ordered pair (a,b) owns len[a]+2 row pointers at rp_off[a*n+b]
// inside rp_pool and nz_cnt cells {column, float bits} at nz_off[a*n+b] inside `cells` (bump-allocated).
struct CsrSetDev {
    int* rp_pool; long long* nz_off; int* nz_cnt; int2* cells; unsigned long long* cursor; long long cap;
};

struct KArgs {
    const PairTask* tasks; int ntasks; int* counter;
    int task_begin;        // register-band kernels (sweep_c.cuh) process tasks[task_begin .. ntasks): one launch per columns-per-lane value
    PairOut* pout;
    const uint8_t* residues; const long long* seq_off; int n;
    int flavour; unsigned mask; float cutoff;
    int Cmax;
    // tables in global memory (staged into shared memory by each CTA)
    const float* match; const float* ins; const double* sub;
    // dense layers (slot layout)
    double* layerZ; float* layerP; float* layerS5; float* layerSL; float* layerVB;
    float* layerLC; int loc_phase; int loc_debug;
    float* rowaux; long long rowaux_stride;   // loc_c.cu, per task: row sums R, prefix bounds PB and candidate counts of the Z chain (three arrays of rowaux_stride/3 floats)   // loc_c.cu: row-major candidate lists of the local model's Z chain; 0 = forward chain (over layerSL), 1 = backward chain (over layerVB)
    int* rowexp; long long rowexp_stride;   // per task: scale exponent of every row of the forward partition layer (cpnp)
    unsigned char* layerTB8; int* vit_ident; int* vit_len; float vit_init0, vit_init1;
    char* vit_aln; const long long* vit_aln_off;   // optional: reversed B/X/Y alignment string of every pair (by pidx)   // Viterbi: packed traceback bytes (slot layout), per-pair results by pidx
    int* layerTB;   // MEA traceback codes (MLP_CPNP_P1 only); aliases a dense layer whose slot has already been consumed
    // boundary-column hand-off between column blocks (only when some pair has nb > 1)
    float* edge_f; double* edge_d; long long edge_stride;
    // sparse sets: rp_off (fixed layout, shared), `out` is written, `in` is read (relax only)
    const long long* rp_off; CsrSetDev out; CsrSetDev in;
    int4* stage; int stage_cap;
    int* tfill; long long tfill_stride;
    float* dist; int* err;
    // optional dense dumps (debug / tests), row-major
    float* dense; float* dense5; float* denseP; float* denseL;
};

// relax_blk.cu
struct RelaxArgs {
    const PairTask* tasks; int ntasks; int* counter;
    int n; int flavour; float cutoff;
    const long long* rp_off;
    CsrSetDev in, out;
    // QP only
    const float* weights; const float* seldist; float selectivity, selfweight;
    float* wk_scratch; long long wk_stride;   // per warp: n floats (weight of z for this pair, < 0 = z not accepted)
    int* err;
    int wide_span;                            // S_yz rows spanning more q than this are merged instead of getting a dense strip
};
// CTA-per-pair relaxation with TMA-staged slices (relax_blk.cu)
cudaError_t relax_blk_launch(const RelaxArgs& a, int grid, cudaStream_t st);
int relax_blk_max_blocks_per_sm();
long long relax_blk_scratch_words(int n);

size_t posterior_smem_bytes(int kernel, int Cmax, int warps);
cudaError_t posterior_set_scalars(const DevScalars& s, cudaStream_t st);
cudaError_t posterior_launch(int kernel, const KArgs& a, int grid, size_t smem, cudaStream_t st);
int posterior_max_blocks_per_sm(int kernel, size_t smem);
// Register-band kernels, compiled per columns-per-lane value C (part_c.cu, hmm_c.cu, final_c.cu)
bool posterior_c_available(int kernel, const KArgs& a);          // is there a C-specialised kernel for this launch?
size_t posterior_c_smem(int kernel, const KArgs& a);
int posterior_c_max_blocks_per_sm(int kernel, int C, const KArgs& a);
cudaError_t posterior_c_launch(int kernel, int C, const KArgs& a, int grid, cudaStream_t st);
cudaError_t part_c_set_scalars(const DevScalars& s, cudaStream_t st);
cudaError_t hmm_c_set_scalars(const DevScalars& s, cudaStream_t st);
void (*part_c_kernel(int kernel, int C))(KArgs);
void (*hmm_c_kernel(int kernel, int C))(KArgs);
void (*final_c_kernel(int C, int mode))(KArgs);
cudaError_t loc_c_set_scalars(const DevScalars& s, cudaStream_t st);
cudaError_t part_sc_set_scalars(const DevScalars& s, cudaStream_t st);
void (*part_sc_kernel(int kernel, int C))(KArgs);
void (*loc_c_kernel(int kernel, int C))(KArgs);
cudaError_t loc_replay_launch(const KArgs& a, cudaStream_t st);
cudaError_t loc_debug_counters(unsigned long long out[4]);
