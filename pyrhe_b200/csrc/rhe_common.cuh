// Shared definitions for the pyrhe_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/pyrhe_b200.h"

struct rhe_ctx {
  rhe_config cfg;
  int Np;        // individuals incl. padding = 4 * pitch_bytes
  int R1;        // total right-hand-side columns = n_sets * Rs
  int n_groups;  // n_ops * n_sets
  int E_reg;     // n_groups * K
  // caller-owned inputs
  const float* rhs = nullptr;
  const float* rowscale = nullptr;
  const uint32_t* keep2 = nullptr;
  const double* uniforms = nullptr;
  int n_uniforms = 0;
  // context-owned workspaces (sized for max_block_snps)
  double* colsum = nullptr;    // [R1]               sum over individuals of every RHS column
  int32_t* counts = nullptr;   // [m][4]             n0 n1 n2 nmiss (kept individuals)
  uint8_t* fill = nullptr;     // [m]                imputation fill value 0/1/2
  double* mu = nullptr;        // [m]                mean A2 count after imputation
  double* f2 = nullptr;        // [m]                fraction of individuals with count 2 (DOM)
  double* t_raw = nullptr;     // [n_ops][m][R1]     G^T R and [G==2]^T R
  double* t_std = nullptr;     // [n_groups][m][Rs]  standardised X^T R
  float* w1 = nullptr;         // [n_groups][m][B]   pass-B weight of [g == 1]
  float* w2 = nullptr;         // [n_groups][m][B]   pass-B weight of [g == 2]
  double* shiftv = nullptr;    // [n_groups][m][B]   per-SNP mean term of pass B
  double* cs = nullptr;        // [E_reg][B]         per-(bin) sum of shiftv
  int32_t* bin_off = nullptr;  // [K + 1]            device copy of the block's bin offsets
  void* tc = nullptr;          // tensor-core path state (rhe_tc.cu)
  int64_t launches = 0;
  bool timing = false;
  std::vector<cudaEvent_t> ev;   // 5 events per timed rhe_block_accumulate call
};

// Annotation-derived metadata of one jackknife block (rhe_block_plan_create): everything rhe_block_accumulate needs
// besides the genotypes is resident before the block is first seen, so the hot call neither allocates nor synchronises.
struct rhe_block_plan {
  int m = 0;                          // SNPs in the block
  const int32_t* bin_rows = nullptr;  // caller-owned device list of block-local SNP rows, bins concatenated
  std::vector<int32_t> off_host;      // [K + 1]
  int32_t* off_dev = nullptr;         // [K + 1]
  void* tc = nullptr;                 // tensor-core path: bin-sorted positions per bin group (rhe_tc.cu)
};

void rhe_set_error(const char* fmt, ...);

#define RHE_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      rhe_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,   \
                    __LINE__);                                                          \
      return RHE_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define RHE_LAUNCH_CHECK(ctx)                                                           \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      rhe_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),         \
                    __FILE__, __LINE__);                                                \
      return RHE_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

// 2-bit PLINK code -> A2 count with the per-SNP fill for the missing code (01).
__device__ __forceinline__ int rhe_code_value(uint32_t code, int fill) {
  // 00 -> 0, 01 -> fill, 10 -> 1, 11 -> 2
  return code == 1u ? fill : (int)(code >> 1) + (int)(code == 3u);
}

// Imputation fill of one SNP from its allele counts {n0, n1, n2, n_miss}: "mean" fills 0 (base.py:287, Q4); "binary"
// replays numpy's float32 arithmetic of base.py:265-285 bit for bit (round-to-nearest intrinsics, no FMA contraction)
// with the uniform drawn for this block-local SNP.  Host specification: pyrhe_b200/hostmath.py:binary_fill_values.
__device__ __forceinline__ int rhe_fill_from_counts(int n1, int n2, int nm, int n_kept, int binary, double uniform) {
  if (!binary) return 0;
  float mean32 = (float)((double)(n1 + 2 * n2) / (double)(n_kept - nm));
  float p = __fmul_rn(mean32, 0.5f);
  float om = __fsub_rn(1.0f, p);
  float d0 = __fmul_rn(om, om);
  float d1 = __fmul_rn(__fmul_rn(2.0f, p), om);
  float u = (float)uniform;
  return (u < d0) ? 0 : ((u < __fadd_rn(d0, d1)) ? 1 : 2);
}

static inline int rhe_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// tensor-core path entry points (rhe_tc.cu)
int rhe_tc_create(rhe_ctx* ctx);
void rhe_tc_destroy(rhe_ctx* ctx);
int rhe_tc_set_rhs(rhe_ctx* ctx, cudaStream_t st);
int rhe_tc_pass_a(rhe_ctx* ctx, const uint8_t* bed, int m, int tiled, cudaStream_t st);
int64_t rhe_tc_tiled_bytes(const rhe_ctx* ctx, int m);                          // bytes of a block's rows once re-tiled (128-row tiles)
int rhe_tc_retile(rhe_ctx* ctx, uint8_t* bed, int m, const int32_t* counts, uint8_t* scratch, cudaStream_t st);
unsigned int* rhe_tc_wmax(rhe_ctx* ctx);   // per-column max |pass-B weight| (float bits), or NULL
int rhe_tc_pass_b(rhe_ctx* ctx, const uint8_t* bed, const uint8_t* gt, const rhe_block_plan* plan, float* P_out,
                  float* S_accum, cudaStream_t st);
int64_t rhe_tc_gt_bytes(const rhe_ctx* ctx, const rhe_block_plan* plan);        // 0: no individual-major fast path
int rhe_tc_transpose(rhe_ctx* ctx, const uint8_t* bed, const rhe_block_plan* plan, const int32_t* counts, uint8_t* gt,
                     cudaStream_t st);
int rhe_tc_plan_create(rhe_ctx* ctx, rhe_block_plan* plan, cudaStream_t st);   // may allocate and synchronise
void rhe_tc_plan_destroy(rhe_block_plan* plan);
int rhe_tc_check(const rhe_config* cfg, int quiet);                            // RHE_OK when the shapes fit the tcgen05 kernels

// Work-skipping ablation switches (PYRHE_TC_DEBUG_*) exist only in the profiling build (-DRHE_TC_DEBUG,
// libpyrhe_b200_prof.so); in the shipped library the tests compile to constants and the variables are never read.
#ifdef RHE_TC_DEBUG
#define RHE_DBG(mask) (dbg & (mask))
#define RHE_DBG_ENV(name, dflt) (getenv(name) ? atoi(getenv(name)) : (dflt))
#else
#define RHE_DBG(mask) 0
#define RHE_DBG_ENV(name, dflt) (dflt)
#endif
