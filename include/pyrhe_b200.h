/* pyrhe_b200 -- C ABI of the B200 (sm_100a) replacement for PyRHE's per-jackknife-block
 * trace-estimation hot path.
 *
 * Reference interfaces replaced (all under /root/reference/pyrhe/src/):
 *   util/mat_mul.py:17-48          mat_mul / elem_mul  (the only tensor call site)
 *   base/base.py:338-359           read_geno           (.bed decode + 0<->2 flip)
 *   base/base.py:265-289           impute_geno         (mean / binary)
 *   base/base.py:291-296           standardize_geno
 *   base/base.py:315-336,362-379   partition_bins / _get_jacknife_subsample
 *   base/base.py:403-417           _compute_XXz / _UXXz / _XXUz / _yXXy
 *   models/rhe_dom/rhe_dom.py:15-41 dominance encoding + scaling
 *   models/genie/genie.py:61-82    GxE row scaling
 *   base/base.py:465-500           aggregate (totals, leave-one-out by subtraction)
 *   base/base.py:578-581           the O(J E^2 B N) Gram of the leave-one-out vectors
 *
 * Conventions: plain C, no C++ types, no exceptions across the boundary.  Every function
 * returns 0 or a negative RHE_ERR_* code; the message is available from rhe_last_error()
 * (thread local).  All device pointers are caller owned (the Python host passes
 * torch tensors' data_ptr()); the context owns only its workspaces.  One context per GPU /
 * rank; calls on one context are serialised by the caller; work is ordered on the
 * `stream` argument (a cudaStream_t passed as void*).  Only the set-up and tear-down calls may allocate or
 * synchronise the host: rhe_ctx_create / rhe_ctx_destroy, rhe_block_plan_create / rhe_block_plan_destroy and
 * rhe_timing_collect.  The per-block calls (rhe_upload_rows, rhe_block_stats, rhe_block_accumulate, rhe_loo_gram*)
 * only enqueue work, so the host runs ahead of the GPU and the first pass over a block costs the same as any later one.
 *
 * Data layout:
 *   packed genotypes  uint8 [n_snps][pitch_bytes]   PLINK-1 SNP-major rows, 4 genotypes/byte,
 *                     LSB first, codes 00->0, 10->1, 11->2 (A2 count), 01->missing; row pitch
 *                     padded to a multiple of 128 bytes, padding zero.  Np = 4 * pitch_bytes.
 *   right-hand sides  float [n_sets * n_cols_set][Np]   column c of set t at row t*Rs + c, file
 *                     order of individuals, rows of dropped / padded individuals zero.
 *   rowscale          float [n_sets][Np]    keep mask (x env for the GxE set)
 *   keep2             uint32 [Np / 16]      2 bits per individual, 0b11 = kept
 *   P / S             float [E_reg][n_vec][Np]   X (X^T Z) per estimate e = group * K + bin
 *   gram              double [E_reg][Rs][Rs]    sum over the (block, bin) SNPs of t t^T,
 *                     t = X_s^T [Z | W | y]
 *   individual-major copy (optional, per resident block; rhe_block_transpose)
 *                     uint8 [Np / 128][n_pos / 512][128][128]: one contiguous 16 KB box per (128 individuals, 512
 *                     bin-sorted positions), 2 bits per genotype holding the imputed A2 COUNT (0, 1, 2; no missing code)
 */
#ifndef PYRHE_B200_H
#define PYRHE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RHE_ABI_VERSION 4

#define RHE_OK 0
#define RHE_ERR_INVALID (-1)   /* bad argument */
#define RHE_ERR_CUDA (-2)      /* CUDA runtime error (message holds cudaGetErrorString) */
#define RHE_ERR_STATE (-3)     /* call order violated (e.g. accumulate before set_rhs) */
#define RHE_ERR_UNSUPPORTED (-4)

#define RHE_ROWS_PLINK 0       /* SNP-major rows exactly as in the .bed file                */
#define RHE_ROWS_TILED 1       /* re-tiled and re-encoded by rhe_block_retile (pass A only) */

#define RHE_PATH_SIMT 0        /* fp32/fp64 CUDA-core kernels (validation path)            */
#define RHE_PATH_TCGEN05 1     /* int8 tcgen05 tensor-core kernels with TMEM accumulators   */

typedef struct rhe_ctx rhe_ctx;
typedef struct rhe_block_plan rhe_block_plan;   /* annotation metadata of one jackknife block */

typedef struct rhe_config {
  int32_t device;          /* CUDA device ordinal */
  int32_t n_indv;          /* N0: individuals in the .bed / .fam (file order) */
  int32_t n_kept;          /* N : individuals after filtering (base.py:156) */
  int32_t pitch_bytes;     /* device row pitch, multiple of 128, >= ceil(N0 / 4) */
  int32_t n_cols_set;      /* Rs = B + C + Ty */
  int32_t n_sets;          /* 1, or 2 when a GxE (env-scaled) set is present */
  int32_t n_ops;           /* 1, or 2 for RHE-DOM (additive + dominance operand) */
  int32_t n_vec;           /* B: leading columns of every set that go through pass B */
  int32_t n_bins;          /* K */
  int32_t max_block_snps;  /* largest jackknife block */
  int32_t impute_binary;   /* 1 = "binary" (base.py:283-285), 0 = "mean" (base.py:287) */
  int32_t kernel_path;     /* RHE_PATH_* */
} rhe_config;

int rhe_version(void);
const char* rhe_last_error(void);

/* mat_mul.py:4-15 / base.py:195-206: device selection and workspace allocation. */
int rhe_ctx_create(rhe_ctx** out, const rhe_config* cfg);
int rhe_ctx_destroy(rhe_ctx* ctx);

/* 1 when the shapes of `cfg` fit the tcgen05 kernels, else 0 with the reason in rhe_last_error() (such a configuration
 * needs kernel_path = RHE_PATH_SIMT).  Shapes beyond one launch's tensor-memory / shared-memory layout run in chunks of
 * vector columns (pass B) or right-hand-side columns (pass A), so every valid configuration is supported today; the
 * call stays the single place where the limits live.  No device is touched. */
int rhe_tc_supported(const rhe_config* cfg);

/* base.py:176-178,396-401 (Z, covariates, regressed phenotype as right-hand sides). */
int rhe_set_rhs(rhe_ctx* ctx, const float* rhs_dev, const float* rowscale_dev,
                const uint32_t* keep2_dev, void* stream);

/* base.py:281-285,510: the uniforms np.random.random() yields after np.random.seed(seed);
 * block-local SNP s uses uniforms[s].  double [count] on the device. */
int rhe_set_uniforms(rhe_ctx* ctx, const double* uniforms_dev, int32_t count);

/* base.py:100,341: host .bed rows -> padded device rows (one pitched async copy). */
int rhe_upload_rows(const void* host_src, int64_t row_bytes, int64_t n_rows,
                    void* dev_dst, int64_t pitch_bytes, void* stream);

/* base.py:277-289 statistics: per SNP {n0, n1, n2, n_missing} over kept individuals.
 * counts_dev int32 [n_snps][4].  The counts depend only on the genotypes and the keep mask, so the ingest path runs
 * this once when a block becomes resident and hands the result to every later rhe_block_accumulate (any trait, any
 * set of random vectors); the imputation fill itself (which needs the per-run uniforms) is re-derived per call. */
int rhe_block_stats(rhe_ctx* ctx, const uint8_t* bed_dev, int32_t n_snps,
                    int32_t* counts_dev, void* stream);

/* base.py:338-359 (+277-289 when apply_impute != 0): decoded A2 counts, one byte per
 * genotype, int8 [n_snps][Np]; missing = 3 when apply_impute == 0.  Test hook for the
 * bit-exact decode check. */
int rhe_decode_block(rhe_ctx* ctx, const uint8_t* bed_dev, int32_t n_snps,
                     int32_t apply_impute, int8_t* out_dev, void* stream);

/* base.py:315-336 (`partition_bins`): the bin row lists of ONE jackknife block, prepared once.
 *   bin_rows_dev     int32 [bin_offsets[K]]  block-local SNP rows of every bin, concatenated (caller owned, must
 *                                            outlive the plan); a SNP may sit in several bins
 *   bin_offsets_host int32 [K + 1]           (host memory, copied)
 * May allocate device memory and synchronise; the plan is then immutable and reusable for any number of calls. */
int rhe_block_plan_create(rhe_ctx* ctx, int32_t n_snps, const int32_t* bin_rows_dev,
                          const int32_t* bin_offsets_host, void* stream, rhe_block_plan** out);
int rhe_block_plan_destroy(rhe_ctx* ctx, rhe_block_plan* plan);

/* Optional second layout of a resident block (tensor-core path): the INDIVIDUAL-MAJOR copy
 *   gt   uint8 [Np / 128][n_pos / 512][128][128]   one contiguous 16 KB box per (128 individuals, 512 bin-sorted
 *                                positions of the plan; every bin padded to 128 positions, the list to 512): row =
 *                                individual, 2 bits per position, imputation applied, 16 positions per 32-bit word
 * lets pass B (X (X^T Z), base.py:403-405) feed the tensor cores from tensor memory like pass A instead of staging every
 * genotype through shared memory as a byte.  It doubles the genotype footprint, so the caller decides per
 * block.  rhe_block_fast_bytes: size of the copy, 0 when the configuration / plan has no such path.
 * rhe_block_transpose: builds it from the packed rows and their counts (ingest; depends on the imputation uniforms). */
int64_t rhe_block_fast_bytes(const rhe_ctx* ctx, const rhe_block_plan* plan);
int rhe_block_transpose(rhe_ctx* ctx, const uint8_t* bed_dev, const rhe_block_plan* plan,
                        const int32_t* counts_dev, uint8_t* gt_dev, void* stream);

/* Optional in-place re-layout of a resident block's SNP-major rows for pass A (X^T [Z | W | y], base.py:403-417), once the
 * counts are taken and the individual-major copy is written (both read the PLINK rows):
 *   [n_snps / 128][pitch / 128][128 rows][128 B]   every TMA box of pass A is one contiguous 16 KB piece (streams from HBM
 *                        like a copy instead of 128 row-strided segments) and the two bits of a genotype hold the imputed
 *                        A2 count (mask-and-shift decode, no per-SNP table).
 * The rows buffer must hold rhe_block_tiled_bytes (the SNP count rounded up to 128 rows; 0 = not available: 2^21 or more
 * individuals, or no tensor-core path); scratch_dev as many bytes.
 * Afterwards the rows are readable ONLY by rhe_block_accumulate(..., rows_layout = RHE_ROWS_TILED) with counts and copy. */
int64_t rhe_block_tiled_bytes(const rhe_ctx* ctx, const rhe_block_plan* plan);
int rhe_block_retile(rhe_ctx* ctx, uint8_t* bed_dev, const rhe_block_plan* plan, const int32_t* counts_dev,
                     uint8_t* scratch_dev, void* stream);

/* rhe.py:13-22 / rhe_dom.py:43-68 / genie.py:46-82 for ONE jackknife block:
 *   plan             from rhe_block_plan_create (its n_snps rows start at bed_dev)
 *   counts_dev       int32 [n_snps][4] from rhe_block_stats on the same rows, or NULL (counted here: one more
 *                    read of the block)
 *   gt_dev           the block's individual-major copy (rhe_block_transpose) or NULL (pass B gathers SNP rows)
 *   rows_layout      RHE_ROWS_PLINK, or RHE_ROWS_TILED after rhe_block_retile (then counts_dev and gt_dev are required)
 *   P_out_dev        float [E_reg][B][Np] or NULL   this block's X (X^T Z)
 *   S_accum_dev      float [E_reg][B][Np] or NULL   running totals (+=)
 *   gram_out_dev     double [E_reg][Rs][Rs]         overwritten
 * Enqueues kernels on `stream` only: no allocation, no host synchronisation. */
int rhe_block_accumulate(rhe_ctx* ctx, const uint8_t* bed_dev, const rhe_block_plan* plan,
                         const int32_t* counts_dev, const uint8_t* gt_dev, int32_t rows_layout, float* P_out_dev,
                         float* S_accum_dev, double* gram_out_dev, void* stream);

/* base.py:483-486 (totals over the blocks): S[i] = sum_j P[j * p_stride + i], j in block order, `len` floats (a multiple
 * of 4).  With stored partials a caller may pass S_accum_dev = NULL to rhe_block_accumulate and sum once here: one
 * streaming read of the partials instead of a read-modify-write of S in every block's pass B, for a total that is
 * bit-reproducible from run to run (the RED accumulation is not; it is slightly faster). */
int rhe_sum_partials(rhe_ctx* ctx, const float* P_dev, int64_t p_stride, int32_t n_blocks, int64_t len, float* S_dev,
                     void* stream);

/* base.py:578-581 after aggregate (base.py:483-486): out[a][c] = sum (S_a - P_a)(S_c - P_c)
 * over `len` floats per estimate; P_dev may be NULL (totals).  out_dev double [n_est][n_est]. */
int rhe_loo_gram(rhe_ctx* ctx, const float* S_dev, const float* P_dev, int32_t n_est,
                 int64_t len, double* out_dev, void* stream);

/* The same for n_blocks partials laid out at P_dev + b * p_stride floats, results at out_dev + b * out_stride
 * doubles: up to four blocks share one read of S per launch (the pass is memory bound). */
int rhe_loo_gram_multi(rhe_ctx* ctx, const float* S_dev, const float* P_dev, int64_t p_stride, int32_t n_blocks,
                       int32_t n_est, int64_t len, double* out_dev, int64_t out_stride, void* stream);

/* Synthetic PLINK rows generated on the device: SNP s has A2 frequency
 * p_s ~ U(0.05, 0.5), genotypes ~ Binomial(2, p_s) i.i.d., code 01 (missing) with probability
 * missing_rate; counter-based hash of (seed, first_snp + row, byte) so any row range can be
 * regenerated.  Bench / test utility: 125 GB of genotypes never has to cross PCIe. */
int rhe_synth_genotypes(uint8_t* bed_dev, int64_t n_rows, int64_t pitch_bytes, int32_t n_indv,
                        int64_t first_snp, uint64_t seed, float missing_rate, void* stream);

/* Per-phase device timing of rhe_block_accumulate (CUDA events on the launch stream).
 * phases_ms double[4] = {stats + imputation parameters, pass A, standardise + Gram, pass B},
 * summed over the calls since the last collect; the call synchronises the events. */
int rhe_timing_enable(rhe_ctx* ctx, int32_t enable);
int rhe_timing_collect(rhe_ctx* ctx, double* phases_ms, int32_t* n_calls);

/* Number of kernels this library has launched on behalf of `ctx` (bench.py gpu_launches). */
int64_t rhe_launch_count(const rhe_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PYRHE_B200_H */
