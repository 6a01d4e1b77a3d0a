// pyrhe_b200 C ABI + CUDA-core (SIMT) kernels.  See include/pyrhe_b200.h for the contract and
// DESIGN.md §3 for the algorithm.  The tensor-core (tcgen05) kernels live in rhe_tc.cu and
// replace pass A / pass B when cfg.kernel_path == RHE_PATH_TCGEN05.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include "rhe_common.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[512] = "";

void rhe_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rhe_last_error(void) { return g_err; }
extern "C" int rhe_version(void) { return RHE_ABI_VERSION; }
extern "C" int64_t rhe_launch_count(const rhe_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------
// kernels

// Column sums of the right-hand sides (used by the rank-1 mean correction of X^T R).
__global__ void k_colsum(const float* __restrict__ rhs, int Np, double* __restrict__ colsum) {
  const float* col = rhs + (size_t)blockIdx.x * Np;
  double acc = 0.0;
  for (int i = threadIdx.x; i < Np; i += blockDim.x) acc += (double)col[i];
  __shared__ double red[32];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) colsum[blockIdx.x] = acc;
  }
}

__device__ __forceinline__ uint4 rhe_ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void rhe_popc_word(uint32_t x, uint32_t keep, int& n1, int& n2, int& nm) {
  const uint32_t k = keep & 0x55555555u, lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
  n2 += __popc(hi & lo & k);
  n1 += __popc(hi & ~lo & k);
  nm += __popc(~hi & lo & k);
}

// Cheaper variant for the fused front end: only S = n1 + 2 n2 and n_miss are needed (plus n2 for RHE-DOM).
// popc(x & keep) = (n1 + n2) + (n2 + n_miss), so S = popc(x & keep) - n_miss: two POPC per word instead of three.
template <bool NEED_N2>
__device__ __forceinline__ void rhe_popc_word2(uint32_t x, uint32_t keep, int& bits, int& nm, int& n2) {
  const uint32_t t = x & keep, sh = t >> 1;
  bits += __popc(t);
  nm += __popc(t & ~sh & 0x55555555u);
  if (NEED_N2) n2 += __popc(t & sh & 0x55555555u);
}

// Masked popcount statistics: one warp per SNP row (base.py:277-289 needs the observed mean).
__global__ void __launch_bounds__(256)
k_stats(const uint8_t* __restrict__ bed, int pitch, int m, const uint32_t* __restrict__ keep2,
        int n_kept, int32_t* __restrict__ counts) {
  int s = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (s >= m) return;
  int lane = threadIdx.x & 31;
  const uint32_t* row = reinterpret_cast<const uint32_t*>(bed + (size_t)s * pitch);
  int n1 = 0, n2 = 0, nm = 0;
  {
    const uint4* row4 = reinterpret_cast<const uint4*>(row);
    const uint4* keep4 = reinterpret_cast<const uint4*>(keep2);
    const int n4 = pitch / 16;                      // pitch is a multiple of 128 bytes
#pragma unroll 4
    for (int w = lane; w < n4; w += 32) {
      const uint4 x = rhe_ldg_stream(row4 + w), k = __ldg(keep4 + w);
      rhe_popc_word(x.x, k.x, n1, n2, nm);
      rhe_popc_word(x.y, k.y, n1, n2, nm);
      rhe_popc_word(x.z, k.z, n1, n2, nm);
      rhe_popc_word(x.w, k.w, n1, n2, nm);
    }
  }
  for (int o = 16; o; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    nm += __shfl_xor_sync(0xffffffffu, nm, o);
  }
  if (lane == 0) {
    int4 c = make_int4(n_kept - n1 - n2 - nm, n1, n2, nm);
    reinterpret_cast<int4*>(counts)[s] = c;
  }
}

// Per-SNP imputation fill + moments.  The fill decision replays numpy's float32 arithmetic of
// base.py:265-285 bit for bit (round-to-nearest intrinsics, no FMA contraction); its host
// specification is pyrhe_b200/hostmath.py:binary_fill_values.
__global__ void k_snp_params(const int32_t* __restrict__ counts, int m, int n_kept, int binary,
                             const double* __restrict__ uniforms, uint8_t* __restrict__ fill,
                             double* __restrict__ mu, double* __restrict__ f2) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= m) return;
  int4 c = reinterpret_cast<const int4*>(counts)[s];
  int n1 = c.y, n2 = c.z, nm = c.w, f = 0;
  if (binary) {
    float mean32 = (float)((double)(n1 + 2 * n2) / (double)(n_kept - nm));
    float p = __fmul_rn(mean32, 0.5f);
    float om = __fsub_rn(1.0f, p);
    float d0 = __fmul_rn(om, om);
    float d1 = __fmul_rn(__fmul_rn(2.0f, p), om);
    float u = (float)uniforms[s];
    f = (u < d0) ? 0 : ((u < __fadd_rn(d0, d1)) ? 1 : 2);
  }
  if (f == 1) n1 += nm;
  if (f == 2) n2 += nm;
  fill[s] = (uint8_t)f;
  mu[s] = (double)(n1 + 2 * n2) / (double)n_kept;
  f2[s] = (double)n2 / (double)n_kept;
}

// Fused front end of rhe_block_accumulate: masked popcounts, the imputation fill and moments of the SNP, and the
// zeroing of everything this block accumulates into (t_raw rows, cs, gram, wmax).  ST_SPLIT warps share one SNP row
// (a quarter of the row each, combined through shared memory): with one warp per row the 10^4 rows of a block are
// only ~1.4 waves of resident warps and the half-empty last wave costs a quarter of the pass.
#define ST_SPLIT 4
#define ST_ROWS (8 / ST_SPLIT)
__global__ void __launch_bounds__(256)
k_stats_params(const uint8_t* __restrict__ bed, int pitch, int m, const uint32_t* __restrict__ keep2, int n_kept,
               int binary, const double* __restrict__ uniforms, int32_t* __restrict__ counts,
               uint8_t* __restrict__ fill, double* __restrict__ mu, double* __restrict__ f2,
               double* __restrict__ t_raw, int t_cols, int n_ops, double* __restrict__ cs, int n_cs,
               double* __restrict__ gram, int n_gram, unsigned int* __restrict__ wmax, int n_wmax) {
  __shared__ int part[8][3];
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < n_cs; i += 256) cs[i] = 0.0;
    for (int i = threadIdx.x; i < n_wmax; i += 256) wmax[i] = 0u;
  }
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n_gram; i += gridDim.x * 256) gram[i] = 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * ST_ROWS + warp / ST_SPLIT, piece = warp % ST_SPLIT;
  const bool live = s < m;
  int bits = 0, n2 = 0, nm = 0;
  if (live) {
    if (piece == 0)
      for (int op = 0; op < n_ops; ++op)
        for (int c = lane; c < t_cols; c += 32) t_raw[((size_t)op * m + s) * t_cols + c] = 0.0;
    const uint4* row4 = reinterpret_cast<const uint4*>(bed + (size_t)s * pitch);
    const uint4* keep4 = reinterpret_cast<const uint4*>(keep2);
    const int n4 = pitch / 16;                        // pitch is a multiple of 128 bytes
    const int per = (n4 + ST_SPLIT - 1) / ST_SPLIT;
    const int w0 = piece * per, w1 = min(n4, w0 + per);
    if (n_ops == 2) {
#pragma unroll 4
      for (int w = w0 + lane; w < w1; w += 32) {
        const uint4 x = rhe_ldg_stream(row4 + w), k = __ldg(keep4 + w);
        rhe_popc_word2<true>(x.x, k.x, bits, nm, n2); rhe_popc_word2<true>(x.y, k.y, bits, nm, n2);
        rhe_popc_word2<true>(x.z, k.z, bits, nm, n2); rhe_popc_word2<true>(x.w, k.w, bits, nm, n2);
      }
    } else {
#pragma unroll 4
      for (int w = w0 + lane; w < w1; w += 32) {
        const uint4 x = rhe_ldg_stream(row4 + w), k = __ldg(keep4 + w);
        rhe_popc_word2<false>(x.x, k.x, bits, nm, n2); rhe_popc_word2<false>(x.y, k.y, bits, nm, n2);
        rhe_popc_word2<false>(x.z, k.z, bits, nm, n2); rhe_popc_word2<false>(x.w, k.w, bits, nm, n2);
      }
    }
    for (int o = 16; o; o >>= 1) {
      bits += __shfl_xor_sync(0xffffffffu, bits, o);
      n2 += __shfl_xor_sync(0xffffffffu, n2, o);
      nm += __shfl_xor_sync(0xffffffffu, nm, o);
    }
  }
  if (lane == 0) { part[warp][0] = bits; part[warp][1] = n2; part[warp][2] = nm; }
  __syncthreads();
  if (live && piece == 0 && lane == 0) {
    for (int q = 1; q < ST_SPLIT; ++q) { bits += part[warp + q][0]; n2 += part[warp + q][1]; nm += part[warp + q][2]; }
    const int ssum = bits - nm;                        // n1 + 2 n2 over kept individuals
    // n1 / n0 are filled in only when n2 was counted (RHE-DOM); the exact four-way counts come from k_stats
    reinterpret_cast<int4*>(counts)[s] = make_int4(n_kept - (ssum - n2) - nm, ssum - 2 * n2, n2, nm);
    int f = 0;
    int n1 = ssum - 2 * n2;
    if (binary) {   // same float32 replay as k_snp_params
      float mean32 = (float)((double)ssum / (double)(n_kept - nm));
      float p = __fmul_rn(mean32, 0.5f);
      float om = __fsub_rn(1.0f, p);
      float d0 = __fmul_rn(om, om);
      float d1 = __fmul_rn(__fmul_rn(2.0f, p), om);
      float u = (float)uniforms[s];
      f = (u < d0) ? 0 : ((u < __fadd_rn(d0, d1)) ? 1 : 2);
    }
    if (f == 1) n1 += nm;
    if (f == 2) n2 += nm;
    fill[s] = (uint8_t)f;
    mu[s] = (double)(n1 + 2 * n2) / (double)n_kept;   // = (S + fill * n_miss) / N, valid without n2 as well
    f2[s] = (double)n2 / (double)n_kept;               // used by the dominance operand only (n2 counted then)
  }
}

// Front end of rhe_block_accumulate when the block's allele counts are already resident (rhe_block_stats at ingest):
// imputation fill + moments from the counts (same float32 replay as k_snp_params) and the zeroing of everything the
// block accumulates into.  No genotype byte is read.
__global__ void __launch_bounds__(256)
k_params_from_counts(const int32_t* __restrict__ counts, int m, int n_kept, int binary, const double* __restrict__ uniforms,
                     uint8_t* __restrict__ fill, double* __restrict__ mu, double* __restrict__ f2,
                     double* __restrict__ t_raw, int n_traw, double* __restrict__ cs, int n_cs,
                     double* __restrict__ gram, int n_gram, unsigned int* __restrict__ wmax, int n_wmax) {
  const int tid = blockIdx.x * 256 + threadIdx.x, nth = gridDim.x * 256;
  for (int i = tid; i < n_traw; i += nth) t_raw[i] = 0.0;
  for (int i = tid; i < n_gram; i += nth) gram[i] = 0.0;
  for (int i = tid; i < n_cs; i += nth) cs[i] = 0.0;
  for (int i = tid; i < n_wmax; i += nth) wmax[i] = 0u;
  for (int s = tid; s < m; s += nth) {
    const int4 c = reinterpret_cast<const int4*>(counts)[s];
    int n1 = c.y, n2 = c.z;
    const int nm = c.w;
    const int f = rhe_fill_from_counts(n1, n2, nm, n_kept, binary, binary ? uniforms[s] : 0.0);
    if (f == 1) n1 += nm;
    if (f == 2) n2 += nm;
    fill[s] = (uint8_t)f;
    mu[s] = (double)(n1 + 2 * n2) / (double)n_kept;
    f2[s] = (double)n2 / (double)n_kept;
  }
}

// Test hook: decoded (optionally imputed) A2 counts, one byte per genotype.
__global__ void k_decode(const uint8_t* __restrict__ bed, int pitch, int m,
                         const uint8_t* __restrict__ fill, int apply_impute, int8_t* __restrict__ out) {
  int Np = pitch * 4;
  size_t total = (size_t)m * pitch;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int s = (int)(idx / pitch), byte = (int)(idx % pitch);
    uint32_t b = bed[idx];
    int f = apply_impute ? fill[s] : 3;
    char4 v;
    v.x = (char)rhe_code_value(b & 3u, f);
    v.y = (char)rhe_code_value((b >> 2) & 3u, f);
    v.z = (char)rhe_code_value((b >> 4) & 3u, f);
    v.w = (char)rhe_code_value((b >> 6) & 3u, f);
    reinterpret_cast<char4*>(out + (size_t)s * Np)[byte] = v;
  }
}

// Pass A (SIMT): t_raw[s][c] += sum_i val(g_is) * R[c][i]   for a 32-SNP x chunk tile and 16 columns.
// MODE 0: val = imputed A2 count; MODE 1: val = [count == 2] (dominance operand, rhe_dom.py:36-39).
template <int MODE>
__global__ void __launch_bounds__(256)
k_pass_a(const uint8_t* __restrict__ bed, int pitch, int m, const float* __restrict__ rhs, int Np,
         int R1, const uint8_t* __restrict__ fill, double* __restrict__ t_raw, int chunk) {
  __shared__ float Rt[16][512];
  __shared__ uint32_t gw[32][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int snp0 = blockIdx.x * 32, col0 = blockIdx.z * 16;
  const int i_begin = blockIdx.y * chunk, i_end = min(Np, i_begin + chunk);
  float acc[4][16];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[q][c] = 0.f;
  int fl[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int s = snp0 + warp * 4 + q;
    fl[q] = s < m ? fill[s] : 0;
  }
  for (int base = i_begin; base < i_end; base += 512) {
    for (int idx = threadIdx.x; idx < 16 * 512; idx += 256) {
      int c = idx >> 9, i = idx & 511;
      Rt[c][i] = (col0 + c < R1) ? __ldg(rhs + (size_t)(col0 + c) * Np + base + i) : 0.f;
    }
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 256) {
      int r = idx >> 5, w = idx & 31;
      int s = snp0 + r;
      gw[r][w] = s < m ? __ldg(reinterpret_cast<const uint32_t*>(bed + (size_t)s * pitch) + (base >> 4) + w) : 0u;
    }
    __syncthreads();
#pragma unroll 2
    for (int t = 0; t < 16; ++t) {
      const int ind = 32 * t + lane;
      const int word = ind >> 4, sh = (ind & 15) * 2;
      float r[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) r[c] = Rt[c][ind];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t code = (gw[warp * 4 + q][word] >> sh) & 3u;
        int g = rhe_code_value(code, fl[q]);
        float v = MODE == 0 ? (float)g : (g == 2 ? 1.f : 0.f);
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[q][c] = fmaf(v, r[c], acc[q][c]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float v = acc[q][c];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      int s = snp0 + warp * 4 + q;
      if (lane == 0 && s < m && col0 + c < R1) atomicAdd(t_raw + (size_t)s * R1 + col0 + c, (double)v);
    }
}

// Standardisation as a rank-1 fix-up of the raw products:
//   additive  t = r (G^T R - mu 1^T R),                      r  = 1/sqrt(mu (1 - mu/2))   base.py:291-296
//   dominance t = r'(mu G^T R - 2 [G==2]^T R - eta 1^T R),   r' = 1/(mu (1 - mu/2)),
//             eta = mu^2 - 2 f2 (mean of the encoded column)                                rhe_dom.py:15-41
// plus the pass-B weights of [g==1], [g==2] and the per-SNP mean term.
__global__ void k_standardize(int m, int Rs, int R1, int B, int n_ops, int n_sets,
                              const double* __restrict__ t_raw, const double* __restrict__ colsum,
                              const double* __restrict__ mu, const double* __restrict__ f2,
                              double* __restrict__ t_std, float* __restrict__ w1, float* __restrict__ w2,
                              double* __restrict__ shiftv, unsigned int* __restrict__ wmax) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int n_groups = n_ops * n_sets;
  if (idx >= n_groups * m * Rs) return;
  int c = idx % Rs, s = (idx / Rs) % m, grp = idx / (Rs * m);
  int op = n_ops == 2 ? grp : 0, st = n_ops == 2 ? 0 : grp;
  int col = st * Rs + c;
  double mean = mu[s], var = mean * (1.0 - 0.5 * mean);
  double ta = t_raw[(size_t)s * R1 + col], cs = colsum[col];
  double t, u, a1, a2, sh;
  if (op == 0) {
    double r = 1.0 / sqrt(var);
    t = r * (ta - mean * cs);
    u = r * t;
    a1 = u;
    a2 = 2.0 * u;
    sh = mean * u;
  } else {
    double r = 1.0 / var, eta = mean * mean - 2.0 * f2[s];
    double t2 = t_raw[(size_t)m * R1 + (size_t)s * R1 + col];
    t = r * (mean * ta - 2.0 * t2 - eta * cs);
    u = r * t;
    a1 = mean * u;
    a2 = 2.0 * mean * u - 2.0 * u;
    sh = eta * u;
  }
  t_std[((size_t)grp * m + s) * Rs + c] = t;
  if (c < B) {
    size_t o = ((size_t)grp * m + s) * B + c;
    w1[o] = (float)a1;
    w2[o] = (float)a2;
    shiftv[o] = sh;
    if (wmax)   // tensor path: per-column quantisation range over both operand weights (count, [g == 2] extra)
      atomicMax(wmax + grp * B + c, __float_as_uint(fmaxf(fabsf((float)a1), fabsf((float)a2 - 2.0f * (float)a1))));
  }
}

// Per-(group, bin) Gram of the standardised products and the pass-B mean term.
//   gram[e][c1][c2] += sum_{s in bin} t[s][c1] t[s][c2];   cs[e][b] += sum_{s in bin} shiftv[s][b]
__global__ void __launch_bounds__(256)
k_bin_gram(int m, int Rs, int B, int K, const int32_t* __restrict__ bin_rows,
           const int32_t* __restrict__ bin_off, const double* __restrict__ t_std,
           const double* __restrict__ shiftv, double* __restrict__ gram, double* __restrict__ cs) {
  const int e = blockIdx.x, k = e % K, grp = e / K;
  const int begin = bin_off[k], end = bin_off[k + 1];
  const int per = (end - begin + gridDim.y - 1) / gridDim.y;
  const int lo = begin + blockIdx.y * per, hi = min(end, lo + per);
  const double* T = t_std + (size_t)grp * m * Rs;
  const double* SH = shiftv + (size_t)grp * m * B;
  for (int p = threadIdx.x; p < Rs * Rs + B; p += blockDim.x) {
    double acc = 0.0;
    if (p < Rs * Rs) {
      int c1 = p / Rs, c2 = p % Rs;
      if (c2 < c1) continue;
      // four independent gather chains (row index -> row) in flight: the loop is pure load latency
      double a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int i = lo;
      for (; i + 4 <= hi; i += 4) {
        const double* r0 = T + (size_t)bin_rows[i] * Rs;
        const double* r1 = T + (size_t)bin_rows[i + 1] * Rs;
        const double* r2 = T + (size_t)bin_rows[i + 2] * Rs;
        const double* r3 = T + (size_t)bin_rows[i + 3] * Rs;
        acc += r0[c1] * r0[c2];
        a1 += r1[c1] * r1[c2];
        a2 += r2[c1] * r2[c2];
        a3 += r3[c1] * r3[c2];
      }
      for (; i < hi; ++i) {
        const double* row = T + (size_t)bin_rows[i] * Rs;
        acc += row[c1] * row[c2];
      }
      acc += a1 + a2 + a3;
      if (lo < hi) {
        atomicAdd(gram + ((size_t)e * Rs + c1) * Rs + c2, acc);
        if (c1 != c2) atomicAdd(gram + ((size_t)e * Rs + c2) * Rs + c1, acc);
      }
    } else {
      int b = p - Rs * Rs;
      for (int i = lo; i < hi; ++i) acc += SH[(size_t)bin_rows[i] * B + b];
      if (lo < hi) atomicAdd(cs + (size_t)e * B + b, acc);
    }
  }
}

// Pass B (SIMT): P[e][b][i] = rowscale[i] * ( sum_{s in bin} [g_is==1] w1[s][b] + [g_is==2] w2[s][b] - cs[e][b] )
// One thread per individual, BG columns per thread, 32-SNP tiles staged in shared memory.
template <int BG>
__global__ void __launch_bounds__(512)
k_pass_b(const uint8_t* __restrict__ bed, int pitch, int m, int Np, int B, int K, int n_ops,
         const int32_t* __restrict__ bin_rows, const int32_t* __restrict__ bin_off,
         const uint8_t* __restrict__ fill, const float* __restrict__ w1, const float* __restrict__ w2,
         const double* __restrict__ cs, const float* __restrict__ rowscale,
         float* __restrict__ P_out, float* __restrict__ S_accum) {
  __shared__ uint32_t gw[32][33];
  __shared__ float s1[32][BG], s2[32][BG];
  __shared__ int sfill[32];
  const int e = blockIdx.y, k = e % K, grp = e / K;
  const int st = n_ops == 2 ? 0 : grp;
  const int b0 = blockIdx.z * BG;
  const int ibase = blockIdx.x * 512, i = ibase + threadIdx.x;
  const int begin = bin_off[k], end = bin_off[k + 1];
  const float* W1 = w1 + (size_t)grp * m * B;
  const float* W2 = w2 + (size_t)grp * m * B;
  double acc[BG];
#pragma unroll
  for (int b = 0; b < BG; ++b) acc[b] = 0.0;
  const int word = threadIdx.x >> 4, sh = (threadIdx.x & 15) * 2;
  for (int base = begin; base < end; base += 32) {
    const int nrow = min(32, end - base);
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 512) {
      int r = idx >> 5, w = idx & 31;
      gw[r][w] = r < nrow ? __ldg(reinterpret_cast<const uint32_t*>(bed + (size_t)bin_rows[base + r] * pitch) + (ibase >> 4) + w) : 0u;
    }
    for (int idx = threadIdx.x; idx < 32 * BG; idx += 512) {
      int r = idx / BG, b = idx % BG;
      bool ok = r < nrow && b0 + b < B;
      size_t o = ok ? (size_t)bin_rows[base + r] * B + b0 + b : 0;
      s1[r][b] = ok ? W1[o] : 0.f;
      s2[r][b] = ok ? W2[o] : 0.f;
    }
    if (threadIdx.x < 32) sfill[threadIdx.x] = threadIdx.x < nrow ? fill[bin_rows[base + threadIdx.x]] : 0;
    __syncthreads();
    float part[BG];
#pragma unroll
    for (int b = 0; b < BG; ++b) part[b] = 0.f;
    for (int r = 0; r < nrow; ++r) {
      int g = rhe_code_value((gw[r][word] >> sh) & 3u, sfill[r]);
      float i1 = g == 1 ? 1.f : 0.f, i2 = g == 2 ? 1.f : 0.f;
#pragma unroll
      for (int b = 0; b < BG; ++b) part[b] = fmaf(i1, s1[r][b], fmaf(i2, s2[r][b], part[b]));
    }
#pragma unroll
    for (int b = 0; b < BG; ++b) acc[b] += (double)part[b];
    __syncthreads();
  }
  if (i < Np) {
    const double rs = (double)rowscale[(size_t)st * Np + i];
#pragma unroll
    for (int b = 0; b < BG; ++b) {
      if (b0 + b >= B) break;
      float v = (float)(rs * (acc[b] - cs[(size_t)e * B + b0 + b]));
      size_t o = ((size_t)e * B + b0 + b) * Np + i;
      if (P_out) P_out[o] = v;
      if (S_accum) S_accum[o] += v;
    }
  }
}

// Gram of the leave-one-out vectors: out[a][c] += sum_p (S_a[p] - P_a[p]) (S_c[p] - P_c[p]).
// (base.py:483-486 forms S - P in fp64; base.py:578-581 sums the products in fp64.)
#define GRAM_CHUNK 128
#define GRAM_MAX_E 32
__global__ void __launch_bounds__(256)
k_loo_gram(const float* __restrict__ S, const float* __restrict__ P, int E, int64_t len,
           double* __restrict__ out) {
  __shared__ double d[GRAM_MAX_E][GRAM_CHUNK + 1];
  const int npairs = E * (E + 1) / 2;
  const int nseg = max(1, min(8, 256 / npairs));
  // a thread owns up to 3 (pair, segment) work items: 528 pairs max for E = 32
  double acc[3] = {0.0, 0.0, 0.0};
  int pa[3], pc[3], seg[3];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    int item = threadIdx.x + w * 256;
    int pair = item % npairs;
    seg[w] = item / npairs;
    if (seg[w] >= nseg || (npairs > 256 && item >= npairs)) { pa[w] = -1; pc[w] = 0; continue; }
    if (npairs > 256) seg[w] = 0;
    // unrank pair -> (a <= c)
    int a = 0, rem = pair;
    while (rem >= E - a) { rem -= E - a; ++a; }
    pa[w] = a;
    pc[w] = a + rem;
  }
  const int nseg_eff = npairs > 256 ? 1 : nseg;
  const int64_t nchunks = (len + GRAM_CHUNK - 1) / GRAM_CHUNK;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t p0 = ch * GRAM_CHUNK;
    for (int idx = threadIdx.x; idx < E * GRAM_CHUNK; idx += 256) {
      int a = idx / GRAM_CHUNK, p = idx % GRAM_CHUNK;
      int64_t g = p0 + p;
      double v = 0.0;
      if (g < len) {
        v = (double)S[(size_t)a * len + g];
        if (P) v -= (double)P[(size_t)a * len + g];
      }
      d[a][p] = v;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 3; ++w) {
      if (pa[w] < 0) continue;
      double s = 0.0;
      for (int p = seg[w]; p < GRAM_CHUNK; p += nseg_eff) s += d[pa[w]][p] * d[pc[w]][p];
      acc[w] += s;
    }
    __syncthreads();
  }
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    if (pa[w] < 0) continue;
    atomicAdd(out + (size_t)pa[w] * E + pc[w], acc[w]);
    if (pa[w] != pc[w]) atomicAdd(out + (size_t)pc[w] * E + pa[w], acc[w]);
  }
}

// Synthetic genotypes: one thread per packed byte (4 genotypes).
__device__ __forceinline__ uint64_t rhe_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void k_synth(uint8_t* __restrict__ bed, int64_t n_rows, int64_t pitch, int n_indv, int64_t first_snp,
                        uint64_t seed, float missing_rate) {
  const int64_t total = n_rows * pitch;
  const int row_bytes = (n_indv + 3) / 4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / pitch;
    const int byte = (int)(idx % pitch);
    uint32_t out = 0;
    if (byte < row_bytes) {
      const uint64_t snp = (uint64_t)(first_snp + r);
      const float p = 0.05f + 0.45f * (float)(rhe_mix64(seed ^ (snp * 0xD1342543DE82EF95ull)) >> 40) * (1.0f / 16777216.0f);
      const uint32_t thr = (uint32_t)(p * 65536.0f), mthr = (uint32_t)(missing_rate * 65536.0f);
      const uint64_t h = rhe_mix64(rhe_mix64(seed + snp) ^ ((uint64_t)byte * 0x2545F4914F6CDD1Dull));
      const uint64_t h2 = rhe_mix64(h);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (byte * 4 + t >= n_indv) break;
        const uint32_t a = (uint32_t)(h >> (16 * t)) & 0xFFFFu, b = (uint32_t)(h2 >> (16 * t)) & 0xFFFFu;
        const uint32_t c = (uint32_t)(rhe_mix64(h2 + t) & 0xFFFFu);
        const int g = (a < thr) + (b < thr);
        uint32_t code = g == 0 ? 0u : (g == 1 ? 2u : 3u);
        if (c < mthr) code = 1u;
        out |= code << (2 * t);
      }
    }
    bed[idx] = (uint8_t)out;
  }
}

// Fast path for E <= 8 (RHE with up to 8 bins): one thread per position, the (at most 36) pair products stay in
// fp64 registers over a grid-stride loop, one warp + block reduction at the end.  Memory bound: reads S and P once.
template <int E>
__global__ void __launch_bounds__(256, 2)
k_loo_gram_small(const float* __restrict__ S, const float* __restrict__ P, int64_t len, double* __restrict__ out) {
  constexpr int NP = E * (E + 1) / 2;
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  // two adjacent positions per thread and iteration (8-byte loads; `len` is a multiple of 512)
  const int64_t len2 = len >> 1;
  const float2* S2 = reinterpret_cast<const float2*>(S);
  const float2* P2 = reinterpret_cast<const float2*>(P);
  for (int64_t g = blockIdx.x * (int64_t)256 + threadIdx.x; g < len2; g += (int64_t)gridDim.x * 256) {
    double d0[E], d1[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
      const float2 v = __ldg(S2 + (size_t)a * len2 + g);
      d0[a] = (double)v.x;
      d1[a] = (double)v.y;
      if (P) {
        const float2 w = __ldg(P2 + (size_t)a * len2 + g);
        d0[a] -= (double)w.x;
        d1[a] -= (double)w.y;
      }
    }
    int p = 0;
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
      for (int c = a; c < E; ++c) { acc[p] = fma(d0[a], d0[c], fma(d1[a], d1[c], acc[p])); ++p; }
  }
  __shared__ double red[8][NP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    double v = acc[p];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][p] = v;
  }
  __syncthreads();
  if (threadIdx.x < NP) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    int a = 0, rem = threadIdx.x;
    while (rem >= E - a) { rem -= E - a; ++a; }
    const int c = a + rem;
    atomicAdd(out + a * E + c, v);
    if (a != c) atomicAdd(out + c * E + a, v);
  }
}

template <int E>
static void launch_loo_small(const float* S, const float* P, int64_t len, double* out, cudaStream_t st) {
  int64_t blocks = (len / 2 + 255) / 256;
  int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  k_loo_gram_small<E><<<grid, 256, 0, st>>>(S, P, len, out);
}

// Leave-one-out Gram on the FP64 tensor cores: out[a][c] += sum_x (S_a - P_a)(x) (S_c - P_c)(x) for up to
// 8 NT estimates.  One mma.m8n8k4.f64 forms all 64 pair products of an 8-row tile over four positions; for a Gram
// the A fragment (row = lane / 4, k = lane % 4) and the B fragment (k = lane % 4, column = lane / 4) of a tile are the
// same register.  Lane (e, c) reads the positions 16 it + 4 c .. + 3 of row e as one float4, so k-step j of an
// iteration covers the positions {4 c + j}.  Differences are formed in fp64 from the fp32 inputs (exact); ~40
// registers per thread leave the whole SM to loads in flight.  Memory bound: reads S and P once.
// NBLK block partials P, P + p_stride, ... share one read of S (outputs out, out + out_stride, ...).
template <int NT, int NBLK, int NX>
__global__ void __launch_bounds__(256)
k_loo_gram_mma(const float* __restrict__ S, const float* __restrict__ P, int64_t p_stride, int E, int64_t len,
               double* __restrict__ out, int64_t out_stride) {
  // E = 8 NT + NX estimates: the first 8 NT rows go through the FP64 tensor cores (for a Gram the A and B fragments of
  // an 8-row tile are the same registers), NX (0 or 1) trailing rows through plain FP64 FMAs -- GENIE's 2 K + 1
  // estimates (K = 8: 17) would otherwise pay for a third, almost empty, row tile in every tile pair.
  constexpr int NPAIR = NT * (NT + 1) / 2;
  double acc[NBLK][NPAIR][2];
  double accx[NBLK][NX ? NT : 1], accxx[NBLK];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) {
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) acc[b][p][0] = acc[b][p][1] = 0.0;
#pragma unroll
    for (int t = 0; t < (NX ? NT : 1); ++t) accx[b][t] = 0.0;
    accxx[b] = 0.0;
  }
  const int lane = threadIdx.x & 31, e = lane >> 2, c = lane & 3;
  const int64_t n16 = len >> 4;                      // len is a multiple of 16 (B * Np, Np a multiple of 512)
  const int64_t warp_id = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t it = warp_id; it < n16; it += n_warps) {
    float4 v[NT + NX], w[NBLK][NT + NX];
#pragma unroll
    for (int t = 0; t < NT + NX; ++t) {
      const int row = t < NT ? 8 * t + e : 8 * NT;   // the trailing row: every lane reads its own four columns of it
      v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < NBLK; ++b) w[b][t] = v[t];
      if (row < E) {
        const size_t o = (size_t)row * len + (size_t)it * 16 + 4 * c;
        v[t] = __ldg(reinterpret_cast<const float4*>(S + o));      // S is re-read by every launch: keep it cacheable
        if (P) {
#pragma unroll
          for (int b = 0; b < NBLK; ++b) w[b][t] = __ldcs(reinterpret_cast<const float4*>(P + (size_t)b * p_stride + o));
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      double d[NT + NX][4];
#pragma unroll
      for (int t = 0; t < NT + NX; ++t) {
        d[t][0] = (double)v[t].x - (double)w[b][t].x; d[t][1] = (double)v[t].y - (double)w[b][t].y;
        d[t][2] = (double)v[t].z - (double)w[b][t].z; d[t][3] = (double)v[t].w - (double)w[b][t].w;
      }
      int p = 0;
#pragma unroll
      for (int ti = 0; ti < NT; ++ti)
#pragma unroll
        for (int tj = ti; tj < NT; ++tj) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(acc[b][p][0]), "+d"(acc[b][p][1]) : "d"(d[ti][j]), "d"(d[tj][j]));
          ++p;
        }
      if (NX) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) accx[b][t] = fma(d[t][j], d[NT][j], accx[b][t]);
#pragma unroll
        for (int j = 0; j < 4; ++j) accxx[b] = fma(d[NT][j], d[NT][j], accxx[b]);
      }
    }
  }
  // C fragment: row = lane / 4, columns 2 (lane % 4), + 1 of the (ti, tj) tile; block reduction, then one atomic per entry
  __shared__ double red[NPAIR][64];
  __shared__ double redx[8 * NT + 1];
#pragma unroll
  for (int b = 0; b < NBLK; ++b) {
    __syncthreads();
    for (int i = threadIdx.x; i < NPAIR * 64; i += blockDim.x) (&red[0][0])[i] = 0.0;
    if (threadIdx.x < 8 * NT + 1) redx[threadIdx.x] = 0.0;
    __syncthreads();
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
      atomicAdd(&red[p][e * 8 + 2 * c], acc[b][p][0]);
      atomicAdd(&red[p][e * 8 + 2 * c + 1], acc[b][p][1]);
    }
    if (NX) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {                 // sum the four column quads of row 8 t + e
        double x = accx[b][t];
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        if (c == 0) atomicAdd(&redx[8 * t + e], x);
      }
      double x = accxx[b];                           // every row group e holds the same four partial sums
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      if (lane == 0) atomicAdd(&redx[8 * NT], x);
    }
    __syncthreads();
    double* ob = out + (size_t)b * out_stride;
    int p = 0;
    for (int ti = 0; ti < NT; ++ti)
      for (int tj = ti; tj < NT; ++tj, ++p)
        for (int i = threadIdx.x; i < 64; i += blockDim.x) {
          const int a = 8 * ti + (i >> 3), bb = 8 * tj + (i & 7);
          if (a < E && bb < E) {
            atomicAdd(ob + a * E + bb, red[p][i]);
            if (ti != tj) atomicAdd(ob + bb * E + a, red[p][i]);
          }
        }
    if (NX && 8 * NT < E) {
      const int x = 8 * NT;
      for (int i = threadIdx.x; i < 8 * NT; i += blockDim.x) {
        atomicAdd(ob + x * E + i, redx[i]);
        atomicAdd(ob + i * E + x, redx[i]);
      }
      if (threadIdx.x == 0) atomicAdd(ob + x * E + x, redx[x]);
    }
  }
}

// Row tiling of n_est estimates: NT 8-row tensor-core tiles + NX trailing rows on the FP64 FMA pipe
static inline void loo_tiling(int n_est, int* nt, int* nx) {
  if (n_est > 8 && n_est % 8 == 1) { *nt = n_est / 8; *nx = 1; }
  else { *nt = (n_est + 7) / 8; *nx = 0; }
}

template <int NT, int NX>
static void launch_loo_mma(const float* S, const float* P, int64_t p_stride, int n_blk, int E, int64_t len, double* out,
                           int64_t out_stride, cudaStream_t st) {
  const int grid = 148 * 8;
  if constexpr (NT <= 2) {
    if (n_blk == 4) { k_loo_gram_mma<NT, 4, NX><<<grid, 256, 0, st>>>(S, P, p_stride, E, len, out, out_stride); return; }
    if (n_blk == 3) { k_loo_gram_mma<NT, 3, NX><<<grid, 256, 0, st>>>(S, P, p_stride, E, len, out, out_stride); return; }
  }
  if (n_blk == 2) k_loo_gram_mma<NT, 2, NX><<<grid, 256, 0, st>>>(S, P, p_stride, E, len, out, out_stride);
  else k_loo_gram_mma<NT, 1, NX><<<grid, 256, 0, st>>>(S, P, p_stride, E, len, out, out_stride);
}

// n_est <= 24; n_blk <= loo_share(n_est) blocks share the read of S
static inline int loo_share(int n_est) {
  int nt, nx;
  loo_tiling(n_est, &nt, &nx);
  return nt <= 2 ? 4 : 2;
}
static void launch_loo(const float* S, const float* P, int64_t p_stride, int n_blk, int E, int64_t len, double* out,
                       int64_t out_stride, cudaStream_t st) {
  int nt, nx;
  loo_tiling(E, &nt, &nx);
  if (nt == 1 && nx == 0) launch_loo_mma<1, 0>(S, P, p_stride, n_blk, E, len, out, out_stride, st);
  else if (nt == 1) launch_loo_mma<1, 1>(S, P, p_stride, n_blk, E, len, out, out_stride, st);
  else if (nt == 2 && nx == 0) launch_loo_mma<2, 0>(S, P, p_stride, n_blk, E, len, out, out_stride, st);
  else if (nt == 2) launch_loo_mma<2, 1>(S, P, p_stride, n_blk, E, len, out, out_stride, st);
  else launch_loo_mma<3, 0>(S, P, p_stride, n_blk, E, len, out, out_stride, st);
}

// ------------------------------------------------------------------------------------------
// C ABI

static int validate(const rhe_config* c) {
  if (!c) { rhe_set_error("config is NULL"); return RHE_ERR_INVALID; }
  if (c->n_indv <= 0 || c->n_kept <= 0 || c->n_kept > c->n_indv) { rhe_set_error("bad n_indv / n_kept"); return RHE_ERR_INVALID; }
  if (c->pitch_bytes % 128 != 0 || (int64_t)c->pitch_bytes * 4 < c->n_indv) { rhe_set_error("pitch_bytes must be a multiple of 128 and cover n_indv"); return RHE_ERR_INVALID; }
  if (c->n_sets < 1 || c->n_sets > 2 || c->n_ops < 1 || c->n_ops > 2 || (c->n_sets == 2 && c->n_ops == 2)) { rhe_set_error("bad n_sets / n_ops"); return RHE_ERR_INVALID; }
  if (c->n_vec < 1 || c->n_vec >= c->n_cols_set) { rhe_set_error("need 1 <= n_vec < n_cols_set"); return RHE_ERR_INVALID; }
  if (c->n_bins < 1 || c->max_block_snps < 1) { rhe_set_error("bad n_bins / max_block_snps"); return RHE_ERR_INVALID; }
  if (c->kernel_path != RHE_PATH_SIMT && c->kernel_path != RHE_PATH_TCGEN05) { rhe_set_error("unknown kernel_path"); return RHE_ERR_INVALID; }
  return RHE_OK;
}

extern "C" int rhe_ctx_create(rhe_ctx** out, const rhe_config* cfg) {
  if (!out) { rhe_set_error("out is NULL"); return RHE_ERR_INVALID; }
  *out = nullptr;
  int rc = validate(cfg);
  if (rc) return rc;
  RHE_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  RHE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    rhe_set_error("pyrhe_b200 targets sm_100a (B200); device %d is sm_%d%d", cfg->device, prop.major, prop.minor);
    return RHE_ERR_UNSUPPORTED;
  }
  rhe_ctx* c = new (std::nothrow) rhe_ctx();
  if (!c) { rhe_set_error("out of host memory"); return RHE_ERR_INVALID; }
  c->cfg = *cfg;
  c->Np = cfg->pitch_bytes * 4;
  c->R1 = cfg->n_sets * cfg->n_cols_set;
  c->n_groups = cfg->n_ops * cfg->n_sets;
  c->E_reg = c->n_groups * cfg->n_bins;
  const size_t m = cfg->max_block_snps, Rs = cfg->n_cols_set, B = cfg->n_vec;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
  alloc((void**)&c->colsum, sizeof(double) * c->R1);
  alloc((void**)&c->counts, sizeof(int32_t) * 4 * m);
  alloc((void**)&c->fill, m);
  alloc((void**)&c->mu, sizeof(double) * m);
  alloc((void**)&c->f2, sizeof(double) * m);
  alloc((void**)&c->t_raw, sizeof(double) * cfg->n_ops * m * c->R1);
  alloc((void**)&c->t_std, sizeof(double) * c->n_groups * m * Rs);
  alloc((void**)&c->w1, sizeof(float) * c->n_groups * m * B);
  alloc((void**)&c->w2, sizeof(float) * c->n_groups * m * B);
  alloc((void**)&c->shiftv, sizeof(double) * c->n_groups * m * B);
  alloc((void**)&c->cs, sizeof(double) * c->E_reg * B);
  alloc((void**)&c->bin_off, sizeof(int32_t) * (cfg->n_bins + 1));
  if (e != cudaSuccess) {
    rhe_set_error("workspace allocation failed: %s", cudaGetErrorString(e));
    rhe_ctx_destroy(c);
    return RHE_ERR_CUDA;
  }
  if (cfg->kernel_path == RHE_PATH_TCGEN05) {
    rc = rhe_tc_create(c);
    if (rc) { rhe_ctx_destroy(c); return rc; }
  }
  *out = c;
  return RHE_OK;
}

extern "C" int rhe_ctx_destroy(rhe_ctx* c) {
  if (!c) return RHE_OK;
  cudaSetDevice(c->cfg.device);
  cudaDeviceSynchronize();
  if (c->tc) rhe_tc_destroy(c);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  void* ptrs[] = {c->colsum, c->counts, c->fill, c->mu, c->f2, c->t_raw, c->t_std, c->w1, c->w2, c->shiftv, c->cs, c->bin_off};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete c;
  return RHE_OK;
}

extern "C" int rhe_set_rhs(rhe_ctx* c, const float* rhs, const float* rowscale, const uint32_t* keep2, void* stream) {
  if (!c || !rhs || !rowscale || !keep2) { rhe_set_error("rhe_set_rhs: NULL argument"); return RHE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  c->rhs = rhs;
  c->rowscale = rowscale;
  c->keep2 = keep2;
  k_colsum<<<c->R1, 256, 0, st>>>(rhs, c->Np, c->colsum);
  RHE_LAUNCH_CHECK(c);
  if (c->tc) return rhe_tc_set_rhs(c, st);
  return RHE_OK;
}

extern "C" int rhe_set_uniforms(rhe_ctx* c, const double* u, int32_t count) {
  if (!c) { rhe_set_error("ctx is NULL"); return RHE_ERR_INVALID; }
  c->uniforms = u;
  c->n_uniforms = count;
  return RHE_OK;
}

extern "C" int rhe_upload_rows(const void* host_src, int64_t row_bytes, int64_t n_rows, void* dev_dst,
                               int64_t pitch_bytes, void* stream) {
  if (!host_src || !dev_dst || row_bytes <= 0 || n_rows < 0 || pitch_bytes < row_bytes) {
    rhe_set_error("rhe_upload_rows: bad argument");
    return RHE_ERR_INVALID;
  }
  if (n_rows == 0) return RHE_OK;
  RHE_CUDA(cudaMemcpy2DAsync(dev_dst, (size_t)pitch_bytes, host_src, (size_t)row_bytes, (size_t)row_bytes,
                             (size_t)n_rows, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return RHE_OK;
}

static int check_block(rhe_ctx* c, const void* bed, int m, const char* who) {
  if (!c || !bed) { rhe_set_error("%s: NULL argument", who); return RHE_ERR_INVALID; }
  if (m < 1 || m > c->cfg.max_block_snps) { rhe_set_error("%s: n_snps %d outside [1, %d]", who, m, c->cfg.max_block_snps); return RHE_ERR_INVALID; }
  if (!c->keep2) { rhe_set_error("%s: rhe_set_rhs has not been called", who); return RHE_ERR_STATE; }
  return RHE_OK;
}

static int run_stats(rhe_ctx* c, const uint8_t* bed, int m, int32_t* counts, cudaStream_t st) {
  k_stats<<<rhe_div_up(m, 8), 256, 0, st>>>(bed, c->cfg.pitch_bytes, m, c->keep2, c->cfg.n_kept, counts);
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

static int run_params(rhe_ctx* c, const uint8_t* bed, int m, cudaStream_t st) {
  int rc = run_stats(c, bed, m, c->counts, st);
  if (rc) return rc;
  if (c->cfg.impute_binary && (!c->uniforms || c->n_uniforms < m)) {
    rhe_set_error("binary imputation needs rhe_set_uniforms with >= %d values", m);
    return RHE_ERR_STATE;
  }
  k_snp_params<<<rhe_div_up(m, 256), 256, 0, st>>>(c->counts, m, c->cfg.n_kept, c->cfg.impute_binary,
                                                    c->uniforms, c->fill, c->mu, c->f2);
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

extern "C" int rhe_block_stats(rhe_ctx* c, const uint8_t* bed, int32_t m, int32_t* counts, void* stream) {
  int rc = check_block(c, bed, m, "rhe_block_stats");
  if (rc) return rc;
  if (!counts) { rhe_set_error("rhe_block_stats: counts is NULL"); return RHE_ERR_INVALID; }
  return run_stats(c, bed, m, counts, (cudaStream_t)stream);
}

extern "C" int rhe_decode_block(rhe_ctx* c, const uint8_t* bed, int32_t m, int32_t apply_impute, int8_t* out, void* stream) {
  int rc = check_block(c, bed, m, "rhe_decode_block");
  if (rc) return rc;
  if (!out) { rhe_set_error("rhe_decode_block: out is NULL"); return RHE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (apply_impute) { rc = run_params(c, bed, m, st); if (rc) return rc; }
  k_decode<<<296, 256, 0, st>>>(bed, c->cfg.pitch_bytes, m, c->fill, apply_impute, out);
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

template <int BG>
static void launch_pass_b(rhe_ctx* c, dim3 grid, cudaStream_t st, const uint8_t* bed, int m, const int32_t* rows,
                          const int32_t* off, float* P_out, float* S_accum) {
  k_pass_b<BG><<<grid, 512, 0, st>>>(bed, c->cfg.pitch_bytes, m, c->Np, c->cfg.n_vec, c->cfg.n_bins, c->cfg.n_ops,
                                     rows, off, c->fill, c->w1, c->w2, c->cs, c->rowscale, P_out, S_accum);
}

extern "C" int rhe_tc_supported(const rhe_config* cfg) {
  if (validate(cfg)) return 0;
  return rhe_tc_check(cfg, 0) == RHE_OK ? 1 : 0;
}

extern "C" int rhe_block_plan_create(rhe_ctx* c, int32_t m, const int32_t* bin_rows, const int32_t* bin_off_host,
                                     void* stream, rhe_block_plan** out) {
  if (!c || !bin_rows || !bin_off_host || !out) { rhe_set_error("rhe_block_plan_create: NULL argument"); return RHE_ERR_INVALID; }
  *out = nullptr;
  const int K = c->cfg.n_bins;
  if (m < 1 || m > c->cfg.max_block_snps) { rhe_set_error("rhe_block_plan_create: n_snps %d outside [1, %d]", m, c->cfg.max_block_snps); return RHE_ERR_INVALID; }
  if (bin_off_host[0] != 0) { rhe_set_error("rhe_block_plan_create: bin_offsets[0] must be 0"); return RHE_ERR_INVALID; }
  for (int k = 0; k < K; ++k)
    if (bin_off_host[k + 1] < bin_off_host[k] || bin_off_host[k + 1] - bin_off_host[k] > m) {
      rhe_set_error("rhe_block_plan_create: bin %d has a bad row range", k);
      return RHE_ERR_INVALID;
    }
  RHE_CUDA(cudaSetDevice(c->cfg.device));
  rhe_block_plan* p = new (std::nothrow) rhe_block_plan();
  if (!p) { rhe_set_error("out of host memory"); return RHE_ERR_INVALID; }
  p->m = m;
  p->bin_rows = bin_rows;
  p->off_host.assign(bin_off_host, bin_off_host + K + 1);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMalloc((void**)&p->off_dev, sizeof(int32_t) * (K + 1));
  if (e == cudaSuccess) e = cudaMemcpyAsync(p->off_dev, p->off_host.data(), sizeof(int32_t) * (K + 1), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    rhe_set_error("rhe_block_plan_create: %s", cudaGetErrorString(e));
    rhe_block_plan_destroy(c, p);
    return RHE_ERR_CUDA;
  }
  if (c->tc) {
    int rc = rhe_tc_plan_create(c, p, st);
    if (rc) { rhe_block_plan_destroy(c, p); return rc; }
  }
  e = cudaStreamSynchronize(st);     // the offsets were copied from the plan's own host vector; be done before returning
  if (e != cudaSuccess) { rhe_set_error("rhe_block_plan_create: %s", cudaGetErrorString(e)); rhe_block_plan_destroy(c, p); return RHE_ERR_CUDA; }
  *out = p;
  return RHE_OK;
}

extern "C" int rhe_block_plan_destroy(rhe_ctx* c, rhe_block_plan* p) {
  if (!p) return RHE_OK;
  if (c) cudaSetDevice(c->cfg.device);
  if (p->tc) rhe_tc_plan_destroy(p);
  if (p->off_dev) cudaFree(p->off_dev);
  delete p;
  return RHE_OK;
}

extern "C" int64_t rhe_block_fast_bytes(const rhe_ctx* c, const rhe_block_plan* plan) {
  if (!c || !plan || !c->tc) return 0;
  return rhe_tc_gt_bytes(c, plan);
}

extern "C" int rhe_block_transpose(rhe_ctx* c, const uint8_t* bed, const rhe_block_plan* plan, const int32_t* counts,
                                   uint8_t* gt, void* stream) {
  if (!c || !bed || !plan || !counts || !gt) { rhe_set_error("rhe_block_transpose: NULL argument"); return RHE_ERR_INVALID; }
  if (!c->tc) { rhe_set_error("rhe_block_transpose: the context runs the CUDA-core path"); return RHE_ERR_UNSUPPORTED; }
  return rhe_tc_transpose(c, bed, plan, counts, gt, (cudaStream_t)stream);
}

extern "C" int64_t rhe_block_tiled_bytes(const rhe_ctx* c, const rhe_block_plan* plan) {
  if (!c || !plan || !c->tc) return 0;
  return rhe_tc_tiled_bytes(c, plan->m);
}

extern "C" int rhe_block_retile(rhe_ctx* c, uint8_t* bed, const rhe_block_plan* plan, const int32_t* counts, uint8_t* scratch,
                                void* stream) {
  if (!c || !bed || !plan || !counts || !scratch) { rhe_set_error("rhe_block_retile: NULL argument"); return RHE_ERR_INVALID; }
  if (!c->tc) { rhe_set_error("rhe_block_retile: the context runs the CUDA-core path"); return RHE_ERR_UNSUPPORTED; }
  return rhe_tc_retile(c, bed, plan->m, counts, scratch, (cudaStream_t)stream);
}

extern "C" int rhe_block_accumulate(rhe_ctx* c, const uint8_t* bed, const rhe_block_plan* plan, const int32_t* counts_in,
                                    const uint8_t* gt, int32_t rows_layout, float* P_out, float* S_accum, double* gram_out,
                                    void* stream) {
  if (!plan) { rhe_set_error("rhe_block_accumulate: plan is NULL"); return RHE_ERR_INVALID; }
  const int m = plan->m;
  int rc = check_block(c, bed, m, "rhe_block_accumulate");
  if (rc) return rc;
  if (!gram_out) { rhe_set_error("rhe_block_accumulate: NULL argument"); return RHE_ERR_INVALID; }
  if (gt && c->cfg.kernel_path != RHE_PATH_TCGEN05) { rhe_set_error("rhe_block_accumulate: the individual-major copy belongs to the tensor-core path"); return RHE_ERR_INVALID; }
  if (rows_layout != RHE_ROWS_PLINK && rows_layout != RHE_ROWS_TILED) { rhe_set_error("rhe_block_accumulate: unknown rows_layout %d", rows_layout); return RHE_ERR_INVALID; }
  if (rows_layout == RHE_ROWS_TILED && (c->cfg.kernel_path != RHE_PATH_TCGEN05 || !counts_in || !gt)) {
    // re-tiled rows carry imputed counts in box order: only pass A of the tensor path reads them, so the allele counts
    // and pass B's operand must come from the ingest-time artefacts
    rhe_set_error("rhe_block_accumulate: re-tiled rows need the tensor-core path, the ingest-time counts and the individual-major copy");
    return RHE_ERR_INVALID;
  }
  const int32_t* bin_rows = plan->bin_rows;
  const int32_t* s_off_dev = plan->off_dev;
  cudaStream_t st = (cudaStream_t)stream;
  const rhe_config& g = c->cfg;
  const int K = g.n_bins, Rs = g.n_cols_set, B = g.n_vec;
  cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  if (c->timing) {
    for (auto& e : tev) { RHE_CUDA(cudaEventCreate(&e)); c->ev.push_back(e); }
    RHE_CUDA(cudaEventRecord(tev[0], st));
  }
  if (g.impute_binary && (!c->uniforms || c->n_uniforms < m)) {
    rhe_set_error("binary imputation needs rhe_set_uniforms with >= %d values", m);
    return RHE_ERR_STATE;
  }
  unsigned int* wmax = g.kernel_path == RHE_PATH_TCGEN05 ? rhe_tc_wmax(c) : nullptr;
  if (counts_in) {     // allele counts resident since ingest: the block is not read here
    k_params_from_counts<<<rhe_div_up(m, 256) < 64 ? 64 : rhe_div_up(m, 256), 256, 0, st>>>(
        counts_in, m, g.n_kept, g.impute_binary, c->uniforms, c->fill, c->mu, c->f2, c->t_raw, g.n_ops * m * c->R1, c->cs,
        c->E_reg * B, gram_out, c->E_reg * Rs * Rs, wmax, wmax ? c->n_groups * B : 0);
  } else {
    k_stats_params<<<rhe_div_up(m, ST_ROWS), 256, 0, st>>>(bed, g.pitch_bytes, m, c->keep2, g.n_kept, g.impute_binary, c->uniforms,
                                                     c->counts, c->fill, c->mu, c->f2, c->t_raw, c->R1, g.n_ops, c->cs,
                                                     c->E_reg * B, gram_out, c->E_reg * Rs * Rs, wmax, wmax ? c->n_groups * B : 0);
  }
  RHE_LAUNCH_CHECK(c);
  if (c->timing) RHE_CUDA(cudaEventRecord(tev[1], st));

  // ---- pass A
  if (g.kernel_path == RHE_PATH_TCGEN05) {
    rc = rhe_tc_pass_a(c, bed, m, rows_layout == RHE_ROWS_TILED, st);
    if (rc) return rc;
  } else {
    int chunk = 8192;
    const int tiles = rhe_div_up(m, 32), cgroups = rhe_div_up(c->R1, 16);
    while (chunk > 512 && (int64_t)tiles * cgroups * rhe_div_up(c->Np, chunk) < 592) chunk >>= 1;
    dim3 grid(tiles, rhe_div_up(c->Np, chunk), cgroups);
    k_pass_a<0><<<grid, 256, 0, st>>>(bed, g.pitch_bytes, m, c->rhs, c->Np, c->R1, c->fill, c->t_raw, chunk);
    RHE_LAUNCH_CHECK(c);
    if (g.n_ops == 2) {
      k_pass_a<1><<<grid, 256, 0, st>>>(bed, g.pitch_bytes, m, c->rhs, c->Np, c->R1, c->fill,
                                         c->t_raw + (size_t)m * c->R1, chunk);
      RHE_LAUNCH_CHECK(c);
    }
  }
  if (c->timing) RHE_CUDA(cudaEventRecord(tev[2], st));
  // ---- standardise, per-bin Gram, pass-B weights
  {
    int total = c->n_groups * m * Rs;
    k_standardize<<<rhe_div_up(total, 256), 256, 0, st>>>(m, Rs, c->R1, B, g.n_ops, g.n_sets, c->t_raw, c->colsum,
                                                           c->mu, c->f2, c->t_std, c->w1, c->w2, c->shiftv, wmax);
    RHE_LAUNCH_CHECK(c);
  }
  k_bin_gram<<<dim3(c->E_reg, 64), 256, 0, st>>>(m, Rs, B, K, bin_rows, s_off_dev, c->t_std, c->shiftv, gram_out, c->cs);
  RHE_LAUNCH_CHECK(c);
  if (c->timing) RHE_CUDA(cudaEventRecord(tev[3], st));
  // ---- pass B
  if (P_out || S_accum) {
    if (g.kernel_path == RHE_PATH_TCGEN05) {
      rc = rhe_tc_pass_b(c, bed, gt, plan, P_out, S_accum, st);
      if (rc) return rc;
    } else {
      int BG = B <= 4 ? 4 : B <= 8 ? 8 : B <= 12 ? 12 : 16;
      dim3 grid(rhe_div_up(c->Np, 512), c->E_reg, rhe_div_up(B, BG));
      switch (BG) {
        case 4: launch_pass_b<4>(c, grid, st, bed, m, bin_rows, s_off_dev, P_out, S_accum); break;
        case 8: launch_pass_b<8>(c, grid, st, bed, m, bin_rows, s_off_dev, P_out, S_accum); break;
        case 12: launch_pass_b<12>(c, grid, st, bed, m, bin_rows, s_off_dev, P_out, S_accum); break;
        default: launch_pass_b<16>(c, grid, st, bed, m, bin_rows, s_off_dev, P_out, S_accum); break;
      }
      RHE_LAUNCH_CHECK(c);
    }
  }
  if (c->timing) RHE_CUDA(cudaEventRecord(tev[4], st));
  return RHE_OK;
}

// S = sum_j P_j over the stored block partials, in block order (fp32, one streaming read of the partials): the totals of
// base.py:483-486 without a read-modify-write of S in every block's pass B, and bit-reproducible from run to run.
__global__ void __launch_bounds__(256)
k_sum_partials(const float4* __restrict__ P, int64_t p_stride4, int n_blocks, int64_t len4, float4* __restrict__ S) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = 0;
    for (; j + 4 <= n_blocks; j += 4) {                // four loads in flight per thread
      const float4 a = __ldcs(P + (size_t)j * p_stride4 + i), b = __ldcs(P + (size_t)(j + 1) * p_stride4 + i);
      const float4 c = __ldcs(P + (size_t)(j + 2) * p_stride4 + i), d = __ldcs(P + (size_t)(j + 3) * p_stride4 + i);
      acc.x = ((acc.x + a.x) + b.x) + c.x + d.x; acc.y = ((acc.y + a.y) + b.y) + c.y + d.y;
      acc.z = ((acc.z + a.z) + b.z) + c.z + d.z; acc.w = ((acc.w + a.w) + b.w) + c.w + d.w;
    }
    for (; j < n_blocks; ++j) {
      const float4 a = __ldcs(P + (size_t)j * p_stride4 + i);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    S[i] = acc;
  }
}

extern "C" int rhe_sum_partials(rhe_ctx* c, const float* P, int64_t p_stride, int32_t n_blocks, int64_t len, float* S,
                                void* stream) {
  if (!c || !P || !S) { rhe_set_error("rhe_sum_partials: NULL argument"); return RHE_ERR_INVALID; }
  if (n_blocks < 1 || len < 0 || len % 4 != 0 || p_stride % 4 != 0) { rhe_set_error("rhe_sum_partials: bad block count / length / stride"); return RHE_ERR_INVALID; }
  if (len == 0) return RHE_OK;
  const int64_t len4 = len / 4;
  const int grid = (int)(len4 / 256 + 1 < 148 * 16 ? len4 / 256 + 1 : 148 * 16);
  k_sum_partials<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(P), p_stride / 4, n_blocks, len4,
                                                         reinterpret_cast<float4*>(S));
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

extern "C" int rhe_loo_gram_multi(rhe_ctx* c, const float* S, const float* P, int64_t p_stride, int32_t n_blocks,
                                  int32_t n_est, int64_t len, double* out, int64_t out_stride, void* stream) {
  if (!c || !S || !P || !out) { rhe_set_error("rhe_loo_gram_multi: NULL argument"); return RHE_ERR_INVALID; }
  if (n_blocks < 1 || p_stride < 0 || out_stride < (int64_t)n_est * n_est) { rhe_set_error("rhe_loo_gram_multi: bad block count / strides"); return RHE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const bool mma = len % 16 == 0 && n_est >= 1 && n_est <= 24 && p_stride % 4 == 0 && !RHE_DBG_ENV("PYRHE_B200_LOO_SIMT", 0);
  const int share = mma ? loo_share(n_est) : 1;      // blocks that share one read of S (three 8-row tiles: registers for two)
  for (int b0 = 0; b0 < n_blocks;) {
    const int nb = mma ? (n_blocks - b0 < share ? n_blocks - b0 : share) : 1;
    const float* Pb = P + (size_t)b0 * p_stride;
    double* ob = out + (size_t)b0 * out_stride;
    if (nb == 1) {
      int rc = rhe_loo_gram(c, S, Pb, n_est, len, ob, stream);
      if (rc) return rc;
    } else {
      for (int b = 0; b < nb; ++b) RHE_CUDA(cudaMemsetAsync(ob + (size_t)b * out_stride, 0, sizeof(double) * n_est * n_est, st));
      launch_loo(S, Pb, p_stride, nb, n_est, len, ob, out_stride, st);
      RHE_LAUNCH_CHECK(c);
    }
    b0 += nb;
  }
  return RHE_OK;
}

extern "C" int rhe_synth_genotypes(uint8_t* bed, int64_t n_rows, int64_t pitch, int32_t n_indv, int64_t first_snp,
                                   uint64_t seed, float missing_rate, void* stream) {
  if (!bed || n_rows < 0 || pitch <= 0 || n_indv <= 0 || (int64_t)(n_indv + 3) / 4 > pitch) {
    rhe_set_error("rhe_synth_genotypes: bad argument");
    return RHE_ERR_INVALID;
  }
  if (n_rows == 0) return RHE_OK;
  k_synth<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(bed, n_rows, pitch, n_indv, first_snp, seed, missing_rate);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rhe_set_error("k_synth launch failed: %s", cudaGetErrorString(e)); return RHE_ERR_CUDA; }
  return RHE_OK;
}

extern "C" int rhe_timing_enable(rhe_ctx* c, int32_t enable) {
  if (!c) { rhe_set_error("ctx is NULL"); return RHE_ERR_INVALID; }
  c->timing = enable != 0;
  return RHE_OK;
}

extern "C" int rhe_timing_collect(rhe_ctx* c, double* phases_ms, int32_t* n_calls) {
  if (!c || !phases_ms || !n_calls) { rhe_set_error("rhe_timing_collect: NULL argument"); return RHE_ERR_INVALID; }
  for (int i = 0; i < 4; ++i) phases_ms[i] = 0.0;
  *n_calls = (int32_t)(c->ev.size() / 5);
  for (size_t k = 0; k + 4 < c->ev.size(); k += 5) {
    RHE_CUDA(cudaEventSynchronize(c->ev[k + 4]));
    for (int i = 0; i < 4; ++i) {
      float ms = 0.f;
      RHE_CUDA(cudaEventElapsedTime(&ms, c->ev[k + i], c->ev[k + i + 1]));
      phases_ms[i] += (double)ms;
    }
  }
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  c->ev.clear();
  return RHE_OK;
}

extern "C" int rhe_loo_gram(rhe_ctx* c, const float* S, const float* P, int32_t n_est, int64_t len, double* out, void* stream) {
  if (!c || !S || !out) { rhe_set_error("rhe_loo_gram: NULL argument"); return RHE_ERR_INVALID; }
  if (n_est < 1 || n_est > GRAM_MAX_E) { rhe_set_error("rhe_loo_gram: n_est %d outside [1, %d]", n_est, GRAM_MAX_E); return RHE_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  RHE_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * n_est * n_est, st));
  // register-resident kernels for up to 8 estimates (8-byte loads: even length, always true for B * Np)
  if (len % 16 == 0 && n_est <= 24 && !RHE_DBG_ENV("PYRHE_B200_LOO_SIMT", 0)) {      // FP64 tensor-core Gram
    launch_loo(S, P, 0, 1, n_est, len, out, 0, st);
    RHE_LAUNCH_CHECK(c);
    return RHE_OK;
  }
  switch (len % 2 == 0 ? n_est : 0) {
    case 1: launch_loo_small<1>(S, P, len, out, st); break;
    case 2: launch_loo_small<2>(S, P, len, out, st); break;
    case 3: launch_loo_small<3>(S, P, len, out, st); break;
    case 4: launch_loo_small<4>(S, P, len, out, st); break;
    case 5: launch_loo_small<5>(S, P, len, out, st); break;
    case 6: launch_loo_small<6>(S, P, len, out, st); break;
    case 7: launch_loo_small<7>(S, P, len, out, st); break;
    case 8: launch_loo_small<8>(S, P, len, out, st); break;
    default: {
      int64_t nchunks = (len + GRAM_CHUNK - 1) / GRAM_CHUNK;
      int grid = (int)(nchunks < 148 * 4 ? nchunks : 148 * 4);
      k_loo_gram<<<grid, 256, 0, st>>>(S, P, n_est, len, out);
    }
  }
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}
