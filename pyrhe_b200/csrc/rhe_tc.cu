// Tensor-core path (sm_100a): fused 2-bit decode + int8 tcgen05.mma with int32 accumulators in TMEM.
//
// Both products of the block path are skinny GEMMs whose big operand is the genotype matrix:
//   pass A   T[s, n] = sum_i G[i, s] * Rq[i, n]      M = 128 SNPs,        K = individuals, N = limb columns
//   pass B   D[i, n] = sum_s G[i, s] * Uq[s, n]      M = 128 individuals, K = SNPs,        N = limb columns
// G is exact in int8 ({0,1,2}); the fp32 right-hand sides are fixed-point numbers split into L signed
// 8-bit limbs stacked along N, so the int32 accumulation is EXACT and the result is independent of how
// the work is split.  One decoded shared-memory tile serves both passes: rows = SNPs,
// 128 bytes = 128 individuals, 128B-swizzled.  Pass A reads it as a K-major A operand, pass B as an
// MN-major A operand (instruction-descriptor bit 15).  The small B operands (Rq / Uq tiles) arrive by
// TMA; accumulators live in TMEM and are read back with tcgen05.ld in the epilogue.
//
// Warp roles: TC_G groups of four decode warps (group g expands sub-tiles q = g mod TC_G of every super-stage into
// A slot g, so the groups overlap each other's load / fence / barrier latencies; afterwards they run the epilogue,
// one TMEM lane quadrant per warp), then one TMA warp and one warp that allocates TMEM and issues tcgen05.mma.
#include <cuda.h>
#include <cstdio>
#include "rhe_common.cuh"

#define TC_TILE_A 16384          // 128 rows x 128 bytes
#define TC_G 2                                  // decode groups of 4 warps; group g owns A slot g
#define TC_DECODE_WARPS (4 * TC_G)
#define TC_THREADS (32 * (TC_DECODE_WARPS + 2))

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  int L = 3;            // limbs per fixed-point value (PYRHE_B200_LIMBS)
  int F = 22;           // fixed-point magnitude bits: |q| <= 2^F, F = 8 L - 2
  int R1p = 0;          // RHS columns rounded up to 4
  int NBa = 0;          // pass A MMA N  = round16(L * R1p)
  int Bp = 0;           // pass-B columns rounded up to 4
  int NCb = 0;          // pass B MMA N per bin = round16(L * Bp)
  int8_t* rq = nullptr;       // [NBa][Np]   limb l of column c at row l * R1p + c, permuted individual order
  double* col_dq = nullptr;   // [R1]        power-of-two dequantisation factor of every RHS column
  int32_t* pos_rows = nullptr;  // [cap_pos]   block-local SNP row of every bin-sorted position (-1 = padding)
  int32_t* pstart = nullptr;    // [K + 1]     first position of every bin (multiples of 128)
  int8_t* uq = nullptr;         // [NCb][cap_pos] quantised pass-B weights
  unsigned int* wmax = nullptr; // [B]         max |weight| per column (float bits)
  int cap_pos = 0;
  CUtensorMap tm_rq, tm_uq;
  PFN_encodeTiled encode = nullptr;
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug traps (sticky launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  uint32_t zero = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// value[j] = sum_l 256^l * limb_l[j] for four adjacent columns; limb l sits `stride` columns further on
__device__ __forceinline__ void tmem_combine4(uint32_t taddr, int L, int stride, double (&val)[4]) {
  int32_t v[4][4];
#pragma unroll
  for (int l = 0; l < 4; ++l)
    if (l < L) tmem_ld4(taddr + (uint32_t)(l * stride), v[l]);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double acc = 0.0;
#pragma unroll
    for (int l = 3; l >= 0; --l)
      if (l < L) acc = acc * 256.0 + (double)v[l][j];
    val[j] = acc;
  }
}

// Shared-memory matrix descriptor, 128B swizzle, version 1 (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::i8: S32 accumulator, signed 8-bit A and B (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N, int a_mn_major) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | (0u << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Individuals are permuted inside every group of 16 so that one 32-bit packed word expands to four
// int8 words with 3 logic ops + 4 byte-permutes (no bit spreading): stored byte p of a 16-byte chunk
// holds individual PERM[p] of the group.
__device__ __forceinline__ int tc_perm16(int p) { return ((p & 7) << 1) | (p >> 3); }        // byte -> individual
__device__ __forceinline__ int tc_invperm16(int x) { return (x >> 1) | ((x & 1) << 3); }     // individual -> byte

// Expand one packed word (16 genotypes) into 16 int8 values with the per-SNP value table
// tab = {0, fill, 1, 2} (byte c = value of code c), in the permuted order above.
__device__ __forceinline__ uint4 tc_expand(uint32_t w, uint32_t tab) {
  const uint32_t e = w & 0x33333333u, o = (w >> 2) & 0x33333333u;
  uint4 r;
  r.x = __byte_perm(tab, 0, e);
  r.y = __byte_perm(tab, 0, e >> 16);
  r.z = __byte_perm(tab, 0, o);
  r.w = __byte_perm(tab, 0, o >> 16);
  return r;
}

// Explicit shared-state-space accesses (32-bit addresses): the dynamic-smem base is re-aligned with integer
// arithmetic, after which the compiler can no longer prove the address space and would emit generic LD/ST.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Decode 32 packed bytes (128 individuals of one SNP row) into row `r` of a swizzled 128 x 128B tile
// (`tile` is a shared-space address).
__device__ __forceinline__ void tc_store_row(uint32_t tile, int r, const uint4& lo, const uint4& hi, uint32_t tab) {
  const uint32_t row = tile + r * 128;
  const int x = r & 7;
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int c = 0; c < 8; ++c) sts128(row + ((c ^ x) << 4), tc_expand(w[c], tab));
}

__device__ __forceinline__ uint4 ldg_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// Ring depths.  A "super-stage" is 128 rows x 128 packed bytes (512 individuals): every decode thread
// owns one row and streams its own 128-byte lines with cp.async (thread-private, so no block barrier
// is needed to consume them), TC_PK super-stages deep.  Each super-stage is decoded into four int8
// A tiles (128 individuals each) that cycle through a ring of TC_AS slots.
#define TC_PK 3
#define TC_AS TC_G
#define TC_BS 4          // B-operand (TMA) ring: deep enough to hide the L2 -> smem latency
#define TC_PACKED (128 * 128)

struct TcSmem {
  uint64_t full_a[TC_AS], empty_a[TC_AS], full_b[TC_BS], empty_b[TC_BS], acc_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Thread t streams 128 bytes of its row into slot layout [chunk 0..7][thread][16 B] (conflict-free reads).
// Group g fetches only the chunks of the sub-tiles it decodes (q = g, g + TC_G, ...: chunks 2q, 2q + 1).
__device__ __forceinline__ void tc_issue_row(uint32_t slot, int t, int g, const uint8_t* src) {
#pragma unroll
  for (int q = 0; q < 4; q += TC_G) {
    const int c = 2 * (q + g);
    cp_async16(slot + c * 2048 + t * 16, src + c * 16);
    cp_async16(slot + (c + 1) * 2048 + t * 16, src + (c + 1) * 16);
  }
}

__device__ __forceinline__ void tc_setup(TcSmem* sm, int warp, uint32_t tmem_cols) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_AS; ++s) { mbar_init(&sm->full_a[s], 128); mbar_init(&sm->empty_a[s], 1); }
    for (int s = 0; s < TC_BS; ++s) { mbar_init(&sm->full_b[s], 1); mbar_init(&sm->empty_b[s], 1); }
    mbar_init(&sm->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == TC_DECODE_WARPS + 1) tmem_alloc(&sm->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}

// ------------------------------------------------------------------------------------------ pass A
// grid = (SNP tiles of 128, splits over individuals).  t_raw[s][c] += dq[c] * sum_i g_is * q_ic  (exact).
__global__ void __launch_bounds__(TC_THREADS, 2)
k_tc_pass_a(const __grid_constant__ CUtensorMap tm_rq, const uint8_t* __restrict__ bed, int pitch, int m, int Np,
            int NB, int R1, int R1p, int L, const uint8_t* __restrict__ fill, const double* __restrict__ col_dq,
            double* __restrict__ t_raw, int chunk, uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tileA = smem;
  uint8_t* tileB = tileA + TC_AS * TC_TILE_A;
  const int tileB_bytes = NB * 128;
  uint8_t* packed = tileB + TC_BS * tileB_bytes;
  TcSmem* sm = reinterpret_cast<TcSmem*>(packed + TC_PK * TC_PACKED);
  const uint32_t tileA_s = smem_u32(tileA), packed_s = smem_u32(packed);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int snp0 = blockIdx.x * 128;
  const int i_begin = blockIdx.y * chunk, i_end = min(Np, i_begin + chunk);
  const int n_ss = (i_end - i_begin) >> 9;          // super-stages of 512 individuals
  if (n_ss <= 0) return;
  const int n_sub = n_ss * 4;

  tc_setup(sm, warp, tmem_cols);
  const uint32_t tmem = sm->tmem_base;

  if (warp < TC_DECODE_WARPS) {
    const int t = threadIdx.x & 127, g = warp >> 2;
    const int s = min(snp0 + t, m - 1);
    const uint32_t tab = ((uint32_t)fill[s] << 8) | (1u << 16) | (2u << 24);
    const uint8_t* src = bed + (size_t)s * pitch + (i_begin >> 2);
#pragma unroll
    for (int pre = 0; pre < TC_PK - 1; ++pre) {
      if (pre < n_ss) tc_issue_row(packed_s + pre * TC_PACKED, t, g, src + pre * 128);
      cp_async_commit();
    }
    for (int ss = 0; ss < n_ss; ++ss) {
      const int nxt = ss + TC_PK - 1;
      if (nxt < n_ss) tc_issue_row(packed_s + (nxt % TC_PK) * TC_PACKED, t, g, src + (size_t)nxt * 128);
      cp_async_commit();
      cp_async_wait<TC_PK - 1>();
      const uint32_t slot = packed_s + (ss % TC_PK) * TC_PACKED + t * 16;
#pragma unroll
      for (int q0 = 0; q0 < 4; q0 += TC_G) {
        const int q = q0 + g;
        const int sub = ss * 4 + q, a = sub % TC_AS, use = sub / TC_AS;
        const uint4 lo = lds128(slot + (2 * q) * 2048);
        const uint4 hi = lds128(slot + (2 * q + 1) * 2048);
        mbar_wait(&sm->empty_a[a], (use & 1) ^ 1);
        tc_store_row(tileA_s + a * TC_TILE_A, t, lo, hi, tab);
        fence_proxy_async();
        mbar_arrive(&sm->full_a[a]);
      }
    }
    // ---- epilogue: lane quadrant `warp` of TMEM, row = SNP
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    const int snp = snp0 + t;
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c0 = 4 * g; c0 < R1p; c0 += 4 * TC_G) {
      double val[4];
      tmem_combine4(trow + (uint32_t)c0, L, R1p, val);
      if (snp < m) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < R1) atomicAdd(t_raw + (size_t)snp * R1 + c0 + j, val[j] * col_dq[c0 + j]);
      }
    }
    tc_fence_before();
  } else if (warp == TC_DECODE_WARPS) {
    if (lane == 0) {
      for (int sub = 0; sub < n_sub; ++sub) {
        const int b = sub % TC_BS, use = sub / TC_BS;
        mbar_wait(&sm->empty_b[b], (use & 1) ^ 1);
        mbar_expect_tx(&sm->full_b[b], (uint32_t)tileB_bytes);
        tma_load_2d(tileB + b * tileB_bytes, &tm_rq, &sm->full_b[b], i_begin + sub * 128, 0);
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = idesc_i8(128, NB, 0);
      for (int sub = 0; sub < n_sub; ++sub) {
        const int a = sub % TC_AS, use = sub / TC_AS, b = sub % TC_BS, useb = sub / TC_BS;
        mbar_wait(&sm->full_b[b], useb & 1);
        mbar_wait(&sm->full_a[a], use & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(tileA + a * TC_TILE_A), b0 = smem_u32(tileB + b * tileB_bytes);
#pragma unroll
        for (int j = 0; j < 4; ++j)   // K = 32 individuals per instruction: advance 32 bytes inside the swizzle atom
          umma_i8(tmem, smem_desc_sw128(a0 + j * 32, 16, 1024), smem_desc_sw128(b0 + j * 32, 16, 1024), idesc,
                  (uint32_t)((sub | j) != 0));
        umma_commit(&sm->empty_a[a]);
        umma_commit(&sm->empty_b[b]);
      }
      umma_commit(&sm->acc_full);
    }
  }
  __syncthreads();
  if (warp == TC_DECODE_WARPS + 1) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// ------------------------------------------------------------------------------------------ pass B
// grid = (tiles of 512 individuals, bins).  Positions = the bin's SNP rows, padded to a multiple of 128
// with zero-weight rows.  Each stage gathers 128 rows x 128 packed bytes (a full DRAM line per row), decodes
// them into four MN-major A tiles (128 individuals each) that share one Uq tile, and accumulates four
// 128 x NC int32 tiles in TMEM.
__global__ void __launch_bounds__(TC_THREADS, 2)
k_tc_pass_b(const __grid_constant__ CUtensorMap tm_uq, const uint8_t* __restrict__ bed, int pitch, int Np,
            const int32_t* __restrict__ pos_rows, const uint8_t* __restrict__ fill, int B, int Bp, int L, int NC, int F,
            const unsigned int* __restrict__ wmax, const int32_t* __restrict__ pstart, const int32_t* __restrict__ bin_off,
            const double* __restrict__ cs, const float* __restrict__ rowscale, float* __restrict__ P_out,
            float* __restrict__ S_accum, uint32_t tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tileA = smem;
  uint8_t* tileB = tileA + TC_AS * TC_TILE_A;
  const int tileB_bytes = NC * 128;
  uint8_t* packed = tileB + TC_BS * tileB_bytes;
  TcSmem* sm = reinterpret_cast<TcSmem*>(packed + TC_PK * TC_PACKED);
  const uint32_t tileA_s = smem_u32(tileA), packed_s = smem_u32(packed);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.y;
  const int i0 = blockIdx.x * 512;
  const int p0 = pstart[k];
  const int n_st = (pstart[k + 1] - p0) >> 7;
  const int n_real = bin_off[k + 1] - bin_off[k];

  if (n_st == 0) {   // empty bin in this block: X_k = 0, so P = 0 (S unchanged)
    if (P_out)
      for (int idx = threadIdx.x; idx < B * 512; idx += TC_THREADS) {
        const int b = idx >> 9, i = i0 + (idx & 511);
        if (i < Np) P_out[((size_t)k * B + b) * Np + i] = 0.f;
      }
    return;
  }

  tc_setup(sm, warp, tmem_cols);
  const uint32_t tmem = sm->tmem_base;

  if (warp < TC_DECODE_WARPS) {
    const int t = threadIdx.x & 127, g = warp >> 2;
    const uint8_t* base = bed + (i0 >> 2);
    const int32_t* rows = pos_rows + p0 + t;
#pragma unroll
    for (int pre = 0; pre < TC_PK - 1; ++pre) {
      if (pre < n_st) {
        const int row = rows[pre * 128];
        if (row >= 0) tc_issue_row(packed_s + pre * TC_PACKED, t, g, base + (size_t)row * pitch);
      }
      cp_async_commit();
    }
    for (int st = 0; st < n_st; ++st) {
      const int nxt = st + TC_PK - 1;
      if (nxt < n_st) {
        const int row = rows[nxt * 128];
        if (row >= 0) tc_issue_row(packed_s + (nxt % TC_PK) * TC_PACKED, t, g, base + (size_t)row * pitch);
      }
      cp_async_commit();
      cp_async_wait<TC_PK - 1>();
      const int row = rows[st * 128];
      const uint32_t tab = 0x02010000u | (row >= 0 ? (uint32_t)fill[row] << 8 : 0u);
      const uint32_t slot = packed_s + (st % TC_PK) * TC_PACKED + t * 16;
#pragma unroll
      for (int q0 = 0; q0 < 4; q0 += TC_G) {
        const int q = q0 + g;
        const int sub = st * 4 + q, a = sub % TC_AS, use = sub / TC_AS;
        const uint4 lo = lds128(slot + (2 * q) * 2048);
        const uint4 hi = lds128(slot + (2 * q + 1) * 2048);
        mbar_wait(&sm->empty_a[a], (use & 1) ^ 1);
        tc_store_row(tileA_s + a * TC_TILE_A, t, lo, hi, tab);
        fence_proxy_async();
        mbar_arrive(&sm->full_a[a]);
      }
    }
    // ---- epilogue: TMEM lane = position inside the 128-individual tile q
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int q = g; q < 4; q += TC_G) {
      const int i = i0 + q * 128 + (t & ~15) + tc_perm16(t & 15);
      const bool in_range = i < Np;
      const double rs = in_range ? (double)rowscale[i] : 0.0;
      for (int c0 = 0; c0 < Bp; c0 += 4) {
        double val[4];
        tmem_combine4(trow + (uint32_t)(q * NC + c0), L, Bp, val);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int b = c0 + j;
          if (b < B && in_range) {
            // power-of-two dequantisation factor 2^(e - F), 2^e > max |w| (same rule as k_tc_quant_w)
            const int e = (int)((wmax[b] >> 23) & 255u) - 126;
            const float xf = (float)(rs * (ldexp(val[j], e - F) - cs[(size_t)k * B + b]));
            const size_t o = ((size_t)k * B + b) * Np + i;
            if (P_out) P_out[o] = xf;
            if (S_accum) atomicAdd(S_accum + o, xf);   // result unused -> RED: no load round trip
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == TC_DECODE_WARPS) {
    if (lane == 0) {
      for (int st = 0; st < n_st; ++st) {
        const int b = st % TC_BS, use = st / TC_BS;
        mbar_wait(&sm->empty_b[b], (use & 1) ^ 1);
        mbar_expect_tx(&sm->full_b[b], (uint32_t)tileB_bytes);
        tma_load_2d(tileB + b * tileB_bytes, &tm_uq, &sm->full_b[b], p0 + st * 128, 0);
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = idesc_i8(128, NC, 1);   // A is MN-major: 128 individuals contiguous per SNP row
      for (int st = 0; st < n_st; ++st) {
        const int b = st % TC_BS, useb = st / TC_BS;
        mbar_wait(&sm->full_b[b], useb & 1);
        const uint32_t b0 = smem_u32(tileB + b * tileB_bytes);
        const int ksteps = min(4, (n_real - st * 128 + 31) >> 5);   // all-padding K-steps are skipped
        for (int q = 0; q < 4; ++q) {
          const int sub = st * 4 + q, a = sub % TC_AS, use = sub / TC_AS;
          mbar_wait(&sm->full_a[a], use & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(tileA + a * TC_TILE_A);
          for (int j = 0; j < ksteps; ++j)   // K = 32 SNP rows per instruction: 32 rows x 128 B further down the tile
            umma_i8(tmem + q * NC, smem_desc_sw128(a0 + j * 4096, TC_TILE_A, 1024), smem_desc_sw128(b0 + j * 32, 16, 1024),
                    idesc, (uint32_t)((st | j) != 0));
          umma_commit(&sm->empty_a[a]);
        }
        umma_commit(&sm->empty_b[b]);
      }
      umma_commit(&sm->acc_full);
    }
  }
  __syncthreads();
  if (warp == TC_DECODE_WARPS + 1) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// ------------------------------------------------------------------------------------------ quantisation kernels
// Balanced base-256 digits of a signed integer: q = sum_l d_l 256^l, d_l in [-128, 127].
__device__ __forceinline__ void tc_limbs(long long q, int L, int8_t* out, size_t stride) {
  for (int l = 0; l < L; ++l) {
    int d = (int)(((q + 128) & 255) - 128);
    out[(size_t)l * stride] = (int8_t)d;
    q = (q - d) >> 8;
  }
}

// One block per RHS column: max |R|, power-of-two scale, int8 limbs in the permuted individual order.
__global__ void __launch_bounds__(256)
k_tc_quant_rhs(const float* __restrict__ rhs, int Np, int R1p, int L, int F, int8_t* __restrict__ rq,
               double* __restrict__ col_dq) {
  const int c = blockIdx.x;
  const float* col = rhs + (size_t)c * Np;
  __shared__ float red[8];
  __shared__ int s_e;
  float mx = 0.f;
  for (int i = threadIdx.x; i < Np; i += 256) mx = fmaxf(mx, fabsf(col[i]));
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    s_e = mx > 0.f ? (int)((__float_as_uint(mx) >> 23) & 255u) - 126 : 0;   // 2^e > max
    col_dq[c] = ldexp(1.0, s_e - F);
  }
  __syncthreads();
  const int e = s_e;
  for (int i = threadIdx.x; i < Np; i += 256) {
    long long q = llrint(ldexp((double)col[i], F - e));
    const int pos = (i & ~15) | tc_invperm16(i & 15);
    tc_limbs(q, L, rq + (size_t)c * Np + pos, (size_t)R1p * Np);
  }
}

__global__ void k_tc_positions(const int32_t* __restrict__ bin_rows, const int32_t* __restrict__ bin_off,
                               const int32_t* __restrict__ pstart, int32_t* __restrict__ pos_rows) {
  const int k = blockIdx.y;
  const int n = bin_off[k + 1] - bin_off[k], p0 = pstart[k], span = pstart[k + 1] - p0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < span; i += gridDim.x * blockDim.x)
    pos_rows[p0 + i] = i < n ? bin_rows[bin_off[k] + i] : -1;
}

__global__ void k_tc_wmax(const float* __restrict__ w1, int m, int B, unsigned int* __restrict__ wmax) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * B) return;
  atomicMax(wmax + (idx % B), __float_as_uint(fabsf(w1[idx])));
}

__global__ void k_tc_quant_w(const float* __restrict__ w1, const int32_t* __restrict__ pos_rows, int n_pos, int cap_pos,
                             int B, int Bp, int L, int F, const unsigned int* __restrict__ wmax, int8_t* __restrict__ uq) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pos * B) return;
  const int p = idx % n_pos, b = idx / n_pos;
  const int row = pos_rows[p];
  const int e = (int)((wmax[b] >> 23) & 255u) - 126;
  const long long q = row >= 0 ? llrint(ldexp((double)w1[(size_t)row * B + b], F - e)) : 0ll;
  tc_limbs(q, L, uq + (size_t)b * cap_pos + p, (size_t)Bp * cap_pos);
}

// ------------------------------------------------------------------------------------------ host side
static int tc_encode_2d(TcState* s, CUtensorMap* map, void* base, uint64_t inner, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {inner};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = s->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rhe_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RHE_ERR_CUDA; }
  return RHE_OK;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int tc_smem_bytes(int n_cols) {
  return TC_AS * TC_TILE_A + TC_BS * n_cols * 128 + TC_PK * TC_PACKED + (int)sizeof(TcSmem) + 1024;
}
static inline uint32_t pow2_cols(int n) { uint32_t c = 32; while ((int)c < n) c <<= 1; return c; }

int rhe_tc_create(rhe_ctx* c) {
  const rhe_config& g = c->cfg;
  if (g.n_ops != 1 || g.n_sets != 1) { rhe_set_error("RHE_PATH_TCGEN05 currently covers the RHE model (one operand, one RHS set)"); return RHE_ERR_UNSUPPORTED; }
  TcState* s = new TcState();
  const char* envL = getenv("PYRHE_B200_LIMBS");
  s->L = envL ? atoi(envL) : 3;
  if (s->L < 2 || s->L > 4) { delete s; rhe_set_error("PYRHE_B200_LIMBS must be 2..4"); return RHE_ERR_INVALID; }
  s->F = 8 * s->L - 2;
  s->R1p = round_up(c->R1, 4);
  s->NBa = round_up(s->L * s->R1p, 16);
  s->Bp = round_up(g.n_vec, 4);
  s->NCb = round_up(s->L * s->Bp, 16);
  if (s->NBa > 256 || 4 * s->NCb > 512 || g.n_bins > 64) {
    delete s;
    rhe_set_error("RHE_PATH_TCGEN05: %d RHS columns / %d bins x %d vectors exceed one TMEM allocation", c->R1, g.n_bins, g.n_vec);
    return RHE_ERR_UNSUPPORTED;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    delete s;
    rhe_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RHE_ERR_CUDA;
  }
  s->encode = (PFN_encodeTiled)fn;
  c->tc = s;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes); } };
  alloc((void**)&s->rq, (size_t)s->NBa * c->Np);
  alloc((void**)&s->col_dq, sizeof(double) * c->R1);
  alloc((void**)&s->pstart, sizeof(int32_t) * (g.n_bins + 1));
  alloc((void**)&s->wmax, sizeof(unsigned int) * g.n_vec);
  if (e != cudaSuccess) { rhe_set_error("tensor-core workspace allocation failed: %s", cudaGetErrorString(e)); return RHE_ERR_CUDA; }
  int rc = tc_encode_2d(s, &s->tm_rq, s->rq, (uint64_t)c->Np, (uint64_t)s->NBa, (uint32_t)s->NBa);
  if (rc) return rc;
  const int smem_a = tc_smem_bytes(s->NBa), smem_b = tc_smem_bytes(s->NCb);
  RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a));
  RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b));
  return RHE_OK;
}

void rhe_tc_destroy(rhe_ctx* c) {
  TcState* s = (TcState*)c->tc;
  if (!s) return;
  void* ptrs[] = {s->rq, s->col_dq, s->pos_rows, s->pstart, s->uq, s->wmax};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete s;
  c->tc = nullptr;
}

int rhe_tc_set_rhs(rhe_ctx* c, cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  k_tc_quant_rhs<<<c->R1, 256, 0, st>>>(c->rhs, c->Np, s->R1p, s->L, s->F, s->rq, s->col_dq);
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

int rhe_tc_pass_a(rhe_ctx* c, const uint8_t* bed, int m, cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  const int tiles = rhe_div_up(m, 128);
  // whole waves of 2 resident CTAs per SM (a nearly empty last wave costs a full CTA time); at least
  // 8 super-stages (4096 individuals) per CTA so that prologue / epilogue stay amortised
  const int slots = 148 * 2;
  int splits = tiles >= slots ? 1 : slots / tiles;
  if (tiles * splits < slots / 2 + slots / 4 && tiles < slots) splits = (2 * slots) / tiles;   // poor fill: use two waves
  int chunk = round_up(rhe_div_up(c->Np, splits), 512);
  if (chunk < 4096) chunk = 4096;
  if (chunk > c->Np) chunk = c->Np;
  splits = rhe_div_up(c->Np, chunk);
  k_tc_pass_a<<<dim3(tiles, splits), TC_THREADS, tc_smem_bytes(s->NBa), st>>>(
      s->tm_rq, bed, c->cfg.pitch_bytes, m, c->Np, s->NBa, c->R1, s->R1p, s->L, c->fill, s->col_dq, c->t_raw, chunk,
      pow2_cols(s->NBa));
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

int rhe_tc_pass_b(rhe_ctx* c, const uint8_t* bed, int m, const int32_t* bin_rows, const int32_t* bin_off,
                  const int32_t* bin_off_host, float* P_out, float* S_accum, cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  const rhe_config& g = c->cfg;
  const int K = g.n_bins, B = g.n_vec;
  // bin-sorted positions: every bin padded to a multiple of 128 rows (one stage never mixes bins)
  int32_t pstart[65];
  pstart[0] = 0;
  for (int k = 0; k < K; ++k) pstart[k + 1] = pstart[k] + round_up(bin_off_host[k + 1] - bin_off_host[k], 128);
  const int n_pos = round_up(pstart[K] > 0 ? pstart[K] : 1, 128);
  if (n_pos > s->cap_pos) {
    RHE_CUDA(cudaStreamSynchronize(st));
    if (s->pos_rows) cudaFree(s->pos_rows);
    if (s->uq) cudaFree(s->uq);
    s->cap_pos = round_up(n_pos + n_pos / 8, 128);
    RHE_CUDA(cudaMalloc((void**)&s->pos_rows, sizeof(int32_t) * s->cap_pos));
    RHE_CUDA(cudaMalloc((void**)&s->uq, (size_t)s->NCb * s->cap_pos));
    RHE_CUDA(cudaMemset(s->uq, 0, (size_t)s->NCb * s->cap_pos));
    int rc = tc_encode_2d(s, &s->tm_uq, s->uq, (uint64_t)s->cap_pos, (uint64_t)s->NCb, (uint32_t)s->NCb);
    if (rc) return rc;
  }
  RHE_CUDA(cudaMemcpyAsync(s->pstart, pstart, sizeof(int32_t) * (K + 1), cudaMemcpyHostToDevice, st));
  RHE_CUDA(cudaMemsetAsync(s->pos_rows, 0xFF, sizeof(int32_t) * n_pos, st));
  RHE_CUDA(cudaMemsetAsync(s->wmax, 0, sizeof(unsigned int) * B, st));
  k_tc_positions<<<dim3(rhe_div_up(m, 256), K), 256, 0, st>>>(bin_rows, bin_off, s->pstart, s->pos_rows);
  RHE_LAUNCH_CHECK(c);
  k_tc_wmax<<<rhe_div_up(m * B, 256), 256, 0, st>>>(c->w1, m, B, s->wmax);
  RHE_LAUNCH_CHECK(c);
  k_tc_quant_w<<<rhe_div_up((int64_t)n_pos * B, 256), 256, 0, st>>>(c->w1, s->pos_rows, n_pos, s->cap_pos, B, s->Bp, s->L, s->F,
                                                                    s->wmax, s->uq);
  RHE_LAUNCH_CHECK(c);
  k_tc_pass_b<<<dim3(rhe_div_up(c->Np, 512), K), TC_THREADS, tc_smem_bytes(s->NCb), st>>>(
      s->tm_uq, bed, g.pitch_bytes, c->Np, s->pos_rows, c->fill, B, s->Bp, s->L, s->NCb, s->F, s->wmax, s->pstart, bin_off,
      c->cs, c->rowscale, P_out, S_accum, pow2_cols(4 * s->NCb));
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}
