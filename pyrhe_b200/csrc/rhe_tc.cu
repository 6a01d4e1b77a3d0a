// Tensor-core path (sm_100a): fused 2-bit decode + int8 tcgen05.mma with int32 accumulators in TMEM.
//
// Both products of the block path are skinny GEMMs whose big operand is the genotype matrix:
//   pass A   T[s, n] = sum_i G[i, s] * Rq[i, n]      M = 128 SNPs,        K = individuals, N = limb columns
//   pass B   D[i, n] = sum_s G[i, s] * Uq[s, n]      M = 128 individuals, K = SNPs,        N = limb columns
// G is exact in int8 ({0,1,2}); the fp32 right-hand sides are fixed-point numbers split into L signed
// 8-bit limbs stacked along N, so the int32 accumulation is EXACT and the result is independent of how
// the work is split.
//
// Warps are specialised (decode groups, TMA producers, one MMA-issue warp per decode group, drain warps); every role
// loop is warp-uniform with elect.sync around the issue.
//   pass A (k_tc_pass_a): packed super-stages (128 SNP rows x 128 B) arrive as 2-D TMA boxes; a decode thread owns one
//           SNP row, a 32-bit packed word expands to sixteen int8 values with 2 logic ops, 4 byte-permutes and 3
//           multiply-high shifts; the expanded A operand goes straight from registers into TENSOR MEMORY (tcgen05.st)
//           and the MMA reads A from TMEM (K-major), so the big operand never touches shared memory as bytes; only the
//           small Rq tiles (TMA, 128B swizzle) do.
//   pass B needs the same bytes as an MN-major operand (128 individuals contiguous per SNP row), which TMEM cannot
//           provide.  Two kernels:
//           k_tc_pass_b2 -- the block also has an INDIVIDUAL-MAJOR copy (rhe_block_transpose, written at ingest: imputed
//           A2 counts, 2 bits each, contiguous 16 KB boxes): then pass B is pass A with the roles swapped -- TMEM lane =
//           individual, K = bin-sorted positions, the A operand from tensor memory, a mask-and-shift decode -- in a
//           persistent CTA per SM whose drain warps write a bin's result while the next bins accumulate.
//           k_tc_pass_b  -- no copy: the tile is decoded into shared memory (128B swizzle) and read by the MMA through
//           an MN-major descriptor; rows are gathered by bin through per-warp cp.async rings; one CTA owns MT x 128
//           individuals and the bins of one bin group (bin k accumulates in its own TMEM columns).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rhe_common.cuh"

#define TC_TILE_A 16384          // 128 rows x 128 bytes
// Right shifts of the decode: plain SHF on the ALU pipe (1), or multiply-high on the FMA pipe (0).  Measured on config-5
// blocks: IMAD.HI costs more issue time than it saves the ALU pipe wherever the expansion is mask-and-shift (pass B from
// tensor memory 0.379 -> 0.340 ms, pass A on re-tiled rows 0.286 -> 0.272 ms) and in pass A's table look-up (0.295 -> 0.289 ms);
// only the gather kernel, whose ALU pipe also writes the tile to shared memory, keeps the multiply-high (0.543 vs 0.554 ms).
#ifndef TC_SHIFT_VARIANT
#define TC_SHIFT_VARIANT 1
#endif

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Annotation-derived metadata of one (jackknife block, bin group): depends only on the bin row lists, so it is built
// by rhe_block_plan_create and owned by the plan.
struct TcBlockMeta {
  int k0 = 0, kn = 0;             // the bins [k0, k0 + kn) of this bin group
  int n_pos = 0;                  // bin-sorted positions, every bin padded to a multiple of 128 rows
  int32_t* pos_rows = nullptr;    // [n_pos]       block-local SNP row (-1 = padding)
  int32_t* stage_info = nullptr;  // [n_pos / 128] bin | first-stage-of-bin << 8 | K-steps << 16
  int32_t* bin_count = nullptr;   // [kn]          rows per bin of the group
};

struct TcPlan {
  std::vector<TcBlockMeta> groups;   // one per bin group of at most KG bins
};

struct TcState {
  int L = 3;            // limbs per fixed-point value (PYRHE_B200_LIMBS)
  int F = 22;           // fixed-point magnitude bits: |q| <= 2^F, F = 8 L - 2
  int Rc = 0;           // RHS columns per pass-A launch (all R1 of them unless their limb columns exceed the MMA / shared-memory limits)
  int R1p = 0;          // Rc rounded up to 4
  int NBa = 0;          // pass A MMA N  = round16(L * R1p)
  int Bc = 0;           // pass-B vector columns per launch (all of them unless groups x vectors exceeds 64: then in chunks)
  int Bp = 0;           // Bc rounded up to 2
  int NCb = 0;          // pass B MMA N per (bin, M-tile) = round16(L * Bp)
  int MT = 2;           // 128-individual M-tiles per pass-B CTA
  int G = 4;            // pass-B decode groups per CTA: 4 (one CTA per SM) or 2 (two co-resident CTAs per SM)
  int KG = 0;           // bins per pass-B launch: the accumulators of KG bins x MT tiles fill the 512 TMEM columns
  int8_t* rq = nullptr;       // [NBa][Np]   limb l of column c at row l * R1p + c, permuted individual order
  double* col_dq = nullptr;   // [R1]        power-of-two dequantisation factor of every RHS column
  int8_t* uq = nullptr;         // [NCb][cap_pos] quantised pass-B weights of the current block
  int32_t* pos_meta = nullptr;  // [SI][chunks][128][4] per-group decode metadata (SNP row | fill << 24 | mode << 26) of the current block
  unsigned int* wmax = nullptr; // [n_groups][B] max |weight| per weight group and column (float bits)
  int cap_pos = 0;
  int n_sm = 148;               // SMs of the device (grid of the persistent kernel)
  int n_ops = 1;                // genotype operands (2: RHE-DOM)
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
#pragma unroll 1
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
// The same primitives on raw shared-space addresses: the single-thread producer / issuer loops are pure latency
// chains, so they keep every address in a register and advance it incrementally (no cvta, no div / mod).
__device__ __forceinline__ void mbar_arrive_s(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t addr, uint32_t parity) {
#pragma unroll 1
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
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
// One elected lane of a converged warp (the role loops run warp-uniformly so that descriptors and barrier
// addresses live in uniform registers; only the issue itself is predicated).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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

__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_commit_s(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
// Loop-invariant form: the start address lives in the low 14 bits of the low word, so advancing the operand by
// `bytes` is one 32-bit add on a precomputed descriptor (no carry: all shared addresses are < 256 KB).
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

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

// Per-SNP value table, byte c = operand value of 2-bit code c.  mode 0: A2 count {0, fill, 1, 2};
// mode 1: the dominance indicator [count == 2] = {0, fill == 2, 0, 1} (rhe_dom.py:36-39 after h = mu g - 2 [g == 2]).
__device__ __forceinline__ uint32_t tc_value_table(uint32_t fill, int mode) {
  return mode ? (0x01000000u | ((fill == 2u ? 1u : 0u) << 8)) : (0x02010000u | (fill << 8));
}

// Expand one packed word (16 genotypes) into 16 int8 values with the per-SNP value table
// tab = {0, fill, 1, 2} (byte c = value of code c), in the permuted order above.
// PRMT reads only the low four nibbles of its selector, and the ALU pipe (LOP3 / SHF / PRMT) is what limits the
// decode, so: the two masks are computed once per word (inline PTX keeps the compiler from re-deriving a mask per
// PRMT) and the right shifts run as multiply-high on the FMA pipe (x >> s == umulhi(x, 2^(32 - s))).
__device__ __forceinline__ uint32_t tc_prmt(uint32_t tab, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(tab), "r"(0u), "r"(sel));
  return r;
}
__device__ __forceinline__ uint32_t tc_shr_fma(uint32_t x, uint32_t pow2) {   // pow2 = 2^(32 - shift)
  uint32_t r;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(pow2));
  return r;
}
// x << s as a multiply: IMAD.SHL on the FMA pipe (the ALU pipe carries the masks and the right shifts of the decode)
__device__ __forceinline__ uint32_t tc_shl_fma(uint32_t x, uint32_t pow2) {
  uint32_t r;
  asm("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(pow2));
  return r;
}
// TC_SHIFT_VARIANT 2: the mask-and-shift expansions deliver 4 x value (the two bits of a field land on bits 2-3 of their
// byte), which turns one of the three right shifts per word into a left shift on the other pipe; the epilogues divide by 4.
// Measured: pass A on re-tiled rows 0.272 -> 0.261 ms, pass B from tensor memory 0.343 -> 0.349 ms -- so pass A takes it
// (TC_PA_SHIFT_VARIANT) and pass B does not.  With values up to 8 the int32 accumulation of pass A is exact for
// 8 x 128 x N < 2^31, i.e. N < 2^21 individuals (rhe_block_tiled_bytes answers 0 beyond that).
#ifndef TC_PA_SHIFT_VARIANT
#define TC_PA_SHIFT_VARIANT 2
#endif
#define TC_VALUE_SCALE (TC_SHIFT_VARIANT == 2 ? 0.25 : 1.0)
#define TC_PA_VALUE_SCALE (TC_PA_SHIFT_VARIANT == 2 ? 0.25 : 1.0)
template <int MULHI = 1>
__device__ __forceinline__ uint4 tc_expand(uint32_t w, uint32_t tab) {
  const uint32_t e = w & 0x33333333u;
  uint4 r;
  if constexpr (MULHI) {
    const uint32_t o = tc_shr_fma(w, 1u << 30) & 0x33333333u;
    r.x = tc_prmt(tab, e);
    r.y = tc_prmt(tab, tc_shr_fma(e, 1u << 16));
    r.z = tc_prmt(tab, o);
    r.w = tc_prmt(tab, tc_shr_fma(o, 1u << 16));
  } else {
    const uint32_t o = (w >> 2) & 0x33333333u;
    r.x = tc_prmt(tab, e);
    r.y = tc_prmt(tab, e >> 16);
    r.z = tc_prmt(tab, o);
    r.w = tc_prmt(tab, o >> 16);
  }
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 32 registers (one 128-byte operand row) -> 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint4 (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0].x), "r"(r[0].y), "r"(r[0].z), "r"(r[0].w), "r"(r[1].x), "r"(r[1].y), "r"(r[1].z), "r"(r[1].w),
        "r"(r[2].x), "r"(r[2].y), "r"(r[2].z), "r"(r[2].w), "r"(r[3].x), "r"(r[3].y), "r"(r[3].z), "r"(r[3].w),
        "r"(r[4].x), "r"(r[4].y), "r"(r[4].z), "r"(r[4].w), "r"(r[5].x), "r"(r[5].y), "r"(r[5].z), "r"(r[5].w),
        "r"(r[6].x), "r"(r[6].y), "r"(r[6].z), "r"(r[6].w), "r"(r[7].x), "r"(r[7].y), "r"(r[7].z), "r"(r[7].w)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 16 registers -> 16 consecutive TMEM columns of this thread's lane; completion is awaited by tmem_st_wait()
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint4 (&r)[4]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0].x), "r"(r[0].y), "r"(r[0].z), "r"(r[0].w), "r"(r[1].x), "r"(r[1].y), "r"(r[1].z), "r"(r[1].w),
        "r"(r[2].x), "r"(r[2].y), "r"(r[2].z), "r"(r[2].w), "r"(r[3].x), "r"(r[3].y), "r"(r[3].z), "r"(r[3].w)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Zero 32 consecutive TMEM columns of this thread's lane (accumulators start at 0 so that every MMA can
// accumulate and the issuing threads need no ordering among themselves).
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  uint4 z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) z[i] = make_uint4(0u, 0u, 0u, 0u);
  tmem_st32(taddr, z);
}

// D[tmem] (+)= A[tmem] * B[smem], int8 x int8 -> int32 (A operand from tensor memory, K-major)
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  uint32_t zero = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero) : "memory");
}

// Optional cycle accounting (-DRHE_TC_PROF): where each role spends its time.  Every thread accumulates in
// registers and lane 0 of each warp adds its totals once at the end, so the timed code is barely disturbed.
#ifdef RHE_TC_PROF
__device__ unsigned long long g_prof[32];
#define PROF_T0() long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long _t0 = clock64()
#define PROF_ADD(slot) do { long long _t1 = clock64(); _acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#define PROF_FLUSH(base) do { if ((threadIdx.x & 31) == 0) { _Pragma("unroll") for (int _i = 0; _i < 8; ++_i) if (_acc[_i]) atomicAdd(&g_prof[(base) + _i], (unsigned long long)_acc[_i]); } } while (0)
extern "C" void rhe_tc_prof_dump() {
  unsigned long long h[32];
  cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
  const char* names[] = {"B.dec.top(meta,tab)", "B.dec.empty_wait", "B.dec.store", "B.dec.fetch", "B.dec.fence+arrive", "B.dec.acc_wait", "B.dec.epilogue", "B.dec.-",
                         "B.mma.info", "B.mma.full_b", "B.mma.full_a", "B.mma.issue+commit", "B.mma.-", "B.mma.-", "B.mma.-", "B.mma.-",
                         "A.dec.lds", "A.dec.empty_wait", "A.dec.expand+st", "A.dec.st_wait+arrive", "A.dec.cpwait", "A.dec.acc_wait", "A.dec.epilogue", "A.dec.-",
                         "A.mma.full_b", "A.mma.full_a", "A.mma.issue+commit", "A.mma.-", "A.mma.-", "A.mma.-", "A.mma.-", "A.mma.-"};
  for (int i = 0; i < 32; ++i) if (h[i]) printf("  %-22s %12.3f Mcyc (sum over warps)\n", names[i], h[i] / 1e6);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(g_prof, z, sizeof(z));
}
#else
#define PROF_T0()
#define PROF_ADD(slot)
#define PROF_FLUSH(base)
#endif

// ------------------------------------------------------------------------------------------ pass A
// grid = (splits over individuals, SNP tiles of 128).  t_raw[s][c] += dq[c] * sum_i g_is * q_ic  (exact).
// The CTA owns the super-stages ss = y + splits * k (512 individuals = one 128-byte line of each of its 128 SNP
// rows).  The producer warp brings every super-stage in with ONE 2-D TMA box (128 rows x 128 B, 128-byte swizzle)
// and the four Rq tiles that go with it; no decode thread issues a global load.  The 4 n_ss sub-tiles (128
// individuals) are taken round-robin by PA_G groups of four decode warps (group g: j = g, g + PA_G, ...), each
// expanding its sub-tile into one of the group's two TMEM A slots.  512 threads x 64 registers x 2 CTAs per SM.
#define PA_G 3
#define PA_DW (4 * PA_G)
#define PA_THREADS (32 * (PA_DW + 2 + PA_G))   // decode warps, two TMA warps (Rq tiles, genotype boxes), one MMA-issue warp per group
#define PA_AS (2 * PA_G)          // TMEM A slots (32 columns each): two per group
#define PA_RS 2                   // smem ring of Rq super-stages (four NB x 128 B tiles each, one barrier pair per slot)
#define PA_MAXGS 8                // deepest smem ring of packed super-stages (128 rows x 128 B each); the launch picks the depth
#define PA_PACKED (128 * 128)

struct PaSmem {
  uint64_t full_a[PA_AS], empty_a[PA_AS], full_b[PA_RS], empty_b[PA_RS], full_g[PA_MAXGS], empty_g[PA_MAXGS], acc_full;
  uint32_t tmem_base;
};

// TILED = 1: the block's rows were re-tiled at ingest (rhe_block_retile): every (128 SNP rows x 128 B) box is one
// contiguous 16 KB piece of memory -- [row tile][512-individual column][128 rows][128 B] -- and the two bits of a genotype
// hold the imputed A2 COUNT, so a box streams from HBM like a plain copy (the row-strided boxes of the PLINK layout top
// out at 4.2-4.5 TB/s on this part, tools/membench.cu) and a word expands with masks and shifts only (no per-SNP table).
template <int PA_GS, int TILED>
__global__ void __launch_bounds__(PA_THREADS, 2)
k_tc_pass_a(const __grid_constant__ CUtensorMap tm_rq, const __grid_constant__ CUtensorMap tm_bed, int m, int Np,
            int NB, int R1, int R1p, int L, int c_lo, int Rv, int rq_row0, const uint8_t* __restrict__ fill,
            const double* __restrict__ col_dq, double* __restrict__ t_raw, uint32_t tmem_cols, uint32_t col_a, int mode, int dbg) {
  // the launch's RHS columns [c_lo, c_lo + Rv) of R1 sit in the Rq rows [rq_row0, rq_row0 + NB)
  // the 128-byte swizzle needs a 1 KB aligned base: declared, not padded for (the pad would cost the two co-resident
  // CTAs their fourth ring slot); a misplaced window traps instead of corrupting the tiles
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) asm volatile("trap;");
  uint8_t* packed = smem;                            // [PA_GS][128 rows][128 B], 128-byte swizzle
  uint8_t* tileB = packed + PA_GS * PA_PACKED;
  const int tileB_bytes = NB * 128;
  PaSmem* sm = reinterpret_cast<PaSmem*>(tileB + PA_RS * 4 * tileB_bytes);
  const uint32_t packed_s = smem_u32(packed);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  // split index fastest: the CTAs of one SNP tile are co-resident and read adjacent lines of the same 128 rows
  const int snp0 = blockIdx.y * 128;
  const int splits = gridDim.x, y = blockIdx.x;
  const int total_ss = Np >> 9;
  const int n_ss = total_ss > y ? (total_ss - y + splits - 1) / splits : 0;   // super-stages of this CTA
  if (n_ss <= 0) return;
  const int n_sub = 4 * n_ss;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PA_AS; ++s) { mbar_init(&sm->full_a[s], 4); mbar_init(&sm->empty_a[s], 1); }   // one arrival per decode warp
    for (int s = 0; s < PA_RS; ++s) { mbar_init(&sm->full_b[s], 1); mbar_init(&sm->empty_b[s], 4); }           // 4 sub-tiles
    for (int s = 0; s < PA_GS; ++s) { mbar_init(&sm->full_g[s], 1); mbar_init(&sm->empty_g[s], 16); } // 4 sub-tiles x 4 warps
    mbar_init(&sm->acc_full, PA_G);
    fence_barrier_init();
  }
  if (warp == PA_DW + 2) tmem_alloc(&sm->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  if (warp < 4)                                        // zero the accumulator columns (lane quadrant per warp)
    for (uint32_t c = 0; c < col_a; c += 32) tmem_zero32(tmem + ((uint32_t)(warp * 32) << 16) + c);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < PA_DW) {
    const int t = (warp & 3) * 32 + lane, g = warp >> 2;
    const int s = min(snp0 + t, m - 1);
    const uint32_t tab = TILED ? 0u : tc_value_table(fill[s], mode);
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t row_s = packed_s + (uint32_t)t * 128;      // this thread's SNP row inside a staged box
    // value-coded rows: register k of a word = fields k, k + 4, k + 8, k + 12 (rhe_block_retile places the individual
    // that the table look-up would deliver at byte 4 k + j into field 4 j + k); [g == 2] is bit 1 of the value
    auto expand = [&](uint32_t w) {
      if constexpr (TILED) {
        uint4 r;
        if (mode == 0) {
#if TC_PA_SHIFT_VARIANT == 2
          const uint32_t mk = 0x0C0C0C0Cu;
          r.x = tc_shl_fma(w, 4u) & mk; r.y = w & mk; r.z = (w >> 2) & mk; r.w = (w >> 4) & mk;
#elif TC_PA_SHIFT_VARIANT == 1
          const uint32_t mk = 0x03030303u;
          r.x = w & mk; r.y = (w >> 2) & mk; r.z = (w >> 4) & mk; r.w = (w >> 6) & mk;
#else
          const uint32_t mk = 0x03030303u;
          r.x = w & mk; r.y = tc_shr_fma(w, 1u << 30) & mk; r.z = tc_shr_fma(w, 1u << 28) & mk; r.w = tc_shr_fma(w, 1u << 26) & mk;
#endif
        } else {
#if TC_PA_SHIFT_VARIANT == 2
          const uint32_t mk = 0x04040404u;
          r.x = tc_shl_fma(w, 2u) & mk; r.y = (w >> 1) & mk; r.z = (w >> 3) & mk; r.w = (w >> 5) & mk;
#elif TC_PA_SHIFT_VARIANT == 1
          const uint32_t mk = 0x01010101u;
          r.x = (w >> 1) & mk; r.y = (w >> 3) & mk; r.z = (w >> 5) & mk; r.w = (w >> 7) & mk;
#else
          const uint32_t mk = 0x01010101u;
          r.x = tc_shr_fma(w, 1u << 31) & mk; r.y = tc_shr_fma(w, 1u << 29) & mk;
          r.z = tc_shr_fma(w, 1u << 27) & mk; r.w = tc_shr_fma(w, 1u << 25) & mk;
#endif
        }
        return r;
      } else {
        return tc_expand<1 - TC_SHIFT_VARIANT>(w, tab);
      }
    };
    const uint32_t sw = (uint32_t)(t & 7);                      // 128-byte swizzle: 16-byte chunk c sits at c ^ (row & 7)
    const uint32_t fg = smem_u32(&sm->full_g[0]), eg = smem_u32(&sm->empty_g[0]);
    const uint32_t fa = smem_u32(&sm->full_a[2 * g]), ea = smem_u32(&sm->empty_a[2 * g]);
    const uint32_t dst0 = lane_base + col_a + 32u * (uint32_t)(2 * g);
    const int n_own = n_sub > g ? (n_sub - g + PA_G - 1) / PA_G : 0;     // own sub-tiles j = g + PA_G * jj
    // Sub-tile j = 4 k + q is chunks 2 q, 2 q + 1 of this thread's row in ring slot k % PA_GS.  The group walks
    // j = g, g + PA_G, ...: q advances by PA_G (mod 4) and k by the carry, so slot, parity and chunk offsets advance
    // incrementally (no division in the loop); the same walk, one sub-tile behind, tells which slot to release.
    static_assert(PA_G < 4, "one carry per step");
    uint32_t fq = (uint32_t)g, fsl = 0, fpar = 0;              // next sub-tile to read: chunk pair, ring slot, parity
    uint32_t cq = (uint32_t)g, csl = 0;                        // sub-tile being expanded
    uint32_t aq = 0, apar = 1;                                 // the group's A slot of the next expansion / parity to wait for
    const uint32_t sw4 = sw << 4;
    auto fetch = [&](uint4& lo, uint4& hi) {
      mbar_wait_s(fg + 8u * fsl, fpar);
      // chunk (2 q) ^ sw of the row: ((2 q) ^ sw) << 4 = (sw << 4) ^ (q << 5), and its partner 2 q + 1 is the same
      // address with bit 4 flipped (the row starts on a 128-byte boundary)
      const uint32_t a_lo = row_s + fsl * PA_PACKED + (sw4 ^ (fq << 5));
      lo = lds128(a_lo);
      hi = lds128(a_lo ^ 16u);
      fq += PA_G;
      if (fq >= 4u) { fq -= 4u; if (++fsl == (uint32_t)PA_GS) { fsl = 0; fpar ^= 1u; } }
    };
    // Software pipeline: the packed words of the next sub-tile are read from the ring before the current one is
    // expanded, and each sub-tile goes to TMEM as two 16-column stores so that the first store overlaps the
    // expansion of the second half.  The ring slot is released once the words are in registers (consumed).
    auto put = [&](const uint4& lo, const uint4& hi) {
      mbar_wait_s(ea + 8u * aq, apar);
      tc_fence_after();
      const uint32_t dst = dst0 + 32u * aq;
      if (!(RHE_DBG(8))) {
        uint4 r[4];
        r[0] = expand(lo.x); r[1] = expand(lo.y); r[2] = expand(lo.z); r[3] = expand(lo.w);
        tmem_st16(dst, r);
        uint4 r2[4];
        r2[0] = expand(hi.x); r2[1] = expand(hi.y); r2[2] = expand(hi.z); r2[3] = expand(hi.w);
        tmem_st16(dst + 16, r2);
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();                                    // every lane's stores are complete and fenced: one arrival per warp
      if (lane == 0) { mbar_arrive_s(eg + 8u * csl); mbar_arrive_s(fa + 8u * aq); }
      cq += PA_G;
      if (cq >= 4u) { cq -= 4u; if (++csl == (uint32_t)PA_GS) csl = 0; }
      aq ^= 1u;
      if (aq == 0u) apar ^= 1u;
    };
    // unrolled by two so that the two register sets swap roles without moves
    uint4 lo0, hi0, lo1, hi1;
    if (n_own > 0) fetch(lo0, hi0);
    for (int jj = 0; jj < n_own; jj += 2) {
      if (jj + 1 < n_own) fetch(lo1, hi1);
      put(lo0, hi0);
      if (jj + 1 < n_own) {
        if (jj + 2 < n_own) fetch(lo0, hi0);
        put(lo1, hi1);
      }
    }
    // ---- epilogue: lane quadrant (warp & 3) of TMEM, row = SNP; the groups split the columns
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    const int snp = snp0 + t;
    for (int c0 = 4 * g; c0 < R1p; c0 += 4 * PA_G) {
      double val[4];
      tmem_combine4(lane_base + (uint32_t)c0, L, R1p, val);
      if (snp < m) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < Rv) atomicAdd(t_raw + (size_t)snp * R1 + c_lo + c0 + j, val[j] * ((TILED ? TC_PA_VALUE_SCALE : 1.0) * col_dq[c_lo + c0 + j]));
      }
    }
    tc_fence_before();
  } else if (warp == PA_DW) {
    {                                                  // TMA producer 1: the four Rq tiles of super-stage k -> slot k % PA_RS
      const uint32_t fb = smem_u32(&sm->full_b[0]), eb = smem_u32(&sm->empty_b[0]);
      const uint32_t tileB_s = smem_u32(tileB);
      static_assert(PA_RS == 2, "slot = k & 1, use = k >> 1");
      for (int k = 0; k < n_ss; ++k) {                 // the whole warp runs the loop; one elected lane issues
        const uint32_t sl = (uint32_t)k & 1u;
        mbar_wait_s(eb + 8u * sl, (((uint32_t)k >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          if (RHE_DBG(1)) mbar_arrive_s(fb + 8u * sl);
          else {
            const int x = (y + splits * k) * 512;
            mbar_expect_tx_s(fb + 8u * sl, 4u * (uint32_t)tileB_bytes);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tma_load_2d_s(tileB_s + (sl * 4u + (uint32_t)q) * (uint32_t)tileB_bytes, &tm_rq, fb + 8u * sl, x + q * 128, rq_row0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == PA_DW + 1) {
    {                                                  // TMA producer 2: genotype box of super-stage k -> slot k % PA_GS
      const uint32_t fg = smem_u32(&sm->full_g[0]), eg = smem_u32(&sm->empty_g[0]);
      for (int k = 0; k < n_ss; ++k) {
        const uint32_t sg = (uint32_t)(k % PA_GS);
        mbar_wait_s(eg + 8u * sg, ((uint32_t)(k / PA_GS) & 1u) ^ 1u);
        if (elect_one()) {
          if (RHE_DBG(4)) mbar_arrive_s(fg + 8u * sg);
          else {
            mbar_expect_tx_s(fg + 8u * sg, PA_PACKED);
            if constexpr (TILED)   // box (row tile, column ss) = rows [(tile n_cc + ss) 128, + 128) of a [..][128 B] tensor
              tma_load_2d_s(packed_s + sg * PA_PACKED, &tm_bed, fg + 8u * sg, 0, ((int)blockIdx.y * total_ss + (y + splits * k)) * 128);
            else
              tma_load_2d_s(packed_s + sg * PA_PACKED, &tm_bed, fg + 8u * sg, (y + splits * k) * 128, snp0);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ---- MMA issue: one warp per decode group (one elected lane issues), so no single thread serialises the block.
    // Every MMA accumulates (the accumulator was zeroed), hence the issuers need no mutual ordering.  Group g takes
    // the sub-tiles j = g, g + PA_G, ...: the Rq tile j & 3 of super-stage slot (j >> 2) % PA_RS, TMEM A slot
    // alternating between 2 g and 2 g + 1.
    {
      const int g = warp - (PA_DW + 2);
      const uint32_t idesc = idesc_i8(128, NB, 0);
      const uint32_t fb = smem_u32(&sm->full_b[0]), eb = smem_u32(&sm->empty_b[0]);
      const uint32_t fa = smem_u32(&sm->full_a[2 * g]), ea = smem_u32(&sm->empty_a[2 * g]);
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(tileB), 16, 1024);
      const int n_own = n_sub > g ? (n_sub - g + PA_G - 1) / PA_G : 0;
      PROF_T0();
      uint32_t ph_a = 0, q = 0;
      for (int jj = 0; jj < n_own; ++jj) {
        const uint32_t j = (uint32_t)(g + PA_G * jj), k = j >> 2, sl = k & 1u;     // Rq slot k % PA_RS, tile j & 3
        mbar_wait_s(fb + 8u * sl, (k >> 1) & 1u);
        PROF_ADD(0);
        mbar_wait_s(fa + 8u * q, ph_a);
        tc_fence_after();
        PROF_ADD(1);
        const uint64_t bdesc = bdesc0 + (uint64_t)((sl * 4u + (j & 3u)) * (uint32_t)(tileB_bytes >> 4));
        const uint32_t acol = tmem + col_a + 32u * (2u * (uint32_t)g + q);
        if (elect_one()) {
          if (!(RHE_DBG(2))) {
#pragma unroll
            for (int i = 0; i < 4; ++i)   // K = 32 individuals per instruction: 8 TMEM columns / 32 bytes of the Rq row
              umma_i8_ts(tmem, acol + 8u * i, bdesc + (uint64_t)(i * 2), idesc, 1u);
          }
          if (RHE_DBG(16)) { mbar_arrive_s(ea + 8u * q); mbar_arrive_s(eb + 8u * sl); }
          else { umma_commit_s(ea + 8u * q); umma_commit_s(eb + 8u * sl); }
        }
        __syncwarp();
        PROF_ADD(2);
        q ^= 1u;
        if (q == 0u) ph_a ^= 1u;
      }
      if (elect_one()) umma_commit(&sm->acc_full);
      __syncwarp();
      PROF_FLUSH(24);
    }
  }
  __syncthreads();
  if (warp == PA_DW + 2) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// ------------------------------------------------------------------------------------------ pass B
// grid = tiles of MT x 128 individuals; one CTA runs over ALL bin-sorted positions of the block.
// Four groups of four decode warps: group g = (par, q) expands M-tile q of the stages st = par (mod 4 / MT)
// into shared-memory A slot g.  Stage st belongs to one bin (bins are padded to 128 rows); bin k, M-tile q
// accumulates in TMEM columns [(k * MT + q) * NC, +NC).
#define PB_G 4                    // decode groups of the one-CTA-per-SM variant (the two-CTA variant runs 2)
#define PB_THREADS_OF(G) (32 * (4 * (G) + 1 + (G)))   // decode warps, one TMA warp, one MMA-issue warp per group
#define PB_BS 16                  // maximum depth of the smem ring of Uq tiles (TMA); the launch picks bs <= PB_BS
#define PB_PKG 3                  // per-warp cp.async ring depth (in the group's own stages): PB_PKG - 1 stages in flight
#define PB_AS 2                   // shared-memory A slots per decode group (decode of tile u+1 overlaps the MMAs of tile u)

#define PB_MAX_KB 512             // K * B entries of the per-bin mean term staged in shared memory
struct PbSmem {
  uint64_t full_a[PB_G * PB_AS], empty_a[PB_G * PB_AS], full_b[PB_BS], empty_b[PB_BS], acc_full;
  uint32_t tmem_base;
  double cs[PB_MAX_KB];           // per-(bin, column) mean term
  double dq[64];                  // per-column dequantisation factor 2^(e - F)
  int32_t cnt[256];               // rows per bin
};

// TMEM reads of the epilogues without a "memory" clobber (with one, the compiler re-derives every address and re-reads
// the kernel parameters after each access: 35 instructions per value instead of 15); the ordering the hardware needs is
// expressed through the registers themselves.
template <int W>
__device__ __forceinline__ void tmem_ldw_nc(uint32_t taddr, int32_t (&v)[W]) {
  static_assert(W == 2 || W == 4, "column chunk");
  if constexpr (W == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
  }
}
// the loaded registers may be used only after the wait: every one is passed through an (empty) volatile asm behind it
template <int L, int W>
__device__ __forceinline__ void tmem_ld_fence(int32_t (&a)[L][W]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < W; ++j) asm volatile("" : "+r"(a[l][j]));
}
__device__ __forceinline__ double lds_f64_nv(uint32_t addr) {     // not volatile: the scheduler may hoist and overlap these
  double r;
  asm("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(addr));
  return r;
}

// One chunk of W adjacent columns of a non-empty bin for one individual (TMEM lane): limbs recombined exactly in fp64,
// scale, mean term, row scale, one store and one RED per value.  `idx` is the individual's entry of the chunk's first
// column in P / S (32-bit element index when the accumulators hold fewer than 2^31 floats, else 64-bit); successive
// columns are `Np` floats apart.
template <int L, int W, typename IDX>
__device__ __forceinline__ void pb_chunk(uint32_t taddr, uint32_t stride, int nvalid, IDX Np, double rs, uint32_t dq_a,
                                         uint32_t cs_a, float* __restrict__ P_out, float* __restrict__ S_accum,
                                         bool wp, bool ws, IDX idx) {
  int32_t a[L][W];
#pragma unroll
  for (int l = 0; l < L; ++l) tmem_ldw_nc<W>(taddr + (uint32_t)l * stride, a[l]);
  tmem_ld_fence<L, W>(a);
  // all W values first, without branches (their dependent chains interleave), then the stores; a padding column
  // (j >= nvalid) reads table entries of its neighbour and is dropped
  float xf[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    double val = (double)a[L - 1][j];
#pragma unroll
    for (int l = L - 2; l >= 0; --l) val = fma(val, 256.0, (double)a[l][j]);   // exact: |value| < 2^53
    xf[j] = (float)(rs * (val * lds_f64_nv(dq_a + 8u * (uint32_t)j) - lds_f64_nv(cs_a + 8u * (uint32_t)j)));
  }
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (j < nvalid) {
      if (wp) P_out[idx] = xf[j];
      if (ws) atomicAdd(S_accum + idx, xf[j]);         // result unused -> RED: no load round trip
      idx += Np;
    }
  }
}

// Pass-B epilogue of one decode thread (TMEM lane = individual i): P[e][b][i] = rs_i * (acc * dq - cs) for the
// bins k = par (mod SI) of M-tile q.  Columns are read four at a time (one tcgen05.ld per limb row), the
// remainder of a row in pairs, so nothing is read beyond the limb rows of the (bin, tile) accumulator.
template <int L, typename IDX>
__device__ __forceinline__ void pb_epilogue(const int32_t* cnt, const double* dq_s, const double* cs_s, uint32_t trow, int i,
                                            int par, int SI, int q, int MT, int K, int K_total, int k0, int WG, int B, int Bs,
                                            int b0, int Bp, int NC, int Np,
                                            const float* __restrict__ rowscale, int rs_stride,
                                            float* __restrict__ P_out, float* __restrict__ S_accum) {
  const bool wp = P_out != nullptr, ws = S_accum != nullptr;
  const uint32_t dq0 = smem_u32(dq_s), cs0 = smem_u32(cs_s);
  for (int wg = 0; wg < WG; ++wg) {                  // weight group = RHS set (GxE) or operand (dominance) -> estimate wg * K + k
    const double rs = (double)rowscale[(size_t)wg * rs_stride + i];
    for (int k = par; k < K; k += SI) {
      const int e = wg * K_total + k0 + k;           // estimate index; k is local to this launch's bin group
      IDX idx = ((IDX)e * (IDX)Bs + (IDX)b0) * (IDX)Np + (IDX)i;
      if (cnt[k] <= 0) {                             // a bin without SNPs in this block: X (X^T Z) = 0
        if (wp)
          for (int b = 0; b < B; ++b, idx += (IDX)Np) P_out[idx] = 0.f;
        continue;
      }
      const uint32_t base = trow + (uint32_t)((k * MT + q) * NC + wg * L * Bp);
      uint32_t dq_a = dq0 + 8u * (uint32_t)(wg * B), cs_a = cs0 + 8u * (uint32_t)((wg * K + k) * B);
      int c0 = 0;
      for (; c0 + 4 <= Bp; c0 += 4) {
        pb_chunk<L, 4, IDX>(base + (uint32_t)c0, (uint32_t)Bp, B - c0, (IDX)Np, rs, dq_a, cs_a, P_out, S_accum, wp, ws, idx);
        dq_a += 32u; cs_a += 32u; idx += (IDX)4 * (IDX)Np;
      }
      for (; c0 < Bp; c0 += 2) {
        pb_chunk<L, 2, IDX>(base + (uint32_t)c0, (uint32_t)Bp, B - c0, (IDX)Np, rs, dq_a, cs_a, P_out, S_accum, wp, ws, idx);
        dq_a += 16u; cs_a += 16u; idx += (IDX)2 * (IDX)Np;
      }
    }
  }
}

// G = 4 decode groups and one CTA per SM, or G = 2 and two co-resident CTAs per SM (then one CTA's prologue and
// epilogue -- the write of its [bins x vectors x 128 individuals] results -- run underneath the other's main loop).
template <int MT, int G>
__global__ void __launch_bounds__(PB_THREADS_OF(G), G == 2 ? 2 : 1)
k_tc_pass_b(const __grid_constant__ CUtensorMap tm_uq, const uint8_t* __restrict__ bed, int pitch, int Np, int n_stage,
            const int32_t* __restrict__ meta_v, int n_chunk, const int32_t* __restrict__ stage_info,
            const int32_t* __restrict__ bin_count, int K, int K_total, int k0, int WG, int B, int Bs, int b0, int Bp, int L,
            int NC, int F, const unsigned int* __restrict__ wmax, const double* __restrict__ cs,
            const float* __restrict__ rowscale, int rs_stride, float* __restrict__ P_out, float* __restrict__ S_accum,
            uint32_t tmem_cols, int a_major, int kcap, int bs, int bzsh, int dbg) {
  if (RHE_DBG(16)) n_stage = 0;
  constexpr int PB_DW = 4 * G, PB_THREADS = PB_THREADS_OF(G);
  constexpr int SI = G / MT;                         // stage interleave between groups
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* tileA = smem;
  uint8_t* tileB = tileA + G * PB_AS * TC_TILE_A;
  const int tileB_bytes = NC * 128;
  uint8_t* packed = tileB + bs * tileB_bytes;     // [decode warp][PB_PKG][32 rows][32 B]
  PbSmem* sm = reinterpret_cast<PbSmem*>(packed + PB_DW * PB_PKG * 1024);
  const uint32_t tileA_s = smem_u32(tileA);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  const int i0 = blockIdx.x * (MT * 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < G * PB_AS; ++s) { mbar_init(&sm->full_a[s], 4); mbar_init(&sm->empty_a[s], 1); }   // one arrival per decode warp
    for (int s = 0; s < (bs >> bzsh); ++s) { mbar_init(&sm->full_b[s], 1); mbar_init(&sm->empty_b[s], MT << bzsh); }   // one pair per batch of 2^bzsh Uq tiles
    mbar_init(&sm->acc_full, G);
    fence_barrier_init();
  }
  if (warp == PB_DW + 1) tmem_alloc(&sm->tmem_base, tmem_cols);
  for (int i = threadIdx.x; i < WG * K * B; i += PB_THREADS) {   // mean terms of this group's bins [k0, k0 + K), columns [b0, b0 + B)
    const int wg = i / (K * B), rem = i - wg * (K * B), kl = rem / B, b = rem - kl * B;
    sm->cs[i] = cs[(size_t)(wg * K_total + k0 + kl) * Bs + b0 + b];
  }
  for (int i = threadIdx.x; i < K; i += PB_THREADS) sm->cnt[i] = bin_count[i];
  // power-of-two dequantisation factor 2^(e - F), 2^e > max |w| (same rule as k_tc_quant_w)
  // A non-finite weight (a monomorphic SNP: base.py:291-296 divides by sqrt(mu (1 - mu / 2)) = 0) has no fixed-point
  // image; it poisons its whole column exactly as the NaN does in the reference's products.
  for (int i = threadIdx.x; i < WG * B; i += PB_THREADS) {
    const int ex = (int)((wmax[(i / B) * Bs + b0 + i % B] >> 23) & 255u);
    sm->dq[i] = ex == 255 ? __longlong_as_double(0x7ff8000000000000ll) : ldexp(1.0, ex - 126 - F);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  if (warp < PB_DW && !(RHE_DBG(32))) {                   // zero the accumulators: quadrant per warp, columns split by group
    const uint32_t used = (uint32_t)(K * MT * NC);
    for (uint32_t c = (uint32_t)(warp >> 2) * 32; c < used; c += 32 * G)
      tmem_zero32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < PB_DW) {
    const int t = threadIdx.x & 127, g = warp >> 2;
    const int par = g / MT, q = g % MT;
    const uint8_t* base = bed + (i0 >> 2) + q * 32;
    const int n_own = n_stage > par ? (n_stage - par + SI - 1) / SI : 0;   // stages st = par + SI * u
    // Decode metadata (SNP row | fill << 24 | mode << 26, or -1 for padding) comes from the per-group layout
    // meta_v[par][chunk][t] = int4 of four consecutive own stages: ONE 16-byte load per four stages, issued two
    // chunks ahead (per-stage 4-byte loads share hardware scoreboards with the newest load and stall on it every
    // stage).  The packed bytes travel through a per-warp cp.async ring, PB_PKG - 1 stages ahead: lane pair
    // (2 i, 2 i + 1) copies the two 16-byte halves of row 16 h + i of the warp's 32 rows (h = 0, 1), so every request
    // is a whole 32-byte sector, and the row's owner thread reads it back after the warp has synchronised.  The row
    // address of the pair comes from the owner's meta by shuffle.
    const int4* mv = reinterpret_cast<const int4*>(meta_v) + (size_t)par * n_chunk * 128 + t;
    const int4 none = make_int4(-1, -1, -1, -1);
    auto chunk_of = [&](int c) { return c < n_chunk ? __ldg(mv + (size_t)c * 128) : none; };
    auto comp = [](const int4& m, int r) { return r == 0 ? m.x : r == 1 ? m.y : r == 2 ? m.z : m.w; };
    const uint32_t ring = smem_u32(packed) + (uint32_t)warp * (PB_PKG * 1024);
    // slot layout [32 rows][32 B]: the two halves of a sector stay adjacent (the sector lands as one shared-memory
    // wavefront instead of two) and swap places in every other group of four rows (conflict-free 16-byte owner reads)
    const uint32_t my_lo = ring + (uint32_t)lane * 32 + (((uint32_t)lane >> 2) & 1u) * 16;    // owner view, half 0
    const uint32_t my_hi = my_lo ^ 16u;
    const uint32_t cp_row = (uint32_t)(lane >> 1);                                              // copier view, h = 0
    const uint32_t cp_dst = ring + cp_row * 32 + ((((uint32_t)lane & 1u) ^ ((cp_row >> 2) & 1u)) * 16);
    const uint8_t* cp_base = base + (lane & 1) * 16;
    auto issue = [&](int meta, uint32_t slot) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mrow = __shfl_sync(0xffffffffu, meta, 16 * h + (lane >> 1));
        if (mrow >= 0 && !(RHE_DBG(1)))
          cp_async16(cp_dst + slot * 1024u + (uint32_t)(h * 512), cp_base + (size_t)(mrow & 0xFFFFFF) * pitch);
      }
      cp_async_commit();
    };
    constexpr int AHEAD = 2;                             // stages of packed bytes in flight
    static_assert(PB_PKG == AHEAD + 1, "ring depth");
    int4 cur = chunk_of(0), nxt = chunk_of(1);
    issue(n_own > 0 ? cur.x : -1, 0u);
    issue(n_own > 1 ? cur.y : -1, 1u);
    uint32_t rd_slot = 0, wr_slot = AHEAD;               // ring slots of stage u and stage u + AHEAD
    const uint32_t tile0 = tileA_s + g * PB_AS * TC_TILE_A;
    uint64_t* const full0 = &sm->full_a[g * PB_AS];
    uint64_t* const empty0 = &sm->empty_a[g * PB_AS];
    PROF_T0();
    static_assert(PB_AS == 2, "slot parity below");
    for (int c = 0; 4 * c < n_own; ++c) {
      const int4 nn = chunk_of(c + 2);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int u = 4 * c + r;
        if (u < n_own) {                                 // warp-uniform
          const int m_iss = u + AHEAD < n_own ? (r < 2 ? comp(cur, r + 2) : comp(nxt, r - 2)) : -1;
          issue(m_iss, wr_slot);                         // stage u + AHEAD
          wr_slot = wr_slot + 1 == PB_PKG ? 0u : wr_slot + 1;
          const int meta = comp(cur, r);
          const uint32_t tab = meta >= 0 ? tc_value_table(((uint32_t)meta >> 24) & 3u, (meta >> 26) & 1) : 0u;
          const int a = r & 1;                           // 4 c is even: slot = u % 2, use index u / 2
          cp_async_wait<AHEAD>();                        // stage u has landed (the AHEAD later ones may be in flight)
          __syncwarp();
          const uint4 lo = lds128(my_lo + rd_slot * 1024u), hi = lds128(my_hi + rd_slot * 1024u);
          rd_slot = rd_slot + 1 == PB_PKG ? 0u : rd_slot + 1;
          PROF_ADD(0);
          if (!(RHE_DBG(128))) mbar_wait(empty0 + a, ((u >> 1) & 1) ^ 1);
          PROF_ADD(1);
          if (!(RHE_DBG(2))) tc_store_row(tile0 + a * TC_TILE_A, t, lo, hi, tab);
          PROF_ADD(2);
          fence_proxy_async();
          __syncwarp();                                  // every lane's rows are fenced: one arrival per warp
          if (lane == 0 && !(RHE_DBG(128))) mbar_arrive(full0 + a);
          PROF_ADD(4);
        }
      }
      cur = nxt;
      nxt = nn;
    }
    // ---- epilogue: TMEM lane = position inside M-tile q; group (par, q) takes the bins k = par (mod SI)
    mbar_wait(&sm->acc_full, 0);
    tc_fence_after();
    PROF_ADD(5);
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int i = i0 + q * 128 + (t & ~15) + tc_perm16(t & 15);
    if (!(RHE_DBG(4))) {
      // 32-bit element indices into P / S whenever the accumulators hold fewer than 2^31 floats (one IMAD.WIDE per access)
      const bool small = (long long)WG * K_total * Bs * (long long)Np < (1ll << 31);
#define PB_EPI(L_, T_) pb_epilogue<L_, T_>(sm->cnt, sm->dq, sm->cs, trow, i, par, SI, q, MT, K, K_total, k0, WG, B, Bs, b0, Bp, NC, Np, rowscale, rs_stride, P_out, S_accum)
      if (small) { if (L == 2) PB_EPI(2, uint32_t); else if (L == 3) PB_EPI(3, uint32_t); else PB_EPI(4, uint32_t); }
      else { if (L == 2) PB_EPI(2, uint64_t); else if (L == 3) PB_EPI(3, uint64_t); else PB_EPI(4, uint64_t); }
#undef PB_EPI
    }
    PROF_ADD(6);
    PROF_FLUSH(0);
    tc_fence_before();
  } else if (warp == PB_DW) {
    {
      // TMA producer: Uq tile of stage st -> ring slot st % bs, in batches of 2^bzsh consecutive stages that share one
      // barrier pair (one wait and one expect_tx per batch: this single thread is otherwise the CTA's serial bottleneck).
      const uint32_t fb = smem_u32(&sm->full_b[0]), eb = smem_u32(&sm->empty_b[0]);
      const int bz = 1 << bzsh;
      uint32_t bar = 0, dst = smem_u32(tileB), wait_par = 1;
      int b = 0;
      for (int st0 = 0; st0 < n_stage; st0 += bz) {    // the whole warp runs the loop; one elected lane issues
        const int n_in = min(bz, n_stage - st0);
        mbar_wait_s(eb + bar, wait_par);
        if (elect_one()) {
          if (RHE_DBG(8)) mbar_arrive_s(fb + bar);
          else {
            mbar_expect_tx_s(fb + bar, (uint32_t)(n_in * tileB_bytes));
            for (int i = 0; i < n_in; ++i)
              tma_load_2d_s(dst + (uint32_t)(i * tileB_bytes), &tm_uq, fb + bar, (st0 + i) * 128, 0);
          }
        }
        __syncwarp();
        bar += 8; dst += (uint32_t)(bz * tileB_bytes); b += bz;
        if (b == bs) { b = 0; bar = 0; dst = smem_u32(tileB); wait_par ^= 1u; }
      }
    }
  } else {
    // ---- MMA issue: one warp (one elected lane) per decode group (par, q); all MMAs accumulate into zeroed TMEM.
    // The loop is one thread's dependent-latency chain: slot indices, parities and descriptors advance incrementally.
    {
      const int g = warp - (PB_DW + 1);
      const int par = g / MT, q = g % MT;
      const uint32_t idesc = idesc_i8(128, NC, a_major);   // A is MN-major: 128 individuals contiguous per SNP row
      const uint32_t fb = smem_u32(&sm->full_b[0]), eb = smem_u32(&sm->empty_b[0]);
      const uint32_t fa = smem_u32(&sm->full_a[g * PB_AS]), ea = smem_u32(&sm->empty_a[g * PB_AS]);
      const uint64_t adesc0 = smem_desc_sw128(tileA_s + g * PB_AS * TC_TILE_A, TC_TILE_A, 1024);
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(tileB), 16, 1024);
      const uint32_t dq = tmem + (uint32_t)(q * NC);
      PROF_T0();
      int b = par;                                     // par < SI <= bs
      uint32_t ph_b = 0, a = 0, ph_a = 0;
      uint32_t info = par < n_stage ? (uint32_t)__ldg(stage_info + par) : 0u;      // one stage ahead, straight from L2
      for (int st = par; st < n_stage; st += SI) {
        const uint32_t k = info & 255u;
        const int ksteps = min((int)(info >> 16), kcap);
        if (st + SI < n_stage) info = (uint32_t)__ldg(stage_info + st + SI);
        PROF_ADD(0);
        mbar_wait_s(fb + 8u * (uint32_t)(b >> bzsh), ph_b);
        PROF_ADD(1);
        if (!(RHE_DBG(128))) mbar_wait_s(fa + 8u * a, ph_a);
        PROF_ADD(2);
        tc_fence_after();
        const uint32_t dcol = dq + k * (uint32_t)(MT * NC);
        const uint64_t adesc = adesc0 + (uint64_t)(a * (TC_TILE_A >> 4));
        const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)b * (uint32_t)(tileB_bytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j)   // K = 32 SNP rows per instruction: 32 rows x 128 B further down the tile
            if (j < ksteps) umma_i8(dcol, adesc + (uint64_t)(j * 256), bdesc + (uint64_t)(j * 2), idesc, 1u);
          const uint32_t ebb = eb + 8u * (uint32_t)(b >> bzsh);
          if (RHE_DBG(128)) { mbar_arrive_s(ebb); }
          else if (RHE_DBG(64)) { mbar_arrive_s(ea + 8u * a); mbar_arrive_s(ebb); }
          else { umma_commit_s(ea + 8u * a); umma_commit_s(ebb); }
        }
        __syncwarp();
        PROF_ADD(3);
        b += SI;
        if (b >= bs) { b -= bs; ph_b ^= 1u; }
        a ^= 1u;
        if (a == 0u) ph_a ^= 1u;
      }
      if (elect_one()) umma_commit(&sm->acc_full);
      __syncwarp();
      PROF_FLUSH(8);
    }
  }
  __syncthreads();
  if (warp == PB_DW + 1) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// ------------------------------------------------------------------------------------------ pass B from TMEM (fast path)
// The gather pass B above is bound by the shared-memory data pipe: every genotype is written to shared memory as a
// byte and read back by the MMA.  When the block also has an INDIVIDUAL-MAJOR copy
//     GT [Np / 128][n_ss][128][128 B]   one contiguous 16 KB box per (M-tile of 128 individuals, super-stage of 512
//                         bin-sorted positions of the block's plan): row = individual, 2 bits per position, imputation
//                         already applied (no missing code left), 16 positions per 32-bit word
// the product D[i, n] = sum_s G[i, s] Uq[s, n] has the shape of pass A with the roles swapped: the TMEM lane is the
// individual, K runs over positions, and the expanded A operand goes from registers straight into TENSOR MEMORY.
// Only the small Uq tiles travel through shared memory.  GT is written once per resident block by k_tc_transpose
// (ingest); it doubles the genotype footprint, so the engine keeps it for as many blocks as the HBM allows.
//
// Positions are bin-sorted, so a CTA (128 individuals, all positions) finishes bin k before it starts bin k + 1: the
// accumulator of a bin lives in one of `n_acc` TMEM buffers of NC columns and is drained (scaled, written to P / S,
// re-zeroed) by the drain warps while the next bins accumulate in the other buffers: n_acc NC + 8 A slots of TMEM
// columns for any number of bins, and the epilogue overlaps the main loop.  The CTA is persistent: it walks the
// M-tiles blockIdx.x, blockIdx.x + gridDim.x, ... and the sequence of sub-tiles, super-stages and bins simply
// continues from one M-tile into the next, so rings and barriers never restart.
// One persistent CTA per SM (960 threads, 64 registers): warps 0-15 decode (four groups of four, one TMEM lane quadrant
// per warp), 16 TMA producer of the genotype boxes, 17 TMA producer of the Uq tiles, 18-21 MMA issue (one per decode
// group), 22-29 drain (two sets of four consecutive warps = four lane quadrants; every set takes part of the columns of
// every bin: each drain warp runs alone on its scheduler slot at the latency of its dependent chain, so the per-bin
// latency, not the instruction count, is what has to stay below the time the issuers need for a bin).  The decode loop
// is the spill-free loop of pass A; the drains never touch it.  The two rings are fed independently: the genotype ring must run several
// super-stages ahead of the decode front (about 2.5 stages are being consumed at any time and HBM latency is ~2 us), while
// a Uq slot only frees up when its MMAs have completed.
#define P2_G 4
#define P2_DW (4 * P2_G)
#define P2_W_PROD P2_DW
#define P2_W_PRODU (P2_DW + 1)
#define P2_W_MMA (P2_DW + 2)
#define P2_W_DRAIN (P2_DW + 2 + P2_G)
#define P2_NDRAIN 8               // drain warps: two sets of four (lane quadrants), each set takes part of the columns
#define P2_THREADS (32 * (P2_W_DRAIN + P2_NDRAIN))
#define P2_AS (2 * P2_G)          // TMEM A slots (32 columns each): two per group
#define P2_MAXRING 8
#define P2_MAXACC 4               // accumulator buffers (bins in flight between the MMA issuers and the drain warps)
#define P2_MAXSUB 1024            // sub-tiles (128 positions) of one M-tile, all operands
#ifndef P2_DRAIN_BY_BIN
#define P2_DRAIN_BY_BIN 1         // the two drain sets alternate BINS (1) or split the columns of every bin (0)
#endif

struct P2Smem {
  uint64_t full_a[P2_AS], empty_a[P2_AS], full_u[P2_MAXRING], empty_u[P2_MAXRING], full_g[P2_MAXRING], empty_g[P2_MAXRING];
  uint64_t bin_full[P2_MAXACC], acc_free[P2_MAXACC];
  uint32_t tmem_base;
  double cs[PB_MAX_KB];           // per-(weight group, bin, column) mean term
  double dq[64];                  // per-(weight group, column) dequantisation factor 2^(e - F)
  int32_t cnt[256];               // rows per bin
  int32_t vbin[P2_MAXSUB];        // virtual bin (operand * K + bin) of every sub-tile of an M-tile
};

template <int W>
__device__ __forceinline__ void tmem_zero_nc(uint32_t taddr) {     // zero W columns of this thread's lane (completion: tmem_st_wait)
  const uint32_t z = 0u;
  if constexpr (W == 4) asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(taddr), "r"(z));
  else asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %1};" ::"r"(taddr), "r"(z));
}

// One chunk of W adjacent columns of a non-empty bin for this thread's individual (TMEM lane): the same arithmetic as
// pb_epilogue -- limbs recombined exactly in fp64, scale, mean term, row scale, one store and one RED per value.  `idx`
// is the individual's entry of the chunk's first column in P / S (32-bit element index: the launch checks that the
// accumulators hold fewer than 2^31 floats); successive columns are `Np` floats apart.  The columns are zeroed again as
// soon as they are in registers (all MMAs accumulate).
template <int L, int W, int SMALL>
__device__ __forceinline__ void p2_drain_chunk(uint32_t taddr, uint32_t stride, int nvalid, uint32_t Np, double rs, uint32_t dq_a,
                                               uint32_t cs_a, float* __restrict__ P_out, float* __restrict__ S_accum,
                                               bool wp, bool ws, uint32_t idx) {
  int32_t a[L][W];
#pragma unroll
  for (int l = 0; l < L; ++l) tmem_ldw_nc<W>(taddr + (uint32_t)l * stride, a[l]);
  tmem_ld_fence<L, W>(a);
#pragma unroll
  for (int l = 0; l < L; ++l) tmem_zero_nc<W>(taddr + (uint32_t)l * stride);
  // all W values first, without branches (the chains of the W columns interleave: the drain warps run alone on their
  // scheduler slots, so their latency is the dependent chain of whatever sits between two branches), then the stores
  float xf[W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    double val;
    if constexpr (SMALL && L >= 2) {
      // bins of at most 2^14 positions: |limb sum| <= 2 x 128 x 2^14 = 2^22, so the two low limbs combine exactly in
      // int32 (one integer multiply-add on the FMA pipe instead of a conversion and an fp64 FMA)
      const int32_t lo = a[1][j] * 256 + a[0][j];
      val = (double)lo;
      if constexpr (L >= 3) val = fma((double)a[2][j], 65536.0, val);
      if constexpr (L >= 4) val = fma((double)a[3][j], 16777216.0, val);
    } else {
      val = (double)a[L - 1][j];
#pragma unroll
      for (int l = L - 2; l >= 0; --l) val = fma(val, 256.0, (double)a[l][j]);   // exact: |value| < 2^53
    }
    xf[j] = (float)(rs * (val * lds_f64_nv(dq_a + 8u * (uint32_t)j) - lds_f64_nv(cs_a + 8u * (uint32_t)j)));
  }
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (j < nvalid) {
      if (wp) P_out[idx] = xf[j];
      if (ws) atomicAdd(S_accum + idx, xf[j]);         // result unused -> RED
      idx += Np;
    }
  }
}

// Drain the columns [c_lo, c_hi) of every weight group of bin k for this thread's individual.
template <int L, int SMALL>
__device__ __forceinline__ void p2_drain(uint32_t tcol, int c_lo, int c_hi, int k, int i, int K, int WG, int B, int Bp,
                                         uint32_t Np, uint32_t dq_s, uint32_t cs_s, float rs0, float rs1,
                                         float* __restrict__ P_out, float* __restrict__ S_accum) {
  const bool wp = P_out != nullptr, ws = S_accum != nullptr;
  for (int wg = 0; wg < WG; ++wg) {                    // four columns per tcgen05.ld, the remainder in pairs
    const uint32_t base = tcol + (uint32_t)(wg * L * Bp);
    const double rs = (double)(wg ? rs1 : rs0);        // row scale of this individual: loaded once per M-tile, not per bin
    uint32_t dq_a = dq_s + 8u * (uint32_t)(wg * B + c_lo), cs_a = cs_s + 8u * (uint32_t)((wg * K + k) * B + c_lo);
    uint32_t idx = (uint32_t)((wg * K + k) * B + c_lo) * Np + (uint32_t)i;
    int c0 = c_lo;
    for (; c0 + 4 <= c_hi; c0 += 4) {
      p2_drain_chunk<L, 4, SMALL>(base + (uint32_t)c0, (uint32_t)Bp, B - c0, Np, rs, dq_a, cs_a, P_out, S_accum, wp, ws, idx);
      dq_a += 32u; cs_a += 32u; idx += 4u * Np;
    }
    for (; c0 < c_hi; c0 += 2) {
      p2_drain_chunk<L, 2, SMALL>(base + (uint32_t)c0, (uint32_t)Bp, B - c0, Np, rs, dq_a, cs_a, P_out, S_accum, wp, ws, idx);
      dq_a += 16u; cs_a += 16u; idx += 2u * Np;
    }
  }
}

// Sub-tile j = 128 positions of one M-tile; super-stage k = 4 sub-tiles = one TMA box of GT (128 individuals x 128 B) +
// four Uq tiles per operand; n_ss super-stages per M-tile.  RHE-DOM has two operands (n_modes = 2: the count and
// [g == 2], rhe_dom.py:36-39): every sub-tile is expanded twice from the same packed words, with the operand's mask and
// against the operand's Uq tile, into the SAME accumulator -- P = rs (dq (D0 + D1) - cs) exactly as the gather kernel
// forms it, and the copy is read once.
template <int GS, int US>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_tc_pass_b2(const __grid_constant__ CUtensorMap tm_uq, const __grid_constant__ CUtensorMap tm_gt, int n_mt, int Np, int n_ss,
             int n_modes, const int32_t* __restrict__ stage_info, const int32_t* __restrict__ bin_count, int K, int WG,
             int B, int Bp, int L, int NC, int F, const unsigned int* __restrict__ wmax, const double* __restrict__ cs,
             const float* __restrict__ rowscale, int rs_stride, float* __restrict__ P_out, float* __restrict__ S_accum,
             uint32_t tmem_cols, uint32_t col_a, int n_acc, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int acc_stride = (NC + 31) & ~31;            // TMEM columns per accumulator buffer (zeroed 32 columns at a time)
  uint8_t* packed = smem;                            // [GS][128 individuals][128 B], 128-byte swizzle
  uint8_t* tileU = packed + GS * PA_PACKED;          // [US][n_modes][4][NC][128 B]
  const int tileU_bytes = NC * 128;
  P2Smem* sm = reinterpret_cast<P2Smem*>(tileU + US * 4 * n_modes * tileU_bytes);
  const uint32_t packed_s = smem_u32(packed);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  const int n_it = n_mt > (int)blockIdx.x ? (n_mt - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;   // own M-tiles
  const int total_ss = n_ss, n_sub = 4 * total_ss, V = K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P2_AS; ++s) { mbar_init(&sm->full_a[s], 4); mbar_init(&sm->empty_a[s], 1); }   // one arrival per decode warp
    for (int s = 0; s < P2_MAXRING; ++s) {
      mbar_init(&sm->full_u[s], 1); mbar_init(&sm->empty_u[s], 4 * n_modes);   // 4 sub-tiles x operands
      mbar_init(&sm->full_g[s], 1); mbar_init(&sm->empty_g[s], 16);     // 4 sub-tiles x 4 warps
    }
    for (int s = 0; s < P2_MAXACC; ++s) {
      mbar_init(&sm->bin_full[s], P2_G);
      mbar_init(&sm->acc_free[s], P2_DRAIN_BY_BIN ? P2_NDRAIN / 2 : P2_NDRAIN);
    }
    fence_barrier_init();
  }
  if (warp == P2_W_PROD) tmem_alloc(&sm->tmem_base, tmem_cols);
  for (int i = threadIdx.x; i < WG * K * B; i += P2_THREADS) sm->cs[i] = cs[i];
  for (int i = threadIdx.x; i < K; i += P2_THREADS) sm->cnt[i] = bin_count[i];
  for (int i = threadIdx.x; i < n_sub; i += P2_THREADS) sm->vbin[i] = stage_info[i] & 255;   // bin of every sub-tile
  for (int i = threadIdx.x; i < WG * B; i += P2_THREADS) {
    const int ex = (int)((wmax[i] >> 23) & 255u);       // same rule as k_tc_quant_w / the gather kernel
    sm->dq[i] = ex == 255 ? __longlong_as_double(0x7ff8000000000000ll) : TC_VALUE_SCALE * ldexp(1.0, ex - 126 - F);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  if (warp < 4)                                        // zero the accumulator buffers (lane quadrant per warp)
    for (uint32_t c = 0; c < col_a; c += 32) tmem_zero32(tmem + ((uint32_t)(warp * 32) << 16) + c);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < P2_DW) {
    // ---- decode: group g expands sub-tile g of EVERY super-stage (P2_G == 4 sub-tiles per box), so the chunk it reads
    // inside a staged box is fixed and every ring index advances by one per iteration: no division, no modulo.
    static_assert(P2_G == 4, "group g = sub-tile g of every super-stage");
    const int t = (warp & 3) * 32 + lane, g = warp >> 2;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t sw = (uint32_t)(t & 7);                      // 128-byte swizzle: 16-byte chunk c sits at c ^ (row & 7)
    const uint32_t row_lo = packed_s + (uint32_t)t * 128 + ((((uint32_t)(2 * g)) ^ sw) << 4);
    const uint32_t row_hi = packed_s + (uint32_t)t * 128 + ((((uint32_t)(2 * g + 1)) ^ sw) << 4);
    const uint32_t fg = smem_u32(&sm->full_g[0]), eg = smem_u32(&sm->empty_g[0]);
    const uint32_t fa = smem_u32(&sm->full_a[2 * g]), ea = smem_u32(&sm->empty_a[2 * g]);
    const uint32_t dst0 = lane_base + col_a + 32u * (uint32_t)(2 * g);
    const int n_k = n_it * total_ss;                          // super-stages of the whole run
    uint32_t fsl = 0, fpar = 0;                                // ring slot / parity of the next box to read
    uint32_t csl = 0;                                          // ring slot of the box being expanded
    uint32_t aq = 0, apar = 1;                                 // the group's A slot of the next expansion / parity to wait for
    // The individual-major copy carries no missing code and stores the operand VALUE in its two bits (imputation and the
    // PLINK code -> count map are applied at ingest), so a word expands with one mask per output register and three
    // multiply-high shifts -- four ALU-pipe and three FMA-pipe instructions per sixteen genotypes -- instead of the
    // table look-up of pass A (two masks, four byte permutes, three shifts).  Register k of a word holds the fields
    // k, k + 4, k + 8, k + 12: k_tc_transpose places position 4 k + j of a word's sixteen at field 4 j + k.
    // [g == 2] is bit 1 of the value: one more shift per word.
    auto fetch = [&](uint4& lo, uint4& hi) {
      mbar_wait_s(fg + 8u * fsl, fpar);
      lo = lds128(row_lo + fsl * PA_PACKED);
      hi = lds128(row_hi + fsl * PA_PACKED);
      if (++fsl == (uint32_t)GS) { fsl = 0; fpar ^= 1u; }
    };
    auto expand0 = [](uint32_t w) {
#if TC_SHIFT_VARIANT == 2
      const uint32_t m = 0x0C0C0C0Cu;
      uint4 r;
      r.x = tc_shl_fma(w, 4u) & m;
      r.y = w & m;
      r.z = (w >> 2) & m;
      r.w = (w >> 4) & m;
      return r;
#else
      const uint32_t m = 0x03030303u;
      uint4 r;
      r.x = w & m;
#if TC_SHIFT_VARIANT == 1
      r.y = (w >> 2) & m;
      r.z = (w >> 4) & m;
      r.w = (w >> 6) & m;
#else
      r.y = tc_shr_fma(w, 1u << 30) & m;
      r.z = tc_shr_fma(w, 1u << 28) & m;
      r.w = tc_shr_fma(w, 1u << 26) & m;
#endif
      return r;
#endif
    };
    auto expand1 = [](uint32_t w) {
#if TC_SHIFT_VARIANT == 2
      const uint32_t m4 = 0x04040404u;
      uint4 r4;
      r4.x = tc_shl_fma(w, 2u) & m4;
      r4.y = (w >> 1) & m4;
      r4.z = (w >> 3) & m4;
      r4.w = (w >> 5) & m4;
      return r4;
#else
      const uint32_t m = 0x01010101u;
      uint4 r;
#if TC_SHIFT_VARIANT == 1
      r.x = (w >> 1) & m;
      r.y = (w >> 3) & m;
      r.z = (w >> 5) & m;
      r.w = (w >> 7) & m;
#else
      r.x = tc_shr_fma(w, 1u << 31) & m;
      r.y = tc_shr_fma(w, 1u << 29) & m;
      r.z = tc_shr_fma(w, 1u << 27) & m;
      r.w = tc_shr_fma(w, 1u << 25) & m;
#endif
      return r;
#endif
    };
    // one operand of one sub-tile -> the group's next A slot; `release`: the packed words are not needed again
    auto put = [&](const uint4& lo, const uint4& hi, auto expand, bool release) {
      mbar_wait_s(ea + 8u * aq, apar);
      tc_fence_after();
      const uint32_t dst = dst0 + 32u * aq;
      uint4 r[4];
      r[0] = expand(lo.x); r[1] = expand(lo.y); r[2] = expand(lo.z); r[3] = expand(lo.w);
      tmem_st16(dst, r);
      uint4 r2[4];
      r2[0] = expand(hi.x); r2[1] = expand(hi.y); r2[2] = expand(hi.z); r2[3] = expand(hi.w);
      tmem_st16(dst + 16, r2);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();                                    // every lane's stores are complete and fenced: one arrival per warp
      if (lane == 0) {
        if (release) mbar_arrive_s(eg + 8u * csl);
        mbar_arrive_s(fa + 8u * aq);
      }
      if (release && ++csl == (uint32_t)GS) csl = 0;
      aq ^= 1u;
      if (aq == 0u) apar ^= 1u;
    };
    // Software pipeline, unrolled by two so that the register sets swap roles without moves: the packed words of the
    // next box are read from the ring before the current one is expanded.
    uint4 lo0, hi0, lo1, hi1;
    if (n_k > 0) fetch(lo0, hi0);
    if (n_modes == 1) {
      for (int k = 0; k < n_k; k += 2) {
        if (k + 1 < n_k) fetch(lo1, hi1);
        put(lo0, hi0, expand0, true);
        if (k + 1 < n_k) {
          if (k + 2 < n_k) fetch(lo0, hi0);
          put(lo1, hi1, expand0, true);
        }
      }
    } else {
      for (int k = 0; k < n_k; k += 2) {
        if (k + 1 < n_k) fetch(lo1, hi1);
        put(lo0, hi0, expand0, false);
        put(lo0, hi0, expand1, true);
        if (k + 1 < n_k) {
          if (k + 2 < n_k) fetch(lo0, hi0);
          put(lo1, hi1, expand0, false);
          put(lo1, hi1, expand1, true);
        }
      }
    }
    tc_fence_before();
  } else if (warp == P2_W_PROD) {
    // TMA producer 1: the genotype box of every super-stage (positions 512 kk .. of the M-tile's 128 individuals) -> ring
    // slot Kg % GS, as far ahead of the decode front as the ring allows.
    const uint32_t fg = smem_u32(&sm->full_g[0]), eg = smem_u32(&sm->empty_g[0]);
    uint32_t gsl = 0, gpar = 1;
    for (int it = 0; it < n_it; ++it) {
      const int mt = (int)blockIdx.x + it * (int)gridDim.x;
      for (int k = 0; k < total_ss; ++k) {             // the whole warp runs the loop; one elected lane issues
        mbar_wait_s(eg + 8u * gsl, gpar);
        if (elect_one()) {
          mbar_expect_tx_s(fg + 8u * gsl, PA_PACKED);
          tma_load_2d_s(packed_s + gsl * PA_PACKED, &tm_gt, fg + 8u * gsl, 0, (mt * n_ss + k) * 128);
        }
        __syncwarp();
        if (++gsl == (uint32_t)GS) { gsl = 0; gpar ^= 1u; }
      }
    }
  } else if (warp == P2_W_PRODU) {
    // TMA producer 2: the four Uq tiles per operand of every super-stage -> ring slot Kg % US (the same tiles for every
    // M-tile: L2 hits); operand md occupies the positions [md n_pos, (md + 1) n_pos) of the weight buffer
    const uint32_t fu = smem_u32(&sm->full_u[0]), eu = smem_u32(&sm->empty_u[0]);
    const uint32_t tileU_s = smem_u32(tileU);
    uint32_t usl = 0, upar = 1;
    for (int it = 0; it < n_it; ++it) {
      for (int k = 0; k < total_ss; ++k) {
        mbar_wait_s(eu + 8u * usl, upar);
        if (elect_one()) {
          mbar_expect_tx_s(fu + 8u * usl, 4u * (uint32_t)(n_modes * tileU_bytes));
          for (int md = 0; md < n_modes; ++md)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tma_load_2d_s(tileU_s + ((usl * (uint32_t)n_modes + (uint32_t)md) * 4u + (uint32_t)q) * (uint32_t)tileU_bytes, &tm_uq,
                            fu + 8u * usl, (md * n_sub + 4 * k + q) * 128, 0);
        }
        __syncwarp();
        if (++usl == (uint32_t)US) { usl = 0; upar ^= 1u; }
      }
    }
  } else if (warp < P2_W_DRAIN) {
    // ---- MMA issue: one warp per decode group = sub-tile g of every super-stage.  Every issuer walks ALL virtual bins of
    // the run in order (bin v of M-tile `it` is Vg = it V + v) -- enter (wait until the bin's accumulator buffer has been
    // drained of the bin n_acc before it), issue its own sub-tiles of the bin, leave (commit: the buffer's bin_full
    // barrier counts the issuers) -- so no barrier ever sees arrivals of two phases.  The loop is one thread's latency
    // chain and sits on the round trip of the A slots (decode -> MMA -> slot free): ring indices, parities and
    // descriptors advance incrementally and the bin of a sub-tile comes from shared memory.  Every sub-tile issues
    // all four K-steps: rows beyond the end of a bin are zero in the copy and in the weights.
    const int g = warp - P2_W_MMA;
    const uint32_t idesc = idesc_i8(128, NC, 0);
    const uint32_t fu = smem_u32(&sm->full_u[0]), eu = smem_u32(&sm->empty_u[0]);
    const uint32_t fa = smem_u32(&sm->full_a[2 * g]), ea = smem_u32(&sm->empty_a[2 * g]);
    const uint32_t bfull = smem_u32(&sm->bin_full[0]), afree = smem_u32(&sm->acc_free[0]);
    const uint64_t bdesc_g = smem_desc_sw128(smem_u32(tileU), 16, 1024) + (uint64_t)((uint32_t)g * (uint32_t)(tileU_bytes >> 4));
    const uint32_t mode_step = 4u * (uint32_t)(tileU_bytes >> 4), slot_step = mode_step * (uint32_t)n_modes;
    const uint32_t acol0 = tmem + col_a + 32u * (2u * (uint32_t)g);
    const uint32_t vb_s = smem_u32(sm->vbin) + 4u * (uint32_t)g;          // vbin[4 kin + g]
    const int n_k = n_it * total_ss;
    const int V_all = n_it * V, am = n_acc - 1, sh = n_acc == 4 ? 2 : n_acc - 1;   // n_acc is 1, 2 or 4: Vg % n_acc = Vg & am, Vg / n_acc = Vg >> sh
    uint32_t ph_a = 0, q = 0, usl = 0, upar = 0, desc_off = 0;
    int cur_v = 0, vbase = 0, kin = 0;
    auto enter = [&](int v) {                          // buffer v % n_acc must have been drained of bin v - n_acc
      if (v >= n_acc) mbar_wait_s(afree + 8u * (uint32_t)(v & am), (uint32_t)((v >> sh) - 1) & 1u);
      tc_fence_after();
    };
    auto leave = [&](int v) {
      if (elect_one()) umma_commit_s(bfull + 8u * (uint32_t)(v & am));
      __syncwarp();
    };
    if (V_all > 0) enter(0);
    for (int k = 0; k < n_k; ++k) {
      const int v = vbase + (int)lds32(vb_s + 16u * (uint32_t)kin);
      if (++kin == total_ss) { kin = 0; vbase += V; }
      while (cur_v < v) { leave(cur_v); ++cur_v; enter(cur_v); }
      mbar_wait_s(fu + 8u * usl, upar);
      const uint32_t dcol = tmem + (uint32_t)((v & am) * acc_stride);
      for (int md = 0; md < n_modes; ++md) {           // the operands of the sub-tile: same accumulator, own A slot and Uq tile
        mbar_wait_s(fa + 8u * q, ph_a);
        tc_fence_after();
        const uint64_t bdesc = bdesc_g + (uint64_t)(desc_off + (uint32_t)md * mode_step);
        const uint32_t acol = acol0 + 32u * q;
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 4; ++i)   // K = 32 positions per instruction: 8 TMEM columns / 32 bytes of the Uq row
            umma_i8_ts(dcol, acol + 8u * i, bdesc + (uint64_t)(i * 2), idesc, 1u);
          umma_commit_s(ea + 8u * q);
          umma_commit_s(eu + 8u * usl);
        }
        __syncwarp();
        q ^= 1u;
        if (q == 0u) ph_a ^= 1u;
      }
      desc_off += slot_step;
      if (++usl == (uint32_t)US) { usl = 0; upar ^= 1u; desc_off = 0; }
    }
    while (cur_v < V_all) { leave(cur_v); ++cur_v; if (cur_v < V_all) enter(cur_v); }
  } else {
    // ---- drain warps (TMEM lane quadrant = warp & 3): bin after bin, in sequence order.  The same eight warps drain every
    // bin, so each accumulator buffer's barriers advance one phase per use and a parity wait is never ambiguous.  Set 0
    // takes the columns [0, c_split) of every weight group, set 1 the columns [c_split, Bp).
    const int t = (warp & 3) * 32 + lane, am = n_acc - 1, sh = n_acc == 4 ? 2 : n_acc - 1;
    // P2_DRAIN_BY_BIN: set s takes the bins Vg = s (mod 2) whole -- an accumulator buffer (Vg mod n_acc, n_acc even) is then
    // always drained by the same set, and a set has two bin periods for its bin: with ten vectors a column split is 4 | 6
    // (two chunks on the critical path of every bin), alternate bins are three chunks per two bin periods.
    const int set = (warp - P2_W_DRAIN) >> 2;
    const int c_split = min(Bp, 4 * ((Bp / 2 + 2) / 4));
    const int c_lo = P2_DRAIN_BY_BIN ? 0 : (set ? c_split : 0), c_hi = P2_DRAIN_BY_BIN ? Bp : (set ? Bp : c_split);
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t bfull = smem_u32(&sm->bin_full[0]), afree = smem_u32(&sm->acc_free[0]);
    const uint32_t dq_s = smem_u32(sm->dq), cs_s = smem_u32(sm->cs);
    int Vg = 0;
    for (int it = 0; it < n_it; ++it) {
      const int i = ((int)blockIdx.x + it * (int)gridDim.x) * 128 + t;
      const float rs0 = rowscale[i], rs1 = WG > 1 ? rowscale[(size_t)rs_stride + i] : 0.f;
      for (int v = 0; v < V; ++v, ++Vg) {
        if (P2_DRAIN_BY_BIN && (Vg & 1) != set) continue;
        const int k = v, buf = Vg & am;
        mbar_wait_s(bfull + 8u * (uint32_t)buf, (uint32_t)(Vg >> sh) & 1u);
        tc_fence_after();
        const bool has = sm->cnt[k] > 0;
        const uint32_t tcol = lane_base + (uint32_t)(buf * acc_stride);
        if (has) {
          // (the operands of RHE-DOM share an accumulator: twice the positions per bin)
          const bool small = sm->cnt[k] * n_modes <= (1 << 14);
          if (L == 3 && small) p2_drain<3, 1>(tcol, c_lo, c_hi, k, i, K, WG, B, Bp, (uint32_t)Np, dq_s, cs_s, rs0, rs1, P_out, S_accum);
          else if (L == 3) p2_drain<3, 0>(tcol, c_lo, c_hi, k, i, K, WG, B, Bp, (uint32_t)Np, dq_s, cs_s, rs0, rs1, P_out, S_accum);
          else if (L == 2) p2_drain<2, 0>(tcol, c_lo, c_hi, k, i, K, WG, B, Bp, (uint32_t)Np, dq_s, cs_s, rs0, rs1, P_out, S_accum);
          else p2_drain<4, 0>(tcol, c_lo, c_hi, k, i, K, WG, B, Bp, (uint32_t)Np, dq_s, cs_s, rs0, rs1, P_out, S_accum);
          tmem_st_wait();                              // the columns this warp read are zero again
        } else if (P_out) {                              // a bin without SNPs in this block: X (X^T Z) = 0
          for (int wg = 0; wg < WG; ++wg)
            for (int b = c_lo; b < min(c_hi, B); ++b) P_out[(size_t)((wg * K + k) * B + b) * (size_t)Np + i] = 0.f;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_s(afree + 8u * (uint32_t)buf);
      }
    }
  }
  __syncthreads();
  if (warp == P2_W_PROD) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols); }
}

// Individual-major copy of one block (ingest): GT[i][p] = imputed A2 count (0, 1, 2 as a 2-bit number; no missing code
// is left) of bin-sorted position p for individual i.  Word w of a row holds positions 16 w .. 16 w + 15, position
// 16 w + 4 k + j in field 4 j + k, so that register k, byte j of the expansion in k_tc_pass_b2 is position 4 k + j: the
// natural position order of the Uq tiles.  Tile = 128 positions x 512 individuals.
__global__ void __launch_bounds__(256)
k_tc_transpose(const uint8_t* __restrict__ bed, int pitch, const int32_t* __restrict__ pos_rows,
               const int32_t* __restrict__ counts, int n_kept, int binary, const double* __restrict__ uniforms,
               uint8_t* __restrict__ gt, int n_ss) {
  __shared__ uint32_t tile[128][33];
  __shared__ uint32_t outw[512][9];
  __shared__ uint8_t fcode[128];
  const int st = blockIdx.y, i0 = blockIdx.x * 512;
  for (int idx = threadIdx.x; idx < 128 * 32; idx += 256) {
    const int r = idx >> 5, w = idx & 31;
    const int row = pos_rows[st * 128 + r];
    tile[r][w] = row >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(bed + (size_t)row * pitch + (i0 >> 2)) + w) : 0u;
  }
  if (threadIdx.x < 128) {
    const int row = pos_rows[st * 128 + threadIdx.x];
    int code = 0;
    if (row >= 0) {
      const int4 c = reinterpret_cast<const int4*>(counts)[row];
      code = rhe_fill_from_counts(c.y, c.z, c.w, n_kept, binary, binary ? uniforms[row] : 0.0);   // value a missing genotype takes
    }
    fcode[threadIdx.x] = (uint8_t)code;
  }
  __syncthreads();
  for (int it = 0; it < 16; ++it) {
    const int idx = threadIdx.x + 256 * it;            // lanes = consecutive individuals, warp-uniform position group
    const int i = idx & 511, pg = idx >> 9;
    const int wsrc = i >> 4, sh = 2 * (i & 15);
    uint32_t out = 0;
#pragma unroll
    for (int f = 0; f < 16; ++f) {
      const int p = 16 * pg + 4 * (f & 3) + (f >> 2);
      const uint32_t code = (tile[p][wsrc] >> sh) & 3u;            // PLINK code: 00 -> 0, 01 -> missing, 10 -> 1, 11 -> 2
      out |= (uint32_t)rhe_code_value(code, fcode[p]) << (2 * f);
    }
    outw[i][pg] = out;
  }
  __syncthreads();
  for (int it = 0; it < 16; ++it) {
    const int idx = threadIdx.x + 256 * it;            // 8 consecutive lanes = one individual's 32-byte sector
    const int i = idx >> 3, pg = idx & 7;
    // box (M-tile, super-stage) = 128 individuals x 128 B, contiguous: [mtile][ss][row][q * 32 + pg * 4]
    const size_t box = (size_t)((i0 + i) >> 7) * n_ss + (size_t)(st >> 2);
    *reinterpret_cast<uint32_t*>(gt + (box * 128 + (size_t)((i0 + i) & 127)) * 128 + (size_t)((st & 3) * 32 + pg * 4)) = outw[i][pg];
  }
}

// Re-tile + re-encode the SNP-major rows of one block for pass A (ingest, rhe_block_retile): source = the PLINK rows
// [m][pitch]; destination (a scratch buffer, copied back over the rows afterwards) = [row tile][column of 128 B][128 rows]
// [128 B], the two bits of a genotype holding the imputed A2 count (missing -> the SNP's fill), fields permuted inside
// every word so that the mask-and-shift expansion of k_tc_pass_a<., 1> delivers the individuals in the order the table
// look-up does (byte 4 k + j of the expansion = individual tc_perm16(4 k + j) of the word's sixteen).
__global__ void __launch_bounds__(256)
k_tc_retile(const uint8_t* __restrict__ bed, int pitch, int m, const int32_t* __restrict__ counts, int n_kept, int binary,
            const double* __restrict__ uniforms, uint8_t* __restrict__ out) {
  const int n_cc = pitch >> 7;                         // 128-byte columns per row
  const int tile = blockIdx.y, cc0 = blockIdx.x * 8;   // 8 columns (1 KB of every row) per CTA
  __shared__ uint8_t fillv[128];
  if (threadIdx.x < 128) {
    const int row = tile * 128 + (int)threadIdx.x;
    int f = 0;
    if (row < m) {
      const int4 c = reinterpret_cast<const int4*>(counts)[row];
      f = rhe_fill_from_counts(c.y, c.z, c.w, n_kept, binary, binary ? uniforms[row] : 0.0);
    }
    fillv[threadIdx.x] = (uint8_t)f;
  }
  __syncthreads();
  // word index inside the CTA's piece: [128 rows][8 columns][32 words]; a warp reads 128 contiguous bytes of one row
  for (int idx = threadIdx.x; idx < 128 * 8 * 32; idx += 256) {
    const int w = idx & 31, c = (idx >> 5) & 7, r = idx >> 8;
    const int row = tile * 128 + r, cc = cc0 + c;
    if (cc >= n_cc) continue;
    uint32_t v = 0u;
    if (row < m) {
      const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(bed + (size_t)row * pitch + (size_t)cc * 128) + w);
      // PLINK code -> count, sixteen fields at once: 00 -> 0, 10 -> 1, 11 -> 2, 01 (missing) -> fill
      const uint32_t b0 = x & 0x55555555u, b1 = (x >> 1) & 0x55555555u;
      const uint32_t miss = b0 & ~b1, f = fillv[r];
      const uint32_t val = b1 + (b1 & b0) + (f == 1u ? miss : (f == 2u ? miss << 1 : 0u));
#pragma unroll
      for (int nf = 0; nf < 16; ++nf) {                // new field 4 j + k <- raw field tc_perm16(4 k + j)
        const int j = nf >> 2, k = nf & 3;
        v |= ((val >> (2 * tc_perm16(4 * k + j))) & 3u) << (2 * nf);
      }
    }
    reinterpret_cast<uint32_t*>(out + (((size_t)tile * n_cc + cc) * 128 + r) * 128)[w] = v;
  }
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
k_tc_quant_rhs(const float* __restrict__ rhs, int Np, int Rc, int R1p, int NB, int L, int F, int8_t* __restrict__ rq,
               double* __restrict__ col_dq) {
  // column c = chunk ch, local column cl: limb l at row ch * NB + l * R1p + cl of rq
  const int c = blockIdx.x, ch = c / Rc, cl = c - ch * Rc;
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
    tc_limbs(q, L, rq + (size_t)(ch * NB + cl) * Np + pos, (size_t)R1p * Np);
  }
}

__global__ void k_tc_positions(const int32_t* __restrict__ bin_rows, const int32_t* __restrict__ bin_off, int k0,
                               const int32_t* __restrict__ pstart, int32_t* __restrict__ pos_rows) {
  const int kl = blockIdx.y, k = k0 + kl;            // local / global bin
  const int n = bin_off[k + 1] - bin_off[k], p0 = pstart[kl], span = pstart[kl + 1] - p0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < span; i += gridDim.x * blockDim.x)
    pos_rows[p0 + i] = i < n ? bin_rows[bin_off[k] + i] : -1;
}

__global__ void k_tc_wmax(const float* __restrict__ w1, int m, int B, unsigned int* __restrict__ wmax) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * B) return;
  atomicMax(wmax + (idx % B), __float_as_uint(fabsf(w1[idx])));
}

// Quantise the pass-B weights of one block into int8 limbs, in bin-sorted position order, and emit the
// per-position decode metadata (SNP row | fill << 24 | mode << 26).  mode 0 multiplies the count operand
// (weight w1, with w2 = 2 w1 + extra); mode 1 multiplies the [g == 2] indicator operand (weight extra = w2 - 2 w1,
// non-zero only for the dominance group) and occupies positions [n_pos, 2 n_pos).
__global__ void k_tc_quant_w(const float* __restrict__ w1, const float* __restrict__ w2, const int32_t* __restrict__ pos_rows,
                             int n_pos, int cap_pos, int m, int WG, int n_modes, int B, int Bs, int b0, int Bp, int L, int F,
                             const unsigned int* __restrict__ wmax, int8_t* __restrict__ uq,
                             const uint8_t* __restrict__ fill, int32_t* __restrict__ meta_v, int SI, int n_chunk) {
  // B columns [b0, b0 + B) of the Bs vector columns
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pos * B * WG * n_modes) return;
  const int p = idx % n_pos, b = (idx / n_pos) % B, wg = (idx / (n_pos * B)) % WG, mode = idx / (n_pos * B * WG);
  const int row = pos_rows[p];
  if (b == 0 && wg == 0) {
    // stage st = gp / 128 belongs to decode group par = st % SI as its own stage u = st / SI: meta_v[par][u / 4][t][u % 4]
    const int gp = mode * n_pos + p, st = gp >> 7, t = gp & 127, par = st % SI, u = st / SI;
    meta_v[((((size_t)par * n_chunk + (u >> 2)) * 128 + t) << 2) + (u & 3)] =
        row >= 0 ? (row | ((int)fill[row] << 24) | (mode << 26)) : -1;
  }
  const int e = (int)((wmax[wg * Bs + b0 + b] >> 23) & 255u) - 126;   // one fixed-point scale per (weight group, column)
  long long q = 0ll;
  if (row >= 0) {
    const size_t o = ((size_t)wg * m + row) * Bs + b0 + b;
    const double w = mode == 0 ? (double)w1[o] : (double)w2[o] - 2.0 * (double)w1[o];
    q = llrint(ldexp(w, F - e));
  }
  tc_limbs(q, L, uq + ((size_t)wg * L * Bp + b) * cap_pos + mode * n_pos + p, (size_t)Bp * cap_pos);
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
static inline uint32_t pow2_cols(int n) { uint32_t c = 32; while ((int)c < n) c <<= 1; return c; }
static inline int pa_smem_bytes(int nb, int gs) { return PA_RS * 4 * nb * 128 + gs * PA_PACKED + (int)sizeof(PaSmem); }
// Depth of pass A's genotype ring: the decode warps stall on HBM latency with fewer than four boxes in flight per CTA.
// Two co-resident CTAs when four slots fit half an SM's shared memory (three as a last resort), else one CTA with as
// deep a ring as fits.
static inline int pa_ring(int nb) {
  const int half = (233472 - 2048) / 2, whole = 232448 - 1024;
  if (pa_smem_bytes(nb, 4) <= half) return 4;
  if (pa_smem_bytes(nb, 3) <= half) return 3;
  if (pa_smem_bytes(nb, 8) <= whole) return 8;
  if (pa_smem_bytes(nb, 6) <= whole) return 6;
  if (pa_smem_bytes(nb, 4) <= whole) return 4;
  return 3;
}
static inline int pb_smem_bytes(int nc, int bs, int G = PB_G) { return G * PB_AS * TC_TILE_A + bs * nc * 128 + 4 * G * PB_PKG * 1024 + (int)sizeof(PbSmem) + 1024; }
// Uq ring: `bs` tile slots in batches of 2^bzsh tiles that share one barrier pair.  Either a batch spans at least
// one stage of every issuer (2^bzsh >= SI = G / MT), or there is no batching and bs is a multiple of SI (a slot is
// then always consumed by the same issuer); both keep every waiter within one mbarrier phase of its barrier.
static inline int pb_budget(int G) { return G == 2 ? 233472 / 2 - 1024 : 232448 - 2048; }
static inline int pb_ring(int nc, int mt, int G, int* bzsh) {
  const int si = G / mt;
  const int budget = pb_budget(G);
  if (RHE_DBG_ENV("PYRHE_TC_DEBUG_RING", 0) != 1) {        // "1" (profiling build only): the unbatched ring
    if (si <= 4 && pb_smem_bytes(nc, 8, G) <= budget) { *bzsh = 2; return 8; }    // two batches of four tiles
    if (si <= 2 && pb_smem_bytes(nc, 4, G) <= budget) { *bzsh = 1; return 4; }    // two batches of two tiles
  }
  *bzsh = 0;
  int bs = PB_BS / si * si;
  while (bs > si && pb_smem_bytes(nc, bs, G) > budget) bs -= si;
  return bs;
}

struct P2Shape;
static bool p2_shape(const TcState* s, P2Shape* o);
static int p2_smem_of(const TcState* s);

// Shape plan of the tensor kernels for one configuration (pure host arithmetic: shared by rhe_tc_supported and
// rhe_tc_create so that the Python side never has to mirror the limits).
struct TcShape { int L, F, Rc, R1p, NBa, Bc, Bp, NCb, MT, G, KG; };

static int tc_shape(const rhe_config& g, TcShape* o, int quiet) {
  const int R1 = g.n_sets * g.n_cols_set, n_groups = g.n_ops * g.n_sets;
  const char* envL = getenv("PYRHE_B200_LIMBS");        // precision knob, read once per context
  o->L = envL ? atoi(envL) : 3;
  if (o->L < 2 || o->L > 4) { if (!quiet) rhe_set_error("PYRHE_B200_LIMBS must be 2..4"); return RHE_ERR_INVALID; }
  o->F = 8 * o->L - 2;
  // pass A takes all RHS columns in one launch when their limb columns fit one MMA (N <= 256) and the Rq ring fits the
  // shared memory; otherwise in balanced chunks of Rc columns, each a pass of its own over the block
  for (int n_ch = 1;; ++n_ch) {
    o->Rc = round_up(rhe_div_up(R1, n_ch), 4);
    o->R1p = o->Rc;
    o->NBa = round_up(o->L * o->R1p, 16);
    if ((o->NBa <= 256 && pa_smem_bytes(o->NBa, 3) <= 232448 - 1024) || o->Rc <= 4) break;
  }
  // pass B runs over all vector columns at once when they fit -- groups x vectors <= 64 (shared-memory tables), MMA
  // N <= 256, the Uq ring within the shared memory the decoded tiles leave -- otherwise in chunks of Bc columns, each a
  // pass of its own over the block's rows (e.g. RHE-DOM / GENIE with the reference's 50 random vectors).
  auto plan_b = [&](int bc) {
    o->Bc = bc;
    o->Bp = round_up(o->Bc, 2);
    o->NCb = round_up(n_groups * o->L * o->Bp, 16);   // weight groups (RHS sets) are stacked along N
    o->G = PB_G;
    // M-tiles per CTA and bins per launch: everything in one launch when the accumulators fit (two tiles if possible),
    // otherwise bin groups of two-tile CTAs.
    if (g.n_bins * 2 * o->NCb <= 512) { o->MT = 2; o->KG = g.n_bins; }
    else if (g.n_bins * o->NCb <= 512) { o->MT = 1; o->KG = g.n_bins; }
    else if (2 * o->NCb <= 512) { o->MT = 2; o->KG = 512 / (2 * o->NCb); }
    else { o->MT = 1; o->KG = o->NCb <= 512 ? 512 / o->NCb : 0; }
    if (o->KG > 0 && n_groups * o->KG * o->Bc > PB_MAX_KB) o->KG = PB_MAX_KB / (n_groups * o->Bc);   // per-(bin, column) mean terms staged in shared memory
    return o->NCb <= 256 && o->KG >= 1 && pb_smem_bytes(o->NCb, o->G / o->MT, o->G) <= pb_budget(o->G);
  };
  int bc = n_groups * g.n_vec <= 64 ? g.n_vec : (64 / n_groups) & ~1;
  while (bc > 2 && !plan_b(bc)) bc = (bc - 1) & ~1;
  plan_b(bc);
  {
    // Optional variant (PYRHE_B200_PASSB_GROUPS=2): two half-size CTAs per SM (one M-tile, two decode groups, 256 TMEM
    // columns each), so that the epilogue of one overlaps the main loop of the other.  Measured equal to the default
    // on config 5, so it stays opt-in.
    const char* envG = getenv("PYRHE_B200_PASSB_GROUPS");
    const bool fits2 = o->Bc == g.n_vec && g.n_bins * o->NCb <= 256 && pb_smem_bytes(o->NCb, 2, 2) <= pb_budget(2);
    if (fits2 && envG && atoi(envG) == 2) { o->G = 2; o->MT = 1; o->KG = g.n_bins; }
  }
  if (pb_smem_bytes(o->NCb, o->G / o->MT, o->G) > pb_budget(o->G) || o->NBa > 256 || pa_smem_bytes(o->NBa, 3) > 232448 - 1024 ||
      o->KG < 1 || o->KG > 255 ||
      n_groups * o->KG * o->Bc > PB_MAX_KB || o->Bc < 1) {
    if (!quiet)
      rhe_set_error("RHE_PATH_TCGEN05: %d RHS columns / %d bins x %d vectors x %d weight groups exceed the tensor kernels' "
                    "TMEM / shared-memory layout (pass A: %d limb columns <= 256 and its Rq ring within 227 KB; pass B: "
                    "%d limb columns per bin <= 512)",
                    R1, g.n_bins, g.n_vec, n_groups, o->NBa, o->NCb);
    return RHE_ERR_UNSUPPORTED;
  }
  return RHE_OK;
}

int rhe_tc_check(const rhe_config* cfg, int quiet) {
  TcShape sh;
  return tc_shape(*cfg, &sh, quiet);
}

// (Re)allocate the per-block pass-B staging (quantised weights, decode metadata) for `need` positions.  Called from the
// set-up entry points only (context / plan creation): the device is idle-synchronised before buffers are replaced.
static int tc_reserve_positions(rhe_ctx* c, TcState* s, int need) {
  if (need <= s->cap_pos) return RHE_OK;
  RHE_CUDA(cudaDeviceSynchronize());
  if (s->uq) cudaFree(s->uq);
  if (s->pos_meta) cudaFree(s->pos_meta);
  s->uq = nullptr;
  s->pos_meta = nullptr;
  s->cap_pos = round_up(need, 128);
  RHE_CUDA(cudaMalloc((void**)&s->pos_meta, sizeof(int32_t) * (s->cap_pos + 8 * 512)));   // + one padded chunk per group
  RHE_CUDA(cudaMalloc((void**)&s->uq, (size_t)s->NCb * s->cap_pos));
  RHE_CUDA(cudaMemset(s->uq, 0, (size_t)s->NCb * s->cap_pos));
  return tc_encode_2d(s, &s->tm_uq, s->uq, (uint64_t)s->cap_pos, (uint64_t)s->NCb, (uint32_t)s->NCb);
}

int rhe_tc_create(rhe_ctx* c) {
  const rhe_config& g = c->cfg;
  TcShape sh;
  int rc = tc_shape(g, &sh, 0);
  if (rc) return rc;
  TcState* s = new TcState();
  s->L = sh.L; s->F = sh.F; s->Rc = sh.Rc; s->R1p = sh.R1p; s->NBa = sh.NBa; s->Bc = sh.Bc; s->Bp = sh.Bp; s->NCb = sh.NCb; s->MT = sh.MT; s->G = sh.G; s->KG = sh.KG;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    delete s;
    rhe_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return RHE_ERR_CUDA;
  }
  s->encode = (PFN_encodeTiled)fn;
  cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, c->cfg.device);
  s->n_ops = c->cfg.n_ops;
  c->tc = s;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes); } };
  alloc((void**)&s->rq, (size_t)s->NBa * rhe_div_up(c->R1, s->Rc) * c->Np);
  alloc((void**)&s->col_dq, sizeof(double) * c->R1);
  alloc((void**)&s->wmax, sizeof(unsigned int) * c->n_groups * g.n_vec);
  if (e != cudaSuccess) { rhe_set_error("tensor-core workspace allocation failed: %s", cudaGetErrorString(e)); return RHE_ERR_CUDA; }
  rc = tc_encode_2d(s, &s->tm_rq, s->rq, (uint64_t)c->Np, (uint64_t)s->NBa * (uint64_t)rhe_div_up(c->R1, s->Rc), (uint32_t)s->NBa);
  if (rc) return rc;
  // pass-B staging for the largest block of a one-bin-per-SNP annotation (every bin padded to 128 rows); plans with
  // overlapping annotations grow it at plan creation
  const int kg = s->KG < g.n_bins ? s->KG : g.n_bins;
  rc = tc_reserve_positions(c, s, g.n_ops * (round_up(g.max_block_snps, 128) + 128 * kg + 512));
  if (rc) return rc;
  {
    const int gs = pa_ring(s->NBa), smem = pa_smem_bytes(s->NBa, gs);
#define PA_ATTR(GS_)                                                                                            \
    do {                                                                                                         \
      RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_a<GS_, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
      RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_a<GS_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
    } while (0)
    if (gs == 8) PA_ATTR(8);
    else if (gs == 6) PA_ATTR(6);
    else if (gs == 4) PA_ATTR(4);
    else PA_ATTR(3);
#undef PA_ATTR
  }
  {
    int shb;
    const int smem = pb_smem_bytes(s->NCb, pb_ring(s->NCb, s->MT, s->G, &shb), s->G);
    if (s->G == 2) RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else if (s->MT == 2) RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  if (p2_smem_of(s) > 0) {
    RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b2<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2_smem_of(s)));
    RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b2<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2_smem_of(s)));
    RHE_CUDA(cudaFuncSetAttribute(k_tc_pass_b2<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2_smem_of(s)));
  }
  return RHE_OK;
}

void rhe_tc_destroy(rhe_ctx* c) {
  TcState* s = (TcState*)c->tc;
  if (!s) return;
  void* ptrs[] = {s->rq, s->col_dq, s->uq, s->wmax, s->pos_meta};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete s;
  c->tc = nullptr;
}

int rhe_tc_set_rhs(rhe_ctx* c, cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  k_tc_quant_rhs<<<c->R1, 256, 0, st>>>(c->rhs, c->Np, s->Rc, s->R1p, s->NBa, s->L, s->F, s->rq, s->col_dq);
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

unsigned int* rhe_tc_wmax(rhe_ctx* c) { return c->tc ? ((TcState*)c->tc)->wmax : nullptr; }

int64_t rhe_tc_tiled_bytes(const rhe_ctx* c, int m) {
  if (TC_PA_SHIFT_VARIANT == 2 && c->Np >= (1 << 21)) return 0;      // the x4 operand values need 8 x 128 x N < 2^31
  return (int64_t)rhe_div_up(m, 128) * 128 * c->cfg.pitch_bytes;
}

int rhe_tc_retile(rhe_ctx* c, uint8_t* bed, int m, const int32_t* counts, uint8_t* scratch, cudaStream_t st) {
  const rhe_config& g = c->cfg;
  if (rhe_tc_tiled_bytes(c, m) == 0) { rhe_set_error("rhe_block_retile: not available for %d individuals (see rhe_block_tiled_bytes)", c->Np); return RHE_ERR_UNSUPPORTED; }
  if (g.impute_binary && (!c->uniforms || c->n_uniforms < m)) { rhe_set_error("binary imputation needs rhe_set_uniforms first"); return RHE_ERR_STATE; }
  const int n_cc = g.pitch_bytes / 128;
  k_tc_retile<<<dim3(rhe_div_up(n_cc, 8), rhe_div_up(m, 128)), 256, 0, st>>>(bed, g.pitch_bytes, m, counts, g.n_kept,
                                                                               g.impute_binary, c->uniforms, scratch);
  RHE_LAUNCH_CHECK(c);
  RHE_CUDA(cudaMemcpyAsync(bed, scratch, (size_t)rhe_tc_tiled_bytes(c, m), cudaMemcpyDeviceToDevice, st));
  return RHE_OK;
}

int rhe_tc_pass_a(rhe_ctx* c, const uint8_t* bed, int m, int tiled, cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  const int tiles = rhe_div_up(m, 128);
  // Split the individuals so that the grid fills whole waves of 2 resident CTAs per SM (a nearly empty last wave
  // costs a full CTA time): pick the split count with the best wave efficiency, keeping at least 64 sub-tiles
  // (8192 individuals) per CTA so that prologue / epilogue stay amortised.
  const int slots = 148 * 2;
  int splits = 1;
  double best = 0.0;
  for (int cand = 1; cand <= 16; ++cand) {
    if (cand > 1 && rhe_div_up(c->Np, cand) < 8192) break;
    const int ctas = tiles * cand;
    const double eff = (double)ctas / ((double)rhe_div_up(ctas, slots) * slots);
    if (eff > best + 0.02) { best = eff; splits = cand; }
  }
  if (splits > c->Np / 512) splits = c->Np / 512 > 0 ? c->Np / 512 : 1;
  const uint32_t col_a = (uint32_t)round_up(s->NBa, 32);
  CUtensorMap tm_bed;                                  // the block's packed rows as a 2-D byte tensor [m][pitch], or its tiles
  int rc = tiled ? tc_encode_2d(s, &tm_bed, const_cast<uint8_t*>(bed), 128, (uint64_t)tiles * 128 * (uint64_t)(c->cfg.pitch_bytes / 128), 128)
                 : tc_encode_2d(s, &tm_bed, const_cast<uint8_t*>(bed), (uint64_t)c->cfg.pitch_bytes, (uint64_t)m, 128);
  if (rc) return rc;
  const int dbg = RHE_DBG_ENV("PYRHE_TC_DEBUG_SKIPA", 0);
  const int gs = pa_ring(s->NBa), smem = pa_smem_bytes(s->NBa, gs);
  const uint32_t cols = pow2_cols((int)col_a + 32 * PA_AS);
  for (int mode = 0; mode < c->cfg.n_ops; ++mode) {      // RHE-DOM: the same pass over the [g == 2] indicator operand
    double* t_out = c->t_raw + (size_t)mode * m * c->R1;
    for (int ch = 0; ch * s->Rc < c->R1; ++ch) {          // one pass per chunk of RHS columns (a single one as a rule)
      const int c_lo = ch * s->Rc, rv = c->R1 - c_lo < s->Rc ? c->R1 - c_lo : s->Rc;
#define PA_LAUNCH(GS_, T_)                                                                                                 \
      k_tc_pass_a<GS_, T_><<<dim3(splits, tiles), PA_THREADS, smem, st>>>(s->tm_rq, tm_bed, m, c->Np, s->NBa, c->R1, s->R1p, \
          s->L, c_lo, rv, ch * s->NBa, c->fill, s->col_dq, t_out, cols, col_a, mode, mode ? 0 : dbg)
      if (tiled) { if (gs == 8) PA_LAUNCH(8, 1); else if (gs == 6) PA_LAUNCH(6, 1); else if (gs == 4) PA_LAUNCH(4, 1); else PA_LAUNCH(3, 1); }
      else { if (gs == 8) PA_LAUNCH(8, 0); else if (gs == 6) PA_LAUNCH(6, 0); else if (gs == 4) PA_LAUNCH(4, 0); else PA_LAUNCH(3, 0); }
#undef PA_LAUNCH
      RHE_LAUNCH_CHECK(c);
    }
  }
  return RHE_OK;
}

// Bin-sorted positions of the bins [k0, k0 + kn) of a block (bins numbered locally 0 .. kn - 1 inside the group).
// Set-up path: allocates, copies from host staging vectors and synchronises before they die.
static int tc_block_meta(rhe_ctx* c, const rhe_block_plan* plan, int k0, int kn, cudaStream_t st, TcBlockMeta* out) {
  const int m = plan->m;
  const int32_t* bin_off_host = plan->off_host.data();
  std::vector<int32_t> pstart(kn + 1, 0), counts(kn), info;
  for (int kl = 0; kl < kn; ++kl) {
    counts[kl] = bin_off_host[k0 + kl + 1] - bin_off_host[k0 + kl];
    pstart[kl + 1] = pstart[kl] + round_up(counts[kl], 128);
    for (int done = 0; done < counts[kl]; done += 128) {
      const int left = counts[kl] - done;
      info.push_back(kl | ((done == 0) << 8) | (((left < 128 ? left : 128) + 31) / 32) << 16);
    }
  }
  TcBlockMeta& b = *out;
  b.k0 = k0;
  b.kn = kn;
  // the list ends on a super-stage boundary (512 positions = one 128-byte TMA box of the individual-major copy): up
  // to three empty tail stages (no rows, no K-steps) that belong to the last bin
  b.n_pos = round_up(pstart[kn], 512);
  for (int p = pstart[kn]; p < b.n_pos; p += 128) info.push_back((kn - 1) | (0 << 8) | (0 << 16));
  const int n_alloc = b.n_pos > 0 ? b.n_pos : 128;
  RHE_CUDA(cudaMalloc((void**)&b.pos_rows, sizeof(int32_t) * n_alloc));
  RHE_CUDA(cudaMemsetAsync(b.pos_rows, 0xFF, sizeof(int32_t) * n_alloc, st));      // -1 = padding
  const int n_modes = c->cfg.n_ops;                  // RHE-DOM runs the position list twice (count, then [g == 2] operand)
  {
    const size_t one = info.size();
    for (int mo = 1; mo < n_modes; ++mo) for (size_t i = 0; i < one; ++i) info.push_back(info[i]);
  }
  RHE_CUDA(cudaMalloc((void**)&b.stage_info, sizeof(int32_t) * (n_alloc / 128) * n_modes));
  RHE_CUDA(cudaMalloc((void**)&b.bin_count, sizeof(int32_t) * kn));
  int32_t* d_pstart = nullptr;
  RHE_CUDA(cudaMalloc((void**)&d_pstart, sizeof(int32_t) * (kn + 1)));
  cudaError_t e = cudaMemcpyAsync(d_pstart, pstart.data(), sizeof(int32_t) * (kn + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b.bin_count, counts.data(), sizeof(int32_t) * kn, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && !info.empty())
    e = cudaMemcpyAsync(b.stage_info, info.data(), sizeof(int32_t) * info.size(), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    k_tc_positions<<<dim3(rhe_div_up(m, 256), kn), 256, 0, st>>>(plan->bin_rows, plan->off_dev, k0, d_pstart, b.pos_rows);
    c->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);     // the host staging vectors die here
  cudaFree(d_pstart);
  if (e != cudaSuccess) { rhe_set_error("tc_block_meta: %s", cudaGetErrorString(e)); return RHE_ERR_CUDA; }
  return RHE_OK;
}

int rhe_tc_plan_create(rhe_ctx* c, rhe_block_plan* plan, cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  TcPlan* tp = new TcPlan();
  plan->tc = tp;
  const int K = c->cfg.n_bins;
  int need = 0;
  for (int k0 = 0; k0 < K; k0 += s->KG) {
    const int kn = K - k0 < s->KG ? K - k0 : s->KG;
    tp->groups.emplace_back();
    int rc = tc_block_meta(c, plan, k0, kn, st, &tp->groups.back());
    if (rc) return rc;
    if (c->cfg.n_ops * tp->groups.back().n_pos > need) need = c->cfg.n_ops * tp->groups.back().n_pos;
  }
  return tc_reserve_positions(c, s, need);
}

void rhe_tc_plan_destroy(rhe_block_plan* plan) {
  TcPlan* tp = (TcPlan*)plan->tc;
  if (!tp) return;
  for (TcBlockMeta& b : tp->groups) {
    if (b.pos_rows) cudaFree(b.pos_rows);
    if (b.stage_info) cudaFree(b.stage_info);
    if (b.bin_count) cudaFree(b.bin_count);
  }
  delete tp;
  plan->tc = nullptr;
}

// One launch per bin group of at most KG bins (the accumulators of KG bins x MT tiles fill the TMEM allocation); every
// group walks only the bin-sorted rows of its own bins, so the block is still read once in total.
static int tc_pass_b_group(rhe_ctx* c, TcState* s, const uint8_t* bed, int m, const TcBlockMeta* meta, float* P_out,
                           float* S_accum, cudaStream_t st) {
  const rhe_config& g = c->cfg;
  const int K = g.n_bins, B = g.n_vec, k0 = meta->k0, kn = meta->kn;
  const int n_pos = meta->n_pos, n_modes = g.n_ops;
  if (n_modes * n_pos > s->cap_pos) { rhe_set_error("pass B staging smaller than the plan (plan of another context?)"); return RHE_ERR_STATE; }
  const int n_stage = n_modes * n_pos / 128, SI = s->G / s->MT;
  const int n_chunk = rhe_div_up(rhe_div_up(n_stage, SI), 4);       // chunks of four own stages per decode group
  const uint32_t cols = pow2_cols(kn * s->MT * s->NCb);
  int bzsh = 0;
  const int bs = pb_ring(s->NCb, s->MT, s->G, &bzsh), smem = pb_smem_bytes(s->NCb, bs, s->G);
  const int a_major = RHE_DBG_ENV("PYRHE_TC_DEBUG_KMAJOR", 0) ? 0 : 1;
  const int kcap = RHE_DBG_ENV("PYRHE_TC_DEBUG_KSTEPS", 4);
  const int dbg = RHE_DBG_ENV("PYRHE_TC_DEBUG_SKIP", 0);
  const int rs_stride = g.n_sets == 2 ? c->Np : 0;
  for (int b0 = 0; b0 < B; b0 += s->Bc) {              // one pass per chunk of vector columns (a single one as a rule)
    const int bc = B - b0 < s->Bc ? B - b0 : s->Bc;
    if (n_pos > 0) {
      k_tc_quant_w<<<rhe_div_up((int64_t)n_pos * bc * c->n_groups * n_modes, 256), 256, 0, st>>>(
          c->w1, c->w2, meta->pos_rows, n_pos, s->cap_pos, m, c->n_groups, n_modes, bc, B, b0, s->Bp, s->L, s->F, s->wmax, s->uq,
          c->fill, s->pos_meta, SI, n_chunk);
      RHE_LAUNCH_CHECK(c);
    }
#define PB_LAUNCH(MT_, G_)                                                                                                   \
    k_tc_pass_b<MT_, G_><<<c->Np / (128 * MT_), PB_THREADS_OF(G_), smem, st>>>(                                              \
        s->tm_uq, bed, g.pitch_bytes, c->Np, n_stage, s->pos_meta, n_chunk, meta->stage_info, meta->bin_count, kn, K, k0,    \
        c->n_groups, bc, B, b0, s->Bp, s->L, s->NCb, s->F, s->wmax, c->cs, c->rowscale, rs_stride, P_out, S_accum, cols,     \
        a_major, kcap, bs, bzsh, dbg)
    if (s->G == 2) PB_LAUNCH(1, 2);
    else if (s->MT == 2) PB_LAUNCH(2, 4);
    else PB_LAUNCH(1, 4);
#undef PB_LAUNCH
    RHE_LAUNCH_CHECK(c);
  }
  return RHE_OK;
}

// ---- individual-major fast path (k_tc_pass_b2): shapes, ingest, launch
struct P2Shape { int n_acc, gs, us, smem; uint32_t col_a, cols; };

static bool p2_shape(const TcState* s, P2Shape* o) {
  if (s->NCb > 96) return false;
  o->n_acc = 4 * round_up(s->NCb, 32) + 32 * P2_AS <= 512 ? 4 : 2;   // bins in flight between the issuers and the drain warps
  o->col_a = (uint32_t)(o->n_acc * round_up(s->NCb, 32));
  o->cols = pow2_cols((int)o->col_a + 32 * P2_AS);
  if (o->cols > 512) return false;
  const int budget = 232448 - 2048;                    // one CTA per SM
  const int rings[3][2] = {{8, 4}, {6, 3}, {4, 2}};    // depth of the genotype-box ring / of the Uq super-stage ring
  for (const auto& r : rings) {
    const int smem = r[0] * PA_PACKED + r[1] * 4 * s->n_ops * s->NCb * 128 + (int)sizeof(P2Smem) + 1024;
    if (smem <= budget) { o->gs = r[0]; o->us = r[1]; o->smem = smem; return true; }
  }
  return false;
}

static int p2_smem_of(const TcState* s) {
  P2Shape sh;
  return p2_shape(s, &sh) ? sh.smem : 0;
}

// Bytes of the individual-major copy of the plan's block (0: this configuration / plan has no fast path and pass B
// gathers from the SNP-major rows).  The copy exists only for plans whose bins form one group.
int64_t rhe_tc_gt_bytes(const rhe_ctx* c, const rhe_block_plan* plan) {
  const TcState* s = (const TcState*)c->tc;
  const TcPlan* tp = (const TcPlan*)plan->tc;
  P2Shape sh;
  if (!s || !tp || tp->groups.size() != 1 || tp->groups[0].n_pos <= 0 || !p2_shape(s, &sh)) return 0;
  if (s->Bc < c->cfg.n_vec) return 0;                                                                   // pass B runs in column chunks
  if ((int64_t)c->n_groups * c->cfg.n_bins * c->cfg.n_vec * (int64_t)c->Np >= (1ll << 31)) return 0;   // 32-bit indices into P / S
  if (c->cfg.n_ops * (tp->groups[0].n_pos / 128) > P2_MAXSUB) return 0;                                 // per-sub-tile bin table in shared memory
  return (int64_t)c->Np * (tp->groups[0].n_pos / 4);
}

int rhe_tc_transpose(rhe_ctx* c, const uint8_t* bed, const rhe_block_plan* plan, const int32_t* counts, uint8_t* gt, cudaStream_t st) {
  const TcPlan* tp = (const TcPlan*)plan->tc;
  if (rhe_tc_gt_bytes(c, plan) == 0) { rhe_set_error("rhe_block_transpose: this plan has no individual-major fast path"); return RHE_ERR_UNSUPPORTED; }
  const TcBlockMeta& meta = tp->groups[0];
  const rhe_config& g = c->cfg;
  if (g.impute_binary && (!c->uniforms || c->n_uniforms < plan->m)) { rhe_set_error("binary imputation needs rhe_set_uniforms first"); return RHE_ERR_STATE; }
  k_tc_transpose<<<dim3(c->Np / 512, meta.n_pos / 128), 256, 0, st>>>(bed, g.pitch_bytes, meta.pos_rows, counts, g.n_kept,
                                                                     g.impute_binary, c->uniforms, gt, meta.n_pos / 512);
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

static int tc_pass_b2(rhe_ctx* c, TcState* s, const uint8_t* gt, int m, const TcBlockMeta* meta, float* P_out, float* S_accum,
                      cudaStream_t st) {
  const rhe_config& g = c->cfg;
  const int K = g.n_bins, B = g.n_vec, n_pos = meta->n_pos, n_modes = g.n_ops;
  P2Shape sh;
  if (!p2_shape(s, &sh)) { rhe_set_error("individual-major pass B: unsupported shape"); return RHE_ERR_UNSUPPORTED; }
  if (n_modes * n_pos > s->cap_pos) { rhe_set_error("pass B staging smaller than the plan (plan of another context?)"); return RHE_ERR_STATE; }
  // the quantised weights in bin-sorted position order: the same kernel and buffers as the gather path (its per-group
  // decode metadata is written as well and simply not read)
  const int n_stage = n_modes * n_pos / 128, SI = s->G / s->MT;
  const int n_chunk = rhe_div_up(rhe_div_up(n_stage, SI), 4);
  k_tc_quant_w<<<rhe_div_up((int64_t)n_pos * B * c->n_groups * n_modes, 256), 256, 0, st>>>(
      c->w1, c->w2, meta->pos_rows, n_pos, s->cap_pos, m, c->n_groups, n_modes, B, B, 0, s->Bp, s->L, s->F, s->wmax, s->uq, c->fill,
      s->pos_meta, SI, n_chunk);
  RHE_LAUNCH_CHECK(c);
  CUtensorMap tm_gt;                                   // GT as a 2-D byte tensor [(Np / 128) n_ss 128 rows][128 B]: one box = 16 KB contiguous
  int rc = tc_encode_2d(s, &tm_gt, const_cast<uint8_t*>(gt), 128, (uint64_t)c->Np * (uint64_t)(n_pos / 512), 128);
  if (rc) return rc;
  const int rs_stride = g.n_sets == 2 ? c->Np : 0;
  const int n_mt = c->Np / 128, grid = n_mt < s->n_sm ? n_mt : s->n_sm;      // persistent: one CTA per SM
  const int dbg2 = RHE_DBG_ENV("PYRHE_TC_DEBUG_SKIPB2", 0);
#define P2_LAUNCH(GS_, US_)                                                                                                 \
  k_tc_pass_b2<GS_, US_><<<grid, P2_THREADS, sh.smem, st>>>(                                                               \
      s->tm_uq, tm_gt, n_mt, c->Np, n_pos / 512, n_modes, meta->stage_info, meta->bin_count, K, c->n_groups, B, s->Bp, s->L, \
      s->NCb, s->F, s->wmax, c->cs, c->rowscale, rs_stride, P_out, S_accum, sh.cols, sh.col_a, sh.n_acc, dbg2)
  if (sh.gs == 8) P2_LAUNCH(8, 4);
  else if (sh.gs == 6) P2_LAUNCH(6, 3);
  else P2_LAUNCH(4, 2);
#undef P2_LAUNCH
  RHE_LAUNCH_CHECK(c);
  return RHE_OK;
}

int rhe_tc_pass_b(rhe_ctx* c, const uint8_t* bed, const uint8_t* gt, const rhe_block_plan* plan, float* P_out, float* S_accum,
                  cudaStream_t st) {
  TcState* s = (TcState*)c->tc;
  const TcPlan* tp = (const TcPlan*)plan->tc;
  if (!tp) { rhe_set_error("rhe_block_accumulate: the plan was created without the tensor-core path"); return RHE_ERR_STATE; }
  if (gt) {
    if (tp->groups.size() != 1) { rhe_set_error("rhe_block_accumulate: individual-major copy given for a plan without fast path"); return RHE_ERR_INVALID; }
    return tc_pass_b2(c, s, gt, plan->m, &tp->groups[0], P_out, S_accum, st);
  }
  for (const TcBlockMeta& meta : tp->groups) {
    int rc = tc_pass_b_group(c, s, bed, plan->m, &meta, P_out, S_accum, st);
    if (rc) return rc;
  }
  return RHE_OK;
}
