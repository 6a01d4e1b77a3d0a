// Probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (profiling / bring-up utility, not part of the library):
// which tensor-map box shape it wants, how the four gathered rows land in shared memory, what happens to
// out-of-range rows, and how fast one warp can issue 32 of them (= one 128-row pass-B stage).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gather4_test tools/gather4_test.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap map, const int* rows, int n_rows, int col, int W, uint8_t* out,
                        long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  const uint32_t bar_s = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x;
  const int n_inst = n_rows / 4;
  const int r0 = rows[4 * lane], r1 = rows[4 * lane + 1], r2 = rows[4 * lane + 2], r3 = rows[4 * lane + 3];
  long long t0 = 0, t1 = 0, t2 = 0;
  const int rounds = 21;                               // round 0 is the cold one; report the average of the rest
  for (int rd = 0; rd < rounds; ++rd) {
    if (rd == 1) t0 = clock64();
    long long ta = clock64();
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(n_rows * W) : "memory");
    __syncwarp();
    if (lane < n_inst) {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
          ::"r"(smem_u32(smem) + lane * 4 * W), "l"(&map), "r"(bar_s), "r"(col + (rd % 4) * 128), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    }
    if (rd >= 1) t1 += clock64() - ta;
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_s), "r"(rd & 1) : "memory");
    if (!ok && lane == 0) printf("TIMEOUT waiting for gather4 bytes\n");
    __syncwarp();
  }
  t2 = clock64() - t0;
  t1 /= (rounds - 1);
  t2 /= (rounds - 1);
  t0 = 0;
  for (int i = lane; i < n_rows * W; i += 32) out[i] = smem[i];
  if (lane == 0) { cycles[0] = t1; cycles[1] = t2; }
}

// NW warps issue 32 gather4 each per round into their own 128 x W tile and barrier: does the rate scale with warps?
__global__ void k_rate(const __grid_constant__ CUtensorMap map, const int* rows, int W, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t bar_s = smem_u32(&bar[warp]);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int r0 = rows[4 * lane], r1 = rows[4 * lane + 1], r2 = rows[4 * lane + 2], r3 = rows[4 * lane + 3];
  const uint32_t tile = smem_u32(smem) + warp * 128 * W + lane * 4 * W;
  const int rounds = 41;
  long long t0 = 0;
  for (int rd = 0; rd < rounds; ++rd) {
    if (rd == 1) { __syncthreads(); t0 = clock64(); }
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(128 * W) : "memory");
    __syncwarp();
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(tile), "l"(&map), "r"(bar_s), "r"((rd % 8) * 128 + warp * 16), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_s), "r"(rd & 1) : "memory");
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[0] = (clock64() - t0) / (rounds - 1);
  (void)nw;
}

int main() {
  const int M = 4096, pitch = 1024;           // rows x bytes
  std::vector<uint8_t> h((size_t)M * pitch);
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < pitch; ++c) h[(size_t)r * pitch + c] = (uint8_t)((r * 7 + c * 3) & 255);
  uint8_t *d, *out;
  int* d_rows;
  long long* d_cyc;
  cudaMalloc(&d, h.size());
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 65536);
  cudaMalloc(&d_cyc, 16);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  std::vector<int> rows(128);
  for (int i = 0; i < 128; ++i) rows[i] = (i * 37 + 11) % M;
  rows[5] = M + 3;                             // one out-of-range row: expect zero fill
  cudaMalloc(&d_rows, sizeof(int) * 128);
  cudaMemcpy(d_rows, rows.data(), sizeof(int) * 128, cudaMemcpyHostToDevice);
  for (int W : {32, 64, 128}) {
    for (int box_rows : {1}) {
      for (int sw = 0; sw < 2; ++sw) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        CUtensorMapSwizzle swz = !sw ? CU_TENSOR_MAP_SWIZZLE_NONE
                                     : (W == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : W == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("W=%d box_rows=%d sw=%d: encode failed (%d)\n", W, box_rows, sw, (int)r); continue; }
        cudaMemset(out, 0xEE, 65536);
        const int col = 0;   // the last round reads columns [(20 % 4) * 128, +W) = [0, W)
        k_probe<<<1, 32, 128 * 128>>>(map, d_rows, 128, col, W, out, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("W=%d box_rows=%d sw=%d: kernel failed: %s\n", W, box_rows, sw, cudaGetErrorString(e)); return 1; }
        std::vector<uint8_t> o(128 * W);
        long long cyc[2];
        cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost);
        // check: row i of the tile == source row rows[i], bytes [col, col + W), with the swizzle of 16-byte chunks applied
        int bad = 0, bad_oob = 0;
        const int chunks = W / 16;
        for (int i = 0; i < 128; ++i)
          for (int b = 0; b < W; ++b) {
            int c16 = b / 16;
            int phys = c16;
            if (sw) {
              const int rows_per_128 = 128 / W;           // swizzle atom: 16-byte chunk index XOR ((128-byte line index) mod chunks)
              phys = c16 ^ ((i / rows_per_128) % chunks);
            }
            const uint8_t got = o[(size_t)i * W + phys * 16 + (b & 15)];
            const uint8_t want = rows[i] < M ? h[(size_t)rows[i] * pitch + col + b] : 0;
            if (got != want) { if (rows[i] < M) ++bad; else ++bad_oob; }
          }
        printf("W=%3d box_rows=%d swizzle=%d: mismatches=%d (oob row: %d)  issue %lld cyc, complete %lld cyc\n", W, box_rows, sw, bad,
               bad_oob, cyc[0], cyc[1]);
      }
    }
  }
  for (int W : {64, 128}) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)W, 1};
    cuuint32_t estr[2] = {1, 1};
    enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        W == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int nw : {1, 2, 4, 8}) {
      cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 128 * 128);
      k_rate<<<1, 32 * nw, nw * 128 * W>>>(map, d_rows, W, d_cyc);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("rate: W=%3d, %d warps x 32 gather4 per round: %lld cycles per round = %.1f cycles per gather4 (SM-wide) %s\n", W, nw, cyc,
             (double)cyc / (32.0 * nw), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
