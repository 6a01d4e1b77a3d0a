// Access-pattern microbenchmark for the packed-genotype reads of the two tensor passes (profiling utility, not part
// of the library).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membench tools/membench.cu
//
//   mode B (pass-B pattern): CTA x owns a column window of W bytes of every SNP row and walks down the rows,
//                            128 rows per step; thread t of group q reads 32 bytes (two 16-byte loads) of row t.
//   mode A (pass-A pattern): CTA (y, tile) owns 128 SNP rows and walks along them, W bytes per row and step;
//                            the `splits` CTAs of a tile take interleaved or contiguous steps.
// Every thread keeps D steps of loads in flight.  Prints the achieved GB/s of each pattern.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int D>
__global__ void k_mode_b(const uint8_t* __restrict__ bed, size_t pitch, int m, int W, uint32_t* out) {
  const int t = threadIdx.x & 127, q = threadIdx.x >> 7;
  const uint8_t* base = bed + (size_t)blockIdx.x * W + q * 32;
  const int n_step = m / 128;
  uint4 lo[D], hi[D];
  uint32_t acc = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)(d * 128 + t) * pitch);
    lo[d] = ldg_nc(p); hi[d] = ldg_nc(p + 1);
  }
  for (int s0 = 0; s0 < n_step; s0 += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      acc ^= lo[d].x ^ lo[d].y ^ lo[d].z ^ lo[d].w ^ hi[d].x ^ hi[d].y ^ hi[d].z ^ hi[d].w;
      const int s = s0 + d + D;
      if (s < n_step) {
        const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)(s * 128 + t) * pitch);
        lo[d] = ldg_nc(p); hi[d] = ldg_nc(p + 1);
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// W bytes per row and step = (W / 32) groups of 128 threads, each thread 32 bytes
template <int D>
__global__ void k_mode_a(const uint8_t* __restrict__ bed, size_t pitch, int row_bytes, int W, int interleave, uint32_t* out) {
  const int t = threadIdx.x & 127, q = threadIdx.x >> 7;
  const int splits = gridDim.x, y = blockIdx.x;
  const int total = row_bytes / W;
  int n_step, first, stride;
  if (interleave) { n_step = total > y ? (total - y + splits - 1) / splits : 0; first = y; stride = splits; }
  else { const int per = (total + splits - 1) / splits; first = y * per; n_step = min(per, total - first); stride = 1; }
  const uint8_t* base = bed + (size_t)(blockIdx.y * 128 + t) * pitch + q * 32;
  uint4 lo[D], hi[D];
  uint32_t acc = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (d < n_step) {
      const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)(first + d * stride) * W);
      lo[d] = ldg_nc(p); hi[d] = ldg_nc(p + 1);
    } else { lo[d] = hi[d] = make_uint4(0, 0, 0, 0); }
  }
  for (int s0 = 0; s0 < n_step; s0 += D) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      acc ^= lo[d].x ^ lo[d].y ^ lo[d].z ^ lo[d].w ^ hi[d].x ^ hi[d].y ^ hi[d].z ^ hi[d].w;
      const int s = s0 + d + D;
      if (s < n_step) {
        const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)(first + s * stride) * W);
        lo[d] = ldg_nc(p); hi[d] = ldg_nc(p + 1);
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// mode W (wide pass-A pattern): CTA (y, tile) owns 128 SNP rows; per step it reads W bytes of every row, LPR = W / 128
// threads per row each taking one 128-byte line (eight 16-byte loads), steps interleaved across the `splits` CTAs.
__global__ void k_mode_w(const uint8_t* __restrict__ bed, size_t pitch, int row_bytes, int W, uint32_t* out) {
  const int lpr = W / 128;
  const int r = threadIdx.x / lpr, c = threadIdx.x % lpr;
  const int splits = gridDim.x, y = blockIdx.x;
  const int total = row_bytes / W;
  const int n_step = total > y ? (total - y + splits - 1) / splits : 0;
  const uint8_t* base = bed + (size_t)(blockIdx.y * 128 + r) * pitch + c * 128;
  uint32_t acc = 0;
  uint4 v[8];
  for (int s = 0; s < n_step; ++s) {
    const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)(y + s * splits) * W);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ldg_nc(p + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static void timeit(const char* name, double bytes, F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  printf("%-52s %8.3f ms  %8.1f GB/s %s\n", name, ms / reps, bytes / (ms / reps) * 1e-6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const int m = 10240;                       // SNP rows of a block
  const size_t pitch = 125056;               // bytes per row (500224 individuals)
  const int row_bytes = 124928;              // multiple of 512 used by the walks
  uint8_t* bed; uint32_t* out;
  // several blocks so that consecutive launches do not hit L2
  const int n_blocks = 4;
  cudaMalloc(&bed, (size_t)n_blocks * m * pitch);
  cudaMemset(bed, 1, (size_t)n_blocks * m * pitch);
  cudaMalloc(&out, 64);
  int blk = 0;
  auto next = [&]() { blk = (blk + 1) % n_blocks; return bed + (size_t)blk * m * pitch; };
  const double bytes = (double)m * row_bytes;
  char name[128];
  for (int W : {32, 64, 128, 256}) {
    const int threads = W / 32 * 128;
    snprintf(name, sizeof name, "B: window %3d B/row/CTA, D=3, %d thr", W, threads);
    timeit(name, bytes, [&]() { k_mode_b<3><<<row_bytes / W, threads>>>(next(), pitch, m, W, out); });
    snprintf(name, sizeof name, "B: window %3d B/row/CTA, D=6, %d thr", W, threads);
    timeit(name, bytes, [&]() { k_mode_b<6><<<row_bytes / W, threads>>>(next(), pitch, m, W, out); });
  }
  for (int W : {128, 256}) {
    for (int inter = 0; inter < 2; ++inter) {
      for (int splits : {4, 11, 22}) {
        const int threads = W / 32 * 128;
        snprintf(name, sizeof name, "A: %3d B/row/step, splits %2d, %s, D=4", W, splits, inter ? "interleaved" : "contiguous");
        timeit(name, bytes, [&]() { k_mode_a<4><<<dim3(splits, m / 128), threads>>>(next(), pitch, row_bytes, W, inter, out); });
      }
    }
  }
  for (int W : {128, 256, 512, 1024}) {
    for (int splits : {11, 22}) {
      const int threads = 128 * (W / 128);
      snprintf(name, sizeof name, "W: %4d B/row/step (line per thread), splits %2d", W, splits);
      timeit(name, bytes, [&]() { k_mode_w<<<dim3(splits, m / 128), threads>>>(next(), pitch, row_bytes, W, out); });
    }
  }
  return 0;
}
