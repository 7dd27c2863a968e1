// Micro-benchmark: throughput of tcgen05.ld (TMEM -> registers) on one SM for the access patterns an epilogue
// can use. 16 warps (warp w may only touch TMEM lanes 32 (w & 3) .. +31), 512 allocated columns, no MMA running.
// Every pattern reads R rounds; a round moves the same 128 lanes x 64 columns x 4 B = 32 KB.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/tmem_ld_bench.cu -o scripts/bin/tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD16(addr, v)                                                                                            \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
               : "r"(addr) : "memory")
#define LD8(addr, v)                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"    \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) \
               : "r"(addr) : "memory")
#define LD32(addr, v)                                                                                            \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"   \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                          \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
               : "r"(addr) : "memory")
// 16 lanes x 256 bits, repeated x4: a warp reads 16 lanes x 32 columns (4 registers per thread per repeat... x4 -> 16 regs)
#define LD16x256_4(addr, v)                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
               : "r"(addr) : "memory")
#define WAIT() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// mode 0: x16, one load in flight, wait after each (the shipped epilogue without its arithmetic)
// mode 1: x16, issue the next load BEFORE waiting for the previous one is impossible (wait::ld waits for all):
//         two loads per wait (32 registers)
// mode 2: x32, one per wait, 16 warps x 32 columns (two rounds' worth per wait)
// mode 3: x16, four loads per wait (64 registers)
// mode 4: x8, one per wait
// mode 5: 16x256b.x4 (16 lanes x 32 columns per instruction, two instructions per 32 lanes), two per wait
// mode 6: x16 one per wait, only 8 warps active (2 per lane quarter)
// mode 7: x16 one per wait, only 4 warps active (1 per lane quarter)
__global__ void __launch_bounds__(512, 1) bench(int mode, int rounds, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t cq = (uint32_t)(warp >> 2);
  uint32_t acc = 0;
  uint32_t a[32], b[32];
  const bool active = mode == 6 ? warp < 8 : (mode == 7 ? warp < 4 : true);
  __syncthreads();
  const long long t0 = clock64();
  if (active) {
    for (int r = 0; r < rounds; ++r) {
      const uint32_t col = (uint32_t)((r & 7) * 64);
      if (mode == 0 || mode == 6 || mode == 7) {
        LD16(base + col + 16 * (cq & 3), a); WAIT(); acc += a[0] ^ a[15];
      } else if (mode == 1) {
        uint32_t* a2 = a + 16;
        LD16(base + col + 16 * cq, a); LD16(base + ((col + 64) & 511) + 16 * cq, a2); WAIT(); acc += a[0] ^ a2[15]; ++r;
      } else if (mode == 2) {
        LD32(base + ((col + 32 * cq) & 511), a); WAIT(); acc += a[0] ^ a[31]; ++r;
      } else if (mode == 3) {
        uint32_t* a2 = a + 16; uint32_t* b2 = b + 16;
        LD16(base + col + 16 * cq, a); LD16(base + ((col + 64) & 511) + 16 * cq, a2);
        LD16(base + ((col + 128) & 511) + 16 * cq, b); LD16(base + ((col + 192) & 511) + 16 * cq, b2);
        WAIT(); acc += a[0] ^ a2[15] ^ b[3] ^ b2[7]; r += 3;
      } else if (mode == 4) {
        LD8(base + col + 16 * cq, a); WAIT(); LD8(base + col + 16 * cq + 8, b); WAIT(); acc += a[0] ^ b[7];
      } else if (mode == 5) {
        uint32_t* a2 = a + 16;
        const uint32_t hb = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
        LD16x256_4(hb + ((col + 32 * (cq & 1)) & 511), a); LD16x256_4(hb + (16u << 16) + ((col + 32 * (cq & 1)) & 511), a2);
        WAIT(); acc += a[0] ^ a2[15]; ++r;
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * 512 + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4 * 512 * 148);
  const int rounds = 4096;
  const char* names[] = {"32x32b.x16, 16 warps, 1 load per wait", "32x32b.x16, 16 warps, 2 loads per wait", "32x32b.x32, 16 warps, 1 load per wait",
                         "32x32b.x16, 16 warps, 4 loads per wait", "32x32b.x8, 16 warps, 1 load per wait", "16x256b.x4, 16 warps, 2 loads per wait",
                         "32x32b.x16, 8 warps, 1 load per wait", "32x32b.x16, 4 warps, 1 load per wait"};
  for (int mode = 0; mode < 8; ++mode) {
    bench<<<1, 512>>>(mode, rounds, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long cyc; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    const int warps = mode == 6 ? 8 : (mode == 7 ? 4 : 16);
    double bytes = (double)rounds * warps * 32 * 16 * 4;      // every warp moves 2 KB per round
    printf("%-44s %8.1f cycles per 16-warp-round-equivalent (32 KB), %6.1f B/clk\n", names[mode], cyc / (bytes / 32768.0), bytes / cyc);
  }
  return 0;
}
