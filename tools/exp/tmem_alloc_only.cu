// minimal tensor-memory allocation (no mbarrier anywhere in the kernel): does compute-sanitizer --tool synccheck flag it?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void alloc_only(unsigned* out)
{
  __shared__ __align__(16) unsigned slot[4];
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((unsigned)__cvta_generic_to_shared(&slot[2])) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = *reinterpret_cast<volatile unsigned*>(&slot[2]);
  if (threadIdx.x == 0) out[blockIdx.x] = base;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(base) : "memory");
  }
}
int main()
{
  unsigned* d;
  cudaMalloc(&d, 1024 * sizeof(unsigned));
  alloc_only<<<512, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("tmem alloc-only kernel: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
