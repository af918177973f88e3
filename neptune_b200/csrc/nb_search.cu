// nb_search.cu -- kernel and launcher of K0, the front-end search (nb_search.cuh).  Its own translation
// unit because it is compiled with --fmad=false: every FP64 result in the search feeds a discrete
// decision (voxel keys, cost ties, admissibility tests), so no product-sum is contracted.
#include <cuda_runtime.h>

#include "nb_search.cuh"
#include "nb_search_launch.h"

static __host__ __device__ inline size_t nb_search_shared_bytes() { return (sizeof(NbSearchShared) + 15) & ~(size_t)15; }

// 25 "child" warps (one per jerk sample) + 7 auxiliary warps that run the collision tests of the popped node
// while the children are evaluated.  sync_aux is a named barrier over the auxiliary warps only.
struct NbCtaDev
{
  int tid, nthreads, warp, nwarps, lane, aux_tid, aux_n;
  bool child, aux;
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ int any(int p) const { return __syncthreads_or(p); }
  __device__ __forceinline__ void sync_aux() const
  {
    __syncwarp();
    asm volatile("bar.sync 2, %0;" ::"r"(NB_SEARCH_THREADS - NB_SEARCH_CHILD_THREADS) : "memory");
  }
};

__global__ void __launch_bounds__(NB_SEARCH_THREADS, 1) k_search(NbSearchArgs a, unsigned arena_bytes)
{
  extern __shared__ double nb_search_smem[];
  NbSearchShared* sh = reinterpret_cast<NbSearchShared*>(nb_search_smem);
  unsigned char* arena = reinterpret_cast<unsigned char*>(nb_search_smem) + nb_search_shared_bytes();
  NbCtaDev cta;
  cta.tid = threadIdx.x, cta.nthreads = blockDim.x, cta.warp = threadIdx.x >> 5, cta.nwarps = NB_SEARCH_CHILD_THREADS >> 5;
  cta.lane = threadIdx.x & 31;
  cta.child = threadIdx.x < NB_SEARCH_CHILD_THREADS, cta.aux = !cta.child;
  cta.aux_tid = threadIdx.x - NB_SEARCH_CHILD_THREADS, cta.aux_n = NB_SEARCH_THREADS - NB_SEARCH_CHILD_THREADS;
  nb_search_task<NbCtaDev, 32>(cta, a, blockIdx.x, blockIdx.x, sh, arena_bytes ? arena : nullptr, arena_bytes);
}

int nb_search_launch(const NbSearchArgs* a, int B, void* stream, int* smem_attr_set, const char** err)
{
  const size_t fixed = nb_search_shared_bytes();
  size_t arena = nb_search_arena_wanted(a->p);
  if (fixed + arena > NB_SEARCH_SMEM_MAX) arena = NB_SEARCH_SMEM_MAX - fixed;
  if (!*smem_attr_set)
  {  // per device / context: remembered in the caller's handle
    cudaError_t e = cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, NB_SEARCH_SMEM_MAX);
    if (e != cudaSuccess)
    {
      *err = cudaGetErrorString(e);
      return -1;
    }
    *smem_attr_set = 1;
  }
  k_search<<<B, NB_SEARCH_THREADS, fixed + arena, (cudaStream_t)stream>>>(*a, (unsigned)arena);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    *err = cudaGetErrorString(e);
    return -1;
  }
  return 0;
}
