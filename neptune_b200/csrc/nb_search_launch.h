// nb_search_launch.h -- host-side interface between nb_capi.cu (ABI) and nb_search.cu (kernel TU)
#pragma once
#include <stddef.h>

struct NbSearchArgs;

#define NB_SEARCH_THREADS 1024           // 25 child warps (one per jerk sample of the 5x5 lattice) + 7 auxiliary warps
#define NB_SEARCH_CHILD_THREADS 800
#define NB_SEARCH_SMEM_MAX (220 * 1024)  // dynamic shared memory the search kernel may ask for

// returns 0, or -1 with *err = CUDA error text
// smem_attr_set: the caller's per-handle flag (the opt-in shared-memory attribute is per device)
int nb_search_launch(const NbSearchArgs* a, int B, void* stream, int* smem_attr_set, const char** err);
