// nb_handle.h -- internal: the library handle (shared by nb_capi.cu and nb_cycle.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/neptune_b200.h"
#include "nb_common.cuh"

struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) return -1;
    // Library-owned scratch and staging buffers start zeroed: a host-space call copies whole output arrays back, and the
    // parts a kernel has no reason to write (list tails beyond their counts, slots of agents without a result) must not
    // carry whatever an earlier allocation of the process left in that memory.
    if (cudaMemset(p, 0, want) != cudaSuccess)
    {
      cudaFree(p);
      p = nullptr;
      return -1;
    }
    cap = want;
    return 0;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct nb_handle
{
  nb_params par;
  NbConsts cs;
  int device;
  long long launches;
  NbQpTable* d_tables;  // [2][NB_NPOL] : mode-major
  double* d_pb;
  int64_t* d_st_ptr;
  double* d_st_xy;
  double* d_strep;
  int strep_per_agent = 0;   // d_strep is [N][M][2][2] (nb_set_static_rep_per_agent), d_st_longest [N][M][2]
  int64_t st_nvert;
  // staging of NB_HOST arguments
  DevBuf in[16], out[8];
  // scratch
  DevBuf lines, line_ok, keep, cl, ncl, rows, err, ent_scratch;
  // front-end search: configuration, staging and workspace
  nb_search_params sp;
  int sp_set = 0;
  int search_smem_set = 0;
  int sprof_B = 0;
  double* d_st_longest = nullptr;
  DevBuf sprof, qprof;
  int qprof_B = 0;
  DevBuf sin[20], sout[12], sw_meta, sw_kin, sw_alpha, sw_beta, sw_bend, sw_hash, sw_heap, sw_gh, sw_ng, sw_chi, sw_chd, sw_fcode;
  int qp_smem_set = 0;
  int num_sms = 148;
  int profiling = 0;
  cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
};


// internal entry points shared between nb_capi.cu and nb_cycle.cu
void nb_set_error(const char* text);
int nb_internal_predict_grouped(nb_handle* h, int B, const int32_t* agent_id, const uint8_t* known, const int32_t* bp_cnt,
                                const double* bp_xy, nb_ent_state stt, const double* prev_pos, const double* prev_pos_agent,
                                const double* cur, const double* samp_g, const int32_t* group, cudaStream_t st);
int nb_internal_hull_aabb(nb_handle* h, size_t n_hulls, const double* hull_xy, const int* hull_cnt, double* aabb, cudaStream_t st);
int nb_internal_postcheck_hulls(nb_handle* h, int B, const int32_t* n_int, const double* coeff, const int32_t* group,
                                const double* hull_xy_g, const int32_t* hull_cnt_g, const double* aabb, const uint8_t* late,
                                int32_t* collide, cudaStream_t st);
int nb_internal_postcheck_entangle(nb_handle* h, int phase, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                                   const uint8_t* late, const int32_t* bp_cnt, const double* bp_xy, const int32_t* bp_cnt_late,
                                   const double* bp_xy_late, nb_ent_state st_in, const double* prev_pos,
                                   const double* prev_pos_agent, const double* cur, const int32_t* n_int, const double* coeff,
                                   const double* t_start, const double* samp, int32_t samp_shared, const int32_t* samp_group,
                                   const double* late_recs, int32_t* entangled, void* stream);
