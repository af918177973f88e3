// nb_publish.cuh -- commit of a replan as a committed-trajectory record (shared by k_commit and the cycle's k_publish).
#pragma once
#include "../../include/neptune_b200.h"
#include "nb_common.cuh"
#include "nb_hull.cuh"

// What NeptuneRos::publishOwnTraj adds to the trajectory (neptune_ros.cpp:436-480): id, is_agent, bbox, the publisher's
// position and its tether as bend points (base, then the contact point of every entry of entangle_state_.bendPointsIdx:
// the base of an agent id, col(case) of a static obstacle's representation).
struct NbPublishHdr
{
  int on;             // 0: header left zero (plain trajectory records)
  int N, M, cap;
  double bbox;        // 2 drone_radius
  const double* pb;   // [N][2]
  const double* strep;  // [M][2][2] or [N][M][2][2]
  int strep_per_agent;
  nb_ent_state es;    // entangle_state_ of the batch ([B] states); cnt may be nullptr: base point only
  double seq;
};

static __device__ __forceinline__ void nb_fill_header(double* r, int agent, int b, const NbPublishHdr& hd, int* err)
{
  r[NB_REC_ID] = (double)agent, r[NB_REC_ISAGENT] = 1.0, r[NB_REC_SEQ] = hd.seq;
  r[NB_REC_BBOX] = r[NB_REC_BBOX + 1] = r[NB_REC_BBOX + 2] = hd.bbox;
  const int n = (int)r[0];
  for (int ax = 0; ax < 3; ax++) r[NB_REC_POS + ax] = n > 0 ? r[1 + (NB_TP + 1) + ax * NB_TP * 4 + 3] : 0.0;  // start of piece 0
  int nb = 0;
  r[NB_REC_BEND] = hd.pb[2 * (agent - 1)], r[NB_REC_BEND + 1] = hd.pb[2 * (agent - 1) + 1];
  nb = 1;
  if (hd.es.cnt)
  {
    const int nbend = hd.es.cnt[2 * b + 1];
    const int* alpha = hd.es.alpha + (size_t)b * hd.cap * 2;
    const int* bend = hd.es.bend + (size_t)b * hd.cap;
    const double* rep = hd.strep + (hd.strep_per_agent ? (size_t)(agent - 1) * 4 * hd.M : 0);
    for (int q = 0; q < nbend; q++)
    {
      if (nb >= NB_REC_BEND_MAX)
      {
        *err = 6;
        break;
      }
      const int id = alpha[2 * bend[q]], cs = alpha[2 * bend[q] + 1];
      if (id <= hd.N)
        r[NB_REC_BEND + 2 * nb] = hd.pb[2 * (id - 1)], r[NB_REC_BEND + 2 * nb + 1] = hd.pb[2 * (id - 1) + 1];
      else
        r[NB_REC_BEND + 2 * nb] = rep[4 * (id - hd.N - 1) + 2 * cs], r[NB_REC_BEND + 2 * nb + 1] = rep[4 * (id - hd.N - 1) + 2 * cs + 1];
      nb++;
    }
  }
  r[NB_REC_NBEND] = (double)nb;
  for (int q = nb; q < NB_REC_BEND_MAX; q++) r[NB_REC_BEND + 2 * q] = r[NB_REC_BEND + 2 * q + 1] = 0.0;
  for (int q = NB_REC_SEQ + 1; q < NB_REC; q++) r[q] = 0.0;
}

// pwp_now of generatePwpOut (times shifted by t_start, solver_gurobi_poly.cpp:892-907) and, when t_now is
// given, pwp_out = composePieceWisePol(time_now, dc, pwp_prev, pwp_now) of Neptune::replanFull
// (neptune.cpp:1689-1699).  An agent whose replan failed, ended entangled or collides in the post-check keeps
// its previous record (replanFull returns before pwp_out is touched).  fe_solved: front end found no path
// ("returning with no solution", neptune.cpp:1473-1478) counts as rejected too.
// One CTA per agent b; r = where the record goes; now = NB_REC doubles of shared memory.
static __device__ __forceinline__ void nb_commit_one(int b, const int* n_int, const double* coeff, const double* t_start, double T, double* r,
                              double* now, const double* t_now, const double* prev, const int* prev_agent,
                              const uint8_t* has_prev, const int* status, const int* entangled, const int* collide,
                              const int* fe_solved, int* n_pieces, const NbPublishHdr& hd, int* err, double* prev_stage = nullptr)
{
  const int n = n_int[b];
  const bool compose = t_now != nullptr;
  const int agent = prev_agent ? prev_agent[b] : b + 1;
  for (int q = threadIdx.x; q < NB_REC; q += blockDim.x)
  {
    double v = 0.0;
    if (q == 0)
      v = (double)n;
    else if (q <= NB_TP + 1)
    {
      const int k = q - 1;
      v = k <= n ? NB_ADD(t_start[b], NB_MUL((double)k, T)) : 0.0;  // pwp_out.times[i] += t_start (:898)
    }
    else if (q < NB_REC_PWP)
    {
      const int e = q - (NB_TP + 2), ax = e / (NB_TP * 4), rem = e % (NB_TP * 4), piece = rem / 4, c = rem % 4;
      v = piece < n ? coeff[(size_t)b * 96 + ax * 32 + piece * 4 + c] : 0.0;
    }
    if (compose)
      now[q] = v;
    else
      r[q] = v;
  }
  __syncthreads();
  if (!compose)
  {
    if (hd.on && threadIdx.x == 0) nb_fill_header(r, agent, b, hd, err);
    return;
  }
  const double* pv = prev + (size_t)(agent - 1) * NB_REC;
  if (prev_stage)
  {  // the previous record into shared memory by the whole CTA: the (sequential) composition then reads it from there
    for (int q = threadIdx.x; q < NB_REC; q += blockDim.x) prev_stage[q] = pv[q];
    __syncthreads();
    pv = prev_stage;
  }
  const bool hp = has_prev == nullptr || has_prev[b];
  const bool ok = !((status && status[b] >= 2) || (entangled && entangled[b]) || (collide && collide[b]) ||
                    (fe_solved && !fe_solved[b]));
  if (!ok || !hp)
  {
    const double* src = (!ok && hp) ? pv : now;  // failed without a previous plan: nothing better to publish than pwp_now
    for (int q = threadIdx.x; q < NB_REC; q += blockDim.x) r[q] = src[q];
    if (threadIdx.x == 0 && n_pieces) n_pieces[b] = (int)src[0];
    __syncthreads();
    if (hd.on && src == now && threadIdx.x == 0) nb_fill_header(r, agent, b, hd, err);
    return;
  }
  if (prev_stage)
  {  // CTA-wide composition: one thread decides which pieces make up the result, all of them copy
    __shared__ NbComposePlan plan;
    if (threadIdx.x == 0)
    {
      int np = nb_compose_plan(t_now[b], pv, now, &plan);
      if (np < 0)
      {
        *err = 4;
        np = 0;
        plan.kind = 0;
      }
      if (n_pieces) n_pieces[b] = np;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NB_REC; q += blockDim.x) r[q] = nb_compose_fill(&plan, pv, now, q);
    __syncthreads();
    if (hd.on && threadIdx.x == 0) nb_fill_header(r, agent, b, hd, err);
    return;
  }
  if (threadIdx.x == 0)
  {
    int np = nb_compose_records(t_now[b], pv, now, r);
    if (np < 0)
    {
      *err = 4;
      np = 0;
    }
    if (n_pieces) n_pieces[b] = np;
    if (hd.on) nb_fill_header(r, agent, b, hd, err);
  }
}

