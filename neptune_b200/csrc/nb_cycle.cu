// nb_cycle.cu -- the replan cycle of one rank, resident on the device (include/neptune_b200.h, "The replan cycle").
//
// Host side of the hot part of Neptune::replanFull (reference neptune/src/neptune.cpp:1430-1448, :1450-1510, :1512-1529,
// :1641-1647, :1685-1699) and of the exchange around it (NeptuneRos::publishOwnTraj / trajCB, neptune_ros.cpp:379-480)
// as ONE launch sequence: side streams for the stages that do not depend on the optimisation, CUDA graphs (one per
// ring phase) for replay, and a commit kernel that writes every record straight into the rings of all ranks over
// NVLink and signals them -- the per-cycle all-gather fused into the kernel that produces its payload.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/neptune_b200.h"
#include "nb_handle.h"
#include "nb_publish.cuh"

#define CY_CUDA(call)                                                        \
  do                                                                         \
  {                                                                          \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess)                                                   \
    {                                                                        \
      nb_set_error((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); \
      return NB_ERR_CUDA;                                                    \
    }                                                                        \
  } while (0)
#define CY_RC(call)              \
  do                             \
  {                              \
    const int rc_ = (call);      \
    if (rc_ != NB_OK) return rc_; \
  } while (0)

#define NB_MAX_WORLD 16

// ------------------------------------------------------------------------------------------ kernels of the cycle
struct NbPeers
{
  double* ring[NB_MAX_WORLD];          // ring base of every rank (own included): [3][N][NB_REC]
  long long* flags[NB_MAX_WORLD];      // [world] arrival flags of every rank: flags[r][q] = last cycle whose records of q are in r's ring
  int world, rank;
};

// Commit + publish (tail of replanFull, publishOwnTraj): the record of agent b goes to slot agent_id[b] - 1 of ring phase
// `phase` -- in this rank's ring and in every peer's ring (plain stores through the peer mapping: NVLink) -- and when the
// last CTA of the grid has finished, the rank raises its arrival flag on every peer (release at system scope).
__global__ void __launch_bounds__(64) k_publish(int B, int N, const int* agent_id, const int* n_int, const double* coeff,
                                                const double* t_start, double T, const double* t_now, const double* prev,
                                                const int* status, const int* entangled, const int* collide,
                                                const int* fe_solved, int* n_pieces, NbPublishHdr hd, NbPeers peers, int phase,
                                                const long long* cycle_no, unsigned int* done, int* err)
{
  __shared__ double now[NB_REC];
  __shared__ double rec[NB_REC];
  __shared__ double pst[NB_REC];
  const int b = blockIdx.x;
  const int agent = agent_id[b];
  hd.seq = (double)*cycle_no;
  nb_commit_one(b, n_int, coeff, t_start, T, rec, now, t_now, prev, agent_id, nullptr, status, entangled, collide, fe_solved,
                n_pieces, hd, err, pst);
  __syncthreads();
  const size_t slot = ((size_t)phase * N + (agent - 1)) * NB_REC;
  for (int r = 0; r < peers.world; r++)
  {
    double* dst = peers.ring[r] + slot;
    for (int q = threadIdx.x; q < NB_REC; q += blockDim.x) dst[q] = rec[q];
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    // one fence per CTA: it is cumulative over the stores of the other threads, which the barrier has ordered before it;
    // system scope only when there are peers to publish to
    if (peers.world > 1)
      __threadfence_system();
    else
      __threadfence();
    const unsigned int prevc = atomicAdd(done, 1u);
    if (prevc == (unsigned int)B - 1)
    {  // every record of this rank is out: tell the peers
      *done = 0;
      const long long k1 = *cycle_no + 1;
      __threadfence_system();
      for (int r = 0; r < peers.world; r++)
      {
        volatile long long* f = peers.flags[r] + peers.rank;
        *f = k1;
      }
      __threadfence_system();
    }
  }
}

// End of the cycle: the records of every rank for this cycle have landed in this rank's ring.  One thread per peer spins
// on its flag (system-scope loads) with a generous bound, so that a lost peer shows up as an error, not as a hang.
__global__ void k_wait_peers(NbPeers peers, long long* cycle_no, int* err)
{
  const int r = threadIdx.x;
  const long long want = *cycle_no + 1;
  if (r < peers.world)
  {
    volatile long long* f = peers.flags[peers.rank] + r;
    long long t0 = clock64();
    while (*f < want)
    {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000LL)
      {  // ~10 s at 2 GHz
        *err = 7;
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (r == 0) *cycle_no = want;
}

// records nobody plans for are carried from the "late" slot of the ring into the "new" one
__global__ void k_carry(int N, const unsigned char* planned, const double* late, double* neu)
{
  const int j = blockIdx.x;
  if (j >= N || planned[j]) return;
  for (int q = threadIdx.x; q < NB_REC; q += blockDim.x) neu[(size_t)j * NB_REC + q] = late[(size_t)j * NB_REC + q];
}

// "returning with no solution" (neptune.cpp:1473-1478): an agent without a front-end path keeps the host-provided one in
// the arrays the back end reads, so that the batch stays dense; its replan is rejected in the commit
__global__ void k_fe_merge(int B, int cap, int NA, const int* solved, const int* n_int_h, const double* coeff_h,
                           const int* ecnt_h, const int* ealpha_h, const int* eact_h, int* n_int, double* coeff, int* ecnt,
                           int* ealpha, int* eact)
{
  const int b = blockIdx.x;
  if (b >= B || solved[b]) return;
  if (threadIdx.x == 0) n_int[b] = n_int_h[b];
  for (int q = threadIdx.x; q < 96; q += blockDim.x) coeff[(size_t)b * 96 + q] = coeff_h[(size_t)b * 96 + q];
  for (int q = threadIdx.x; q < 18; q += blockDim.x) ecnt[(size_t)b * 18 + q] = ecnt_h[(size_t)b * 18 + q];
  for (int q = threadIdx.x; q < 9 * cap * 2; q += blockDim.x) ealpha[(size_t)b * 9 * cap * 2 + q] = ealpha_h[(size_t)b * 9 * cap * 2 + q];
  for (int q = threadIdx.x; q < 9 * NA; q += blockDim.x) eact[(size_t)b * 9 * NA + q] = eact_h[(size_t)b * 9 * NA + q];
}

// Rank barrier on the device (nb_cycle_align): every rank raises its flag for `epoch` on every peer, then waits until all
// peers have raised theirs here.  Same transport as the commit (plain peer stores + system-scope fences).
__global__ void k_align(NbPeers peers, long long epoch, int* err)
{
  const int r = threadIdx.x;
  if (r < peers.world)
  {
    volatile long long* mine = peers.flags[r] + NB_MAX_WORLD + peers.rank;
    *mine = epoch;
    __threadfence_system();
    volatile long long* f = peers.flags[peers.rank] + NB_MAX_WORLD + r;
    const long long t0 = clock64();
    while (*f < epoch)
    {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000LL)
      {
        *err = 7;
        break;
      }
    }
  }
}

// active_cases rebuilt from the crossing list: active[id - 1] = number of entries of the list that carry id (every change
// of active_cases in eu::addAlphaBetaToList is paired with an append / erase of alphas, entangle_utils.cpp:1402-1534).
// One CTA per state; states: [n_states] lists of stride cap, counts cnt[2 * state] (n_alpha first).
__global__ void k_active_from_lists(int n_states, int cap, int NA, const int* cnt, const int* alpha, int* active)
{
  const int q = blockIdx.x;
  if (q >= n_states) return;
  int* act = active + (size_t)q * NA;
  for (int e = threadIdx.x; e < NA; e += blockDim.x) act[e] = 0;
  __syncthreads();
  const int n = cnt[2 * q];
  const int* al = alpha + (size_t)q * cap * 2;
  for (int e = threadIdx.x; e < n && e < cap; e += blockDim.x)
  {
    const int id = al[2 * e];
    if (id >= 1 && id <= NA) atomicAdd(act + id - 1, 1);
  }
}

// ------------------------------------------------------------------------------------------ the object
struct Field
{
  const char* name;
  void* ptr;
  size_t bytes;
};

struct nb_cycle
{
  nb_handle* h = nullptr;
  nb_cycle_desc d;
  std::vector<int32_t> agent_id;
  int N = 0, M = 0, NA = 0, cap = 0, S = 0, P = 0, bp_max = 0, G = 0;
  nb_cycle_layout lay;
  char *h_in = nullptr, *h_out = nullptr;       // pinned
  char *d_in = nullptr, *d_out = nullptr;
  // ring + flags (one allocation: IPC-exported as a whole)
  char* ring_alloc = nullptr;
  double* ring = nullptr;       // [3][N][NB_REC]
  long long* flags = nullptr;   // [world]
  NbPeers peers;
  void* peer_base[NB_MAX_WORLD] = { nullptr };
  long long* d_cycle = nullptr;  // device copy of k
  long long align_epoch = 0;
  long long last_upload_bytes = 0;   // what the last nb_cycle_upload* moved over PCIe
  bool split_postcheck = true;   // first half of the entanglement post-check beside the QP (NB_CYCLE_NO_SPLIT=1: one kernel after it)
  unsigned int* d_done = nullptr;
  long long k = 0;
  // intermediates
  int32_t* d_agent_id = nullptr;
  unsigned char *d_planned = nullptr, *d_ones = nullptr;
  char* esA = nullptr;  // packed entangle_state_A: cnt | alpha | beta | bend | active
  size_t es_bytes = 0, es_off[5] = { 0, 0, 0, 0, 0 };
  int32_t *bp_cnt = nullptr, *bp_cnt_l = nullptr;
  double *bp_xy = nullptr, *bp_xy_l = nullptr, *latest_pos = nullptr;
  double *hull_xy = nullptr, *hull_xy_l = nullptr, *nih0 = nullptr, *nih0_l = nullptr, *samp = nullptr, *aabb_l = nullptr;
  int32_t *hull_cnt = nullptr, *hull_cnt_l = nullptr;
  int64_t *hull_ptr = nullptr, *hull_ptr_l = nullptr;
  // front-end outputs
  int32_t *fe_n_int = nullptr, *fe_ecnt = nullptr, *fe_ealpha = nullptr, *fe_ebend = nullptr, *fe_eact = nullptr;
  double *fe_coeff = nullptr, *fe_ebeta = nullptr, *fe_cost = nullptr;
  // streams, events, graphs
  cudaStream_t sB = nullptr, sC = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_hulls = nullptr, ev_B = nullptr, ev_C = nullptr, ev_qp = nullptr, ev_C2 = nullptr, ev_pre = nullptr, ev_unpack = nullptr, ev_prof[10] = { nullptr };
  cudaGraphExec_t graph[3] = { nullptr, nullptr, nullptr };
  int graph_G = -1;
  long long launches_per_step = 0;
  std::vector<Field> fields;
  std::vector<void*> to_free;
};

namespace
{
template <typename T>
int dalloc(nb_cycle* c, T** p, size_t count)
{
  void* q = nullptr;
  if (cudaMalloc(&q, count * sizeof(T) + 16) != cudaSuccess)
  {
    nb_set_error("nb_cycle: cudaMalloc failed");
    return NB_ERR_CUDA;
  }
  cudaMemset(q, 0, count * sizeof(T) + 16);
  c->to_free.push_back(q);
  *p = (T*)q;
  return NB_OK;
}

int64_t place(int64_t& off, size_t bytes)
{
  const int64_t at = off;
  off += (int64_t)((bytes + 255) / 256 * 256);
  return at;
}

double* ring_slot(nb_cycle* c, int phase) { return c->ring + (size_t)phase * c->N * NB_REC; }
int ph_new(long long k) { return (int)(k % 3); }
int ph_late(long long k) { return (int)((k + 2) % 3); }
int ph_known(long long k) { return (int)((k + 1) % 3); }

template <typename T>
T* in_at(nb_cycle* c, int64_t off)
{
  return off < 0 ? nullptr : (T*)(c->d_in + off);
}
template <typename T>
T* out_at(nb_cycle* c, int64_t off)
{
  return off < 0 ? nullptr : (T*)(c->d_out + off);
}

nb_ent_state es_view(char* base, const size_t* off)
{
  nb_ent_state e;
  e.cnt = (int32_t*)(base + off[0]), e.alpha = (int32_t*)(base + off[1]), e.beta = (double*)(base + off[2]);
  e.bend = (int32_t*)(base + off[3]), e.active = (int32_t*)(base + off[4]);
  return e;
}

// The launch sequence of one cycle.  prof != nullptr: single stream, an event after every stage.
int step_body(nb_cycle* c, cudaStream_t st, bool prof)
{
  nb_handle* h = c->h;
  const int B = c->d.B, N = c->N, G = c->G;
  const nb_cycle_layout& L = c->lay;
  cudaStream_t sB = prof ? st : c->sB, sC = prof ? st : c->sC;
  const long long k = c->k;
  double *r_new = ring_slot(c, ph_new(k)), *r_late = ring_slot(c, ph_late(k)), *r_known = ring_slot(c, ph_known(k));
  const int32_t* agent_id = c->d_agent_id;
  const int32_t* group = in_at<int32_t>(c, L.group);
  const double* t_group = in_at<double>(c, L.t_group);
  const unsigned char *known = in_at<unsigned char>(c, L.known), *late = in_at<unsigned char>(c, L.late);
  const long long l0 = h->launches;
  int pe = 0;
  auto mark = [&]() {
    if (prof) cudaEventRecord(c->ev_prof[pe++], st);
  };
  mark();
  if (!prof)
  {
    CY_CUDA(cudaEventRecord(c->ev_fork, st));
    CY_CUDA(cudaStreamWaitEvent(sC, c->ev_fork, 0));
  }
  // (stream C) what arrived during the optimisation: bend points of the late messages (trajCB) and the hulls of the late
  // trajectories over the planning windows (neptune.cpp:737-741, :792) -- independent of the optimisation
  CY_RC(nb_unpack_records_batch(h, NB_DEVICE, r_late, c->bp_cnt_l, c->bp_xy_l, nullptr, sC));
  CY_RC(nb_hulls_batch(h, G, NB_DEVICE, t_group, r_late, c->d_ones, c->d.delta, c->hull_xy_l, c->hull_cnt_l, c->hull_ptr_l, c->nih0_l,
                       nullptr, nullptr, sC));
  CY_RC(nb_internal_hull_aabb(h, (size_t)G * N * NB_NPOL, c->hull_xy_l, c->hull_cnt_l, c->aabb_l, sC));
  if (!prof) CY_CUDA(cudaEventRecord(c->ev_C, sC));
  mark();
  // (stream B) trajCB bookkeeping of the trajectories the agents plan against (bend points, latest positions) and the
  // records of the agents this rank does not plan for, beside the hulls; (main) hulls and samples (neptune.cpp:1433-1434)
  if (!prof) CY_CUDA(cudaStreamWaitEvent(sB, c->ev_fork, 0));
  CY_RC(nb_unpack_records_batch(h, NB_DEVICE, r_known, c->bp_cnt, c->bp_xy, c->latest_pos, sB));
  k_carry<<<N, 64, 0, sB>>>(N, c->d_planned, r_late, r_new);
  h->launches += 1;
  if (!prof) CY_CUDA(cudaEventRecord(c->ev_unpack, sB));
  CY_RC(nb_hulls_batch(h, G, NB_DEVICE, t_group, r_known, c->d_ones, c->d.delta, c->hull_xy, c->hull_cnt, c->hull_ptr, c->nih0,
                       c->samp, nullptr, st));
  mark();
  // (stream B) entangle_state_A = PredictAlphasBetas(entangle_state_) (:1445): needs the samples only
  if (!prof)
  {
    CY_CUDA(cudaEventRecord(c->ev_hulls, st));
    CY_CUDA(cudaStreamWaitEvent(sB, c->ev_hulls, 0));
    CY_CUDA(cudaStreamWaitEvent(st, c->ev_unpack, 0));   // the back end reads the bend points
  }
  CY_CUDA(cudaMemcpyAsync(c->esA, c->d_in + L.es_cnt, c->es_bytes, cudaMemcpyDeviceToDevice, sB));
  nb_ent_state esA = es_view(c->esA, c->es_off);
  CY_RC(nb_internal_predict_grouped(h, B, agent_id, known, c->bp_cnt, c->bp_xy, esA, in_at<double>(c, L.prev_pos),
                                    in_at<double>(c, L.prev_pos_agent), in_at<double>(c, L.cur), c->samp, group, sB));
  if (!prof) CY_CUDA(cudaEventRecord(c->ev_B, sB));
  mark();
  const int32_t* n_int = in_at<int32_t>(c, L.n_int);
  const double* coeff_init = in_at<double>(c, L.coeff_init);
  const int32_t *esv_cnt = in_at<int32_t>(c, L.esv_cnt), *esv_alpha = in_at<int32_t>(c, L.esv_alpha), *esv_active = in_at<int32_t>(c, L.esv_active);
  const int32_t* fe_solved = nullptr;
  if (c->d.front_end)
  {  // KinodynamicSearch::setUp + run (:1450-1453) from entangle_state_A and the shared hulls / samples; pwp_init and
     // entStateVec stay on the device for the back end (:1509-1517)
    if (!prof) CY_CUDA(cudaStreamWaitEvent(st, c->ev_B, 0));
    nb_search_args sa;
    memset(&sa, 0, sizeof(sa));
    sa.B = B, sa.space = NB_DEVICE, sa.agent_id = agent_id, sa.init = in_at<double>(c, L.fe_init), sa.goal = in_at<double>(c, L.fe_goal);
    sa.coeffs_z = in_at<double>(c, L.fe_coeffs_z), sa.n_groups = G, sa.group = group, sa.hull_xy = c->hull_xy, sa.hull_cnt = c->hull_cnt;
    sa.samp = c->samp, sa.known = known, sa.es = esA, sa.bp_cnt = c->bp_cnt, sa.bp_xy = c->bp_xy;
    sa.comb = in_at<uint8_t>(c, L.fe_comb), sa.comb_shared = 0;
    sa.status = out_at<int32_t>(c, L.fe_status), sa.solved = out_at<int32_t>(c, L.fe_solved), sa.n_int = c->fe_n_int, sa.coeff = c->fe_coeff;
    sa.esv.cnt = c->fe_ecnt, sa.esv.alpha = c->fe_ealpha, sa.esv.beta = c->fe_ebeta, sa.esv.bend = c->fe_ebend, sa.esv.active = c->fe_eact;
    sa.stats = out_at<int32_t>(c, L.fe_stats), sa.cost = c->fe_cost;
    CY_RC(nb_search_batch(h, &sa, st));
    fe_solved = sa.solved;
    k_fe_merge<<<B, 128, 0, st>>>(B, c->cap, c->NA, fe_solved, n_int, coeff_init, esv_cnt, esv_alpha, esv_active, c->fe_n_int,
                                  c->fe_coeff, c->fe_ecnt, c->fe_ealpha, c->fe_eact);
    h->launches += 1;
    n_int = c->fe_n_int, coeff_init = c->fe_coeff, esv_cnt = c->fe_ecnt, esv_alpha = c->fe_ealpha, esv_active = c->fe_eact;
    if (L.fe_n_int >= 0)
      CY_CUDA(cudaMemcpyAsync(c->d_out + L.fe_n_int, c->fe_n_int, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  mark();
  // (stream B) first half of the entanglement post-check (neptune.cpp:735-749): the late agents' samples over the span of
  // the trajectory being optimised, their late bend points, PredictAlphasBetas afresh -- none of it needs the optimiser's
  // answer, so it runs beside the lines + QP kernels
  nb_ent_state es0 = es_view(c->d_in + L.es_cnt, c->es_off);
  double* coeff_out = out_at<double>(c, L.coeff_out);
  if (!prof)
  {
    CY_CUDA(cudaEventRecord(c->ev_pre, st));
    CY_CUDA(cudaStreamWaitEvent(sB, c->ev_pre, 0));
    CY_CUDA(cudaStreamWaitEvent(sB, c->ev_C, 0));   // bend points of the late messages (stream C)
  }
  if (c->split_postcheck)
    CY_RC(nb_internal_postcheck_entangle(h, 1, B, NB_DEVICE, agent_id, known, late, c->bp_cnt, c->bp_xy, c->bp_cnt_l, c->bp_xy_l, es0,
                                         in_at<double>(c, L.prev_pos), in_at<double>(c, L.prev_pos_agent), in_at<double>(c, L.cur),
                                         n_int, coeff_out, in_at<double>(c, L.t_start), c->samp, 0, group, r_late,
                                         out_at<int32_t>(c, L.entangled), sB));
  if (!prof) CY_CUDA(cudaEventRecord(c->ev_B, sB));
  // back end: separating lines + trajectory QP (:1514-1519); shared-window mode of nb_replan_batch
  nb_replan_args a;
  memset(&a, 0, sizeof(a));
  a.B = B, a.space = NB_DEVICE, a.agent_id = agent_id, a.n_int = n_int, a.coeff_init = coeff_init, a.n_hull_slots = N;
  a.hull_ptr = nullptr, a.hull_xy = c->hull_xy, a.hull_cnt = c->hull_cnt, a.hull_nvert = (int64_t)G * N * NB_NPOL * NB_HULL_STRIDE;
  a.nih0 = c->nih0, a.nih0_group = group, a.hull_known = known;
  a.esv_cnt = esv_cnt, a.esv_alpha = esv_alpha, a.esv_active = esv_active, a.bp_cnt = c->bp_cnt, a.bp_xy = c->bp_xy;
  a.coeff_out = out_at<double>(c, L.coeff_out), a.obj = out_at<double>(c, L.obj), a.status = out_at<int32_t>(c, L.status);
  a.iters = out_at<int32_t>(c, L.iters);
  if (c->d.front_end == 2)
  {  // Neptune::replanKinodynamic (neptune.cpp:1010-1300): no back end -- the front-end path itself (generatePwpOut of the
     // search, :1189) goes to safetyCheckAfterReplan (:1198) and is committed (:1244-1255); status 0 = "optimised" path
    CY_CUDA(cudaMemcpyAsync(a.coeff_out, coeff_init, (size_t)B * 96 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CY_CUDA(cudaMemsetAsync(a.status, 0, (size_t)B * sizeof(int32_t), st));
    CY_CUDA(cudaMemsetAsync(a.obj, 0, (size_t)B * sizeof(double), st));
    CY_CUDA(cudaMemsetAsync(a.iters, 0, (size_t)B * 2 * sizeof(int32_t), st));
  }
  else
    CY_RC(nb_replan_batch(h, &a, st));
  mark();
  // safetyCheckAfterReplan (:719-752): GJK against the late hulls, then the gated entanglement re-check
  // the two halves are independent: the GJK half runs on stream C beside the entanglement half on the main stream
  if (!prof)
  {
    CY_CUDA(cudaEventRecord(c->ev_qp, st));
    CY_CUDA(cudaStreamWaitEvent(sC, c->ev_qp, 0));
    CY_CUDA(cudaStreamWaitEvent(st, c->ev_B, 0));
  }
  CY_RC(nb_internal_postcheck_hulls(h, B, n_int, a.coeff_out, group, c->hull_xy_l, c->hull_cnt_l, c->aabb_l, late,
                                    out_at<int32_t>(c, L.collide), sC));
  if (!prof) CY_CUDA(cudaEventRecord(c->ev_C2, sC));
  CY_RC(nb_internal_postcheck_entangle(h, c->split_postcheck ? 2 : 0, B, NB_DEVICE, agent_id, known, late, c->bp_cnt, c->bp_xy, c->bp_cnt_l, c->bp_xy_l, es0,
                                       in_at<double>(c, L.prev_pos), in_at<double>(c, L.prev_pos_agent), in_at<double>(c, L.cur),
                                       n_int, a.coeff_out, in_at<double>(c, L.t_start), c->samp, 0, group, r_late,
                                       out_at<int32_t>(c, L.entangled), st));
  if (!prof) CY_CUDA(cudaStreamWaitEvent(st, c->ev_C2, 0));
  mark();
  // commit: compose with the previous plan, DynTraj header, records into the ring of every rank (:1685-1699, publishOwnTraj)
  NbPublishHdr hd;
  hd.on = 1, hd.N = N, hd.M = c->M, hd.cap = c->cap, hd.bbox = c->d.bbox, hd.pb = h->d_pb, hd.strep = h->d_strep;
  hd.strep_per_agent = h->strep_per_agent, hd.es = es0, hd.seq = 0.0;
  k_publish<<<B, 64, 0, st>>>(B, N, agent_id, n_int, a.coeff_out, in_at<double>(c, L.t_start), h->cs.T, in_at<double>(c, L.t_now),
                              r_late, a.status, out_at<int32_t>(c, L.entangled), out_at<int32_t>(c, L.collide), fe_solved,
                              out_at<int32_t>(c, L.n_pieces), hd, c->peers, ph_new(k), c->d_cycle, c->d_done, (int*)h->err.p);
  h->launches += 1;
  mark();
  k_wait_peers<<<1, 32, 0, st>>>(c->peers, c->d_cycle, (int*)h->err.p);
  h->launches += 1;
  mark();
  CY_CUDA(cudaGetLastError());
  c->launches_per_step = h->launches - l0;
  return NB_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------ ABI
extern "C" int nb_cycle_create(nb_handle* h, const nb_cycle_desc* d, nb_cycle** out)
{
  if (!h || !d || !out || d->B < 1 || !d->agent_id || d->world < 1 || d->world > NB_MAX_WORLD || d->rank < 0 || d->rank >= d->world)
  {
    nb_set_error("nb_cycle_create: invalid descriptor");
    return NB_ERR_ARG;
  }
  if (d->front_end && !h->sp_set)
  {
    nb_set_error("nb_cycle_create: front_end needs nb_search_configure first");
    return NB_ERR_ARG;
  }
  CY_CUDA(cudaSetDevice(h->device));
  nb_cycle* c = new nb_cycle();
  c->h = h, c->d = *d;
  // measured (tools/cycle_rank0_of.py, bench.py grid1024 block): beside the QP the first half saves 35-50 us per cycle while
  // SMs are idle (64 agents against 64 ... 512: 0.544 -> 0.496 ms in the 512-agent world); in the 1224-tether world it is a
  // wash with 1024 agents per GPU (3.69 vs 3.75 ms) and a loss with 128 (1.56 vs 1.74 ms per cycle on eight GPUs: the first
  // half alone is as long as the QP there), so from 1024 tethers on it stays one kernel after the QP
  const int NA_ = h->par.num_agents + h->par.num_static;
  c->split_postcheck = getenv("NB_CYCLE_NO_SPLIT") == nullptr &&
                       ((NA_ < 1024 && (NA_ < 512 || d->B <= h->num_sms)) || getenv("NB_CYCLE_FORCE_SPLIT"));
  c->agent_id.assign(d->agent_id, d->agent_id + d->B);
  c->d.agent_id = c->agent_id.data();
  const int B = d->B, N = h->par.num_agents, M = h->par.num_static, NA = N + M, cap = h->par.ent_cap, S = h->par.samples;
  const int P = h->par.num_pol, bm = h->par.bp_max;
  c->N = N, c->M = M, c->NA = NA, c->cap = cap, c->S = S, c->P = P, c->bp_max = bm;
  for (int b = 0; b < B; b++)
    if (d->agent_id[b] < 1 || d->agent_id[b] > N)
    {
      nb_set_error("nb_cycle_create: agent_id out of range");
      delete c;
      return NB_ERR_ARG;
    }
  // ---- packed layouts
  nb_cycle_layout& L = c->lay;
  int64_t off = 0;
  L.n_int = place(off, (size_t)B * 4), L.coeff_init = place(off, (size_t)B * 96 * 8), L.t_start = place(off, (size_t)B * 8);
  L.t_now = place(off, (size_t)B * 8), L.t_group = place(off, (size_t)B * 8), L.group = place(off, (size_t)B * 4);
  L.known = place(off, (size_t)B * N), L.late = place(off, (size_t)B * N);
  L.esv_cnt = place(off, (size_t)B * 18 * 4), L.esv_alpha = place(off, (size_t)B * 9 * cap * 8), L.esv_active = place(off, (size_t)B * 9 * NA * 4);
  // entangle_state_: five arrays back to back with the padding of place(), copied as one block into entangle_state_A
  const int64_t es0 = off;
  L.es_cnt = place(off, (size_t)B * 2 * 4), L.es_alpha = place(off, (size_t)B * cap * 8), L.es_beta = place(off, (size_t)B * cap * 8);
  L.es_bend = place(off, (size_t)B * cap * 4), L.es_active = place(off, (size_t)B * NA * 4);
  c->es_bytes = (size_t)(off - es0);
  c->es_off[0] = 0, c->es_off[1] = (size_t)(L.es_alpha - es0), c->es_off[2] = (size_t)(L.es_beta - es0);
  c->es_off[3] = (size_t)(L.es_bend - es0), c->es_off[4] = (size_t)(L.es_active - es0);
  L.prev_pos = place(off, (size_t)B * (N + 1) * 16), L.prev_pos_agent = place(off, (size_t)B * N * 16), L.cur = place(off, (size_t)B * 16);
  L.fe_init = L.fe_goal = L.fe_coeffs_z = L.fe_comb = -1;
  if (d->front_end)
  {
    const int ns2 = h->sp.num_samples * h->sp.num_samples;
    L.fe_init = place(off, (size_t)B * 48), L.fe_goal = place(off, (size_t)B * 16), L.fe_coeffs_z = place(off, (size_t)B * 32 * 8);
    L.fe_comb = place(off, (size_t)B * ns2);
  }
  L.in_bytes = off;
  off = 0;
  L.coeff_out = place(off, (size_t)B * 96 * 8), L.obj = place(off, (size_t)B * 8), L.status = place(off, (size_t)B * 4);
  L.iters = place(off, (size_t)B * 8), L.entangled = place(off, (size_t)B * 4), L.collide = place(off, (size_t)B * 4);
  L.n_pieces = place(off, (size_t)B * 4);
  L.fe_status = L.fe_solved = L.fe_n_int = L.fe_stats = -1;
  if (d->front_end)
    L.fe_status = place(off, (size_t)B * 4), L.fe_solved = place(off, (size_t)B * 4), L.fe_n_int = place(off, (size_t)B * 4),
    L.fe_stats = place(off, (size_t)B * 16);
  L.out_bytes = off;
  int rc = [&]() -> int {
    CY_CUDA(cudaMallocHost((void**)&c->h_in, (size_t)L.in_bytes));
    CY_CUDA(cudaMallocHost((void**)&c->h_out, (size_t)L.out_bytes));
    memset(c->h_in, 0, (size_t)L.in_bytes), memset(c->h_out, 0, (size_t)L.out_bytes);
    CY_RC(dalloc(c, &c->d_in, (size_t)L.in_bytes));
    CY_RC(dalloc(c, &c->d_out, (size_t)L.out_bytes));
    // ring + flags in one allocation (IPC export)
    const size_t ring_bytes = (size_t)3 * N * NB_REC * 8, flag_bytes = (size_t)2 * NB_MAX_WORLD * 8;   // arrival flags, then the flags of nb_cycle_align
    CY_CUDA(cudaMalloc((void**)&c->ring_alloc, ring_bytes + flag_bytes));
    CY_CUDA(cudaMemset(c->ring_alloc, 0, ring_bytes + flag_bytes));
    c->ring = (double*)c->ring_alloc, c->flags = (long long*)(c->ring_alloc + ring_bytes);
    memset(&c->peers, 0, sizeof(c->peers));
    c->peers.world = 1, c->peers.rank = 0;   // until nb_cycle_open_peers: a world of one (own ring only)
    c->peers.ring[0] = c->ring, c->peers.flags[0] = c->flags;
    CY_RC(dalloc(c, &c->d_cycle, 1));
    CY_RC(dalloc(c, &c->d_done, 1));
    CY_RC(dalloc(c, &c->d_agent_id, (size_t)B));
    CY_CUDA(cudaMemcpy(c->d_agent_id, d->agent_id, (size_t)B * 4, cudaMemcpyHostToDevice));
    CY_RC(dalloc(c, &c->d_planned, (size_t)N));
    std::vector<unsigned char> pl(N, 0);
    if (d->planned)
      pl.assign(d->planned, d->planned + N);
    else
      for (int b = 0; b < B; b++) pl[d->agent_id[b] - 1] = 1;
    CY_CUDA(cudaMemcpy(c->d_planned, pl.data(), (size_t)N, cudaMemcpyHostToDevice));
    CY_RC(dalloc(c, &c->esA, c->es_bytes));
    CY_RC(dalloc(c, &c->bp_cnt, (size_t)N));
    CY_RC(dalloc(c, &c->bp_cnt_l, (size_t)N));
    CY_RC(dalloc(c, &c->bp_xy, (size_t)N * bm * 2));
    CY_RC(dalloc(c, &c->bp_xy_l, (size_t)N * bm * 2));
    CY_RC(dalloc(c, &c->latest_pos, (size_t)N * 2));
    if (d->front_end)
    {
      CY_RC(dalloc(c, &c->fe_n_int, (size_t)B));
      CY_RC(dalloc(c, &c->fe_coeff, (size_t)B * 96));
      CY_RC(dalloc(c, &c->fe_ecnt, (size_t)B * 18));
      CY_RC(dalloc(c, &c->fe_ealpha, (size_t)B * 9 * cap * 2));
      CY_RC(dalloc(c, &c->fe_ebeta, (size_t)B * 9 * cap));
      CY_RC(dalloc(c, &c->fe_ebend, (size_t)B * 9 * cap));
      CY_RC(dalloc(c, &c->fe_eact, (size_t)B * 9 * NA));
      CY_RC(dalloc(c, &c->fe_cost, (size_t)B));
    }
    CY_CUDA(cudaStreamCreateWithFlags(&c->sB, cudaStreamNonBlocking));
    CY_CUDA(cudaStreamCreateWithFlags(&c->sC, cudaStreamNonBlocking));
    for (cudaEvent_t* e : { &c->ev_fork, &c->ev_hulls, &c->ev_B, &c->ev_C, &c->ev_qp, &c->ev_C2, &c->ev_pre, &c->ev_unpack }) CY_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (auto& e : c->ev_prof) CY_CUDA(cudaEventCreate(&e));
    return NB_OK;
  }();
  if (rc != NB_OK)
  {
    nb_cycle_destroy(c);
    return rc;
  }
  *out = c;
  return NB_OK;
}

extern "C" void nb_cycle_destroy(nb_cycle* c)
{
  if (!c) return;
  cudaSetDevice(c->h->device);
  cudaDeviceSynchronize();
  for (auto& g : c->graph)
    if (g) cudaGraphExecDestroy(g);
  for (int r = 0; r < NB_MAX_WORLD; r++)
    if (c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  for (void* p : c->to_free) cudaFree(p);
  cudaFree(c->ring_alloc);
  if (c->h_in) cudaFreeHost(c->h_in);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->sB) cudaStreamDestroy(c->sB);
  if (c->sC) cudaStreamDestroy(c->sC);
  for (cudaEvent_t e : { c->ev_fork, c->ev_hulls, c->ev_B, c->ev_C, c->ev_qp, c->ev_C2, c->ev_pre, c->ev_unpack })
    if (e) cudaEventDestroy(e);
  for (auto e : c->ev_prof)
    if (e) cudaEventDestroy(e);
  delete c;
}

extern "C" int nb_cycle_layout_get(const nb_cycle* c, nb_cycle_layout* out)
{
  if (!c || !out) return NB_ERR_ARG;
  *out = c->lay;
  return NB_OK;
}
extern "C" void* nb_cycle_host_in(nb_cycle* c) { return c ? c->h_in : nullptr; }
extern "C" void* nb_cycle_host_out(nb_cycle* c) { return c ? c->h_out : nullptr; }
extern "C" long long nb_cycle_launches_per_step(const nb_cycle* c) { return c ? c->launches_per_step : 0; }
extern "C" long long nb_cycle_index(const nb_cycle* c) { return c ? c->k : 0; }

namespace
{
// buffers whose size depends on the number of window groups
int ensure_groups(nb_cycle* c, int G)
{
  if (G == c->G) return NB_OK;
  if (G < 1 || G > c->d.B)
  {
    nb_set_error("nb_cycle_upload: n_groups out of range (1..B)");
    return NB_ERR_ARG;
  }
  CY_CUDA(cudaDeviceSynchronize());
  const size_t nh = (size_t)G * c->N * NB_NPOL;
  // (the old buffers stay in to_free until the cycle is destroyed: regrouping is rare)
  CY_RC(dalloc(c, &c->hull_xy, nh * NB_HULL_STRIDE * 2));
  CY_RC(dalloc(c, &c->hull_xy_l, nh * NB_HULL_STRIDE * 2));
  CY_RC(dalloc(c, &c->hull_cnt, nh));
  CY_RC(dalloc(c, &c->hull_cnt_l, nh));
  CY_RC(dalloc(c, &c->hull_ptr, nh));
  CY_RC(dalloc(c, &c->hull_ptr_l, nh));
  CY_RC(dalloc(c, &c->nih0, nh * 2));
  CY_RC(dalloc(c, &c->nih0_l, nh * 2));
  CY_RC(dalloc(c, &c->aabb_l, nh * 4));
  CY_RC(dalloc(c, &c->samp, (size_t)G * c->N * c->P * (c->S + 1) * 2));
  CY_RC(dalloc(c, &c->d_ones, (size_t)G * c->N));
  CY_CUDA(cudaMemset(c->d_ones, 1, (size_t)G * c->N));
  c->G = G;
  return NB_OK;
}
}  // namespace

extern "C" int nb_cycle_seed_records(nb_cycle* c, const double* known_recs, const double* late_recs, void* stream)
{
  if (!c || !known_recs || !late_recs) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)c->N * NB_REC * 8;
  CY_CUDA(cudaMemcpyAsync(ring_slot(c, ph_known(c->k)), known_recs, bytes, cudaMemcpyHostToDevice, st));
  CY_CUDA(cudaMemcpyAsync(ring_slot(c, ph_late(c->k)), late_recs, bytes, cudaMemcpyHostToDevice, st));
  return NB_OK;
}

extern "C" int nb_cycle_upload_from(nb_cycle* c, const void* host_in, int32_t n_groups, void* stream);
extern "C" int nb_cycle_upload(nb_cycle* c, int32_t n_groups, void* stream)
{
  if (!c) return NB_ERR_ARG;
  return nb_cycle_upload_from(c, c->h_in, n_groups, stream);
}

extern "C" int nb_cycle_download(nb_cycle* c, void* stream)
{
  if (!c) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  CY_CUDA(cudaMemcpyAsync(c->h_out, c->d_out, (size_t)c->lay.out_bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return NB_OK;
}

extern "C" int nb_cycle_upload_from(nb_cycle* c, const void* host_in, int32_t n_groups, void* stream)
{
  if (!c || !host_in) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  CY_RC(ensure_groups(c, n_groups));
  cudaStream_t st = (cudaStream_t)stream;
  const nb_cycle_layout& L = c->lay;
  const char* hp = (const char*)host_in;
  const int B = c->d.B, cap = c->cap, NA = c->NA;
  if ((L.in_bytes < (4 << 20) && !getenv("NB_CYCLE_SPARSE_UPLOAD")) || getenv("NB_CYCLE_DENSE_UPLOAD"))
  {  // small worlds: one copy of the packed buffer
    CY_CUDA(cudaMemcpyAsync(c->d_in, host_in, (size_t)L.in_bytes, cudaMemcpyHostToDevice, st));
    c->last_upload_bytes = L.in_bytes;
    return NB_OK;
  }
  // Large worlds: the entanglement lists are allocated for 3 (N + M) entries per state and hold a few dozen, and the
  // active-case arrays ([N + M] per state) are a function of the lists.  Only the used prefix of every list row crosses
  // PCIe (2-D copies out of the same packed host buffer) and active_cases is rebuilt on the device; everything else goes
  // as before.  Nothing downstream reads a list beyond its count.
  const int32_t* esv_cnt = (const int32_t*)(hp + L.esv_cnt);
  const int32_t* es_cnt = (const int32_t*)(hp + L.es_cnt);
  int ma = 0, me = 0, mb = 0;
  for (int q = 0; q < B * 9; q++) ma = esv_cnt[2 * q] > ma ? esv_cnt[2 * q] : ma;
  for (int b = 0; b < B; b++) me = es_cnt[2 * b] > me ? es_cnt[2 * b] : me, mb = es_cnt[2 * b + 1] > mb ? es_cnt[2 * b + 1] : mb;
  ma = ma > cap ? cap : ma, me = me > cap ? cap : me, mb = mb > cap ? cap : mb;
  long long moved = 0;
  auto flat = [&](int64_t from, int64_t to) -> int {
    if (to > from) CY_CUDA(cudaMemcpyAsync(c->d_in + from, hp + from, (size_t)(to - from), cudaMemcpyHostToDevice, st));
    moved += to - from;
    return NB_OK;
  };
  auto rows = [&](int64_t off, size_t pitch, size_t width, size_t n) -> int {
    if (width > 0 && n > 0)
      CY_CUDA(cudaMemcpy2DAsync(c->d_in + off, pitch, hp + off, pitch, width, n, cudaMemcpyHostToDevice, st));
    moved += (long long)(width * n);
    return NB_OK;
  };
  CY_RC(flat(0, L.esv_alpha));                                              // n_int ... esv_cnt
  CY_RC(rows(L.esv_alpha, (size_t)cap * 8, (size_t)ma * 8, (size_t)B * 9));
  CY_RC(flat(L.es_cnt, L.es_alpha));                                        // es_cnt
  CY_RC(rows(L.es_alpha, (size_t)cap * 8, (size_t)me * 8, (size_t)B));
  CY_RC(rows(L.es_beta, (size_t)cap * 8, (size_t)me * 8, (size_t)B));
  CY_RC(rows(L.es_bend, (size_t)cap * 4, (size_t)mb * 4, (size_t)B));
  CY_RC(flat(L.prev_pos, L.in_bytes));                                      // prev_pos ... front-end inputs
  k_active_from_lists<<<B * 9, 128, 0, st>>>(B * 9, cap, NA, (const int*)(c->d_in + L.esv_cnt), (const int*)(c->d_in + L.esv_alpha),
                                           (int*)(c->d_in + L.esv_active));
  k_active_from_lists<<<B, 128, 0, st>>>(B, cap, NA, (const int*)(c->d_in + L.es_cnt), (const int*)(c->d_in + L.es_alpha),
                                       (int*)(c->d_in + L.es_active));
  CY_CUDA(cudaGetLastError());
  c->last_upload_bytes = moved;
  return NB_OK;
}

extern "C" long long nb_cycle_last_upload_bytes(const nb_cycle* c) { return c ? c->last_upload_bytes : 0; }

extern "C" int nb_cycle_download_to(nb_cycle* c, void* host_out, void* stream)
{
  if (!c || !host_out) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  CY_CUDA(cudaMemcpyAsync(host_out, c->d_out, (size_t)c->lay.out_bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return NB_OK;
}

extern "C" void* nb_pinned_alloc(int64_t bytes)
{
  void* p = nullptr;
  if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) return nullptr;
  memset(p, 0, (size_t)bytes);
  return p;
}
extern "C" void nb_pinned_free(void* p)
{
  if (p) cudaFreeHost(p);
}

extern "C" int nb_cycle_capture(nb_cycle* c, void* stream)
{
  if (!c || c->G < 1) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!st)
  {
    nb_set_error("nb_cycle_capture: capture needs a non-default stream");
    return NB_ERR_ARG;
  }
  for (auto& g : c->graph)
    if (g) cudaGraphExecDestroy(g), g = nullptr;
  // one REAL cycle outside capture first (it counts: k advances): every library buffer (scratch, search workspace)
  // reaches its size, so that nothing allocates while the stream is capturing
  CY_RC(step_body(c, st, false));
  c->k += 1;
  CY_CUDA(cudaStreamSynchronize(st));
  const long long k0 = c->k;
  const long long before = c->h->launches;
  for (int ph = 0; ph < 3; ph++)
  {
    c->k = k0 + ph;   // only k % 3 matters to the captured pointers
    cudaGraph_t g = nullptr;
    CY_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = step_body(c, st, false);
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != NB_OK || e != cudaSuccess)
    {
      if (g) cudaGraphDestroy(g);
      c->k = k0;
      if (rc == NB_OK) nb_set_error(cudaGetErrorString(e));
      return rc != NB_OK ? rc : NB_ERR_CUDA;
    }
    CY_CUDA(cudaGraphInstantiate(&c->graph[(k0 + ph) % 3], g, 0));
    cudaGraphDestroy(g);
  }
  c->h->launches = before;   // captured, not launched
  c->k = k0;
  c->graph_G = c->G;
  return NB_OK;
}

extern "C" int nb_cycle_step(nb_cycle* c, void* stream)
{
  if (!c || c->G < 1) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  cudaStream_t st = (cudaStream_t)stream;
  cudaGraphExec_t g = c->graph_G == c->G ? c->graph[c->k % 3] : nullptr;
  if (g)
  {
    CY_CUDA(cudaGraphLaunch(g, st));
    c->h->launches += c->launches_per_step;
  }
  else
    CY_RC(step_body(c, st, false));
  c->k += 1;
  return NB_OK;
}

extern "C" int nb_cycle_step_profiled(nb_cycle* c, void* stream, double* ms)
{
  if (!c || !ms || c->G < 1) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  cudaStream_t st = (cudaStream_t)stream;
  CY_RC(step_body(c, st, true));
  c->k += 1;
  CY_CUDA(cudaStreamSynchronize(st));
  for (int q = 0; q < 8; q++)
  {
    float f = 0.f;
    CY_CUDA(cudaEventElapsedTime(&f, c->ev_prof[q], c->ev_prof[q + 1]));
    ms[q] = f;
  }
  return NB_OK;
}

extern "C" int nb_cycle_fetch(nb_cycle* c, const char* name, void* dst, int64_t bytes, void* stream)
{
  if (!c || !name || !dst) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  const std::string n(name);
  const size_t rec = (size_t)c->N * NB_REC * 8, G = (size_t)(c->G > 0 ? c->G : 1);
  const void* src = nullptr;
  size_t have = 0;
  // the last completed cycle is k - 1: its new / late / known slots
  const long long kk = c->k > 0 ? c->k - 1 : 0;
  if (n == "ring_new") src = ring_slot(c, ph_new(kk)), have = rec;
  else if (n == "ring_late") src = ring_slot(c, ph_late(kk)), have = rec;
  else if (n == "ring_known") src = ring_slot(c, ph_known(kk)), have = rec;
  else if (n == "in_esv_active") src = c->d_in + c->lay.esv_active, have = (size_t)c->d.B * 9 * c->NA * 4;
  else if (n == "in_es_active") src = c->d_in + c->lay.es_active, have = (size_t)c->d.B * c->NA * 4;
  else if (n == "esA_cnt") src = c->esA + c->es_off[0], have = (size_t)c->d.B * 8;
  else if (n == "esA_alpha") src = c->esA + c->es_off[1], have = (size_t)c->d.B * c->cap * 8;
  else if (n == "esA_beta") src = c->esA + c->es_off[2], have = (size_t)c->d.B * c->cap * 8;
  else if (n == "esA_bend") src = c->esA + c->es_off[3], have = (size_t)c->d.B * c->cap * 4;
  else if (n == "esA_active") src = c->esA + c->es_off[4], have = (size_t)c->d.B * c->NA * 4;
  else if (n == "hull_cnt") src = c->hull_cnt, have = G * c->N * NB_NPOL * 4;
  else if (n == "hull_xy") src = c->hull_xy, have = G * c->N * NB_NPOL * NB_HULL_STRIDE * 16;
  else if (n == "samp") src = c->samp, have = G * c->N * c->P * (c->S + 1) * 16;
  else if (n == "bp_cnt") src = c->bp_cnt, have = (size_t)c->N * 4;
  else if (n == "bp_xy") src = c->bp_xy, have = (size_t)c->N * c->bp_max * 16;
  else if (n == "latest_pos") src = c->latest_pos, have = (size_t)c->N * 16;
  else if (n == "fe_coeff" && c->fe_coeff) src = c->fe_coeff, have = (size_t)c->d.B * 96 * 8;
  else if (n == "fe_esv_cnt" && c->fe_ecnt) src = c->fe_ecnt, have = (size_t)c->d.B * 18 * 4;
  else if (n == "fe_esv_alpha" && c->fe_ealpha) src = c->fe_ealpha, have = (size_t)c->d.B * 9 * c->cap * 8;
  else if (n == "fe_esv_active" && c->fe_eact) src = c->fe_eact, have = (size_t)c->d.B * 9 * c->NA * 4;
  else if (n == "fe_esv_beta" && c->fe_ebeta) src = c->fe_ebeta, have = (size_t)c->d.B * 9 * c->cap * 8;
  else if (n == "fe_esv_bend" && c->fe_ebend) src = c->fe_ebend, have = (size_t)c->d.B * 9 * c->cap * 4;
  else if (n == "fe_cost" && c->fe_cost) src = c->fe_cost, have = (size_t)c->d.B * 8;
  else if (n == "fe_n_int" && c->fe_n_int) src = c->fe_n_int, have = (size_t)c->d.B * 4;
  else if (n == "hull_cnt_late") src = c->hull_cnt_l, have = G * c->N * NB_NPOL * 4;
  if (!src || (size_t)bytes > have)
  {
    nb_set_error("nb_cycle_fetch: unknown array or too many bytes");
    return NB_ERR_ARG;
  }
  CY_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CY_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return NB_OK;
}

extern "C" int nb_cycle_align(nb_cycle* c, void* stream)
{
  if (!c) return NB_ERR_ARG;
  if (c->d.world <= 1) return NB_OK;
  CY_CUDA(cudaSetDevice(c->h->device));
  c->align_epoch += 1;   // every rank calls this the same number of times
  k_align<<<1, 32, 0, (cudaStream_t)stream>>>(c->peers, c->align_epoch, (int*)c->h->err.p);
  CY_CUDA(cudaGetLastError());
  return NB_OK;
}

extern "C" int nb_cycle_ipc_handle(nb_cycle* c, void* handle64)
{
  if (!c || !handle64) return NB_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CY_CUDA(cudaSetDevice(c->h->device));
  cudaIpcMemHandle_t hd;
  CY_CUDA(cudaIpcGetMemHandle(&hd, c->ring_alloc));
  memcpy(handle64, &hd, 64);
  return NB_OK;
}

extern "C" int nb_cycle_open_peers(nb_cycle* c, const void* handles)
{
  if (!c || !handles) return NB_ERR_ARG;
  CY_CUDA(cudaSetDevice(c->h->device));
  const int W = c->d.world, me = c->d.rank;
  const size_t ring_bytes = (size_t)3 * c->N * NB_REC * 8;
  c->peers.world = W, c->peers.rank = me;
  for (int r = 0; r < W; r++)
  {
    char* base = c->ring_alloc;
    if (r != me)
    {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, (const char*)handles + (size_t)r * 64, 64);
      void* p = nullptr;
      CY_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
      c->peer_base[r] = p;
      base = (char*)p;
    }
    c->peers.ring[r] = (double*)base;
    c->peers.flags[r] = (long long*)(base + ring_bytes);
  }
  for (auto& g : c->graph)   // the captured commit kernel carries the peer table by value
    if (g) cudaGraphExecDestroy(g), g = nullptr;
  c->graph_G = -1;
  return NB_OK;
}
