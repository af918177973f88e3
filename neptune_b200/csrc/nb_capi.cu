// nb_capi.cu -- C-ABI (include/neptune_b200.h) and kernel launches of libneptune_b200.so.
// sm_100a only; no CPU fallback: every entry point fails with NB_ERR_NO_DEVICE / NB_ERR_CUDA when no
// B200-class device is usable.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/neptune_b200.h"
#include "nb_common.cuh"
#include "nb_entangle.cuh"
#include "nb_hull.cuh"
#include "nb_lines.cuh"
#include "nb_publish.cuh"
#include "nb_qp.cuh"
#include "nb_search.cuh"
#include "nb_search_launch.h"
#include "nb_sep.cuh"
#include "nb_tables.h"

static thread_local std::string g_err;
extern "C" const char* nb_last_error(void) { return g_err.c_str(); }
void nb_set_error(const char* text) { g_err = text; }  // internal (nb_cycle.cu)

#define NB_CUDA(call)                                                                         \
  do                                                                                          \
  {                                                                                           \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
    {                                                                                         \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
      return NB_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#include "nb_handle.h"

// ------------------------------------------------------------------------------------------ kernels

// one CTA per (agent, interval): 128 threads over the LP slots, 512 in worlds with a thousand slots per interval
template <int NT>
__global__ void __launch_bounds__(NT) k_lines(NbConsts cs, NbLinesIn in, int LS, double* lines, uint8_t* ok,
                                              uint8_t* keep, int* err, double* cl, int* ncl)
{
  extern __shared__ double smem_d[];
  NbPruneShared ps;
  ps.px = smem_d;
  ps.py = ps.px + (LS + 1);
  ps.red = (int*)(ps.py + (LS + 1));
  ps.hull = ps.red + 512;
  ps.misc = ps.hull + (NB_PRUNE_KMAX + 1);
  ps.valid = (uint8_t*)(ps.misc + 24);  // misc holds one entry per warp (up to 16)
  const int b = blockIdx.x / NB_NPOL, i = blockIdx.x % NB_NPOL;
  nb_lines_task<NT>(threadIdx.x, b, i, cs, in, lines + (size_t)blockIdx.x * LS * 3, ok + (size_t)blockIdx.x * LS,
                     keep + (size_t)blockIdx.x * LS, ps, err, cl + (size_t)blockIdx.x * LS * 3, ncl + blockIdx.x);
}

static size_t lines_smem_bytes(int LS)
{
  return (size_t)(LS + 1) * 16 + (512 + NB_PRUNE_KMAX + 1 + 8 + 16) * sizeof(int) + (size_t)(LS + 1) + 16;
}

struct NbQpArgs
{
  const int* n_int;
  const double* coeff_init;
  const double* cl;     // [B][8][LS][3] kept lines of every (agent, interval), compact (k_lines)
  const int* ncl;       // [B][8] their number
  int LS;
  double* rows;         // [B][5][RS]     (only used when an agent keeps more than NB_QP_SMEM_LINES lines)
  int RS;
  double* coeff_out;
  double* obj;
  int* status;
  int* iters;
  int* err;             // device error latch (5 = n_int out of range in a device-resident batch)
  long long* prof;      // optional [B][16] phase cycles (nb_set_profiling)
};

#define NB_QP_SMEM_LINES 160     // kept lines whose rows live in shared memory; more: rows in global scratch
#define NB_QP_SMEM_ROWS (NB_ROW_LINE0 + 4 * NB_QP_SMEM_LINES)

#define NB_QP_THREADS 128        // four warps per agent; NB_QP_THREADS_WIDE when the batch leaves SMs idle (B <= SM count)
#define NB_QP_THREADS_WIDE 256
struct NbQpSmem
{
  NbQpTable tb;  // first: destination of the bulk copy (16-byte aligned)
  NbQpShared sh;
  double red[8 * NB_QP_THREADS_WIDE / 32];
  double rows[5 * NB_QP_SMEM_ROWS];
  double cl[3 * NB_QP_SMEM_LINES];
  double xout[96];
  unsigned long long mbar;
  const double* clb[NB_NPOL];
  int lstart[12];
  unsigned char l2i[NB_QP_SMEM_LINES];
};

// ---- bulk asynchronous copy global -> shared (TMA unit, no tensor map: cp.async.bulk) completing on an mbarrier
NB_DEV unsigned nb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
NB_DEV void nb_mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nb_smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
NB_DEV void nb_bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nb_smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   nb_smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(nb_smem_addr(bar))
               : "memory");
}
NB_DEV void nb_mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NB_DONE_%=;\n"
      "bra NB_WAIT_%=;\n"
      "NB_DONE_%=:\n"
      "}\n" ::"r"(nb_smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// K4: one CTA of four warps per agent.  The (n, mode) table (32 KB) is staged by one bulk copy that overlaps the
// gathering of the agent's lines (nb_qp.cuh has the solver).
template <int NT>
__global__ void __launch_bounds__(NT) k_qp(NbConsts cs, const NbQpTable* tables, NbQpArgs a)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  NbQpSmem* sm = reinterpret_cast<NbQpSmem*>(smem_raw);
  const int b = blockIdx.x, lane = threadIdx.x;
  Group<NT> g(lane, sm->red);
  const int n = a.n_int[b];
  const double* ci = a.coeff_init + (size_t)b * 96;
  if (n < 1 || n > NB_NPOL)
  {  // device-resident batch with a corrupt n_int: nothing is indexed with it; reported by nb_check_async_errors
    for (int q = lane; q < 96; q += NT) a.coeff_out[(size_t)b * 96 + q] = ci[q];
    if (lane == 0)
    {
      *a.err = 5;
      a.obj[b] = 0.0, a.status[b] = NB_STATUS_FAILED, a.iters[2 * b] = 0, a.iters[2 * b + 1] = 0;
    }
    return;
  }
  if (lane == 0)
  {
    nb_mbar_init(&sm->mbar, 1);
    nb_bulk_load(&sm->tb, tables + (n - 1), (unsigned)sizeof(NbQpTable), &sm->mbar);
  }
  // line runs of the intervals: exclusive prefix of the counts (lanes 0..7 hold one interval each)
  int cnt = (lane < n) ? a.ncl[(size_t)b * NB_NPOL + lane] : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1)
  {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane <= NB_NPOL) sm->lstart[lane] = 0;
  __syncthreads();
  if (lane < n) sm->lstart[lane + 1] = incl;
  __syncthreads();
  const int nl = sm->lstart[n];
  const bool in_smem = nl <= NB_QP_SMEM_LINES;
  NbQpRows R;
  const size_t rs = in_smem ? NB_QP_SMEM_ROWS : a.RS;
  double* rb = in_smem ? sm->rows : a.rows + (size_t)b * 5 * a.RS;
  R.s = rb, R.lam = rb + rs, R.dsa = rb + 2 * rs, R.dla = rb + 3 * rs, R.inv = rb + 4 * rs;
  R.lstart = sm->lstart;
  for (int i = 0; i < NB_NPOL; i++)
  {
    const double* src = a.cl + ((size_t)b * NB_NPOL + i) * a.LS * 3;
    const int l0 = i < n ? sm->lstart[i] : 0, l1 = i < n ? sm->lstart[i + 1] : 0;
    if (in_smem)
    {
      for (int q = lane; q < 3 * (l1 - l0); q += NT) sm->cl[3 * l0 + q] = src[q];
      for (int q = lane; q < l1 - l0; q += NT) sm->l2i[l0 + q] = (unsigned char)i;
      if (lane == 0) sm->clb[i] = sm->cl;
    }
    else if (lane == 0)
      sm->clb[i] = src - 3 * (ptrdiff_t)l0;
  }
  R.clb = sm->clb;
  R.l2i = in_smem ? sm->l2i : nullptr;
  __syncthreads();
  int status = NB_STATUS_FAILED, it0 = 0, it1 = 0;
  double obj = 0.0;
  bool ok = false;
  for (int mode = 0; mode < 2 && !ok; mode++)
  {
    if (mode == 1)
    {  // fallback table: the first one is not read any more; order those reads before the asynchronous overwrite
      __syncthreads();
      if (lane == 0)
      {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        nb_bulk_load(&sm->tb, tables + NB_NPOL + (n - 1), (unsigned)sizeof(NbQpTable), &sm->mbar);
      }
    }
    nb_mbar_wait(&sm->mbar, (unsigned)mode);
    ok = nb_qp_solve<NT>(g, cs, &sm->tb, &sm->sh, R, ci, nl, sm->xout, mode == 0 ? &it0 : &it1, &obj,
                         a.prof ? a.prof + (size_t)b * 16 : nullptr);
    if (ok) status = mode == 0 ? NB_STATUS_OK : NB_STATUS_FALLBACK;
  }
  __syncthreads();
  // copy the solution (:866-876) or keep the initial path (:858); z override (:879-880)
  const double T = cs.T;
  double pfx = 0, pfy = 0;
  {
    const double qp[4] = { T * T * T, T * T, T, 1.0 };
    for (int r = 0; r < 4; r++)
    {
      pfx += qp[r] * ci[4 * (n - 1) + r];
      pfy += qp[r] * ci[32 + 4 * (n - 1) + r];
    }
  }
  const double dx = ci[3] - pfx, dy = ci[32 + 3] - pfy;
  const bool keep_z = sqrt(dx * dx + dy * dy) < 1.0;
  double* co = a.coeff_out + (size_t)b * 96;
  for (int q = lane; q < 96; q += NT)
  {
    const int ax = q / 32, r = q % 32;
    double v = ci[q];
    if (ok && r < 4 * n && !(ax == 2 && keep_z)) v = sm->xout[q];
    co[q] = v;
  }
  if (lane == 0)
  {
    a.obj[b] = ok ? obj : 0.0;
    a.status[b] = status;
    a.iters[2 * b] = it0;
    a.iters[2 * b + 1] = it1;
  }
}

__global__ void k_separate(int L, const int64_t* a_ptr, const double* a_xy, const int64_t* b_ptr, const double* b_xy,
                           int a_polygon, double* out_line, uint8_t* out_ok)
{
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double o[3];
  const bool ok = nb_separate(a_xy + 2 * a_ptr[l], (int)(a_ptr[l + 1] - a_ptr[l]), a_polygon != 0, b_xy + 2 * b_ptr[l],
                              (int)(b_ptr[l + 1] - b_ptr[l]), o);
  out_line[3 * l] = o[0];
  out_line[3 * l + 1] = o[1];
  out_line[3 * l + 2] = o[2];
  out_ok[l] = ok ? 1 : 0;
}

// PolySolverGurobi::generatePwpOut sampling loop (solver_gurobi_poly.cpp:911-934).  Operation order as written there
// (power vectors first, then coefficients times powers summed from the left), every operation rounded on its own, so the
// samples are bit-identical to the reference's.
__global__ void k_traj(int B, const int* n_int, const double* coeff, double T, double dc, int max_states, double* states,
                       int* n_states)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = n_int[b];
  const double* cf = coeff + (size_t)b * 96;
  double t = 0;
  int i = 0, cnt = 0;
  while (i < n && cnt < max_states)
  {
    const double dt = NB_SUB(t, NB_MUL((double)i, T));
    const double tp1 = NB_MUL(dt, dt), tp0 = NB_MUL(tp1, dt);
    const double tv0 = NB_MUL(NB_MUL(3.0, dt), dt), tv1 = NB_MUL(2.0, dt), ta0 = NB_MUL(6.0, dt);
    double* st = states + ((size_t)b * max_states + cnt) * 12;
    for (int ax = 0; ax < 3; ax++)
    {
      const double* c = cf + ax * 32 + 4 * i;
      st[ax] = NB_ADD(NB_ADD(NB_ADD(NB_MUL(c[0], tp0), NB_MUL(c[1], tp1)), NB_MUL(c[2], dt)), c[3]);
      st[3 + ax] = NB_ADD(NB_ADD(NB_MUL(c[0], tv0), NB_MUL(c[1], tv1)), c[2]);
      st[6 + ax] = NB_ADD(NB_MUL(c[0], ta0), NB_MUL(c[1], 2.0));
      st[9 + ax] = NB_MUL(c[0], 6.0);
    }
    cnt++;
    t = NB_ADD(t, dc);
    if (t > NB_MUL((double)(i + 1), T)) i++;
  }
  n_states[b] = cnt;
}

// one warp per agent; in worlds with hundreds of tethers four warps (N + M >= 256) or sixteen (>= 1024), lanes over the tethers
template <int NT>
__global__ void __launch_bounds__(NT) k_entangle(NbEntArgs a)
{
  extern __shared__ int smem_i[];
  Group<NT> g(threadIdx.x);
  nb_entangle_task<NT>(g, blockIdx.x, a, smem_i, smem_i + 4 * a.tcap);
}

// ---- K1 / K5 kernels
// One CTA of eight warps per (b, j): warp i builds the hull of window i (nb_hull_of_window_warp), then the threads take one
// sample each of Neptune::SamplePointsOfCurves for the same trajectory (the record is read once per CTA).
__global__ void __launch_bounds__(32 * NB_NPOL) k_hulls(NbConsts cs, int B, const double* t_start, const double* recs,
                                                        const uint8_t* known, double delta, double* hull_xy, int* hull_cnt,
                                                        int64_t* hull_ptr, double* nih0, int* idx, double* samp, int* err)
{
  __shared__ double scr[NB_NPOL][NB_HULL_WARP_SCRATCH];
  __shared__ double rec_s[NB_REC_PWP];
  const size_t bj = blockIdx.x;   // (b, j)
  const int j = (int)(bj % cs.N), b = (int)(bj / cs.N);
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool kn = known[bj] != 0;
  const double ts = t_start[b];
  if (kn)
    for (int q = threadIdx.x; q < NB_REC_PWP; q += blockDim.x) rec_s[q] = recs[(size_t)j * NB_REC + q];
  __syncthreads();
  const size_t k = bj * NB_NPOL + i;
  int cnt = 0, id2[2] = { -1, -1 };
  double n0[2];
  n0[0] = n0[1] = __longlong_as_double(0x7ff8000000000000LL);
  if (i < cs.num_pol && kn)
  {
    const double t0 = NB_ADD(ts, NB_MUL((double)i, cs.T)), t1 = NB_ADD(ts, NB_MUL((double)(i + 1), cs.T));
    cnt = nb_hull_of_window_warp(cs, rec_s, t0, t1, delta, hull_xy + k * NB_HMAX * 2, n0, id2, scr[i], lane);
    if (cnt < 0 || cnt > NB_HMAX)
    {
      if (lane == 0) *err = 3;
      cnt = 0;
    }
  }
  if (lane == 0)
  {
    hull_ptr[k] = (int64_t)k * NB_HMAX;
    hull_cnt[k] = cnt;
    nih0[2 * k] = n0[0];
    nih0[2 * k + 1] = n0[1];
    if (idx)
    {
      idx[2 * k] = id2[0];
      idx[2 * k + 1] = id2[1];
    }
  }
  if (!samp) return;
  const int S1 = cs.S + 1, ns = cs.num_pol * S1;
  double* out = samp + bj * ns * 2;
  for (int q = threadIdx.x; q < ns; q += blockDim.x)
  {
    if (!kn)
      out[2 * q] = 0.0, out[2 * q + 1] = 0.0;
    else
      nb_sample_one(rec_s, ts, NB_ADD(ts, NB_MUL(cs.T, (double)cs.num_pol)), cs.num_pol, cs.S, q / S1, q % S1, out + 2 * q);
  }
}

__global__ void k_postcheck(NbConsts cs, int B, const int* n_int, const double* coeff, const double* t_start,
                            const double* recs, const uint8_t* late, double delta, int* collide, int* err)
{
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (b, j, i)
  if (k >= (size_t)B * cs.N * NB_NPOL) return;
  const int i = (int)(k % NB_NPOL);
  const size_t bj = k / NB_NPOL;
  const int j = (int)(bj % cs.N), b = (int)(bj / cs.N);
  if (!late[bj] || i >= n_int[b]) return;
  const int r = nb_pwp_collides_interval(cs, coeff + (size_t)b * 96, n_int[b], i, t_start[b], recs + (size_t)j * NB_REC, delta);
  if (r < 0) *err = 3;
  if (r > 0) atomicOr(collide + b, 1);
}

__global__ void k_hull_index(int B, int N, const int* agent_id, const int* group, const uint8_t* known,
                             const int* hull_cnt_g, int64_t* hull_ptr, int* hull_cnt)
{
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (b, j, i)
  if (k >= (size_t)B * N * NB_NPOL) return;
  const int i = (int)(k % NB_NPOL);
  const size_t bj = k / NB_NPOL;
  const int j = (int)(bj % N), b = (int)(bj / N);
  const size_t src = ((size_t)group[b] * N + j) * NB_NPOL + i;
  hull_ptr[k] = (int64_t)src * NB_HMAX;
  hull_cnt[k] = (j == agent_id[b] - 1 || !known[bj]) ? 0 : hull_cnt_g[src];
}

// axis-aligned bounding box of every hull: (xmin, ymin, xmax, ymax); empty hulls get an empty box
__global__ void k_hull_aabb(size_t n, const double* hull_xy, const int* hull_cnt, double* aabb)
{
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int c = hull_cnt[k];
  const double* v = hull_xy + k * NB_HMAX * 2;
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int q = 0; q < c; q++)
  {
    x0 = fmin(x0, v[2 * q]), x1 = fmax(x1, v[2 * q]);
    y0 = fmin(y0, v[2 * q + 1]), y1 = fmax(y1, v[2 * q + 1]);
  }
  aabb[4 * k] = x0, aabb[4 * k + 1] = y0, aabb[4 * k + 2] = x1, aabb[4 * k + 3] = y1;
}

// aabb (nullable): boxes of the hulls.  Two convex sets whose boxes are disjoint are disjoint, which is what gjk::collision
// answers for them (gjk.cpp:76-148 finds a separating direction in its first iterations): those pairs -- almost all of
// them in a large world -- skip the GJK and the 384-byte hull read.
__global__ void k_postcheck_hulls(NbConsts cs, int B, const int* n_int, const double* coeff, const int* group,
                                  const double* hull_xy_g, const int* hull_cnt_g, const uint8_t* late, int* collide,
                                  const double* aabb)
{
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (b, j, i)
  if (k >= (size_t)B * cs.N * NB_NPOL) return;
  const int i = (int)(k % NB_NPOL);
  const size_t bj = k / NB_NPOL;
  const int j = (int)(bj % cs.N), b = (int)(bj / cs.N);
  if (!late[bj] || i >= n_int[b]) return;
  const size_t src = ((size_t)group[b] * cs.N + j) * NB_NPOL + i;
  const int hn = hull_cnt_g[src];
  if (hn <= 0) return;
  const double* cf = coeff + (size_t)b * 96;
  double A[8];
  for (int q = 0; q < 4; q++)
  {  // pointsA = P * A_rest_pos_basis_t_inverse_ (neptune.cpp:786-789)
    double x = 0, y = 0;
    for (int r = 0; r < 4; r++)
    {
      x = NB_ADD(x, NB_MUL(cf[4 * i + r], cs.Ainv[r * 4 + q]));
      y = NB_ADD(y, NB_MUL(cf[32 + 4 * i + r], cs.Ainv[r * 4 + q]));
    }
    A[2 * q] = x, A[2 * q + 1] = y;
  }
  if (aabb)
  {
    const double* bx = aabb + 4 * src;
    const double ax0 = fmin(fmin(A[0], A[2]), fmin(A[4], A[6])), ax1 = fmax(fmax(A[0], A[2]), fmax(A[4], A[6]));
    const double ay0 = fmin(fmin(A[1], A[3]), fmin(A[5], A[7])), ay1 = fmax(fmax(A[1], A[3]), fmax(A[5], A[7]));
    if (ax1 < bx[0] || bx[2] < ax0 || ay1 < bx[1] || bx[3] < ay0) return;
  }
  if (nb_gjk_collision(hull_xy_g + src * NB_HMAX * 2, hn, A, 4)) atomicOr(collide + b, 1);
}

__global__ void k_compose(int B, const double* t, const uint8_t* has_prev, const double* prev, const double* now,
                          double* out, int* n_pieces, int* err)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* nw = now + (size_t)b * NB_REC;
  double* o = out + (size_t)b * NB_REC;
  int np;
  if (!has_prev[b])
  {
    for (int q = 0; q < NB_REC; q++) o[q] = nw[q];
    np = (int)nw[0];
  }
  else
    np = nb_compose_records(t[b], prev + (size_t)b * NB_REC, nw, o);
  if (np < 0)
  {
    *err = 4;
    np = 0;
  }
  n_pieces[b] = np;
}

__global__ void k_commit(int B, const int* n_int, const double* coeff, const double* t_start, double T, double* recs,
                         const double* t_now, const double* prev, const int* prev_agent, const uint8_t* has_prev,
                         const int* status, const int* entangled, const int* collide, const int* fe_solved, int* n_pieces,
                         NbPublishHdr hd, int* err)
{
  __shared__ double now[NB_REC];
  const int b = blockIdx.x;
  if (b >= B) return;
  nb_commit_one(b, n_int, coeff, t_start, T, recs + (size_t)b * NB_REC, now, t_now, prev, prev_agent, has_prev, status, entangled,
                collide, fe_solved, n_pieces, hd, err);
}

// NeptuneRos::trajCB (neptune_ros.cpp:379-432), the bookkeeping a received DynTraj triggers besides
// Neptune::updateTrajObstacles: bendPtsForAgents_[id-1] = msg.bendpt, latestCheckingPosAgent_[id-1] = msg.pos.
// One thread per record.  bp_cnt [N], bp_xy [N][bp_max][2], latest_pos [N][2] (nullable).
__global__ void k_unpack_records(int N, int bp_max, const double* recs, int* bp_cnt, double* bp_xy, double* latest_pos, int* err)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const double* r = recs + (size_t)j * NB_REC;
  int nb = (int)r[NB_REC_NBEND];
  if (nb > bp_max)
  {
    *err = 6;
    nb = bp_max;
  }
  bp_cnt[j] = nb;
  for (int q = 0; q < bp_max; q++)
  {
    bp_xy[((size_t)j * bp_max + q) * 2] = q < nb ? r[NB_REC_BEND + 2 * q] : 0.0;
    bp_xy[((size_t)j * bp_max + q) * 2 + 1] = q < nb ? r[NB_REC_BEND + 2 * q + 1] : 0.0;
  }
  if (latest_pos) latest_pos[2 * j] = r[NB_REC_POS], latest_pos[2 * j + 1] = r[NB_REC_POS + 1];
}

// ------------------------------------------------------------------------------------------ ABI

extern "C" int nb_create(const nb_params* par, const double* pb, int device, nb_handle** out)
{
  if (!par || !pb || !out) return NB_ERR_ARG;
  if (par->num_pol < 1 || par->num_pol > NB_NPOL || par->deg_pol != 3 || par->use_linear_constraints != 1 ||
      par->num_agents < 1 || par->T_span <= 0)
  {
    g_err = "nb_create: unsupported parameters (num_pol<=8, deg_pol==3, use_linear_constraints==1)";
    return NB_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
  {
    g_err = "no CUDA device visible: libneptune_b200 has no CPU fallback";
    return NB_ERR_NO_DEVICE;
  }
  if (device < 0) NB_CUDA(cudaGetDevice(&device));
  NB_CUDA(cudaSetDevice(device));
  nb_handle* h = new nb_handle();
  h->par = *par;
  h->device = device;
  h->launches = 0;
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) h->num_sms = sms;
  }
  nb_build_consts(par, &h->cs);
  std::vector<NbQpTable> tabs(2 * NB_NPOL);
  for (int mode = 0; mode < 2; mode++)
    for (int n = 1; n <= NB_NPOL; n++)
      if (!nb_build_table(&h->cs, n, mode, &tabs[mode * NB_NPOL + n - 1]))
      {
        g_err = "nb_create: degenerate QP table";
        delete h;
        return NB_ERR_ARG;
      }
  h->d_tables = nullptr;
  h->d_pb = nullptr;
  h->d_st_ptr = nullptr;
  h->d_st_xy = nullptr;
  h->d_strep = nullptr;
  h->st_nvert = 0;
  const int rc = [&]() -> int {
    NB_CUDA(cudaMalloc(&h->d_tables, sizeof(NbQpTable) * tabs.size()));
    NB_CUDA(cudaMemcpy(h->d_tables, tabs.data(), sizeof(NbQpTable) * tabs.size(), cudaMemcpyHostToDevice));
    NB_CUDA(cudaMalloc(&h->d_pb, sizeof(double) * 2 * par->num_agents));
    NB_CUDA(cudaMemcpy(h->d_pb, pb, sizeof(double) * 2 * par->num_agents, cudaMemcpyHostToDevice));
    NB_CUDA(cudaMalloc(&h->d_st_ptr, sizeof(int64_t) * (par->num_static + 1)));
    NB_CUDA(cudaMemset(h->d_st_ptr, 0, sizeof(int64_t) * (par->num_static + 1)));
    if (h->err.ensure(sizeof(int)))
    {
      g_err = "cudaMalloc failed";
      return NB_ERR_CUDA;
    }
    NB_CUDA(cudaMemset(h->err.p, 0, sizeof(int)));
    return NB_OK;
  }();
  if (rc != NB_OK)
  {  // nothing leaks on a failed construction
    nb_destroy(h);
    return rc;
  }
  *out = h;
  return NB_OK;
}

extern "C" void nb_destroy(nb_handle* h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->d_tables);
  cudaFree(h->d_pb);
  cudaFree(h->d_st_ptr);
  cudaFree(h->d_st_xy);
  cudaFree(h->d_strep);
  for (auto& b : h->in) b.release();
  for (auto& b : h->out) b.release();
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  h->lines.release(), h->line_ok.release(), h->keep.release(), h->cl.release(), h->ncl.release(), h->rows.release(), h->err.release(), h->ent_scratch.release();
  cudaFree(h->d_st_longest);
  for (auto& b : h->sin) b.release();
  for (auto& b : h->sout) b.release();
  h->sw_meta.release(), h->sw_kin.release(), h->sw_alpha.release(), h->sw_beta.release(), h->sw_bend.release();
  h->sprof.release(), h->qprof.release();
  h->sw_ng.release(), h->sw_fcode.release();
  h->sw_hash.release(), h->sw_heap.release(), h->sw_gh.release(), h->sw_chi.release(), h->sw_chd.release();
  delete h;
}

extern "C" long long nb_launch_count(const nb_handle* h) { return h ? h->launches : 0; }

extern "C" int nb_set_profiling(nb_handle* h, int on)
{
  if (!h) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  if (on && !h->ev[0])
    for (auto& e : h->ev) NB_CUDA(cudaEventCreate(&e));
  h->profiling = on ? 1 : 0;
  return NB_OK;
}

extern "C" int nb_kernel_times(nb_handle* h, double* ms, int n)
{
  if (!h || !ms || !h->ev[0]) return NB_ERR_ARG;
  NB_CUDA(cudaEventSynchronize(h->ev[2]));
  for (int k = 0; k < n && k < 2; k++)
  {
    float f = 0.f;
    NB_CUDA(cudaEventElapsedTime(&f, h->ev[k], h->ev[k + 1]));
    ms[k] = f;
  }
  return NB_OK;
}

extern "C" int nb_check_async_errors(nb_handle* h, void* stream)
{
  if (!h) return NB_ERR_ARG;
  if (!h->err.p) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  int err = 0;
  NB_CUDA(cudaMemcpyAsync(&err, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  NB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  NB_CUDA(cudaMemsetAsync(h->err.p, 0, sizeof(int), (cudaStream_t)stream));
  if (err == 5)
  {
    g_err = "nb_replan_batch: n_int out of range (1..8) in a device-resident batch";
    return NB_ERR_ARG;
  }
  if (err == 6)
  {
    g_err = "a tether has more bend points than a record (8) or bp_max can hold";
    return NB_ERR_CAPACITY;
  }
  if (err == 7)
  {
    g_err = "nb_cycle: the records of a peer rank did not arrive (exchange timeout)";
    return NB_ERR_CUDA;
  }
  if (err)
  {
    g_err = "a fixed-capacity list overflowed in an asynchronous call (ent_slots, ent_cap or a 16-piece record)";
    return NB_ERR_CAPACITY;
  }
  return NB_OK;
}

extern "C" int nb_kept_lines(nb_handle* h, int32_t* out, int32_t B)
{
  if (!h || !out || B < 1 || !h->ncl.p || (size_t)B * NB_NPOL * sizeof(int) > h->ncl.cap) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  NB_CUDA(cudaDeviceSynchronize());
  NB_CUDA(cudaMemcpy(out, h->ncl.p, (size_t)B * NB_NPOL * sizeof(int), cudaMemcpyDeviceToHost));
  return NB_OK;
}

extern "C" int nb_line_slots(const nb_handle* h, int n_hull_slots)
{
  return n_hull_slots + h->par.num_agents + h->par.num_static + h->par.ent_slots;
}

extern "C" int nb_set_static(nb_handle* h, const int64_t* st_ptr, const double* st_xy, const double* strep)
{
  if (!h) return NB_ERR_ARG;
  const int M = h->par.num_static;
  if (M == 0) return NB_OK;
  if (!st_ptr || !st_xy) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  h->st_nvert = st_ptr[M];
  cudaFree(h->d_st_xy);
  cudaFree(h->d_strep);
  h->d_st_xy = nullptr;
  h->d_strep = nullptr;
  NB_CUDA(cudaMemcpy(h->d_st_ptr, st_ptr, sizeof(int64_t) * (M + 1), cudaMemcpyHostToDevice));
  NB_CUDA(cudaMalloc(&h->d_st_xy, sizeof(double) * 2 * h->st_nvert));
  NB_CUDA(cudaMemcpy(h->d_st_xy, st_xy, sizeof(double) * 2 * h->st_nvert, cudaMemcpyHostToDevice));
  h->strep_per_agent = 0;
  if (strep)
  {
    NB_CUDA(cudaMalloc(&h->d_strep, sizeof(double) * 4 * M));
    NB_CUDA(cudaMemcpy(h->d_strep, strep, sizeof(double) * 4 * M, cudaMemcpyHostToDevice));
  }
  return NB_OK;
}

extern "C" int nb_set_static_rep_per_agent(nb_handle* h, const double* strep_all, const double* longest_all)
{
  if (!h) return NB_ERR_ARG;
  const int M = h->par.num_static, N = h->par.num_agents;
  if (M == 0) return NB_OK;
  if (!strep_all) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  cudaFree(h->d_strep);
  h->d_strep = nullptr;
  NB_CUDA(cudaMalloc(&h->d_strep, sizeof(double) * 4 * M * N));
  NB_CUDA(cudaMemcpy(h->d_strep, strep_all, sizeof(double) * 4 * M * N, cudaMemcpyHostToDevice));
  if (longest_all)
  {
    cudaFree(h->d_st_longest);
    h->d_st_longest = nullptr;
    NB_CUDA(cudaMalloc(&h->d_st_longest, sizeof(double) * 2 * M * N));
    NB_CUDA(cudaMemcpy(h->d_st_longest, longest_all, sizeof(double) * 2 * M * N, cudaMemcpyHostToDevice));
  }
  h->strep_per_agent = 1;
  return NB_OK;
}

namespace
{
// stage one NB_HOST input array (or pass an NB_DEVICE pointer through)
template <typename T>
int stage_in(nb_handle* h, int slot, int space, const T* src, size_t count, cudaStream_t st, const T** dst)
{
  if (space == NB_DEVICE || src == nullptr || count == 0)
  {
    *dst = src;
    return NB_OK;
  }
  if (h->in[slot].ensure(count * sizeof(T)))
  {
    g_err = "cudaMalloc failed while staging inputs";
    return NB_ERR_CUDA;
  }
  NB_CUDA(cudaMemcpyAsync(h->in[slot].p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
  *dst = (const T*)h->in[slot].p;
  return NB_OK;
}
template <typename T>
int stage_out(nb_handle* h, int slot, int space, T* user, size_t count, T** dst)
{
  if (space == NB_DEVICE || user == nullptr)
  {
    *dst = user;
    return NB_OK;
  }
  if (h->out[slot].ensure(count * sizeof(T)))
  {
    g_err = "cudaMalloc failed while staging outputs";
    return NB_ERR_CUDA;
  }
  *dst = (T*)h->out[slot].p;
  return NB_OK;
}
}  // namespace

extern "C" int nb_replan_batch(nb_handle* h, const nb_replan_args* a, void* stream)
{
  if (!h || !a || a->B < 0) return NB_ERR_ARG;
  if (a->B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->B, N = h->par.num_agents, M = h->par.num_static, NA = N + M, cap = h->par.ent_cap;
  const int NH = a->n_hull_slots, LS = nb_line_slots(h, NH);
  const int sp = a->space;
  if (!a->agent_id || !a->n_int || !a->coeff_init || !a->coeff_out || !a->obj || !a->status || !a->iters)
  {
    g_err = "nb_replan_batch: null argument";
    return NB_ERR_ARG;
  }
  if (sp == NB_HOST)
  {  // host arrays are checked before anything is launched; device arrays are checked by k_lines / k_qp (h->err)
    for (int b = 0; b < B; b++)
    {
      if (a->n_int[b] < 1 || a->n_int[b] > h->par.num_pol)
      {
        g_err = "nb_replan_batch: n_int out of range (1..num_pol)";
        return NB_ERR_ARG;
      }
      if (a->agent_id[b] < 1 || a->agent_id[b] > N)
      {
        g_err = "nb_replan_batch: agent_id out of range (1..num_agents)";
        return NB_ERR_ARG;
      }
    }
  }
  int rc;
  NbLinesIn in;
  in.NH = NH;
  if ((rc = stage_in(h, 0, sp, a->agent_id, (size_t)B, st, &in.agent_id))) return rc;
  if ((rc = stage_in(h, 1, sp, a->n_int, (size_t)B, st, &in.n_int))) return rc;
  if ((rc = stage_in(h, 2, sp, a->coeff_init, (size_t)B * 96, st, &in.coeff_init))) return rc;
  if ((rc = stage_in(h, 3, sp, a->hull_ptr, a->hull_known ? 0 : (size_t)B * NH * 8 + (a->hull_cnt ? 0 : 1), st, &in.hull_ptr))) return rc;
  if ((rc = stage_in(h, 4, sp, a->hull_xy, (size_t)a->hull_nvert * 2, st, &in.hull_xy))) return rc;
  if ((rc = stage_in(h, 11, sp, a->hull_cnt, (size_t)B * NH * 8, st, &in.hull_cnt))) return rc;
  if (a->nih0_group && sp == NB_HOST)
  {
    g_err = "nb_replan_batch: nih0_group is only supported with device pointers";
    return NB_ERR_ARG;
  }
  in.nih0_group = a->nih0_group;
  in.hull_known = a->hull_known;
  if (a->hull_known && (!a->nih0_group || !a->hull_cnt || NH != N))
  {
    g_err = "nb_replan_batch: hull_known needs nih0_group, hull_cnt and n_hull_slots == num_agents";
    return NB_ERR_ARG;
  }
  if ((rc = stage_in(h, 5, sp, a->nih0, (size_t)B * N * 16, st, &in.nih0))) return rc;
  if ((rc = stage_in(h, 6, sp, a->esv_cnt, (size_t)B * 18, st, &in.esv_cnt))) return rc;
  if ((rc = stage_in(h, 7, sp, a->esv_alpha, (size_t)B * 9 * cap * 2, st, &in.esv_alpha))) return rc;
  if ((rc = stage_in(h, 8, sp, a->esv_active, (size_t)B * 9 * NA, st, &in.esv_active))) return rc;
  if ((rc = stage_in(h, 9, sp, a->bp_cnt, (size_t)N, st, &in.bp_cnt))) return rc;
  if ((rc = stage_in(h, 10, sp, a->bp_xy, (size_t)N * h->par.bp_max * 2, st, &in.bp_xy))) return rc;
  in.st_ptr = h->d_st_ptr;
  in.st_xy = h->d_st_xy;
  in.pb = h->d_pb;
  if (M > 0 && !h->d_st_xy)
  {
    g_err = "nb_replan_batch: num_static > 0 but nb_set_static was never called";
    return NB_ERR_ARG;
  }
  // scratch
  const size_t nslots = (size_t)B * NB_NPOL * LS;
  const int RS = NB_ROW_LINE0 + 4 * NB_NPOL * LS;
  if (h->lines.ensure(nslots * 3 * sizeof(double)) || h->line_ok.ensure(nslots) ||
      h->cl.ensure(nslots * 3 * sizeof(double)) || h->keep.ensure(nslots) || h->ncl.ensure((size_t)B * NB_NPOL * sizeof(int)) ||
      h->rows.ensure((size_t)B * 5 * RS * sizeof(double)) || h->err.ensure(sizeof(int)))
  {
    g_err = "cudaMalloc failed for scratch buffers";
    return NB_ERR_CUDA;
  }
  if (sp == NB_HOST) NB_CUDA(cudaMemsetAsync(h->err.p, 0, sizeof(int), st));
  NbQpArgs q;
  q.n_int = in.n_int;
  q.coeff_init = in.coeff_init;
  q.cl = (const double*)h->cl.p;
  q.ncl = (const int*)h->ncl.p;
  q.LS = LS;
  q.rows = (double*)h->rows.p;
  q.RS = RS;
  q.err = (int*)h->err.p;
  q.prof = nullptr;
  if (h->profiling)
  {
    if (h->qprof.ensure((size_t)B * 16 * sizeof(long long)))
    {
      g_err = "cudaMalloc failed";
      return NB_ERR_CUDA;
    }
    NB_CUDA(cudaMemsetAsync(h->qprof.p, 0, (size_t)B * 16 * sizeof(long long), st));
    q.prof = (long long*)h->qprof.p;
    h->qprof_B = B;
  }
  if ((rc = stage_out(h, 0, sp, a->coeff_out, (size_t)B * 96, &q.coeff_out))) return rc;
  if ((rc = stage_out(h, 1, sp, a->obj, (size_t)B, &q.obj))) return rc;
  if ((rc = stage_out(h, 2, sp, a->status, (size_t)B, &q.status))) return rc;
  if ((rc = stage_out(h, 3, sp, a->iters, (size_t)B * 2, &q.iters))) return rc;

  if (h->profiling) cudaEventRecord(h->ev[0], st);
  const size_t lsm = lines_smem_bytes(LS);
  if (!h->qp_smem_set)
  {
    NB_CUDA(cudaFuncSetAttribute(k_qp<NB_QP_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NbQpSmem)));
    NB_CUDA(cudaFuncSetAttribute(k_qp<NB_QP_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NbQpSmem)));
    NB_CUDA(cudaFuncSetAttribute(k_lines<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    NB_CUDA(cudaFuncSetAttribute(k_lines<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    h->qp_smem_set = 1;
  }
  if (lsm > 200 * 1024)
  {
    g_err = "nb_replan_batch: too many line slots per interval for the pruning stage";
    return NB_ERR_CAPACITY;
  }
  if (LS >= 1024 && B * NB_NPOL <= 4 * h->num_sms)   // a grid that leaves the GPU half empty: more threads per (agent, interval)
    k_lines<512><<<B * NB_NPOL, 512, lsm, st>>>(h->cs, in, LS, (double*)h->lines.p, (uint8_t*)h->line_ok.p,
                                                (uint8_t*)h->keep.p, (int*)h->err.p, (double*)h->cl.p, (int*)h->ncl.p);
  else
    k_lines<128><<<B * NB_NPOL, 128, lsm, st>>>(h->cs, in, LS, (double*)h->lines.p, (uint8_t*)h->line_ok.p,
                                                (uint8_t*)h->keep.p, (int*)h->err.p, (double*)h->cl.p, (int*)h->ncl.p);
  if (h->profiling) cudaEventRecord(h->ev[1], st);
  if (B <= h->num_sms && !getenv("NB_QP_NARROW"))   // the batch leaves SMs idle: eight warps per agent, one row per thread in every sweep
    k_qp<NB_QP_THREADS_WIDE><<<B, NB_QP_THREADS_WIDE, sizeof(NbQpSmem), st>>>(h->cs, h->d_tables, q);
  else
    k_qp<NB_QP_THREADS><<<B, NB_QP_THREADS, sizeof(NbQpSmem), st>>>(h->cs, h->d_tables, q);
  if (h->profiling) cudaEventRecord(h->ev[2], st);
  h->launches += 2;
  NB_CUDA(cudaGetLastError());

  if (sp == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(a->coeff_out, q.coeff_out, (size_t)B * 96 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(a->obj, q.obj, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(a->status, q.status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(a->iters, q.iters, (size_t)B * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  if (a->lines)
    NB_CUDA(cudaMemcpyAsync(a->lines, h->lines.p, nslots * 3 * sizeof(double),
                            sp == NB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  if (a->line_ok)
    NB_CUDA(cudaMemcpyAsync(a->line_ok, h->line_ok.p, nslots,
                            sp == NB_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  if (sp == NB_HOST)
  {
    int err = 0;
    NB_CUDA(cudaMemcpyAsync(&err, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
    if (err)
    {
      g_err = "nb_replan_batch: more non-entangling LPs than ent_slots in some interval";
      return NB_ERR_CAPACITY;
    }
  }
  return NB_OK;
}

extern "C" int nb_separate_batch(nb_handle* h, int32_t L, int32_t space, const int64_t* a_ptr, const double* a_xy,
                                 const int64_t* b_ptr, const double* b_xy, int32_t a_polygon, double* out_line,
                                 uint8_t* out_ok, void* stream)
{
  if (!h || L < 0) return NB_ERR_ARG;
  if (L == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t *dap, *dbp;
  const double *dax, *dbx;
  double* dline;
  uint8_t* dok;
  int rc;
  size_t na = 0, nb = 0;
  if (space == NB_HOST)
  {
    na = (size_t)a_ptr[L];
    nb = (size_t)b_ptr[L];
  }
  if ((rc = stage_in(h, 0, space, a_ptr, (size_t)L + 1, st, &dap))) return rc;
  if ((rc = stage_in(h, 1, space, a_xy, na * 2, st, &dax))) return rc;
  if ((rc = stage_in(h, 2, space, b_ptr, (size_t)L + 1, st, &dbp))) return rc;
  if ((rc = stage_in(h, 3, space, b_xy, nb * 2, st, &dbx))) return rc;
  if ((rc = stage_out(h, 0, space, out_line, (size_t)L * 3, &dline))) return rc;
  if ((rc = stage_out(h, 1, space, out_ok, (size_t)L, &dok))) return rc;
  k_separate<<<(L + 127) / 128, 128, 0, st>>>(L, dap, dax, dbp, dbx, a_polygon, dline, dok);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(out_line, dline, (size_t)L * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(out_ok, dok, (size_t)L, cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
  }
  return NB_OK;
}

extern "C" int nb_generate_traj_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                                      double dc, int32_t max_states, double* states, int32_t* n_states, void* stream)
{
  if (!h || B < 0 || max_states <= 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int* dn;
  const double* dc_;
  double* dstates;
  int* dns;
  int rc;
  if ((rc = stage_in(h, 0, space, n_int, (size_t)B, st, &dn))) return rc;
  if ((rc = stage_in(h, 1, space, coeff, (size_t)B * 96, st, &dc_))) return rc;
  if ((rc = stage_out(h, 0, space, states, (size_t)B * max_states * 12, &dstates))) return rc;
  if ((rc = stage_out(h, 1, space, n_states, (size_t)B, &dns))) return rc;
  k_traj<<<(B + 63) / 64, 64, 0, st>>>(B, dn, dc_, h->cs.T, dc, max_states, dstates, dns);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(states, dstates, (size_t)B * max_states * 12 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(n_states, dns, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
  }
  return NB_OK;
}

// ------------------------------------------------------------------------------------------ K3 ABI
namespace
{
struct EntStage
{
  nb_ent_state dev;
};

int stage_state_in(nb_handle* h, int slot0, int space, const nb_ent_state& u, size_t nstates, cudaStream_t st,
                   nb_ent_state* d, bool copy)
{
  const int cap = h->par.ent_cap, NA = h->par.num_agents + h->par.num_static;
  if (space == NB_DEVICE)
  {
    *d = u;
    return NB_OK;
  }
  const size_t sz[5] = { nstates * 2 * sizeof(int), nstates * cap * 2 * sizeof(int), nstates * cap * sizeof(double),
                         nstates * cap * sizeof(int), nstates * NA * sizeof(int) };
  void* up[5] = { u.cnt, u.alpha, u.beta, u.bend, u.active };
  void* dp[5];
  for (int k = 0; k < 5; k++)
  {
    if (h->in[slot0 + k].ensure(sz[k]))
    {
      g_err = "cudaMalloc failed while staging entanglement state";
      return NB_ERR_CUDA;
    }
    dp[k] = h->in[slot0 + k].p;
    if (copy) NB_CUDA(cudaMemcpyAsync(dp[k], up[k], sz[k], cudaMemcpyHostToDevice, st));
  }
  d->cnt = (int32_t*)dp[0], d->alpha = (int32_t*)dp[1], d->beta = (double*)dp[2], d->bend = (int32_t*)dp[3];
  d->active = (int32_t*)dp[4];
  return NB_OK;
}

int state_out(nb_handle* h, int space, const nb_ent_state& u, const nb_ent_state& d, size_t nstates, cudaStream_t st)
{
  if (space == NB_DEVICE) return NB_OK;
  const int cap = h->par.ent_cap, NA = h->par.num_agents + h->par.num_static;
  NB_CUDA(cudaMemcpyAsync(u.cnt, d.cnt, nstates * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaMemcpyAsync(u.alpha, d.alpha, nstates * cap * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaMemcpyAsync(u.beta, d.beta, nstates * cap * sizeof(double), cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaMemcpyAsync(u.bend, d.bend, nstates * cap * sizeof(int), cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaMemcpyAsync(u.active, d.active, nstates * NA * sizeof(int), cudaMemcpyDeviceToHost, st));
  return NB_OK;
}

int ent_common(nb_handle* h, NbEntArgs* a, int mode, int B, int space, const int32_t* agent_id, const uint8_t* known,
               const int32_t* bp_cnt, const double* bp_xy, cudaStream_t st)
{
  const int N = h->par.num_agents, M = h->par.num_static;
  memset(a, 0, sizeof(*a));
  a->mode = mode, a->N = N, a->M = M, a->cap = h->par.ent_cap, a->bp_max = h->par.bp_max;
  a->num_pol = h->par.num_pol, a->S = h->par.samples, a->T = h->par.T_span;
  int tcap = 4 * (N + M) + 16;
  a->tcap = tcap > 1024 ? 1024 : tcap;
  a->pb = h->d_pb;
  a->strep = h->d_strep;
  a->strep_per_agent = h->strep_per_agent;
  if (M > 0 && !h->d_strep)
  {
    g_err = "entanglement call with static obstacles but nb_set_static(strep) was never called";
    return NB_ERR_ARG;
  }
  int rc;
  if ((rc = stage_in(h, 0, space, agent_id, (size_t)B, st, &a->agent_id))) return rc;
  if ((rc = stage_in(h, 1, space, known, (size_t)B * N, st, &a->known))) return rc;
  if ((rc = stage_in(h, 2, space, bp_cnt, (size_t)N, st, &a->bp_cnt))) return rc;
  if ((rc = stage_in(h, 3, space, bp_xy, (size_t)N * h->par.bp_max * 2, st, &a->bp_xy))) return rc;
  a->err = (int*)h->err.p;
  return NB_OK;
}

int ent_launch(nb_handle* h, const NbEntArgs& a_in, int B, cudaStream_t st)
{
  NbEntArgs a = a_in;
  size_t sm = (size_t)(4 * a.tcap + 4) * sizeof(int);
  a.stage = 0;
  if (a.mode == 4)
  {  // the post-check's work state in shared memory when it fits
    const size_t extra = nb_ent_stage_bytes(a.N + a.M, a.cap);
    if (sm + extra <= 200 * 1024) a.stage = 1, sm += extra;
  }
  // few agents, many tethers: sixteen warps per agent (512 tethers: measured on rank 0's share of the 8-GPU world)
  if (a.N + a.M >= 512 && a.mode != 3 && B <= 2 * h->num_sms)
  {
    if (sm > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(k_entangle<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_entangle<512><<<B, 512, sm, st>>>(a);
  }
  else if (a.N + a.M >= 256 && a.mode != 3)
  {
    if (sm > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(k_entangle<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_entangle<128><<<B, 128, sm, st>>>(a);
  }
  else
  {
    if (sm > 48 * 1024) NB_CUDA(cudaFuncSetAttribute(k_entangle<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_entangle<32><<<B, 32, sm, st>>>(a);
  }
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}

int ent_finish(nb_handle* h, int space, cudaStream_t st)
{
  if (space != NB_HOST) return NB_OK;
  int err = 0;
  NB_CUDA(cudaMemcpyAsync(&err, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaStreamSynchronize(st));
  if (err)
  {
    NB_CUDA(cudaMemsetAsync(h->err.p, 0, sizeof(int), st));
    g_err = "entanglement chain: fixed-capacity list overflow (ent_cap / crossing buffer)";
    return NB_ERR_CAPACITY;
  }
  return NB_OK;
}
}  // namespace

extern "C" int nb_entangle_predict_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id,
                                         const uint8_t* known, const int32_t* bp_cnt, const double* bp_xy,
                                         nb_ent_state stt, const double* prev_pos, const double* prev_pos_agent,
                                         const double* cur, const double* samp0, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents;
  NbEntArgs a;
  int rc;
  if ((rc = ent_common(h, &a, 0, B, space, agent_id, known, bp_cnt, bp_xy, st))) return rc;
  nb_ent_state d;
  if ((rc = stage_state_in(h, 4, space, stt, (size_t)B, st, &d, true))) return rc;
  a.st = d;
  if ((rc = stage_in(h, 9, space, prev_pos, (size_t)B * (N + 1) * 2, st, &a.prev_pos))) return rc;
  if ((rc = stage_in(h, 10, space, prev_pos_agent, (size_t)B * N * 2, st, &a.prev_pos_agent))) return rc;
  if ((rc = stage_in(h, 11, space, cur, (size_t)B * 2, st, &a.cur))) return rc;
  if ((rc = stage_in(h, 12, space, samp0, (size_t)B * N * 2, st, &a.samp0))) return rc;
  if ((rc = ent_launch(h, a, B, st))) return rc;
  if ((rc = state_out(h, space, stt, d, (size_t)B, st))) return rc;
  return ent_finish(h, space, st);
}

// internal (nb_cycle.cu): PredictAlphasBetas with SampledPointsForAll[j][0].col(0) read in place from the group-shaped
// samples of nb_hulls_batch (agent b looks at block group[b]); device pointers
int nb_internal_predict_grouped(nb_handle* h, int B, const int32_t* agent_id, const uint8_t* known, const int32_t* bp_cnt,
                                const double* bp_xy, nb_ent_state stt, const double* prev_pos, const double* prev_pos_agent,
                                const double* cur, const double* samp_g, const int32_t* group, cudaStream_t st)
{
  NbEntArgs a;
  int rc;
  if ((rc = ent_common(h, &a, 0, B, NB_DEVICE, agent_id, known, bp_cnt, bp_xy, st))) return rc;
  a.st = stt;
  a.prev_pos = prev_pos, a.prev_pos_agent = prev_pos_agent, a.cur = cur;
  a.samp0 = samp_g, a.samp_group = group, a.samp0_stride = h->par.num_pol * (h->par.samples + 1) * 2;
  return ent_launch(h, a, B, st);
}

extern "C" int nb_entangle_track_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const int32_t* bp_cnt,
                                       const double* bp_xy, const int32_t* bp_cnt_prev, const double* bp_xy_prev,
                                       nb_ent_state stt, double* prev_pos, double* prev_pos_agent,
                                       const double* latest_pos_agent, const double* cur, const double* elapsed_ms,
                                       int32_t* result, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents;
  NbEntArgs a;
  int rc;
  if ((rc = ent_common(h, &a, 3, B, space, agent_id, nullptr, bp_cnt, bp_xy, st))) return rc;
  nb_ent_state d;
  if ((rc = stage_state_in(h, 4, space, stt, (size_t)B, st, &d, true))) return rc;
  a.st = d;
  const double *pp = nullptr, *ppa = nullptr;
  if ((rc = stage_in(h, 9, space, (const double*)prev_pos, (size_t)B * (N + 1) * 2, st, &pp))) return rc;
  if ((rc = stage_in(h, 10, space, (const double*)prev_pos_agent, (size_t)B * N * 2, st, &ppa))) return rc;
  a.prev_pos_rw = (double*)pp, a.prev_pos_agent_rw = (double*)ppa;
  if ((rc = stage_in(h, 11, space, cur, (size_t)B * 2, st, &a.cur))) return rc;
  if ((rc = stage_in(h, 12, space, latest_pos_agent, (size_t)B * N * 2, st, &a.latest))) return rc;
  if ((rc = stage_in(h, 13, space, bp_cnt_prev, (size_t)N, st, &a.bp_cnt_prev))) return rc;
  if ((rc = stage_in(h, 14, space, bp_xy_prev, (size_t)N * h->par.bp_max * 2, st, &a.bp_xy_prev))) return rc;
  if ((rc = stage_in(h, 15, space, elapsed_ms, (size_t)B, st, &a.elapsed_ms))) return rc;
  if ((rc = stage_out(h, 7, space, result, (size_t)B, &a.result))) return rc;
  if ((rc = ent_launch(h, a, B, st))) return rc;
  if ((rc = state_out(h, space, stt, d, (size_t)B, st))) return rc;
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(prev_pos, pp, (size_t)B * (N + 1) * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(prev_pos_agent, ppa, (size_t)B * N * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(result, a.result, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  return ent_finish(h, space, st);
}

extern "C" int nb_entangle_rollout_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id,
                                         const uint8_t* known, const int32_t* bp_cnt, const double* bp_xy,
                                         nb_ent_state in, const int32_t* n_int, const double* coeff, const double* samp,
                                         int32_t samp_shared, nb_ent_state out, int32_t* done, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents, S = h->par.samples;
  NbEntArgs a;
  int rc;
  if ((rc = ent_common(h, &a, 1, B, space, agent_id, known, bp_cnt, bp_xy, st))) return rc;
  nb_ent_state din, dout;
  if ((rc = stage_state_in(h, 4, space, in, (size_t)B, st, &din, true))) return rc;
  a.st = din;
  if ((rc = stage_in(h, 9, space, n_int, (size_t)B, st, &a.n_int))) return rc;
  if ((rc = stage_in(h, 10, space, coeff, (size_t)B * 96, st, &a.coeff))) return rc;
  if ((rc = stage_in(h, 11, space, samp, (size_t)(samp_shared ? 1 : B) * N * h->par.num_pol * (S + 1) * 2, st, &a.samp)))
    return rc;
  a.samp_shared = samp_shared;
  // outputs: staged in out[] slots 2..6 + 7
  const int cap = h->par.ent_cap, NA = N + h->par.num_static;
  if (space == NB_HOST)
  {
    const size_t ns = (size_t)B * 9;
    const size_t sz[5] = { ns * 2 * sizeof(int), ns * cap * 2 * sizeof(int), ns * cap * sizeof(double),
                           ns * cap * sizeof(int), ns * NA * sizeof(int) };
    for (int k = 0; k < 5; k++)
      if (h->out[2 + k].ensure(sz[k]))
      {
        g_err = "cudaMalloc failed";
        return NB_ERR_CUDA;
      }
    dout.cnt = (int32_t*)h->out[2].p, dout.alpha = (int32_t*)h->out[3].p, dout.beta = (double*)h->out[4].p;
    dout.bend = (int32_t*)h->out[5].p, dout.active = (int32_t*)h->out[6].p;
  }
  else
    dout = out;
  a.out = dout;
  if ((rc = stage_out(h, 7, space, done, (size_t)B, &a.result))) return rc;
  if ((rc = ent_launch(h, a, B, st))) return rc;
  if ((rc = state_out(h, space, out, dout, (size_t)B * 9, st))) return rc;
  if (space == NB_HOST) NB_CUDA(cudaMemcpyAsync(done, a.result, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  return ent_finish(h, space, st);
}

extern "C" int nb_entangle_check_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id,
                                       const uint8_t* known, const int32_t* bp_cnt, const double* bp_xy, nb_ent_state stt,
                                       const int32_t* n_int, const double* coeff, const double* samp, int32_t samp_shared,
                                       int32_t* entangled, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents, S = h->par.samples;
  NbEntArgs a;
  int rc;
  if ((rc = ent_common(h, &a, 2, B, space, agent_id, known, bp_cnt, bp_xy, st))) return rc;
  nb_ent_state d;
  if ((rc = stage_state_in(h, 4, space, stt, (size_t)B, st, &d, true))) return rc;
  a.st = d;
  if ((rc = stage_in(h, 9, space, n_int, (size_t)B, st, &a.n_int))) return rc;
  if ((rc = stage_in(h, 10, space, coeff, (size_t)B * 96, st, &a.coeff))) return rc;
  if ((rc = stage_in(h, 11, space, samp, (size_t)(samp_shared ? 1 : B) * N * h->par.num_pol * (S + 1) * 2, st, &a.samp)))
    return rc;
  a.samp_shared = samp_shared;
  if ((rc = stage_out(h, 7, space, entangled, (size_t)B, &a.result))) return rc;
  if ((rc = ent_launch(h, a, B, st))) return rc;
  if ((rc = state_out(h, space, stt, d, (size_t)B, st))) return rc;
  if (space == NB_HOST) NB_CUDA(cudaMemcpyAsync(entangled, a.result, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  return ent_finish(h, space, st);
}

// ------------------------------------------------------------------------------------------ K1 / K5 ABI
extern "C" int nb_hulls_batch(nb_handle* h, int32_t B, int32_t space, const double* t_start, const double* recs,
                              const uint8_t* known, double delta, double* hull_xy, int32_t* hull_cnt, int64_t* hull_ptr,
                              double* nih0, double* samp, int32_t* idx, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  static_assert(NB_REC == NB_REC_DOUBLES && NB_HMAX == NB_HULL_STRIDE, "record layout");
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents, S = h->par.samples, P = h->par.num_pol;
  const size_t nh = (size_t)B * N * NB_NPOL;
  const double *dt, *dr;
  const uint8_t* dk;
  double *dxy, *dn0, *dsamp;
  int *dcnt, *didx;
  int64_t* dptr;
  int rc;
  if ((rc = stage_in(h, 0, space, t_start, (size_t)B, st, &dt))) return rc;
  if ((rc = stage_in(h, 1, space, recs, (size_t)N * NB_REC, st, &dr))) return rc;
  if ((rc = stage_in(h, 2, space, known, (size_t)B * N, st, &dk))) return rc;
  if ((rc = stage_out(h, 0, space, hull_xy, nh * NB_HMAX * 2, &dxy))) return rc;
  if ((rc = stage_out(h, 1, space, hull_cnt, nh, &dcnt))) return rc;
  if ((rc = stage_out(h, 2, space, hull_ptr, nh, &dptr))) return rc;
  if ((rc = stage_out(h, 3, space, nih0, nh * 2, &dn0))) return rc;
  if ((rc = stage_out(h, 4, space, samp, (size_t)B * N * P * (S + 1) * 2, &dsamp))) return rc;
  if ((rc = stage_out(h, 5, space, idx, nh * 2, &didx))) return rc;
  k_hulls<<<(unsigned)((size_t)B * N), 32 * NB_NPOL, 0, st>>>(h->cs, B, dt, dr, dk, delta, dxy, dcnt, dptr, dn0, didx, dsamp,
                                                           (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(hull_xy, dxy, nh * NB_HMAX * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(hull_cnt, dcnt, nh * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(hull_ptr, dptr, nh * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(nih0, dn0, nh * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (samp) NB_CUDA(cudaMemcpyAsync(samp, dsamp, (size_t)B * N * P * (S + 1) * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (idx) NB_CUDA(cudaMemcpyAsync(idx, didx, nh * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    int err = 0;
    NB_CUDA(cudaMemcpyAsync(&err, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
    if (err)
    {
      NB_CUDA(cudaMemsetAsync(h->err.p, 0, sizeof(int), st));
      g_err = "nb_hulls_batch: a window overlaps too many pieces or a hull has too many vertices";
      return NB_ERR_CAPACITY;
    }
  }
  return NB_OK;
}

extern "C" int nb_postcheck_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                                  const double* t_start, const double* recs, const uint8_t* late, double delta,
                                  int32_t* collide, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents;
  const int* dn;
  const double *dc, *dt, *dr;
  const uint8_t* dl;
  int* dcol;
  int rc;
  if ((rc = stage_in(h, 0, space, n_int, (size_t)B, st, &dn))) return rc;
  if ((rc = stage_in(h, 1, space, coeff, (size_t)B * 96, st, &dc))) return rc;
  if ((rc = stage_in(h, 2, space, t_start, (size_t)B, st, &dt))) return rc;
  if ((rc = stage_in(h, 3, space, recs, (size_t)N * NB_REC, st, &dr))) return rc;
  if ((rc = stage_in(h, 4, space, late, (size_t)B * N, st, &dl))) return rc;
  if ((rc = stage_out(h, 0, space, collide, (size_t)B, &dcol))) return rc;
  NB_CUDA(cudaMemsetAsync(dcol, 0, (size_t)B * sizeof(int), st));
  k_postcheck<<<(unsigned)(((size_t)B * N * NB_NPOL + 127) / 128), 128, 0, st>>>(h->cs, B, dn, dc, dt, dr, dl, delta, dcol, (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(collide, dcol, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
  }
  return NB_OK;
}

extern "C" int nb_commit_records_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                                       const double* t_start, double* recs_out, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int* dn;
  const double *dc, *dt;
  double* dr;
  int rc;
  if ((rc = stage_in(h, 0, space, n_int, (size_t)B, st, &dn))) return rc;
  if ((rc = stage_in(h, 1, space, coeff, (size_t)B * 96, st, &dc))) return rc;
  if ((rc = stage_in(h, 2, space, t_start, (size_t)B, st, &dt))) return rc;
  if ((rc = stage_out(h, 0, space, recs_out, (size_t)B * NB_REC, &dr))) return rc;
  NbPublishHdr hd;
  memset(&hd, 0, sizeof(hd));
  k_commit<<<B, 64, 0, st>>>(B, dn, dc, dt, h->cs.T, dr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                             nullptr, hd, (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(recs_out, dr, (size_t)B * NB_REC * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
  }
  return NB_OK;
}

extern "C" int nb_hull_index_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const int32_t* group,
                                   const uint8_t* known, const int32_t* hull_cnt_g, int64_t* hull_ptr, int32_t* hull_cnt,
                                   void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  if (space != NB_DEVICE)
  {
    g_err = "nb_hull_index_batch: device pointers only (it indexes device-resident hulls)";
    return NB_ERR_ARG;
  }
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * h->par.num_agents * NB_NPOL;
  k_hull_index<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(B, h->par.num_agents, agent_id, group, known, hull_cnt_g, hull_ptr,
                                                          hull_cnt);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}

extern "C" int nb_postcheck_hulls_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                                        const int32_t* group, const double* hull_xy_g, const int32_t* hull_cnt_g,
                                        const uint8_t* late, int32_t* collide, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  if (space != NB_DEVICE)
  {
    g_err = "nb_postcheck_hulls_batch: device pointers only";
    return NB_ERR_ARG;
  }
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * h->par.num_agents * NB_NPOL;
  NB_CUDA(cudaMemsetAsync(collide, 0, (size_t)B * sizeof(int), st));
  k_postcheck_hulls<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(h->cs, B, n_int, coeff, group, hull_xy_g, hull_cnt_g, late,
                                                               collide, nullptr);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}

// internal (nb_cycle.cu): the same with bounding boxes of the late hulls, built once per cycle by nb_internal_hull_aabb
int nb_internal_hull_aabb(nb_handle* h, size_t n_hulls, const double* hull_xy, const int* hull_cnt, double* aabb, cudaStream_t st)
{
  k_hull_aabb<<<(unsigned)((n_hulls + 127) / 128), 128, 0, st>>>(n_hulls, hull_xy, hull_cnt, aabb);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}
int nb_internal_postcheck_hulls(nb_handle* h, int B, const int32_t* n_int, const double* coeff, const int32_t* group,
                                const double* hull_xy_g, const int32_t* hull_cnt_g, const double* aabb, const uint8_t* late,
                                int32_t* collide, cudaStream_t st)
{
  const size_t n = (size_t)B * h->par.num_agents * NB_NPOL;
  NB_CUDA(cudaMemsetAsync(collide, 0, (size_t)B * sizeof(int), st));
  k_postcheck_hulls<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(h->cs, B, n_int, coeff, group, hull_xy_g, hull_cnt_g, late,
                                                               collide, aabb);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}

extern "C" int nb_compose_records_batch(nb_handle* h, int32_t B, int32_t space, const double* t, const uint8_t* has_prev,
                                        const double* prev, const double* now, double* out, int32_t* n_pieces,
                                        void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const double *dt, *dp, *dn;
  const uint8_t* dh;
  double* dout;
  int* dnp;
  int rc;
  if ((rc = stage_in(h, 0, space, t, (size_t)B, st, &dt))) return rc;
  if ((rc = stage_in(h, 1, space, has_prev, (size_t)B, st, &dh))) return rc;
  if ((rc = stage_in(h, 2, space, prev, (size_t)B * NB_REC, st, &dp))) return rc;
  if ((rc = stage_in(h, 3, space, now, (size_t)B * NB_REC, st, &dn))) return rc;
  if ((rc = stage_out(h, 0, space, out, (size_t)B * NB_REC, &dout))) return rc;
  if ((rc = stage_out(h, 1, space, n_pieces, (size_t)B, &dnp))) return rc;
  k_compose<<<(B + 63) / 64, 64, 0, st>>>(B, dt, dh, dp, dn, dout, dnp, (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(out, dout, (size_t)B * NB_REC * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(n_pieces, dnp, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    int err = 0;
    NB_CUDA(cudaMemcpyAsync(&err, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
    if (err)
    {
      NB_CUDA(cudaMemsetAsync(h->err.p, 0, sizeof(int), st));
      g_err = "nb_compose_records_batch: composed trajectory exceeds 16 pieces";
      return NB_ERR_CAPACITY;
    }
  }
  return NB_OK;
}

extern "C" int nb_commit_compose_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* n_int, const double* coeff,
                                       const double* t_start, const double* t_now, const double* prev,
                                       const int32_t* prev_agent, const uint8_t* has_prev, const int32_t* status,
                                       const int32_t* entangled, const int32_t* collide, double* recs_out,
                                       int32_t* n_pieces, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  if (space != NB_DEVICE)
  {
    g_err = "nb_commit_compose_batch: device pointers only (it runs inside the resident replan cycle)";
    return NB_ERR_ARG;
  }
  if (!t_now || !prev)
  {
    g_err = "nb_commit_compose_batch: t_now and prev are required";
    return NB_ERR_ARG;
  }
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  NbPublishHdr hd;
  memset(&hd, 0, sizeof(hd));
  k_commit<<<B, 64, 0, st>>>(B, n_int, coeff, t_start, h->cs.T, recs_out, t_now, prev, prev_agent, has_prev, status, entangled,
                             collide, nullptr, n_pieces, hd, (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}

extern "C" int nb_unpack_records_batch(nb_handle* h, int32_t space, const double* recs, int32_t* bp_cnt, double* bp_xy,
                                       double* latest_pos, void* stream)
{
  if (!h || !recs || !bp_cnt || !bp_xy) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents, bm = h->par.bp_max;
  const double* dr;
  int* dc;
  double *dx, *dl;
  int rc;
  if ((rc = stage_in(h, 0, space, recs, (size_t)N * NB_REC, st, &dr))) return rc;
  if ((rc = stage_out(h, 0, space, bp_cnt, (size_t)N, &dc))) return rc;
  if ((rc = stage_out(h, 1, space, bp_xy, (size_t)N * bm * 2, &dx))) return rc;
  if ((rc = stage_out(h, 2, space, latest_pos, (size_t)N * 2, &dl))) return rc;
  k_unpack_records<<<(N + 127) / 128, 128, 0, st>>>(N, bm, dr, dc, dx, dl, (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(bp_cnt, dc, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(bp_xy, dx, (size_t)N * bm * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (latest_pos) NB_CUDA(cudaMemcpyAsync(latest_pos, dl, (size_t)N * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    return ent_finish(h, space, st);
  }
  return NB_OK;
}

extern "C" int nb_publish_records_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const int32_t* n_int,
                                        const double* coeff, const double* t_start, const double* t_now, const double* prev,
                                        const uint8_t* has_prev, const int32_t* status, const int32_t* entangled,
                                        const int32_t* collide, const int32_t* fe_solved, nb_ent_state es, double bbox,
                                        double seq, double* recs_out, int32_t* n_pieces, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  if (space != NB_DEVICE)
  {
    g_err = "nb_publish_records_batch: device pointers only (it runs inside the resident replan cycle)";
    return NB_ERR_ARG;
  }
  if (!agent_id || !n_int || !coeff || !t_start || !recs_out || (t_now && !prev))
  {
    g_err = "nb_publish_records_batch: null argument";
    return NB_ERR_ARG;
  }
  if (h->par.num_static > 0 && es.cnt && !h->d_strep)
  {
    g_err = "nb_publish_records_batch: static obstacles without nb_set_static(strep)";
    return NB_ERR_ARG;
  }
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  NbPublishHdr hd;
  hd.on = 1, hd.N = h->par.num_agents, hd.M = h->par.num_static, hd.cap = h->par.ent_cap, hd.bbox = bbox, hd.pb = h->d_pb;
  hd.strep = h->d_strep, hd.strep_per_agent = h->strep_per_agent, hd.es = es, hd.seq = seq;
  k_commit<<<B, 64, 0, st>>>(B, n_int, coeff, t_start, h->cs.T, recs_out, t_now, prev, agent_id, has_prev, status, entangled,
                             collide, fe_solved, n_pieces, hd, (int*)h->err.p);
  h->launches += 1;
  NB_CUDA(cudaGetLastError());
  return NB_OK;
}

// phase 0: the whole post-check; 1 / 2 (device pointers, nb_cycle.cu): the part before / after the optimised trajectory is
// needed -- phase 1 runs beside the QP
int nb_internal_postcheck_entangle(nb_handle* h, int phase, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                                   const uint8_t* late, const int32_t* bp_cnt, const double* bp_xy, const int32_t* bp_cnt_late,
                                   const double* bp_xy_late, nb_ent_state st_in, const double* prev_pos,
                                   const double* prev_pos_agent, const double* cur, const int32_t* n_int, const double* coeff,
                                   const double* t_start, const double* samp, int32_t samp_shared, const int32_t* samp_group,
                                   const double* late_recs, int32_t* entangled, void* stream);

extern "C" int nb_postcheck_entangle_batch(nb_handle* h, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                                           const uint8_t* late, const int32_t* bp_cnt, const double* bp_xy,
                                           const int32_t* bp_cnt_late, const double* bp_xy_late, nb_ent_state st_in,
                                           const double* prev_pos, const double* prev_pos_agent, const double* cur,
                                           const int32_t* n_int, const double* coeff, const double* t_start, const double* samp,
                                           int32_t samp_shared, const int32_t* samp_group, const double* late_recs,
                                           int32_t* entangled, void* stream)
{
  return nb_internal_postcheck_entangle(h, 0, B, space, agent_id, known, late, bp_cnt, bp_xy, bp_cnt_late, bp_xy_late, st_in, prev_pos,
                                        prev_pos_agent, cur, n_int, coeff, t_start, samp, samp_shared, samp_group, late_recs,
                                        entangled, stream);
}

int nb_internal_postcheck_entangle(nb_handle* h, int phase, int32_t B, int32_t space, const int32_t* agent_id, const uint8_t* known,
                                   const uint8_t* late, const int32_t* bp_cnt, const double* bp_xy, const int32_t* bp_cnt_late,
                                   const double* bp_xy_late, nb_ent_state st_in, const double* prev_pos,
                                   const double* prev_pos_agent, const double* cur, const int32_t* n_int, const double* coeff,
                                   const double* t_start, const double* samp, int32_t samp_shared, const int32_t* samp_group,
                                   const double* late_recs, int32_t* entangled, void* stream)
{
  if (!h || B < 0) return NB_ERR_ARG;
  if (B == 0) return NB_OK;
  if (!agent_id || !known || !late || !bp_cnt || !bp_xy || !bp_cnt_late || !bp_xy_late || !st_in.cnt || !prev_pos ||
      !prev_pos_agent || !cur || !n_int || !coeff || !t_start || !samp || !late_recs || !entangled)
  {
    g_err = "nb_postcheck_entangle_batch: null argument";
    return NB_ERR_ARG;
  }
  if (space == NB_HOST && samp_group)
  {
    g_err = "nb_postcheck_entangle_batch: samp_group is only supported with device pointers";
    return NB_ERR_ARG;
  }
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->par.num_agents, S = h->par.samples, cap = h->par.ent_cap, NA = N + h->par.num_static;
  NbEntArgs a;
  int rc;
  if ((rc = ent_common(h, &a, 4, B, space, agent_id, known, bp_cnt, bp_xy, st))) return rc;
  nb_ent_state d;
  if ((rc = stage_state_in(h, 4, space, st_in, (size_t)B, st, &d, true))) return rc;
  a.st = d;
  if ((rc = stage_in(h, 9, space, n_int, (size_t)B, st, &a.n_int))) return rc;
  if ((rc = stage_in(h, 10, space, coeff, (size_t)B * 96, st, &a.coeff))) return rc;
  if ((rc = stage_in(h, 11, space, samp, (size_t)(samp_shared ? 1 : B) * N * h->par.num_pol * (S + 1) * 2, st, &a.samp))) return rc;
  a.samp_shared = samp_shared;
  a.samp_group = samp_group;
  if ((rc = stage_in(h, 12, space, prev_pos, (size_t)B * (N + 1) * 2, st, &a.prev_pos))) return rc;
  if ((rc = stage_in(h, 13, space, prev_pos_agent, (size_t)B * N * 2, st, &a.prev_pos_agent))) return rc;
  if ((rc = stage_in(h, 14, space, cur, (size_t)B * 2, st, &a.cur))) return rc;
  if ((rc = stage_in(h, 15, space, late, (size_t)B * N, st, &a.late))) return rc;
  // further inputs and the scratch live in one library buffer: late records, late bend points, t_start | work states,
  // per-agent samples of interval 0, per-agent known flags
  const size_t o_rec = 0, o_bpc = o_rec + (size_t)N * NB_REC * 8, o_bpx = o_bpc + (((size_t)N * 4 + 7) & ~(size_t)7);
  const size_t o_ts = o_bpx + (size_t)N * h->par.bp_max * 16, o_cnt = o_ts + (size_t)B * 8;
  const size_t o_alpha = o_cnt + (size_t)B * 8, o_beta = o_alpha + (size_t)B * cap * 8, o_bend = o_beta + (size_t)B * cap * 8;
  const size_t o_act = o_bend + (((size_t)B * cap * 4 + 7) & ~(size_t)7), o_ps = o_act + (((size_t)B * NA * 4 + 7) & ~(size_t)7);
  const size_t o_pk = o_ps + (size_t)B * N * (S + 1) * 16, total = o_pk + (size_t)B * N + 16;
  if (h->ent_scratch.ensure(total))
  {
    g_err = "cudaMalloc failed for the post-check scratch";
    return NB_ERR_CUDA;
  }
  char* base = (char*)h->ent_scratch.p;
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(base + o_rec, late_recs, (size_t)N * NB_REC * 8, cudaMemcpyHostToDevice, st));
    NB_CUDA(cudaMemcpyAsync(base + o_bpc, bp_cnt_late, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    NB_CUDA(cudaMemcpyAsync(base + o_bpx, bp_xy_late, (size_t)N * h->par.bp_max * 16, cudaMemcpyHostToDevice, st));
    NB_CUDA(cudaMemcpyAsync(base + o_ts, t_start, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    a.late_recs = (const double*)(base + o_rec), a.bp_cnt_late = (const int*)(base + o_bpc);
    a.bp_xy_late = (const double*)(base + o_bpx), a.t_start = (const double*)(base + o_ts);
  }
  else
    a.late_recs = late_recs, a.bp_cnt_late = bp_cnt_late, a.bp_xy_late = bp_xy_late, a.t_start = t_start;
  a.out.cnt = (int32_t*)(base + o_cnt), a.out.alpha = (int32_t*)(base + o_alpha), a.out.beta = (double*)(base + o_beta);
  a.out.bend = (int32_t*)(base + o_bend), a.out.active = (int32_t*)(base + o_act);
  a.psamp = (double*)(base + o_ps), a.pknown = (unsigned char*)(base + o_pk);
  a.phase = phase;
  if ((rc = stage_out(h, 7, space, entangled, (size_t)B, &a.result))) return rc;
  if ((rc = ent_launch(h, a, B, st))) return rc;
  if (space == NB_HOST) NB_CUDA(cudaMemcpyAsync(entangled, a.result, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  return ent_finish(h, space, st);
}

// ------------------------------------------------------------------------------------------ K0 ABI (front end)
extern "C" int nb_search_configure(nb_handle* h, const nb_search_params* sp)
{
  if (!h || !sp) return NB_ERR_ARG;
  if (sp->num_samples < 2 || sp->num_samples * sp->num_samples > NB_SEARCH_MAXCHILD || sp->max_nodes < 64 ||
      sp->max_expansions < 0 || sp->ecap < 1 || sp->ecap > 0x7fff || !(sp->voxel_size > 0) || !(sp->j_max > 0))
  {
    g_err = "nb_search_configure: unsupported parameters (2 <= num_samples <= 5, max_nodes >= 64, 1 <= ecap < 32768)";
    return NB_ERR_ARG;
  }
  h->sp = *sp;
  h->sp_set = 1;
  return NB_OK;
}

extern "C" int nb_set_static_longest(nb_handle* h, const double* longest)
{
  if (!h) return NB_ERR_ARG;
  const int M = h->par.num_static;
  if (M == 0) return NB_OK;
  if (!longest) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  if (!h->d_st_longest) NB_CUDA(cudaMalloc(&h->d_st_longest, sizeof(double) * 2 * M));
  NB_CUDA(cudaMemcpy(h->d_st_longest, longest, sizeof(double) * 2 * M, cudaMemcpyHostToDevice));
  return NB_OK;
}

namespace
{
template <typename T>
int sstage_in(nb_handle* h, int slot, int space, const T* src, size_t count, cudaStream_t st, const T** dst)
{
  if (space == NB_DEVICE || src == nullptr || count == 0)
  {
    *dst = src;
    return NB_OK;
  }
  if (h->sin[slot].ensure(count * sizeof(T)))
  {
    g_err = "cudaMalloc failed while staging search inputs";
    return NB_ERR_CUDA;
  }
  NB_CUDA(cudaMemcpyAsync(h->sin[slot].p, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
  *dst = (const T*)h->sin[slot].p;
  return NB_OK;
}
template <typename T>
int sstage_out(nb_handle* h, int slot, int space, T* user, size_t count, T** dst)
{
  if (space == NB_DEVICE)
  {
    *dst = user;
    return NB_OK;
  }
  if (h->sout[slot].ensure(count * sizeof(T)))
  {
    g_err = "cudaMalloc failed while staging search outputs";
    return NB_ERR_CUDA;
  }
  *dst = (T*)h->sout[slot].p;
  return NB_OK;
}
}  // namespace

extern "C" int nb_search_batch(nb_handle* h, const nb_search_args* u, void* stream)
{
  if (!h || !u || u->B < 0) return NB_ERR_ARG;
  if (u->B == 0) return NB_OK;
  if (!h->sp_set)
  {
    g_err = "nb_search_batch before nb_search_configure";
    return NB_ERR_ARG;
  }
  NB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const nb_search_params& sp = h->sp;
  const int B = u->B, N = h->par.num_agents, M = h->par.num_static, NA = N + M, S = h->par.samples, np = h->par.num_pol;
  const int cap = h->par.ent_cap, space = u->space, G = u->group ? u->n_groups : B;
  if (S > 8)
  {
    g_err = "nb_search_batch: num_sample_per_interval > 8 is not supported";
    return NB_ERR_ARG;
  }
  if (!u->agent_id || !u->init || !u->goal || !u->coeffs_z || !u->hull_xy || !u->hull_cnt || !u->samp || !u->known ||
      !u->es.cnt || !u->es.alpha || !u->es.beta || !u->es.bend || !u->es.active || !u->bp_cnt || !u->bp_xy || !u->comb ||
      !u->status || !u->solved || !u->n_int || !u->coeff || !u->esv.cnt || !u->esv.alpha || !u->esv.beta || !u->esv.bend ||
      !u->esv.active || !u->stats || !u->cost || (u->group && u->n_groups < 1))
  {
    g_err = "nb_search_batch: null argument";
    return NB_ERR_ARG;
  }
  if (space == NB_HOST)
  {  // host arrays can be checked before anything is launched
    for (int b = 0; b < B; b++)
    {
      if (u->agent_id[b] < 1 || u->agent_id[b] > N || (u->group && (u->group[b] < 0 || u->group[b] >= u->n_groups)))
      {
        g_err = "nb_search_batch: agent_id or group out of range";
        return NB_ERR_ARG;
      }
      const uint8_t* cb = u->comb + (u->comb_shared ? 0 : (size_t)b * sp.num_samples * sp.num_samples);
      unsigned seen = 0;
      for (int k = 0; k < sp.num_samples * sp.num_samples; k++)
        if (cb[k] < 32) seen |= 1u << cb[k];
      if (seen != (sp.num_samples * sp.num_samples >= 32 ? 0xffffffffu : (1u << (sp.num_samples * sp.num_samples)) - 1))
      {
        g_err = "nb_search_batch: comb is not a permutation of the jerk samples";
        return NB_ERR_ARG;
      }
    }
  }
  if (M > 0 && (!h->d_strep || !h->d_st_longest || !h->d_st_xy))
  {
    g_err = "nb_search_batch with static obstacles needs nb_set_static(strep) and nb_set_static_longest first";
    return NB_ERR_ARG;
  }
  NbSearchArgs a;
  memset(&a, 0, sizeof(a));
  NbSearchPar& p = a.p;
  nb_search_fill_par(h->par, sp, h->cs, &p);
  int rc;
  if ((rc = sstage_in(h, 0, space, u->agent_id, (size_t)B, st, &a.agent_id))) return rc;
  if ((rc = sstage_in(h, 1, space, u->init, (size_t)B * 6, st, &a.init))) return rc;
  if ((rc = sstage_in(h, 2, space, u->goal, (size_t)B * 2, st, &a.goal))) return rc;
  if ((rc = sstage_in(h, 3, space, u->coeffs_z, (size_t)B * NB_NPOL * 4, st, &a.coeffs_z))) return rc;
  if ((rc = sstage_in(h, 4, space, u->group, (size_t)B, st, &a.group))) return rc;
  if ((rc = sstage_in(h, 5, space, u->hull_xy, (size_t)G * N * NB_NPOL * NB_HULL_STRIDE * 2, st, &a.hull_xy))) return rc;
  if ((rc = sstage_in(h, 6, space, u->hull_cnt, (size_t)G * N * NB_NPOL, st, &a.hull_cnt))) return rc;
  if ((rc = sstage_in(h, 7, space, u->samp, (size_t)G * N * np * (S + 1) * 2, st, &a.samp))) return rc;
  if ((rc = sstage_in(h, 8, space, u->known, (size_t)B * N, st, &a.known))) return rc;
  if ((rc = sstage_in(h, 9, space, (const int32_t*)u->es.cnt, (size_t)B * 2, st, (const int32_t**)&a.es.cnt))) return rc;
  if ((rc = sstage_in(h, 10, space, (const int32_t*)u->es.alpha, (size_t)B * cap * 2, st, (const int32_t**)&a.es.alpha))) return rc;
  if ((rc = sstage_in(h, 11, space, (const double*)u->es.beta, (size_t)B * cap, st, (const double**)&a.es.beta))) return rc;
  if ((rc = sstage_in(h, 12, space, (const int32_t*)u->es.bend, (size_t)B * cap, st, (const int32_t**)&a.es.bend))) return rc;
  if ((rc = sstage_in(h, 13, space, (const int32_t*)u->es.active, (size_t)B * NA, st, (const int32_t**)&a.es.active))) return rc;
  if ((rc = sstage_in(h, 14, space, u->bp_cnt, (size_t)N, st, &a.bp_cnt))) return rc;
  if ((rc = sstage_in(h, 15, space, u->bp_xy, (size_t)N * h->par.bp_max * 2, st, &a.bp_xy))) return rc;
  if ((rc = sstage_in(h, 16, space, u->comb, (size_t)(u->comb_shared ? 1 : B) * p.nchild, st, &a.comb))) return rc;
  a.comb_shared = u->comb_shared;
  a.st_ptr = h->d_st_ptr, a.st_xy = h->d_st_xy, a.strep = h->d_strep, a.st_longest = h->d_st_longest, a.pb = h->d_pb;
  a.strep_per_agent = h->strep_per_agent;
  const size_t ns9 = (size_t)B * 9;
  if ((rc = sstage_out(h, 0, space, u->status, (size_t)B, &a.status))) return rc;
  if ((rc = sstage_out(h, 1, space, u->solved, (size_t)B, &a.solved))) return rc;
  if ((rc = sstage_out(h, 2, space, u->n_int, (size_t)B, &a.n_int))) return rc;
  if ((rc = sstage_out(h, 3, space, u->coeff, (size_t)B * 3 * NB_NPOL * 4, &a.coeff))) return rc;
  if ((rc = sstage_out(h, 4, space, u->esv.cnt, ns9 * 2, &a.esv.cnt))) return rc;
  if ((rc = sstage_out(h, 5, space, u->esv.alpha, ns9 * cap * 2, &a.esv.alpha))) return rc;
  if ((rc = sstage_out(h, 6, space, u->esv.beta, ns9 * cap, &a.esv.beta))) return rc;
  if ((rc = sstage_out(h, 7, space, u->esv.bend, ns9 * cap, &a.esv.bend))) return rc;
  if ((rc = sstage_out(h, 8, space, u->esv.active, ns9 * NA, &a.esv.active))) return rc;
  if ((rc = sstage_out(h, 9, space, u->stats, (size_t)B * 4, &a.stats))) return rc;
  if ((rc = sstage_out(h, 10, space, u->cost, (size_t)B, &a.cost))) return rc;
  // workspace
  const size_t mn = (size_t)sp.max_nodes;
  const size_t ch_stride = (size_t)nb_search_ch_stride(p);
  bool bad = false;
  bad |= h->sw_meta.ensure(B * mn * sizeof(NbInt4)) != 0;
  bad |= h->sw_kin.ensure(B * mn * NB_SEARCH_KIN * sizeof(double)) != 0;
  bad |= h->sw_alpha.ensure(B * mn * p.ecap * 2 * sizeof(int)) != 0;
  bad |= h->sw_beta.ensure(B * mn * p.ecap * sizeof(double)) != 0;
  bad |= h->sw_bend.ensure(B * mn * p.ecap * sizeof(int)) != 0;
  bad |= h->sw_hash.ensure((size_t)B * p.hcap * sizeof(NbInt4)) != 0;
  // global homes of everything the kernel tries to keep in shared memory (used when it does not fit)
  bad |= h->sw_heap.ensure(B * mn * sizeof(int)) != 0;
  bad |= h->sw_gh.ensure(B * mn * 2 * sizeof(double)) != 0;
  bad |= h->sw_ng.ensure(B * mn * sizeof(double)) != 0;
  bad |= h->sw_chi.ensure((size_t)B * (p.nchild * ch_stride + NA) * sizeof(int)) != 0;
  bad |= h->sw_chd.ensure((size_t)B * nb_search_chd_stride(p) * sizeof(double)) != 0;
  bad |= h->sw_fcode.ensure((size_t)B * nb_search_fcode_bytes(p)) != 0;
  if (bad)
  {
    g_err = "cudaMalloc failed for the search workspace";
    return NB_ERR_CUDA;
  }
  a.nd_meta = (NbInt4*)h->sw_meta.p, a.nd_kin = (double*)h->sw_kin.p, a.nd_alpha = (int*)h->sw_alpha.p;
  a.nd_beta = (double*)h->sw_beta.p, a.nd_bend = (int*)h->sw_bend.p, a.hash = (NbInt4*)h->sw_hash.p;
  a.heap_g = (int*)h->sw_heap.p, a.gh_g = (double*)h->sw_gh.p, a.nd_g = (double*)h->sw_ng.p, a.ch_int = (int*)h->sw_chi.p, a.ch_dbl = (double*)h->sw_chd.p;
  a.fcode_g = (uint8_t*)h->sw_fcode.p;
  a.err = (int*)h->err.p;
  a.prof = nullptr;
  if (h->profiling)
  {
    if (h->sprof.ensure((size_t)B * 16 * sizeof(long long)))
    {
      g_err = "cudaMalloc failed";
      return NB_ERR_CUDA;
    }
    NB_CUDA(cudaMemsetAsync(h->sprof.p, 0, (size_t)B * 16 * sizeof(long long), st));
    a.prof = (long long*)h->sprof.p;
    h->sprof_B = B;
  }
  const char* etxt = nullptr;
  if (nb_search_launch(&a, B, stream, &h->search_smem_set, &etxt))
  {
    g_err = std::string("k_search launch: ") + (etxt ? etxt : "?");
    return NB_ERR_CUDA;
  }
  h->launches += 1;
  if (space == NB_HOST)
  {
    NB_CUDA(cudaMemcpyAsync(u->status, a.status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->solved, a.solved, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->n_int, a.n_int, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->coeff, a.coeff, (size_t)B * 96 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->esv.cnt, a.esv.cnt, ns9 * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->esv.alpha, a.esv.alpha, ns9 * cap * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->esv.beta, a.esv.beta, ns9 * cap * sizeof(double), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->esv.bend, a.esv.bend, ns9 * cap * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->esv.active, a.esv.active, ns9 * NA * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->stats, a.stats, (size_t)B * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaMemcpyAsync(u->cost, a.cost, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
    int err = 0;
    NB_CUDA(cudaMemcpyAsync(&err, h->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NB_CUDA(cudaStreamSynchronize(st));
    if (err)
    {
      NB_CUDA(cudaMemsetAsync(h->err.p, 0, sizeof(int), st));
      g_err = "front-end search: a node's alphas list exceeded search ecap / ent_cap";
      return NB_ERR_CAPACITY;
    }
  }
  return NB_OK;
}

// Measurement hook: SM cycles thread 0 of every search CTA spent per phase in the last profiled nb_search_batch
// (nb_set_profiling on): [0] children, [1] sequential resolve, [2] pool copy, [3] open-list pop, [4] collision
// tests, [5] endpoint tests, [6] set-up.  out: [B][8].
extern "C" int nb_search_phase_cycles(nb_handle* h, long long* out, int B)
{
  if (!h || !out || !h->sprof.p || B > h->sprof_B) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  NB_CUDA(cudaDeviceSynchronize());
  NB_CUDA(cudaMemcpy(out, h->sprof.p, (size_t)B * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return NB_OK;
}

// Measurement hook: SM cycles lane 0 of every QP warp spent per phase in the last profiled nb_replan_batch, out [B][16]:
// [0] set-up, [1] residual sweep, [2] dual residual + stopping test, [3] normal-matrix assembly, [4] factorisation,
// [5] predictor solve, [6] predictor sweep, [7] corrector solve, [8] final sweep.
extern "C" int nb_qp_phase_cycles(nb_handle* h, long long* out, int B)
{
  if (!h || !out || !h->qprof.p || B > h->qprof_B) return NB_ERR_ARG;
  NB_CUDA(cudaSetDevice(h->device));
  NB_CUDA(cudaDeviceSynchronize());
  NB_CUDA(cudaMemcpy(out, h->qprof.p, (size_t)B * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return NB_OK;
}
