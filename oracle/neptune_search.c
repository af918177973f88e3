/*
 * neptune_search.c -- CPU ORACLE (test infrastructure, NOT the product) for the front end of the
 * replan: KinodynamicSearch (reference neptune/src/kinodynamic_search.cpp), SURVEY.md section 8(f) #1.
 *
 * Plain-C restatement of
 *   KinodynamicSearch::setUp (4-argument entangle form)   kinodynamic_search.cpp:190-249
 *   KinodynamicSearch::expandAndAddToQueue (root, node)   :1240-1385, :1045-1228
 *   KinodynamicSearch::entanglesWithOtherAgents           :805-895
 *   eu::getTetherLength / getBendPt2dwIdx                 entangle_utils.cpp:1724-1743, :1681-1707
 *   KinodynamicSearch::collidesWithObstacles2dSolve       :1514-1580
 *   KinodynamicSearch::collidesWithBases2d                :1583-1627
 *   KinodynamicSearch::run                                :1629-1827
 *   recoverPwpOut / recoverEntStateVector                 :521-553, :582-603
 *   getIz / power_int                                     :2006-2031
 *   CompareCost                                           kinodynamic_search.hpp:163-179
 *
 * Third-party behaviour restated because it decides results:
 *   std::priority_queue = std::push_heap / std::pop_heap of libstdc++ (bits/stl_heap.h: __push_heap,
 *   __adjust_heap); CompareCost is not a strict weak order (ties within 1e-5 go by h), so the pop order
 *   depends on that exact sift algorithm.  std::unordered_map is used only as key -> node (find / insert
 *   that keeps the first), so any map gives the same answers.
 *
 * Two inputs replace non-deterministic state of the reference, both injected by the caller:
 *   comb[]          the order of the 25 jerk samples (the reference shuffles all_combinations_ with a
 *                   wall-clock seed, :321-322, :1462-1463)
 *   max_expansions  pops of the open list allowed before "Max Runtime was reached" (:1646-1652 uses a
 *                   wall-clock timer); max_nodes is node_num_max_ (:370-371).
 *
 * PARITY STATUS: PINNED against the reference's own kinodynamic_search.cpp, compiled unmodified into oracle/_ref
 * (Eigen replaced by oracle/eigen_shim): tests/test_reference_pin.py::test_search_matches_reference compares whole
 * runs field by field, bit for bit; recorded outputs in tests/golden/reference/ref_search.npz.  Also checked by
 * properties and against the real std::priority_queue in tests/test_search.py.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "neptune_oracle.h"

typedef struct
{
  int prev, index, state;
  double end[6], cx[4], cy[4], Q[8], g, h;
  orc_ent es; /* lists live in the pool arrays */
} knode;

typedef struct
{
  const orc_search_par* par;
  const orc_search_in* in;
  orc_ectx cx;
  double Ainv[16], V[9];
  knode* pool;
  int *alpha, *bend, *active;
  double* beta;
  int n_used;
  int* heap;
  int heap_n;
  int* hkey; /* [hcap][4]: ix, iy, iz, id+1 */
  int hcap;
  int ran_trigger;
  int goal_occupied;
  int *toadd, *act_old;
  int tcap;
  int overflow;
} ksearch;

static double norm2(double x, double y) { return sqrt(x * x + y * y); }

/* CompareCost (kinodynamic_search.hpp:163-179): true when `l` has LOWER priority than `r` */
static int cmp_cost(const ksearch* s, int l, int r)
{
  const double bias = s->par->bias;
  double cl = s->pool[l].g + bias * s->pool[l].h;
  double cr = s->pool[r].g + bias * s->pool[r].h;
  if (fabs(cl - cr) < 1e-5) return s->pool[l].h > s->pool[r].h;
  return cl > cr;
}

/* libstdc++ std::__push_heap */
static void heap_push_at(ksearch* s, int hole, int top, int value)
{
  int parent = (hole - 1) / 2;
  while (hole > top && cmp_cost(s, s->heap[parent], value))
  {
    s->heap[hole] = s->heap[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  s->heap[hole] = value;
}

static void heap_push(ksearch* s, int id)
{
  s->heap[s->heap_n++] = id;
  heap_push_at(s, s->heap_n - 1, 0, id);
}

/* top() then pop(): std::pop_heap = __pop_heap + __adjust_heap, then pop_back */
static int heap_pop(ksearch* s)
{
  int top = s->heap[0];
  if (s->heap_n > 1)
  {
    int last = s->heap_n - 1;
    int value = s->heap[last];
    s->heap[last] = s->heap[0];
    int len = last, hole = 0, child = 0;
    while (child < (len - 1) / 2)
    {
      child = 2 * (child + 1);
      if (cmp_cost(s, s->heap[child], s->heap[child - 1])) child--;
      s->heap[hole] = s->heap[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
      child = 2 * (child + 1);
      s->heap[hole] = s->heap[child - 1];
      hole = child - 1;
    }
    heap_push_at(s, hole, 0, value);
  }
  s->heap_n--;
  return top;
}

static uint32_t hash3(int ix, int iy, int iz)
{
  uint32_t h = (uint32_t)ix * 0x9E3779B1u;
  h ^= (uint32_t)iy * 0x85EBCA77u + (h << 6) + (h >> 2);
  h ^= (uint32_t)iz * 0xC2B2AE3Du + (h << 6) + (h >> 2);
  return h;
}

static int hash_find(const ksearch* s, int ix, int iy, int iz)
{
  uint32_t p = hash3(ix, iy, iz) & (uint32_t)(s->hcap - 1);
  for (;;)
  {
    const int* e = s->hkey + 4 * p;
    if (e[3] == 0) return -1;
    if (e[0] == ix && e[1] == iy && e[2] == iz) return e[3] - 1;
    p = (p + 1) & (uint32_t)(s->hcap - 1);
  }
}

/* unordered_map::insert: keeps the existing mapping when the key is present */
static void hash_insert(ksearch* s, int ix, int iy, int iz, int id)
{
  uint32_t p = hash3(ix, iy, iz) & (uint32_t)(s->hcap - 1);
  for (;;)
  {
    int* e = s->hkey + 4 * p;
    if (e[3] == 0)
    {
      e[0] = ix, e[1] = iy, e[2] = iz, e[3] = id + 1;
      return;
    }
    if (e[0] == ix && e[1] == iy && e[2] == iz) return;
    p = (p + 1) & (uint32_t)(s->hcap - 1);
  }
}

/* KinodynamicSearch::power_int (:2016-2031), unsigned 32-bit arithmetic */
static uint32_t power_int(uint32_t base, uint32_t exponent)
{
  if (exponent == 0) return 1;
  if (base < 2) return base;
  uint32_t result = 1;
  for (uint32_t term = base;; term = term * term)
  {
    if (exponent % 2 != 0) result *= term;
    exponent /= 2;
    if (exponent == 0) break;
  }
  return result;
}

/* KinodynamicSearch::getIz (:2006-2014) */
static int get_iz(const orc_ent* es)
{
  uint32_t iz = 0;
  for (int i = 0; i < es->n_alpha; i++) iz += (uint32_t)(i + 1) * power_int((uint32_t)es->alpha[2 * i], (uint32_t)es->alpha[2 * i + 1]);
  return (int)iz;
}

/* unsigned int ix = round(x / voxel) then Eigen::Vector2i(ix, iy): two's-complement wrap (x86-64 GCC) */
static int voxel_index(double x, double voxel) { return (int)(uint32_t)(int64_t)round(x / voxel); }

static void node_lists(ksearch* s, int id)
{
  const int cap = s->par->ecap, NA = s->cx.N + s->cx.M;
  s->pool[id].es.alpha = s->alpha + (size_t)id * 2 * cap;
  s->pool[id].es.beta = s->beta + (size_t)id * cap;
  s->pool[id].es.bend = s->bend + (size_t)id * cap;
  s->pool[id].es.active = s->active + (size_t)id * NA;
}

static void copy_state(ksearch* s, orc_ent* dst, const orc_ent* src)
{
  const int cap = s->par->ecap, NA = s->cx.N + s->cx.M;
  dst->n_alpha = src->n_alpha;
  dst->n_bend = src->n_bend;
  memcpy(dst->alpha, src->alpha, sizeof(int) * 2 * (size_t)(src->n_alpha < cap ? src->n_alpha : cap));
  memcpy(dst->beta, src->beta, sizeof(double) * (size_t)(src->n_alpha < cap ? src->n_alpha : cap));
  memcpy(dst->bend, src->bend, sizeof(int) * (size_t)(src->n_bend < cap ? src->n_bend : cap));
  memcpy(dst->active, src->active, sizeof(int) * NA);
}

/* eu::getTetherLength (entangle_utils.cpp:1724-1743) with getBendPt2dwIdx (:1681-1707) */
static double tether_length(const ksearch* s, const orc_ent* es, const double pk1[2])
{
  const int N = s->cx.N;
  double length = 0.0;
  double prev[2] = { s->cx.pb[2 * s->cx.self], s->cx.pb[2 * s->cx.self + 1] };
  for (int i = 0; i < es->n_bend; i++)
  {
    const int q = es->bend[i];
    const int id = es->alpha[2 * q], cs = es->alpha[2 * q + 1];
    double bp[2], comp = 0.0;
    if (id <= N)
    {
      bp[0] = s->cx.pb[2 * (id - 1)], bp[1] = s->cx.pb[2 * (id - 1) + 1];
    }
    else
    {
      bp[0] = s->cx.strep[4 * (id - N - 1) + 2 * cs], bp[1] = s->cx.strep[4 * (id - N - 1) + 2 * cs + 1];
      comp = s->in->st_longest[2 * (id - N - 1) + cs];
    }
    length += norm2(bp[0] - prev[0], bp[1] - prev[1]) + 2 * comp;
    prev[0] = bp[0], prev[1] = bp[1];
  }
  length += norm2(pk1[0] - prev[0], pk1[1] - prev[1]);
  return length;
}

/* KinodynamicSearch::entanglesWithOtherAgents (:805-895).  Returns 1 entangles, 0 fine. */
static int entangles(ksearch* s, knode* nd, double* arc_length)
{
  const orc_search_par* par = s->par;
  const orc_search_in* in = s->in;
  const orc_ectx* cx = &s->cx;
  const int N = cx->N, M = cx->M, NA = N + M, S = par->S, num_pol = par->num_pol;
  orc_ent* es = &nd->es;
  double pk[2] = { nd->cx[3], nd->cy[3] }, pk1[2] = { pk[0], pk[1] };
  memcpy(s->act_old, es->active, sizeof(int) * NA);
  for (int j = 1; j <= S; j++)
  {
    int nadd = 0;
    if (j < S)
    {
      const double t = par->T * j / S;
      const double t3 = t * t * t, t2 = t * t;
      pk1[0] = nd->cx[0] * t3 + nd->cx[1] * t2 + nd->cx[2] * t + nd->cx[3];
      pk1[1] = nd->cy[0] * t3 + nd->cy[1] * t2 + nd->cy[2] * t + nd->cy[3];
    }
    else
      pk1[0] = nd->end[0], pk1[1] = nd->end[1];
    *arc_length += norm2(pk1[0] - pk[0], pk1[1] - pk[1]);
    for (int a = 0; a < N; a++)
    {
      if (a == cx->self || !in->known[a]) continue;
      const double *pik, *pik1;
      if (nd->index > num_pol)
      {
        pik = in->samp + ((size_t)(a * num_pol + (num_pol - 1)) * (S + 1) + S) * 2;
        pik1 = pik;
      }
      else
      {
        pik = in->samp + ((size_t)(a * num_pol + (nd->index - 1)) * (S + 1) + (j - 1)) * 2;
        pik1 = in->samp + ((size_t)(a * num_pol + (nd->index - 1)) * (S + 1) + j) * 2;
      }
      if (nadd + cx->bp_cnt[a] + 2 > s->tcap) return 1; /* more crossings than the list bound allows (:844) */
      nadd = orc_hsig_agent(s->toadd, nadd, pk, pk1, pik, pik1, cx->pb + 2 * cx->self, cx->bp_xy + 2 * cx->bp_max * a,
                            cx->bp_cnt[a], a + 1);
    }
    if (nadd + M > s->tcap) return 1;
    nadd = orc_hsig_static(s->toadd, nadd, pk, pk1, cx->strep, M, N);
    if (es->n_alpha + nadd > NA) return 1; /* :844-848 */
    if (orc_add_alpha_beta(s->toadd, nadd, es, pk, cx))
    {
      s->overflow = 1; /* storage capacity ecap exceeded: not a reference outcome */
      return 1;
    }
    for (int a = 0; a < N; a++)
    {
      if (s->act_old[a] < 2 && es->active[a] >= 2) return 1;
      if (s->act_old[a] >= 2 && es->active[a] > s->act_old[a]) return 1;
    }
    orc_update_bend_pts(es, pk1, cx);
    memcpy(s->act_old, es->active, sizeof(int) * NA);
    pk[0] = pk1[0], pk[1] = pk1[1];
  }
  if (tether_length(s, es, pk1) > par->tether) return 1; /* :884-891 */
  return 0;
}

/* KinodynamicSearch::collidesWithObstacles2dSolve (:1514-1580) */
static int collides_solve(const ksearch* s, const double* Q /*[2][4]*/, int index)
{
  const orc_search_par* par = s->par;
  const orc_search_in* in = s->in;
  double cps[8];
  for (int i = 0; i < 4; i++) cps[2 * i] = Q[i], cps[2 * i + 1] = Q[4 + i];
  if (index > par->num_pol) index = par->num_pol;
  for (int o = 0; o < par->N; o++)
  {
    const int hn = in->hull_cnt[o * ORC_NPOL_MAX + index - 1];
    if (hn <= 0) continue;
    if (orc_gjk_collision(in->hull_xy + ((size_t)(o * ORC_NPOL_MAX + index - 1) * ORC_SEARCH_HSTRIDE) * 2, hn, cps, 4)) return 1;
  }
  for (int m = 0; m < par->M; m++)
  {
    const long long p0 = in->st_ptr[m], p1 = in->st_ptr[m + 1];
    if (orc_gjk_collision(in->st_xy + 2 * p0, (int)(p1 - p0), cps, 4)) return 1;
  }
  return 0;
}

/* KinodynamicSearch::collidesWithBases2d (:1583-1627) */
static int collides_bases(const ksearch* s, const double* Q)
{
  const orc_search_par* par = s->par;
  if (!par->enable_entangle) return 0;
  const double radius = 0.7, safe_dist = par->T * par->v_max * 2;
  double cps[8];
  for (int i = 0; i < 4; i++) cps[2 * i] = Q[i], cps[2 * i + 1] = Q[4 + i];
  for (int a = 0; a < par->N; a++)
  {
    if (a == s->cx.self) continue;
    const double bx = s->cx.pb[2 * a], by = s->cx.pb[2 * a + 1];
    const double d1 = norm2(cps[0] - bx, cps[1] - by);
    if (d1 > safe_dist) continue;
    const double sq[8] = { bx + radius, by + radius, bx + radius, by - radius, bx - radius, by - radius, bx - radius, by + radius };
    if (orc_gjk_collision(sq, 4, cps, 4)) return 1;
  }
  return 0;
}

/* one jerk sample of expandAndAddToQueue: kinematics and admissibility (:1071-1155 / :1260-1339).
 * Returns 1 when the primitive survives; fills end, cx, cy, Q of `nb`. */
static int primitive(const ksearch* s, const double* ist, int comb, int root, knode* nb)
{
  const orc_search_par* par = s->par;
  const double tau = par->T, j_max = par->j_max, j_min = -par->j_max, a_max = par->a_max, a_min = -par->a_max;
  const double v_max = par->v_max, v_min = -par->v_max;
  const int ns = par->num_samples;
  const double delta_x = (j_max - j_min) / (ns - 1);
  const int jx = comb / ns, jy = comb % ns;
  const double ji[2] = { j_min + jx * delta_x, j_min + jy * delta_x };
  double* e = nb->end;
  for (int d = 0; d < 2; d++)
  {
    e[d] = ist[d] + ist[2 + d] * tau + ist[4 + d] * tau * tau / 2 + ji[d] * tau * tau * tau / 6;
    e[2 + d] = ist[2 + d] + ist[4 + d] * tau + ji[d] * tau * tau / 2;
    e[4 + d] = ist[4 + d] + ji[d] * tau;
  }
  double n2 = 0;
  for (int k = 0; k < 6; k++) n2 += (e[k] - ist[k]) * (e[k] - ist[k]);
  if (sqrt(n2) < 0.00001) return 0;
  if (e[5] > a_max || e[5] < a_min || e[4] > a_max || e[4] < a_min) return 0;
  nb->cx[0] = ji[0] / 6, nb->cx[1] = ist[4] / 2, nb->cx[2] = ist[2], nb->cx[3] = ist[0];
  nb->cy[0] = ji[1] / 6, nb->cy[1] = ist[5] / 2, nb->cy[2] = ist[3], nb->cy[3] = ist[1];
  for (int i = 0; i < 4; i++)
  {
    double qx = 0, qy = 0;
    for (int k = 0; k < 4; k++) qx += nb->cx[k] * s->Ainv[k * 4 + i], qy += nb->cy[k] * s->Ainv[k * 4 + i];
    nb->Q[i] = qx, nb->Q[4 + i] = qy;
  }
  const double bx = s->cx.pb[2 * s->cx.self], by = s->cx.pb[2 * s->cx.self + 1];
  for (int i = 0; i < 4; i++)
  {
    if (nb->Q[i] < par->x_min || nb->Q[i] > par->x_max || nb->Q[4 + i] < par->y_min || nb->Q[4 + i] > par->y_max ||
        norm2(nb->Q[i] - bx, nb->Q[4 + i] - by) > par->tether)
      return 0;
  }
  if (!root)
  {
    for (int i = 0; i < 3; i++)
    {
      double vx = 0, vy = 0;
      for (int k = 0; k < 3; k++) vx += nb->cx[k] * s->V[k * 3 + i], vy += nb->cy[k] * s->V[k * 3 + i];
      if (vx < v_min || vx > v_max || vy < v_min || vy > v_max) return 0;
    }
  }
  if (e[4] > 0 && e[2] - 0.5 * e[4] * e[4] / j_min > v_max) return 0;
  else if (e[4] < 0 && e[2] - 0.5 * e[4] * e[4] / j_max < v_min) return 0;
  if (e[5] > 0 && e[3] - 0.5 * e[5] * e[5] / j_min > v_max) return 0;
  else if (e[5] < 0 && e[3] - 0.5 * e[5] * e[5] / j_max < v_min) return 0;
  return 1;
}

/* expandAndAddToQueue: cur < 0 is the root form (:1240-1385), otherwise the node form (:1045-1228) */
static void expand(ksearch* s, int cur)
{
  const orc_search_par* par = s->par;
  const orc_search_in* in = s->in;
  const int root = cur < 0;
  const int nchild = par->num_samples * par->num_samples;
  double ist[6];
  memcpy(ist, root ? in->init : s->pool[cur].end, sizeof(ist));
  orc_ent es0 = { in->es_cnt[0], in->es_cnt[1], (int*)in->es_alpha, (double*)in->es_beta, (int*)in->es_bend, (int*)in->es_active };
  for (int c = 0; c < nchild; c++)
  {
    if (!root && s->n_used == par->max_nodes - 1) return; /* "run out of memory" :1060-1064 */
    knode* nb = &s->pool[s->n_used];
    node_lists(s, s->n_used);
    nb->index = root ? 1 : s->pool[cur].index + 1;
    nb->prev = cur;
    if (!primitive(s, ist, in->comb[c], root, nb)) continue;
    copy_state(s, &nb->es, root ? &es0 : &s->pool[cur].es);
    double arc = 0.0;
    if (par->enable_entangle)
    {
      if (entangles(s, nb, &arc)) continue;
    }
    else
      arc = norm2(nb->end[0] - ist[0], nb->end[1] - ist[1]);
    const int iz = get_iz(&nb->es);
    const int ix = voxel_index(nb->end[0], par->voxel_size), iy = voxel_index(nb->end[1], par->voxel_size);
    nb->g = (root ? 0.0 : s->pool[cur].g) + arc;
    nb->h = norm2(nb->end[0] - in->goal[0], nb->end[1] - in->goal[1]) + 0.3 * (double)nb->es.n_alpha + 1.0 * (double)nb->es.n_bend;
    if (!root)
    {
      const int f = hash_find(s, ix, iy, iz);
      if (f >= 0)
      {
        knode* fn = &s->pool[f];
        if (fn->state == 1 && fn->index == nb->index)
        {
          if (nb->g + par->bias * nb->h < fn->g + par->bias * fn->h && s->ran_trigger % 2 == 0)
          { /* :1193-1205: kinematics replaced, entangle state and heap position kept */
            fn->prev = cur;
            fn->g = nb->g, fn->h = nb->h;
            memcpy(fn->end, nb->end, sizeof(fn->end));
            memcpy(fn->cx, nb->cx, sizeof(fn->cx));
            memcpy(fn->cy, nb->cy, sizeof(fn->cy));
            memcpy(fn->Q, nb->Q, sizeof(fn->Q));
          }
          s->ran_trigger++;
        }
        continue;
      }
    }
    nb->state = 1;
    heap_push(s, s->n_used);
    hash_insert(s, ix, iy, iz, s->n_used);
    s->n_used++;
  }
}

int orc_search(const orc_search_par* par, const orc_search_in* in, orc_search_out* out)
{
  const int N = par->N, M = par->M, NA = N + M, cap = par->ecap, ocap = par->out_cap;
  ksearch s;
  memset(&s, 0, sizeof(s));
  s.par = par, s.in = in;
  s.cx.N = N, s.cx.M = M, s.cx.self = in->agent_id - 1, s.cx.cap = cap;
  s.cx.pb = in->pb, s.cx.strep = in->strep, s.cx.bp_cnt = in->bp_cnt, s.cx.bp_xy = in->bp_xy, s.cx.bp_max = par->bp_max;
  orc_basis(par->T, s.Ainv, s.V, 0);
  const int maxn = par->max_nodes;
  s.pool = (knode*)calloc((size_t)maxn, sizeof(knode));
  s.alpha = (int*)malloc(sizeof(int) * 2 * (size_t)cap * maxn);
  s.beta = (double*)malloc(sizeof(double) * (size_t)cap * maxn);
  s.bend = (int*)malloc(sizeof(int) * (size_t)cap * maxn);
  s.active = (int*)malloc(sizeof(int) * (size_t)NA * maxn);
  s.heap = (int*)malloc(sizeof(int) * (size_t)maxn);
  s.hcap = 64;
  while (s.hcap < 2 * maxn) s.hcap *= 2;
  s.hkey = (int*)calloc((size_t)s.hcap * 4, sizeof(int));
  s.tcap = NA + par->bp_max + 2;
  s.toadd = (int*)malloc(sizeof(int) * 2 * (size_t)s.tcap);
  s.act_old = (int*)malloc(sizeof(int) * NA);

  /* setUp: is the goal occupied by some other agent's last hull? (:213-229) */
  {
    const double r = 0.5, gx = in->goal[0], gy = in->goal[1];
    const double gh[8] = { gx + r, gy + r, gx + r, gy - r, gx - r, gy + r, gx - r, gy - r };
    for (int o = 0; o < N && !s.goal_occupied; o++)
    {
      const int hn = in->hull_cnt[o * ORC_NPOL_MAX + par->num_pol - 1];
      if (hn > 0 && orc_gjk_collision(in->hull_xy + ((size_t)(o * ORC_NPOL_MAX + par->num_pol - 1) * ORC_SEARCH_HSTRIDE) * 2, hn, gh, 4))
        s.goal_occupied = 1;
    }
  }

  /* run (:1629-1827) */
  int status = 2, pops = 0, cur = -1, closest = -1;
  double smallest = DBL_MAX;
  if (in->es_cnt[0] > cap || in->es_cnt[1] > cap)
    s.overflow = 1; /* entangle_state_A does not fit a search node: storage error, nothing searched */
  else
    expand(&s, -1);
  while (s.heap_n > 0)
  {
    if (pops >= par->max_expansions)
    {
      status = 0;
      break;
    }
    pops++;
    cur = heap_pop(&s);
    knode* nd = &s.pool[cur];
    nd->state = -1;
    const double dist = norm2(nd->end[0] - in->goal[0], nd->end[1] - in->goal[1]);
    const double dist_init = norm2(nd->end[0] - in->init[0], nd->end[1] - in->init[1]);
    if (collides_solve(&s, nd->Q, nd->index)) continue;
    if (collides_bases(&s, nd->Q)) continue;
    int valid = 1;
    for (int i = 0; i < N; i++)
      if (nd->es.active[i] > 1)
      {
        valid = 0;
        break;
      }
    const double dcmp = s.goal_occupied ? dist * dist : dist_init;
    const double dti = dcmp * (double)nd->index;
    if (dti < smallest && valid)
    {
      smallest = dti;
      closest = cur;
    }
    if (dist < par->goal_size && valid)
    {
      status = 1;
      break;
    }
    expand(&s, cur);
  }
  int best = -1;
  if (status == 1)
    best = cur;
  else if (closest >= 0 && par->use_not_reaching)
    best = closest;

  out->status[0] = status;
  out->solved[0] = best >= 0;
  out->stats[0] = s.n_used, out->stats[1] = pops, out->stats[2] = best >= 0 ? s.pool[best].index : 0;
  out->stats[3] = s.goal_occupied;
  out->cost[0] = best >= 0 ? s.pool[best].g : 0.0;
  /* recoverPwpOut (:521-553), recoverEntStateVector (:582-603) */
  memset(out->coeff, 0, sizeof(double) * 3 * ORC_NPOL_MAX * 4);
  int n = 0;
  if (best >= 0)
  {
    int path[ORC_NPOL_MAX];
    for (int t = best; t >= 0; t = s.pool[t].prev)
      if (s.pool[t].index <= par->num_pol) path[s.pool[t].index - 1] = t, n = n > s.pool[t].index ? n : s.pool[t].index;
    for (int i = 0; i <= ORC_NPOL_MAX; i++)
    {
      const int src = i == 0 ? -1 : path[(i <= n ? i : n) - 1];
      if (n == 0 && i > 0) break;
      const int na = src < 0 ? in->es_cnt[0] : s.pool[src].es.n_alpha, nbd = src < 0 ? in->es_cnt[1] : s.pool[src].es.n_bend;
      const int* al = src < 0 ? in->es_alpha : s.pool[src].es.alpha;
      const double* be = src < 0 ? in->es_beta : s.pool[src].es.beta;
      const int* bd = src < 0 ? in->es_bend : s.pool[src].es.bend;
      const int* ac = src < 0 ? in->es_active : s.pool[src].es.active;
      out->esv_cnt[2 * i] = na, out->esv_cnt[2 * i + 1] = nbd;
      memset(out->esv_alpha + (size_t)i * 2 * ocap, 0, sizeof(int) * 2 * ocap);
      memset(out->esv_beta + (size_t)i * ocap, 0, sizeof(double) * ocap);
      memset(out->esv_bend + (size_t)i * ocap, 0, sizeof(int) * ocap);
      if (na > ocap || nbd > ocap)
      {
        s.overflow = 1;
        continue;
      }
      memcpy(out->esv_alpha + (size_t)i * 2 * ocap, al, sizeof(int) * 2 * na);
      memcpy(out->esv_beta + (size_t)i * ocap, be, sizeof(double) * na);
      memcpy(out->esv_bend + (size_t)i * ocap, bd, sizeof(int) * nbd);
      memcpy(out->esv_active + (size_t)i * NA, ac, sizeof(int) * NA);
    }
    for (int i = 0; i < n; i++)
    {
      const knode* nd = &s.pool[path[i]];
      memcpy(out->coeff + (0 * ORC_NPOL_MAX + i) * 4, nd->cx, sizeof(double) * 4);
      memcpy(out->coeff + (1 * ORC_NPOL_MAX + i) * 4, nd->cy, sizeof(double) * 4);
      memcpy(out->coeff + (2 * ORC_NPOL_MAX + i) * 4, in->coeffs_z + 4 * i, sizeof(double) * 4);
    }
  }
  out->n_int[0] = n;
  const int ovf = s.overflow;
  free(s.pool), free(s.alpha), free(s.beta), free(s.bend), free(s.active), free(s.heap), free(s.hkey), free(s.toadd), free(s.act_old);
  return ovf ? -3 : 0;
}

/* batch driver: every array carries a leading [B] dimension (shared ones noted in the header) */
int orc_search_batch(const orc_search_par* par, const orc_search_batch_t* b, int nthreads)
{
  const int N = par->N, M = par->M, NA = N + M, S = par->S, np = par->num_pol, ocap = par->out_cap;
  const int nchild = par->num_samples * par->num_samples;
  int rc = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
  for (int i = 0; i < b->B; i++)
  {
    orc_search_in in;
    orc_search_out out;
    const int g = b->group ? b->group[i] : i;
    in.agent_id = b->agent_id[i];
    memcpy(in.init, b->init + 6 * (size_t)i, sizeof(in.init));
    in.goal[0] = b->goal[2 * i], in.goal[1] = b->goal[2 * i + 1];
    in.coeffs_z = b->coeffs_z + (size_t)i * ORC_NPOL_MAX * 4;
    /* hulls of the window group, with the agent's own slot and unknown agents masked out */
    int* hc = (int*)malloc(sizeof(int) * N * ORC_NPOL_MAX);
    memcpy(hc, b->hull_cnt + (size_t)g * N * ORC_NPOL_MAX, sizeof(int) * N * ORC_NPOL_MAX);
    const unsigned char* known = b->known + (size_t)i * N;
    for (int o = 0; o < N; o++)
      if (o == in.agent_id - 1 || !known[o])
        for (int k = 0; k < ORC_NPOL_MAX; k++) hc[o * ORC_NPOL_MAX + k] = 0;
    in.hull_cnt = hc;
    in.hull_xy = b->hull_xy + (size_t)g * N * ORC_NPOL_MAX * ORC_SEARCH_HSTRIDE * 2;
    in.samp = b->samp + (size_t)g * N * np * (S + 1) * 2;
    in.known = known;
    in.st_ptr = b->st_ptr, in.st_xy = b->st_xy, in.strep = b->strep, in.st_longest = b->st_longest;
    in.pb = b->pb, in.bp_cnt = b->bp_cnt, in.bp_xy = b->bp_xy;
    in.es_cnt = b->es_cnt + 2 * (size_t)i;
    in.es_alpha = b->es_alpha + (size_t)i * 2 * b->es_cap;
    in.es_beta = b->es_beta + (size_t)i * b->es_cap;
    in.es_bend = b->es_bend + (size_t)i * b->es_cap;
    in.es_active = b->es_active + (size_t)i * NA;
    in.comb = b->comb + (b->comb_shared ? 0 : (size_t)i * nchild);
    out.status = b->status + i, out.solved = b->solved + i, out.n_int = b->n_int + i;
    out.coeff = b->coeff + (size_t)i * 3 * ORC_NPOL_MAX * 4;
    out.esv_cnt = b->esv_cnt + (size_t)i * 9 * 2;
    out.esv_alpha = b->esv_alpha + (size_t)i * 9 * 2 * ocap;
    out.esv_beta = b->esv_beta + (size_t)i * 9 * ocap;
    out.esv_bend = b->esv_bend + (size_t)i * 9 * ocap;
    out.esv_active = b->esv_active + (size_t)i * 9 * NA;
    out.stats = b->stats + 4 * (size_t)i;
    out.cost = b->cost + i;
    int r = orc_search(par, &in, &out);
    free(hc);
    if (r)
    {
#pragma omp critical
      rc = r;
    }
  }
  return rc;
}

/* Test hook: replays a script of open-list operations through the heap restatement above so that tests can
 * compare it with the real std::priority_queue of this image's libstdc++ (tests/cpp/heap_check.cpp).
 * ops[k] = {kind, id}: 0 push id with (g,h) = vals[k]; 1 pop (id appended to out); 2 overwrite (g,h) of a node
 * that may be inside the heap (what the "better node" rule of :1193-1205 does).  Returns pops written. */
int orc_heap_replay(int n_ops, const int* ops, const double* vals, int n_ids, double bias, int* out)
{
  ksearch s;
  orc_search_par par;
  memset(&s, 0, sizeof(s));
  memset(&par, 0, sizeof(par));
  par.bias = bias;
  s.par = &par;
  s.pool = (knode*)calloc((size_t)n_ids, sizeof(knode));
  s.heap = (int*)malloc(sizeof(int) * (size_t)(n_ops + 1));
  int n_out = 0;
  for (int k = 0; k < n_ops; k++)
  {
    const int kind = ops[2 * k], id = ops[2 * k + 1];
    if (kind == 0)
    {
      s.pool[id].g = vals[2 * k], s.pool[id].h = vals[2 * k + 1];
      heap_push(&s, id);
    }
    else if (kind == 1)
    {
      if (s.heap_n > 0) out[n_out++] = heap_pop(&s);
    }
    else
      s.pool[id].g = vals[2 * k], s.pool[id].h = vals[2 * k + 1];
  }
  free(s.pool), free(s.heap);
  return n_out;
}

/* NeptuneRos::setUpCheckingPosAndStaticObs (neptune_ros.cpp:852-1019): representative points of every static
 * obstacle for one agent -- where a line through the obstacle's centre, at an angle found by a one-degree sweep that
 * starts along base -> pos, leaves the polygon -- and staticObsLongestDist.  poly: raw (un-inflated) obstacles.
 * Returns 0, or -1 where the reference prints "cannot find a feasible vertex representation" and exits. */
int orc_static_obst_rep(int M, const long long* ptr, const double* xy, const double base[2], const double pos[2],
                        double voxel, double* strep, double* longest)
{
  double* cen = (double*)malloc(sizeof(double) * 2 * (M > 0 ? M : 1));
  for (int i = 0; i < M; i++)
  {
    double sx = 0.0, sy = 0.0;
    const int nv = (int)(ptr[i + 1] - ptr[i]);
    for (int j = 0; j < nv; j++) sx += xy[2 * (ptr[i] + j)], sy += xy[2 * (ptr[i] + j) + 1];
    cen[2 * i] = sx / nv, cen[2 * i + 1] = sy / nv;
  }
  const double c11 = base[0], c22 = base[1], d11 = pos[0], d22 = pos[1];
  double theta = atan2(d22 - c22, d11 - c11);
  int ok = 0, rc = 0;
  while (!ok)
  {
    if (theta > atan2(d22 - c22, d11 - c11) + 3.14)
    {
      rc = -1;
      break;
    }
    const double m = tan(theta);
    int fail = 0;
    for (int i = 0; i < M && !fail; i++)
    {
      const double e1 = cen[2 * i + 1] - cen[2 * i] * m;
      for (int j = i + 1; j < M; j++)
      {
        const double e2 = cen[2 * j + 1] - cen[2 * j] * m;
        if (fabs(e1 - e2) / sqrt(1 + m * m) < voxel * 1.6) fail = 1;
      }
      if (fail) break;
      const double e2 = base[1] - base[0] * m;
      if (fabs(e1 - e2) / sqrt(1 + m * m) < voxel * 1.6)
      {
        fail = 1;
        break;
      }
      const double a1 = cen[2 * i], a2 = cen[2 * i + 1], b1 = cen[2 * i] + 10, b2 = cen[2 * i + 1] + 10 * m;
      const double den = (b2 - a2) * (d11 - c11) - (b1 - a1) * (d22 - c22);
      const double y = ((c22 - a2) * (b1 - a1) - (b2 - a2) * (c11 - a1)) / den;
      if (fabs(den) < 0.01 || y < 0 || y > 1.0)
      {
      }
      else
        fail = 1;
    }
    for (int i = 0; i < M && !fail; i++)
    {
      const int nv = (int)(ptr[i + 1] - ptr[i]);
      const double* v = xy + 2 * ptr[i];
      double x_max = -1e-5, x_min = 1e-5, maxv[2] = { 0, 0 }, minv[2] = { 0, 0 };
      for (int j = 0; j < nv; j++)
      {
        const double a1 = cen[2 * i], a2 = cen[2 * i + 1], b1 = cen[2 * i] + 10, b2 = cen[2 * i + 1] + 10 * m;
        const double c1 = v[2 * j], c2 = v[2 * j + 1], d1 = v[2 * ((j + 1) % nv)], d2 = v[2 * ((j + 1) % nv) + 1];
        const double den = (b2 - a2) * (d1 - c1) - (b1 - a1) * (d2 - c2);
        const double y = ((c2 - a2) * (b1 - a1) - (b2 - a2) * (c1 - a1)) / den;
        if (fabs(den) < 0.01 || y < 0 || y > 1.0) continue;
        double x;
        if (fabs(b1 - a1) < 0.01)
          x = (c2 - a2) / (b2 - a2) + y * (d2 - c2) / (b2 - a2);
        else
          x = (c1 - a1) / (b1 - a1) + y * (d1 - c1) / (b1 - a1);
        if (x > x_max)
          x_max = x, maxv[0] = c1 + y * (d1 - c1), maxv[1] = c2 + y * (d2 - c2);
        else if (x < x_min)
          x_min = x, minv[0] = c1 + y * (d1 - c1), minv[1] = c2 + y * (d2 - c2);
      }
      if (x_max < 0 || x_min > 0)
      {
        fail = 1;
        break;
      }
      strep[4 * i] = minv[0], strep[4 * i + 1] = minv[1], strep[4 * i + 2] = maxv[0], strep[4 * i + 3] = maxv[1];
      double l1 = 0, l2 = 0;
      for (int j = 0; j < nv; j++)
      {
        const double q1 = sqrt((v[2 * j] - minv[0]) * (v[2 * j] - minv[0]) + (v[2 * j + 1] - minv[1]) * (v[2 * j + 1] - minv[1]));
        const double q2 = sqrt((v[2 * j] - maxv[0]) * (v[2 * j] - maxv[0]) + (v[2 * j + 1] - maxv[1]) * (v[2 * j + 1] - maxv[1]));
        if (q1 > l1) l1 = q1;
        if (q2 > l2) l2 = q2;
      }
      longest[2 * i] = l1, longest[2 * i + 1] = l2;
    }
    if (!fail) ok = 1;
    theta = theta + 1.0 / 180.0 * 3.1415927;
  }
  free(cen);
  return rc;
}

/* eu::getTetherLength on an explicit state (test hook for tests/test_reference_pin.py) */
double orc_tether_length_state(const orc_ent* es, const orc_ectx* cx, const double* st_longest, const double pk1[2])
{
  ksearch s;
  orc_search_in in;
  memset(&s, 0, sizeof(s));
  memset(&in, 0, sizeof(in));
  in.st_longest = st_longest;
  s.in = &in;
  s.cx = *cx;
  return tether_length(&s, es, pk1);
}
