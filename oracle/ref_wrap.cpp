// ref_wrap.cpp -- C entry points around the REFERENCE's own entanglement and GJK code (TEST INFRASTRUCTURE).
//
// oracle/Makefile (target _ref) compiles neptune/src/entangle_utils.cpp and neptune/src/gjk.cpp where they lie under
// /root/reference, unmodified, against the Eigen stand-in in oracle/eigen_shim (the image has no Eigen), together with
// this file, into oracle/_ref/libneptune_ref.so.  The wrappers only convert plain arrays to the reference's argument
// types and call eu:: / gjk:: functions; ref_chain additionally replays the caller loop of
// KinodynamicSearch::entanglesWithOtherAgents (kinodynamic_search.cpp:813-891, which itself needs ROS and cannot be
// compiled) around those calls.  tests/test_reference_pin.py uses the library to pin the oracle's restatement, and
// tests/golden/make_ref_golden.py to record golden vectors that travel to machines without /root/reference.
#include <vector>

#include "entangle_utils.hpp"
#include "gjk.hpp"

typedef Eigen::Vector2d V2;

static std::vector<V2> pts(const double* p, int n)
{
  std::vector<V2> v;
  for (int i = 0; i < n; i++) v.push_back(V2(p[2 * i], p[2 * i + 1]));
  return v;
}

static int dump(const std::vector<Eigen::Vector2i>& a, int* out, int cap)
{
  for (size_t i = 0; i < a.size() && (int)i < cap; i++) out[2 * i] = a[i](0), out[2 * i + 1] = a[i](1);
  return (int)a.size();
}

extern "C" int ref_gjk_collision(const double* v1, int n1, const double* v2, int n2)
{
  Eigen::Matrix<double, 2, Eigen::Dynamic> a(2, n1), b(2, n2);
  for (int i = 0; i < n1; i++) a(0, i) = v1[2 * i], a(1, i) = v1[2 * i + 1];
  for (int i = 0; i < n2; i++) b(0, i) = v2[2 * i], b(1, i) = v2[2 * i + 1];
  return gjk::collision(a, b) ? 1 : 0;
}

extern "C" int ref_hsig_agent(int* out, int cap, const double* pk, const double* pk1, const double* pik, const double* pik1,
                              const double* pb, const double* bend, int nbend, int agent_id)
{
  std::vector<Eigen::Vector2i> add;
  std::vector<V2> bp = pts(bend, nbend);
  V2 p1(pik1[0], pik1[1]);
  eu::entangleHSigToAddAgentInd(add, V2(pk[0], pk[1]), V2(pk1[0], pk1[1]), V2(pik[0], pik[1]), p1, V2(pb[0], pb[1]), bp, agent_id);
  return dump(add, out, cap);
}

// the 9-argument form; the caller must not pass inputs on which the reference calls exit(-1)
extern "C" int ref_hsig_agent9(int* out, int cap, const double* pk, const double* pk1, const double* pik, const double* pik1,
                               const double* pb, const double* bend, int nbend, const double* prev, int nprev, int agent_id)
{
  std::vector<Eigen::Vector2i> add;
  std::vector<V2> bp = pts(bend, nbend), bpp = pts(prev, nprev);
  V2 p1(pik1[0], pik1[1]);
  eu::entangleHSigToAddAgentInd(add, V2(pk[0], pk[1]), V2(pk1[0], pk1[1]), V2(pik[0], pik[1]), p1, V2(pb[0], pb[1]), bp, bpp, agent_id);
  return dump(add, out, cap);
}

static std::vector<Eigen::Matrix<double, 2, 2>> reps(const double* strep, int M)
{
  std::vector<Eigen::Matrix<double, 2, 2>> r;
  for (int m = 0; m < M; m++)
  {
    Eigen::Matrix<double, 2, 2> q;
    q(0, 0) = strep[4 * m], q(1, 0) = strep[4 * m + 1], q(0, 1) = strep[4 * m + 2], q(1, 1) = strep[4 * m + 3];
    r.push_back(q);
  }
  return r;
}

extern "C" int ref_hsig_static(int* out, int cap, const double* pk, const double* pk1, const double* strep, int M, int N)
{
  std::vector<Eigen::Vector2i> add;
  std::vector<Eigen::Matrix<double, 2, 2>> rep = reps(strep, M);
  eu::entangleHSigToAddStatic(add, V2(pk[0], pk[1]), V2(pk1[0], pk1[1]), rep, N);
  return dump(add, out, cap);
}

// The chain along a piecewise-cubic path, n intervals of S steps (entanglesWithOtherAgents :813-891 per interval,
// sample times of :116-127).  State in / out: alphas [cap][2], betas [cap], bend [cap], active [N+M], counts.
// Per interval i (0-based) the state AFTER it is written to out_* [i+1] (out_*[0] = the input state), the tether length
// to out_len[i].  Returns the number of intervals completed before the first entangling step (n if none).
extern "C" int ref_chain(int N, int M, int self, const double* pb, const double* strep, const double* longest,
                         const int* bp_cnt, const double* bp_xy, int bp_max, const unsigned char* known, const double* samp,
                         int num_pol, int S, double T, int n, const double* cxy /*[2][n][4]*/, int cap, const int* cnt0,
                         const int* alpha0, const double* beta0, const int* bend0, const int* active0, int* out_cnt,
                         int* out_alpha, double* out_beta, int* out_bend, int* out_active, double* out_len)
{
  std::vector<V2> vpb = pts(pb, N);
  std::vector<Eigen::Matrix<double, 2, 2>> rep = reps(strep, M);
  std::vector<V2> vlong = pts(longest, M);
  std::vector<std::vector<V2>> bends(N);
  for (int j = 0; j < N; j++) bends[j] = pts(bp_xy + (size_t)2 * bp_max * j, bp_cnt[j]);
  V2 base = vpb[self];
  eu::ent_state st;
  for (int i = 0; i < cnt0[0]; i++) st.alphas.push_back(Eigen::Vector2i(alpha0[2 * i], alpha0[2 * i + 1])), st.betas.push_back(beta0[i]);
  for (int i = 0; i < cnt0[1]; i++) st.bendPointsIdx.push_back(bend0[i]);
  for (int i = 0; i < N + M; i++) st.active_cases.push_back(active0[i]);
  const int NA = N + M;
  auto store = [&](int slot) {
    out_cnt[2 * slot] = (int)st.alphas.size(), out_cnt[2 * slot + 1] = (int)st.bendPointsIdx.size();
    for (size_t i = 0; i < st.alphas.size() && (int)i < cap; i++)
    {
      out_alpha[((size_t)slot * cap + i) * 2] = st.alphas[i](0), out_alpha[((size_t)slot * cap + i) * 2 + 1] = st.alphas[i](1);
      out_beta[(size_t)slot * cap + i] = st.betas[i];
    }
    for (size_t i = 0; i < st.bendPointsIdx.size() && (int)i < cap; i++) out_bend[(size_t)slot * cap + i] = st.bendPointsIdx[i];
    for (int i = 0; i < NA; i++) out_active[(size_t)slot * NA + i] = st.active_cases[i];
  };
  store(0);
  int done = n;
  for (int ii = 0; ii < n; ii++)
  {
    const double* x = cxy + 4 * ii;
    const double* y = cxy + 4 * n + 4 * ii;
    V2 pk(x[3], y[3]), pk1 = pk;
    std::vector<int> act_old = st.active_cases;
    bool ent = false;
    for (int j = 1; j <= S && !ent; j++)
    {
      const double t = (j < S) ? T * j / S : T;
      const double t3 = t * t * t, t2 = t * t;
      pk1 = V2(x[0] * t3 + x[1] * t2 + x[2] * t + x[3], y[0] * t3 + y[1] * t2 + y[2] * t + y[3]);
      std::vector<Eigen::Vector2i> add;
      for (int a = 0; a < N; a++)
      {
        if (a == self || !known[a]) continue;
        const double *p0, *p1;
        if (ii > num_pol - 1)
          p0 = p1 = samp + ((size_t)(a * num_pol + (num_pol - 1)) * (S + 1) + S) * 2;
        else
          p0 = samp + ((size_t)(a * num_pol + ii) * (S + 1) + (j - 1)) * 2, p1 = samp + ((size_t)(a * num_pol + ii) * (S + 1) + j) * 2;
        V2 pik(p0[0], p0[1]), pik1(p1[0], p1[1]);
        eu::entangleHSigToAddAgentInd(add, pk, pk1, pik, pik1, base, bends[a], a + 1);
      }
      eu::entangleHSigToAddStatic(add, pk, pk1, rep, N);
      if ((int)(st.alphas.size() + add.size()) > NA)
      {
        ent = true;
        break;
      }
      eu::addAlphaBetaToList(add, st, pk, vpb, base, rep, N, bends);
      for (int a = 0; a < N; a++)
      {
        if (act_old[a] < 2 && st.active_cases[a] >= 2) ent = true;
        else if (act_old[a] >= 2 && st.active_cases[a] > act_old[a]) ent = true;
      }
      if (ent) break;
      eu::updateBendPts(st, pk1, vpb, base, rep, N);
      act_old = st.active_cases;
      pk = pk1;
    }
    if (ent)
    {
      if (done == n) done = ii;
      break;
    }
    out_len[ii] = eu::getTetherLength(st, vpb, base, pk1, rep, vlong, N);
    store(ii + 1);
  }
  return done;
}

// One tick of the online tracker: the loop of NeptuneRos::updateEntStateStaticObs (neptune_ros.cpp:798-850, a ROS file that
// cannot be compiled) replayed around the reference's own functions -- the 9-argument entangleHSigToAddAgentInd,
// entangleHSigToAddStatic, addAlphaBetaToList, updateBendPts.  elapsed_ms stands for the wall-clock timer of :803-804.
// Returns 0 updated, 1 skipped by the gate; where the reference calls exit(-1) this process exits too (the test forks).
extern "C" int ref_track(int N, int M, int self, const double* pb, const double* strep, const int* bp_cnt, const double* bp_xy,
                         const int* bp_cnt_prev, const double* bp_xy_prev, int bp_max, int cap, int* cnt, int* alpha, double* beta,
                         int* bend, int* active, double* prev_pos /*[N+1][2]*/, double* prev_pos_agent /*[N][2]*/,
                         const double* latest /*[N][2]*/, const double* cur, double elapsed_ms)
{
  V2 pk1(cur[0], cur[1]);
  {
    V2 last(prev_pos[2 * N], prev_pos[2 * N + 1]);
    if ((last - pk1).norm() < 0.05 && elapsed_ms < 100) return 1;
  }
  std::vector<V2> vpb = pts(pb, N);
  std::vector<Eigen::Matrix<double, 2, 2>> rep = reps(strep, M);
  std::vector<std::vector<V2>> bends(N), bends_prev(N);
  for (int j = 0; j < N; j++) bends[j] = pts(bp_xy + (size_t)2 * bp_max * j, bp_cnt[j]);
  for (int j = 0; j < N; j++) bends_prev[j] = pts(bp_xy_prev + (size_t)2 * bp_max * j, bp_cnt_prev[j]);
  eu::ent_state st;
  for (int i = 0; i < cnt[0]; i++) st.alphas.push_back(Eigen::Vector2i(alpha[2 * i], alpha[2 * i + 1])), st.betas.push_back(beta[i]);
  for (int i = 0; i < cnt[1]; i++) st.bendPointsIdx.push_back(bend[i]);
  for (int i = 0; i < N + M; i++) st.active_cases.push_back(active[i]);

  std::vector<Eigen::Vector2i> add;
  for (int i = 0; i < N; i++)
  {
    if (i == self) continue;
    if (prev_pos_agent[2 * i] < -900 || bends[i].empty()) continue;
    V2 lat(latest[2 * i], latest[2 * i + 1]);
    eu::entangleHSigToAddAgentInd(add, V2(prev_pos[2 * i], prev_pos[2 * i + 1]), pk1, V2(prev_pos_agent[2 * i], prev_pos_agent[2 * i + 1]),
                                  lat, vpb[self], bends[i], bends_prev[i], i + 1);
    prev_pos[2 * i] = cur[0], prev_pos[2 * i + 1] = cur[1];
    prev_pos_agent[2 * i] = latest[2 * i], prev_pos_agent[2 * i + 1] = latest[2 * i + 1];
  }
  V2 last(prev_pos[2 * N], prev_pos[2 * N + 1]);
  eu::entangleHSigToAddStatic(add, last, pk1, rep, N);
  eu::addAlphaBetaToList(add, st, last, vpb, vpb[self], rep, N, bends);
  eu::updateBendPts(st, pk1, vpb, vpb[self], rep, N);
  prev_pos[2 * N] = cur[0], prev_pos[2 * N + 1] = cur[1];

  cnt[0] = (int)st.alphas.size(), cnt[1] = (int)st.bendPointsIdx.size();
  for (size_t i = 0; i < st.alphas.size() && (int)i < cap; i++) alpha[2 * i] = st.alphas[i](0), alpha[2 * i + 1] = st.alphas[i](1), beta[i] = st.betas[i];
  for (size_t i = 0; i < st.bendPointsIdx.size() && (int)i < cap; i++) bend[i] = st.bendPointsIdx[i];
  for (int i = 0; i < N + M; i++) active[i] = st.active_cases[i];
  return 0;
}

// Neptune::PredictAlphasBetas (neptune.cpp:976-1008; neptune.cpp needs CGAL, Gurobi and ROS) replayed around the reference's
// own 8-argument crossing test, static crossing test, addAlphaBetaToList and updateBendPts.  samp0 [N][2] is
// SampledPointsForAll[i][0].col(0), known[i] == 0 an empty SampledPointsForAll[i].  State in / out as in ref_track.
extern "C" int ref_predict(int N, int M, int self, const double* pb, const double* strep, const int* bp_cnt, const double* bp_xy,
                           int bp_max, int cap, int* cnt, int* alpha, double* beta, int* bend, int* active, const double* prev_pos,
                           const double* prev_pos_agent, const double* cur, const double* samp0, const unsigned char* known)
{
  V2 pk1(cur[0], cur[1]);
  std::vector<V2> vpb = pts(pb, N);
  std::vector<Eigen::Matrix<double, 2, 2>> rep = reps(strep, M);
  std::vector<std::vector<V2>> bends(N);
  for (int j = 0; j < N; j++) bends[j] = pts(bp_xy + (size_t)2 * bp_max * j, bp_cnt[j]);
  eu::ent_state st;
  for (int i = 0; i < cnt[0]; i++) st.alphas.push_back(Eigen::Vector2i(alpha[2 * i], alpha[2 * i + 1])), st.betas.push_back(beta[i]);
  for (int i = 0; i < cnt[1]; i++) st.bendPointsIdx.push_back(bend[i]);
  for (int i = 0; i < N + M; i++) st.active_cases.push_back(active[i]);
  std::vector<Eigen::Vector2i> add;
  for (int i = 0; i < N; i++)
  {
    if (i == self || !known[i]) continue;
    V2 pik1(samp0[2 * i], samp0[2 * i + 1]);
    eu::entangleHSigToAddAgentInd(add, V2(prev_pos[2 * i], prev_pos[2 * i + 1]), pk1, V2(prev_pos_agent[2 * i], prev_pos_agent[2 * i + 1]), pik1,
                                  vpb[self], bends[i], i + 1);
  }
  V2 last(prev_pos[2 * N], prev_pos[2 * N + 1]);
  eu::entangleHSigToAddStatic(add, last, pk1, rep, N);
  eu::addAlphaBetaToList(add, st, last, vpb, vpb[self], rep, N, bends);
  eu::updateBendPts(st, pk1, vpb, vpb[self], rep, N);
  cnt[0] = (int)st.alphas.size(), cnt[1] = (int)st.bendPointsIdx.size();
  for (size_t i = 0; i < st.alphas.size() && (int)i < cap; i++) alpha[2 * i] = st.alphas[i](0), alpha[2 * i + 1] = st.alphas[i](1), beta[i] = st.betas[i];
  for (size_t i = 0; i < st.bendPointsIdx.size() && (int)i < cap; i++) bend[i] = st.bendPointsIdx[i];
  for (int i = 0; i < N + M; i++) active[i] = st.active_cases[i];
  return 0;
}
