// records_check.cpp -- nb::pwp2Record / record2Pwp / dynTraj2Record / updateTrajObstacles (neptune_b200/cpp/records_b200.hpp)
// against the REFERENCE's own mu::pwp2PwpMsg / mu::pwpMsg2Pwp (neptune/src/utils.cpp:180-261, compiled where it lies):
// the trajectory part of a record carries exactly the fields of the PieceWisePolTraj message the reference builds, and
// converting back gives exactly the trajectory the reference's pwpMsg2Pwp gives.  Host only (no GPU).  Prints "ok N".
#include <cstdio>
#include <cstdlib>
#include <random>

#include "utils.hpp"  // the reference's (mu::)
#include "../../neptune_b200/cpp/records_b200.hpp"

static bool same(const mt::PieceWisePol& a, const mt::PieceWisePol& b)
{
  if (a.times != b.times || a.coeff_x.size() != b.coeff_x.size()) return false;
  for (size_t i = 0; i < a.coeff_x.size(); i++)
    for (int c = 0; c < 4; c++)
      if (a.coeff_x[i](c) != b.coeff_x[i](c) || a.coeff_y[i](c) != b.coeff_y[i](c) || a.coeff_z[i](c) != b.coeff_z[i](c)) return false;
  return true;
}

int main()
{
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> u(-20.0, 20.0);
  int checked = 0;
  for (int it = 0; it < 500; it++)
  {
    const int n = 1 + (int)(rng() % 16);
    mt::PieceWisePol pwp;
    double t = u(rng);
    for (int i = 0; i <= n; i++) pwp.times.push_back(t), t += 0.05 + 0.5 * (rng() % 3);
    for (int i = 0; i < n; i++)
    {
      pwp.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(u(rng), u(rng), u(rng), u(rng)));
      pwp.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(u(rng), u(rng), u(rng), u(rng)));
      pwp.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(u(rng), u(rng), u(rng), u(rng)));
    }
    const mader_msgs::PieceWisePolTraj msg = mu::pwp2PwpMsg(pwp);
    double rec[NB_REC_DOUBLES];
    nb::DynTrajHeader h;
    h.id = 1 + (int)(rng() % 8), h.bbox[0] = h.bbox[1] = h.bbox[2] = 1.2, h.pos[0] = u(rng), h.pos[1] = u(rng), h.pos[2] = 1.0, h.seq = it;
    for (int q = 0; q < (int)(rng() % 5); q++) h.bendpt.push_back(Eigen::Vector2d(u(rng), u(rng)));
    if (!nb::dynTraj2Record(h, pwp, rec)) return 1;
    // field for field the reference's message
    if ((int)rec[0] != (int)msg.coeff_x.size() || msg.times.size() != (size_t)n + 1) return 2;
    for (int i = 0; i <= n; i++)
      if (rec[1 + i] != msg.times[i]) return 3;
    const double* co = rec + 18;
    for (int i = 0; i < n; i++)
    {
      const mader_msgs::CoeffPoly3 *x = &msg.coeff_x[i], *y = &msg.coeff_y[i], *z = &msg.coeff_z[i];
      const double want[12] = { x->a, x->b, x->c, x->d, y->a, y->b, y->c, y->d, z->a, z->b, z->c, z->d };
      for (int c = 0; c < 4; c++)
        if (co[4 * i + c] != want[c] || co[64 + 4 * i + c] != want[4 + c] || co[128 + 4 * i + c] != want[8 + c]) return 4;
    }
    // and back: what the reference's pwpMsg2Pwp returns
    if (!same(nb::record2Pwp(rec), mu::pwpMsg2Pwp(msg)) || !same(nb::record2Pwp(rec), pwp)) return 5;
    nb::DynTrajHeader h2;
    mt::PieceWisePol p2;
    nb::record2DynTraj(rec, h2, p2);
    if (h2.id != h.id || h2.bendpt.size() != h.bendpt.size() || h2.pos[1] != h.pos[1] || h2.seq != h.seq || !same(p2, pwp)) return 6;
    for (size_t q = 0; q < h.bendpt.size(); q++)
      if (h2.bendpt[q](0) != h.bendpt[q](0) || h2.bendpt[q](1) != h.bendpt[q](1)) return 7;
    // Neptune::updateTrajObstacles: replace by id, add when new
    std::vector<double> table((size_t)8 * NB_REC_DOUBLES, 0.0);
    std::vector<uint8_t> known(8, 0);
    if (!nb::updateTrajObstacles(table.data(), known.data(), 8, rec) || !known[h.id - 1]) return 8;
    if (table[(size_t)(h.id - 1) * NB_REC_DOUBLES + NB_REC_OFF_ID] != h.id) return 9;
    rec[NB_REC_OFF_ID] = 9;   // an id beyond num_of_agents_ is dropped (trajCB :381-384)
    if (nb::updateTrajObstacles(table.data(), known.data(), 8, rec)) return 10;
    checked++;
  }
  // a trajectory with more pieces than a record holds is refused, not truncated
  mt::PieceWisePol big;
  for (int i = 0; i <= 17; i++) big.times.push_back(i);
  for (int i = 0; i < 17; i++)
    big.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(0, 0, 0, 0)), big.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(0, 0, 0, 0)),
        big.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(0, 0, 0, 0));
  double rec[NB_REC_DOUBLES];
  if (nb::pwp2Record(big, rec)) return 11;
  printf("ok %d\n", checked);
  return 0;
}
