// records_b200.hpp -- the wire format of the exchange on the host side: committed-trajectory records
// (include/neptune_b200.h, NB_REC_DOUBLES = 256 doubles per agent) to and from the reference's types.
//
//   nb::pwp2Record / nb::record2Pwp          <- mu::pwp2PwpMsg / mu::pwpMsg2Pwp (neptune/src/utils.cpp:180-261): the
//                                               trajectory part of a record IS the PieceWisePolTraj message, field for field
//   nb::dynTraj2Record / nb::record2DynTraj  <- the mader_msgs/DynTraj a NeptuneRos::publishOwnTraj builds and a
//                                               NeptuneRos::trajCB reads (neptune_ros.cpp:379-480; DynTraj.msg:1-9)
//   nb::updateTrajObstacles                  <- Neptune::updateTrajObstacles (neptune.cpp:158-220) on a record table: the
//                                               trajectory of an id replaces the one held for that id, a new id is
//                                               added (the reference never removes one: ids_to_remove stays empty)
// The device side of the same conversions is nb_unpack_records_batch / nb_publish_records_batch.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/neptune_b200.h"
#include "nb_types.hpp"

namespace nb
{
// false when the trajectory has more than NB_REC_PIECES pieces (the record's capacity) or is malformed
inline bool pwp2Record(const mt::PieceWisePol& pwp, double* rec)
{
  const size_t n = pwp.coeff_x.size();
  if (n > NB_REC_PIECES || pwp.coeff_y.size() != n || pwp.coeff_z.size() != n || pwp.times.size() != n + 1) return false;
  for (int q = 0; q < NB_REC_PWP_DOUBLES; q++) rec[q] = 0.0;
  rec[0] = (double)n;
  for (size_t i = 0; i <= n; i++) rec[1 + i] = pwp.times[i];
  double* co = rec + 1 + (NB_REC_PIECES + 1);
  for (size_t i = 0; i < n; i++)
    for (int c = 0; c < 4; c++)
    {
      co[4 * i + c] = pwp.coeff_x[i](c);
      co[NB_REC_PIECES * 4 + 4 * i + c] = pwp.coeff_y[i](c);
      co[2 * NB_REC_PIECES * 4 + 4 * i + c] = pwp.coeff_z[i](c);
    }
  return true;
}

inline mt::PieceWisePol record2Pwp(const double* rec)
{
  mt::PieceWisePol pwp;
  const int n = (int)rec[0];
  if (n < 1 || n > NB_REC_PIECES) return pwp;
  for (int i = 0; i <= n; i++) pwp.times.push_back(rec[1 + i]);
  const double* co = rec + 1 + (NB_REC_PIECES + 1);
  for (int i = 0; i < n; i++)
  {
    pwp.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(co[4 * i], co[4 * i + 1], co[4 * i + 2], co[4 * i + 3]));
    const double* cy = co + NB_REC_PIECES * 4;
    pwp.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(cy[4 * i], cy[4 * i + 1], cy[4 * i + 2], cy[4 * i + 3]));
    const double* cz = co + 2 * NB_REC_PIECES * 4;
    pwp.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(cz[4 * i], cz[4 * i + 1], cz[4 * i + 2], cz[4 * i + 3]));
  }
  return pwp;
}

// the fields of mader_msgs/DynTraj besides pwp (and `function`, which agents leave empty, neptune_ros.cpp:438-443)
struct DynTrajHeader
{
  int id = 0;
  bool is_agent = true;
  double bbox[3] = { 0, 0, 0 };
  double pos[3] = { 0, 0, 0 };
  std::vector<Eigen::Vector2d> bendpt;  // tether: base first (publishOwnTraj :455-476)
  double seq = 0.0;                      // stands in for header.stamp / time_received
};

inline bool dynTraj2Record(const DynTrajHeader& h, const mt::PieceWisePol& pwp, double* rec)
{
  if (h.bendpt.size() > 8 || !pwp2Record(pwp, rec)) return false;
  for (int q = NB_REC_PWP_DOUBLES; q < NB_REC_DOUBLES; q++) rec[q] = 0.0;
  rec[NB_REC_OFF_ID] = (double)h.id, rec[NB_REC_OFF_ISAGENT] = h.is_agent ? 1.0 : 0.0, rec[NB_REC_OFF_SEQ] = h.seq;
  for (int k = 0; k < 3; k++) rec[NB_REC_OFF_BBOX + k] = h.bbox[k], rec[NB_REC_OFF_POS + k] = h.pos[k];
  rec[NB_REC_OFF_NBEND] = (double)h.bendpt.size();
  for (size_t q = 0; q < h.bendpt.size(); q++)
    rec[NB_REC_OFF_BEND + 2 * q] = h.bendpt[q](0), rec[NB_REC_OFF_BEND + 2 * q + 1] = h.bendpt[q](1);
  return true;
}

inline void record2DynTraj(const double* rec, DynTrajHeader& h, mt::PieceWisePol& pwp)
{
  pwp = record2Pwp(rec);
  h.id = (int)rec[NB_REC_OFF_ID], h.is_agent = rec[NB_REC_OFF_ISAGENT] != 0.0, h.seq = rec[NB_REC_OFF_SEQ];
  for (int k = 0; k < 3; k++) h.bbox[k] = rec[NB_REC_OFF_BBOX + k], h.pos[k] = rec[NB_REC_OFF_POS + k];
  h.bendpt.clear();
  for (int q = 0; q < (int)rec[NB_REC_OFF_NBEND]; q++)
    h.bendpt.push_back(Eigen::Vector2d(rec[NB_REC_OFF_BEND + 2 * q], rec[NB_REC_OFF_BEND + 2 * q + 1]));
}

// table [num_agents][NB_REC_DOUBLES], known [num_agents]: what Neptune::trajs_ holds (one entry per id) and which ids it
// holds.  Returns false for a record of an id outside 1..num_agents (trajCB drops those, neptune_ros.cpp:381-384).
inline bool updateTrajObstacles(double* table, uint8_t* known, int num_agents, const double* rec)
{
  const int id = (int)rec[NB_REC_OFF_ID];
  if (id < 1 || id > num_agents) return false;
  std::memcpy(table + (size_t)(id - 1) * NB_REC_DOUBLES, rec, sizeof(double) * NB_REC_DOUBLES);
  known[id - 1] = 1;
  return true;
}
}  // namespace nb
