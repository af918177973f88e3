// heap_check.cpp -- runs a script of open-list operations through the REAL std::priority_queue of this
// image's libstdc++, with the comparator of KinodynamicSearch (kinodynamic_search.hpp:163-179) on pointers to
// nodes whose (g, h) may be overwritten while they sit in the queue (kinodynamic_search.cpp:1193-1205).
// tests/test_search.py compares the pop order with the oracle's heap restatement (orc_heap_replay).
// stdin: n_ops n_ids bias, then n_ops lines "kind id g h"; stdout: popped ids.
#include <cmath>
#include <cstdio>
#include <queue>
#include <vector>

struct Node
{
  double g = 0, h = 0;
  int id = 0;
};

struct CompareCost
{
  double bias;
  bool operator()(const Node* left, const Node* right) const
  {
    double cost_left = left->g + bias * left->h;
    double cost_right = right->g + bias * right->h;
    if (fabs(cost_left - cost_right) < 1e-5)
      return left->h > right->h;
    else
      return cost_left > cost_right;
  }
};

int main()
{
  int n_ops, n_ids;
  double bias;
  if (scanf("%d %d %lf", &n_ops, &n_ids, &bias) != 3) return 2;
  std::vector<Node> nodes(n_ids);
  for (int i = 0; i < n_ids; i++) nodes[i].id = i;
  CompareCost cmp;
  cmp.bias = bias;
  std::priority_queue<Node*, std::vector<Node*>, CompareCost> q(cmp);
  for (int k = 0; k < n_ops; k++)
  {
    int kind, id;
    double g, h;
    if (scanf("%d %d %lf %lf", &kind, &id, &g, &h) != 4) return 2;
    if (kind == 0)
    {
      nodes[id].g = g, nodes[id].h = h;
      q.push(&nodes[id]);
    }
    else if (kind == 1)
    {
      if (!q.empty())
      {
        printf("%d\n", q.top()->id);
        q.pop();
      }
    }
    else
      nodes[id].g = g, nodes[id].h = h;
  }
  return 0;
}
