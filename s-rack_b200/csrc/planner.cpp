// Cycle-breaking topological planner with the reference's exact ordering and
// cut semantics (src/synth.rs:107-212), on dense indices instead of
// Arc<RwLock<..>> keyed hash maps.  Deterministic for a given module-list order.
#include <algorithm>

#include "patch.hpp"

namespace srk {
namespace {

// is_loop (src/synth.rs:107-126): breadth-first walk over the remaining source
// edges starting at `module`; returns the first node met whose sources contain
// `module` itself, i.e. the reader of a wire that closes a cycle through `module`.
// The reference rescans its growing `to_search` list for the first unvisited
// entry; since entries are only ever appended that is a queue with a moving head.
int find_cycle_reader(int module, const std::vector<std::vector<int>>& edges, std::vector<int>& queue,
                      std::vector<uint8_t>& seen) {
  queue.clear();
  std::fill(seen.begin(), seen.end(), 0);
  queue.push_back(module);
  for (size_t head = 0; head < queue.size(); ++head) {
    int cur = queue[head];
    if (seen[cur]) continue;
    seen[cur] = 1;
    for (int dep : edges[cur]) {
      if (dep == module) return cur;
      queue.push_back(dep);
    }
  }
  return -1;
}

}  // namespace

void plan_execution(int output, const std::vector<std::vector<int>>& deps, std::vector<int>& plan,
                    std::vector<std::pair<int, int>>& cuts) {
  const int n = (int)deps.size();
  plan.clear();
  cuts.clear();
  // Phase 1 (synth.rs:134-163): sink -> sources.  Every module is in the list, so
  // the reachability walk of the reference visits exactly the list.
  std::vector<std::vector<int>> edges = deps;

  // Phase 2 (synth.rs:164-192): depth-first from the back of [all_modules..., output];
  // the first module of a cycle that is reached loses its outgoing wire into the cycle.
  std::vector<int> stack(n);
  for (int i = 0; i < n; ++i) stack[i] = i;
  stack.push_back(output);
  std::vector<uint8_t> visited(n, 0), seen(n, 0);
  std::vector<int> queue;
  while (!stack.empty()) {
    int m = stack.back();
    stack.pop_back();
    if (visited[m]) continue;
    visited[m] = 1;
    for (int dep : edges[m]) stack.push_back(dep);
    for (int reader; (reader = find_cycle_reader(m, edges, queue, seen)) >= 0;) {
      auto& src = edges[reader];
      src.erase(std::remove(src.begin(), src.end(), m), src.end());
      cuts.emplace_back(reader, m);
    }
  }

  // Phase 3 (synth.rs:193-211): repeatedly the first not-yet-planned module, in list
  // order, all of whose remaining sources are planned.
  std::fill(visited.begin(), visited.end(), 0);
  for (;;) {
    int next = -1;
    for (int m = 0; m < n && next < 0; ++m) {
      if (visited[m]) continue;
      bool ready = true;
      for (int d : edges[m])
        if (!visited[d]) { ready = false; break; }
      if (ready) next = m;
    }
    if (next < 0) break;
    visited[next] = 1;
    plan.push_back(next);
  }
}

}  // namespace srk
