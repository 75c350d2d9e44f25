// Headless stand-in for the viewer's run loop (/root/reference/eggshell/eggshell_view.cc:540-554):
// SimulationInitialization(), then N x SimulationStep(), then a state dump of the hanging chain
// (one line per body: index p.x p.y p.z v.x v.y v.z).  Usage: host_demo [steps]
#include <cstdio>
#include <cstdlib>

#include "eggshell/ensembles.h"
#include "eggshell/model.h"

const Ensemble& EggshellHangingChain();
const Ensemble& EggshellCairn();

int main(int argc, char** argv) {
  int steps = argc > 1 ? std::atoi(argv[1]) : 10;
  SimulationInitialization();
  for (int s = 0; s < steps; s++)
    if (!SimulationStep()) break;
  const Ensemble& ch = EggshellHangingChain();
  std::printf("chain %d\n", (int)ch.components().size());
  int i = 0;
  for (const auto& b : ch.components()) {
    std::printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", i++, b->p()(0), b->p()(1), b->p()(2), b->v()(0), b->v()(1), b->v()(2));
  }
  std::printf("chain_constraints %d status %d\n", (int)ch.constraints().size(), ch.status());
  const Ensemble& ca = EggshellCairn();
  std::printf("cairn %d\n", (int)ca.components().size());
  i = 0;
  for (const auto& b : ca.components())
    std::printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", i++, b->p()(0), b->p()(1), b->p()(2), b->v()(0), b->v()(1), b->v()(2));
  return 0;
}
