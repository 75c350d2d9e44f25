// Mirror of /root/reference/eggshell/model.cc:28-115: the file-static scenes and the two entry
// points the viewer calls, plus weak no-op Draw* callbacks and Panic().
#include "eggshell/model.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <unistd.h>

#include "eggshell/ensembles.h"

namespace {
constexpr double kSimTimeStep = 0.001;   // constants.h:6
Chain& ch1() { static Chain c(10, Vector3d(2, 2, 1)); return c; }                        // model.cc:28
Cairn& cairn() { static Cairn c(4, {-0.2, 0.2}, {-0.2, 0.2}, {1, 8}); return c; }        // model.cc:31
}  // namespace

void Panic(const char* message, ...) {
  va_list ap;
  va_start(ap, message);
  std::fprintf(stderr, "Panic: ");
  std::vfprintf(stderr, message, ap);
  std::fprintf(stderr, "\n");
  va_end(ap);
  std::fflush(stderr);
  _exit(1);                                                                              // toolkit/error.cc:44-49
}

__attribute__((weak)) void DrawSphere(const Vector3d&, const Matrix3d&, double, int) {}
__attribute__((weak)) void DrawBox(const Vector3d&, const Matrix3d&, const Vector3d&, int) {}
__attribute__((weak)) void DrawCapsule(const Vector3d&, const Matrix3d&, double, double, int) {}
__attribute__((weak)) void DrawPoint(const Vector3d&, int) {}
__attribute__((weak)) void DrawLine(const Vector3d&, const Vector3d&, int) {}
__attribute__((weak)) void EggPlot(const VectorXd&, const MatrixXd&, const char*) {}

void Body::Draw() const { DrawBox(p(), R(), side_lengths_); }

void SimulationInitialization() {                       // model.cc:33-36
  SimulationInitialization_HangingChain();
  SimulationInitialization_Cairn();
}
bool SimulationStep() {                                 // model.cc:38-71
  SimulationStep_HangingChain();
  SimulationStep_Cairn();
  return true;
}
void SimulationInitialization_Cairn() {                 // model.cc:73-76
  cairn().Init();
  cairn().InitStabilize();
}
bool SimulationStep_Cairn() {                           // model.cc:78-95
  cairn().Draw();
  cairn().Step(kSimTimeStep * 5, Ensemble::Integrator::OPEN_DYNAMICS_ENGINE);
  return true;
}
void SimulationInitialization_HangingChain() { ch1().Init(); }   // model.cc:97-100
bool SimulationStep_HangingChain() {                    // model.cc:102-115
  ch1().Draw();
  ch1().Step(kSimTimeStep, Ensemble::Integrator::OPEN_DYNAMICS_ENGINE);
  return true;
}

// Accessors for the headless driver (host_demo.cc).
const Ensemble& EggshellHangingChain() { return ch1(); }
const Ensemble& EggshellCairn() { return cairn(); }
