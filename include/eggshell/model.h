// Mirror of the drop-in boundary /root/reference/eggshell/model.h:8-40: the two entry points the
// viewer calls and the callbacks a step may call.  The Draw*/EggPlot symbols are weak no-ops here
// (the Qt/OpenGL viewer that defines them, eggshell_view.cc:375-422, is out of scope); a host
// application may define strong versions.
#ifndef EGGSHELL_MODEL_H_
#define EGGSHELL_MODEL_H_
#include "body.h"

void SimulationInitialization();
bool SimulationStep();

void DrawSphere(const Vector3d& center, const Matrix3d& rotation, double radius, int color = 0xffffff);
void DrawBox(const Vector3d& center, const Matrix3d& rotation, const Vector3d& side_lengths, int color = 0xffffff);
void DrawCapsule(const Vector3d& center, const Matrix3d& rotation, double radius, double length, int color = 0xffffff);
void DrawPoint(const Vector3d& position, int color = 0xffff00);
void DrawLine(const Vector3d& pos1, const Vector3d& pos2, int color = 0xffff00);
void EggPlot(const VectorXd& x, const MatrixXd& data, const char* title = "");

// A fatal error: print the message and stop execution (toolkit/error.cc:44-49: _exit(1)).
void Panic(const char* message, ...) __attribute__((noreturn));

void SimulationInitialization_HangingChain();
bool SimulationStep_HangingChain();
void SimulationInitialization_Cairn();
bool SimulationStep_Cairn();
#endif
