// Iteration cadence and tolerances of the host trainers.
//
// The names are the ones code written against the reference expects (they are macros there, const.h:4-12) and
// the values must equal the reference's: they decide when the objective is evaluated, when progress is printed,
// when factors are saved and when two consecutive objectives count as converged — all of which show in the
// stdout lines and output files the parity tests compare.  Here they are typed constants; the GKlib read flag is the
// only one that stays a macro because GKlib.h uses it in a default argument.
#pragma once

namespace matfac_cadence {
constexpr int evaluate_every = 1;     // epochs between objective / validation evaluations (model.cpp:1471 call sites)
constexpr int display_every = 50;     // epochs between progress lines
constexpr int save_every = 50;        // epochs between factor dumps
constexpr int chance_epochs = 500;    // patience of the long-run stopping rule (model.cpp:1506-1519)
constexpr double objective_eps = 1e-5;  // |delta objective| below this stops training (model.cpp:1521)
}  // namespace matfac_cadence

constexpr int OBJ_ITER = matfac_cadence::evaluate_every;
constexpr int DISP_ITER = matfac_cadence::display_every;
constexpr int SAVE_ITER = matfac_cadence::save_every;
constexpr int CHANCE_ITER = matfac_cadence::chance_epochs;
constexpr double EPS = matfac_cadence::objective_eps;

#define GK_CSR_IS_VAL 1
