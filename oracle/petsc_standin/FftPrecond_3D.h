/* tests/FFTDirectSolver/testFftSolver_3D.c includes the solver header under this earlier name */
#include "FftLinearSolver_3D.h"
