#include "solverlab_standin.hxx"
