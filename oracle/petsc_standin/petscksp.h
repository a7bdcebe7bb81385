#include "petsc_standin.h"
