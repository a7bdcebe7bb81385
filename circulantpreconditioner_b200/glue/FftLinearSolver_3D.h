/* forwards to the B200 implementation of the reference interface (see circulantpc_petsc.h) */
#include "circulantpc_petsc.h"
