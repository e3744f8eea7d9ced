// petsc/private/pcimpl.h of the stand-in (tests/petsc_stub/petsc.h): the private PC layout (_p_PC, _PCOps) lives in petsc.h.
#ifndef PETSC_STUB_PCIMPL_H
#define PETSC_STUB_PCIMPL_H
#include <petsc.h>
#endif
