// amie_b200_shim.h -- glue shared by the two drop-in translation units.
//
// Assembly::cgsolve constructs a fresh solver object per call (solvers/assembly.cpp:1841, :1914), so
// the device context cannot live in the solver: it is kept per Assembly in a registry and reused as
// long as the sparsity pattern is the same (structure is uploaded once per topology change, values
// on every solve -- they change each damage step).
#pragma once
#include <cstdint>
#include <valarray>
#include "../../../include/amie_b200.h"
#include "solvers/assembly.h"

namespace AmieB200Shim
{
// context of this assembly with the current matrix uploaded; nullptr + message on cerr if no device
amie_b200_ctx * context_for(Amie::Assembly * a) ;
void release(Amie::Assembly * a) ;
// what the caller passed as Preconditionner*: AMIE_B200_PRECOND_* or -1 (not available on the device)
int precond_kind(Amie::Preconditionner * p) ;
}
