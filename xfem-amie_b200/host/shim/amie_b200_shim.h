// amie_b200_shim.h -- glue shared by the two drop-in translation units.
//
// Assembly::cgsolve constructs a fresh solver object per call (solvers/assembly.cpp:1841, :1914), so
// the device context cannot live in the solver: it is kept per Assembly in a registry and reused as
// long as the sparsity pattern is the same (structure is uploaded once per topology change, values
// on every solve -- they change each damage step).
#pragma once
#include <cstdint>
#include <valarray>
#include <vector>
#include "../../../include/amie_b200.h"
#include "solvers/assembly.h"

namespace AmieB200Shim
{
// context of this assembly with the current matrix uploaded; nullptr + message on cerr if no device
amie_b200_ctx * context_for(Amie::Assembly * a) ;
void release(Amie::Assembly * a) ;
// what the caller passed as Preconditionner*: AMIE_B200_PRECOND_* or -1 (not available on the device).
// For the reference's diagonal classes (InverseDiagonal, InverseDiagonalSquared, InverseLumpedDiagonal:
// precondition() is t = v .* diagonal, solvers/inversediagonal.cpp:44-82) *diagonal_out points at the OBJECT's own
// vector -- it may have been built from another matrix than the one being solved -- and the kind is
// AMIE_B200_PRECOND_DIAGONAL; pass it to upload_diagonal() once the context exists.
int precond_kind(Amie::Preconditionner * p, const Vector ** diagonal_out) ;
// Renumbering (env AMIE_B200_RENUMBER=1, assemblies without rowstart/colstart): the device works on the matrix
// renumbered by reverse Cuthill-McKee -- AMIE's mesher numbering has no locality -- and the solver shims permute b, x0
// and x at the boundary.  permutation_for: perm[old node] = new node of the context's current structure, or nullptr.
const std::vector<uint32_t> * permutation_for(Amie::Assembly * a) ;
// out = in in the device numbering (a shorter `in` is a prefix in AMIE's numbering, zero-filled) / back
void to_device_order(const std::vector<uint32_t> & perm, size_t stride, const Vector & in, Vector & out) ;
void from_device_order(const std::vector<uint32_t> & perm, size_t stride, const Vector & in, Vector & out) ;
// false + message on cerr if the vector does not have one entry per degree of freedom or the upload fails
bool upload_diagonal(amie_b200_ctx * ctx, const Vector * diagonal, size_t ndof, Amie::Assembly * a = nullptr) ;
}
