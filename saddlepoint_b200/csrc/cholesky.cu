// cholesky.cu -- supernodal LL^T (placeholder until the numeric phase lands).
#include "spb_internal.h"
namespace spb {
spb_status chol_factorize(spb_context* h) { throw StateFail{"Cholesky solver not available yet in this build"}; }
spb_status chol_solve(spb_context* h, const double*, double*) { throw StateFail{"Cholesky solver not available yet in this build"}; }
void chol_release(spb_context* h) { h->chol = nullptr; }
}  // namespace spb
