// pgb_pass_b.cu -- the second half of the instances of the O(k) pass kernel (pgb_pass.cuh): predict (no y) and the
// IRLS step of a prepared Gram pass; the first half is in pgb_pass.cu.
#include "pgb_pass.cuh"

template int pass_dispatch<false, PASS_STATS>(const PassLaunch&);
template int pass_dispatch<true, PASS_IRLS>(const PassLaunch&);
