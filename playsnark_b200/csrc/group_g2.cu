// G2 (over Fp2) instantiation of the per-group device operations; its bucket-accumulation kernel lives in
// accum_g2.cu (fully inlined base-field products, the longest single compile of the library).
#include "group_impl.cuh"

namespace ps {
template struct GroupOps<Fp2>;
}
