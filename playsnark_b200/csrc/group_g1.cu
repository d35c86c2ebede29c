// G1 (over Fp) instantiation of the per-group device operations: Pippenger pipeline, codec, setup kernels.
#include "group_impl.cuh"

namespace ps {
template struct GroupOps<Fp>;
}
