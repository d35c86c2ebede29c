// G2 bucket accumulation kernel (MsmAccumK over Fp2I: base-field products inlined) in its own
// translation unit so that it compiles in parallel with capi.cu.
#include "msm.cuh"

namespace ps {

int launch_accum_g2(ps_stream_t st, size_t T1, uint32_t nb, uint32_t L, const Affine<Fp2>* tab, const uint32_t* ent, const uint32_t* off,
                    XYZZ<Fp2>* buckets, XYZZ<Fp2>* slot_pt, int32_t* slot_bid, uint8_t* slot_fl) {
  static_assert(sizeof(Fp2I) == sizeof(Fp2) && sizeof(XYZZ<Fp2I>) == sizeof(XYZZ<Fp2>) && sizeof(Affine<Fp2I>) == sizeof(Affine<Fp2>),
                "Fp2 and Fp2I must share their layout");
  PS_LAUNCH(MsmAccumK<Fp2I>, st, T1, nb, L, reinterpret_cast<const Affine<Fp2I>*>(tab), ent, off, reinterpret_cast<XYZZ<Fp2I>*>(buckets),
            reinterpret_cast<XYZZ<Fp2I>*>(slot_pt), slot_bid, slot_fl);
  return PS_OK;
}

}  // namespace ps
