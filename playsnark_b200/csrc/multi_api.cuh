// Internal (non-ABI) entry points the multi-GPU orchestration (capi_multi.cu) uses on each device's
// context, next to the public per-device calls of include/playsnark_b200.h.
#pragma once
#include "poly_api.cuh"

#include <functional>

namespace ps {

// Index ranges of the proving key one device holds: points [x_lo, x_hi) of Xi and Xi2, [t_lo, t_hi) of XiT,
// [n_lo, n_hi) of NioLP; `consts` marks the device that also holds the single points
// (Delta, Alpha | Delta2, Beta2 | Alpha, Beta, Delta).
struct KeySlice { size_t x_lo, x_hi, t_lo, t_hi, n_lo, n_hi; bool consts; };

// the per-device part of a sharded Groth16 key: base sets A_d = [Xi_d | Delta Alpha], B_d = [Xi2_d | Delta2 Beta2],
// C_d = [NioLP_d | XiT_d | Xi_d | Alpha Beta Delta] (single points on the `consts` device only); one window for the G1
// pair (they share a pipeline), one for B_d
int g16_key_load_slice(ps_ctx* ctx, const KeySlice& sl, int format, int window_g1, int window_g2, const uint8_t* xi, const uint8_t* xi2,
                       const uint8_t* xit, const uint8_t* niolp, const uint8_t* alpha, const uint8_t* beta, const uint8_t* delta,
                       const uint8_t* beta2, const uint8_t* delta2, ps_g16_key** key);

// scalar vectors of that device (standard form, device memory):
//   scA = [a[x_lo..x_hi) | r 1], scB = [b[x_lo..x_hi) | s 1],
//   scC = [w_nio[n_lo..n_hi) | h[t_lo..t_hi) (left untouched here) | (s a + r b)[x_lo..x_hi) | s r rs]
// a, b: n Montgomery coefficients; w: the whole witness (Montgomery), w_nio = w[diff..)
int g16_slice_scalars(ps_ctx* ctx, const KeySlice& sl, const uint8_t* r_be, const uint8_t* s_be, const Fr* d_a, const Fr* d_b,
                      const Fr* d_w, size_t diff, Fr* scA, Fr* scB, Fr* scC);

// The MSMs of one device of a sharded proof: B_d (G2) goes out first, on the second (high-priority) stream -- it needs
// neither h nor the witness tail; then `before_g1` runs on the host (the pipelined prover makes the primary stream wait
// for device 0's h there), then A_d and the WHOLE of C_d as one G1 pipeline with two outputs.  B_d's accumulation fills
// the time device 0 spends dividing, its latency-bound tail hides under the G1 accumulation, and every device pays the
// fixed sort / merge / reduction cost of a G1 pipeline once per proof instead of twice (early + late).
//   d_partials: [A 192 B | C 192 B | B 384 B | 192 B zero]   (record layout of ps_g16_combine, C's late slot = infinity)
int g16_slice_msm_all(ps_ctx* ctx, const ps_g16_key* key, const KeySlice& sl, const Fr* scA, const Fr* scB, const Fr* scC,
                      void* d_partials, const std::function<int()>& before_g1);

// Poly.BlindEval over the whole of `b` with host scalars (wire format), leaving the XYZZ partial on the device
// (one device's share of ps_mmsm); *d_err_out points at the device flag for scalars >= r
int msm_partial_host_scalars(ps_ctx* ctx, const ps_bases* b, const uint8_t* scalars_be, size_t n, void* d_out_xyzz, uint32_t** d_err_out);

}  // namespace ps
