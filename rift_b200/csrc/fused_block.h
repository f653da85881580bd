// Fused transformer sub-block kernels (tcgen05 + thread-block clusters), host interface.  See fused_block.cu.
#pragma once
#include "gemm_tc.h"

namespace rift {

// Pre-LN MLP sub-block  Y = X + fc2(act(fc1(LN(X))))  (layers/transformer.py:83-94, modules/planning_decoder.py:80-86,
// layers/embedding.py NATLayer MLP) in ONE kernel.  Optional outputs are what the backward needs (all may be null).
struct FusedMlpArgs {
    const float* X = nullptr; long long ldx = 0;     // [rows, D] fp32 residual stream
    float* Y = nullptr; long long ldy = 0;           // [rows, D] fp32
    int rows = 0, D = 0, Hd = 0, act = 0;
    const float* ln_g = nullptr; const float* ln_b = nullptr;
    const float* b1 = nullptr; const float* b2 = nullptr;
    float* ln_mean = nullptr; float* ln_rstd = nullptr;      // [rows]
    Planes t2p;                                      // LN(X) as split-bf16 planes [rows, pitch(D)]
    float* hpre = nullptr;                           // fc1 output before the activation, fp32 [rows, Hd]
    Planes hmp;                                      // act(fc1) as planes [rows, pitch(Hd)]
};
// profiling aid: CTA 0 of every following fused launch writes %globaltimer stamps into dev_buf (16 u64); null = off
void set_fused_trace(void* dev_buf);
bool fused_mlp_shape_ok(int rows, int D, int Hd);
bool fused_blocks_enabled();      // RIFT_B200_FUSED=1
int launch_fused_mlp(const FusedMlpArgs& a, const TcWeight& w1, const TcWeight& w2, cudaStream_t st);

}  // namespace rift
