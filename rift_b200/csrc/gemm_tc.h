// tcgen05 GEMM host interface (see gemm_tc.cu).
#pragma once
#include "ops.h"

namespace rift {

// Pre-split bf16 planes of one fp32 weight matrix [N, K]: hi = bf16(w), lo = bf16(w - hi), row pitch Kp
// (K rounded up to 64, zero padded), plus the two TMA descriptors (box 64 x 64, SWIZZLE_128B).
struct TcWeight {
    const float* src = nullptr; long long ld_src = 0;
    int N = 0, K = 0, Kp = 0;
    void* hi = nullptr; void* lo = nullptr;
    bool trainable = false;
    alignas(64) unsigned char tm_hi[128];
    alignas(64) unsigned char tm_lo[128];
};

inline int tc_pitch(int K) { return (K + 63) / 64 * 64; }
bool gemm_tc_shape_ok(int M, int N, int K);
bool gemm_tc_eligible(const GemmArgs& a);
// fp32 [M, K] -> bf16 planes hi / lo [M, Kp] (each M * Kp * 2 bytes, 256 B aligned)
int launch_pack_split(const float* src, long long ld, int M, int K, int Kp, void* hi, void* lo, cudaStream_t st);
// a.A / a.B are ignored: the A operand is the plane pair (a_hi, a_lo), the B operand is `w`
// rows [n0, n0 + a.N), columns [k0, k0 + a.K)
int launch_gemm_tc(const GemmArgs& a, const void* a_hi, const void* a_lo, int Kp, const TcWeight& w, int n0, int k0,
                   cudaStream_t st);

int make_weight_tensor_map(void* map_out, const void* plane, int N, int Kp);
size_t split_job_bytes();
void fill_split_job(void* dst, const float* src, long long ld, int N, int K, int Kp, void* hi, void* lo, long long first);
int launch_split_weights(const void* jobs_dev, int n_jobs, long long total, cudaStream_t st);

}  // namespace rift
