// tcgen05 GEMM host interface (see gemm_tc.cu).
#pragma once
#include "ops.h"

namespace rift {

// Pre-split bf16 planes of one fp32 weight matrix: hi = bf16(w), lo = bf16(w - hi), zero padded to the pitch.
// transpose == false: planes [N, Kp] of W [N, K] (forward operand);  true: planes [K, Np] of W^T (dX = dY W).
struct TcWeight {
    const float* src = nullptr; long long ld_src = 0;
    int N = 0, K = 0, Kp = 0;       // plane rows, valid columns, pitch
    void* hi = nullptr; void* lo = nullptr;
    bool trainable = false;
    bool transpose = false;
};

// A pair of bf16 planes [rows, pitch] and the origin of the operand inside them
// (mn0: offset along the M / N index of the GEMM, k0: offset along the reduction index)
struct PlaneOp { const void* hi; const void* lo; int rows; int pitch; int mn0; int k0; };

inline int tc_pitch(int K) { return (K + 63) / 64 * 64; }
// One product of a grouped weight-gradient launch: C[M, N] += A^T B over K rows (A = dY planes [K, >= M], B = X planes [K, >= N]),
// colsum[M] += column sums of A (bias gradient; may be null)
struct WgradItem { PlaneOp A, B; float* C; long long ldc; float* colsum; int M, N, K; };
bool wgrad_group_takes(int M, int N, int K, long long ldc, const void* C);
constexpr int WGRAD_GROUP_MAX = 12;
int launch_wgrad_group(const WgradItem* items, int n, int terms, cudaStream_t st);      // n <= WGRAD_GROUP_MAX
bool gemm_tc_shape_ok(int M, int N, int K);
bool gemm_tc_eligible(const GemmArgs& a);
// fp32 [M, K] -> bf16 planes hi / lo [M, Kp] (each M * Kp * 2 bytes, 256 B aligned)
int launch_pack_split(const float* src, long long ld, int M, int K, int Kp, void* hi, void* lo, cudaStream_t st);
// the same pass also produces per-slab column partials [slabs][K] of src (bias gradient, finished by launch_colsum_final);
// *slabs_out == 0 means the fused form does not apply (unaligned source) and nothing was launched
int pack_colsum_slabs(int M, int Kp);       // rows of the partial buffer launch_pack_split_colsum writes (each K floats)
int launch_pack_split_colsum(const float* src, long long ld, int M, int K, int Kp, void* hi, void* lo, float* partial,
                             int* slabs_out, cudaStream_t st);
// C = epilogue(A B^T) with A = planes (a_hi, a_lo) [M, Kp] and B = weight planes rows [n0, n0+N), cols [k0, k0+K)
int launch_gemm_tc(const GemmArgs& a, const void* a_hi, const void* a_lo, int Kp, const TcWeight& w, int n0, int k0,
                   cudaStream_t st);
// General form.  mn_major == false: A planes [M.., K..], B planes [N.., K..] (rows = M/N index).
// mn_major == true: A planes [K.., M..], B planes [K.., N..] (rows = reduction index; dW = dY^T X).
// splits > 1: split-K into `partials` [splits, M, N] followed by a fixed-order reduce that applies the epilogue.
int launch_gemm_tc_ex(const GemmArgs& a, const PlaneOp& A, const PlaneOp& B, bool mn_major, int splits, float* partials,
                      cudaStream_t st);

// profiling aid: CTA 0 of every following tcgen05 GEMM writes %globaltimer stamps into dev_buf (64 u64); null = off
void set_gemm_tc_trace(void* dev_buf);

// work units (32 x 64 plane tiles) of one weight; `first` / `total` of the job table count these
long long split_job_units(int N, int Kp);
size_t split_job_bytes();
void fill_split_job(void* dst, const float* src, long long ld, int N, int K, int Kp, void* hi, void* lo, long long first,
                    int transpose);
int launch_split_weights(const void* jobs_dev, int n_jobs, long long total, cudaStream_t st);

}  // namespace rift
