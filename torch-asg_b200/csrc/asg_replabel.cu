// asg_replabel.cu -- target preprocessing on the device: replabel packing / unpacking.
//
// ASG transcripts cannot hold the same label twice in a row (consecutive equal frames collapse: doc/tech_report.tex:126), so a
// run "jj" is written "j r" with extra symbols r_k = "repeat the previous label".  The reference leaves this to the user
// (README.md:93-94 lists what it omits; its tests feed raw random targets).  The convention here is wav2letter++'s
// (flashlight lib/text `packReplabels` / `unpackReplabels`, not part of /root/reference): with max_reps = m the token
// first_rep_label + k - 1 (1 <= k <= m) means "the previous label k MORE times", runs longer than m + 1 start over:
//     a a a a  (m = 2)  ->  a r2 a          i j j k l l l  (m >= 2)  ->  i j r1 k l r2   (the report's  i j r_2 k l r_3)
// Integer-only, a few hundred bytes per utterance: one warp per utterance, the row staged through shared memory with coalesced
// loads and stores, the greedy pass itself by one lane (it is inherently sequential in the cap rule and costs microseconds).
#include "asg_common.cuh"
#include "asg_internal.h"

namespace asg {

__global__ void __launch_bounds__(32) pack_replabels_kernel(const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int S,
                                                            int64_t first_rep, int max_reps, int64_t fill, int64_t *out, int64_t *out_len) {
    extern __shared__ int64_t rp_smem[];                 // [S] in, [S] out
    int64_t *in = rp_smem, *res = rp_smem + S;
    const int n = blockIdx.x, lane = threadIdx.x;
    int Sn = tl ? (int)tl[n] : S;
    Sn = max(0, min(Sn, S));
    const int64_t *yn = y + (int64_t)n * y_s0;
    for (int s = lane; s < Sn; s += 32) in[s] = yn[(int64_t)s * y_s1];
    __syncwarp();
    int m = 0;
    if (lane == 0) {
        int64_t prev = 0;
        int reps = 0;
        bool have = false;
        for (int s = 0; s < Sn; ++s) {
            const int64_t tok = in[s];
            if (have && tok == prev && reps < max_reps) {
                ++reps;
            } else {
                if (reps > 0) { res[m++] = first_rep + reps - 1; reps = 0; }
                res[m++] = tok;
                prev = tok;
                have = true;
            }
        }
        if (reps > 0) res[m++] = first_rep + reps - 1;
        out_len[n] = m;
    }
    m = __shfl_sync(ASG_FULL_MASK, m, 0);
    __syncwarp();
    int64_t *on = out + (int64_t)n * S;
    for (int s = lane; s < S; s += 32) on[s] = s < m ? res[s] : fill;
}

__global__ void __launch_bounds__(32) unpack_replabels_kernel(const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int S,
                                                              int64_t first_rep, int max_reps, int64_t fill, int64_t *out, int S_out,
                                                              int64_t *out_len) {
    extern __shared__ int64_t rp_smem[];                 // [S] in, [S_out] out
    int64_t *in = rp_smem, *res = rp_smem + S;
    const int n = blockIdx.x, lane = threadIdx.x;
    int Sn = tl ? (int)tl[n] : S;
    Sn = max(0, min(Sn, S));
    const int64_t *yn = y + (int64_t)n * y_s0;
    for (int s = lane; s < Sn; s += 32) in[s] = yn[(int64_t)s * y_s1];
    __syncwarp();
    long long m = 0;
    if (lane == 0) {
        int64_t prev = 0;
        bool have = false;
        for (int s = 0; s < Sn; ++s) {
            const int64_t tok = in[s];
            const int64_t k = tok - first_rep + 1;                  // 1..max_reps for a replabel
            if (k >= 1 && k <= max_reps) {
                if (have) {                                         // a leading replabel has nothing to repeat: dropped, like flashlight
                    for (int64_t i = 0; i < k; ++i) { if (m < S_out) res[m] = prev; ++m; }
                }
            } else {
                if (m < S_out) res[m] = tok;
                ++m;
                prev = tok;
                have = true;
            }
        }
        out_len[n] = m;                                             // the full length, even when it did not fit S_out
    }
    m = __shfl_sync(ASG_FULL_MASK, m, 0);
    __syncwarp();
    int64_t *on = out + (int64_t)n * S_out;
    for (int s = lane; s < S_out; s += 32) on[s] = s < m ? res[s] : fill;
}

int replabel_dispatch(bool pack, const int64_t *y, int64_t y_s0, int64_t y_s1, const int64_t *tl, int64_t N, int64_t S, int64_t first_rep,
                      int64_t max_reps, int64_t fill, int64_t *out, int64_t S_out, int64_t *out_len, cudaStream_t st) {
    const size_t smem = sizeof(int64_t) * (size_t)(S + (pack ? S : S_out));
    cudaError_t e;
    if (pack) {
        if ((e = cudaFuncSetAttribute(pack_replabels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        LaunchScope ls(K_HOST_MISC, st);
        pack_replabels_kernel<<<(unsigned)N, 32, smem, st>>>(y, y_s0, y_s1, tl, (int)S, first_rep, (int)max_reps, fill, out, out_len);
    } else {
        if ((e = cudaFuncSetAttribute(unpack_replabels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        LaunchScope ls(K_HOST_MISC, st);
        unpack_replabels_kernel<<<(unsigned)N, 32, smem, st>>>(y, y_s0, y_s1, tl, (int)S, first_rep, (int)max_reps, fill, out, (int)S_out, out_len);
    }
    return (int)cudaGetLastError();
}

}  // namespace asg
