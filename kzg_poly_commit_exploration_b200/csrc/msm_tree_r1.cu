// K2' round 1 with the field product OUT OF LINE (experiment, option tree_r1_ool).
//
// ncu's source page (profiles/prof_seg_add_lean_r02_instruction_mix.txt) attributes 12 % of the round-1 addition kernel's
// stall samples to instruction fetch (5 % in the dense rounds): the table gathers desynchronise the warps of an SM, and an
// addition is ~85 KB of straight-line code with the five products inlined.  This translation unit compiles the same kernel
// (msm_tree_kernels.cuh) with ONE copy of the product (KZG_FQ_NOINLINE: arguments in registers, as the bucket-reduction
// tail does), which shrinks the loop body to ~15 KB.
#define KZG_FQ_NOINLINE 1
#include "msm_tree_kernels.cuh"

namespace kzg {

void launch_seg_add_r1_ool(cudaStream_t stream, unsigned grid, const uint32_t* seg_in, const uint32_t* seg_out, uint32_t M,
                           const uint32_t* vals, const G1Affine* src, uint64_t T, int K, const Fq* prefix, Fq* tot,
                           G1Affine* out, uint32_t pitch) {
  k_seg_add_lean<true, 4, 1><<<grid, 128, 0, stream>>>(seg_in, seg_out, M, vals, src, T, K, prefix, tot, out, pitch);
}

}  // namespace kzg
